#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
    'sm__cycles_elapsed.max', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_red.sum',
    'lts__t_sectors_op_atom.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_global_red.sum',
    'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_selected_per_issue_active.ratio',
]


def main(path, grep=None):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('# kernel:', name[:100])
        for i, h in enumerate(hdr):
            if h in WANT or (grep and grep in h):
                print(f'{h:80s} {vals[i]:>18s} {units[i]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
