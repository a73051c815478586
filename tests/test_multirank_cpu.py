"""CPU, world_size 2 (gloo): the N>1 plumbing of bench.py.

The op has no collective (images shard over ranks), so what there is to check on
the host side is: every rank draws different images, the reported time is the max over ranks, and
under torchrun the CPU reference arm runs on rank 0 only while the other ranks exit 0 silently.
"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    ms = bench.max_over_ranks(10.0 + 5.0 * rank, "cpu")
    seeds = [None] * world
    dist.all_gather_object(seeds, bench.rank_seed(rank))
    if rank == 0:
        with open(out, "w") as f:
            json.dump({"ms": ms, "seeds": seeds}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_max_over_ranks_and_per_rank_images(tmp_path):
    out = str(tmp_path / "r.json")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = json.load(open(out))
    assert res["ms"] == 15.0                      # the slower rank decides
    assert len(set(res["seeds"])) == 2            # different images per rank


@pytest.mark.timeout(600)
def test_reference_arm_under_torchrun_prints_one_line():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"]
    env = dict(os.environ, OMP_NUM_THREADS=str(max(1, (os.cpu_count() or 2) // 2)))
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=580)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, proc.stdout
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["n_gpus"] == 2 and rec["unit"] == "Gsamples/s"
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["cpu_baseline"]["kind"] == "port"
    assert rec["value"] > 0
