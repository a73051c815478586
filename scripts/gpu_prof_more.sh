#!/bin/bash
# ncu captures of the non-headline kernels changed in round 2.  usage: gpurun -- 'bash scripts/gpu_prof_more.sh <tag>'
TAG=$1
OUT=gpurun_out/$TAG
mkdir -p $OUT
prof() {  # name regex driver-args...
  name=$1; re=$2; shift 2
  timeout 400 ncu --set full --clock-control none -k "regex:$re" -s 2 -c 1 -f -o $OUT/$name python scripts/prof_driver.py "$@" > $OUT/ncu_$name.log 2>&1
  if [ -f $OUT/$name.ncu-rep ]; then
    python scripts/ncu_summary.py $OUT/$name.ncu-rep > $OUT/$name.summary.txt 2>&1
    rm -f $OUT/$name.ncu-rep
    echo "== $name"; grep "Kernel\|kernel:\|time_duration\|inst_executed.sum\|issue_active\|long_score\|short_score\|hit_rate\|dram__bytes\|warps_active\|registers" $OUT/$name.summary.txt
  else
    tail -3 $OUT/ncu_$name.log
  fi
}
prof fwd_enc_K4_trained "box_fwd_win" --workload enc --K 4 --dist trained
prof bwd_enc_K4_trained "box_bwd_win" --workload enc --K 4 --dist trained
prof fwd_enc_K4_bf16 "box_fwd_win" --workload enc --K 4 --dtype bf16
prof bwd_enc_K4_bf16 "box_bwd_win" --workload enc --K 4 --dtype bf16
prof inst_fwd_K14_f32 "inst_fwd" --workload mask --K 14
prof inst_bwd_K14_f32 "inst_bwd" --workload mask --K 14
prof fwd_enc_K2 "box_fwd_win" --workload enc --K 2
