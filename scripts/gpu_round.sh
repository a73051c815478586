#!/bin/bash
# One gpurun call: GPU tests, bench (+variants), ncu launch list, ncu full captures of the two hot kernels.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag> [quick]'
TAG=${1:-r01}
MODE=${2:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $OUT/gpu.csv 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 1300 python -m pytest tests -m gpu -q --maxfail=12 --timeout 900 --durations=10 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
timeout 600 python bench.py --variants > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json; tail -30 $OUT/bench.err
cp gpurun_out/variants.json $OUT/variants.json 2>/dev/null
[ -x scripts/microbench/red_throughput ] && scripts/microbench/red_throughput | tee $OUT/red_throughput.json
python scripts/host_overhead.py $OUT/host_overhead.json 2>&1 | tail -6
timeout 300 python scripts/compare_reference_cuda.py $OUT/vs_reference_cuda.json > $OUT/vs_reference_cuda.log 2>&1; tail -8 $OUT/vs_reference_cuda.log
if [ "$MODE" != "quick" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:attn_|box_|absmax|finalize|det_scale" -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd_vec|box_fwd_win" -s 2 -c 1 -f -o $OUT/fwd_enc_K4 \
    python scripts/prof_driver.py --workload enc --K 4 > $OUT/ncu_fwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:attn_bwd_vec|box_bwd_win" -s 2 -c 1 -f -o $OUT/bwd_enc_K4 \
    python scripts/prof_driver.py --workload enc --K 4 > $OUT/ncu_bwd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:attn_fwd_vec|box_fwd_win" -s 2 -c 1 -f -o $OUT/fwd_enc_K4_uniform \
    python scripts/prof_driver.py --workload enc --K 4 --dist uniform > $OUT/ncu_fwd_u.log 2>&1
# the .ncu-rep files have grown past the 64 MiB copy-back limit: digest them here, keep only text
for r in fwd_enc_K4 bwd_enc_K4 fwd_enc_K4_uniform; do
  if [ -f $OUT/$r.ncu-rep ]; then
    python scripts/ncu_summary.py $OUT/$r.ncu-rep > $OUT/$r.summary.txt 2>&1
    ncu -i $OUT/$r.ncu-rep --page source --csv --print-source sass > $OUT/$r.source_sass.csv 2>/dev/null
    ncu -i $OUT/$r.ncu-rep --page details > $OUT/$r.details.txt 2>/dev/null
    rm -f $OUT/$r.ncu-rep
  fi
done
fi
ls -la $OUT
