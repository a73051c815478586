#!/bin/bash
# GPU tests (+ optional bench) in one gpurun call.  usage: gpurun --timeout 1500 -- 'bash scripts/gpu_tests.sh <tag> [bench]'
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.csv 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 --durations=15 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -40 $OUT/pytest.log
if [ "$2" == "bench" ]; then
  timeout 600 python bench.py --variants > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
  cat $OUT/bench.json; tail -40 $OUT/bench.err
  cp gpurun_out/variants.json $OUT/variants.json 2>/dev/null
fi
