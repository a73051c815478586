#!/bin/bash
# Selected GPU tests + bench in one call.  usage: gpurun -- 'bash scripts/gpu_tests_sel.sh <tag> "<pytest -k expr or files>" [bench]'
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest $2 -m gpu -q --maxfail=20 --timeout 900 --durations=8 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -25 $OUT/pytest.log
if [ "$3" == "bench" ]; then
  timeout 900 python bench.py --variants > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
  cat $OUT/bench.json; tail -60 $OUT/bench.err
  cp gpurun_out/variants.json $OUT/variants.json 2>/dev/null
fi
