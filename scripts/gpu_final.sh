#!/bin/bash
# last call of a round: full GPU suite + smoke + bench of HEAD, then initcheck / synccheck.  usage: gpurun -- 'bash scripts/gpu_final.sh <tag>'
TAG=${1:-fin}
bash scripts/gpu_tests.sh $TAG bench
OUT=gpurun_out/${TAG}_san
mkdir -p $OUT
SEL2='test_window_kernels_vs_oracle and (enc_box_K4 or enc_box_K2 or enc_uniform_K4) and f32'
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k "$SEL2" -p no:cacheprovider > $OUT/initcheck.log 2>&1; echo "initcheck rc=$?" | tee -a $OUT/initcheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k "$SEL2" -p no:cacheprovider > $OUT/synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $OUT/synccheck.log
grep -E "ERROR SUMMARY|passed|failed" $OUT/initcheck.log $OUT/synccheck.log
