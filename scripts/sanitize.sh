#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (memcheck + racecheck); logs under gpurun_out/<tag>/
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SEL='test_window_kernels_vs_oracle or test_instance_mask_head_vs_oracle or test_box_decoder_like_vs_oracle or test_misaligned or test_all_out_of_range or test_empty or test_six_d or test_nonfinite'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k "$SEL" -p no:cacheprovider > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/memcheck.log
# the fused entry points (grid / softmax / instance weights / value epilogue), fp32 and bf16 cases
SELF='(test_fused_matches_reference_formulation or test_fused_softmax_matches_reference_formulation or test_instance_weights or test_value_epilogue) and not f64'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fused.py -q -x -k "$SELF" -p no:cacheprovider > $OUT/memcheck_fused.log 2>&1; echo "memcheck(fused) rc=$?" | tee -a $OUT/memcheck_fused.log
tail -4 $OUT/memcheck_fused.log
tail -4 $OUT/memcheck.log
SEL2='test_window_kernels_vs_oracle and (enc_box_K4 or enc_box_K2 or enc_uniform_K4) and f32'
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k "$SEL2" -p no:cacheprovider > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/racecheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k 'test_instance_mask_head_vs_oracle and f32' -p no:cacheprovider > $OUT/racecheck_inst.log 2>&1; echo "racecheck(instance) rc=$?" | tee -a $OUT/racecheck_inst.log
grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck_inst.log | sort | uniq -c | head
grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck.log | sort | uniq -c | head -20
grep -E "hazard" $OUT/racecheck.log | sed -E 's/0x[0-9a-f]+/X/g; s/[0-9]+ bytes/N bytes/' | sort | uniq -c | sort -rn | head -12
# optional third pass: uninitialised global reads and barrier misuse on the window / tile kernels
if [ "$2" == "more" ]; then
  timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k "$SEL2" -p no:cacheprovider > $OUT/initcheck.log 2>&1; echo "initcheck rc=$?" | tee -a $OUT/initcheck.log
  timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -x -k "$SEL2" -p no:cacheprovider > $OUT/synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a $OUT/synccheck.log
  grep -E "ERROR SUMMARY|passed|failed" $OUT/initcheck.log $OUT/synccheck.log
fi
