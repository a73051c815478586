#!/bin/bash
OUT=gpurun_out/r01u
mkdir -p $OUT
BOXER_B200_LIB=boxer_b200/_C/ab/lib_noflags.so timeout 240 compute-sanitizer --tool memcheck --print-limit 5 python scripts/prof_driver.py --workload enc --K 4 --iters 1 > $OUT/noflags_memcheck.log 2>&1
grep -m 12 -A12 "Invalid\|ERROR SUMMARY" $OUT/noflags_memcheck.log | head -60
bash scripts/gpu_ab_multi.sh r01u default boxer_b200/_C/ab/lib_t128f7.so
