#!/bin/bash
# ncu full capture (+ SASS page for scripts/ncu_lines.py) of one kernel: gpu_prof1.sh <tag> <kernel regex> [prof_driver args...]
TAG=$1; KRE=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s 2 -c 1 -f -o $OUT/k \
    python scripts/prof_driver.py "$@" > $OUT/ncu.log 2>&1
if [ -f $OUT/k.ncu-rep ]; then
    python scripts/ncu_summary.py $OUT/k.ncu-rep > $OUT/summary.txt 2>&1
    ncu -i $OUT/k.ncu-rep --page source --csv --print-source sass > $OUT/source_sass.csv 2>/dev/null
    ncu -i $OUT/k.ncu-rep --page details > $OUT/details.txt 2>/dev/null
    rm -f $OUT/k.ncu-rep
fi
tail -3 $OUT/ncu.log
