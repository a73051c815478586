#!/bin/bash
# usage: gpurun -- 'bash scripts/gpu_path_ab.sh <tag> "<path_ab args>" [ncu]'
TAG=$1
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python scripts/path_ab.py $OUT/path_ab.json $2 > $OUT/path_ab.log 2>&1; cat $OUT/path_ab.log | cut -c1-700
if [ "$3" == "ncu" ]; then
for k in fwd bwd; do
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:box_${k}_win" -s 2 -c 1 -f -o $OUT/${k}_win \
      python scripts/prof_driver.py --workload enc --K 4 > $OUT/ncu_$k.log 2>&1
  if [ -f $OUT/${k}_win.ncu-rep ]; then
    python scripts/ncu_summary.py $OUT/${k}_win.ncu-rep > $OUT/${k}_win.summary.txt 2>&1
    ncu -i $OUT/${k}_win.ncu-rep --page source --csv --print-source sass > $OUT/${k}_win.source_sass.csv 2>/dev/null
    rm -f $OUT/${k}_win.ncu-rep
    grep -v "stalled_\(drain\|membar\|tex\|lg_thr\|dispatch\|mio\)" $OUT/${k}_win.summary.txt
  else
    tail -5 $OUT/ncu_$k.log
  fi
done
fi
