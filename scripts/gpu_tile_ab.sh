#!/bin/bash
# A/B of tile-kernel builds + ncu of one of them.  usage: gpurun -- 'PROF_LIB=<lib> bash scripts/gpu_tile_ab.sh <tag> <lib> [<lib> ...]'
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for lib in "$@"; do
  name=$(basename $lib .so)
  BOXER_B200_LIB=$lib timeout 300 python scripts/tile_check.py $OUT/check_$name.json --bwd --only ${ONLY:-K4_box,K4_trained,K4_uniform,K2_box,small_K4_oob} > $OUT/check_$name.log 2>&1
  echo "== $name"; cat $OUT/check_$name.log
done
if [ -n "$PROF_LIB" ]; then
for k in fwd bwd; do
  BOXER_B200_LIB=$PROF_LIB timeout 400 ncu --set full --clock-control none --import-source on -k "regex:box_${k}_tile" -s 2 -c 1 -f -o $OUT/${k}_tile \
      python scripts/prof_driver.py --workload enc --K 4 --path tile > $OUT/ncu_$k.log 2>&1
  if [ -f $OUT/${k}_tile.ncu-rep ]; then
    python scripts/ncu_summary.py $OUT/${k}_tile.ncu-rep > $OUT/${k}_tile.summary.txt 2>&1
    ncu -i $OUT/${k}_tile.ncu-rep --page source --csv --print-source sass > $OUT/${k}_tile.source_sass.csv 2>/dev/null
    rm -f $OUT/${k}_tile.ncu-rep
    grep -v "stalled_\(drain\|membar\|tex\|lg_thr\|dispatch\|mio\)" $OUT/${k}_tile.summary.txt
  else
    tail -5 $OUT/ncu_$k.log
  fi
done
fi
