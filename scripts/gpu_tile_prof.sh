#!/bin/bash
# Tile-kernel development round: agreement + timing of the in-tree build and of alternative builds, ncu capture of the
# tile forward (and backward with "bwd").  usage: gpurun -- 'bash scripts/gpu_tile_prof.sh <tag> [bwd] [alt.so ...]'
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
BWD=""
if [ "$1" == "bwd" ]; then BWD="--bwd"; shift; fi
timeout 300 python scripts/tile_check.py $OUT/check_default.json $BWD > $OUT/check_default.log 2>&1; tail -20 $OUT/check_default.log
for lib in "$@"; do
  name=$(basename $lib .so)
  BOXER_B200_LIB=$lib timeout 300 python scripts/tile_check.py $OUT/check_$name.json $BWD --only K4_box,K4_trained,K4_uniform,K2_box > $OUT/check_$name.log 2>&1
  echo "== $name"; cat $OUT/check_$name.log
done
for k in fwd $( [ -n "$BWD" ] && echo bwd ); do
  BOXER_B200_LIB=${PROF_LIB:-} timeout 400 ncu --set full --clock-control none --import-source on -k "regex:box_${k}_tile" -s 2 -c 1 -f -o $OUT/${k}_tile \
      python scripts/prof_driver.py --workload enc --K 4 --path tile > $OUT/ncu_$k.log 2>&1
  if [ -f $OUT/${k}_tile.ncu-rep ]; then
    python scripts/ncu_summary.py $OUT/${k}_tile.ncu-rep > $OUT/${k}_tile.summary.txt 2>&1
    ncu -i $OUT/${k}_tile.ncu-rep --page source --csv --print-source sass > $OUT/${k}_tile.source_sass.csv 2>/dev/null
    ncu -i $OUT/${k}_tile.ncu-rep --page details > $OUT/${k}_tile.details.txt 2>/dev/null
    rm -f $OUT/${k}_tile.ncu-rep
    cat $OUT/${k}_tile.summary.txt
  else
    tail -5 $OUT/ncu_$k.log
  fi
done
ls -la $OUT
