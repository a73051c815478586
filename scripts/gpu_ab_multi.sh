#!/bin/bash
# A/B several builds of the library in one gpurun call: scripts/ab_bench.py with each, two interleaved rounds.
# usage: gpurun --timeout 600 -- 'bash scripts/gpu_ab_multi.sh <tag> <lib> [<lib> ...]'   ("default" = the in-tree build)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for round in 1 2; do
  for lib in "$@"; do
    name=$(basename $lib .so)
    if [ "$lib" = "default" ]; then
      python scripts/ab_bench.py > $OUT/${name}_$round.json 2>> $OUT/err.log
    else
      BOXER_B200_LIB=$lib python scripts/ab_bench.py > $OUT/${name}_$round.json 2>> $OUT/err.log
    fi
    echo "$name $round: $(cat $OUT/${name}_$round.json)"
  done
done
# correctness of every alternative on the window-kernel tests
for lib in "$@"; do
  [ "$lib" = "default" ] && continue
  name=$(basename $lib .so)
  BOXER_B200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "window or golden or full_size or zero or c1_scale" > $OUT/${name}_pytest.log 2>&1
  echo "$name pytest: $(tail -1 $OUT/${name}_pytest.log)"
done
