#!/bin/bash
# usage: gpurun -- 'bash scripts/gpu_round2.sh <tag>'  : hook A/B builds, host overhead (shim vs ctypes vs reference), full tests
TAG=$1
OUT=gpurun_out/$TAG
mkdir -p $OUT
python scripts/host_overhead.py $OUT/host_overhead_shim.json 2>&1 | tail -8
BOXER_B200_NO_SHIM=1 python scripts/host_overhead.py $OUT/host_overhead_ctypes.json 2>&1 | tail -8
if ls gpurun_in/*.so >/dev/null 2>&1; then
  bash scripts/gpu_ab_multi.sh $TAG default gpurun_in/*.so 2>&1 | cut -c1-400
fi
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 --timeout 900 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
