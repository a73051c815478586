#!/bin/bash
# A/B one alternative build against the in-tree one, then the whole GPU suite with the alternative.
# usage: gpurun --timeout 600 -- 'bash scripts/gpu_ab_full.sh <tag> <alt lib>'
TAG=$1; ALT=$2
bash scripts/gpu_ab_multi.sh $TAG default $ALT
BOXER_B200_LIB=$ALT timeout 300 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/$TAG/pytest_alt_full.log 2>&1
echo "alt full pytest: $(tail -1 gpurun_out/$TAG/pytest_alt_full.log)"
timeout 300 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/$TAG/pytest_default_full.log 2>&1
echo "default full pytest: $(tail -1 gpurun_out/$TAG/pytest_default_full.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
