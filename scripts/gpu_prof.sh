#!/bin/bash
# One gpurun call: smoke + GPU tests + ncu full captures (with per-source-line pages) of the two hot kernels.
# usage: gpurun --timeout 1200 -- 'bash scripts/gpu_prof.sh <tag> [notests]'
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "$2" != "notests" ]; then
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
tail -5 $OUT/pytest.log
fi
for d in fwd bwd; do
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:box_${d}_win" -s 2 -c 1 -f -o $OUT/${d}_enc_K4 \
    python scripts/prof_driver.py --workload enc --K 4 > $OUT/ncu_$d.log 2>&1
r=${d}_enc_K4
if [ -f $OUT/$r.ncu-rep ]; then
    python scripts/ncu_summary.py $OUT/$r.ncu-rep > $OUT/$r.summary.txt 2>&1
    ncu -i $OUT/$r.ncu-rep --page source --csv --print-source cuda > $OUT/$r.source_cuda.csv 2>/dev/null
    ncu -i $OUT/$r.ncu-rep --page source --csv --print-source sass > $OUT/$r.source_sass.csv 2>/dev/null
    ncu -i $OUT/$r.ncu-rep --page details > $OUT/$r.details.txt 2>/dev/null
    rm -f $OUT/$r.ncu-rep
fi
done
ls -la $OUT
