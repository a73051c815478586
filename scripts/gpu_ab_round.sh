#!/bin/bash
# One gpurun call: A/B an alternative build of the library against the in-tree one (scripts/ab_test.py), carry on
# with whichever is faster on the backward (BOXER_B200_LIB), then the full round (scripts/gpu_round.sh) and a few
# extra ncu captures of the non-headline kernels.
# usage: gpurun --timeout 1800 -- 'bash scripts/gpu_ab_round.sh <tag> <alt lib path>'
TAG=${1:-r01}
ALT=$2
OUT=gpurun_out/$TAG
mkdir -p $OUT
python scripts/ab_test.py > $OUT/ab_default.json 2> $OUT/ab_default.err
BOXER_B200_LIB=$ALT python scripts/ab_test.py > $OUT/ab_alt.json 2> $OUT/ab_alt.err
python scripts/ab_test.py > $OUT/ab_default2.json 2>> $OUT/ab_default.err
BOXER_B200_LIB=$ALT python scripts/ab_test.py > $OUT/ab_alt2.json 2>> $OUT/ab_alt.err
cat $OUT/ab_default.json $OUT/ab_alt.json $OUT/ab_default2.json $OUT/ab_alt2.json
WIN=$(python - $OUT <<'P'
import json, sys, os
o = sys.argv[1]
def bwd(names):
    t = 0.0
    for n in names:
        r = json.load(open(os.path.join(o, n)))
        t += sum(r[k][1] for k in ("K4_box", "K2_box", "K4_uni"))
    return t
d, a = bwd(["ab_default.json", "ab_default2.json"]), bwd(["ab_alt.json", "ab_alt2.json"])
print("alt" if a < 0.985 * d else "default")
P
)
echo "winner: $WIN" | tee $OUT/ab_winner.txt
if [ "$WIN" = "alt" ]; then export BOXER_B200_LIB=$ALT; fi
bash scripts/gpu_round.sh $TAG full
# extra captures (diagnostics for the non-headline kernels)
bash scripts/gpu_prof1.sh $TAG/inst_fwd_K28_bf16 "inst_fwd" --workload mask --K 28 --dtype bf16
bash scripts/gpu_prof1.sh $TAG/inst_bwd_K28_bf16 "inst_bwd" --workload mask --K 28 --dtype bf16
bash scripts/gpu_prof1.sh $TAG/det_bwd_K4 "box_bwd_win" --workload enc --K 4 --det
bash scripts/gpu_prof1.sh $TAG/bwd_K4_uniform "box_bwd_win" --workload enc --K 4 --dist uniform
bash scripts/gpu_prof1.sh $TAG/bev_bwd "attn_bwd|box_bwd" --workload bev
du -sh $OUT
