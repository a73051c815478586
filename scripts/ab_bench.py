"""Time fwd/bwd of a few workloads with whatever library BOXER_B200_LIB points to (A/B builds)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boxer_b200 import ops, _native
from boxer_b200 import workloads as W

def time_call(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

res = {"lib": os.path.basename(_native.LIB_PATH)}
for name, w in (("K4_box", W.coco_encoder(K=4, device="cuda")), ("K4_uni", W.coco_encoder(K=4, dist="uniform", device="cuda")),
                ("K2_box", W.coco_encoder(K=2, device="cuda")), ("K4_trained", W.coco_encoder(K=4, dist="trained", device="cuda")),
                ("K2_uni", W.coco_encoder(K=2, dist="uniform", device="cuda"))):
    go = torch.randn(1, w.value.shape[1], 256, device="cuda")
    a = (w.value, w.shapes, w.level_start, w.loc, w.weights[0])
    res[name] = (round(time_call(lambda: ops.box_attn_forward(*a, 64)), 4), round(time_call(lambda: ops.box_attn_backward(*a, go, 64)), 4))
w = W.coco_encoder(K=4, device="cuda")
vb = w.value.to(torch.bfloat16)
gb = torch.randn(1, vb.shape[1], 256, device="cuda", dtype=torch.bfloat16)
ab = (vb, w.shapes, w.level_start, w.loc, w.weights[0])
res["K4_box_bf16"] = (round(time_call(lambda: ops.box_attn_forward(*ab, 64)), 4), round(time_call(lambda: ops.box_attn_backward(*ab, gb, 64)), 4))
print(json.dumps(res))
