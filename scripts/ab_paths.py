"""A/B of kernel families on one GPU: fwd/bwd ms of a few workloads per ops.set_kernel_path() choice."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boxer_b200 import ops
from boxer_b200 import workloads as W

def time_call(fn, reps=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

paths = sys.argv[1:] or ["auto", "staged"]
res = {}
for name, mk in (("K4_box", lambda: W.coco_encoder(K=4, device="cuda")), ("K4_uni", lambda: W.coco_encoder(K=4, dist="uniform", device="cuda")),
                 ("K2_box", lambda: W.coco_encoder(K=2, device="cuda")), ("K4_box_bf16", lambda: W.coco_encoder(K=4, device="cuda"))):
    w = mk()
    v = w.value.bfloat16() if name.endswith("bf16") else w.value
    go = torch.randn(1, v.shape[1], 256, device="cuda", dtype=v.dtype)
    a = (v, w.shapes, w.level_start, w.loc, w.weights[0])
    for path in paths:
        ops.set_kernel_path(path)
        res[f"{name}:{path}"] = (round(time_call(lambda: ops.box_attn_forward(*a, 64)), 4),
                                 round(time_call(lambda: ops.box_attn_backward(*a, go, 64)), 4))
    ops.set_kernel_path("auto")
print(json.dumps(res, indent=1))
