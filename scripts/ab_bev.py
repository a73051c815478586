"""BEV / decoder-like (wide boxes: every level takes the per-point path) timing per kernel family."""
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
r = W.bev_rotated(B=8, device="cuda")
go = torch.randn(8, 1000, 128, device="cuda")
a = (r.value, r.shapes, r.level_start, r.loc, r.weights[0])
for path in ("auto", "window", "point"):
    ops.set_kernel_path(path)
    res[f"bev_B8:{path}"] = (round(time_call(lambda: ops.box_attn_forward(*a, 64)), 4), round(time_call(lambda: ops.box_attn_backward(*a, go, 64)), 4))
ops.set_kernel_path("auto")
print(json.dumps(res))
