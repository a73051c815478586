"""fwd/bwd ms of the mask-head (instance attention) workloads with whatever library BOXER_B200_LIB points to."""
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
for K in (14, 28):
    for dt in (torch.float32, torch.bfloat16):
        m = W.coco_mask_head(K=K, device="cuda")
        v = m.value.to(dt)
        go = torch.randn(1, 300, 256, device="cuda", dtype=dt)
        gm = torch.randn(1, 300, K * K, 256, device="cuda", dtype=dt)
        sw, lw = m.weights
        tf = time_call(lambda: ops.instance_attn_forward(v, m.shapes, m.level_start, m.loc, sw, lw, 64))
        tb = time_call(lambda: ops.instance_attn_backward(v, m.shapes, m.level_start, m.loc, sw, lw, go, gm, 64))
        res[f"mask_K{K}_{'f32' if dt == torch.float32 else 'bf16'}"] = (round(tf, 4), round(tb, 4))
print(json.dumps(res))
