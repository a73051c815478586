"""Deterministic backward (BXR_FLAG_DETERMINISTIC) at the headline size, fp32 and bf16: ms per launch.  A/B helper."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boxer_b200 import ops
from boxer_b200 import workloads as W


def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n, 4)


res = {"lib": os.path.basename(os.environ.get("BOXER_B200_LIB", "libboxattn_b200.so"))}
ops.set_deterministic(True)
for name, dt, kw in (("K4_box_f32", torch.float32, {}), ("K4_box_bf16", torch.bfloat16, {}), ("K4_trained_f32", torch.float32, {"dist": "trained"}),
                     ("K2_box_f32", torch.float32, {"K": 2})):
    w = W.coco_encoder(K=kw.pop("K", 4), device="cuda", **kw)
    go = torch.randn(1, w.loc.shape[1], 256, device="cuda", dtype=dt)
    a = (w.value.to(dt), w.shapes, w.level_start, w.loc, w.weights[0])
    res[name + "_det_bwd"] = t(lambda: ops.box_attn_backward(*a, go, 64))
ops.set_deterministic(None)
print(json.dumps(res))
