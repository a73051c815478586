"""Tile kernels vs window kernels: agreement and timing (development aid; the parity tests proper are in tests/)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boxer_b200 import ops
from boxer_b200 import workloads as W

BWD = "--bwd" in sys.argv


def time_call(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


cases = {
    "small_K4_oob": lambda: W.coco_encoder(K=4, image=(72, 100), oob=0.05, device="cuda"),
    "small_K3_B2": lambda: W.coco_encoder(K=3, image=(96, 128), oob=0.1, B=2, device="cuda"),
    "small_uniform": lambda: W.coco_encoder(K=4, dist="uniform", image=(72, 100), device="cuda"),
    "K4_box": lambda: W.coco_encoder(K=4, device="cuda"),
    "K4_trained": lambda: W.coco_encoder(K=4, dist="trained", device="cuda"),
    "K4_uniform": lambda: W.coco_encoder(K=4, dist="uniform", device="cuda"),
    "K2_box": lambda: W.coco_encoder(K=2, device="cuda"),
    "box3d_K2": lambda: W.box3d_encoder(device="cuda"),
}
only = None
for i, a_ in enumerate(sys.argv):
    if a_ == "--only":
        only = sys.argv[i + 1].split(",")
res = {}
for name, mk in cases.items():
    if only and name not in only:
        continue
    w = mk()
    for dt in (torch.float32, torch.bfloat16):
        v = w.value.to(dt)
        a = (v, w.shapes, w.level_start, w.loc, w.weights[0])
        go = torch.randn(v.shape[0], v.shape[1], 256, device="cuda", dtype=dt)
        r = {}
        outs = {}
        for path in ("window", "tile"):
            ops.set_kernel_path(path)
            try:
                outs[path] = ops.box_attn_forward(*a, 64)
                torch.cuda.synchronize()
                r[path + "_fwd_ms"] = round(time_call(lambda: ops.box_attn_forward(*a, 64)), 4)
                if BWD:
                    outs[path + "_g"] = ops.box_attn_backward(*a, go, 64)
                    torch.cuda.synchronize()
                    r[path + "_bwd_ms"] = round(time_call(lambda: ops.box_attn_backward(*a, go, 64)), 4)
            except Exception as e:
                r[path + "_error"] = str(e)[:200]
            ops.set_kernel_path("auto")
        if "window" in outs and "tile" in outs:
            r["fwd_rel_diff"] = rel(outs["tile"].float(), outs["window"].float())
            if BWD and "tile_g" in outs:
                for i, k in enumerate(("gv", "gl", "ga")):
                    r[k + "_rel_diff"] = rel(outs["tile_g"][i].float(), outs["window_g"][i].float())
        res[f"{name}_{'f32' if dt == torch.float32 else 'bf16'}"] = r
        print(name, dt, r, flush=True)
out = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None
if out:
    json.dump(res, open(out, "w"), indent=1)
