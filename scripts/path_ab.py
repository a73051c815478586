"""A/B of kernel paths inside one library build: agreement and timing (development aid).
usage: path_ab.py <out.json> --paths window-memory-order,auto [--only case,case] [--reps N]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boxer_b200 import ops
from boxer_b200 import workloads as W


def arg(name, default=None):
    for i, a in enumerate(sys.argv):
        if a == name:
            return sys.argv[i + 1]
    return default


def time_call(fn, reps):
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
    "K4_box": lambda: W.coco_encoder(K=4, device="cuda"),
    "K4_trained": lambda: W.coco_encoder(K=4, dist="trained", device="cuda"),
    "K4_uniform": lambda: W.coco_encoder(K=4, dist="uniform", device="cuda"),
    "K2_box": lambda: W.coco_encoder(K=2, device="cuda"),
    "K2_uniform": lambda: W.coco_encoder(K=2, dist="uniform", device="cuda"),
    "K4_box_B2": lambda: W.coco_encoder(K=4, B=2, device="cuda"),
    "box3d_K2": lambda: W.box3d_encoder(device="cuda"),
    "small_K4_oob": lambda: W.coco_encoder(K=4, image=(72, 100), oob=0.05, device="cuda"),
}
paths = arg("--paths", "window-memory-order,auto").split(",")
only = arg("--only")
only = only.split(",") if only else None
reps = int(arg("--reps", "30"))
res = {}
for rnd in (1, 2):
    for name, mk in cases.items():
        if only and name not in only:
            continue
        w = mk()
        for dt in (torch.float32, torch.bfloat16):
            v = w.value.to(dt)
            a = (v, w.shapes, w.level_start, w.loc, w.weights[0])
            go = torch.randn(v.shape[0], v.shape[1], 256, device="cuda", dtype=dt)
            key = f"{name}_{'f32' if dt == torch.float32 else 'bf16'}"
            r = res.setdefault(key, {})
            outs = {}
            for path in paths:
                ops.set_kernel_path(path)
                try:
                    outs[path] = (ops.box_attn_forward(*a, 64), ops.box_attn_backward(*a, go, 64))
                    torch.cuda.synchronize()
                    r.setdefault(path + "_fwd_ms", []).append(round(time_call(lambda: ops.box_attn_forward(*a, 64), reps), 4))
                    r.setdefault(path + "_bwd_ms", []).append(round(time_call(lambda: ops.box_attn_backward(*a, go, 64), reps), 4))
                except Exception as e:
                    r[path + "_error"] = str(e)[:200]
                ops.set_kernel_path("auto")
            if rnd == 1 and len(outs) == len(paths):
                base = outs[paths[0]]
                for path in paths[1:]:
                    r[path + "_diff"] = [rel(outs[path][0].float(), base[0].float())] + [rel(x.float(), y.float()) for x, y in zip(outs[path][1], base[1])]
for k, r in res.items():
    print(k, r, flush=True)
out = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None
if out:
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    json.dump(res, open(out, "w"), indent=1)
