#!/usr/bin/env python
"""Same-GPU timing of the UNMODIFIED reference CUDA kernels (oracle/_ref) next to ours.
Not part of bench.py (whose numbers never involve oracle/); writes gpurun_out/<tag>/vs_reference_cuda.json.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from boxer_b200 import ops
from boxer_b200 import workloads as W
from oracle import ref_cuda


def time_call(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main(out_path, headline_only=False):
    ref = ref_cuda.load()
    res = {}
    cases = [("enc_K4_box", lambda: W.coco_encoder(K=4, device="cuda"))]
    if not headline_only:
        cases += [("enc_K2_box", lambda: W.coco_encoder(K=2, device="cuda")),
                  ("enc_K4_uniform", lambda: W.coco_encoder(K=4, dist="uniform", device="cuda")),
                  ("enc_K4_trained", lambda: W.coco_encoder(K=4, dist="trained", device="cuda")),
                  ("dec_K2", lambda: W.coco_decoder(K=2, device="cuda")), ("bev_B8_K3", lambda: W.bev_rotated(B=8, device="cuda")),
                  ("box3d_enc_K2", lambda: W.box3d_encoder(device="cuda"))]
    for name, mk in cases:
        w = mk()
        B, Nq = w.loc.shape[:2]
        go = torch.randn(B, Nq, w.value.shape[2] * w.value.shape[3], device="cuda")
        a = (w.value, w.shapes, w.level_start, w.loc, w.weights[0])
        r = {"n_samples": w.n_samples}
        r["ours_fwd_ms"] = time_call(lambda: ops.box_attn_forward(*a, 64))
        r["ours_bwd_ms"] = time_call(lambda: ops.box_attn_backward(*a, go, 64))
        r["ref_fwd_ms"] = time_call(lambda: ref.box_attn_forward(*a, 64))
        r["ref_bwd_ms"] = time_call(lambda: ref.box_attn_backward(*a, go, 64), reps=5)
        r["speedup_fwd"] = r["ref_fwd_ms"] / r["ours_fwd_ms"]
        r["speedup_bwd"] = r["ref_bwd_ms"] / r["ours_bwd_ms"]
        r["speedup_fwdbwd"] = (r["ref_fwd_ms"] + r["ref_bwd_ms"]) / (r["ours_fwd_ms"] + r["ours_bwd_ms"])
        res[name] = r
    for K in (() if headline_only else (14, 28)):
        m = W.coco_mask_head(K=K, device="cuda")
        go = torch.randn(1, 300, 256, device="cuda")
        gm = torch.randn(1, 300, K * K, 256, device="cuda")
        a = (m.value, m.shapes, m.level_start, m.loc, m.weights[0], m.weights[1])
        r = {"n_samples": m.n_samples}
        r["ours_fwd_ms"] = time_call(lambda: ops.instance_attn_forward(*a, 64))
        r["ours_bwd_ms"] = time_call(lambda: ops.instance_attn_backward(*a, go, gm, 64))
        r["ref_fwd_ms"] = time_call(lambda: ref.instance_attn_forward(*a, 64))
        r["ref_bwd_ms"] = time_call(lambda: ref.instance_attn_backward(*a, go, gm, 64), reps=5)
        r["speedup_fwd"] = r["ref_fwd_ms"] / r["ours_fwd_ms"]
        r["speedup_bwd"] = r["ref_bwd_ms"] / r["ours_bwd_ms"]
        res[f"mask_K{K}"] = r
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    for k, r in res.items():
        print(k, {a: round(b, 3) for a, b in r.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else os.path.join(ROOT, "gpurun_out", "vs_reference_cuda.json"),
         headline_only="--headline-only" in sys.argv)
