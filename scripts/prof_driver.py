"""Tiny driver for ncu: a few fwd+bwd launches of one workload (no timing here -- numbers taken
under a profiler are never bench values)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from boxer_b200 import ops
from boxer_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="enc")     # enc | mask | dec | bev
ap.add_argument("--K", type=int, default=4)
ap.add_argument("--dist", default="box")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--det", action="store_true")
ap.add_argument("--path", default="auto")
a = ap.parse_args()
dt = {"f32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
ops.set_deterministic(a.det)
ops.set_kernel_path(a.path)
if a.workload == "enc":
    w = W.coco_encoder(K=a.K, dist=a.dist, device="cuda")
elif a.workload == "dec":
    w = W.coco_decoder(K=a.K, device="cuda")
elif a.workload == "bev":
    w = W.bev_rotated(B=8, device="cuda")
else:
    w = W.coco_mask_head(K=a.K, device="cuda")
v = w.value.to(dt)
B, Nq = w.loc.shape[:2]
C = v.shape[2] * v.shape[3]
go = torch.randn(B, Nq, C, device="cuda", dtype=dt)
for _ in range(a.iters):
    if w.instance:
        gm = torch.randn(B, Nq, a.K * a.K, C, device="cuda", dtype=dt)
        ops.instance_attn_forward(v, w.shapes, w.level_start, w.loc, w.weights[0], w.weights[1], 64)
        ops.instance_attn_backward(v, w.shapes, w.level_start, w.loc, w.weights[0], w.weights[1], go, gm, 64)
    else:
        ops.box_attn_forward(v, w.shapes, w.level_start, w.loc, w.weights[0], 64)
        ops.box_attn_backward(v, w.shapes, w.level_start, w.loc, w.weights[0], go, 64)
torch.cuda.synchronize()
