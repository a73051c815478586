"""Fused box -> grid -> attention op (bxr_box_grid_attn_*) and the deterministic backward at the headline size: ms per launch.
A/B helper (BOXER_B200_LIB picks the library)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import boxer_b200
from boxer_b200 import ops
from boxer_b200 import workloads as W


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / n, 4)


dev = "cuda"
res = {"lib": os.path.basename(os.environ.get("BOXER_B200_LIB", "libboxattn_b200.so"))}
w = W.coco_encoder(K=4, device=dev)
S = w.value.shape[1]
for K in (4, 3):
    torch.manual_seed(0)
    mod = boxer_b200.BoxAttention(256, 4, 8, K).to(dev)
    with torch.no_grad():
        mod.linear_box_weight.normal_(0, 0.01)
        q = torch.randn(1, S, 256, device=dev)
        refw = W.encoder_ref_windows(W.fpn_levels(), 1, dev)
        boxes, _ = mod._boxes_and_angles(q, refw)
        boxes = boxes.contiguous()
        logits = torch.randn(1, S, 8, 4, K, K, device=dev)
        attn = torch.softmax(logits.view(1, S, 8, -1), -1).view_as(logits)
        kidx = mod.kernel_indices
        go = torch.randn(1, S, 256, device=dev)
        a = (w.value, w.shapes, w.level_start, boxes, None, None, kidx)
        res[f"fused_K{K}"] = [t(lambda: ops.box_grid_attn_forward(*a, attn, 64)), t(lambda: ops.box_grid_attn_backward(*a, attn, go, 64))]
        res[f"fused_softmax_K{K}"] = [t(lambda: ops.box_grid_attn_forward(*a, logits, 64, softmax=True)),
                                      t(lambda: ops.box_grid_attn_backward(*a, attn, go, 64, softmax=True))]
go = torch.randn(1, S, 256, device=dev)
ops.set_deterministic(True)
res["det_bwd_K4"] = t(lambda: ops.box_attn_backward(w.value, w.shapes, w.level_start, w.loc, w.weights[0], go, 64), 8)
ops.set_deterministic(None)
print(json.dumps(res))
