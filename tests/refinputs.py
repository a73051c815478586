"""Seeded inputs shared by the golden generator and the tests.

``box_test_sequence`` / ``instance_test_sequence`` replay the global-RNG draw
order of the reference scripts' ``__main__`` blocks
(/root/reference/tests/box_attn_test.py:192-201,
/root/reference/tests/instance_attn_test.py:295-304): ``torch.manual_seed(3)``,
then seven ``check_gradient_numerical`` draws (D = 30, 32, 64, 71, 1025, 2048,
3096), then ``check_forward("float")``, ``check_forward("double")`` and
``check_forward_and_backward()``.  The reference draws on the CPU generator and
then ``.cuda()``s, so the same numbers are reproducible without a GPU.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch

# reference test constants (box_attn_test.py:45-49, instance_attn_test.py:66-73)
N, M = 1, 2
LQ, L = 2, 2
REF_SHAPES = [(6, 4), (3, 2)]
GRADCHECK_D = [30, 32, 64, 71, 1025, 2048, 3096]


def shapes_tensor(shapes):
    return torch.tensor(shapes, dtype=torch.long)


def level_start_index(shapes):
    shapes = shapes_tensor(shapes) if not torch.is_tensor(shapes) else shapes
    return torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))


def digest(inp) -> np.ndarray:
    """sha256 of the input tensors' bytes (detects RNG drift when inputs are regenerated)."""
    h = hashlib.sha256()
    for k in sorted(inp):
        v = inp[k]
        if torch.is_tensor(v):
            h.update(k.encode())
            h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


def side_rand(shape, seed, dtype=torch.float64, lo=0.0, hi=1.0):
    g = torch.Generator().manual_seed(int(seed))
    return torch.rand(*shape, generator=g, dtype=dtype) * (hi - lo) + lo


# ------------------------------------------------------------------ box op
def _box_draw(D, P=2):
    S = sum(h * w for h, w in REF_SHAPES)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, LQ, M, L, P, 2)
    attn = torch.rand(N, LQ, M, L, P) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, loc, attn


def _pack_box(value, loc, attn, double):
    if double:
        value, loc, attn = value.double(), loc.double(), attn.double()
    sh = shapes_tensor(REF_SHAPES)
    return {"value": value, "loc": loc, "attn": attn, "shapes": sh, "level_start": level_start_index(sh)}


def box_test_sequence():
    torch.manual_seed(3)
    out = {}
    for D in GRADCHECK_D:
        out[f"gradcheck_D{D}"] = _pack_box(*_box_draw(D), double=True)
    out["fwd_float"] = _pack_box(*_box_draw(2), double=False)
    out["fwd_double"] = _pack_box(*_box_draw(2), double=True)
    out["fwdbwd_double"] = _pack_box(*_box_draw(2), double=True)
    return out


def _wide_box(seed, B, shapes, H, D, Nq, K, lo, hi, six_d, snap=False):
    S = sum(h * w for h, w in shapes)
    Lw = len(shapes)
    value = side_rand((B, S, H, D), seed, lo=-1.0, hi=1.0)
    loc = side_rand((B, Nq, H, Lw, K * K, 2), seed + 1, lo=lo, hi=hi)
    if snap:
        # a quarter of the coordinates exactly on pixel centres / borders:
        # exercises floor() at integers and the open window test
        flat = loc.view(-1)
        idx = torch.arange(0, flat.numel(), 4)
        flat[idx] = torch.round(flat[idx] * 8) / 8
    logits = side_rand((B, Nq, H, Lw * K * K), seed + 2, lo=-2.0, hi=2.0)
    attn = torch.softmax(logits, -1).view(B, Nq, H, Lw, K, K)
    if not six_d:
        attn = attn.reshape(B, Nq, H, Lw, K * K)
    sh = shapes_tensor(shapes)
    return {"value": value, "loc": loc, "attn": attn, "shapes": sh, "level_start": level_start_index(sh)}


def box_wide_cases():
    return {
        "wide_oob_3lvl": _wide_box(11, 2, [(16, 12), (8, 6), (4, 3)], 4, 32, 37, 2, -0.25, 1.25, True),
        "wide_snap_k3": _wide_box(21, 1, [(9, 7)], 2, 16, 11, 3, -0.125, 1.125, False, snap=True),
        "wide_d8_4lvl": _wide_box(31, 3, [(10, 10), (5, 5), (3, 2), (1, 1)], 8, 8, 5, 4, -0.1, 1.1, True),
    }


# ------------------------------------------------------------- instance op
def _inst_pack(value, loc, sw, lw, double, K=2):
    if double:
        value, loc, sw, lw = value.double(), loc.double(), sw.double(), lw.double()
    sh = shapes_tensor(REF_SHAPES)
    return {"value": value, "loc": loc, "spatial_w": sw, "level_w": lw, "shapes": sh,
            "level_start": level_start_index(sh), "mask_size": K}


def _inst_draw_gradcheck(D, P=4):
    S = sum(h * w for h, w in REF_SHAPES)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, LQ, M, L, P, 2)
    attn = torch.rand(N, LQ, M, L, P) + 1e-5
    sw = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).clone()
    lw = (attn / attn.sum(-2, keepdim=True)).clone()
    return value, loc, sw, lw


def _inst_draw_check(P=4, MS=2):
    S = sum(h * w for h, w in REF_SHAPES)
    value = torch.rand(N, S, M, 2) * 0.01
    loc = torch.rand(N, LQ, M, L, P, 2)
    attn = torch.rand(N, LQ, M, L, MS, MS) + 1e-5
    sw = attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True).sum(-3, keepdim=True)
    lw = attn / attn.sum(-3, keepdim=True)
    return value, loc, sw, lw


def instance_test_sequence():
    torch.manual_seed(3)
    out = {}
    for D in GRADCHECK_D:
        out[f"gradcheck_D{D}"] = _inst_pack(*_inst_draw_gradcheck(D), double=True)
    out["fwd_float"] = _inst_pack(*_inst_draw_check(), double=False)
    out["fwd_double"] = _inst_pack(*_inst_draw_check(), double=True)
    out["fwdbwd_double"] = _inst_pack(*_inst_draw_check(), double=True)
    return out


def _wide_inst(seed, B, shapes, H, D, Nq, K, lo, hi):
    S = sum(h * w for h, w in shapes)
    Lw = len(shapes)
    value = side_rand((B, S, H, D), seed, lo=-1.0, hi=1.0)
    loc = side_rand((B, Nq, H, Lw, K * K, 2), seed + 1, lo=lo, hi=hi)
    logits = side_rand((B, Nq, H, Lw, K, K), seed + 2, lo=-2.0, hi=2.0)
    sw = torch.softmax(logits.view(B, Nq, H, -1), -1).view_as(logits)
    lw = torch.softmax(logits, 3)
    sh = shapes_tensor(shapes)
    return {"value": value, "loc": loc, "spatial_w": sw, "level_w": lw, "shapes": sh,
            "level_start": level_start_index(sh), "mask_size": K}


def instance_wide_cases():
    return {
        "wide_oob_3lvl_k4": _wide_inst(41, 2, [(16, 12), (8, 6), (4, 3)], 4, 32, 9, 4, -0.25, 1.25),
        "wide_1lvl_k2": _wide_inst(51, 1, [(7, 5)], 2, 16, 6, 2, -0.1, 1.1),
        "wide_d8_k6": _wide_inst(61, 1, [(10, 10), (5, 5)], 8, 8, 3, 6, 0.0, 1.0),
    }


# ----------------------------------------------------------------- modules
def module_cases():
    shapes = [(12, 10), (6, 5), (3, 3)]
    return {
        "box_3d_refs_masked": dict(cls="BoxAttention", seed=5, ctor=dict(d_model=64, num_level=3, num_head=4, kernel_size=2),
                                   shapes=shapes, B=2, Nq=13, ref_dim=3, ref_last=4, mask=True, ratios=True),
        "box_4d_refs_k3": dict(cls="BoxAttention", seed=6, ctor=dict(d_model=32, num_level=3, num_head=2, kernel_size=3),
                               shapes=shapes, B=1, Nq=7, ref_dim=4, ref_last=4, mask=False, ratios=False),
        "inst_train": dict(cls="InstanceAttention", seed=7, ctor=dict(d_model=64, num_level=3, num_head=4, kernel_size=4),
                           shapes=shapes, B=2, Nq=5, ref_dim=3, ref_last=4, mask=True, ratios=True, inferencing=False),
        "inst_infer": dict(cls="InstanceAttention", seed=8, ctor=dict(d_model=64, num_level=3, num_head=4, kernel_size=4),
                           shapes=shapes, B=2, Nq=5, ref_dim=4, ref_last=4, mask=False, ratios=False, inferencing=True),
        "box3d_rot": dict(cls="Box3dAttention", seed=9, ctor=dict(d_model=64, num_level=2, num_head=4, with_rotation=True, kernel_size=2),
                          shapes=shapes[:2], B=2, Nq=9, ref_dim=3, ref_last=7, mask=False, ratios=False),
        "box3d_norot_4d": dict(cls="Box3dAttention", seed=10, ctor=dict(d_model=64, num_level=2, num_head=4, with_rotation=False, kernel_size=3),
                               shapes=shapes[:2], B=1, Nq=9, ref_dim=4, ref_last=5, mask=False, ratios=False),
    }


def randomize_module(mod, seed):
    """The reference init zeroes the box/attn projections; use non-trivial values instead."""
    g = torch.Generator().manual_seed(int(seed))
    with torch.no_grad():
        for name, p in mod.named_parameters():
            scale = 0.5 if "bias" in name else 0.2
            p.copy_((torch.rand(p.shape, generator=g, dtype=torch.float64) - 0.5) * 2 * scale)


def module_inputs(spec, dtype=torch.float64):
    B, Nq = spec["B"], spec["Nq"]
    C = spec["ctor"]["d_model"]
    H = spec["ctor"]["num_head"]
    shapes = spec["shapes"]
    Lm = len(shapes)
    S = sum(h * w for h, w in shapes)
    seed = spec["seed"] * 100
    query = side_rand((B, Nq, C), seed, lo=-1, hi=1).to(dtype)
    value = side_rand((B, S, C), seed + 1, lo=-1, hi=1).to(dtype)
    v_shape = shapes_tensor(shapes)
    v_start = level_start_index(v_shape)
    v_mask = None
    if spec["mask"]:
        v_mask = side_rand((B, S), seed + 2) > 0.8
    ratios = None
    if spec["ratios"]:
        ratios = side_rand((B, 1, 1, Lm, 1, 2), seed + 3, lo=0.6, hi=1.0).to(dtype)
    last = spec["ref_last"]
    if spec["ref_dim"] == 3:
        ref = side_rand((B, Nq, last), seed + 4)
    else:
        ref = side_rand((B, Nq, H, last), seed + 4)
    ref = ref.clone()
    ref[..., :2] = 0.1 + 0.8 * ref[..., :2]
    ref[..., 2:4] = 0.05 + 0.4 * ref[..., 2:4]
    return query, value, v_shape, v_mask, v_start, ratios, ref.to(dtype)
