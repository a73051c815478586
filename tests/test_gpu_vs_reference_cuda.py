"""GPU: our kernels vs the UNMODIFIED reference CUDA kernels (oracle/_ref, compiled for sm_100a
from /root/reference by oracle/build_ref.py) on the same device and inputs.

This is the reference itself run here -- the strongest parity evidence available: the reference's
fp32 / fp64 kernels and ours must agree to round-off (both compute in the same precision; the
atomic scatter order is the only legitimate difference)."""
import pytest
import torch

from oracle import ref_cuda
from tests import helpers

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")]
DEV = "cuda"


def _both_box(w, go, dtype):
    import boxer_b200
    ref = ref_cuda.load()
    v = w.value.to(dtype).contiguous()
    loc = w.loc.to(dtype).contiguous()
    a = w.weights[0].to(dtype).contiguous()
    g = go.to(dtype).contiguous()
    ours = (boxer_b200.ops.box_attn_forward(v, w.shapes, w.level_start, loc, a, 64),
            *boxer_b200.ops.box_attn_backward(v, w.shapes, w.level_start, loc, a, g, 64))
    theirs = (ref.box_attn_forward(v, w.shapes, w.level_start, loc, a, 64),
              *ref.box_attn_backward(v, w.shapes, w.level_start, loc, a, g, 64))
    return ours, theirs


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float64, 1e-12)], ids=["f32", "f64"])
@pytest.mark.parametrize("path", ["auto", "window", "point"])
@pytest.mark.parametrize("case", ["enc_K4", "enc_K2_uniform_oob", "dec_K2", "bev_rot_K3"])
def test_box_matches_reference_cuda(case, path, dtype, tol):
    import boxer_b200
    from boxer_b200 import workloads as W
    w = {
        "enc_K4": lambda: W.coco_encoder(K=4, image=(160, 256), device=DEV, oob=0.02),
        "enc_K2_uniform_oob": lambda: W.coco_encoder(K=2, dist="uniform", image=(160, 256), device=DEV, oob=0.2, B=2),
        "dec_K2": lambda: W.coco_decoder(Nq=300, K=2, image=(160, 256), device=DEV),
        "bev_rot_K3": lambda: W.bev_rotated(Nq=500, K=3, size=117, device=DEV, head_dim=32),
    }[case]()
    B, Nq = w.loc.shape[:2]
    go = torch.randn(B, Nq, w.value.shape[2] * w.value.shape[3], device=DEV)
    boxer_b200.ops.set_kernel_path(path)
    try:
        ours, theirs = _both_box(w, go, dtype)
    finally:
        boxer_b200.ops.set_kernel_path("auto")
    # fp32: both sides round loc*size-0.5 in fp32, but FMA contraction may differ by an ulp, which can
    # flip the cell of a sample sitting on a grid line -> compare grad_loc away from grid lines only
    from tests.test_gpu_ops import _near_cell_boundary
    keep = (~_near_cell_boundary(w.loc, w.shapes, eps=1e-3 if dtype == torch.float32 else 1e-9))[..., None].to(DEV)
    names = ("out", "grad_value", "grad_loc", "grad_attn")
    for n, o, t in zip(names, ours, theirs):
        assert o.shape == t.shape or o.numel() == t.numel(), n
        t = t.reshape(o.shape)
        if n == "grad_loc":
            o, t = o * keep, t * keep
        err = helpers.rel_err(o, t)
        assert err <= tol, f"{case}/{path}/{n}: {err:.3e}"


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float64, 1e-12)], ids=["f32", "f64"])
@pytest.mark.parametrize("K,Nq", [(14, 40), (2, 300)])
def test_instance_matches_reference_cuda(K, Nq, dtype, tol):
    import boxer_b200
    from boxer_b200 import workloads as W
    ref = ref_cuda.load()
    w = W.coco_mask_head(Nq=Nq, K=K, image=(160, 256), device=DEV)
    v = w.value.to(dtype)
    loc, sw, lw = (t.to(dtype).contiguous() for t in (w.loc, *w.weights))
    go = torch.randn(1, Nq, 256, device=DEV, dtype=dtype)
    gm = torch.randn(1, Nq, K * K, 256, device=DEV, dtype=dtype)
    ours = (*boxer_b200.ops.instance_attn_forward(v, w.shapes, w.level_start, loc, sw, lw, 64),
            *boxer_b200.ops.instance_attn_backward(v, w.shapes, w.level_start, loc, sw, lw, go, gm, 64))
    theirs = (*ref.instance_attn_forward(v, w.shapes, w.level_start, loc, sw, lw, 64),
              *ref.instance_attn_backward(v, w.shapes, w.level_start, loc, sw, lw, go, gm, 64))
    from tests.test_gpu_ops import _near_cell_boundary
    keep = (~_near_cell_boundary(w.loc, w.shapes, eps=1e-3 if dtype == torch.float32 else 1e-9))[..., None].to(DEV)
    names = ("out", "mask_out", "grad_value", "grad_loc", "grad_spatial_w", "grad_level_w")
    for n, o, t in zip(names, ours, theirs):
        t = t.reshape(o.shape)
        if n == "grad_loc":
            o, t = o * keep, t * keep
        err = helpers.rel_err(o, t)
        assert err <= tol, f"K={K}/{n}: {err:.3e}"
