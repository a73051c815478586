"""Parity at BASELINE.json's own sizes, element by element, against the C oracle (oracle/kernel_ref.c).

BASELINE.md section 5 asks for parity "on every config above"; these are the configs at their stated sizes:
  c2  BoxeR-2D encoder, 4 levels of 800x1333, Nq = S = 22 223, K = 4 (BASELINE) and K = 2 (reference-exact),
      fp32 and bf16, box-structured / trained-like / uniform locations -- every output element, every gradient
      element (grad_value over the whole 5.7 M-element tensor, not through an adjoint scalar);
  c4  mask head: InstanceAttnFunction, 300 queries x 28x28 (BASELINE) and 14x14 (reference-exact), fp32 and bf16;
  c5  BEV: 468x468, C = 128, Nq = 1000, 3x3 rotated grid, B = 8 frames (BASELINE) and the reference-exact BoxeR-3D
      encoder call (234^2 + 117^2, D = 32, Nq = S = 68 445, 2x2).

Two bars per tensor, both over EVERY element: the max-norm bar of BASELINE.md (fp32 1e-4, bf16 1e-2 of the tensor's
largest magnitude) and an element-wise one, |got - want| <= tol * (|want| + 16 mean|want != 0|), which holds small entries
to a scale several times tighter than the max-norm bar (the tensor's typical non-zero magnitude instead of its largest).  (Not to their own magnitude alone: grad_loc is a difference of 32-term dot
products, hy (d01 - d00) + ly (d11 - d10), so an entry that cancels to ~0 still carries the fp32 rounding of its terms --
for any fp32 implementation, the reference's kernels included.  Runs r02c / r02i: with 1 x mean, 634 of 22.6 M grad_loc
entries of the K=4 encoder sat up to 5.2x outside, with 8 x mean one entry at 1.08x; every other tensor passed at 1 x mean.
Inputs are seeded, so a run is reproducible.)
bf16 runs hand the oracle the bf16-rounded value / grad_out (the storage type is the test's input, not its error).
"""
import pytest
import torch

from tests import helpers
from tests.test_gpu_ops import DEV, TOL, _near_cell_boundary, _ops, _run_box, _run_inst, _wl_inputs

pytestmark = pytest.mark.gpu


def _check(got, want, tol, what, keep=None):
    got = torch.as_tensor(got).detach().double().cpu().reshape(-1)
    want = torch.as_tensor(want).detach().double().cpu().reshape(-1)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if keep is not None:
        keep = keep.reshape(-1)
        got, want = got[keep], want[keep]
    err = (got - want).abs()
    scale = want.abs().max().clamp_min(1e-30)
    assert float(err.max() / scale) <= tol, f"{what}: max-norm relative error {float(err.max() / scale):.3e} > {tol:g}"
    nz = want != 0
    typical = want[nz].abs().mean() if bool(nz.any()) else want.new_tensor(0.0)      # grad_value of the BEV case is 99 % zeros
    bound = tol * (want.abs() + 16 * typical)
    bad = err > bound
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())} of {bad.numel()} elements outside "
                                 f"|err| <= {tol:g} (|want| + 16 mean|want != 0|); worst ratio {float((err / bound.clamp_min(1e-300)).max()):.2f}")


def _bf16_round(t):
    return t.bfloat16().float()


def _loc_keep(w):
    return (~_near_cell_boundary(w.loc, w.shapes))[..., None].expand(*w.loc.shape[:-1], 2)


def _oracle_box_per_image(w, go):
    """fp64 oracle, one image at a time (bounds host memory for the 8-frame BEV case)."""
    from oracle import kernel_ref
    outs, gvs, gls, gas = [], [], [], []
    sh, st = w.shapes.cpu(), w.level_start.cpu()
    for b in range(w.value.shape[0]):
        v, l, a = (t[b:b + 1].detach().double().cpu() for t in (w.value, w.loc, w.weights[0]))
        outs.append(kernel_ref.box_attn_forward(v, sh, st, l, a))
        gv, gl, ga = kernel_ref.box_attn_backward(v, sh, st, l, a, go[b:b + 1].double().cpu())
        gvs.append(gv); gls.append(gl); gas.append(ga)
    return torch.cat(outs), (torch.cat(gvs), torch.cat(gls), torch.cat(gas))


def _box_case(w, dtype):
    B, Nq = w.loc.shape[:2]
    C = w.value.shape[2] * w.value.shape[3]
    go = torch.randn(B, Nq, C, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1234))
    if dtype == torch.bfloat16:
        w.value, go = _bf16_round(w.value), _bf16_round(go)
    out, grads = _run_box(_wl_inputs(w), dtype, go)
    ref_out, ref = _oracle_box_per_image(w, go)
    tol = TOL[dtype]
    _check(out, ref_out, tol, "out")
    _check(grads[0], ref[0], tol, "grad_value")
    _check(grads[1], ref[1], tol, "grad_loc", keep=_loc_keep(w))
    _check(grads[2], ref[2], tol, "grad_attn")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("dist", ["box", "trained", "uniform"])
@pytest.mark.parametrize("K", [4, 2])
def test_c2_encoder_full_size_elementwise(K, dist, dtype):
    from boxer_b200 import workloads as W
    _box_case(W.coco_encoder(K=K, dist=dist, oob=0.02 if dist != "box" else 0.0, device=DEV), dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("dist,K", [("box", 4), ("trained", 4), ("box", 2)])
def test_c2_encoder_full_size_tile_kernels(dist, K, dtype):
    """The same full-size comparison through the query-tile x value-tile kernels (boxattn_tile.cuh): staged windows
    (box), windows that left their tile's halo and wide footprints (trained-like)."""
    from boxer_b200 import workloads as W
    b = _ops()
    b.ops.set_kernel_path("tile")
    try:
        _box_case(W.coco_encoder(K=K, dist=dist, oob=0.02 if dist != "box" else 0.0, device=DEV), dtype)
    finally:
        b.ops.set_kernel_path("auto")


def test_c2_encoder_full_size_batch2_deterministic():
    """B = 2 images per GPU (BASELINE configs[2]'s per-GPU batch) through the deterministic scatter, element-wise."""
    from boxer_b200 import workloads as W
    w = W.coco_encoder(B=2, K=2, dist="box", device=DEV)
    go = torch.randn(2, w.loc.shape[1], 256, device=DEV)
    out, grads = _run_box(_wl_inputs(w), torch.float32, go, deterministic=True)
    ref_out, ref = _oracle_box_per_image(w, go)
    _check(out, ref_out, 1e-4, "out")
    _check(grads[0], ref[0], 1e-4, "grad_value (deterministic)")
    _check(grads[2], ref[2], 1e-4, "grad_attn")
    b = _ops()
    b.ops.set_kernel_path("tile")       # and the tile kernels' deterministic scatter, B = 2
    try:
        out_t, grads_t = _run_box(_wl_inputs(w), torch.float32, go, deterministic=True)
        grads_t2 = _run_box(_wl_inputs(w), torch.float32, go, deterministic=True)[1]
    finally:
        b.ops.set_kernel_path("auto")
    _check(out_t, ref_out, 1e-4, "out (tile)")
    _check(grads_t[0], ref[0], 1e-4, "grad_value (tile, deterministic)")
    assert torch.equal(grads_t[0], grads_t2[0]), "deterministic scatter must be bit-reproducible"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("K", [28, 14])
def test_c4_mask_head_full_size(K, dtype):
    """BASELINE configs[3]: 300 decoder queries x KxK RoI grid over the four levels of a 1333x800 image."""
    from boxer_b200 import workloads as W
    from oracle import kernel_ref
    w = W.coco_mask_head(Nq=300, K=K, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(1235)
    go = torch.randn(1, 300, 256, device=DEV, generator=gen)
    gm = torch.randn(1, 300, K, K, 256, device=DEV, generator=gen)
    if dtype == torch.bfloat16:
        w.value, go, gm = _bf16_round(w.value), _bf16_round(go), _bf16_round(gm)
    out, mask, grads = _run_inst(_wl_inputs(w), dtype, go, gm)
    cpu = w.to("cpu", torch.float64)
    args = (cpu.value, cpu.shapes, cpu.level_start, cpu.loc, cpu.weights[0], cpu.weights[1])
    ro, rm = kernel_ref.instance_attn_forward(*args)
    rg = kernel_ref.instance_attn_backward(*args, go.double().cpu(), gm.double().cpu())
    tol = TOL[dtype]
    assert mask.shape == (1, 300, K, K, 256)
    _check(out, ro, tol, "out")
    _check(mask, rm, tol, "mask_out")
    _check(grads[0], rg[0], tol, "grad_value")
    _check(grads[1], rg[1], tol, "grad_loc", keep=_loc_keep(w))
    _check(grads[2], rg[2], tol, "grad_spatial_w")
    _check(grads[3], rg[3], tol, "grad_level_w")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_c5_bev_full_size(dtype):
    """BASELINE configs[4]: 468x468 BEV map, C = 128 (D = 16), Nq = 1000, 3x3 rotated grid, 8 frames."""
    from boxer_b200 import workloads as W
    _box_case(W.bev_rotated(B=8, Nq=1000, K=3, size=468, device=DEV), dtype)


@pytest.mark.parametrize("path", ["auto", "tile"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_c5_box3d_encoder_reference_exact(dtype, path):
    """The call BoxeR-3D's encoder layers actually make (box3d_transformer.py:233, base_boxer3d_detection.yaml:132-146):
    234^2 + 117^2 BEV levels, D = 32, Nq = S = 68 445, 2x2 grid with the /2 divisor, per-head reference angles.
    path "tile": the same through the query-tile x value-tile kernels (two levels, rotated 2 x 2 grids)."""
    from boxer_b200 import workloads as W
    b = _ops()
    w = W.box3d_encoder(device=DEV)
    assert w.dims == dict(B=1, S=68445, H=8, D=32, L=2, Nq=68445, P=4)
    b.ops.set_kernel_path(path)
    try:
        _box_case(w, dtype)
    finally:
        b.ops.set_kernel_path("auto")
