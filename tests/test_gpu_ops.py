"""GPU parity tests proper: the CUDA path (through the C ABI) vs the oracle.

Bars: fp32 <= 1e-4, bf16 <= 1e-2 (max-norm relative and, for the
reference-scale inputs, absolute), fp64 ~ round-off.  The reference's own tolerance is the much
looser rtol=1e-2, atol=1e-3 (tests/box_attn_test.py:87).
"""
import numpy as np
import pytest
import torch

from tests import helpers, refinputs

pytestmark = pytest.mark.gpu

DEV = "cuda"
TOL = {torch.float64: 1e-10, torch.float32: 1e-4, torch.bfloat16: 1e-2}


def _ops():
    import boxer_b200
    return boxer_b200


def _cuda(inp, dtype=None):
    out = {}
    for k, v in inp.items():
        if torch.is_tensor(v):
            if v.is_floating_point() and dtype is not None:
                v = v.to(dtype)
            v = v.to(DEV).contiguous()
        out[k] = v
    return out


def _run_box(inp, dtype, grad_out=None, deterministic=False):
    """inputs (any float dtype, CPU) -> our op in `dtype` (value) on the GPU."""
    b = _ops()
    tw = torch.float64 if dtype == torch.float64 else torch.float32
    g = _cuda(inp)
    # fresh leaves every call (.to() is a no-op for a matching dtype and .grad would accumulate)
    value = g["value"].detach().to(dtype).clone().requires_grad_(grad_out is not None)
    loc = g["loc"].detach().to(tw).clone().requires_grad_(grad_out is not None)
    attn = g["attn"].detach().to(tw).clone().requires_grad_(grad_out is not None)
    b.set_deterministic(deterministic)
    try:
        out = b.BoxAttnFunction.apply(value, g["shapes"], g["level_start"], loc, attn, 64)
        if grad_out is None:
            return out.detach(), None
        out.backward(torch.as_tensor(grad_out).to(DEV, dtype))
    finally:
        b.set_deterministic(None)
    return out.detach(), (value.grad, loc.grad, attn.grad)


def _run_inst(inp, dtype, grad_out=None, grad_mask=None, deterministic=False):
    b = _ops()
    tw = torch.float64 if dtype == torch.float64 else torch.float32
    need = grad_out is not None
    g = _cuda(inp)
    value = g["value"].detach().to(dtype).clone().requires_grad_(need)
    loc = g["loc"].detach().to(tw).clone().requires_grad_(need)
    sw = g["spatial_w"].detach().to(tw).clone().requires_grad_(need)
    lw = g["level_w"].detach().to(tw).clone().requires_grad_(need)
    b.set_deterministic(deterministic)
    try:
        out, mask = b.InstanceAttnFunction.apply(value, g["shapes"], g["level_start"], loc, sw, lw, inp["mask_size"], 64)
        if not need:
            return out.detach(), mask.detach(), None
        torch.autograd.backward([out, mask], [torch.as_tensor(grad_out).to(DEV, dtype),
                                              torch.as_tensor(grad_mask).to(DEV, dtype)])
    finally:
        b.set_deterministic(None)
    return out.detach(), mask.detach(), (value.grad, loc.grad, sw.grad, lw.grad)


def _near_cell_boundary(loc, shapes, eps=2e-3):
    """(B,Nq,H,L,P) mask of sample points within `eps` pixels of a pixel-grid line.  The bilinear
    interpolant is continuous there but its derivative w.r.t. the location is not: an fp32 kernel
    and an fp64 oracle can legitimately pick different cells (the fp32 rounding of loc*size-0.5 is
    ~1.5e-5 px at x~167), which changes grad_loc of that one point by O(1).  Such points are left out
    of the grad_loc comparison only; every other output is continuous and is compared everywhere."""
    loc = loc.detach().double().cpu()
    sh = torch.as_tensor(shapes).cpu()
    near = torch.zeros(loc.shape[:-1], dtype=torch.bool)
    for l in range(sh.shape[0]):
        for c, size in ((0, float(sh[l, 1])), (1, float(sh[l, 0]))):
            x = loc[:, :, :, l, :, c] * size - 0.5
            f = x - torch.floor(x)
            near[:, :, :, l] |= (f < eps) | (f > 1 - eps)
    return near


def _close_grad_loc(got, want, loc, shapes, tol, what="grad_loc"):
    keep = (~_near_cell_boundary(loc, shapes))[..., None]
    got = torch.as_tensor(got).detach().double().cpu().view(*keep.shape[:-1], 2) * keep
    want = torch.as_tensor(want).double().cpu().view(*keep.shape[:-1], 2) * keep
    assert float(keep.double().mean()) > 0.9
    _close(got, want, tol, what)


def _close(got, want, tol, what):
    want = torch.as_tensor(want)
    err = helpers.rel_err(got, want)
    assert err <= tol, f"{what}: max-norm relative error {err:.3e} > {tol:g}"


# ============================================================== golden fixtures (reference outputs)
def _gold_tol(inp, dtype):
    # the reference's check_forward("float") case was itself computed in fp32
    return max(TOL[dtype], 2e-6) if inp["value"].dtype == torch.float32 else TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32, torch.bfloat16], ids=["f64", "f32", "bf16"])
@pytest.mark.parametrize("case", list(helpers.box_inputs()))
def test_box_golden(case, dtype):
    inp, gold = helpers.box_inputs()[case], helpers.golden("box_attn_golden")[case]
    go = gold.get("grad_out")
    out, grads = _run_box(inp, dtype, go)
    tol = _gold_tol(inp, dtype)
    assert out.dtype == dtype and out.shape == gold["out"].shape
    _close(out, gold["out"], tol, "out")
    if dtype == torch.float32:   # the reference-scale inputs: absolute bar as BASELINE.md states it
        assert helpers.max_err(out, gold["out"]) <= 1e-4
    if grads is not None:
        gv, gl, ga = grads
        assert gv.dtype == dtype and gl.shape == inp["loc"].shape and ga.shape == inp["attn"].shape
        _close(helpers.slim_like(gv, gold["grad_value"]), gold["grad_value"], tol, "grad_value")
        _close(gl, gold["grad_loc"], tol, "grad_loc")
        _close(ga, gold["grad_attn"], tol, "grad_attn")


def _inst_grads_in(inp, gold):
    if "grad_out" not in gold:
        return None, None
    if "grad_mask" in gold:
        return gold["grad_out"], gold["grad_mask"]
    D = inp["value"].shape[-1]
    return gold["grad_out"], refinputs.side_rand((1, refinputs.LQ, 2, 2, refinputs.M * D), 3000 + D, torch.float64)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32, torch.bfloat16], ids=["f64", "f32", "bf16"])
@pytest.mark.parametrize("case", list(helpers.instance_inputs()))
def test_instance_golden(case, dtype):
    inp, gold = helpers.instance_inputs()[case], helpers.golden("instance_attn_golden")[case]
    go, gm = _inst_grads_in(inp, gold)
    out, mask, grads = _run_inst(inp, dtype, go, gm)
    tol = _gold_tol(inp, dtype)
    assert mask.shape == gold["mask_out"].shape
    _close(out, gold["out"], tol, "out")
    _close(mask, gold["mask_out"], tol, "mask_out")
    if grads is not None:
        gv, gl, gs, gw = grads
        _close(helpers.slim_like(gv, gold["grad_value"]), gold["grad_value"], tol, "grad_value")
        _close(gl, gold["grad_loc"], tol, "grad_loc")
        _close(gs, gold["grad_spatial_w"], tol, "grad_spatial_w")
        _close(gw, gold["grad_level_w"], tol, "grad_level_w")


# ============================================================== the reference's own test protocol
@pytest.mark.parametrize("D", refinputs.GRADCHECK_D)
def test_box_gradcheck_like_reference(D):
    """tests/box_attn_test.py:162-189: fp64 gradcheck over D in {30,...,3096} (every backward path)."""
    b = _ops()
    inp = _cuda(helpers.box_inputs()[f"gradcheck_D{D}"])
    args = (inp["value"].requires_grad_(True), inp["shapes"], inp["level_start"],
            inp["loc"].requires_grad_(True), inp["attn"].requires_grad_(True), 2)
    # default scatter = floating-point atomics (as in the reference): tolerate round-off-level
    # run-to-run differences in grad_value; the deterministic scatter must be exactly re-entrant
    assert torch.autograd.gradcheck(b.BoxAttnFunction.apply, args, fast_mode=D > 128, nondet_tol=1e-12)
    b.set_deterministic(True)
    try:
        assert torch.autograd.gradcheck(b.BoxAttnFunction.apply, args, fast_mode=True, nondet_tol=0.0)
    finally:
        b.set_deterministic(None)


@pytest.mark.parametrize("D", refinputs.GRADCHECK_D)
def test_instance_gradcheck_like_reference(D):
    """tests/instance_attn_test.py:255-292 (5-D weights, mask_size 2)."""
    b = _ops()
    inp = _cuda(helpers.instance_inputs()[f"gradcheck_D{D}"])
    args = (inp["value"].requires_grad_(True), inp["shapes"], inp["level_start"], inp["loc"].requires_grad_(True),
            inp["spatial_w"].requires_grad_(True), inp["level_w"].requires_grad_(True), 2, 2)
    assert torch.autograd.gradcheck(b.InstanceAttnFunction.apply, args, fast_mode=D > 128, nondet_tol=1e-12)


def test_reference_allclose_protocol():
    """check_forward / check_forward_and_backward of tests/box_attn_test.py with the reference's own
    (loose) criterion, on the inputs that script draws."""
    gold = helpers.golden("box_attn_golden")
    for case, dtype in (("fwd_float", torch.float32), ("fwd_double", torch.float64), ("fwdbwd_double", torch.float64)):
        inp = helpers.box_inputs()[case]
        go = gold[case].get("grad_out")
        out, grads = _run_box(inp, dtype, go)
        assert torch.allclose(out.cpu().double(), torch.from_numpy(gold[case]["out"]).double(), rtol=1e-2, atol=1e-3)
        if grads is not None:
            for g, k in zip(grads, ("grad_value", "grad_loc", "grad_attn")):
                assert torch.allclose(g.cpu(), torch.from_numpy(gold[case][k]), rtol=1e-2, atol=1e-3)


# ============================================================== larger seeded cases vs the C oracle
def _oracle_box(w, grad_out):
    from oracle import kernel_ref
    cpu = w.to("cpu", torch.float64)
    attn = cpu.weights[0]
    out = kernel_ref.box_attn_forward(cpu.value, cpu.shapes, cpu.level_start, cpu.loc, attn)
    grads = kernel_ref.box_attn_backward(cpu.value, cpu.shapes, cpu.level_start, cpu.loc, attn, grad_out.double().cpu())
    return out, grads


def _wl_inputs(w):
    d = {"value": w.value, "loc": w.loc, "shapes": w.shapes, "level_start": w.level_start}
    if w.instance:
        d.update(spatial_w=w.weights[0], level_w=w.weights[1], mask_size=w.kernel_size)
    else:
        d["attn"] = w.weights[0]
    return d


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("K,oob", [(2, 0.0), (4, 0.05), (3, 0.3)])
def test_box_c1_scale_vs_oracle(K, oob, dtype):
    """BASELINE configs[0] scale: B=1, one 64x64 level, C=256, Nq=300, 8 heads."""
    from boxer_b200 import workloads as W
    dev = torch.device(DEV)
    gen = torch.Generator(device=dev).manual_seed(100 + K)
    sh, start = W._level_meta([(64, 64)], dev)
    value = torch.randn(1, 64 * 64, 8, 32, device=dev, generator=gen)
    if dtype == torch.bfloat16:
        value = value.bfloat16().float()          # the oracle sees the same (rounded) values
    loc = W._apply_oob(torch.rand(1, 300, 8, 1, K * K, 2, device=dev, generator=gen), oob, gen)
    attn = W._softmax_weights(1, 300, 8, 1, K, gen, dev)
    w = W.Workload("c1", value, sh, start, loc, (attn,), K)
    go = torch.randn(1, 300, 256, device=dev, generator=gen)
    if dtype == torch.bfloat16:
        go = go.bfloat16().float()
    out, grads = _run_box(_wl_inputs(w), dtype, go)
    ref_out, ref_grads = _oracle_box(w, go)
    tol = TOL[dtype]
    _close(out, ref_out, tol, "out")
    _compare_box_grads(grads, ref_grads, w, tol)


def _compare_box_grads(grads, ref_grads, w, tol, tag=""):
    _close(grads[0], ref_grads[0].view_as(grads[0]), tol, "grad_value" + tag)
    _close_grad_loc(grads[1], ref_grads[1], w.loc, w.shapes, tol, "grad_loc" + tag)
    _close(grads[2], ref_grads[2].view_as(grads[2]), tol, "grad_attn" + tag)


@pytest.mark.parametrize("maker,kw", [
    ("coco_decoder", dict(Nq=300, K=2)),          # rows << SMs: split path
    ("coco_decoder", dict(Nq=37, K=4, B=2)),
    ("bev_rotated", dict(Nq=100, K=3, size=96)),  # D=16 (G=4), rotated grid, one level
])
def test_box_decoder_like_vs_oracle(maker, kw):
    from boxer_b200 import workloads as W
    w = getattr(W, maker)(device=DEV, **kw)
    B, Nq = w.loc.shape[:2]
    C = w.value.shape[2] * w.value.shape[3]
    go = torch.randn(B, Nq, C, device=DEV)
    out, grads = _run_box(_wl_inputs(w), torch.float32, go)
    ref_out, ref_grads = _oracle_box(w, go)
    _close(out, ref_out, 1e-4, "out")
    _compare_box_grads(grads, ref_grads, w, 1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("K,Nq", [(14, 20), (4, 300), (2, 5)])
def test_instance_mask_head_vs_oracle(K, Nq, dtype):
    """configs[3] shape family (reference K=14): point-split CTAs, per-point mask rows."""
    from boxer_b200 import workloads as W
    from oracle import kernel_ref
    w = W.coco_mask_head(Nq=Nq, K=K, device=DEV, image=(200, 336))
    if dtype == torch.bfloat16:
        w.value = w.value.bfloat16().float()
    C = 256
    go = torch.randn(1, Nq, C, device=DEV)
    gm = torch.randn(1, Nq, K, K, C, device=DEV)
    if dtype == torch.bfloat16:
        go, gm = go.bfloat16().float(), gm.bfloat16().float()
    out, mask, grads = _run_inst(_wl_inputs(w), dtype, go, gm)
    cpu = w.to("cpu", torch.float64)
    args = (cpu.value, cpu.shapes, cpu.level_start, cpu.loc, cpu.weights[0], cpu.weights[1])
    ro, rm = kernel_ref.instance_attn_forward(*args)
    rg = kernel_ref.instance_attn_backward(*args, go.double().cpu(), gm.double().cpu())
    tol = TOL[dtype]
    _close(out, ro, tol, "out")
    _close(mask, rm.view_as(mask), tol, "mask_out")
    for g, r, k in zip(grads, rg, ("grad_value", "grad_loc", "grad_spatial_w", "grad_level_w")):
        if k == "grad_loc":
            _close_grad_loc(g, r, w.loc, w.shapes, tol)
        else:
            _close(g, r.view_as(g), tol, k)


# ============================================================== footprint-window kernels, forced at small sizes
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("path", ["window", "window-staged", "tile"])
@pytest.mark.parametrize("case", ["enc_box_K4", "enc_box_K2", "enc_uniform_K4", "enc_box_K3_oob", "dec_K4", "bev_K3", "enc_box_K5"])
def test_window_kernels_vs_oracle(case, dtype, path):
    """Same op through the footprint-window kernels (forced with set_kernel_path) on small inputs:
    window mode (box-structured), per-point fallback (uniform / wide boxes), borders and padding.
    "tile": the query-tile x value-tile kernels (TMA-staged value halos; they apply to the encoder-shaped cases with
    head_dim 32 and fall through to the default kernels elsewhere)."""
    from boxer_b200 import workloads as W
    b = _ops()
    img = (72, 100)
    w = {
        "enc_box_K4": lambda: W.coco_encoder(K=4, dist="box", image=img, device=DEV, oob=0.03),
        "enc_box_K2": lambda: W.coco_encoder(K=2, dist="box", image=img, device=DEV),
        "enc_uniform_K4": lambda: W.coco_encoder(K=4, dist="uniform", image=img, device=DEV, oob=0.1),
        "enc_box_K3_oob": lambda: W.coco_encoder(K=3, dist="box", image=img, device=DEV, oob=0.3, B=2),
        "dec_K4": lambda: W.coco_decoder(Nq=64, K=4, image=img, device=DEV),
        "bev_K3": lambda: W.bev_rotated(Nq=128, K=3, size=40, device=DEV),
        "enc_box_K5": lambda: W.coco_encoder(K=5, dist="box", image=img, device=DEV, heads=4, head_dim=64),
    }[case]()
    if dtype == torch.bfloat16:
        w.value = w.value.bfloat16().float()
    B, Nq = w.loc.shape[:2]
    C = w.value.shape[2] * w.value.shape[3]
    go = torch.randn(B, Nq, C, device=DEV)
    if dtype == torch.bfloat16:
        go = go.bfloat16().float()
    b.ops.set_kernel_path(path)
    try:
        out, grads = _run_box(_wl_inputs(w), dtype, go)
        det = _run_box(_wl_inputs(w), dtype, go, deterministic=True)[1]
    finally:
        b.ops.set_kernel_path("auto")
    ref_out, ref_grads = _oracle_box(w, go)
    tol = TOL[dtype]
    _close(out, ref_out, tol, "out")
    _compare_box_grads(grads, ref_grads, w, tol)
    _compare_box_grads(det, ref_grads, w, tol, " (deterministic)")
    # and the point kernels agree with the window kernels
    b.ops.set_kernel_path("point")
    try:
        out_p, grads_p = _run_box(_wl_inputs(w), dtype, go)
    finally:
        b.ops.set_kernel_path("auto")
    _close(out_p, out.float(), 2 * tol, "point vs window out")


# ============================================================== full size: slices + size-independent properties
@pytest.mark.parametrize("dist", ["box", "uniform"])
def test_coco_encoder_full_size(dist):
    """BASELINE configs[1] (B=1, 4 levels of 800x1333, Nq=S=22223, 8 heads, K=4, fp32).
    Forward / grad_loc / grad_attn are per-query, so the oracle checks a random slice of queries;
    grad_value is checked through the adjoint identity <out(value), go> == <value, grad_value>."""
    from boxer_b200 import workloads as W
    from oracle import kernel_ref
    b = _ops()
    w = W.coco_encoder(K=4, dist=dist, oob=0.02, device=DEV)
    S = w.value.shape[1]
    go = torch.randn(1, S, 256, device=DEV)
    value = w.value.clone().requires_grad_(True)
    loc = w.loc.clone().requires_grad_(True)
    attn = w.weights[0].clone().requires_grad_(True)
    out = b.BoxAttnFunction.apply(value, w.shapes, w.level_start, loc, attn, 64)
    out.backward(go)

    idx = torch.randperm(S, device=DEV)[:257].sort().values
    cpu = lambda t: t.detach().double().cpu()
    sl = (cpu(w.value), w.shapes.cpu(), w.level_start.cpu(), cpu(w.loc[:, idx]), cpu(w.weights[0][:, idx]))
    ref_out = kernel_ref.box_attn_forward(*sl)
    _, ref_gl, ref_ga = kernel_ref.box_attn_backward(*sl, cpu(go[:, idx]))
    _close(out[:, idx], ref_out, 1e-4, "out[slice]")
    _close_grad_loc(loc.grad[:, idx], ref_gl, w.loc[:, idx], w.shapes, 1e-4, "grad_loc[slice]")
    _close(attn.grad[:, idx], ref_ga.view_as(attn.grad[:, idx]), 1e-4, "grad_attn[slice]")

    lhs = (out.detach().double() * go.double()).sum()
    rhs = (w.value.double() * value.grad.double()).sum().to(lhs.device)
    assert abs(float(lhs - rhs)) <= 1e-4 * max(1.0, abs(float(lhs))), (float(lhs), float(rhs))

    # linearity in value and in the weights (size-independent properties of the op)
    out2 = b.BoxAttnFunction.apply(2.5 * w.value, w.shapes, w.level_start, w.loc, 0.5 * w.weights[0], 64)
    _close(out2, 1.25 * out.detach(), 1e-5, "linearity")


def test_deterministic_flag_bit_reproducible():
    from boxer_b200 import workloads as W
    w = W.coco_encoder(K=2, dist="uniform", device=DEV)
    go = torch.randn(1, w.value.shape[1], 256, device=DEV)
    inp = _wl_inputs(w)
    runs = [_run_box(inp, torch.float32, go, deterministic=True)[1] for _ in range(3)]
    for g0, g in zip(runs[0], runs[1]):
        assert torch.equal(g0, g)
    for g0, g in zip(runs[0], runs[2]):
        assert torch.equal(g0, g)
    atomic = _run_box(inp, torch.float32, go, deterministic=False)[1]
    _close(runs[0][0], atomic[0], 1e-5, "deterministic vs atomic grad_value")
    assert torch.equal(runs[0][1], atomic[1]) and torch.equal(runs[0][2], atomic[2])


def test_deterministic_instance_and_bf16():
    from boxer_b200 import workloads as W
    w = W.coco_mask_head(Nq=40, K=4, device=DEV, image=(200, 336))
    go = torch.randn(1, 40, 256, device=DEV)
    gm = torch.randn(1, 40, 4, 4, 256, device=DEV)
    for dtype in (torch.float32, torch.bfloat16, torch.float64):
        a = _run_inst(_wl_inputs(w), dtype, go, gm, deterministic=True)[2]
        b_ = _run_inst(_wl_inputs(w), dtype, go, gm, deterministic=True)[2]
        c = _run_inst(_wl_inputs(w), dtype, go, gm, deterministic=False)[2]
        assert all(torch.equal(x, y) for x, y in zip(a, b_))
        _close(a[0], c[0], 2e-2 if dtype == torch.bfloat16 else 1e-5, f"det vs atomic {dtype}")


# ============================================================== edge cases
def _tiny(B=1, Nq=3, H=2, D=32, shapes=((5, 4), (3, 2)), K=2, dtype=torch.float32):
    sh = torch.tensor(shapes, dtype=torch.long, device=DEV)
    start = torch.cat((sh.new_zeros(1), sh.prod(1).cumsum(0)[:-1]))
    S = int(sh.prod(1).sum())
    L = len(shapes)
    value = torch.randn(B, S, H, D, device=DEV, dtype=dtype)
    loc = torch.rand(B, Nq, H, L, K * K, 2, device=DEV, dtype=dtype)
    attn = torch.rand(B, Nq, H, L, K * K, device=DEV, dtype=dtype)
    return value, sh, start, loc, attn


def test_empty_queries_and_batch():
    b = _ops()
    value, sh, start, loc, attn = _tiny(Nq=0)
    out = b.ops.box_attn_forward(value, sh, start, loc, attn, 64)
    assert out.shape == (1, 0, 64)
    gv, gl, ga = b.ops.box_attn_backward(value, sh, start, loc, attn, out, 64)
    assert gv.shape == value.shape and float(gv.abs().sum()) == 0.0 and gl.numel() == 0
    value, sh, start, loc, attn = _tiny(B=0)
    assert b.ops.box_attn_forward(value, sh, start, loc, attn, 64).shape == (0, 3, 64)
    value, sh, start, loc, attn = _tiny(Nq=0, K=2)
    o, m = b.ops.instance_attn_forward(value, sh, start, loc, attn, attn.clone(), 64)
    assert o.shape == (1, 0, 64) and m.shape == (1, 0, 4, 64)


def test_all_out_of_range_and_nonfinite_locations():
    b = _ops()
    value, sh, start, loc, attn = _tiny()
    far = loc + 7.0
    out = b.ops.box_attn_forward(value, sh, start, far, attn, 64)
    assert float(out.abs().max()) == 0.0
    gv, gl, ga = b.ops.box_attn_backward(value, sh, start, far, attn, torch.ones_like(out), 64)
    assert float(gv.abs().max()) == 0.0 and float(gl.abs().max()) == 0.0 and float(ga.abs().max()) == 0.0
    # NaN / inf locations fail the window test (comparisons are false) and contribute nothing
    bad = loc.clone()
    bad[0, 0, 0, 0, 0, 0] = float("nan")
    bad[0, 1, 1, 1, 1, 1] = float("inf")
    bad[0, 2, 0, 0, 2, 0] = -float("inf")
    good = loc.clone()
    for ix in ((0, 0, 0, 0, 0), (0, 1, 1, 1, 1), (0, 2, 0, 0, 2)):
        good[ix] = 9.0   # plainly outside
    o1 = b.ops.box_attn_forward(value, sh, start, bad, attn, 64)
    o2 = b.ops.box_attn_forward(value, sh, start, good, attn, 64)
    assert torch.equal(o1, o2) and bool(torch.isfinite(o1).all())


def test_nonfinite_and_zero_weights_window_vs_point_paths():
    """The window kernels accumulate pixel weights in fixed point; non-finite weights must still
    propagate (they take the float path) and all-zero weights must give exact zeros, like the point kernels."""
    from boxer_b200 import workloads as W
    b = _ops()
    w = W.coco_encoder(K=4, image=(72, 100), device=DEV)
    attn = w.weights[0].clone()
    # queries 58..62 = pixels (4, 6..10) of the 9 x 13 level 0: every sample point of theirs is inside
    attn[0, 58, 1, 0, 1, 2] = float("nan")
    attn[0, 59, 2, 1, 0, 0] = float("inf")
    attn[0, 60] = 0.0                      # a query whose weights are all exactly zero
    attn[0, 61, 4, 2] = 0.0                # one (row, level) with zero weights
    attn[0, 62, 0, 0] *= -1.0              # negative weights are legal inputs
    go = torch.randn(1, w.value.shape[1], 256, device=DEV)
    res = {}
    for path in ("window", "window-staged", "point"):
        b.ops.set_kernel_path(path)
        try:
            out = b.ops.box_attn_forward(w.value, w.shapes, w.level_start, w.loc, attn, 64)
            gv, gl, ga = b.ops.box_attn_backward(w.value, w.shapes, w.level_start, w.loc, attn, go, 64)
        finally:
            b.ops.set_kernel_path("auto")
        res[path] = (out, gl, ga)
        assert bool(torch.isnan(out[0, 58, 32:64]).all()) and bool(torch.isfinite(out[0, 58, :32]).all())
        assert not bool(torch.isfinite(out[0, 59, 64:96]).any())
        assert float(out[0, 60].abs().max()) == 0.0
        assert float(gl[0, 60].abs().max()) == 0.0            # d out / d loc carries the weight as a factor
        assert float(ga[0, 60].abs().max()) > 0.0             # d out / d attn does not
    ow, op_ = res["window"][0], res["point"][0]
    fin = torch.isfinite(op_)
    assert torch.equal(torch.isfinite(ow), fin)
    _close(ow[fin], op_[fin], 2e-5, "window vs point out")
    gaw, gap = res["window"][2], res["point"][2]
    fin = torch.isfinite(gap) & torch.isfinite(gaw)
    _close(gaw[fin], gap[fin], 2e-5, "window vs point grad_attn")


def test_border_semantics_match_grid_sample():
    """Half-pixel border: loc in (-0.5/W, 0) still gets weight from pixel 0 (window test is -1 < x)."""
    from oracle import plain
    b = _ops()
    value, sh, start, _, _ = _tiny(Nq=1, H=1, D=4, shapes=((3, 3),), K=1, dtype=torch.float64)
    xs = torch.tensor([-0.4, -0.2, -0.01, 0.0, 1 / 6, 0.5, 5 / 6, 0.999, 1.0, 1.1, 1.3, 4 / 3 - 1e-9],
                      device=DEV, dtype=torch.float64)
    for x in xs:
        for y in xs:
            loc = torch.stack([x, y]).view(1, 1, 1, 1, 1, 2).contiguous()
            attn = torch.ones(1, 1, 1, 1, 1, device=DEV, dtype=torch.float64)
            got = b.ops.box_attn_forward(value, sh, start, loc, attn, 64).cpu()
            want = plain.plain_box_attn(value.cpu().view(1, 9, 4), sh.cpu(), 2 * loc.cpu() - 1, attn.cpu())
            assert helpers.max_err(got, want) <= 1e-12, (float(x), float(y))


def test_misaligned_and_odd_shapes_take_generic_path():
    from oracle import kernel_ref
    b = _ops()
    # head_dim 30 / 71: not a multiple of the vector width; P = 1 and a single 1x1 level too
    for D, shapes, K in ((30, ((4, 3),), 1), (71, ((1, 1), (2, 2)), 2), (4, ((3, 5),), 3)):
        value, sh, start, loc, attn = _tiny(D=D, shapes=shapes, K=K)
        out = b.ops.box_attn_forward(value, sh, start, loc, attn, 64)
        ref = kernel_ref.box_attn_forward(value.double().cpu(), sh.cpu(), start.cpu(), loc.double().cpu(), attn.double().cpu())
        _close(out, ref, 1e-5, f"D={D}")
    # a value tensor whose storage is only 4-byte aligned must not take the float4 path
    value, sh, start, loc, attn = _tiny()
    buf = torch.empty(value.numel() + 1, device=DEV)
    v2 = buf[1:].view_as(value).copy_(value)
    assert v2.data_ptr() % 16 != 0 and v2.is_contiguous()
    assert torch.allclose(b.ops.box_attn_forward(v2, sh, start, loc, attn, 64),
                          b.ops.box_attn_forward(value, sh, start, loc, attn, 64), atol=1e-6)


def test_six_d_weights_and_batch_gt_one():
    b = _ops()
    value, sh, start, loc, attn = _tiny(B=4, Nq=6)
    six = attn.view(4, 6, 2, 2, 2, 2).clone().requires_grad_(True)
    out = b.BoxAttnFunction.apply(value, sh, start, loc, six, 2)
    out.sum().backward()
    assert six.grad.shape == six.shape
    per_image = torch.cat([b.ops.box_attn_forward(value[i:i + 1].contiguous(), sh, start, loc[i:i + 1].contiguous(),
                                                   attn[i:i + 1].contiguous(), 64) for i in range(4)])
    assert torch.equal(out.detach(), per_image)     # images are independent: no cross-image traffic


def test_error_behaviour():
    b = _ops()
    value, sh, start, loc, attn = _tiny(B=3)
    with pytest.raises(RuntimeError, match="must divide im2col_step"):
        b.ops.box_attn_forward(value, sh, start, loc, attn, 2)          # box_attn.cu:40-42
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        b.ops.box_attn_forward(value.cpu(), sh, start, loc, attn, 64)   # box_attn.h:53 / CHECK_CUDA
    with pytest.raises(RuntimeError, match="must be contiguous"):
        b.ops.box_attn_forward(value.transpose(2, 3), sh, start, loc, attn, 64)
    with pytest.raises(RuntimeError, match="sampling_loc must be"):
        b.ops.box_attn_forward(value, sh, start, loc[:, :, :1].contiguous(), attn, 64)
    with pytest.raises(RuntimeError, match="float32, float64 and bfloat16"):
        b.ops.box_attn_forward(value.half(), sh, start, loc, attn, 64)
    with pytest.raises(RuntimeError, match="int64"):
        b.ops.box_attn_forward(value, sh.int(), start, loc, attn, 64)


def test_amp_contract_and_bf16_opt_in():
    """custom_fwd(cast_inputs=float32): under autocast the op still runs (and returns) fp32."""
    b = _ops()
    value, sh, start, loc, attn = _tiny()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = b.BoxAttnFunction.apply(value.bfloat16(), sh, start, loc.bfloat16(), attn, 64)
    assert out.dtype == torch.float32
    v = value.clone().requires_grad_(True)
    out16 = b.BoxAttnBf16Function.apply(v, sh, start, loc, attn, 64)
    assert out16.dtype == torch.bfloat16
    out16.float().sum().backward()
    assert v.grad.dtype == torch.float32
    ref = b.BoxAttnFunction.apply(value, sh, start, loc, attn, 64)
    _close(out16, ref, 1e-2, "bf16 opt-in")


def test_runs_on_a_side_stream():
    b = _ops()
    value, sh, start, loc, attn = _tiny(Nq=64)
    ref = b.ops.box_attn_forward(value, sh, start, loc, attn, 64)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out = b.ops.box_attn_forward(value, sh, start, loc, attn, 64)
    s.synchronize()
    assert torch.equal(out, ref)


# ============================================================== modules on the GPU vs reference modules
@pytest.mark.parametrize("case", list(refinputs.module_cases()))
def test_modules_match_reference_modules(case):
    import boxer_b200
    spec = refinputs.module_cases()[case]
    gold = helpers.golden("modules_golden")[case]
    mod = getattr(boxer_b200, spec["cls"])(**spec["ctor"]).double()
    state = {k[len("param_"):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("param_")}
    mod.load_state_dict(state, strict=True)
    mod = mod.to(DEV)
    if spec["cls"] == "InstanceAttention":
        mod.inferencing = spec["inferencing"]
    args = [a.to(DEV) if torch.is_tensor(a) else a for a in refinputs.module_inputs(spec)]
    outs = mod(*args)
    flat = []
    for o in outs:
        if o is None:
            continue
        flat.extend(o if isinstance(o, tuple) else [o])
    assert len(flat) == int(gold["n_out"])
    for i, o in enumerate(flat):
        _close(o, gold[f"out{i}"], 1e-9, f"{case} out{i}")


# ============================================================== runtime properties
def test_cuda_graph_capture_and_replay():
    """The library only enqueues work on the stream it is given (no allocation, no sync), so a
    launch-bound decoder-sized forward+backward can be captured in a CUDA graph and replayed."""
    from boxer_b200 import workloads as W
    b = _ops()
    w = W.coco_decoder(Nq=300, K=2, image=(200, 336), device=DEV)
    go = torch.randn(1, 300, 256, device=DEV)
    args = (w.value, w.shapes, w.level_start, w.loc, w.weights[0].contiguous())
    ref_out = b.ops.box_attn_forward(*args, 64)                      # also warms the per-kernel occupancy cache
    ref_g = b.ops.box_attn_backward(*args, go, 64)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            out = b.ops.box_attn_forward(*args, 64)
            grads = b.ops.box_attn_backward(*args, go, 64)
    torch.cuda.current_stream().wait_stream(s)
    w.value.mul_(2.0)                                                # replay must see the new inputs
    g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(out, 2.0 * ref_out, rtol=1e-5, atol=1e-6)
    assert torch.allclose(grads[0], ref_g[0], rtol=1e-4, atol=1e-5)  # d/dvalue does not depend on value
    assert torch.allclose(grads[2], 2.0 * ref_g[2], rtol=1e-4, atol=1e-5)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_a_non_current_device():
    """One process, tensors on cuda:1 while cuda:0 is current: the wrapper switches device for the call
    (the reference has no device guard at all, box_attn.cu:56)."""
    from boxer_b200 import workloads as W
    b = _ops()
    w0 = W.coco_encoder(K=4, image=(72, 100), device="cuda:0")
    w1 = w0.to("cuda:1")
    assert torch.cuda.current_device() == 0
    out0 = b.ops.box_attn_forward(w0.value, w0.shapes, w0.level_start, w0.loc, w0.weights[0], 64)
    out1 = b.ops.box_attn_forward(w1.value, w1.shapes, w1.level_start, w1.loc, w1.weights[0], 64)
    assert out1.device.index == 1 and torch.cuda.current_device() == 0
    assert torch.equal(out0.cpu(), out1.cpu())
    go = torch.randn_like(out0)
    g0 = b.ops.box_attn_backward(w0.value, w0.shapes, w0.level_start, w0.loc, w0.weights[0], go, 64)
    g1 = b.ops.box_attn_backward(w1.value, w1.shapes, w1.level_start, w1.loc, w1.weights[0], go.to("cuda:1"), 64)
    assert torch.equal(g0[1].cpu(), g1[1].cpu()) and torch.equal(g0[2].cpu(), g1[2].cpu())
    with pytest.raises(RuntimeError, match="same device"):
        b.ops.box_attn_forward(w0.value, w1.shapes, w0.level_start, w0.loc, w0.weights[0], 64)
