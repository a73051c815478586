"""GPU: the fused box -> grid -> attention op vs the reference formulation.

Oracle = the reference's grid construction (BoxAttention._where_to_attend, box_attention.py:196-214;
Box3dAttention, :304-338 -- restated in oracle/plain.py:grid_from_boxes) followed by its grid_sample
oracle, differentiated by autograd in fp64.  Also: the three nn.Modules with set_fused_grid(True)
against outputs of the reference's own module classes (tests/golden/modules_golden.npz).
"""
import math

import pytest
import torch

from tests import helpers, refinputs
from tests.test_gpu_ops import _near_cell_boundary

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(name):
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    cfg = {
        # name: (B, shapes, H, D, Nq, K, rotation, valid_ratios, index divisor)
        "enc_like_K4": (1, [(23, 30), (12, 15), (6, 8), (3, 4)], 8, 32, 200, 4, False, False, 4),
        "k2_ratios": (2, [(20, 24), (10, 12), (5, 6)], 4, 32, 61, 2, False, True, 2),
        "rot_k3_1lvl": (2, [(40, 40)], 8, 16, 77, 3, True, False, 2),
        "rot_k2_ratios_d64": (1, [(16, 20), (8, 10)], 4, 64, 33, 2, True, True, 2),
        "odd_d30_k2": (1, [(9, 7), (5, 4)], 2, 30, 9, 2, True, True, 2),        # generic path: grid materialised inside
        "k6_p36": (1, [(30, 30), (15, 15)], 8, 32, 50, 6, False, False, 6),       # P = 36 > 4*G: falls back inside
        "neg_sizes": (1, [(12, 12)], 4, 32, 40, 2, False, False, 2),
    }[name]
    B, shapes, H, D, Nq, K, rot, ratios, div = cfg
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    sh = refinputs.shapes_tensor(shapes)
    start = refinputs.level_start_index(sh)
    value = torch.rand(B, S, H, D, generator=g, dtype=torch.float64) * 2 - 1
    boxes = torch.rand(B, Nq, H, L, 4, generator=g, dtype=torch.float64)
    boxes[..., :2] = 0.05 + 0.9 * boxes[..., :2]
    boxes[..., 2:] = 0.02 + 0.4 * boxes[..., 2:]
    if name == "neg_sizes":
        boxes[..., 2:] -= 0.15          # some widths / heights <= 0: relu and its zero gradient
    angles = (torch.rand(B, Nq, H, L, 1, generator=g, dtype=torch.float64) * 2 * math.pi) if rot else None
    vr = (0.6 + 0.4 * torch.rand(B, 1, 1, L, 1, 2, generator=g, dtype=torch.float64)) if ratios else None
    half = K / 2.0
    if K % 2 == 0:
        ticks = torch.linspace(-half + 0.5, half - 0.5, K, dtype=torch.float64)
    else:
        ticks = torch.linspace(-(K - 1) // 2, (K - 1) // 2, K, dtype=torch.float64)
    yy, xx = torch.meshgrid(ticks, ticks, indexing="ij")
    kidx = torch.stack([xx, yy], -1).reshape(-1, 2) / div
    attn = torch.softmax(torch.randn(B, Nq, H, L * K * K, generator=g, dtype=torch.float64), -1).view(B, Nq, H, L, K, K)
    go = torch.randn(B, Nq, H * D, generator=g, dtype=torch.float64)
    return dict(value=value, shapes=sh, start=start, boxes=boxes, angles=angles, vr=vr, kidx=kidx, attn=attn, go=go)


def _oracle(c):
    from oracle import plain
    value = c["value"].clone().requires_grad_(True)
    boxes = c["boxes"].clone().requires_grad_(True)
    angles = c["angles"].clone().requires_grad_(True) if c["angles"] is not None else None
    attn = c["attn"].clone().requires_grad_(True)
    grid = plain.grid_from_boxes(boxes, angles, c["vr"], c["kidx"])
    B, S = value.shape[:2]
    out = plain.plain_box_attn(value.view(B, S, -1), c["shapes"], 2 * grid - 1, attn)
    out.backward(c["go"])
    return out.detach(), grid.detach(), (value.grad, boxes.grad, angles.grad if angles is not None else None, attn.grad)


def _ours(c, dtype, deterministic=False):
    import boxer_b200
    tw = torch.float64 if dtype == torch.float64 else torch.float32
    mv = lambda t, dt: None if t is None else t.to(DEV, dt).contiguous()
    value = mv(c["value"], dtype).requires_grad_(True)
    boxes = mv(c["boxes"], tw).requires_grad_(True)
    angles = mv(c["angles"], tw)
    if angles is not None:
        angles.requires_grad_(True)
    attn = mv(c["attn"], tw).requires_grad_(True)
    boxer_b200.set_deterministic(deterministic)
    try:
        fn = boxer_b200.BoxGridAttnBf16Function if dtype == torch.bfloat16 else boxer_b200.BoxGridAttnFunction
        out = fn.apply(value, c["shapes"].to(DEV), c["start"].to(DEV), boxes, angles, mv(c["vr"], tw), mv(c["kidx"], tw), attn, 64)
        out.backward(mv(c["go"], out.dtype))
    finally:
        boxer_b200.set_deterministic(None)
    return out.detach(), (value.grad, boxes.grad, angles.grad if angles is not None else None, attn.grad)


CASES = ["enc_like_K4", "k2_ratios", "rot_k3_1lvl", "rot_k2_ratios_d64", "odd_d30_k2", "k6_p36", "neg_sizes"]


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4), (torch.bfloat16, 1e-2)],
                         ids=["f64", "f32", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_fused_matches_reference_formulation(name, dtype, tol):
    c = _case(name)
    if dtype == torch.bfloat16:
        c["value"] = c["value"].bfloat16().double()
        c["go"] = c["go"].bfloat16().double()
    ref_out, grid, ref_g = _oracle(c)
    out, g = _ours(c, dtype)
    assert helpers.rel_err(out, ref_out) <= tol, "out"
    assert helpers.rel_err(g[0], ref_g[0]) <= tol, "grad_value"
    assert helpers.rel_err(g[3], ref_g[3]) <= tol, "grad_attn"
    # box / angle gradients sum the location gradients of the level's points; a point that sits on a pixel
    # grid line has a discontinuous d/dloc (fp32 may pick the other cell): compare levels without such points
    if dtype == torch.float64:
        assert helpers.rel_err(g[1], ref_g[1]) <= tol, "grad_boxes"
        if ref_g[2] is not None:
            assert helpers.rel_err(g[2], ref_g[2]) <= tol, "grad_angles"
    else:
        keep = (~_near_cell_boundary(grid, c["shapes"]).any(-1))[..., None]      # (B,Nq,H,L,1)
        assert float(keep.double().mean()) > 0.8
        gb, rb = g[1].double().cpu() * keep, ref_g[1] * keep
        assert helpers.rel_err(gb, rb) <= tol, "grad_boxes"
        if ref_g[2] is not None:
            assert helpers.rel_err(g[2].double().cpu() * keep, ref_g[2] * keep) <= tol, "grad_angles"


def test_fused_equals_unfused_on_the_same_device():
    """grid materialised with torch + BoxAttnFunction vs the fused op: same kernels underneath (fp32)."""
    import boxer_b200
    from oracle import plain
    c = _case("enc_like_K4")
    out_f, g_f = _ours(c, torch.float32)
    dev = lambda t: None if t is None else t.to(DEV, torch.float32)
    value = dev(c["value"]).requires_grad_(True)
    boxes = dev(c["boxes"]).requires_grad_(True)
    attn = dev(c["attn"]).requires_grad_(True)
    grid = plain.grid_from_boxes(boxes, None, None, dev(c["kidx"])).contiguous()
    out = boxer_b200.BoxAttnFunction.apply(value, c["shapes"].to(DEV), c["start"].to(DEV), grid, attn, 64)
    out.backward(dev(c["go"]))
    assert helpers.rel_err(out_f, out.detach()) <= 2e-5
    assert helpers.rel_err(g_f[0], value.grad) <= 2e-5
    assert helpers.rel_err(g_f[3], attn.grad) <= 2e-5


def test_fused_accepts_odd_offset_views_of_the_small_operands():
    """kernel_indices / valid_ratios / boxes that are contiguous views at an odd element offset (4-byte aligned only) used to
    size the workspace for the fused kernels and then take the general path: BXR_ERR_WORKSPACE.  ops.py re-aligns them."""
    import boxer_b200

    def odd(t):         # same values, storage offset of one element
        buf = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
        v = buf[1:].view(t.shape)
        v.copy_(t)
        assert v.data_ptr() % 8 != 0 and v.is_contiguous()
        return v

    for name in ("k2_ratios", "rot_k2_ratios_d64"):
        c = _case(name)
        out_ref, g_ref = _ours(c, torch.float32)
        mv = lambda t: None if t is None else t.to(DEV, torch.float32).contiguous()
        value = mv(c["value"]).requires_grad_(True)
        boxes = odd(mv(c["boxes"])).requires_grad_(True)
        angles = mv(c["angles"])
        if angles is not None:
            angles = odd(angles).requires_grad_(True)
        attn = mv(c["attn"]).requires_grad_(True)
        out = boxer_b200.BoxGridAttnFunction.apply(value, c["shapes"].to(DEV), c["start"].to(DEV), boxes, angles,
                                                   odd(mv(c["vr"])), odd(mv(c["kidx"])), attn, 64)
        out.backward(mv(c["go"]))
        assert torch.equal(out.detach(), out_ref)
        assert helpers.rel_err(boxes.grad, g_ref[1]) <= 1e-6
        assert helpers.rel_err(attn.grad, g_ref[3]) <= 1e-6
        assert helpers.rel_err(value.grad, g_ref[0]) <= 1e-5      # float atomics: order differs run to run


def test_fused_deterministic_and_paths_agree():
    import boxer_b200
    c = _case("rot_k2_ratios_d64")
    a = _ours(c, torch.float32, deterministic=True)
    b = _ours(c, torch.float32, deterministic=True)
    assert all(torch.equal(x, y) for x, y in zip(a[1], b[1]) if x is not None)
    boxer_b200.ops.set_kernel_path("point")            # grid materialised inside + point kernels
    try:
        p = _ours(c, torch.float32)
    finally:
        boxer_b200.ops.set_kernel_path("auto")
    assert helpers.rel_err(a[0], p[0]) <= 2e-5
    assert helpers.rel_err(a[1][0], p[1][0]) <= 2e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4)], ids=["f64", "f32"])
@pytest.mark.parametrize("case", ["box_3d_refs_masked", "box_4d_refs_k3", "inst_infer", "box3d_rot", "box3d_norot_4d"])
def test_modules_with_fused_grid_match_reference_modules(case, dtype, tol):
    import boxer_b200
    spec = refinputs.module_cases()[case]
    gold = helpers.golden("modules_golden")[case]
    mod = getattr(boxer_b200, spec["cls"])(**spec["ctor"]).double()
    state = {k[len("param_"):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("param_")}
    mod.load_state_dict(state, strict=True)
    mod = mod.to(DEV, dtype)
    if spec["cls"] == "InstanceAttention":
        mod.inferencing = spec["inferencing"]
    args = [a.to(DEV, dtype) if (torch.is_tensor(a) and a.is_floating_point()) else (a.to(DEV) if torch.is_tensor(a) else a)
            for a in refinputs.module_inputs(spec)]
    boxer_b200.set_fused_grid(True)
    try:
        outs = mod(*args)
    finally:
        boxer_b200.set_fused_grid(False)
    flat = []
    for o in outs:
        if o is None:
            continue
        flat.extend(o if isinstance(o, tuple) else [o])
    assert len(flat) == int(gold["n_out"])
    for i, o in enumerate(flat):
        assert helpers.rel_err(o, gold[f"out{i}"]) <= tol, (case, i)


# ================================================================== fused softmax
def _ours_softmax(c, logits, dtype, deterministic=False):
    import boxer_b200
    tw = torch.float64 if dtype == torch.float64 else torch.float32
    mv = lambda t, dt: None if t is None else t.to(DEV, dt).contiguous()
    value = mv(c["value"], dtype).requires_grad_(True)
    boxes = mv(c["boxes"], tw).requires_grad_(True)
    angles = mv(c["angles"], tw)
    if angles is not None:
        angles.requires_grad_(True)
    z = mv(logits, tw).requires_grad_(True)
    boxer_b200.set_deterministic(deterministic)
    try:
        fn = boxer_b200.BoxGridSoftmaxAttnBf16Function if dtype == torch.bfloat16 else boxer_b200.BoxGridSoftmaxAttnFunction
        out, attn = fn.apply(value, c["shapes"].to(DEV), c["start"].to(DEV), boxes, angles, mv(c["vr"], tw), mv(c["kidx"], tw), z, 64)
        assert not attn.requires_grad
        out.backward(mv(c["go"], out.dtype))
    finally:
        boxer_b200.set_deterministic(None)
    return out.detach(), attn.detach(), (value.grad, boxes.grad, angles.grad if angles is not None else None, z.grad)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4), (torch.bfloat16, 1e-2)],
                         ids=["f64", "f32", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_fused_softmax_matches_reference_formulation(name, dtype, tol):
    """logits -> F.softmax over (L,K,K) (box_attention.py:227-231) -> grid -> grid_sample oracle, autograd in fp64,
    vs the op that takes the logits (in-kernel softmax where the window kernels apply, small kernels elsewhere)."""
    from oracle import plain
    c = _case(name)
    g = torch.Generator().manual_seed(len(name))
    B, Nq, H, L, K, _ = c["attn"].shape
    logits = 2.0 * torch.randn(B, Nq, H, L, K, K, generator=g, dtype=torch.float64)
    if dtype == torch.bfloat16:
        c["value"] = c["value"].bfloat16().double()
        c["go"] = c["go"].bfloat16().double()
    # oracle
    value = c["value"].clone().requires_grad_(True)
    boxes = c["boxes"].clone().requires_grad_(True)
    angles = c["angles"].clone().requires_grad_(True) if c["angles"] is not None else None
    z = logits.clone().requires_grad_(True)
    attn_ref = torch.softmax(z.view(B, Nq, H, -1), -1).view(B, Nq, H, L, K, K)
    grid = plain.grid_from_boxes(boxes, angles, c["vr"], c["kidx"])
    ref_out = plain.plain_box_attn(value.view(B, value.shape[1], -1), c["shapes"], 2 * grid - 1, attn_ref)
    ref_out.backward(c["go"])

    out, attn, gr = _ours_softmax(c, logits, dtype)
    wtol = 1e-9 if dtype == torch.float64 else 2e-6
    assert helpers.rel_err(attn, attn_ref.detach()) <= wtol, "attention weights"
    assert helpers.rel_err(out, ref_out.detach()) <= tol, "out"
    assert helpers.rel_err(gr[0], value.grad) <= tol, "grad_value"
    assert helpers.rel_err(gr[3], z.grad) <= tol, "grad_logits"
    if dtype == torch.float64:
        assert helpers.rel_err(gr[1], boxes.grad) <= tol, "grad_boxes"
        if angles is not None:
            assert helpers.rel_err(gr[2], angles.grad) <= tol, "grad_angles"
    # and deterministic mode chains the same softmax
    if dtype == torch.float32:
        d = _ours_softmax(c, logits, dtype, deterministic=True)
        assert helpers.rel_err(d[2][3], z.grad) <= tol, "grad_logits (deterministic)"


@pytest.mark.parametrize("case", ["box_3d_refs_masked", "box_4d_refs_k3", "box3d_rot", "box3d_norot_4d"])
def test_modules_with_fused_softmax_match_reference_modules(case):
    """BoxAttention / Box3dAttention with set_fused_grid + set_fused_softmax vs outputs of the reference's classes,
    and the parameter gradients of that path vs the default (reference op-for-op) path on the same device."""
    import boxer_b200
    spec = refinputs.module_cases()[case]
    gold = helpers.golden("modules_golden")[case]
    mod = getattr(boxer_b200, spec["cls"])(**spec["ctor"]).double()
    state = {k[len("param_"):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("param_")}
    mod.load_state_dict(state, strict=True)
    mod = mod.to(DEV, torch.float32)
    args = [a.to(DEV, torch.float32) if (torch.is_tensor(a) and a.is_floating_point()) else (a.to(DEV) if torch.is_tensor(a) else a)
            for a in refinputs.module_inputs(spec)]

    def run(fused):
        mod.zero_grad()
        boxer_b200.set_fused_grid(fused)
        boxer_b200.set_fused_softmax(fused)
        try:
            out, attn = mod(*args)
            out.square().sum().backward()
        finally:
            boxer_b200.set_fused_grid(False)
            boxer_b200.set_fused_softmax(False)
        return out.detach(), attn.detach(), {n: p.grad.clone() for n, p in mod.named_parameters()}

    out_f, attn_f, g_f = run(True)
    out_d, attn_d, g_d = run(False)
    assert helpers.rel_err(out_f, gold["out0"]) <= 1e-4
    assert helpers.rel_err(attn_f, gold["out1"]) <= 1e-5
    assert helpers.rel_err(out_f, out_d) <= 2e-5
    for n in g_d:
        assert helpers.rel_err(g_f[n], g_d[n]) <= 2e-4, n


# ================================================================== InstanceAttention weights (row f2, second half)
def _torch_instance_weights(z, K):
    """the reference's chain, box_attention.py:93-110"""
    b, nq, h, L = z.shape[:4]
    a = z.repeat_interleave(K // 2, dim=-1).repeat_interleave(K // 2, dim=-2)
    sw = torch.softmax(a.reshape(b, nq, h, -1), -1).view(b, nq, h, L, K, K)
    lw = torch.softmax(a.view(b, nq, h, L, K, K), dim=3)
    return sw, lw


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-6)], ids=["f64", "f32"])
@pytest.mark.parametrize("L,K", [(4, 14), (1, 2), (3, 4), (4, 28), (9, 6)])
def test_instance_weights_match_reference_chain(L, K, dtype, tol):
    import boxer_b200
    g = torch.Generator().manual_seed(100 * L + K)
    z64 = 3.0 * torch.randn(2, 7, 3, L, 2, 2, generator=g, dtype=torch.float64)
    gs = torch.randn(2, 7, 3, L, K, K, generator=g, dtype=torch.float64)
    gl = torch.randn(2, 7, 3, L, K, K, generator=g, dtype=torch.float64)
    zr = z64.clone().requires_grad_(True)
    sw_r, lw_r = _torch_instance_weights(zr, K)
    (sw_r * gs).sum().add((lw_r * gl).sum()).backward()

    z = z64.to(DEV, dtype).requires_grad_(True)
    sw, lw = boxer_b200.InstanceWeightsFunction.apply(z, K)
    assert sw.shape == sw_r.shape and lw.shape == lw_r.shape
    ((sw * gs.to(DEV, dtype)).sum() + (lw * gl.to(DEV, dtype)).sum()).backward()
    assert helpers.rel_err(sw, sw_r.detach()) <= tol
    assert helpers.rel_err(lw, lw_r.detach()) <= tol
    assert helpers.rel_err(z.grad, zr.grad) <= max(tol, 1e-12) * 10


def test_instance_weights_reject_odd_kernel_and_empty_is_fine():
    import boxer_b200
    with pytest.raises(RuntimeError, match="even"):
        boxer_b200.ops.instance_weights_forward(torch.zeros(1, 1, 1, 1, 2, 2, device=DEV), 3)
    sw, lw = boxer_b200.ops.instance_weights_forward(torch.zeros(1, 0, 2, 3, 2, 2, device=DEV), 4)
    assert sw.shape == (1, 0, 2, 3, 4, 4) and lw.numel() == 0


@pytest.mark.parametrize("case", ["inst_train", "inst_infer"])
def test_instance_module_with_fused_weights_matches_reference_module(case):
    import boxer_b200
    spec = refinputs.module_cases()[case]
    gold = helpers.golden("modules_golden")[case]
    mod = getattr(boxer_b200, spec["cls"])(**spec["ctor"]).double()
    state = {k[len("param_"):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("param_")}
    mod.load_state_dict(state, strict=True)
    mod = mod.to(DEV, torch.float32)
    mod.inferencing = spec["inferencing"]
    args = [a.to(DEV, torch.float32) if (torch.is_tensor(a) and a.is_floating_point()) else (a.to(DEV) if torch.is_tensor(a) else a)
            for a in refinputs.module_inputs(spec)]

    def run(fused):
        mod.zero_grad()
        boxer_b200.set_fused_softmax(fused)
        try:
            outs = mod(*args)
            loss = outs[0].square().sum()
            if outs[1] is not None:
                loss = loss + outs[1].square().sum()
            loss.backward()
        finally:
            boxer_b200.set_fused_softmax(False)
        flat = []
        for o in outs:
            if o is not None:
                flat.extend(o if isinstance(o, tuple) else [o])
        return [f.detach() for f in flat], {n: p.grad.clone() for n, p in mod.named_parameters()}

    flat_f, g_f = run(True)
    flat_d, g_d = run(False)
    assert len(flat_f) == int(gold["n_out"])
    for i, o in enumerate(flat_f):
        assert helpers.rel_err(o, gold[f"out{i}"]) <= 1e-4, (case, i)
    for n in g_d:
        assert helpers.rel_err(g_f[n], g_d[n]) <= 2e-4, n


# ================================================================== value_proj epilogue (row f3)
@pytest.mark.parametrize("C", [256, 30])
@pytest.mark.parametrize("din,dout", [(torch.float32, torch.bfloat16), (torch.float32, torch.float32),
                                      (torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("masked", [True, False])
def test_value_epilogue_is_masked_fill_plus_cast(C, din, dout, masked):
    """bit-exact vs ``value.masked_fill(mask[..., None], 0).to(dtype)`` (box_attention.py:222-225), forward and backward."""
    import boxer_b200
    g = torch.Generator().manual_seed(C)
    x = torch.randn(2, 301, C, generator=g).to(DEV, din).requires_grad_(True)
    mask = (torch.rand(2, 301, generator=g) < 0.3).to(DEV) if masked else None
    out = boxer_b200.ValueEpilogueFunction.apply(x, mask, dout)
    ref_in = x.detach().clone().requires_grad_(True)
    ref = (ref_in.masked_fill(mask[..., None], 0.0) if masked else ref_in).to(dout)
    assert out.dtype == dout and torch.equal(out, ref)
    go = torch.randn(2, 301, C, generator=g).to(DEV, dout)
    out.backward(go)
    ref.backward(go)
    assert x.grad.dtype == din and torch.equal(x.grad, ref_in.grad)


def test_bf16_native_module_uses_the_value_epilogue_and_matches_fp32():
    """BoxAttention under set_amp_native(True) + autocast (value masked and cast by the epilogue kernel) vs the
    default fp32 path: bf16 tolerance on the output, same masked pixels."""
    import boxer_b200
    spec = refinputs.module_cases()["box_3d_refs_masked"]
    gold = helpers.golden("modules_golden")["box_3d_refs_masked"]
    mod = boxer_b200.BoxAttention(**spec["ctor"]).double()
    mod.load_state_dict({k[len("param_"):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("param_")}, strict=True)
    mod = mod.to(DEV, torch.float32)
    args = [a.to(DEV, torch.float32) if (torch.is_tensor(a) and a.is_floating_point()) else (a.to(DEV) if torch.is_tensor(a) else a)
            for a in refinputs.module_inputs(spec)]
    ref = mod(*args)[0]
    boxer_b200.set_amp_native(True)
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = mod(*args)[0]
            out.float().sum().backward()
    finally:
        boxer_b200.set_amp_native(False)
    assert helpers.rel_err(out.float(), ref) <= 2e-2
    assert mod.value_proj.weight.grad is not None and bool(torch.isfinite(mod.value_proj.weight.grad).all())
