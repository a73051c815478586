"""CPU: pin both oracle restatements against fixtures produced by the reference's own code.

tests/golden/*.npz hold outputs of /root/reference's PlainBoxAttnFunction /
PlainInstanceAttnFunction (+ autograd) on the reference tests' seeded inputs
(see tests/golden/make_golden.py).  oracle/plain.py (grid_sample formulation)
and oracle/kernel_ref.c (CUDA-kernel formulation) must both reproduce them.
"""
import pytest
import torch

from oracle import kernel_ref, plain
from tests import helpers

F64_TOL = 1e-12   # fp64 vs fp64, different summation order only
F32_TOL = 1e-6    # fp32 case of the reference test (values ~1e-2)


def _box_oracle_plain(inp, grad_out=None):
    value, loc, attn = (inp[k].clone().requires_grad_(grad_out is not None) for k in ("value", "loc", "attn"))
    B, S = value.shape[:2]
    out = plain.plain_box_attn(value.view(B, S, -1), inp["shapes"], 2 * loc - 1, attn)
    if grad_out is None:
        return out.detach(), None
    out.backward(grad_out)
    return out.detach(), (value.grad, loc.grad, attn.grad)


@pytest.mark.parametrize("case", list(helpers.box_inputs()))
def test_box_plain_matches_reference(case):
    inp, gold = helpers.box_inputs()[case], helpers.golden("box_attn_golden")[case]
    helpers.check_digest(inp, gold)
    go = torch.from_numpy(gold["grad_out"]) if "grad_out" in gold else None
    out, grads = _box_oracle_plain(inp, go)
    tol = F32_TOL if inp["value"].dtype == torch.float32 else F64_TOL
    assert helpers.max_err(out, gold["out"]) <= tol
    if grads is not None:
        gv, gl, ga = grads
        assert helpers.max_err(helpers.slim_like(gv, gold["grad_value"]), gold["grad_value"]) <= tol
        assert helpers.max_err(gl, gold["grad_loc"]) <= 1e-10
        assert helpers.max_err(ga, gold["grad_attn"]) <= tol


@pytest.mark.parametrize("case", list(helpers.box_inputs()))
def test_box_kernel_ref_matches_reference(case):
    inp, gold = helpers.box_inputs()[case], helpers.golden("box_attn_golden")[case]
    args = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["attn"])
    out = kernel_ref.box_attn_forward(*args)
    f32 = inp["value"].dtype == torch.float32
    assert out.dtype == inp["value"].dtype
    assert helpers.max_err(out, gold["out"]) <= (F32_TOL if f32 else 1e-11)
    if "grad_out" in gold:
        gv, gl, ga = kernel_ref.box_attn_backward(*args, gold["grad_out"])
        assert gl.shape == inp["loc"].shape and ga.shape == inp["attn"].shape
        assert helpers.max_err(helpers.slim_like(gv, gold["grad_value"]), gold["grad_value"]) <= 1e-11
        assert helpers.max_err(gl, gold["grad_loc"]) <= 1e-9
        assert helpers.max_err(ga, gold["grad_attn"]) <= 1e-11


def _inst_oracle_plain(inp, grad_out=None, grad_mask=None):
    need = grad_out is not None
    value, loc, sw, lw = (inp[k].clone().requires_grad_(need) for k in ("value", "loc", "spatial_w", "level_w"))
    B, S = value.shape[:2]
    out, mask = plain.plain_instance_attn(value.view(B, S, -1), inp["shapes"], 2 * loc - 1, sw, lw, inp["mask_size"])
    if not need:
        return out.detach(), mask.detach(), None
    torch.autograd.backward([out, mask], [grad_out, grad_mask])
    return out.detach(), mask.detach(), (value.grad, loc.grad, sw.grad, lw.grad)


def _inst_grads_in(case, inp, gold):
    if "grad_out" not in gold:
        return None, None
    go = torch.from_numpy(gold["grad_out"])
    if "grad_mask" in gold:
        gm = torch.from_numpy(gold["grad_mask"])
    else:  # big-D gradcheck cases: regenerated from its seed (make_golden.py)
        D = inp["value"].shape[-1]
        from tests import refinputs
        gm = refinputs.side_rand((1, refinputs.LQ, 2, 2, refinputs.M * D), 3000 + D, torch.float64)
    return go, gm


@pytest.mark.parametrize("case", list(helpers.instance_inputs()))
def test_instance_plain_matches_reference(case):
    inp, gold = helpers.instance_inputs()[case], helpers.golden("instance_attn_golden")[case]
    helpers.check_digest(inp, gold)
    go, gm = _inst_grads_in(case, inp, gold)
    out, mask, grads = _inst_oracle_plain(inp, go, gm)
    tol = F32_TOL if inp["value"].dtype == torch.float32 else F64_TOL
    assert helpers.max_err(out, gold["out"]) <= tol
    assert helpers.max_err(mask, gold["mask_out"]) <= tol
    if grads is not None:
        gv, gl, gs, gw = grads
        assert helpers.max_err(helpers.slim_like(gv, gold["grad_value"]), gold["grad_value"]) <= tol
        assert helpers.max_err(gl, gold["grad_loc"]) <= 1e-10
        assert helpers.max_err(gs, gold["grad_spatial_w"]) <= tol
        assert helpers.max_err(gw, gold["grad_level_w"]) <= tol


@pytest.mark.parametrize("case", list(helpers.instance_inputs()))
def test_instance_kernel_ref_matches_reference(case):
    inp, gold = helpers.instance_inputs()[case], helpers.golden("instance_attn_golden")[case]
    K = inp["mask_size"]
    args = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["spatial_w"], inp["level_w"])
    out, mask = kernel_ref.instance_attn_forward(*args)
    B, Nq = inp["loc"].shape[:2]
    mask = mask.view(B, Nq, K, K, -1)
    f32 = inp["value"].dtype == torch.float32
    assert helpers.max_err(out, gold["out"]) <= (F32_TOL if f32 else 1e-11)
    assert helpers.max_err(mask, gold["mask_out"]) <= (F32_TOL if f32 else 1e-11)
    go, gm = _inst_grads_in(case, inp, gold)
    if go is not None:
        gv, gl, gs, gw = kernel_ref.instance_attn_backward(*args, go, gm)
        assert helpers.max_err(helpers.slim_like(gv, gold["grad_value"]), gold["grad_value"]) <= 1e-11
        assert helpers.max_err(gl, gold["grad_loc"]) <= 1e-9
        assert helpers.max_err(gs, gold["grad_spatial_w"]) <= 1e-11
        assert helpers.max_err(gw, gold["grad_level_w"]) <= 1e-11
