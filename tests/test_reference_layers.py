"""The reference's UNMODIFIED transformer layers on the drop-in.

``BoxTransformer`` (e2edet/module/box_transformer.py:16-465: encoder layers with ``BoxAttention``, decoder layers with
``InstanceAttention`` / ``BoxAttention``, ``use_mask``, padded batch + masks + valid ratios) and ``Box3dTransformer``
(e2edet/module/box3d_transformer.py:18-322: ``Box3dAttention`` without rotation in the encoder, with rotation in the
decoder) are imported from bytecode compiled from the reference where it lies (baseline/build_ref_layers.py) -- once
with ``boxer_b200.compat.install()`` providing the attention modules and ops (the drop-in path), once with the
reference's own ``box_attention.py`` whose two autograd Functions are backed by the CPU oracle (the reference has no
CPU op).  Same weights, same inputs: outputs and gradients must agree.

CPU leg (fp64 both sides; our modules run with the native op swapped for the oracle): <= 1e-9.
GPU leg (ours: fp32 on the GPU through the C ABI; reference side: fp64 on the CPU): <= 5e-4 max-norm over the whole
encoder + decoder stack (the op itself is held to 1e-4 elsewhere).
"""
import pytest
import torch

from baseline import ref_layers as R
from tests import helpers

pytestmark = pytest.mark.skipif(not R.available(), reason="baseline/_ref not built (python -m baseline.build_ref_layers)")


class _OracleOps:
    """Stands in for the reference's e2edet.module.ops (its Functions call the CUDA extension; box_attention_func.py:9-150)."""

    class BoxAttnFunction:
        @staticmethod
        def apply(value, shapes, lsi, loc, w, step):
            from oracle import plain
            b, s = value.shape[:2]
            return plain.plain_box_attn(value.reshape(b, s, -1), shapes, 2 * loc - 1, w)

    class InstanceAttnFunction:
        @staticmethod
        def apply(value, shapes, lsi, loc, sw, lw, k, step):
            from oracle import plain
            b, s = value.shape[:2]
            return plain.plain_instance_attn(value.reshape(b, s, -1), shapes, 2 * loc - 1, sw, lw, k)


def _canon(ref_windows):
    """permutation that sorts the decoder queries by their reference window (topk(sorted=False) orders them differently
    on CPU and GPU; everything downstream is permutation-equivariant)"""
    key = ref_windows[..., 0].double() * 1e3 + ref_windows[..., 1].double()
    return key.argsort(dim=1)


def _gather_q(t, perm, dim):
    idx = perm
    for _ in range(t.dim() - dim - 1):
        idx = idx.unsqueeze(-1)
    shape = list(t.shape)
    view = [1] * dim + list(perm.shape) if dim == 0 else None
    if dim == 1:        # (B, Nq, ...)
        return torch.gather(t, 1, idx.expand(*t.shape))
    # (nl, B, Nq, ...)
    return torch.gather(t, 2, idx.unsqueeze(0).expand(*t.shape))


def _run2d(model, src, mask, pos, wts):
    src = [s.clone().requires_grad_(True) for s in src]
    hs, roi, dec_ref, out_embed, src_ref, src_mask = model(src, mask, pos)
    perm = _canon(dec_ref)
    hs_c = _gather_q(hs, perm, 2)
    loss = (hs_c * wts["hs"].to(hs_c)).sum() + (out_embed * wts["mem"].to(out_embed)).sum()
    res = {"hs": hs_c, "dec_ref": _gather_q(dec_ref, perm, 1), "out_embed": out_embed, "src_ref": src_ref}
    if roi is not None:
        roi_c = _gather_q(roi, perm, 2)
        loss = loss + (roi_c * wts["roi"].to(roi_c)).sum()
        res["roi"] = roi_c
    loss.backward()
    res.update({f"grad_src{i}": s.grad for i, s in enumerate(src)})
    a0 = model.encoder.layers[0].self_attn
    d0 = model.decoder.layers[-1].multihead_attn
    res.update(g_enc_box_w=a0.linear_box_weight.grad, g_enc_attn_w=a0.linear_attn_weight.grad, g_enc_vproj=a0.value_proj.weight.grad,
               g_dec_box_w=d0.linear_box_weight.grad, g_dec_attn_w=d0.linear_attn_weight.grad, g_dec_out=d0.out_proj.weight.grad)
    return {k: v.detach().double().cpu() for k, v in res.items()}


def _run3d(model, src, pos, wts):
    src = [s.clone().requires_grad_(True) for s in src]
    hs, dec_ref, out_embed, src_ref = model(src, pos)
    perm = _canon(dec_ref)
    hs_c = _gather_q(hs, perm, 2)
    loss = (hs_c * wts["hs"].to(hs_c)).sum() + (out_embed * wts["mem"].to(out_embed)).sum()
    loss.backward()
    res = {"hs": hs_c, "dec_ref": _gather_q(dec_ref, perm, 1), "out_embed": out_embed}
    res.update({f"grad_src{i}": s.grad for i, s in enumerate(src)})
    a0 = model.encoder.layers[0].self_attn
    d0 = model.decoder.layers[-1].multihead_attn
    res.update(g_enc_box_w=a0.linear_box_weight.grad, g_enc_attn_w=a0.linear_attn_weight.grad,
               g_dec_box_w=d0.linear_box_weight.grad, g_dec_attn_w=d0.linear_attn_weight.grad)
    return {k: v.detach().double().cpu() for k, v in res.items()}


def _randomize(model, seed):
    """trained-like parameters: the init zeroes linear_box_weight / linear_attn_weight, which would hide their paths"""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "linear_box_weight" in n or "linear_attn_weight" in n:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif "linear_attn_bias" in n:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))


SHAPES_2D = [(20, 28), (10, 14), (5, 7), (3, 4)]
SHAPES_3D = [(24, 24), (12, 12)]


def _pair2d(dev, dtype, use_mask, residual_mode="v1"):
    bt_ref, _ = R.import_layers("reference", ops_module=_OracleOps)
    bt_our, _ = R.import_layers("boxer_b200")
    kw = dict(d_model=64, nhead=8, nlevel=4, enc=2, dec=2, ffn=96, num_queries=12, use_mask=use_mask, residual_mode=residual_mode)
    ref = R.make_boxer2d(bt_ref, **kw).double()
    _randomize(ref, 5)
    our = R.make_boxer2d(bt_our, **kw)
    assert set(our.state_dict()) == set(ref.state_dict())
    our.load_state_dict(ref.state_dict(), strict=True)
    import boxer_b200
    assert isinstance(our.encoder.layers[0].self_attn, boxer_b200.BoxAttention)
    assert isinstance(our.decoder.layers[0].multihead_attn, boxer_b200.InstanceAttention if use_mask else boxer_b200.BoxAttention)
    assert type(ref.encoder.layers[0].self_attn).__module__ == "e2edet.module.box_attention"      # the reference's class
    return ref, our.to(device=dev, dtype=dtype)


def _inputs2d(dev, dtype):
    src, mask, pos = R.padded_batch(SHAPES_2D, 2, 64, valid=[(1.0, 1.0), (0.7, 0.8)], dtype=torch.float64)
    mv = lambda ts: [t.to(device=dev, dtype=dtype) if t.is_floating_point() else t.to(dev) for t in ts]
    g = torch.Generator().manual_seed(9)
    S = sum(h * w for h, w in SHAPES_2D)
    wts = {"hs": torch.randn(2, 2, 12, 64, generator=g, dtype=torch.float64), "mem": torch.randn(2, S, 64, generator=g, dtype=torch.float64),
           "roi": torch.randn(2, 2, 12, 14, 14, 64, generator=g, dtype=torch.float64)}
    return (src, mask, pos), (mv(src), mv(mask), mv(pos)), wts


def _compare(a, b, tol):
    assert set(a) == set(b)
    for k in a:
        err = helpers.rel_err(b[k], a[k])
        assert err <= tol, f"{k}: max-norm relative error {err:.3e} > {tol:g}"


def _oracle_box_fn(value, v_shape, v_start, grid, weights, step):
    return _OracleOps.BoxAttnFunction.apply(value, v_shape, v_start, grid, weights, step)


def _oracle_inst_fn(value, v_shape, v_start, grid, sw, lw, k, step):
    return _OracleOps.InstanceAttnFunction.apply(value, v_shape, v_start, grid, sw, lw, k, step)


@pytest.mark.parametrize("use_mask,residual_mode", [(True, "v1"), (True, "v2"), (False, "v1")])
def test_box_transformer_on_the_module_mirror_cpu(use_mask, residual_mode, monkeypatch):
    from boxer_b200 import box_attention as BA
    monkeypatch.setattr(BA, "_box_attn", _oracle_box_fn)
    monkeypatch.setattr(BA, "_instance_attn", _oracle_inst_fn)
    ref, our = _pair2d("cpu", torch.float64, use_mask, residual_mode)
    (src, mask, pos), _, wts = _inputs2d("cpu", torch.float64)
    _compare(_run2d(ref, src, mask, pos, wts), _run2d(our, src, mask, pos, wts), 1e-9)


def _pair3d(dev, dtype):
    _, b3_ref = R.import_layers("reference", ops_module=_OracleOps)
    _, b3_our = R.import_layers("boxer_b200")
    kw = dict(d_model=64, nhead=8, nlevel=2, enc=2, dec=2, ffn=96, num_queries=12)
    ref = R.make_boxer3d(b3_ref, **kw).double()
    _randomize(ref, 6)
    our = R.make_boxer3d(b3_our, **kw)
    our.load_state_dict(ref.state_dict(), strict=True)
    import boxer_b200
    assert isinstance(our.encoder.layers[0].self_attn, boxer_b200.Box3dAttention) and not our.encoder.layers[0].self_attn.with_rotation
    assert our.decoder.layers[0].multihead_attn.with_rotation
    return ref, our.to(device=dev, dtype=dtype)


def _inputs3d(dev, dtype):
    src, _, pos = R.padded_batch(SHAPES_3D, 2, 64, dtype=torch.float64, seed=3)
    mv = lambda ts: [t.to(device=dev, dtype=dtype) for t in ts]
    g = torch.Generator().manual_seed(10)
    S = sum(h * w for h, w in SHAPES_3D)
    wts = {"hs": torch.randn(2, 2, 12, 64, generator=g, dtype=torch.float64), "mem": torch.randn(2, S, 64, generator=g, dtype=torch.float64)}
    return (src, pos), (mv(src), mv(pos)), wts


def test_box3d_transformer_on_the_module_mirror_cpu(monkeypatch):
    from boxer_b200 import box_attention as BA
    monkeypatch.setattr(BA, "_box_attn", _oracle_box_fn)
    ref, our = _pair3d("cpu", torch.float64)
    (src, pos), _, wts = _inputs3d("cpu", torch.float64)
    _compare(_run3d(ref, src, pos, wts), _run3d(our, src, pos, wts), 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True], ids=["default", "fused-grid+softmax"])
@pytest.mark.parametrize("use_mask,residual_mode", [(True, "v1"), (True, "v2"), (False, "v1")])
def test_box_transformer_on_the_drop_in_gpu(use_mask, residual_mode, fused):
    import boxer_b200
    ref, our = _pair2d("cuda", torch.float32, use_mask, residual_mode)
    (src, mask, pos), (gsrc, gmask, gpos), wts = _inputs2d("cuda", torch.float32)
    want = _run2d(ref, src, mask, pos, wts)
    boxer_b200.set_fused_grid(fused)
    boxer_b200.set_fused_softmax(fused)
    try:
        got = _run2d(our, gsrc, gmask, gpos, wts)
    finally:
        boxer_b200.set_fused_grid(False)
        boxer_b200.set_fused_softmax(False)
    _compare(want, got, 5e-4)


@pytest.mark.gpu
def test_box3d_transformer_on_the_drop_in_gpu():
    ref, our = _pair3d("cuda", torch.float32)
    (src, pos), (gsrc, gpos), wts = _inputs3d("cuda", torch.float32)
    _compare(_run3d(ref, src, pos, wts), _run3d(our, gsrc, gpos, wts), 5e-4)
