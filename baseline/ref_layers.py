"""Loader of the reference's UNMODIFIED transformer layers (bytecode under baseline/_ref/, see build_ref_layers.py)
and the few stand-ins their forward needs from the rest of BoxeR.

Used by tests/test_gpu_reference_layers.py (the layers on boxer_b200 vs the same layers on the reference's own
attention modules backed by the CPU oracle) and by bench.py's configs[2] leg (the layers as the caller of the op).
This is harness code: it never computes box attention itself.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(ROOT, "e2edet", "module", "box_transformer.pyc"))


def _purge():
    for n in [n for n in sys.modules if n == "e2edet" or n.startswith("e2edet.")]:
        sys.modules.pop(n, None)


def _shims():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if "torch._six" not in sys.modules:          # general.py:12 (removed from torch 2.x)
        six = types.ModuleType("torch._six")
        six.string_classes = (str, bytes)
        sys.modules["torch._six"] = six


def import_layers(attention: str = "boxer_b200", ops_module=None):
    """-> (box_transformer module, box3d_transformer module), freshly imported.

    attention="boxer_b200": ``e2edet.module.box_attention`` / ``e2edet.module.ops`` / ``e2edet.ops`` resolve to this
    repo through ``boxer_b200.compat.install(lightweight=True)`` -- the drop-in path.
    attention="reference": the reference's own ``box_attention.py``; ``ops_module`` must provide ``BoxAttnFunction``
    and ``InstanceAttnFunction`` (the reference's come from its CUDA extension; the tests hand in oracle-backed ones).
    """
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python -m baseline.build_ref_layers` where /root/reference exists")
    import boxer_b200
    boxer_b200.compat.uninstall()
    _purge()
    _shims()
    if attention == "boxer_b200":
        boxer_b200.compat.install(lightweight=True)
    elif attention == "reference":
        for name, sub in (("e2edet", ""), ("e2edet.module", "module"), ("e2edet.utils", "utils")):
            boxer_b200.compat._bare_package(name, os.path.join(ROOT, "e2edet", sub) if sub else os.path.join(ROOT, "e2edet"), placeholder=False)
        ops_pkg = types.ModuleType("e2edet.module.ops")
        ops_pkg.__path__ = []
        ops_pkg.BoxAttnFunction = ops_module.BoxAttnFunction
        ops_pkg.InstanceAttnFunction = ops_module.InstanceAttnFunction
        sys.modules["e2edet.module.ops"] = ops_pkg
    else:
        raise ValueError(attention)
    try:
        bt = importlib.import_module("e2edet.module.box_transformer")
        b3 = importlib.import_module("e2edet.module.box3d_transformer")
    finally:
        boxer_b200.compat.uninstall()
        _purge()
    return bt, b3


class _MLP(nn.Module):
    def __init__(self, d_in, d_hidden, d_out, n):
        super().__init__()
        dims = [d_in] + [d_hidden] * (n - 1) + [d_out]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i + 1 < len(self.layers):
                x = torch.relu(x)
        return x


class DetectorStub2d(nn.Module):
    """What BoxTransformerEncoder._get_enc_proposals reads from ``self.detector[0]`` (box_transformer.py:195-203; the
    real one is predictor.py:60-68): a class head and a 3-layer box head."""

    def __init__(self, d_model, num_classes=91):
        super().__init__()
        self.class_embed = nn.Linear(d_model, num_classes)
        self.bbox_embed = _MLP(d_model, d_model, 4, 3)


class DetectorStub3d(nn.Module):
    """box3d_transformer.py:143-175 reads ``num_references``, ``bbox_embed`` (7 per reference) and ``class_embed``."""

    def __init__(self, d_model, num_references=8, num_classes=3):
        super().__init__()
        self.num_references = num_references
        self.class_embed = nn.Linear(d_model, num_references * num_classes)
        self.bbox_embed = _MLP(d_model, d_model, num_references * 7, 3)


def set_inferencing(model: nn.Module, mode: bool):
    """base_model.py:49-67: BoxeR injects ``inferencing`` into every sub-module."""
    for m in model.modules():
        m.inferencing = mode


def make_boxer2d(bt, d_model=256, nhead=8, nlevel=4, enc=6, dec=6, ffn=1024, num_queries=300, use_mask=True,
                 dropout=0.0, residual_mode="v1", seed=0):
    torch.manual_seed(seed)
    t = bt.BoxTransformer(d_model=d_model, nhead=nhead, nlevel=nlevel, num_encoder_layers=enc, num_decoder_layers=dec,
                          dim_feedforward=ffn, dropout=dropout, num_queries=num_queries, use_mask=use_mask,
                          residual_mode=residual_mode)
    det = nn.ModuleList([DetectorStub2d(d_model)])
    t.encoder.detector = det
    t.decoder.detector = det
    set_inferencing(t, False)
    return t


def make_boxer3d(b3, d_model=256, nhead=8, nlevel=2, enc=3, dec=3, ffn=1024, num_queries=300, dropout=0.0, seed=0):
    torch.manual_seed(seed)
    t = b3.Box3dTransformer(d_model=d_model, nhead=nhead, nlevel=nlevel, num_encoder_layers=enc, num_decoder_layers=dec,
                            dim_feedforward=ffn, dropout=dropout, num_queries=num_queries)
    t.encoder.detector = nn.ModuleList([DetectorStub3d(d_model, num_references=nhead)])
    set_inferencing(t, False)
    return t


def padded_batch(shapes, B, d_model, valid=None, device="cpu", dtype=torch.float32, seed=1):
    """src / mask / pos lists as the backbone + collate_fn hand them to BoxTransformer.forward
    (dataset/helper/collate_fn.py:66-84: images padded to the batch maximum, mask True on padding).
    valid[b] = (fraction of height, fraction of width) that is image."""
    g = torch.Generator().manual_seed(seed)
    src, mask, pos = [], [], []
    for (h, w) in shapes:
        src.append(torch.randn(B, d_model, h, w, generator=g).to(device=device, dtype=dtype))
        pos.append((0.1 * torch.randn(B, d_model, h, w, generator=g)).to(device=device, dtype=dtype))
        m = torch.zeros(B, h, w, dtype=torch.bool)
        if valid is not None:
            for b, (fh, fw) in enumerate(valid):
                m[b, max(1, int(round(h * fh))):, :] = True
                m[b, :, max(1, int(round(w * fw))):] = True
        mask.append(m.to(device))
    return src, (mask if valid is not None else [None] * len(shapes)), pos
