"""Golden-fixture access shared by the CPU and GPU tests."""
from __future__ import annotations

import os

import numpy as np
import torch

from tests import refinputs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache: dict = {}


def golden(name: str) -> dict:
    """{case: {key: ndarray}} of tests/golden/<name>.npz"""
    if name not in _cache:
        tree: dict = {}
        with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
            for k in z.files:
                case, key = k.split("/", 1)
                tree.setdefault(case, {})[key] = z[k]
        _cache[name] = tree
    return _cache[name]


_inputs: dict = {}


def box_inputs() -> dict:
    if "box" not in _inputs:
        d = dict(refinputs.box_test_sequence())
        d.update(refinputs.box_wide_cases())
        _inputs["box"] = d
    return _inputs["box"]


def instance_inputs() -> dict:
    if "inst" not in _inputs:
        d = dict(refinputs.instance_test_sequence())
        d.update(refinputs.instance_wide_cases())
        _inputs["inst"] = d
    return _inputs["inst"]


def check_digest(inp: dict, gold: dict):
    assert np.array_equal(refinputs.digest(inp), gold["input_digest"]), \
        "regenerated inputs differ from the ones the golden outputs were computed on (RNG drift)"


def slim_like(gv: torch.Tensor, gold_gv: np.ndarray) -> torch.Tensor:
    """Golden grad_value of the big-D cases holds the first/last 8 channels only."""
    if gv.shape[-1] == gold_gv.shape[-1]:
        return gv
    return torch.cat([gv[..., :8], gv[..., -8:]], dim=-1)


def max_err(a, b) -> float:
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) if a.numel() else 0.0


def rel_err(a, b) -> float:
    """max |a-b| / max(|b|): scale-free error used with the 1e-4 / 1e-2 bars."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if not a.numel():
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
