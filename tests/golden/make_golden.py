#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the REFERENCE's own Python code.

Run in the authoring container only (needs /root/reference, CPU only):

    python tests/golden/make_golden.py

Nothing from the reference is copied into the repo: the source text of the
reference functions / classes is read from /root/reference at run time,
exec'd in a scratch namespace, run on seeded inputs, and only the resulting
tensors are stored.  What is executed:

* ``PlainBoxAttnFunction``       tests/box_attn_test.py:9-42
* ``PlainInstanceAttnFunction``  tests/instance_attn_test.py:11-63
* ``view_with_shape``            e2edet/utils/general.py:289-324
* ``BoxAttention``, ``InstanceAttention``, ``Box3dAttention``
                                 e2edet/module/box_attention.py:10,140,242
  (with ``e2edet.module.ops`` stubbed so that the two autograd Functions
  resolve to the reference's Plain* oracles -- the reference has no CPU op).

Inputs: the reference tests' inputs *as those scripts draw them*
(``torch.manual_seed(3)`` then the exact order of ``torch.rand`` calls of
``__main__``, see ``tests/refinputs.py``) plus a few wider seeded cases
(several levels, B > 1, out-of-range locations, 6-D weights).
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from tests import refinputs  # noqa: E402


def _extract(path: str, names: list[str], prelude: str = "") -> dict:
    """exec the named top-level defs/classes of a reference file in a fresh namespace."""
    src = open(path).read()
    tree = ast.parse(src)
    picked = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(picked) == len(names), (path, names)
    ns: dict = {}
    exec(prelude, ns)
    for node in picked:
        exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns


def load_reference():
    general = _extract(f"{REF}/e2edet/utils/general.py", ["view_with_shape"], "import torch\n")
    prelude = "import math\nimport torch\nimport torch.nn as nn\nimport torch.nn.functional as F\n"
    box = _extract(f"{REF}/tests/box_attn_test.py", ["PlainBoxAttnFunction"], prelude)
    box["view_with_shape"] = general["view_with_shape"]
    inst = _extract(f"{REF}/tests/instance_attn_test.py", ["PlainInstanceAttnFunction"], prelude)
    inst["view_with_shape"] = general["view_with_shape"]
    plain_box, plain_inst = box["PlainBoxAttnFunction"], inst["PlainInstanceAttnFunction"]

    class _BoxFn:  # stands in for e2edet.module.ops.BoxAttnFunction on CPU
        @staticmethod
        def apply(value, shapes, lsi, loc, w, step):
            b, s = value.shape[:2]
            w5 = w.reshape(*loc.shape[:5])
            return plain_box(value.reshape(b, s, -1), shapes, 2 * loc - 1, w5)

    class _InstFn:
        @staticmethod
        def apply(value, shapes, lsi, loc, sw, lw, mask_size, step):
            b, s = value.shape[:2]
            return plain_inst(value.reshape(b, s, -1), shapes, 2 * loc - 1, sw, lw, mask_size)

    mods = _extract(
        f"{REF}/e2edet/module/box_attention.py",
        ["BoxAttention", "InstanceAttention", "Box3dAttention"],
        prelude,
    )
    mods["BoxAttnFunction"] = _BoxFn
    mods["InstanceAttnFunction"] = _InstFn
    return plain_box, plain_inst, mods


def _np(t):
    return t.detach().cpu().numpy()


def box_case(plain_box, inp, with_grad, grad_out=None):
    """Run the reference oracle as tests/box_attn_test.py does (grid = 2*loc-1)."""
    value, loc, attn = (inp[k].clone() for k in ("value", "loc", "attn"))
    shapes = inp["shapes"]
    B, S = value.shape[:2]
    if with_grad:
        for t in (value, loc, attn):
            t.requires_grad_(True)
    out = plain_box(value.view(B, S, -1), shapes, 2 * loc - 1, attn.reshape(*loc.shape[:5]))
    res = {"out": _np(out)}
    if with_grad:
        go = torch.ones_like(out) if grad_out is None else grad_out
        out.backward(go)
        res.update(grad_out=_np(go), grad_value=_np(value.grad), grad_loc=_np(loc.grad), grad_attn=_np(attn.grad))
    return res


def inst_case(plain_inst, inp, with_grad, grad_out=None, grad_mask=None):
    value, loc, sw, lw = (inp[k].clone() for k in ("value", "loc", "spatial_w", "level_w"))
    shapes, K = inp["shapes"], inp["mask_size"]
    B, S = value.shape[:2]
    if with_grad:
        for t in (value, loc, sw, lw):
            t.requires_grad_(True)
    # the reference's gradcheck hands the CUDA op 5-D weights (instance_attn_test.py:259-268); its
    # Plain oracle indexes them as (..., L, K, K), so view them that way (grads land on the leaf)
    six = (*loc.shape[:4], K, K)
    out, mask = plain_inst(value.view(B, S, -1), shapes, 2 * loc - 1, sw.view(six), lw.view(six), K)
    res = {"out": _np(out), "mask_out": _np(mask)}
    if with_grad:
        go = torch.ones_like(out) if grad_out is None else grad_out
        gm = torch.ones_like(mask) if grad_mask is None else grad_mask
        torch.autograd.backward([out, mask], [go, gm])
        res.update(grad_out=_np(go), grad_mask=_np(gm), grad_value=_np(value.grad), grad_loc=_np(loc.grad),
                   grad_spatial_w=_np(sw.grad), grad_level_w=_np(lw.grad))
    return res


def _slim_grad_value(gv, D):
    """Big-D gradcheck cases: keep the first/last 8 channels of each head only."""
    if D <= 128:
        return gv
    return np.concatenate([gv[..., :8], gv[..., -8:]], axis=-1)


def save(name, tree):
    flat = {}
    for case, d in tree.items():
        for k, v in d.items():
            flat[f"{case}/{k}"] = np.asarray(v)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **flat)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, {len(flat)} arrays")


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    plain_box, plain_inst, mods = load_reference()

    # ---------------------------------------------------------------- box op
    tree = {}
    seq = refinputs.box_test_sequence()
    for name, inp in seq.items():
        with_grad = name.startswith("gradcheck") or name == "fwdbwd_double"
        go = None
        if name.startswith("gradcheck"):
            D = inp["value"].shape[-1]
            go = refinputs.side_rand((1, refinputs.LQ, refinputs.M * D), 1000 + D, torch.float64)
        res = box_case(plain_box, inp, with_grad, go)
        if name.startswith("gradcheck"):
            res["grad_value"] = _slim_grad_value(res["grad_value"], inp["value"].shape[-1])
        res["input_digest"] = refinputs.digest(inp)
        if not name.startswith("gradcheck"):     # small: store the inputs too
            res.update({f"in_{k}": _np(v) for k, v in inp.items() if torch.is_tensor(v)})
        tree[name] = res
    for name, inp in refinputs.box_wide_cases().items():
        go = refinputs.side_rand((inp["value"].shape[0], inp["loc"].shape[1], inp["value"].shape[2] * inp["value"].shape[3]),
                                 77, torch.float64)
        res = box_case(plain_box, inp, True, go)
        res["input_digest"] = refinputs.digest(inp)
        tree[name] = res
    save("box_attn_golden.npz", tree)

    # ----------------------------------------------------------- instance op
    tree = {}
    seq = refinputs.instance_test_sequence()
    for name, inp in seq.items():
        with_grad = name.startswith("gradcheck") or name == "fwdbwd_double"
        go = gm = None
        if name.startswith("gradcheck"):
            D = inp["value"].shape[-1]
            go = refinputs.side_rand((1, refinputs.LQ, refinputs.M * D), 2000 + D, torch.float64)
            gm = refinputs.side_rand((1, refinputs.LQ, 2, 2, refinputs.M * D), 3000 + D, torch.float64)
        res = inst_case(plain_inst, inp, with_grad, go, gm)
        if name.startswith("gradcheck"):
            res["grad_value"] = _slim_grad_value(res["grad_value"], inp["value"].shape[-1])
            D = inp["value"].shape[-1]
            if D > 128:
                res.pop("grad_mask")  # regenerated from its seed in the test
        res["input_digest"] = refinputs.digest(inp)
        if not name.startswith("gradcheck"):
            res.update({f"in_{k}": _np(v) for k, v in inp.items() if torch.is_tensor(v)})
        tree[name] = res
    for name, inp in refinputs.instance_wide_cases().items():
        B, Nq = inp["loc"].shape[:2]
        C = inp["value"].shape[2] * inp["value"].shape[3]
        K = inp["mask_size"]
        go = refinputs.side_rand((B, Nq, C), 78, torch.float64)
        gm = refinputs.side_rand((B, Nq, K, K, C), 79, torch.float64)
        res = inst_case(plain_inst, inp, True, go, gm)
        res["input_digest"] = refinputs.digest(inp)
        tree[name] = res
    save("instance_attn_golden.npz", tree)

    # --------------------------------------------------------------- modules
    tree = {}
    for name, spec in refinputs.module_cases().items():
        cls = mods[spec["cls"]]
        torch.manual_seed(spec["seed"])
        mod = cls(**spec["ctor"]).double()
        refinputs.randomize_module(mod, spec["seed"] + 1)
        if spec["cls"] == "InstanceAttention":
            mod.inferencing = spec["inferencing"]
        args = refinputs.module_inputs(spec)
        outs = mod(*args)
        res = {f"param_{k}": _np(v) for k, v in mod.state_dict().items()}
        flat = []
        for o in outs:
            if o is None:
                continue
            flat.extend(o if isinstance(o, tuple) else [o])
        for i, o in enumerate(flat):
            res[f"out{i}"] = _np(o)
        res["n_out"] = np.int64(len(flat))
        tree[name] = res
    save("modules_golden.npz", tree)


if __name__ == "__main__":
    main()
