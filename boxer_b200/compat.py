"""Make the reference's import paths resolve to this package.

BoxeR's transformer layers do (box_transformer.py:6, box3d_transformer.py:6)

    from e2edet.module.box_attention import BoxAttention, InstanceAttention, Box3dAttention

its autograd layer does ``from e2edet import ops`` (box_attention_func.py:6) and its
tests ``from e2edet.module.ops import BoxAttnFunction`` (tests/box_attn_test.py:5).
``install()`` registers modules under exactly those names in ``sys.modules``:

* ``e2edet.ops``                   -> boxer_b200.ops            (the pybind module's four functions)
* ``e2edet.module.ops``            -> BoxAttnFunction, InstanceAttnFunction
* ``e2edet.module.ops.box_attention_func`` -> same
* ``e2edet.module.box_attention``  -> BoxAttention, InstanceAttention, Box3dAttention

If the real ``e2edet`` package is importable, only those sub-modules are
overridden (so a BoxeR checkout picks up the B200 op without any edit); if it is
not, light-weight placeholder parents are created so the imports still work.
"""
from __future__ import annotations

import importlib
import sys
import types


def _parent(name: str) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        try:
            mod = importlib.import_module(name)
        except Exception:
            mod = types.ModuleType(name)
            mod.__path__ = []          # behaves as a package
            mod.__boxer_b200_placeholder__ = True
            sys.modules[name] = mod
    return mod


def install() -> None:
    from . import box_attention, box_attention_func, ops

    e2edet = _parent("e2edet")
    module = _parent("e2edet.module")
    setattr(e2edet, "module", module)

    sys.modules["e2edet.ops"] = ops
    setattr(e2edet, "ops", ops)

    ops_pkg = types.ModuleType("e2edet.module.ops")
    ops_pkg.__path__ = []
    ops_pkg.BoxAttnFunction = box_attention_func.BoxAttnFunction
    ops_pkg.InstanceAttnFunction = box_attention_func.InstanceAttnFunction
    ops_pkg.__all__ = ["BoxAttnFunction", "InstanceAttnFunction"]
    ops_pkg.box_attention_func = box_attention_func
    sys.modules["e2edet.module.ops"] = ops_pkg
    sys.modules["e2edet.module.ops.box_attention_func"] = box_attention_func
    setattr(module, "ops", ops_pkg)

    sys.modules["e2edet.module.box_attention"] = box_attention
    setattr(module, "box_attention", box_attention)


def uninstall() -> None:
    for name in ("e2edet.module.box_attention", "e2edet.module.ops.box_attention_func",
                 "e2edet.module.ops", "e2edet.ops"):
        sys.modules.pop(name, None)
    for name in ("e2edet.module", "e2edet"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__boxer_b200_placeholder__", False):
            sys.modules.pop(name, None)
