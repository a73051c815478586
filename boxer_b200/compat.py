"""Make the reference's import paths resolve to this package.

BoxeR's transformer layers do (box_transformer.py:6, box3d_transformer.py:6)

    from .box_attention import BoxAttention, InstanceAttention        # e2edet.module.box_attention
    from .box_attention import Box3dAttention

its autograd layer does ``from e2edet import ops`` (box_attention_func.py:6) and its
tests ``from e2edet.module.ops import BoxAttnFunction`` (tests/box_attn_test.py:5).
``install()`` registers modules under exactly those names in ``sys.modules``:

* ``e2edet.ops``                   -> boxer_b200.ops            (the pybind module's four functions)
* ``e2edet.module.ops``            -> BoxAttnFunction, InstanceAttnFunction
* ``e2edet.module.ops.box_attention_func`` -> same
* ``e2edet.module.box_attention``  -> BoxAttention, InstanceAttention, Box3dAttention

How it binds with a real BoxeR checkout on ``sys.path``:

* **install() before ``import e2edet``** (the intended order).  Only the four leaf names are
  seeded; ``e2edet`` itself is NOT imported here.  When the checkout's ``e2edet/__init__`` later
  runs (it imports model -> module -> transformer -> box_transformer), every
  ``from .box_attention import ...`` / ``from e2edet import ops`` finds the seeded leaf in
  ``sys.modules`` first -- the import system consults ``sys.modules`` by full name before any
  finder -- so the reference's layers are built from this package's classes and never load the
  pybind extension.
* **install() after ``import e2edet``**: the reference modules have already bound the
  reference classes by value.  ``install()`` then also re-binds those names on every loaded
  ``e2edet.module.*`` module that holds them (box_transformer, box3d_transformer, transformer,
  ...) and the ``ops`` attribute of a loaded reference ``box_attention_func``.  Layers that were
  *instantiated* before the call keep their old attention objects -- build the model after.
* **no ``e2edet`` importable**: bare placeholder parents are created so the imports still work.
* ``install(lightweight=True)``: a checkout is present but its ``e2edet/__init__`` should not run
  (it pulls in the trainer, datasets, omegaconf, pycocotools ...): ``e2edet``, ``e2edet.module``
  and ``e2edet.utils`` are registered as bare packages that carry the checkout's real
  ``__path__``, so ``from e2edet.module.box_transformer import BoxTransformer`` imports just the
  files it names.

Import errors of a real checkout are never swallowed into an empty placeholder.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys
import types

_LEAVES = ("e2edet.module.box_attention", "e2edet.module.ops.box_attention_func", "e2edet.module.ops", "e2edet.ops")
_REBIND = ("BoxAttention", "InstanceAttention", "Box3dAttention", "BoxAttnFunction", "InstanceAttnFunction")
_saved: dict = {}


def _real_root():
    """Directory of an importable e2edet package (regular or namespace), found WITHOUT importing it."""
    mod = sys.modules.get("e2edet")
    if mod is not None and not getattr(mod, "__boxer_b200_placeholder__", False):
        paths = list(getattr(mod, "__path__", []) or [])
        return paths[0] if paths else None
    try:
        spec = importlib.util.find_spec("e2edet")      # top level: runs no package code
    except (ImportError, ValueError):
        spec = None
    if spec is None or not spec.submodule_search_locations:
        return None
    return list(spec.submodule_search_locations)[0]


def _bare_package(name: str, path: str | None, placeholder: bool) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = [path] if path else []
    mod.__package__ = name
    mod.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    mod.__spec__.submodule_search_locations = mod.__path__
    mod.__boxer_b200_placeholder__ = True
    mod.__boxer_b200_bare__ = not placeholder
    sys.modules[name] = mod
    return mod


def install(lightweight: bool = False) -> None:
    from . import box_attention, box_attention_func, ops

    ops_pkg = types.ModuleType("e2edet.module.ops")
    ops_pkg.__path__ = []
    ops_pkg.BoxAttnFunction = box_attention_func.BoxAttnFunction
    ops_pkg.InstanceAttnFunction = box_attention_func.InstanceAttnFunction
    ops_pkg.__all__ = ["BoxAttnFunction", "InstanceAttnFunction"]
    ops_pkg.box_attention_func = box_attention_func
    leaves = {
        "e2edet.ops": ops,
        "e2edet.module.ops": ops_pkg,
        "e2edet.module.ops.box_attention_func": box_attention_func,
        "e2edet.module.box_attention": box_attention,
    }

    # modules of a checkout that was imported before us: they bound the reference's classes by value
    already = {n: m for n, m in sys.modules.items()
               if (n == "e2edet" or n.startswith("e2edet.")) and m is not None and n not in leaves
               and not getattr(m, "__boxer_b200_placeholder__", False)}
    ref_func = sys.modules.get("e2edet.module.ops.box_attention_func")
    if ref_func is not None and ref_func is not box_attention_func and hasattr(ref_func, "ops"):
        _saved.setdefault(("attr", "e2edet.module.ops.box_attention_func", "ops"), (ref_func, ref_func.ops))
        ref_func.ops = ops      # reference Functions someone still holds now call into this library

    for name, mod in leaves.items():
        prev = sys.modules.get(name)
        if prev is not None and prev is not mod:
            _saved.setdefault(("module", name), prev)
        sys.modules[name] = mod

    root = _real_root()
    if root is None:
        # no checkout anywhere: placeholder parents so that the reference's import statements work
        for name in ("e2edet", "e2edet.module"):
            if name not in sys.modules:
                _bare_package(name, None, placeholder=True)
    elif lightweight:
        for name, sub in (("e2edet", ""), ("e2edet.module", "module"), ("e2edet.utils", "utils")):
            if name not in sys.modules:
                _bare_package(name, os.path.join(root, sub) if sub else root, placeholder=False)
    # else: a real checkout, not imported yet -- its own __init__ files run when the user imports it, and every
    # `from .box_attention import ...` inside them resolves to the leaves seeded above

    for pname, attr, mod in (("e2edet", "ops", ops), ("e2edet.module", "ops", ops_pkg),
                             ("e2edet.module", "box_attention", box_attention), ("e2edet", "module", None)):
        parent = sys.modules.get(pname)
        if parent is None:
            continue
        if mod is None:
            mod = sys.modules.get("e2edet.module")
            if mod is None:
                continue
        setattr(parent, attr, mod)

    ours = {n: getattr(box_attention, n, None) or getattr(box_attention_func, n) for n in _REBIND}
    for mname, m in already.items():
        for n, cls in ours.items():
            cur = m.__dict__.get(n)
            if isinstance(cur, type) and cur is not cls:
                _saved.setdefault(("attr", mname, n), (m, cur))
                setattr(m, n, cls)


def uninstall() -> None:
    for key, val in list(_saved.items()):
        if key[0] == "attr":
            obj, old = val
            setattr(obj, key[2], old)
    for name in _LEAVES:
        prev = _saved.get(("module", name))
        if prev is not None:
            sys.modules[name] = prev
        else:
            sys.modules.pop(name, None)
    _saved.clear()
    for name in ("e2edet.utils", "e2edet.module", "e2edet"):
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "__boxer_b200_placeholder__", False):
            sys.modules.pop(name, None)
            if getattr(mod, "__boxer_b200_bare__", False):      # drop what was imported through the bare parents
                for sub in [n for n in sys.modules if n.startswith(name + ".")]:
                    sys.modules.pop(sub, None)
