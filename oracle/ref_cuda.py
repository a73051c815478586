"""Loader of the UNMODIFIED reference CUDA extension built by oracle/build_ref.py.
TEST INFRASTRUCTURE ONLY (same-GPU comparison point; never on the product path)."""
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_mod = None


def available() -> bool:
    return bool(glob.glob(os.path.join(_HERE, "_ref", "e2edet_ref_ops*.so")))


def load():
    """The reference's pybind module (vision.cpp:7-12): box_attn_forward/backward,
    instance_attn_forward/backward taking at::Tensor."""
    global _mod
    if _mod is None:
        import torch  # noqa: F401  (libtorch must be loaded first)
        path = glob.glob(os.path.join(_HERE, "_ref", "e2edet_ref_ops*.so"))[0]
        spec = importlib.util.spec_from_file_location("e2edet_ref_ops", path)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod
