import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, loudly
    try:
        from boxer_b200 import _native
        if _native.is_stale():
            print("[conftest] building libboxattn_b200.so with nvcc (sm_100a) ...", flush=True)
            _native.build()
    except Exception as e:  # the tests that need the library will report it
        print(f"[conftest] could not build the CUDA library: {e}", flush=True)


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when no device is visible, so a bare
    # `pytest tests/` on the CPU container stays green.
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
