"""Loader / builder of the C-ABI library ``libboxattn_b200.so`` (include/boxattn_b200.h).

The library is built in-tree (``boxer_b200/_C/``) by ``build()`` with nvcc for
sm_100a only.  There is no CPU or PyTorch fallback: if the library is missing
or a symbol the header declares is absent, importing callers get an
``ImportError`` / ``RuntimeError`` -- never a silently different code path.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "_C")
# BOXER_B200_LIB: experiment hook (A/B builds of the same ABI); the default is the in-tree build
LIB_PATH = os.environ.get("BOXER_B200_LIB") or os.path.join(LIB_DIR, "libboxattn_b200.so")
HEADER = os.path.join(ROOT, "include", "boxattn_b200.h")
SOURCES = [os.path.join(CSRC, "boxattn_abi.cu")]
DEPENDS = SOURCES + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + [HEADER]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--compiler-options", "-fPIC",
]
# boxattn_abi.cu is compiled once per (dtype, direction) slice, in parallel, and the objects are linked into one
# library (the BXR_TU_* macros at the top of that file); compiled without -D it is one complete translation unit.
SLICES = [("common", 0, 0, 1)] + [(f"{dt}_{dn}", db, dd, 0)
                                  for dt, db in (("f32", 1), ("f64", 2), ("bf16", 4))
                                  for dn, dd in (("fwd", 1), ("bwd", 2))]

DTYPES = ("f32", "f64", "bf16")
OPS = ("box_attn_fwd", "box_attn_bwd", "instance_attn_fwd", "instance_attn_bwd")
FUSED_OPS = ("box_grid_attn_fwd", "box_grid_attn_bwd", "box_grid_softmax_attn_fwd", "box_grid_softmax_attn_bwd")
EXPORTS = (
    ["bxr_abi_version", "bxr_status_string", "bxr_last_error_detail", "bxr_last_launch_count",
     "bxr_attn_bwd_workspace_bytes", "bxr_box_grid_attn_workspace_bytes", "bxr_value_epilogue"]
    + [f"bxr_{op}_{dt}" for op in OPS + FUSED_OPS for dt in DTYPES]
    + [f"bxr_instance_weights_{d}_{dt}" for d in ("fwd", "bwd") for dt in ("f32", "f64")]
)

_lock = threading.RLock()      # re-entrant: load_shim() loads the C-ABI library under it
_lib = None


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libboxattn_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    if os.environ.get("BOXER_B200_LIB"):      # an A/B build carries its own -D flags: never rebuild it implicitly
        return False
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPENDS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into boxer_b200/_C/libboxattn_b200.so."""
    with _lock:
        if not force and not is_stale():
            return LIB_PATH
        os.makedirs(LIB_DIR, exist_ok=True)
        objdir = os.path.join(os.environ.get("TMPDIR", "/tmp"), "boxattn_b200_build" + ("_alt" if os.environ.get("BOXER_B200_LIB") else ""))   # objects stay out of the tree
        os.makedirs(objdir, exist_ok=True)
        nvcc = _nvcc()
        extra = os.environ.get("BOXER_B200_NVCC_EXTRA", "").split()      # experiment hook: -D tuning macros for A/B builds
        procs = []
        for name, dtypes, dirs, common in SLICES:
            obj = os.path.join(objdir, name + ".o")
            cmd = [nvcc, *NVCC_FLAGS, *extra, f"-DBXR_TU_DTYPES={dtypes}", f"-DBXR_TU_DIRS={dirs}", f"-DBXR_TU_COMMON={common}", f"-DBXR_TU_NAME={name}",
                   "-c", "-o", obj, *SOURCES]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        logs, failed = [], False
        for obj, proc in procs:
            out, _ = proc.communicate()
            logs.append(out)
            failed |= proc.returncode != 0
        if failed:
            raise RuntimeError("nvcc failed:\n" + "\n".join(logs))
        if verbose:
            print("\n".join(logs))
        tmp = LIB_PATH + ".tmp"
        # the soname lets the pybind shim's NEEDED entry resolve to the copy ctypes already loaded, wherever the tree lives
        link = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xlinker", "-soname=libboxattn_b200.so",
                               "-o", tmp, *[o for o, _ in procs]], capture_output=True, text=True)
        if link.returncode != 0:
            raise RuntimeError("link failed:\n" + link.stdout + link.stderr)
        os.replace(tmp, LIB_PATH)
        return LIB_PATH


SHIM_NAME = "_boxattn_torch"
SHIM_SOURCE = os.path.join(CSRC, "boxattn_torch.cpp")
SHIM_PATH = os.path.join(LIB_DIR, SHIM_NAME + ".so")
_shim = None
_shim_tried = False


def shim_is_stale() -> bool:
    if not os.path.exists(SHIM_PATH):
        return True
    t = os.path.getmtime(SHIM_PATH)
    return any(os.path.getmtime(d) > t for d in (SHIM_SOURCE, HEADER))


def build_torch_shim(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/boxattn_torch.cpp (the pybind fast path over the same C ABI) into boxer_b200/_C/_boxattn_torch.so.
    Host-only C++: needs g++ and the torch headers, no GPU.  The shim links against libboxattn_b200.so: load_shim() loads that
    library first (its soname satisfies the shim's NEEDED entry); an $ORIGIN rpath covers a direct import."""
    with _lock:
        if not force and not shim_is_stale():
            return SHIM_PATH
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("build libboxattn_b200.so first")
        import shutil as _sh
        from torch.utils.cpp_extension import load
        bdir = os.path.join(os.environ.get("TMPDIR", "/tmp"), "boxattn_b200_shim_build")
        os.makedirs(bdir, exist_ok=True)
        os.environ.setdefault("MAX_JOBS", "4")
        built = os.path.join(bdir, SHIM_NAME + ".so")
        if os.path.exists(built):
            os.remove(built)
        try:
            load(name=SHIM_NAME, sources=[SHIM_SOURCE], extra_include_paths=[os.path.join(ROOT, "include")],
                 extra_cflags=["-O2"], extra_ldflags=[f"-L{LIB_DIR}", "-l:libboxattn_b200.so", "-Wl,-rpath,'$$ORIGIN'"],
                 with_cuda=True, build_directory=bdir, is_python_module=False, verbose=verbose)
        except OSError:
            # cpp_extension tries to dlopen the result inside the build directory, where the $ORIGIN rpath does not
            # find libboxattn_b200.so yet; the file itself is complete
            if not os.path.exists(built):
                raise
        tmp = SHIM_PATH + ".tmp"
        _sh.copy2(os.path.join(bdir, SHIM_NAME + ".so"), tmp)
        os.replace(tmp, SHIM_PATH)
        return SHIM_PATH


def load_shim():
    """The pybind fast path, or None (then ops.py uses ctypes on the same C ABI).  Never used with BOXER_B200_LIB set:
    the shim is linked against the in-tree library."""
    global _shim, _shim_tried
    if _shim_tried:
        return _shim
    with _lock:
        if not _shim_tried:
            _shim_tried = True
            if os.environ.get("BOXER_B200_LIB") or os.environ.get("BOXER_B200_NO_SHIM") or not os.path.exists(SHIM_PATH):
                return None
            try:
                import importlib.util
                load()                      # the C-ABI library first (one instance per process)
                import torch  # noqa: F401
                spec = importlib.util.spec_from_file_location(SHIM_NAME, SHIM_PATH)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                if mod.abi_version() == 1:
                    _shim = mod
            except Exception:               # a shim built against another torch: the ctypes route still works
                _shim = None
    return _shim


def _declare(lib):
    c = ctypes
    lib.bxr_abi_version.restype = c.c_int
    lib.bxr_status_string.restype = c.c_char_p
    lib.bxr_status_string.argtypes = [c.c_int]
    lib.bxr_last_error_detail.restype = c.c_char_p
    lib.bxr_last_launch_count.restype = c.c_int
    lib.bxr_attn_bwd_workspace_bytes.restype = c.c_size_t
    lib.bxr_attn_bwd_workspace_bytes.argtypes = [c.c_int] * 5 + [c.c_uint]
    vp, i, u, sz = c.c_void_p, c.c_int, c.c_uint, c.c_size_t
    dims = [i] * 7
    for dt in DTYPES:
        f = getattr(lib, f"bxr_box_attn_fwd_{dt}")
        f.restype, f.argtypes = i, [vp] * 5 + dims + [vp, u, vp]
        f = getattr(lib, f"bxr_box_attn_bwd_{dt}")
        f.restype, f.argtypes = i, [vp] * 6 + dims + [vp] * 3 + [vp, sz, u, vp]
        f = getattr(lib, f"bxr_instance_attn_fwd_{dt}")
        f.restype, f.argtypes = i, [vp] * 6 + dims + [vp, vp, u, vp]
        f = getattr(lib, f"bxr_instance_attn_bwd_{dt}")
        f.restype, f.argtypes = i, [vp] * 8 + dims + [vp] * 4 + [vp, sz, u, vp]
        f = getattr(lib, f"bxr_box_grid_attn_fwd_{dt}")
        f.restype, f.argtypes = i, [vp] * 8 + dims + [vp] + [vp, sz, u, vp]
        f = getattr(lib, f"bxr_box_grid_attn_bwd_{dt}")
        f.restype, f.argtypes = i, [vp] * 9 + dims + [vp] * 4 + [vp, sz, u, vp]
        f = getattr(lib, f"bxr_box_grid_softmax_attn_fwd_{dt}")
        f.restype, f.argtypes = i, [vp] * 8 + dims + [vp, vp] + [vp, sz, u, vp]
        f = getattr(lib, f"bxr_box_grid_softmax_attn_bwd_{dt}")
        f.restype, f.argtypes = i, [vp] * 9 + dims + [vp] * 4 + [vp, sz, u, vp]
    for dt in ("f32", "f64"):
        f = getattr(lib, f"bxr_instance_weights_fwd_{dt}")
        f.restype, f.argtypes = i, [vp, c.c_longlong, i, i, vp, vp, vp]
        f = getattr(lib, f"bxr_instance_weights_bwd_{dt}")
        f.restype, f.argtypes = i, [vp, vp, vp, c.c_longlong, i, i, vp, vp]
    lib.bxr_value_epilogue.restype = i
    lib.bxr_value_epilogue.argtypes = [vp, i, vp, vp, i, c.c_longlong, i, vp]
    lib.bxr_box_grid_attn_workspace_bytes.restype = c.c_size_t
    lib.bxr_box_grid_attn_workspace_bytes.argtypes = [c.c_int] * 9 + [c.c_uint]
    return lib


def load():
    """dlopen the in-tree library; raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: the CUDA extension has not been built. "
                    "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                    "boxer_b200 has no CPU / PyTorch fallback."
                )
            lib = ctypes.CDLL(LIB_PATH)
            missing = [s for s in EXPORTS if not hasattr(lib, s)]
            if missing:
                raise ImportError(f"{LIB_PATH} does not export {missing}; rebuild it")
            if lib.bxr_abi_version() != 1:
                raise ImportError("libboxattn_b200.so ABI version mismatch; rebuild it")
            _lib = _declare(lib)
    return _lib


def check(status: int, what: str):
    if status != 0:
        lib = load()
        name = lib.bxr_status_string(status).decode()
        detail = lib.bxr_last_error_detail().decode()
        raise RuntimeError(f"{what} failed: {name}" + (f" ({detail})" if detail else ""))
