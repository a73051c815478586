"""Tensor-level entry points: the mirror of the reference's pybind module ``e2edet.ops``.

Same four functions, same argument order and return values as
``/root/reference/e2edet/module/ops/src/vision.cpp:7-12``:

    box_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step) -> output
    box_attn_backward(..., grad_output, im2col_step) -> [grad_value, grad_sampling_loc, grad_attn_weight]
    instance_attn_forward(value, shapes, lsi, loc, spatial_attn_weight, level_attn_weight, im2col_step) -> [output, mask_output]
    instance_attn_backward(..., grad_output, grad_mask_output, im2col_step) -> [grad_value, grad_loc, grad_spatial, grad_level]

PyTorch is used here only to allocate device memory and to name the stream; the
arithmetic is in ``libboxattn_b200.so`` (include/boxattn_b200.h) and nowhere
else -- there is no CPU implementation (the reference has none either:
``box_attn.h:53`` "Not implemented on the CPU") and no fallback.

Behavioural notes vs the reference host code (box_attn.cu:15-135):
* non-CUDA or non-contiguous inputs raise ``RuntimeError`` (reference: AT_ASSERTM);
* ``batch % min(batch, im2col_step) == 0`` is still required (box_attn.cu:40-42), but the
  chunk loop is gone: every image goes in one launch;
* kernel launch failures raise (the reference printf()s them, box_attn_kernel.cuh:1118-1122);
* shape mismatches between loc / weights / value raise instead of reading out of bounds;
* dtypes: float32, float64 and -- beyond the reference -- bfloat16 ``value`` (loc / weights
  stay float32; accumulation fp32).
"""
from __future__ import annotations

import torch

from . import _native

_DET_OVERRIDE = None          # None: follow torch.are_deterministic_algorithms_enabled()
FLAG_DETERMINISTIC = 0x1

FLAG_PATH_WINDOW = 0x2
FLAG_PATH_POINT = 0x4
FLAG_STAGED = 0x8
FLAG_PATH_TILE = 0x10
FLAG_NO_TILE_ORDER = 0x20
_PATH_FLAGS = 0               # tuning / testing override of the kernel family (see set_kernel_path)

_SUFFIX = {torch.float32: "f32", torch.float64: "f64", torch.bfloat16: "bf16"}


def set_deterministic(mode):
    """True / False: force the bit-reproducible (fixed-point) or the atomic grad_value scatter.
    None (default): follow ``torch.are_deterministic_algorithms_enabled()``."""
    global _DET_OVERRIDE
    _DET_OVERRIDE = mode


def deterministic() -> bool:
    if _DET_OVERRIDE is not None:
        return bool(_DET_OVERRIDE)
    return torch.are_deterministic_algorithms_enabled()


def set_kernel_path(path: str = "auto"):
    """"auto" (default): footprint-window kernels for large box-attention calls, point kernels otherwise;
    "window" / "point": force one family where it applies (A/B benchmarking and tests);
    "staged" / "window-staged": the window forward with TMA-staged row operands and a pooled multi-level window
    (experiment, boxattn_staged.cuh; measured no faster than the plain window kernels);
    "tile": the query-tile x value-tile kernels (boxattn_tile.cuh) whenever they apply (Nq == S, head_dim 32, P <= 16);
    "window-memory-order": the window kernels with work units in memory order (A/B of the tile-ordered units they use for
    self-attention-shaped calls)."""
    global _PATH_FLAGS
    _PATH_FLAGS = {"auto": 0, "window": FLAG_PATH_WINDOW, "point": FLAG_PATH_POINT, "staged": FLAG_STAGED,
                   "window-staged": FLAG_PATH_WINDOW | FLAG_STAGED, "tile": FLAG_PATH_TILE,
                   "window-memory-order": FLAG_NO_TILE_ORDER}[path]


def last_launch_count() -> int:
    """Kernels enqueued by this thread's last native call (bench.py's gpu_launches)."""
    return int(_native.load().bxr_last_launch_count())


# ------------------------------------------------------------------ checks
def _check_input(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


_GEOMETRY_NAMES = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight[0]", "attn_weight[1]")


def _geometry(value, shapes, lsi, loc, weights):
    tensors = (value, shapes, lsi, loc, *weights)
    try:        # one pass on the hot path; the per-tensor diagnosis only when something is off
        ok = all(t.is_cuda and t.is_contiguous() for t in tensors)
    except AttributeError:
        ok = False
    if not ok:
        for t, n in zip(tensors, _GEOMETRY_NAMES):
            _check_input(t, n)
    if value.dim() != 4:
        raise RuntimeError(f"value must be (B, S, heads, head_dim), got {tuple(value.shape)}")
    if shapes.dtype != torch.int64 or lsi.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64")
    if shapes.dim() != 2 or shapes.shape[1] != 2:
        raise RuntimeError(f"spatial_shapes must be (L, 2), got {tuple(shapes.shape)}")
    B, S, H, D = value.shape
    L = shapes.shape[0]
    if lsi.numel() != L:
        raise RuntimeError("level_start_index must have one entry per level")
    if loc.dim() != 6 or loc.shape[0] != B or loc.shape[2] != H or loc.shape[3] != L or loc.shape[5] != 2:
        raise RuntimeError(
            f"sampling_loc must be (B={B}, Nq, heads={H}, L={L}, P, 2), got {tuple(loc.shape)}")
    Nq, P = loc.shape[1], loc.shape[4]
    for w in weights:
        if w.numel() != B * Nq * H * L * P:
            raise RuntimeError(
                f"attention weights must hold B*Nq*heads*L*P = {B * Nq * H * L * P} elements, got {tuple(w.shape)}")
    if L > 32:
        raise RuntimeError("at most 32 levels are supported")
    dev = value.device
    for t in tensors:
        if t.device != dev:
            raise RuntimeError(f"all tensors must be on the same device, got {sorted({str(x.device) for x in tensors})}")
    return B, S, H, D, L, Nq, P


def _dtypes(value, loc, weights):
    """(suffix, TW dtype).  value decides; loc/weights must already be the matching TW."""
    suf = _SUFFIX.get(value.dtype)
    if suf is None:
        raise RuntimeError(f"box attention supports float32, float64 and bfloat16 value, got {value.dtype}")
    tw = torch.float64 if value.dtype == torch.float64 else torch.float32
    for t in (loc, *weights):
        if t.dtype != tw:
            raise RuntimeError(f"with {value.dtype} value, sampling_loc and weights must be {tw}, got {t.dtype}")
    return suf, tw


def _step_check(B, im2col_step):
    step = min(B, int(im2col_step))
    if B > 0 and (step <= 0 or B % step != 0):
        raise RuntimeError(f"batch({B}) must divide im2col_step({step})")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(dev):
    # decoder-sized calls are launch-latency bound: the raw-handle query is ~20x cheaper than building a Stream object
    if _raw_stream is not None:
        return _raw_stream(dev.index if dev.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(dev).cuda_stream


class _on_device:
    """Device guard that costs nothing when the tensors already live on the current device
    (the common case: one process per GPU); the reference has no guard at all."""
    __slots__ = ("dev", "prev")

    def __init__(self, dev):
        self.dev = dev

    def __enter__(self):
        idx = self.dev.index
        self.prev = torch.cuda.current_device()
        if idx is not None and idx != self.prev:
            torch.cuda.set_device(idx)
        else:
            self.prev = None

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


_fn_cache = {}


def _fn(name):
    f = _fn_cache.get(name)
    if f is None:
        f = _fn_cache[name] = getattr(_native.load(), name)
    return f


def _workspace(lib, value, flags):
    B, S, H, D = value.shape
    n = lib.bxr_attn_bwd_workspace_bytes(value.element_size(), B, S, H, D, flags)
    if not n:
        return None, 0
    return torch.empty(n, dtype=torch.uint8, device=value.device), n


# ------------------------------------------------------------------ box op
# Fast path: the pybind shim (csrc/boxattn_torch.cpp) validates, allocates and calls the same C-ABI entry points in C++
# (~6 us per call instead of ~15: decoder-sized calls are launch-latency bound).  It returns None for anything but the
# plain case, and then -- or when the shim is not built -- the Python route below runs, with its diagnostics.
_SHIM = None
_SHIM_READY = False


def _shim():
    """Resolved on the first op call (importing the package must not dlopen anything)."""
    global _SHIM, _SHIM_READY
    if not _SHIM_READY:
        _SHIM = _native.load_shim()
        _SHIM_READY = True
    return _SHIM


def _bwd_flags():
    return (FLAG_DETERMINISTIC if deterministic() else 0) | _PATH_FLAGS


def box_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=64):
    if (_SHIM if _SHIM_READY else _shim()) is not None:
        try:
            r = _SHIM.box_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step, _PATH_FLAGS)
        except TypeError:
            r = None
        if r is not None:
            return r
    B, S, H, D, L, Nq, P = _geometry(value, spatial_shapes, level_start_index, sampling_loc, (attn_weight,))
    suf, _ = _dtypes(value, sampling_loc, (attn_weight,))
    _step_check(B, im2col_step)
    lib = _native.load()
    out = torch.empty((B, Nq, H * D), dtype=value.dtype, device=value.device)
    with _on_device(value.device):
        st = _fn(f"bxr_box_attn_fwd_{suf}")(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), B, S, H, D, L, Nq, P,
            out.data_ptr(), _PATH_FLAGS, _stream(value.device))
    _native.check(st, "box_attn_forward")
    return out


def box_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step=64):
    if (_SHIM if _SHIM_READY else _shim()) is not None:
        try:
            r = _SHIM.box_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                        im2col_step, _bwd_flags())
        except TypeError:
            r = None
        if r is not None:
            return r
    B, S, H, D, L, Nq, P = _geometry(value, spatial_shapes, level_start_index, sampling_loc, (attn_weight,))
    suf, _ = _dtypes(value, sampling_loc, (attn_weight,))
    _check_input(grad_output, "grad_output")
    if grad_output.dtype != value.dtype or grad_output.numel() != B * Nq * H * D:
        raise RuntimeError("grad_output must match the forward output's dtype and size")
    _step_check(B, im2col_step)
    lib = _native.load()
    flags = (FLAG_DETERMINISTIC if deterministic() else 0) | _PATH_FLAGS
    grad_value = torch.empty_like(value)
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    with _on_device(value.device):
        ws, ws_bytes = _workspace(lib, value, flags)
        st = _fn(f"bxr_box_attn_bwd_{suf}")(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(),
            B, S, H, D, L, Nq, P,
            grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
            ws.data_ptr() if ws is not None else None, ws_bytes, flags, _stream(value.device))
    _native.check(st, "box_attn_backward")
    return [grad_value, grad_loc, grad_attn]


# ------------------------------------------------------------- instance op
def instance_attn_forward(value, spatial_shapes, level_start_index, sampling_loc,
                          spatial_attn_weight, level_attn_weight, im2col_step=64):
    if (_SHIM if _SHIM_READY else _shim()) is not None:
        try:
            r = _SHIM.instance_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, spatial_attn_weight,
                                            level_attn_weight, im2col_step, _PATH_FLAGS)
        except TypeError:
            r = None
        if r is not None:
            return r
    ws_ = (spatial_attn_weight, level_attn_weight)
    B, S, H, D, L, Nq, P = _geometry(value, spatial_shapes, level_start_index, sampling_loc, ws_)
    suf, _ = _dtypes(value, sampling_loc, ws_)
    _step_check(B, im2col_step)
    lib = _native.load()
    out = torch.empty((B, Nq, H * D), dtype=value.dtype, device=value.device)
    mask_out = torch.empty((B, Nq, P, H * D), dtype=value.dtype, device=value.device)
    with _on_device(value.device):
        st = _fn(f"bxr_instance_attn_fwd_{suf}")(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), spatial_attn_weight.data_ptr(), level_attn_weight.data_ptr(),
            B, S, H, D, L, Nq, P, out.data_ptr(), mask_out.data_ptr(), _PATH_FLAGS, _stream(value.device))
    _native.check(st, "instance_attn_forward")
    return [out, mask_out]


def instance_attn_backward(value, spatial_shapes, level_start_index, sampling_loc,
                           spatial_attn_weight, level_attn_weight, grad_output, grad_mask_output, im2col_step=64):
    if (_SHIM if _SHIM_READY else _shim()) is not None:
        try:
            r = _SHIM.instance_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, spatial_attn_weight,
                                             level_attn_weight, grad_output, grad_mask_output, im2col_step, _bwd_flags())
        except TypeError:
            r = None
        if r is not None:
            return r
    ws_ = (spatial_attn_weight, level_attn_weight)
    B, S, H, D, L, Nq, P = _geometry(value, spatial_shapes, level_start_index, sampling_loc, ws_)
    suf, _ = _dtypes(value, sampling_loc, ws_)
    _check_input(grad_output, "grad_output")
    _check_input(grad_mask_output, "grad_mask_output")
    if grad_output.dtype != value.dtype or grad_output.numel() != B * Nq * H * D:
        raise RuntimeError("grad_output must match the forward output's dtype and size")
    if grad_mask_output.dtype != value.dtype or grad_mask_output.numel() != B * Nq * P * H * D:
        raise RuntimeError("grad_mask_output must match the forward mask output's dtype and size")
    _step_check(B, im2col_step)
    lib = _native.load()
    flags = (FLAG_DETERMINISTIC if deterministic() else 0) | _PATH_FLAGS
    grad_value = torch.empty_like(value)
    grad_loc = torch.empty_like(sampling_loc)
    grad_sw = torch.empty_like(spatial_attn_weight)
    grad_lw = torch.empty_like(level_attn_weight)
    with _on_device(value.device):
        ws, ws_bytes = _workspace(lib, value, flags)
        st = _fn(f"bxr_instance_attn_bwd_{suf}")(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
            sampling_loc.data_ptr(), spatial_attn_weight.data_ptr(), level_attn_weight.data_ptr(),
            grad_output.data_ptr(), grad_mask_output.data_ptr(),
            B, S, H, D, L, Nq, P,
            grad_value.data_ptr(), grad_loc.data_ptr(), grad_sw.data_ptr(), grad_lw.data_ptr(),
            ws.data_ptr() if ws is not None else None, ws_bytes, flags, _stream(value.device))
    _native.check(st, "instance_attn_backward")
    return [grad_value, grad_loc, grad_sw, grad_lw]


# ------------------------------------------------- fused box -> grid -> attention
def _fused_geometry(value, shapes, lsi, boxes, angles, valid_ratios, kernel_indices, attn):
    for t, n in ((value, "value"), (shapes, "spatial_shapes"), (lsi, "level_start_index"), (boxes, "boxes"),
                 (kernel_indices, "kernel_indices"), (attn, "attn_weight")):
        _check_input(t, n)
    if value.dim() != 4:
        raise RuntimeError(f"value must be (B, S, heads, head_dim), got {tuple(value.shape)}")
    if shapes.dtype != torch.int64 or lsi.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64")
    B, S, H, D = value.shape
    L = shapes.shape[0]
    if boxes.dim() != 5 or boxes.shape[0] != B or boxes.shape[2] != H or boxes.shape[3] != L or boxes.shape[4] != 4:
        raise RuntimeError(f"boxes must be (B={B}, Nq, heads={H}, L={L}, 4), got {tuple(boxes.shape)}")
    Nq = boxes.shape[1]
    if kernel_indices.dim() != 2 or kernel_indices.shape[1] != 2:
        raise RuntimeError(f"kernel_indices must be (P, 2), got {tuple(kernel_indices.shape)}")
    P = kernel_indices.shape[0]
    if attn.numel() != B * Nq * H * L * P:
        raise RuntimeError(f"attention weights must hold B*Nq*heads*L*P = {B * Nq * H * L * P} elements, got {tuple(attn.shape)}")
    if angles is not None:
        _check_input(angles, "angles")
        if angles.numel() != B * Nq * H * L:
            raise RuntimeError(f"angles must hold B*Nq*heads*L elements, got {tuple(angles.shape)}")
    if valid_ratios is not None:
        _check_input(valid_ratios, "valid_ratios")
        if valid_ratios.numel() != B * L * 2:
            raise RuntimeError(f"valid_ratios must hold B*L*2 elements, got {tuple(valid_ratios.shape)}")
    if L > 32 or lsi.numel() != L:
        raise RuntimeError("level_start_index must have one entry per level (at most 32 levels)")
    return B, S, H, D, L, Nq, P


def _aligned16(t):
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


def _aligned16_opt(t):
    return None if t is None else _aligned16(t)


def box_grid_attn_forward(value, spatial_shapes, level_start_index, boxes, angles, valid_ratios, kernel_indices,
                          attn_weight, im2col_step=64, softmax=False):
    """out = box_attn_forward(value, ..., grid(boxes, angles, valid_ratios, kernel_indices), attn_weight) with the
    K x K grid of BoxAttention._where_to_attend (box_attention.py:196-214; rotated: :304-338) generated in-kernel.

    softmax=True: ``attn_weight`` holds the attention LOGITS; their softmax over a row's L*P
    points (box_attention.py:227-231) is taken in the kernel and returned too -> ``(out, attention_weights)``."""
    B, S, H, D, L, Nq, P = _fused_geometry(value, spatial_shapes, level_start_index, boxes, angles, valid_ratios,
                                           kernel_indices, attn_weight)
    opt = tuple(t for t in (angles, valid_ratios) if t is not None)
    suf, _ = _dtypes(value, boxes, (attn_weight, kernel_indices, *opt))
    _step_check(B, im2col_step)
    lib = _native.load()
    # the workspace query below assumes the fused kernels' alignment requirements hold (boxattn_abi.cu fused_forward):
    # repair odd-offset views of the small operands here instead of sizing a location workspace for every call
    value, boxes = _aligned16(value), _aligned16(boxes)
    kernel_indices, valid_ratios, angles = _aligned16(kernel_indices), _aligned16_opt(valid_ratios), _aligned16_opt(angles)
    out = torch.empty((B, Nq, H * D), dtype=value.dtype, device=value.device)
    attn_out = torch.empty_like(attn_weight) if softmax else None
    with _on_device(value.device):
        n = lib.bxr_box_grid_attn_workspace_bytes(value.element_size(), 0, B, S, H, D, L, Nq, P, _PATH_FLAGS)
        ws = torch.empty(n, dtype=torch.uint8, device=value.device) if n else None
        head = (value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), boxes.data_ptr(),
                angles.data_ptr() if angles is not None else None,
                valid_ratios.data_ptr() if valid_ratios is not None else None,
                kernel_indices.data_ptr(), attn_weight.data_ptr(), B, S, H, D, L, Nq, P, out.data_ptr())
        tail = (ws.data_ptr() if ws is not None else None, n, _PATH_FLAGS, _stream(value.device))
        if softmax:
            st = _fn(f"bxr_box_grid_softmax_attn_fwd_{suf}")(*head, attn_out.data_ptr(), *tail)
        else:
            st = _fn(f"bxr_box_grid_attn_fwd_{suf}")(*head, *tail)
    _native.check(st, "box_grid_attn_forward")
    return (out, attn_out) if softmax else out


def box_grid_attn_backward(value, spatial_shapes, level_start_index, boxes, angles, valid_ratios, kernel_indices,
                           attn_weight, grad_output, im2col_step=64, softmax=False):
    """-> [grad_value, grad_boxes, grad_angles | None, grad_attn_weight]
    softmax=True: ``attn_weight`` are the softmax weights the softmax forward returned; the last gradient is the
    gradient of the LOGITS."""
    B, S, H, D, L, Nq, P = _fused_geometry(value, spatial_shapes, level_start_index, boxes, angles, valid_ratios,
                                           kernel_indices, attn_weight)
    opt = tuple(t for t in (angles, valid_ratios) if t is not None)
    suf, _ = _dtypes(value, boxes, (attn_weight, kernel_indices, *opt))
    _check_input(grad_output, "grad_output")
    if grad_output.dtype != value.dtype or grad_output.numel() != B * Nq * H * D:
        raise RuntimeError("grad_output must match the forward output's dtype and size")
    _step_check(B, im2col_step)
    lib = _native.load()
    flags = (FLAG_DETERMINISTIC if deterministic() else 0) | _PATH_FLAGS
    value, boxes, grad_output = _aligned16(value), _aligned16(boxes), _aligned16(grad_output)
    kernel_indices, valid_ratios, angles = _aligned16(kernel_indices), _aligned16_opt(valid_ratios), _aligned16_opt(angles)
    grad_value = torch.empty_like(value)
    grad_boxes = torch.empty_like(boxes)
    grad_angles = torch.empty_like(angles) if angles is not None else None
    grad_attn = torch.empty_like(attn_weight)
    with _on_device(value.device):
        n = lib.bxr_box_grid_attn_workspace_bytes(value.element_size(), 1, B, S, H, D, L, Nq, P, flags)
        ws = torch.empty(n, dtype=torch.uint8, device=value.device) if n else None
        st = _fn(f"bxr_box_grid_softmax_attn_bwd_{suf}" if softmax else f"bxr_box_grid_attn_bwd_{suf}")(
            value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), boxes.data_ptr(),
            angles.data_ptr() if angles is not None else None,
            valid_ratios.data_ptr() if valid_ratios is not None else None,
            kernel_indices.data_ptr(), attn_weight.data_ptr(), grad_output.data_ptr(), B, S, H, D, L, Nq, P,
            grad_value.data_ptr(), grad_boxes.data_ptr(), grad_angles.data_ptr() if grad_angles is not None else None,
            grad_attn.data_ptr(), ws.data_ptr() if ws is not None else None, n, flags, _stream(value.device))
    _native.check(st, "box_grid_attn_backward")
    return [grad_value, grad_boxes, grad_angles, grad_attn]


# ============================================================== InstanceAttention weights from the 2x2 logit maps
def _instw_geometry(logits, K):
    _check_input(logits, "logits")
    if logits.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f"instance weights support float32 and float64 logits, got {logits.dtype}")
    if logits.dim() != 6 or logits.shape[-1] != 2 or logits.shape[-2] != 2:
        raise RuntimeError(f"logits must be (B, Nq, heads, L, 2, 2), got {tuple(logits.shape)}")
    K = int(K)
    if K < 0 or K % 2:
        raise RuntimeError(f"kernel size must be even (the 2x2 map is upsampled by K/2), got {K}")
    B, Nq, H, L = logits.shape[:4]
    if L > 32:
        raise RuntimeError("at most 32 levels are supported")
    return B * Nq * H, L, K, "f32" if logits.dtype == torch.float32 else "f64"


def instance_weights_forward(logits, K):
    """(spatial_w, level_w), each (B,Nq,H,L,K,K), from InstanceAttention's (B,Nq,H,L,2,2) logit maps:
    nearest-upsample to K x K, softmax over (L,K,K) / over L (box_attention.py:93-110) -- one kernel."""
    rows, L, K, suf = _instw_geometry(logits, K)
    shape = tuple(logits.shape[:4]) + (K, K)
    sw = torch.empty(shape, dtype=logits.dtype, device=logits.device)
    lw = torch.empty_like(sw)
    with _on_device(logits.device):
        st = _fn(f"bxr_instance_weights_fwd_{suf}")(logits.data_ptr(), rows, L, K, sw.data_ptr(), lw.data_ptr(),
                                                    _stream(logits.device))
    _native.check(st, "instance_weights_forward")
    return sw, lw


def instance_weights_backward(logits, grad_spatial_w, grad_level_w, K):
    """gradient of the logits given the gradients of both weight tensors -- one kernel."""
    rows, L, K, suf = _instw_geometry(logits, K)
    for g, n in ((grad_spatial_w, "grad_spatial_w"), (grad_level_w, "grad_level_w")):
        _check_input(g, n)
        if g.dtype != logits.dtype or g.numel() != rows * L * K * K or g.device != logits.device:
            raise RuntimeError(f"{n} must match the weights' dtype, size and device")
    grad = torch.empty_like(logits)
    with _on_device(logits.device):
        st = _fn(f"bxr_instance_weights_bwd_{suf}")(logits.data_ptr(), grad_spatial_w.data_ptr(), grad_level_w.data_ptr(),
                                                    rows, L, K, grad.data_ptr(), _stream(logits.device))
    _native.check(st, "instance_weights_backward")
    return grad


# ============================================================== value_proj epilogue: mask fill + storage cast
def value_epilogue(value, mask, out_dtype):
    """``value.masked_fill(mask[..., None], 0).to(out_dtype)`` in one pass (box_attention.py:222-225 + the cast the
    bf16 ops want).  value (..., C) float32 / bfloat16; mask (...) bool or None."""
    _check_input(value, "value")
    if value.dtype not in (torch.float32, torch.bfloat16) or out_dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("value epilogue supports float32 and bfloat16")
    C = value.shape[-1] if value.dim() else 1
    rows = value.numel() // C if C else 0
    if mask is not None:
        _check_input(mask, "mask")
        if mask.dtype != torch.bool or mask.numel() != rows or mask.device != value.device:
            raise RuntimeError("mask must be a bool tensor with one entry per pixel, on value's device")
    out = torch.empty(value.shape, dtype=out_dtype, device=value.device)
    with _on_device(value.device):
        st = _fn("bxr_value_epilogue")(value.data_ptr(), value.element_size(), mask.data_ptr() if mask is not None else None,
                                       out.data_ptr(), out.element_size(), rows, C, _stream(value.device))
    _native.check(st, "value_epilogue")
    return out
