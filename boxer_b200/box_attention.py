"""``BoxAttention`` / ``InstanceAttention`` / ``Box3dAttention``: the module surface.

Mirror of ``/root/reference/e2edet/module/box_attention.py:10-363`` so that
BoxeR-2D / BoxeR-3D load it unchanged: same constructor arguments, same
parameter / buffer names and shapes (``linear_box_weight``, ``linear_box_bias``,
``linear_attn_weight``, ``linear_attn_bias``, ``value_proj.*``, ``out_proj.*``,
buffer ``kernel_indices``) -- released checkpoints load with ``strict=True`` --
same initialisation (``_reset_parameters``), same ``forward`` signature and return
tuples.  The dense projections stay PyTorch (cuBLAS); the gather-reduce goes to
the native op through ``BoxAttnFunction`` / ``InstanceAttnFunction``.

The three classes share one implementation here (the reference repeats it three
times); the differences are data:

=================  ==========  =====================  ==============================
class              box params  kernel_indices scale   attention logits
=================  ==========  =====================  ==============================
BoxAttention       4           1 / kernel_size        H*L*K*K, softmax over (L,K,K)
InstanceAttention  4           1 / kernel_size        H*L*2*2 repeated to K*K
Box3dAttention     4 or 5      1 / 2   (:291)         H*L*K*K, softmax over (L,K,K)
=================  ==========  =====================  ==============================
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .box_attention_func import (BoxAttnBf16Function, BoxAttnFunction, BoxGridAttnBf16Function, BoxGridAttnFunction,
                                 BoxGridSoftmaxAttnBf16Function, BoxGridSoftmaxAttnFunction,
                                 InstanceAttnBf16Function, InstanceAttnFunction, InstanceWeightsFunction,
                                 ValueEpilogueFunction)

_AMP_NATIVE = False
_FUSED_GRID = False
_FUSED_SOFTMAX = False


def set_fused_softmax(flag: bool):
    """With ``set_fused_grid(True)``: ``BoxAttention`` / ``Box3dAttention`` also hand the attention LOGITS to the
    op, which takes the softmax over a row's L*P points in its prologue and chains its gradient in its backward
    (SURVEY.md 8 row f2).  The modules still return the attention weights (written by the kernel); they are
    not differentiable through the module in this mode (BoxeR's callers discard them, box_transformer.py:346-354)."""
    global _FUSED_SOFTMAX
    _FUSED_SOFTMAX = bool(flag)


def set_fused_grid(flag: bool):
    """Opt in to the fused box -> grid -> attention op for ``BoxAttention`` / ``Box3dAttention`` (and
    ``InstanceAttention`` when ``inferencing``): the modules hand the boxes (+ angles) to the op instead of the
    materialised (B,Nq,H,L,P,2) sampling grid of ``_where_to_attend``.  Same outputs and gradients
    (tests/test_gpu_fused.py); off by default so that the default path is the reference's, op for op."""
    global _FUSED_GRID
    _FUSED_GRID = bool(flag)


def set_amp_native(flag: bool):
    """Opt in to bf16 value / outputs under ``torch.autocast`` instead of the reference's
    force-cast to fp32 (box_attention_func.py:11).  Off by default (reference behaviour)."""
    global _AMP_NATIVE
    _AMP_NATIVE = bool(flag)


def _use_bf16(value):
    return _AMP_NATIVE and value.is_cuda and (torch.is_autocast_enabled() or value.dtype == torch.bfloat16)


def _box_attn(value, v_shape, v_start_index, grid, weights, im2col_step):
    fn = BoxAttnBf16Function if _use_bf16(value) else BoxAttnFunction
    return fn.apply(value, v_shape, v_start_index, grid, weights, im2col_step)


def _box_grid_attn(value, v_shape, v_start_index, boxes, angles, valid_ratios, kernel_indices, weights, im2col_step):
    fn = BoxGridAttnBf16Function if _use_bf16(value) else BoxGridAttnFunction
    return fn.apply(value, v_shape, v_start_index, boxes, angles, valid_ratios, kernel_indices, weights, im2col_step)


def _box_grid_softmax_attn(value, v_shape, v_start_index, boxes, angles, valid_ratios, kernel_indices, logits, im2col_step):
    fn = BoxGridSoftmaxAttnBf16Function if _use_bf16(value) else BoxGridSoftmaxAttnFunction
    return fn.apply(value, v_shape, v_start_index, boxes, angles, valid_ratios, kernel_indices, logits, im2col_step)


def _instance_attn(value, v_shape, v_start_index, grid, sw, lw, k, im2col_step):
    fn = InstanceAttnBf16Function if _use_bf16(value) else InstanceAttnFunction
    return fn.apply(value, v_shape, v_start_index, grid, sw, lw, k, im2col_step)


def _kernel_offsets(kernel_size: int, divisor: float) -> torch.Tensor:
    """(K*K, 2) grid of (x, y) offsets centred on 0 with unit pitch, divided by ``divisor``
    (box_attention.py:38-51)."""
    half = kernel_size / 2.0
    if kernel_size % 2 == 0:
        ticks = torch.linspace(-half + 0.5, half - 0.5, kernel_size)
    else:
        r = (kernel_size - 1) // 2
        ticks = torch.linspace(-r, r, kernel_size)
    yy, xx = torch.meshgrid(ticks, ticks, indexing="ij")
    return torch.stack([xx, yy], dim=-1).reshape(-1, 2) / divisor


class _BoxAttentionBase(nn.Module):
    _num_box_variable = 4

    def _build(self, d_model, num_level, num_head, kernel_size, n_attn_logits, index_divisor):
        assert d_model % num_head == 0, "d_model should be divided by num_head"
        self.im2col_step = 64
        self.d_model = d_model
        self.num_head = num_head
        self.num_level = num_level
        self.head_dim = d_model // num_head
        self.kernel_size = kernel_size

        nv = self._num_box_variable
        self.linear_box_weight = nn.Parameter(torch.zeros(num_level * num_head * nv, d_model))
        self.linear_box_bias = nn.Parameter(torch.zeros(num_head * num_level * nv))
        self.linear_attn_weight = nn.Parameter(torch.zeros(num_head * num_level * n_attn_logits, d_model))
        self.linear_attn_bias = nn.Parameter(torch.zeros(num_head * num_level * n_attn_logits))
        self.value_proj = nn.Linear(d_model, d_model)
        self.out_proj = nn.Linear(d_model, d_model)
        self._index_divisor = index_divisor
        self._create_kernel_indices(kernel_size, "kernel_indices")
        self._reset_parameters()

    def _create_kernel_indices(self, kernel_size, module_name):
        self.register_buffer(module_name, _kernel_offsets(kernel_size, self._index_divisor))

    def _reset_parameters(self):
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.out_proj.bias, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.0)
        nn.init.constant_(self.linear_attn_weight, 0.0)
        nn.init.constant_(self.linear_attn_bias, 0.0)
        nn.init.constant_(self.linear_box_weight, 0.0)
        nn.init.uniform_(self.linear_box_bias)

    # box -> K*K sampling grid (box_attention.py:196-214)
    def _where_to_attend(self, query, v_valid_ratios, ref_windows):
        b, l = ref_windows.shape[:2]
        offset_boxes = F.linear(query, self.linear_box_weight, self.linear_box_bias)
        offset_boxes = offset_boxes.view(b, l, self.num_head, self.num_level, 4)
        if ref_windows.dim() == 3:
            ref_windows = ref_windows[:, :, None, None]
        else:
            ref_windows = ref_windows.unsqueeze(3)
        boxes = ref_windows + offset_boxes / 8 * ref_windows[..., [2, 3, 2, 3]]
        center, size = boxes.unsqueeze(-2).split(2, dim=-1)
        grid = center + self.kernel_indices * torch.relu(size)
        if v_valid_ratios is not None:
            grid = grid * v_valid_ratios
        return grid.contiguous()

    # the boxes _where_to_attend turns into a grid (box_attention.py:199-207); angles: None (no rotation)
    def _boxes_and_angles(self, query, ref_windows):
        b, l = ref_windows.shape[:2]
        offset_boxes = F.linear(query, self.linear_box_weight, self.linear_box_bias)
        offset_boxes = offset_boxes.view(b, l, self.num_head, self.num_level, 4)
        if ref_windows.dim() == 3:
            ref_windows = ref_windows[:, :, None, None]
        else:
            ref_windows = ref_windows.unsqueeze(3)
        boxes = ref_windows + offset_boxes / 8 * ref_windows[..., [2, 3, 2, 3]]
        return boxes, None

    def _attend(self, query, value, v_shape, v_start_index, v_valid_ratios, ref_windows, weights):
        """box attention of `value` at the K x K grids of the query boxes: fused op or grid + op."""
        if _FUSED_GRID and value.is_cuda:
            boxes, angles = self._boxes_and_angles(query, ref_windows)
            vr = None
            if v_valid_ratios is not None:
                vr = v_valid_ratios.reshape(v_valid_ratios.shape[0], self.num_level, 2)
            return _box_grid_attn(value, v_shape, v_start_index, boxes, angles, vr, self.kernel_indices, weights,
                                  self.im2col_step)
        sampled_grid = self._where_to_attend(query, v_valid_ratios, ref_windows)
        return _box_attn(value, v_shape, v_start_index, sampled_grid, weights, self.im2col_step)

    def _attend_logits(self, query, value, v_shape, v_start_index, v_valid_ratios, ref_windows, logits):
        """softmax over (L, K, K) + box attention; returns (output, attention weights (B,Nq,H,L,K,K))."""
        b, l1 = query.shape[:2]
        shape6 = (b, l1, self.num_head, self.num_level, self.kernel_size, self.kernel_size)
        if _FUSED_GRID and _FUSED_SOFTMAX and value.is_cuda:
            boxes, angles = self._boxes_and_angles(query, ref_windows)
            vr = None
            if v_valid_ratios is not None:
                vr = v_valid_ratios.reshape(v_valid_ratios.shape[0], self.num_level, 2)
            out, attn = _box_grid_softmax_attn(value, v_shape, v_start_index, boxes, angles, vr, self.kernel_indices,
                                               logits.view(shape6), self.im2col_step)
            return out, attn.view(shape6)
        attn = F.softmax(logits.view(b, l1, self.num_head, -1), dim=-1).view(shape6)
        return self._attend(query, value, v_shape, v_start_index, v_valid_ratios, ref_windows, attn), attn

    def _project_value(self, value, v_mask):
        b, l2 = value.shape[:2]
        value = self.value_proj(value)
        if _use_bf16(value) and value.dtype in (torch.float32, torch.bfloat16):
            # bf16-native ops: mask fill + cast to the gather's storage type in one pass
            value = ValueEpilogueFunction.apply(value, v_mask, torch.bfloat16)
        elif v_mask is not None:
            value = value.masked_fill(v_mask[..., None], float(0))
        return value.view(b, l2, self.num_head, self.head_dim)


class BoxAttention(_BoxAttentionBase):
    def __init__(self, d_model, num_level, num_head, kernel_size=2):
        super().__init__()
        self.num_point = kernel_size ** 2
        self._build(d_model, num_level, num_head, kernel_size, self.num_point, kernel_size)

    def forward(self, query, value, v_shape, v_mask, v_start_index, v_valid_ratios, ref_windows):
        b, l1 = query.shape[:2]
        value = self._project_value(value, v_mask)
        logits = F.linear(query, self.linear_attn_weight, self.linear_attn_bias)
        output, attn_weights = self._attend_logits(query, value, v_shape, v_start_index, v_valid_ratios, ref_windows, logits)
        output = self.out_proj(output)
        return output, attn_weights


class InstanceAttention(_BoxAttentionBase):
    """``self.inferencing`` is injected by the model (base_model.py:49-67); it is deliberately
    not set in ``__init__`` -- exactly like the reference, forward raises AttributeError without it."""

    def __init__(self, d_model, num_level, num_head, kernel_size):
        super().__init__()
        self._build(d_model, num_level, num_head, kernel_size, 4, kernel_size)

    def forward(self, query, value, v_shape, v_mask, v_start_index, v_valid_ratios, ref_windows):
        b, l1 = query.shape[:2]
        k = self.kernel_size
        value = self._project_value(value, v_mask)

        # a 2x2 logit map per (head, level), nearest-upsampled to KxK (box_attention.py:93-97)
        attn_weights = F.linear(query, self.linear_attn_weight, self.linear_attn_bias)
        attn_weights = attn_weights.view(b, l1, self.num_head, self.num_level, 2, 2)
        fused_w = _FUSED_SOFTMAX and value.is_cuda and k % 2 == 0
        if fused_w:     # both weight tensors from the 2x2 maps in one kernel
            spatial_attn_weights, level_attn_weights = InstanceWeightsFunction.apply(attn_weights, k)
        else:
            attn_weights = attn_weights.repeat_interleave(k // 2, dim=-1).repeat_interleave(k // 2, dim=-2)
            spatial_attn_weights = F.softmax(attn_weights.view(b, l1, self.num_head, -1), dim=-1)
            spatial_attn_weights = spatial_attn_weights.view(b, l1, self.num_head, self.num_level, k, k)

        if not self.inferencing:
            sampled_grid = self._where_to_attend(query, v_valid_ratios, ref_windows)
            if not fused_w:
                level_attn_weights = attn_weights.view(b, l1, self.num_head, self.num_level, k, k)
                level_attn_weights = F.softmax(level_attn_weights, dim=3)
            output, mask_output = _instance_attn(value, v_shape, v_start_index, sampled_grid,
                                                 spatial_attn_weights, level_attn_weights, k, self.im2col_step)
            attn_weights = (spatial_attn_weights, level_attn_weights)
            mask_output = self.out_proj(mask_output)
        else:
            output = self._attend(query, value, v_shape, v_start_index, v_valid_ratios, ref_windows, spatial_attn_weights)
            attn_weights = (spatial_attn_weights,)
            mask_output = None
        output = self.out_proj(output)
        return output, mask_output, attn_weights


class Box3dAttention(_BoxAttentionBase):
    def __init__(self, d_model, num_level, num_head, with_rotation=True, kernel_size=2):
        super().__init__()
        self.with_rotation = with_rotation
        self.num_variable = 5 if with_rotation else 4
        self._num_box_variable = self.num_variable
        self.num_point = kernel_size ** 2
        # NB the 3-D variant divides the offsets by 2, not by kernel_size (box_attention.py:291)
        self._build(d_model, num_level, num_head, kernel_size, self.num_point, 2)

    # box + angle -> rotated K*K grid (box_attention.py:304-338)
    def _where_to_attend(self, query, v_valid_ratios, ref_windows):
        b, l = ref_windows.shape[:2]
        offset_boxes = F.linear(query, self.linear_box_weight, self.linear_box_bias)
        offset_boxes = offset_boxes.view(b, l, self.num_head, self.num_level, self.num_variable)
        if ref_windows.dim() == 3:
            ref_windows = ref_windows[:, :, None, None]
            ref_windows, ref_angles, _ = ref_windows.split((4, 1, 2), dim=-1)
        else:
            ref_windows = ref_windows.unsqueeze(3)
            ref_windows, ref_angles = ref_windows.split((4, 1), dim=-1)

        if self.with_rotation:
            offset_boxes, offset_angles = offset_boxes.split(4, dim=-1)
            angles = (ref_angles + offset_angles / 16) * 2 * math.pi
        else:
            angles = ref_angles.expand(b, l, self.num_head, self.num_level, 1)

        boxes = ref_windows + offset_boxes / 8 * ref_windows[..., [2, 3, 2, 3]]
        center, size = boxes.unsqueeze(-2).split(2, dim=-1)
        cos_a, sin_a = torch.cos(angles), torch.sin(angles)
        rot = torch.stack([cos_a, -sin_a, sin_a, cos_a], dim=-1).view(b, l, self.num_head, self.num_level, 1, 2, 2)
        grid = self.kernel_indices * torch.relu(size)
        grid = center + (grid.unsqueeze(-2) * rot).sum(-1)
        if v_valid_ratios is not None:
            grid = grid * v_valid_ratios
        return grid.contiguous()

    def _boxes_and_angles(self, query, ref_windows):
        b, l = ref_windows.shape[:2]
        offset_boxes = F.linear(query, self.linear_box_weight, self.linear_box_bias)
        offset_boxes = offset_boxes.view(b, l, self.num_head, self.num_level, self.num_variable)
        if ref_windows.dim() == 3:
            ref_windows = ref_windows[:, :, None, None]
            ref_windows, ref_angles, _ = ref_windows.split((4, 1, 2), dim=-1)
        else:
            ref_windows = ref_windows.unsqueeze(3)
            ref_windows, ref_angles = ref_windows.split((4, 1), dim=-1)
        if self.with_rotation:
            offset_boxes, offset_angles = offset_boxes.split(4, dim=-1)
            angles = (ref_angles + offset_angles / 16) * 2 * math.pi
        else:
            angles = ref_angles.expand(b, l, self.num_head, self.num_level, 1)
        boxes = ref_windows + offset_boxes / 8 * ref_windows[..., [2, 3, 2, 3]]
        return boxes, angles

    def forward(self, query, value, v_shape, v_mask, v_start_index, v_valid_ratios, ref_windows):
        b, l1 = query.shape[:2]
        value = self._project_value(value, v_mask)
        logits = F.linear(query, self.linear_attn_weight, self.linear_attn_bias)
        output, attn_weights = self._attend_logits(query, value, v_shape, v_start_index, v_valid_ratios, ref_windows, logits)
        output = self.out_proj(output)
        return output, attn_weights
