"""grid_sample restatement of the reference oracle.  TEST INFRASTRUCTURE ONLY.

Behavioural restatement (not a copy) of the pure-PyTorch oracle the reference
tests compare their CUDA op against:

* ``plain_box_attn``       <- ``/root/reference/tests/box_attn_test.py:9-42``      (PlainBoxAttnFunction)
* ``plain_instance_attn``  <- ``/root/reference/tests/instance_attn_test.py:11-63`` (PlainInstanceAttnFunction)
* ``split_levels``         <- ``/root/reference/e2edet/utils/general.py:289-324``   (view_with_shape, tensor half)

Conventions are the reference's: ``value`` is ``(B, S, C)`` with the L levels
concatenated along S in row-major (y, x) order, ``C = heads * head_dim`` with
the head index major; ``grid`` is in grid_sample's [-1, 1] convention, i.e.
the caller passes ``2 * sampling_locations - 1`` exactly as the reference tests
do; sampling is bilinear, ``align_corners=False``, zero padding.

Differentiable through autograd, so fp64 gradients of this function are the
gradient oracle as well (the reference's ``check_forward_and_backward``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def split_levels(value: torch.Tensor, shapes) -> list[torch.Tensor]:
    """(B, S, C) -> [(B, C, H_l, W_l)] per level (general.py:289-324)."""
    if torch.is_tensor(shapes):
        shapes = shapes.tolist()
    sizes = [int(h) * int(w) for h, w in shapes]
    B, S, C = value.shape
    assert sum(sizes) == S, f"levels cover {sum(sizes)} positions, value has {S}"
    out = []
    for chunk, (h, w) in zip(torch.split(value, sizes, dim=1), shapes):
        out.append(chunk.reshape(B, int(h), int(w), C).permute(0, 3, 1, 2).contiguous())
    return out


def _sample_level(level_map: torch.Tensor, grid_l: torch.Tensor, heads: int) -> torch.Tensor:
    """Bilinear samples of one level.

    level_map: (B, C, h, w); grid_l: (B, Nq, heads, P, 2) in [-1, 1] (x, y).
    Returns (B, Nq, heads, D, P).
    """
    B, C, h, w = level_map.shape
    D = C // heads
    Nq, P = grid_l.shape[1], grid_l.shape[3]
    maps = level_map.reshape(B * heads, D, h, w)
    g = grid_l.permute(0, 2, 1, 3, 4).reshape(B * heads, Nq, P, 2)
    s = F.grid_sample(maps, g, mode="bilinear", padding_mode="zeros", align_corners=False)
    # (B*heads, D, Nq, P) -> (B, Nq, heads, D, P)
    return s.reshape(B, heads, D, Nq, P).permute(0, 3, 1, 2, 4)


def plain_box_attn(value, shapes, grid, attention_weights):
    """out[b,q,(h,d)] = sum_l sum_p w[b,q,h,l,p] * bilinear(value_l[b,:,h,d], grid[b,q,h,l,p]).

    value (B,S,C); grid (B,Nq,H,L,P,2) in [-1,1]; attention_weights (B,Nq,H,L,P)
    (a trailing (K,K) pair is accepted and flattened).  Returns (B,Nq,C).
    """
    B, Nq, H, L, P = grid.shape[:5]
    w = attention_weights.reshape(B, Nq, H, L, P)
    total = None
    for l, level_map in enumerate(split_levels(value, shapes)):
        s = _sample_level(level_map, grid[:, :, :, l], H)          # (B,Nq,H,D,P)
        contrib = (s * w[:, :, :, l].unsqueeze(-2)).sum(-1)         # (B,Nq,H,D)
        total = contrib if total is None else total + contrib
    return total.reshape(B, Nq, -1)


def plain_instance_attn(value, shapes, grid, spatial_weights, level_weights, mask_size):
    """Two outputs (instance_attn_test.py:11-63).

    out[b,q,(h,d)]        = sum_l sum_p spatial_w[b,q,h,l,p] * sample
    mask[b,q,ky,kx,(h,d)] = sum_l       level_w[b,q,h,l,p]   * sample,   p = ky*K + kx
    """
    B, Nq, H, L, P = grid.shape[:5]
    K = int(mask_size)
    if P != K * K:
        raise AssertionError(f"mask_points: {P}, mask_size: {K}")
    if K % 2:
        raise AssertionError("Only support even mask_size!")
    sw = spatial_weights.reshape(B, Nq, H, L, P)
    lw = level_weights.reshape(B, Nq, H, L, P)
    out = None
    mask = None
    for l, level_map in enumerate(split_levels(value, shapes)):
        s = _sample_level(level_map, grid[:, :, :, l], H)          # (B,Nq,H,D,P)
        o = (s * sw[:, :, :, l].unsqueeze(-2)).sum(-1)
        m = s * lw[:, :, :, l].unsqueeze(-2)
        out = o if out is None else out + o
        mask = m if mask is None else mask + m
    out = out.reshape(B, Nq, -1)
    # (B,Nq,H,D,P) -> (B,Nq,K,K,H*D)
    mask = mask.permute(0, 1, 4, 2, 3).reshape(B, Nq, K, K, -1)
    return out, mask


def grid_from_boxes(boxes, angles, valid_ratios, kernel_indices):
    """The K x K sampling grid the reference modules build from boxes (TEST INFRASTRUCTURE ONLY).

    Restates ``BoxAttention._where_to_attend`` (/root/reference/e2edet/module/box_attention.py:207-212)
    and the rotated variant of ``Box3dAttention`` (:321-336) from the point where ``boxes`` / ``angles``
    are known: boxes (B,Nq,H,L,4) cx,cy,w,h; angles (B,Nq,H,L,1) radians or None; valid_ratios
    (B,1,1,L,1,2) or None; kernel_indices (P,2).  Returns (B,Nq,H,L,P,2) in [0,1] coordinates."""
    center, size = boxes.unsqueeze(-2).split(2, dim=-1)
    grid = kernel_indices * torch.relu(size)
    if angles is not None:
        cos_a, sin_a = torch.cos(angles), torch.sin(angles)
        rot = torch.stack([cos_a, -sin_a, sin_a, cos_a], dim=-1).view(*angles.shape[:4], 1, 2, 2)
        grid = (grid.unsqueeze(-2) * rot).sum(-1)
    grid = center + grid
    if valid_ratios is not None:
        grid = grid * valid_ratios
    return grid
