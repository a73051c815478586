"""Synthetic inputs of the shapes BASELINE.json names (SURVEY.md section 8, table "Shapes").

Used by ``bench.py`` and by the full-size tests; no dataset, no checkpoint.

Location distributions:

* ``"box"``     -- box-structured, what the model actually feeds the op.  Encoder: the
  reference windows of ``BoxTransformer._create_ref_windows``
  (/root/reference/e2edet/module/box_transformer.py:70-116: centre = pixel centre of the
  query's own level, size = ref_size / level size, ref_size = 4) moved by
  ``offset / 8 * size`` with offsets ~ U[0,1) per (head, level, coord) -- the state of
  ``linear_box_bias`` at init (box_attention.py:186-194) -- then the K x K grid of
  ``_where_to_attend`` (box_attention.py:196-214).  Decoder / mask head: random boxes.
* ``"uniform"`` -- every sample point U[0,1)^2 independently (what the reference unit tests
  draw; worst case for locality).
* ``"trained"`` -- trained-like encoder boxes: what ``"box"`` becomes once ``linear_box_weight`` is no
  longer zero.  Every (query, head, level) has its own box: centre = the query's pixel centre moved by
  U(-1/2, 1/2) of the box size, width and height log-uniform in [2, 64] pixels of the finest level
  (the init state is 4 px of the query's own level), then the same K x K grid.  Between the two
  extremes above: neighbouring queries still look at neighbouring pixels, but footprints differ per
  row and many exceed what a 64-pixel window holds.

``oob`` moves that fraction of the points outside [0,1] to exercise zero padding.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


def fpn_levels(height=800, width=1333, strides=(8, 16, 32), extra=1):
    """Level shapes of BoxeR-2D for an image: ResNet C3-C5 strides + `extra` stride-2 convs
    (resnet.py:365-391, boxer2d.py:68-76).  800x1333 -> (100,167),(50,84),(25,42),(13,21)."""
    shapes = [(math.ceil(height / s), math.ceil(width / s)) for s in strides]
    for _ in range(extra):
        h, w = shapes[-1]
        shapes.append(((h + 1) // 2, (w + 1) // 2))
    return shapes


@dataclass
class Workload:
    name: str
    value: torch.Tensor          # (B,S,H,D)
    shapes: torch.Tensor         # (L,2) int64
    level_start: torch.Tensor    # (L,) int64
    loc: torch.Tensor            # (B,Nq,H,L,P,2)
    weights: tuple               # (attn,) or (spatial_w, level_w), each (B,Nq,H,L,K,K)
    kernel_size: int
    instance: bool = False

    @property
    def n_samples(self) -> int:
        B, Nq, H, L, P = self.loc.shape[:5]
        return B * Nq * H * L * P

    @property
    def dims(self):
        B, S, H, D = self.value.shape
        L = self.shapes.shape[0]
        Nq, P = self.loc.shape[1], self.loc.shape[4]
        return dict(B=B, S=S, H=H, D=D, L=L, Nq=Nq, P=P)

    def to(self, device=None, dtype=None):
        def mv(t, dt=None):
            return t.to(device=device, dtype=dt if (dt is not None and t.is_floating_point()) else None)
        return Workload(self.name, mv(self.value, dtype), mv(self.shapes), mv(self.level_start), mv(self.loc),
                        tuple(mv(w) for w in self.weights), self.kernel_size, self.instance)


def _level_meta(shapes, device):
    sh = torch.tensor(shapes, dtype=torch.long, device=device)
    start = torch.cat((sh.new_zeros(1), sh.prod(1).cumsum(0)[:-1]))
    return sh, start


def _kernel_offsets(K, divisor, device):
    half = K / 2.0
    if K % 2 == 0:
        ticks = torch.linspace(-half + 0.5, half - 0.5, K, device=device)
    else:
        r = (K - 1) // 2
        ticks = torch.linspace(-r, r, K, device=device)
    yy, xx = torch.meshgrid(ticks, ticks, indexing="ij")
    return torch.stack([xx, yy], -1).reshape(-1, 2) / divisor


def encoder_ref_windows(shapes, B, device, ref_size=4.0):
    """(B,S,4) cx,cy,w,h -- box_transformer.py:70-116 without padding masks."""
    refs = []
    for h, w in shapes:
        ys = (torch.arange(1, h + 1, device=device, dtype=torch.float32) - 0.5) / (h + 1e-6)
        xs = (torch.arange(1, w + 1, device=device, dtype=torch.float32) - 0.5) / (w + 1e-6)
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        size = torch.tensor([ref_size / w, ref_size / h], device=device).expand(h, w, 2)
        refs.append(torch.cat([torch.stack([xx, yy], -1), size], -1).reshape(h * w, 4))
    return torch.cat(refs, 0).unsqueeze(0).expand(B, -1, -1).contiguous()


def _grid_from_boxes(ref, offsets, K, divisor, angles=None):
    """ref (B,Nq,4) + offsets (B,Nq,H,L,4) -> (B,Nq,H,L,K*K,2); optional rotation (box_attention.py:304-338)."""
    ref = ref[:, :, None, None]
    boxes = ref + offsets / 8 * ref[..., [2, 3, 2, 3]]
    center, size = boxes.unsqueeze(-2).split(2, dim=-1)
    g = _kernel_offsets(K, divisor, ref.device) * torch.relu(size)
    if angles is not None:
        c, s = torch.cos(angles), torch.sin(angles)
        rot = torch.stack([c, -s, s, c], -1).view(*angles.shape[:4], 1, 2, 2)
        g = (g.unsqueeze(-2) * rot).sum(-1)
    return (center + g).contiguous()


def _apply_oob(loc, frac, gen):
    if frac <= 0:
        return loc
    m = torch.rand(loc.shape[:-1], device=loc.device, generator=gen) < frac
    shift = torch.where(torch.rand(loc.shape, device=loc.device, generator=gen) < 0.5, -1.25, 1.25)
    return torch.where(m[..., None], loc + shift, loc)


def _softmax_weights(B, Nq, H, L, K, gen, device):
    logits = torch.randn(B, Nq, H, L * K * K, device=device, generator=gen)
    return torch.softmax(logits, -1).view(B, Nq, H, L, K, K)


def coco_encoder(B=1, K=4, dist="box", oob=0.0, heads=8, head_dim=32, image=(800, 1333), seed=3,
                 device="cuda", value_scale=1.0) -> Workload:
    """BASELINE.json configs[1]: BoxeR-2D encoder box-attn, 4 FPN levels of 1333x800, C=256,
    Nq = S = 22223 queries (every pixel of every level), 8 heads, KxK grid (reference K=2,
    BASELINE K=4), fp32."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    shapes = fpn_levels(*image)
    sh, start = _level_meta(shapes, dev)
    S = int(sh.prod(1).sum())
    L = len(shapes)
    value = torch.randn(B, S, heads, head_dim, device=dev, generator=gen) * value_scale
    if dist == "box":
        ref = encoder_ref_windows(shapes, B, dev)
        offsets = torch.rand(heads, L, 4, device=dev, generator=gen).expand(B, S, heads, L, 4)
        loc = _grid_from_boxes(ref, offsets, K, K)
    elif dist == "uniform":
        loc = torch.rand(B, S, heads, L, K * K, 2, device=dev, generator=gen)
    elif dist == "trained":
        loc = _trained_like_grid(encoder_ref_windows(shapes, B, dev), shapes[0], heads, L, K, gen)
    else:
        raise ValueError(dist)
    loc = _apply_oob(loc, oob, gen)
    attn = _softmax_weights(B, S, heads, L, K, gen, dev)
    return Workload(f"coco_encoder_K{K}_{dist}", value, sh, start, loc, (attn,), K)


def _trained_like_grid(ref, shape0, heads, L, K, gen, lo=2.0, hi=64.0):
    """ref (B,S,4) -> (B,S,H,L,K*K,2): per-(query, head, level) boxes, sizes log-uniform in [lo, hi] pixels of level 0."""
    B, S = ref.shape[:2]
    dev = ref.device
    h0, w0 = shape0
    px = torch.exp(math.log(lo) + (math.log(hi) - math.log(lo)) * torch.rand(B, S, heads, L, 2, device=dev, generator=gen))
    size = px / torch.tensor([w0, h0], device=dev, dtype=torch.float32)
    centre = ref[:, :, None, None, :2] + (torch.rand(B, S, heads, L, 2, device=dev, generator=gen) - 0.5) * size
    return (centre.unsqueeze(-2) + _kernel_offsets(K, K, dev) * size.unsqueeze(-2)).contiguous()


def window_mode_fraction(w: Workload, cap: int = 64):
    """Fraction of (row, level) pairs whose touched pixel range fits a `cap`-pixel footprint window (the window
    kernels' fast mode, boxattn_window.cuh) -- the rest take the per-point walk.  Restates the kernels' range rule:
    points inside the window test, floor(x) .. floor(x)+1 clamped to the level."""
    loc = w.loc.float()
    fits = []
    for l, (h, wd) in enumerate(w.shapes.tolist()):
        x = loc[:, :, :, l, :, 0] * wd - 0.5
        y = loc[:, :, :, l, :, 1] * h - 0.5
        inside = (x > -1) & (y > -1) & (x < wd) & (y < h)
        big = 1 << 30
        x0 = torch.where(inside, torch.floor(x), torch.full_like(x, big)).amin(-1).clamp_min(0)
        x1 = torch.where(inside, torch.floor(x) + 1, torch.full_like(x, -big)).amax(-1).clamp_max(wd - 1)
        y0 = torch.where(inside, torch.floor(y), torch.full_like(y, big)).amin(-1).clamp_min(0)
        y1 = torch.where(inside, torch.floor(y) + 1, torch.full_like(y, -big)).amax(-1).clamp_max(h - 1)
        nx, ny = x1 - x0 + 1, y1 - y0 + 1
        fits.append(((nx * ny <= cap) | (nx <= 0) | (ny <= 0)).float().mean())
    return float(torch.stack(fits).mean())


def backward_reduction_bytes(w: Workload, cap: int = 64, line_bytes: int = None):
    """Bytes of `red.global.add` payload the window backward issues into `grad_value` for this workload: per
    (row, level) one reduction of a head's channel slice (D accumulators) per *unique* touched pixel with a non-zero
    weight when the touched range fits the `cap`-slot window, one per in-range corner otherwise (boxattn_window.cuh,
    phases B / C).  The L2 performs reductions at a fixed payload rate (scripts/microbench/red_throughput.cu) far below
    its load bandwidth, which makes this the backward's roofline numerator.  Checked against ncu's count of executed
    `RED.128`s on the headline workload (profiles/README.md r02zz: 1.552 GB measured, 1.58 GB here)."""
    loc = w.loc.float()
    attn = w.weights[0].reshape(loc.shape[:-1])
    D = w.value.shape[-1]
    line = line_bytes if line_bytes is not None else 4 * D          # fp32 accumulators (bf16 values accumulate in fp32 too)
    total = 0
    for l, (h, wd) in enumerate(w.shapes.tolist()):
        x = loc[:, :, :, l, :, 0] * wd - 0.5
        y = loc[:, :, :, l, :, 1] * h - 0.5
        inside = (x > -1) & (y > -1) & (x < wd) & (y < h)
        x0, y0 = torch.floor(x).long(), torch.floor(y).long()
        big = 1 << 30
        bx0 = torch.where(inside, x0, torch.full_like(x0, big)).amin(-1).clamp_min(0)
        bx1 = torch.where(inside, x0 + 1, torch.full_like(x0, -big)).amax(-1).clamp_max(wd - 1)
        by0 = torch.where(inside, y0, torch.full_like(y0, big)).amin(-1).clamp_min(0)
        by1 = torch.where(inside, y0 + 1, torch.full_like(y0, -big)).amax(-1).clamp_max(h - 1)
        area = (bx1 - bx0 + 1).clamp_min(0) * (by1 - by0 + 1).clamp_min(0)
        lx, ly = x - x0, y - y0
        pix, corners = [], 0
        for dy in (0, 1):
            for dx in (0, 1):
                a, b = x0 + dx, y0 + dy
                ok = inside & (a >= 0) & (a < wd) & (b >= 0) & (b < h)
                wgt = attn[:, :, :, l] * (lx if dx else 1 - lx) * (ly if dy else 1 - ly)
                pix.append(torch.where(ok & (wgt != 0), b * wd + a, torch.full_like(a, -1)))
                corners = corners + ok.long().sum(-1)
        srt = torch.cat(pix, -1).sort(-1).values
        uniq = (srt[..., 1:] != srt[..., :-1]).sum(-1) + 1 - (srt[..., 0] == -1).long()
        total += int(torch.where(area <= cap, uniq, corners).sum())
    return total * line


def box3d_encoder(B=1, K=2, heads=8, head_dim=32, levels=((234, 234), (117, 117)), seed=7, device="cuda", ref_size=4.0) -> Workload:
    """The reference-exact BoxeR-3D encoder call (note N1 / config c5): two BEV levels 234x234 + 117x117
    (base_boxer3d_detection.yaml:132-146), C=256 (D=32), Nq = S = 68 445, 2x2 grid with the /2 index divisor
    (box_attention.py:291), ``Box3dAttention(with_rotation=False)`` (box3d_transformer.py:233).  Reference windows as
    ``Box3dTransformer._create_ref_windows`` (:57-110): centre = (i + 0.5) / size, size = ref_size / size, one window per
    head whose fifth entry is the head's reference angle in *normalised* units -- which the non-rotating attention hands
    to cos / sin as it is (box_attention.py:318-327), so every head's grid is turned by a fixed 0.5 .. 1.0 rad."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    sh, start = _level_meta(list(levels), dev)
    S = int(sh.prod(1).sum())
    L = len(levels)
    value = torch.randn(B, S, heads, head_dim, device=dev, generator=gen)
    ang = torch.tensor([0, 2 * math.pi / 3, -2 * math.pi / 3, 0, 2 * math.pi / 3, -2 * math.pi / 3, 0, math.pi], device=dev)
    ang = ((ang + 0.5 * 2 * math.pi) / (2 * math.pi))[:heads]
    refs = []
    for h, w in levels:
        ys = (torch.arange(h, device=dev, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, device=dev, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        box = torch.stack([xx, yy, torch.full_like(xx, ref_size / w), torch.full_like(yy, ref_size / h)], -1)
        refs.append(box.reshape(h * w, 4))
    ref = torch.cat(refs, 0)[None, :, None, None, :].expand(B, S, heads, 1, 4)
    offsets = torch.rand(heads, L, 4, device=dev, generator=gen).expand(B, S, heads, L, 4)
    boxes = ref + offsets / 8 * ref[..., [2, 3, 2, 3]]
    center, size = boxes.unsqueeze(-2).split(2, dim=-1)
    g = _kernel_offsets(K, 2, dev) * torch.relu(size)
    angles = ang.view(1, 1, heads, 1).expand(B, S, heads, L)
    c, s_ = torch.cos(angles), torch.sin(angles)
    rot = torch.stack([c, -s_, s_, c], -1).view(B, S, heads, L, 1, 2, 2)
    loc = (center + (g.unsqueeze(-2) * rot).sum(-1)).contiguous()
    attn = _softmax_weights(B, S, heads, L, K, gen, dev)
    return Workload(f"box3d_encoder_K{K}", value, sh, start, loc, (attn,), K)


def random_boxes(B, Nq, gen, device):
    cxcy = 0.1 + 0.8 * torch.rand(B, Nq, 2, device=device, generator=gen)
    wh = 0.05 + 0.45 * torch.rand(B, Nq, 2, device=device, generator=gen)
    return torch.cat([cxcy, wh], -1)


def coco_decoder(B=1, Nq=300, K=2, heads=8, head_dim=32, image=(800, 1333), seed=4, device="cuda") -> Workload:
    """Decoder box-attn: Nq object queries with arbitrary boxes (box_transformer.py:385)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    shapes = fpn_levels(*image)
    sh, start = _level_meta(shapes, dev)
    S = int(sh.prod(1).sum())
    L = len(shapes)
    value = torch.randn(B, S, heads, head_dim, device=dev, generator=gen)
    offsets = torch.rand(B, Nq, heads, L, 4, device=dev, generator=gen)
    loc = _grid_from_boxes(random_boxes(B, Nq, gen, dev), offsets, K, K)
    attn = _softmax_weights(B, Nq, heads, L, K, gen, dev)
    return Workload(f"coco_decoder_K{K}", value, sh, start, loc, (attn,), K)


def coco_mask_head(B=1, Nq=300, K=14, heads=8, head_dim=32, image=(800, 1333), seed=5, device="cuda") -> Workload:
    """BASELINE.json configs[3]: InstanceAttention, 300 queries x KxK RoI grid over the 4 levels
    (reference K=14, box_transformer.py:383; BASELINE K=28).  Weights as InstanceAttention.forward
    builds them (box_attention.py:93-110): 2x2 logits repeated to KxK, softmax over (L,K,K) and over L."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    shapes = fpn_levels(*image)
    sh, start = _level_meta(shapes, dev)
    S = int(sh.prod(1).sum())
    L = len(shapes)
    value = torch.randn(B, S, heads, head_dim, device=dev, generator=gen)
    offsets = torch.rand(B, Nq, heads, L, 4, device=dev, generator=gen)
    loc = _grid_from_boxes(random_boxes(B, Nq, gen, dev), offsets, K, K)
    logits = torch.randn(B, Nq, heads, L, 2, 2, device=dev, generator=gen)
    logits = logits.repeat_interleave(K // 2, -1).repeat_interleave(K // 2, -2)
    sw = torch.softmax(logits.reshape(B, Nq, heads, -1), -1).view(B, Nq, heads, L, K, K)
    lw = torch.softmax(logits, 3).contiguous()
    return Workload(f"coco_mask_head_K{K}", value, sh, start, loc, (sw, lw), K, instance=True)


def bev_rotated(B=1, Nq=1000, K=3, heads=8, head_dim=16, size=468, seed=6, device="cuda") -> Workload:
    """BASELINE.json configs[4]: BoxeR-3D BEV box-attn with rotation, one 468x468 level, C=128."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    shapes = [(size, size)]
    sh, start = _level_meta(shapes, dev)
    S = size * size
    value = torch.randn(B, S, heads, head_dim, device=dev, generator=gen)
    offsets = torch.rand(B, Nq, heads, 1, 4, device=dev, generator=gen)
    angles = torch.rand(B, Nq, heads, 1, 1, device=dev, generator=gen) * 2 * math.pi
    loc = _grid_from_boxes(random_boxes(B, Nq, gen, dev), offsets, K, 2, angles=angles)
    attn = _softmax_weights(B, Nq, heads, 1, K, gen, dev)
    return Workload(f"bev_rotated_K{K}", value, sh, start, loc, (attn,), K)


# ----------------------------------------------------------------------------- roofline model
def bytes_per_sample(instance: bool, backward: bool, D: int, L: int, P: int, sz: int = 4) -> float:
    """Algorithmic bytes per sample, SURVEY.md 8(d) / BASELINE.md section 2 (no-reuse traffic model)."""
    if not instance:
        return sz * ((12 * D + 6 + D / (L * P)) if backward else (4 * D + 3 + D / (L * P)))
    return sz * ((12 * D + 8 + D / (L * P) + D / L) if backward else (4 * D + 4 + D / (L * P) + D / L))
