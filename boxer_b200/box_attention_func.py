"""``BoxAttnFunction`` / ``InstanceAttnFunction``: the autograd boundary.

Drop-in for ``/root/reference/e2edet/module/ops/box_attention_func.py:9-150``:
same class names, same ``apply`` argument order, same outputs, same
``(grad, None, None, grad, grad, None...)`` backward tuples, same AMP contract
(``custom_fwd(cast_inputs=torch.float32)``: under autocast the op runs in fp32),
``once_differentiable``.  The native calls go to ``boxer_b200.ops`` (C ABI).

Beyond the reference: ``BoxAttnBf16Function`` / ``InstanceAttnBf16Function`` keep
``value`` / outputs in bfloat16 (locations and weights fp32, fp32 accumulation) --
an explicit opt-in used by the modules when ``boxer_b200.set_amp_native(True)``.
"""
from __future__ import annotations

import torch
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops


class BoxAttnFunction(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = ops.box_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                      sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_contiguous():
            grad_output = grad_output.contiguous()
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = ops.box_attn_backward(
            value, shapes, lsi, loc, attn, grad_output.to(value.dtype), ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


class InstanceAttnFunction(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                spatial_attention_weights, level_attention_weights, mask_size, im2col_step):
        ctx.im2col_step = im2col_step
        output, mask_output = ops.instance_attn_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations,
            spatial_attention_weights, level_attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              spatial_attention_weights, level_attention_weights)
        b, l, _, c = mask_output.shape
        mask_output = mask_output.view(b, l, mask_size, mask_size, c)
        return output, mask_output

    @staticmethod
    @custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, grad_output, grad_mask_output):
        if not grad_output.is_contiguous():
            grad_output = grad_output.contiguous()
        if not grad_mask_output.is_contiguous():
            grad_mask_output = grad_mask_output.contiguous()
        value, shapes, lsi, loc, sw, lw = ctx.saved_tensors
        grad_value, grad_loc, grad_sw, grad_lw = ops.instance_attn_backward(
            value, shapes, lsi, loc, sw, lw, grad_output.to(value.dtype), grad_mask_output.to(value.dtype),
            ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_sw, grad_lw, None, None


class BoxGridAttnFunction(Function):
    """Fused box -> K x K grid -> box attention (beyond the reference's surface).

    ``apply(value, shapes, level_start_index, boxes, angles, valid_ratios, kernel_indices, attention_weights,
    im2col_step)`` equals ``BoxAttnFunction.apply(value, shapes, lsi, grid, attention_weights, im2col_step)`` with
    ``grid = (centre + R(angles) (kernel_indices * relu(size))) * valid_ratios`` as built by
    ``BoxAttention._where_to_attend`` (box_attention.py:196-214) / ``Box3dAttention`` (:304-338), but the
    (B,Nq,H,L,P,2) grid and its gradient are never materialised; gradients flow to ``boxes`` (B,Nq,H,L,4) and
    ``angles`` (B,Nq,H,L[,1]).  ``valid_ratios`` (B,L,2 or the reference's B,1,1,L,1,2) and ``kernel_indices`` are
    treated as constants, as they are in BoxeR."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, shapes, lsi, boxes, angles, valid_ratios, kernel_indices, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.has = (angles is not None, valid_ratios is not None)
        boxes = boxes.contiguous()
        angles_c = angles.contiguous() if angles is not None else None
        vr = valid_ratios.contiguous() if valid_ratios is not None else None
        kidx = kernel_indices.to(boxes.dtype).contiguous()
        out = ops.box_grid_attn_forward(value, shapes, lsi, boxes, angles_c, vr, kidx, attention_weights, im2col_step)
        saved = [value, shapes, lsi, boxes, kidx, attention_weights] + [t for t in (angles_c, vr) if t is not None]
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    @custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, grad_output):
        if not grad_output.is_contiguous():
            grad_output = grad_output.contiguous()
        value, shapes, lsi, boxes, kidx, attn = ctx.saved_tensors[:6]
        rest = list(ctx.saved_tensors[6:])
        angles = rest.pop(0) if ctx.has[0] else None
        vr = rest.pop(0) if ctx.has[1] else None
        gv, gb, ga, gw = ops.box_grid_attn_backward(value, shapes, lsi, boxes, angles, vr, kidx, attn,
                                                    grad_output.to(value.dtype), ctx.im2col_step)
        return gv, None, None, gb, ga, None, None, gw, None


class BoxGridSoftmaxAttnFunction(Function):
    """Fused softmax -> box -> grid -> attention: as ``BoxGridAttnFunction`` but takes the
    attention LOGITS (B,Nq,H,L*P or B,Nq,H,L,K,K) the modules feed to ``F.softmax(dim=-1)``
    (box_attention.py:227-231) and returns ``(output, attention_weights)``; ``attention_weights`` (what the module
    returns to its caller) is a by-product and not differentiable through this Function."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, shapes, lsi, boxes, angles, valid_ratios, kernel_indices, logits, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.has = (angles is not None, valid_ratios is not None)
        boxes = boxes.contiguous()
        angles_c = angles.contiguous() if angles is not None else None
        vr = valid_ratios.contiguous() if valid_ratios is not None else None
        kidx = kernel_indices.to(boxes.dtype).contiguous()
        out, attn = ops.box_grid_attn_forward(value, shapes, lsi, boxes, angles_c, vr, kidx, logits.contiguous(),
                                              im2col_step, softmax=True)
        ctx.save_for_backward(*([value, shapes, lsi, boxes, kidx, attn] + [t for t in (angles_c, vr) if t is not None]))
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    @custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, grad_output, _grad_attn):
        if not grad_output.is_contiguous():
            grad_output = grad_output.contiguous()
        value, shapes, lsi, boxes, kidx, attn = ctx.saved_tensors[:6]
        rest = list(ctx.saved_tensors[6:])
        angles = rest.pop(0) if ctx.has[0] else None
        vr = rest.pop(0) if ctx.has[1] else None
        gv, gb, ga, gz = ops.box_grid_attn_backward(value, shapes, lsi, boxes, angles, vr, kidx, attn,
                                                    grad_output.to(value.dtype), ctx.im2col_step, softmax=True)
        return gv, None, None, gb, ga, None, None, gz, None


class ValueEpilogueFunction(Function):
    """``apply(projected_value (B,S,C), v_mask (B,S) bool | None, out_dtype)``: padding-mask fill + storage cast in one
    pass (box_attention.py:222-225); the gradient is masked and cast back the same way."""

    @staticmethod
    def forward(ctx, value, mask, out_dtype):
        ctx.in_dtype = value.dtype
        ctx.mask = mask
        return ops.value_epilogue(value.contiguous(), mask.contiguous() if mask is not None else None, out_dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        mask = ctx.mask
        return ops.value_epilogue(grad.contiguous(), mask.contiguous() if mask is not None else None, ctx.in_dtype), None, None


class InstanceWeightsFunction(Function):
    """``apply(logits (B,Nq,H,L,2,2), kernel_size) -> (spatial_w, level_w)``, each (B,Nq,H,L,K,K): InstanceAttention's
    ``repeat_interleave`` x2 + softmax over (L,K,K) + softmax over L (box_attention.py:93-110) and their backward as
    one kernel each way."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, logits, kernel_size):
        logits = logits.contiguous()
        ctx.kernel_size = int(kernel_size)
        ctx.save_for_backward(logits)
        return ops.instance_weights_forward(logits, kernel_size)

    @staticmethod
    @custom_bwd(device_type="cuda")
    @once_differentiable
    def backward(ctx, grad_sw, grad_lw):
        (logits,) = ctx.saved_tensors
        return ops.instance_weights_backward(logits, grad_sw.to(logits.dtype).contiguous(),
                                             grad_lw.to(logits.dtype).contiguous(), ctx.kernel_size), None


# --------------------------------------------------------------------------- bf16 opt-in
def _bf16_inputs(value, loc, *weights):
    return (value.to(torch.bfloat16).contiguous(), loc.float().contiguous(),
            *[w.float().contiguous() for w in weights])


class BoxAttnBf16Function(Function):
    """value / out in bf16, loc and weights fp32, fp32 accumulation.  Gradients come back in
    the dtypes of the tensors that were passed in."""

    @staticmethod
    def forward(ctx, value, shapes, lsi, loc, attn, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.in_dtypes = (value.dtype, loc.dtype, attn.dtype)
        with torch.autocast("cuda", enabled=False):
            v, l, a = _bf16_inputs(value, loc, attn)
            out = ops.box_attn_forward(v, shapes, lsi, l, a, im2col_step)
        ctx.save_for_backward(v, shapes, lsi, l, a)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        v, shapes, lsi, l, a = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            gv, gl, ga = ops.box_attn_backward(v, shapes, lsi, l, a, grad_output.to(torch.bfloat16).contiguous(),
                                               ctx.im2col_step)
        dv, dl, da = ctx.in_dtypes
        return gv.to(dv), None, None, gl.to(dl), ga.to(da).view_as(a), None


class InstanceAttnBf16Function(Function):
    @staticmethod
    def forward(ctx, value, shapes, lsi, loc, sw, lw, mask_size, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.in_dtypes = (value.dtype, loc.dtype, sw.dtype, lw.dtype)
        with torch.autocast("cuda", enabled=False):
            v, l, s, w = _bf16_inputs(value, loc, sw, lw)
            out, mask = ops.instance_attn_forward(v, shapes, lsi, l, s, w, im2col_step)
        ctx.save_for_backward(v, shapes, lsi, l, s, w)
        b, nq, _, c = mask.shape
        return out, mask.view(b, nq, mask_size, mask_size, c)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, grad_mask_output):
        v, shapes, lsi, l, s, w = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            gv, gl, gs, gw = ops.instance_attn_backward(
                v, shapes, lsi, l, s, w, grad_output.to(torch.bfloat16).contiguous(),
                grad_mask_output.to(torch.bfloat16).contiguous(), ctx.im2col_step)
        dv, dl, ds, dw = ctx.in_dtypes
        return gv.to(dv), None, None, gl.to(dl), gs.to(ds), gw.to(dw), None, None


class BoxGridSoftmaxAttnBf16Function(Function):
    """bf16 value / output variant of BoxGridSoftmaxAttnFunction (boxes, angles, logits fp32)."""

    @staticmethod
    def forward(ctx, value, shapes, lsi, boxes, angles, valid_ratios, kernel_indices, logits, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.has = (angles is not None, valid_ratios is not None)
        ctx.in_dtypes = (value.dtype, boxes.dtype, angles.dtype if angles is not None else None, logits.dtype)
        with torch.autocast("cuda", enabled=False):
            v = value.to(torch.bfloat16).contiguous()
            bx = boxes.float().contiguous()
            an = angles.float().contiguous() if angles is not None else None
            vr = valid_ratios.float().contiguous() if valid_ratios is not None else None
            kidx = kernel_indices.float().contiguous()
            out, attn = ops.box_grid_attn_forward(v, shapes, lsi, bx, an, vr, kidx, logits.float().contiguous(),
                                                  im2col_step, softmax=True)
        ctx.save_for_backward(*([v, shapes, lsi, bx, kidx, attn] + [t for t in (an, vr) if t is not None]))
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, _grad_attn):
        v, shapes, lsi, bx, kidx, a = ctx.saved_tensors[:6]
        rest = list(ctx.saved_tensors[6:])
        an = rest.pop(0) if ctx.has[0] else None
        vr = rest.pop(0) if ctx.has[1] else None
        with torch.autocast("cuda", enabled=False):
            gv, gb, ga, gz = ops.box_grid_attn_backward(v, shapes, lsi, bx, an, vr, kidx, a,
                                                        grad_output.to(torch.bfloat16).contiguous(), ctx.im2col_step,
                                                        softmax=True)
        dv, db, da, dz = ctx.in_dtypes
        return (gv.to(dv), None, None, gb.to(db), ga.to(da) if ga is not None else None, None, None, gz.to(dz), None)


class BoxGridAttnBf16Function(Function):
    """bf16 value / output variant of BoxGridAttnFunction (boxes, angles, weights fp32)."""

    @staticmethod
    def forward(ctx, value, shapes, lsi, boxes, angles, valid_ratios, kernel_indices, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        ctx.has = (angles is not None, valid_ratios is not None)
        ctx.in_dtypes = (value.dtype, boxes.dtype, angles.dtype if angles is not None else None, attention_weights.dtype)
        with torch.autocast("cuda", enabled=False):
            v = value.to(torch.bfloat16).contiguous()
            bx = boxes.float().contiguous()
            an = angles.float().contiguous() if angles is not None else None
            vr = valid_ratios.float().contiguous() if valid_ratios is not None else None
            kidx = kernel_indices.float().contiguous()
            a = attention_weights.float().contiguous()
            out = ops.box_grid_attn_forward(v, shapes, lsi, bx, an, vr, kidx, a, im2col_step)
        ctx.save_for_backward(*([v, shapes, lsi, bx, kidx, a] + [t for t in (an, vr) if t is not None]))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        v, shapes, lsi, bx, kidx, a = ctx.saved_tensors[:6]
        rest = list(ctx.saved_tensors[6:])
        an = rest.pop(0) if ctx.has[0] else None
        vr = rest.pop(0) if ctx.has[1] else None
        with torch.autocast("cuda", enabled=False):
            gv, gb, ga, gw = ops.box_grid_attn_backward(v, shapes, lsi, bx, an, vr, kidx, a,
                                                        grad_output.to(torch.bfloat16).contiguous(), ctx.im2col_step)
        dv, db, da, dw = ctx.in_dtypes
        return (gv.to(dv), None, None, gb.to(db), ga.to(da) if ga is not None else None, None, None, gw.to(dw), None)
