"""what the op hands to its consumers under bf16 autocast.

``InstanceAttention.forward`` feeds the op's two outputs to ``out_proj`` (box_attention.py:123 ``out_proj(mask_output)``
over the (B, Nq, K, K, C) mask tensor -- 60 MB at K = 14, 241 MB at K = 28 in fp32 -- and :135 ``out_proj(output)``).
With the reference's AMP contract the op returns fp32 and autocast then converts both tensors to bf16 in front of the
GEMM (an extra read of the fp32 tensor + a write of the bf16 one).  With ``boxer_b200.set_amp_native(True)`` the op
writes ``mask_out`` / ``out`` in bf16 once and ``out_proj`` consumes them as they are; ``value_proj``'s output goes
through the one-pass mask-fill + cast epilogue.  The mask head then takes ``mask_output`` channel-last without a copy
(predictor.py:49-54).
"""
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inst_module(K, seed=0):
    import boxer_b200
    torch.manual_seed(seed)
    m = boxer_b200.InstanceAttention(256, 4, 8, K).to(DEV)
    with torch.no_grad():
        m.linear_box_weight.normal_(0, 0.02)
        m.linear_attn_weight.normal_(0, 0.05)
    m.inferencing = False
    return m


def _inputs(Nq=50, image=(200, 336), B=2, seed=1):
    from boxer_b200 import workloads as W
    g = torch.Generator(device=DEV).manual_seed(seed)
    shapes = W.fpn_levels(*image)
    sh, start = W._level_meta(shapes, torch.device(DEV))
    S = int(sh.prod(1).sum())
    q = torch.randn(B, Nq, 256, device=DEV, generator=g)
    v = torch.randn(B, S, 256, device=DEV, generator=g)
    mask = torch.zeros(B, S, dtype=torch.bool, device=DEV)
    mask[1, S // 2:] = True
    ref = W.random_boxes(B, Nq, g, torch.device(DEV))
    vr = 0.5 + 0.5 * torch.rand(B, 1, 1, 4, 1, 2, device=DEV, generator=g)
    return q, v, sh, mask, start, vr, ref


@pytest.mark.parametrize("K", [14, 28])
def test_amp_native_feeds_out_proj_bf16_without_a_round_trip(K):
    import boxer_b200
    m = _inst_module(K)
    args = _inputs()
    seen = []
    hook = m.out_proj.register_forward_pre_hook(lambda mod, inp: seen.append((inp[0].dtype, tuple(inp[0].shape))))
    vseen = []
    from boxer_b200 import box_attention as BA
    orig = BA._instance_attn

    def spy(value, *a):
        vseen.append(value.dtype)
        return orig(value, *a)

    BA._instance_attn = spy
    try:
        out32, mask32, _ = m(*args)                                  # fp32, no autocast: the yardstick
        seen.clear(); vseen.clear()
        with torch.autocast("cuda", dtype=torch.bfloat16):            # the reference's AMP contract: op in fp32
            m(*args)
        assert [d for d, _ in seen] == [torch.float32, torch.float32] and vseen == [torch.bfloat16] or vseen == [torch.float32]
        seen.clear(); vseen.clear()
        boxer_b200.set_amp_native(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out16, mask16, _ = m(*args)
    finally:
        boxer_b200.set_amp_native(False)
        BA._instance_attn = orig
        hook.remove()
    # the op received bf16 value (one-pass epilogue) and handed bf16 straight to both out_proj calls
    assert vseen == [torch.bfloat16]
    assert [d for d, _ in seen] == [torch.bfloat16, torch.bfloat16]
    assert seen[0][1] == (2, 50, K, K, 256) and seen[1][1] == (2, 50, 256)
    assert mask16.dtype == torch.bfloat16 and out16.dtype == torch.bfloat16
    assert helpers.rel_err(out16.float(), out32) <= 2e-2 and helpers.rel_err(mask16.float(), mask32) <= 2e-2


def test_mask_output_reaches_the_mask_head_without_a_copy():
    """SegmentMLP.forward (predictor.py:49-54) permutes mask_output to NCHW and copies it; the permuted view already
    IS a channels_last tensor, which cuDNN's (transposed) convolutions take natively -- same numbers, no copy."""
    m = _inst_module(14)
    out, roi, _ = m(*_inputs(Nq=20))
    n, l, s, _, c = roi.shape
    x = roi.view(-1, s, s, c).permute(0, 3, 1, 2)
    assert x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
    xc = x.contiguous(memory_format=torch.channels_last)
    assert xc.data_ptr() == roi.data_ptr()                            # no copy: the op's own storage
    torch.manual_seed(3)
    head = torch.nn.Sequential(torch.nn.ConvTranspose2d(c, 32, 2, stride=2), torch.nn.ReLU(), torch.nn.Conv2d(32, 1, 1)).to(DEV)
    with torch.no_grad():
        a = head(xc)
        b = head(x.contiguous())                                      # the reference's NCHW copy
    assert helpers.rel_err(a, b) <= 1e-5
