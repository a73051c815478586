"""boxer_b200 -- B200-native (sm_100a) box-attention operators behind BoxeR's own op API.

Public surface (mirrors kienduynguyen/BoxeR's ``e2edet.module.ops`` / ``e2edet.module.box_attention``):

    from boxer_b200 import BoxAttnFunction, InstanceAttnFunction            # autograd Functions
    from boxer_b200 import BoxAttention, InstanceAttention, Box3dAttention  # nn.Modules
    from boxer_b200 import ops                                             # tensor-level mirror of the pybind module
    boxer_b200.compat.install()                                            # make `import e2edet.module.ops` resolve here

Everything computes in ``boxer_b200/_C/libboxattn_b200.so`` (C ABI: include/boxattn_b200.h);
there is no CPU or PyTorch fallback.
"""
from . import _native, compat, ops
from .box_attention import (Box3dAttention, BoxAttention, InstanceAttention, set_amp_native, set_fused_grid,
                            set_fused_softmax)
from .box_attention_func import (BoxAttnBf16Function, BoxAttnFunction, BoxGridAttnBf16Function, BoxGridAttnFunction,
                                 BoxGridSoftmaxAttnBf16Function, BoxGridSoftmaxAttnFunction,
                                 InstanceAttnBf16Function, InstanceAttnFunction, InstanceWeightsFunction,
                                 ValueEpilogueFunction)
from .ops import set_deterministic

__all__ = [
    "BoxAttnFunction", "InstanceAttnFunction", "BoxAttnBf16Function", "InstanceAttnBf16Function",
    "BoxAttention", "InstanceAttention", "Box3dAttention",
    "BoxGridAttnFunction", "BoxGridAttnBf16Function", "BoxGridSoftmaxAttnFunction", "BoxGridSoftmaxAttnBf16Function", "InstanceWeightsFunction", "ValueEpilogueFunction",
    "ops", "compat", "set_deterministic", "set_amp_native", "set_fused_grid", "set_fused_softmax",
]
__version__ = "0.1.0"
