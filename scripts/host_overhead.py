"""Host-side cost per call at a decoder-sized (launch-bound) call: this library (pybind shim, or ctypes with
BOXER_B200_NO_SHIM=1) and, when oracle/_ref is built, the reference's own pybind function on the same tensors.
usage: host_overhead.py [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boxer_b200 import _native, ops
from boxer_b200 import workloads as W

w = W.coco_decoder(Nq=300, K=2, device="cuda")
a = (w.value, w.shapes, w.level_start, w.loc, w.weights[0])
go = torch.randn(1, 300, 256, device="cuda")
res = {"route": "pybind shim" if _native.load_shim() is not None else "ctypes"}


def measure(name, fn, n=3000):
    for _ in range(300): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    res[name] = {"host_us_per_call": 1e6 * (t1 - t0) / n, "incl_drain_us_per_call": 1e6 * (t2 - t0) / n,
                 "cuda_event_us_per_call": 1e3 * e0.elapsed_time(e1) / n}
    print(name, {k: round(v, 2) for k, v in res[name].items()}, flush=True)


measure("ours_fwd", lambda: ops.box_attn_forward(*a, 64))
measure("ours_bwd", lambda: ops.box_attn_backward(*a, go, 64))
# through autograd (what a decoder layer pays): Function.apply + backward
import boxer_b200
v = w.value.clone().requires_grad_(True); l = w.loc.clone().requires_grad_(True); at = w.weights[0].clone().requires_grad_(True)
def fb():
    out = boxer_b200.BoxAttnFunction.apply(v, w.shapes, w.level_start, l, at, 64)
    out.backward(go)
    v.grad = l.grad = at.grad = None
measure("ours_autograd_fwd_bwd", fb, n=1000)
try:
    from oracle import ref_cuda
    if ref_cuda.available():
        ref = ref_cuda.load()
        measure("reference_fwd", lambda: ref.box_attn_forward(*a, 64))
        measure("reference_bwd", lambda: ref.box_attn_backward(*a, go, 64))
except Exception as e:
    res["reference_error"] = str(e)[:200]
if len(sys.argv) > 1:
    os.makedirs(os.path.dirname(sys.argv[1]) or ".", exist_ok=True)
    json.dump(res, open(sys.argv[1], "w"), indent=1)
