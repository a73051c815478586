"""Host-side cost per call of the Python wrappers at a decoder-sized (launch-bound) call."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boxer_b200 import ops
from boxer_b200 import workloads as W

w = W.coco_decoder(Nq=300, K=2, device="cuda")
a = (w.value, w.shapes, w.level_start, w.loc, w.weights[0])
go = torch.randn(1, 300, 256, device="cuda")
for name, fn in (("fwd", lambda: ops.box_attn_forward(*a, 64)), ("bwd", lambda: ops.box_attn_backward(*a, go, 64))):
    for _ in range(200): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3000): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: host {1e6 * (t1 - t0) / 3000:.1f} us/call, incl. drain {1e6 * (t2 - t0) / 3000:.1f} us/call")
pr = cProfile.Profile()
pr.enable()
for _ in range(3000): ops.box_attn_forward(*a, 64)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
