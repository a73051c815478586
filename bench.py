#!/usr/bin/env python
"""bench.py -- box-attn fwd+bwd Gsamples/s at the COCO 4-level encoder config (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--variants]

A *step* is one forward + one backward pass of the box-attention op over one batch of synthetic
multi-scale feature maps (BASELINE.json configs[1]: 4 FPN levels of a 1333x800 image, C=256,
8 heads, Nq = S = 22 223 encoder queries, 4x4 grid, fp32, B=1 image per GPU).  A *sample* is
one bilinear sample point for one (image, query, head, level, point): N = B*Nq*H*L*P per pass.

Printed (rank 0, ONE JSON line on stdout):
  value      whole-job Gsamples/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API (BoxAttnFunction.apply + backward) with HOST
             (pinned) inputs: H2D of value/loc/weights/grad_out and D2H of out + all gradients
             inside the timed region
  roofline   dominant kernel (backward) : algorithmic bytes (a no-reuse model) / its CUDA-event
             time vs the measured HBM copy bandwidth of MEASURED_PEAKS.json; roofline_fwd likewise
  roofline_compulsory  what can physically be held against HBM: compulsory bytes (inputs read once, outputs written
             once) and ncu-measured DRAM bytes over the same kernel times
  trained_like_locations / uniform_locations   the same sizes with the two other location distributions
  reference_cuda  the reference's own CUDA kernels (oracle/_ref) at this config, timed in a subprocess
  per_rank_ms_per_step  min / median / max over ranks of the timed regions
  cpu_baseline  the reference's pure-PyTorch grid_sample formulation (oracle/plain.py) on the
             host CPU, all threads, full config, median of 3 passes
--variants adds (stderr + gpurun_out/variants.json): the other BASELINE configs, the fused entry points, the
reference's unmodified 6+6-layer BoxTransformer on the drop-in (configs[2]; baseline/ref_layers.py), module timings.

--impl reference times that CPU formulation as the "reference arm": the reference has no CPU op
(box_attn.h:53) and its CUDA extension cannot be installed offline against torch 2.11, so its own
test oracle (tests/box_attn_test.py:9-42), restated in oracle/plain.py, is the CPU implementation.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "box-attn fwd+bwd Gsamples/s @ COCO 4-level"
UNIT = "Gsamples/s"
CFG = dict(workload="BoxeR-2D encoder box-attn: 4 FPN levels (100x167,50x84,25x42,13x21) of 1333x800, C=256, "
                    "Nq=S=22223, 8 heads, 4x4 grid, fp32, fwd+bwd",
           B_per_gpu=1, K=4, heads=8, head_dim=32, levels=4, Nq=22223, dist="box")


def ncu_traffic():
    """Measured DRAM bytes per launch of the two kernels, from the committed ncu capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get("fwd_bytes"), t.get("bwd_bytes"), t.get("source")
    except Exception:
        return None, None, None


def ncu_issue():
    """Issue-slot utilisation (%) of the two kernels in the committed ncu capture: the resource that actually binds."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get("fwd_issue_active_pct"), t.get("bwd_issue_active_pct")
    except Exception:
        return None, None


def l2_reduction_peak():
    """Payload rate (GB/s) of `red.global.add.v4.f32` at the backward scatter's access shape: measured live with the
    microbenchmark binary when it is there (scripts/microbench/red_throughput, built by __graft_entry__.build()), else
    the committed measurement in profiles/traffic.json."""
    exe = os.path.join(ROOT, "scripts", "microbench", "red_throughput")
    try:
        if os.access(exe, os.X_OK):
            out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
            return float(json.loads(out)["red_v4_random_22.8MB"]), "measured live (scripts/microbench/red_throughput)"
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return float(t["l2_red_peak_gbs"]), t.get("l2_red_peak_source")
    except Exception:
        return None, None


def ncu_red_bytes():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t.get("bwd_red_bytes"), t.get("bwd_red_bytes_source")
    except Exception:
        return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:      # nvidia-smi takes a moment to start
                time.sleep(0.02)
            self.idle_rows = len(self.rows)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows[getattr(self, "idle_rows", 0):]:      # samples taken while the GPU was under load
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ our arm
def make_sets(dev, n_sets, seed0, K=4, dist="box", B=1):
    from boxer_b200 import workloads as W
    sets = []
    for i in range(n_sets):
        w = W.coco_encoder(B=B, K=K, dist=dist, device=dev, seed=seed0 + i)
        go = torch.randn(B, w.value.shape[1], w.value.shape[2] * w.value.shape[3], device=dev)
        sets.append((w, go))
    return sets


def step_resident(ops, w, go):
    """fwd + bwd straight through the tensor-level API (= what autograd calls), device-resident inputs."""
    out = ops.box_attn_forward(w.value, w.shapes, w.level_start, w.loc, w.weights[0], 64)
    n = ops.last_launch_count()
    grads = ops.box_attn_backward(w.value, w.shapes, w.level_start, w.loc, w.weights[0], go, 64)
    return out, grads, n + ops.last_launch_count()


PER_RANK_MS = {}       # label -> this run's per-rank device times (filled under torchrun; reported in the JSON line)


def max_over_ranks(ms: float, device, label: str = "") -> float:
    """A multi-GPU number is the slowest rank's device time (never wall clock, never the mean)."""
    import torch.distributed as dist
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(allt, t)
    vals = [float(x.item()) for x in allt]
    if label:
        PER_RANK_MS[label] = vals
    return max(vals)


def rank_seed(rank: int, base: int = 3) -> int:
    """Images shard over ranks (no collective in the op): every rank draws its own images."""
    return base + 10 * rank


def timed(fn, steps, warmup, dist_on):
    """W warm-up steps, then exactly K steps between barrier+synchronize pairs, CUDA events; max over ranks."""
    import torch.distributed as dist
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist_on:
        ms = max_over_ranks(ms, "cuda", "resident")
    else:
        PER_RANK_MS["resident"] = [ms]
    return ms


def kernel_times(ops, sets, reps):
    """Average CUDA-event duration of the forward call and of the backward call (memset + kernel)."""
    f_ms, b_ms = [], []
    for i in range(reps):
        w, go = sets[i % len(sets)]
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        ops.box_attn_forward(w.value, w.shapes, w.level_start, w.loc, w.weights[0], 64)
        b.record()
        ops.box_attn_backward(w.value, w.shapes, w.level_start, w.loc, w.weights[0], go, 64)
        c.record()
        torch.cuda.synchronize()
        f_ms.append(a.elapsed_time(b))
        b_ms.append(b.elapsed_time(c))
    return statistics.mean(f_ms), statistics.mean(b_ms)


def gpu_local_cpus(index):
    """CPUs on the NUMA node the GPU hangs off (sysfs), or None.  Pinned host buffers are placed by first
    touch, so the e2e leg runs its host side there: H2D / D2H then do not cross the socket interconnect."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


def e2e_run(sets, steps, warmup, dist_on):
    """Public API with HOST buffers.  Every step: pinned host tensors -> H2D (value, loc, weights, grad_out)
    -> BoxAttnFunction.apply + backward -> D2H (out, grad_value, grad_loc, grad_attn).
    The three stages run on three streams and are double-buffered, so step i+1's upload and step i-1's
    download overlap step i's kernels (what a CUDA-stream prefetcher does, cf. the reference's
    dataset/helper/prefetcher.py:11-54); all copies of all K steps are inside the timed region."""
    import torch.distributed as dist
    import boxer_b200
    dev = sets[0][0].value.device
    w0 = sets[0][0]
    host_in = [tuple(t.detach().cpu().pin_memory() for t in (w.value, w.loc, w.weights[0], go)) for w, go in sets]
    dev_in = [tuple(torch.empty_like(t, device=dev) for t in host_in[0]) for _ in range(2)]
    host_out = [[torch.empty_like(t, device="cpu").pin_memory() for t in (sets[0][1], w0.value, w0.loc, w0.weights[0])]
                for _ in range(2)]                                              # out, gV, gLoc, gW
    h2d = sum(t.numel() * t.element_size() for t in host_in[0])
    d2h = sum(t.numel() * t.element_size() for t in host_out[0])
    s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    up_done = [torch.cuda.Event() for _ in range(2)]
    comp_done = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        slot = i % 2
        with torch.cuda.stream(s_up):
            s_up.wait_event(comp_done[slot])          # the kernels of step i-2 are done with this slot
            for d, h in zip(dev_in[slot], host_in[i % len(host_in)]):
                d.copy_(h, non_blocking=True)
            up_done[slot].record(s_up)

    def compute_and_download(i):
        slot = i % 2
        main.wait_event(up_done[slot])
        v, l, a, g = dev_in[slot]
        v = v.detach().requires_grad_(True)
        l = l.detach().requires_grad_(True)
        a = a.detach().requires_grad_(True)
        out = boxer_b200.BoxAttnFunction.apply(v, w0.shapes, w0.level_start, l, a, 64)
        out.backward(g)
        comp_done[slot].record(main)
        res = (out.detach(), v.grad, l.grad, a.grad)
        with torch.cuda.stream(s_down):
            s_down.wait_event(comp_done[slot])
            for h, d in zip(host_out[slot], res):
                d.record_stream(s_down)
                h.copy_(d, non_blocking=True)

    def run(n, first):
        upload(first)
        for i in range(first, first + n):
            if i + 1 < first + n:
                upload(i + 1)
            compute_and_download(i)
        main.wait_stream(s_down)

    for e in comp_done:
        e.record(main)
    run(warmup, 0)
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    run(steps, warmup)
    e1.record(main)                                   # after main has waited for the last download
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist_on:
        ms = max_over_ranks(ms, "cuda", "e2e")
    else:
        PER_RANK_MS["e2e"] = [ms]
    return ms, h2d, d2h


# ------------------------------------------------------------------------------------ CPU arm
def cpu_pass(cpu_w, go, frac=1.0):
    """One fwd+bwd of the reference's grid_sample formulation on the host CPU; returns (seconds, samples)."""
    from oracle import plain
    value, loc, attn = cpu_w.value, cpu_w.loc, cpu_w.weights[0]
    nq = max(1, int(round(loc.shape[1] * frac)))
    value = value.clone().requires_grad_(True)
    loc = loc[:, :nq].clone().requires_grad_(True)
    attn = attn[:, :nq].clone().requires_grad_(True)
    g = go[:, :nq]
    B, S = value.shape[:2]
    t0 = time.perf_counter()
    out = plain.plain_box_attn(value.view(B, S, -1), cpu_w.shapes, 2 * loc - 1, attn)
    out.backward(g)
    dt = time.perf_counter() - t0
    n = B * nq * loc.shape[2] * loc.shape[3] * loc.shape[4]
    return dt, n


def cpu_workload(K=4):
    from boxer_b200 import workloads as W
    torch.set_num_threads(os.cpu_count() or 1)
    w = W.coco_encoder(B=1, K=K, dist="box", device="cpu", seed=3)
    go = torch.randn(1, w.value.shape[1], 256)
    return w, go


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, rank):
    if rank != 0:
        return
    w, go = cpu_workload(CFG["K"])
    cores = torch.get_num_threads()
    frac = 1.0      # every step is one pass over the FULL config (all queries): ~0.8 s on the box's host cores,
                    # so the driver's 20 + 5 steps take ~20 s and `config` equals the repo arm's
    for _ in range(args.warmup):
        cpu_pass(w, go, frac)
    t, n = 0.0, 0
    for _ in range(args.steps):
        dt, ns = cpu_pass(w, go, frac)
        t += dt
        n += ns
    val = n / t / 1e9
    sample = (f"each step = fwd+bwd of the reference's grid_sample formulation (oracle/plain.py) over all {CFG['Nq']} queries "
              f"(4 levels, K=4, fp32) = the full config, {args.steps} steps, {cores} threads, {cpu_model()}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(CFG),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def compulsory(n_samples, d, ms, peak, backward, dram_bytes):
    """Bytes that must cross HBM once for one launch (fp32): value + loc + weights (+ grad_out) in, out (or the three
    gradients, with grad_value zero-filled first) out."""
    rows = d["B"] * d["Nq"] * d["H"]
    value = d["B"] * d["S"] * d["H"] * d["D"] * 4
    loc, wts, out = n_samples * 8, n_samples * 4, rows * d["D"] * 4
    b = value + loc + wts + (out + 2 * value + loc + wts if backward else out)
    r = {"bytes": b, "GBs": b / (ms * 1e-3) / 1e9, "frac": b / (ms * 1e-3) / 1e9 / peak, "ms_per_launch": ms}
    if dram_bytes:
        r["measured_dram_bytes"] = dram_bytes
        r["frac_measured_dram"] = dram_bytes / (ms * 1e-3) / 1e9 / peak
    return r


def reference_cuda_subprocess():
    """fwd / bwd time of the UNMODIFIED reference CUDA kernels (oracle/_ref, compiled for sm_100a from /root/reference by
    oracle/build_ref.py) at the headline config, measured in a SEPARATE process so that this process never loads
    anything but libboxattn_b200.so.  None when the comparison build is absent."""
    script = os.path.join(ROOT, "scripts", "compare_reference_cuda.py")
    out = os.path.join(ROOT, "gpurun_out", "reference_cuda_headline.json")
    try:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        proc = subprocess.run([sys.executable, script, out, "--headline-only"], capture_output=True, text=True, timeout=240)
        if proc.returncode != 0:
            return {"unavailable": (proc.stderr or proc.stdout).strip().splitlines()[-1][:200] if (proc.stderr or proc.stdout) else "failed"}
        with open(out) as f:
            r = json.load(f)["enc_K4_box"]
        return {"fwd_ms": r["ref_fwd_ms"], "bwd_ms": r["ref_bwd_ms"], "fwdbwd_Gsamples_per_s": r["n_samples"] / (r["ref_fwd_ms"] + r["ref_bwd_ms"]) / 1e6,
                "ours_same_process": {"fwd_ms": r["ours_fwd_ms"], "bwd_ms": r["ours_bwd_ms"]},
                "speedup_fwd": r["speedup_fwd"], "speedup_bwd": r["speedup_bwd"], "speedup_fwdbwd": r["speedup_fwdbwd"],
                "how": "subprocess scripts/compare_reference_cuda.py --headline-only: the reference's own box_attn_forward / "
                       "box_attn_backward (vision.cpp:8-9) from oracle/_ref on the same GPU, same tensors, CUDA events, mean of 20 / 5 launches"}
    except Exception as e:       # a comparison point, never a reason to lose the bench line
        return {"unavailable": str(e)[:200]}


def model_leg(dev):
    """BASELINE configs[2] on one GPU's share: the reference's UNMODIFIED BoxTransformer (6 encoder + 6 decoder layers,
    d_model 256, 8 heads, 4 levels, FFN 1024, 300 queries, use_mask -> InstanceAttention 14 x 14 in every decoder
    layer; box_transformer.py:16-465, loaded from the bytecode baseline/build_ref_layers.py compiled) running on the
    drop-in, forward + backward, B = 2 images per GPU padded to the batch maximum with masks (collate_fn.py:66-84),
    synthetic 800 x 1333 feature pyramids.  Three precisions: fp32; bf16 autocast with the reference's AMP contract (the
    op force-casts to fp32, box_attention_func.py:11); bf16 autocast with boxer_b200.set_amp_native (bf16 value / outputs).
    Reports ms per iteration and the share of it spent inside the four native ops (CUDA events around every call)."""
    import boxer_b200
    from boxer_b200 import ops as ops_mod
    from boxer_b200 import workloads as W
    try:
        from baseline import ref_layers as R
        if not R.available():
            return {"unavailable": "baseline/_ref not built (python -m baseline.build_ref_layers where /root/reference exists)"}
        bt, _ = R.import_layers("boxer_b200")
    except Exception as e:
        return {"unavailable": str(e)[:200]}
    model = R.make_boxer2d(bt, d_model=256, nhead=8, nlevel=4, enc=6, dec=6, ffn=1024, num_queries=300, use_mask=True).to(dev)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():        # trained-like: non-zero box / attention projections
        for n, p_ in model.named_parameters():
            if "linear_box_weight" in n or "linear_attn_weight" in n:
                p_.copy_((0.02 * torch.randn(p_.shape, generator=g)).to(dev))
    src, mask, pos = R.padded_batch(W.fpn_levels(), 2, 256, valid=[(1.0, 1.0), (0.8, 0.75)], device=dev)
    names = ("box_attn_forward", "box_attn_backward", "instance_attn_forward", "instance_attn_backward")
    orig = {n: getattr(ops_mod, n) for n in names}
    events = []

    def wrap(n):
        def f(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig[n](*a, **k)
            e1.record()
            events.append((n, e0, e1))
            return r
        return f

    def step(amp, native):
        boxer_b200.set_amp_native(native)
        for p_ in model.parameters():
            p_.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            hs, roi = model(src, mask, pos)[:2]
            loss = hs.float().square().mean() + roi.float().square().mean()
        loss.backward()

    out = {"config": "6+6 layers, d_model 256, 8 heads, FFN 1024, 300 queries, use_mask, B=2 (second image 0.8 x 0.75 of the padded size), "
                     "S=22223, encoder K=2, decoder InstanceAttention K=14; reference layers unmodified on boxer_b200"}
    try:
        for label, amp, native in (("fp32", False, False), ("bf16_autocast_reference_amp", True, False), ("bf16_autocast_native", True, True)):
            for _ in range(2):
                step(amp, native)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            a.record()
            for _ in range(reps):
                step(amp, native)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            for n in names:
                setattr(ops_mod, n, wrap(n))
            events.clear()
            step(amp, native)
            torch.cuda.synchronize()
            for n in names:
                setattr(ops_mod, n, orig[n])
            per = {}
            for n, e0, e1 in events:
                per.setdefault(n, [0, 0.0])
                per[n][0] += 1
                per[n][1] += e0.elapsed_time(e1)
            op_ms = sum(v[1] for v in per.values())
            out[label] = {"ms_per_iteration": ms, "native_op_ms": op_ms, "native_op_share": op_ms / ms,
                          "ops": {n: {"calls": v[0], "ms": v[1]} for n, v in per.items()},
                          "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
    except Exception as e:
        out["error"] = str(e)[:300]
    finally:
        for n in names:
            setattr(ops_mod, n, orig[n])
        boxer_b200.set_amp_native(False)
    del model
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------ main
def variants(ops, dev, bw_peak):
    """Extra measurements (stderr + gpurun_out/variants.json); not part of the contract line."""
    from boxer_b200 import workloads as W
    res = {}
    red_peak, _ = l2_reduction_peak()

    def red_frac(w, P, ms, instance=False):
        """backward against the L2 reduction rate (see roofline_l2_reduction): window kernels merge a (row, level)'s corners per
        unique pixel where the range fits the window (64 slots, 32 when two levels share a pass: P <= 4); the instance
        kernels issue one reduction per in-range corner (cap=0)."""
        if not red_peak:
            return None
        nbytes = W.backward_reduction_bytes(w, cap=0 if instance else (32 if P <= 4 else 64))
        return nbytes / (ms * 1e-3) / 1e9 / red_peak

    def time_call(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    for K in (2, 4):
        for dist in ("box", "uniform"):
            for dt in (torch.float32, torch.bfloat16):
                w = W.coco_encoder(K=K, dist=dist, device=dev)
                v = w.value.to(dt)
                go = torch.randn(1, v.shape[1], 256, device=dev, dtype=dt)
                a = w.weights[0]
                tf = time_call(lambda: ops.box_attn_forward(v, w.shapes, w.level_start, w.loc, a, 64))
                tb = time_call(lambda: ops.box_attn_backward(v, w.shapes, w.level_start, w.loc, a, go, 64))
                n = w.n_samples
                sz = v.element_size()
                bf = W.bytes_per_sample(False, False, 32, 4, K * K, sz)
                bb = W.bytes_per_sample(False, True, 32, 4, K * K, sz)
                res[f"enc_K{K}_{dist}_{'f32' if dt == torch.float32 else 'bf16'}"] = {
                    "fwd_ms": tf, "bwd_ms": tb, "fwd_Gs": n / tf / 1e6, "fwdbwd_Gs": n / (tf + tb) / 1e6,
                    "fwd_frac": n * bf / (tf * 1e-3) / 1e9 / bw_peak, "bwd_frac": n * bb / (tb * 1e-3) / 1e9 / bw_peak,
                    "bwd_l2_reduction_frac": red_frac(w, K * K, tb)}
    for K in (14, 28):
        for dt in (torch.float32, torch.bfloat16):
            m = W.coco_mask_head(K=K, device=dev)
            v = m.value.to(dt)
            go = torch.randn(1, 300, 256, device=dev, dtype=dt)
            gm = torch.randn(1, 300, K * K, 256, device=dev, dtype=dt)
            sw, lw = m.weights
            tf = time_call(lambda: ops.instance_attn_forward(v, m.shapes, m.level_start, m.loc, sw, lw, 64))
            tb = time_call(lambda: ops.instance_attn_backward(v, m.shapes, m.level_start, m.loc, sw, lw, go, gm, 64))
            n = m.n_samples
            sz = v.element_size()
            bf = W.bytes_per_sample(True, False, 32, 4, K * K, sz)
            bb = W.bytes_per_sample(True, True, 32, 4, K * K, sz)
            res[f"mask_K{K}_{'f32' if dt == torch.float32 else 'bf16'}"] = {
                "fwd_ms": tf, "bwd_ms": tb, "fwd_Gs": n / tf / 1e6, "fwdbwd_Gs": n / (tf + tb) / 1e6,
                "fwd_frac": n * bf / (tf * 1e-3) / 1e9 / bw_peak, "bwd_frac": n * bb / (tb * 1e-3) / 1e9 / bw_peak,
                "bwd_l2_reduction_frac": red_frac(m, K * K, tb, instance=True)}
    r = W.bev_rotated(B=8, device=dev)
    go = torch.randn(8, 1000, 128, device=dev)
    tf = time_call(lambda: ops.box_attn_forward(r.value, r.shapes, r.level_start, r.loc, r.weights[0], 64))
    tb = time_call(lambda: ops.box_attn_backward(r.value, r.shapes, r.level_start, r.loc, r.weights[0], go, 64))
    res["bev_B8"] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_Gs": r.n_samples / tf / 1e6, "fwdbwd_Gs": r.n_samples / (tf + tb) / 1e6}
    w = W.coco_encoder(K=4, dist="box", device=dev)
    go = torch.randn(1, w.value.shape[1], 256, device=dev)
    ops.set_deterministic(True)
    tb = time_call(lambda: ops.box_attn_backward(w.value, w.shapes, w.level_start, w.loc, w.weights[0], go, 64), reps=5)
    ops.set_deterministic(None)
    res["enc_K4_box_f32_deterministic"] = {"bwd_ms": tb}

    # module level: BoxAttention (projections + softmax + grid + op) forward+backward, reference-style grid
    # materialisation vs the fused box->grid op
    import boxer_b200
    S = w.value.shape[1]
    for K in (2, 4):
        torch.manual_seed(0)
        mod = boxer_b200.BoxAttention(256, 4, 8, K).to(dev)
        with torch.no_grad():
            mod.linear_box_weight.normal_(0, 0.01)
            mod.linear_attn_weight.normal_(0, 0.01)
        q = torch.randn(1, S, 256, device=dev, requires_grad=True)
        src = torch.randn(1, S, 256, device=dev, requires_grad=True)
        refw = W.encoder_ref_windows(W.fpn_levels(), 1, dev)

        def run_mod():
            out, _ = mod(q, src, w.shapes, None, w.level_start, None, refw)
            out.sum().backward()

        r = {}
        for fused in (False, True):
            boxer_b200.set_fused_grid(fused)
            r["fused_ms" if fused else "grid_ms"] = time_call(run_mod, reps=10)
        boxer_b200.set_fused_softmax(True)          # + softmax in the op (row f2)
        r["fused_softmax_ms"] = time_call(run_mod, reps=10)
        boxer_b200.set_fused_softmax(False)
        boxer_b200.set_fused_grid(False)
        r["speedup"] = r["grid_ms"] / r["fused_ms"]
        r["speedup_softmax"] = r["grid_ms"] / r["fused_softmax_ms"]
        res[f"module_BoxAttention_K{K}_fwdbwd"] = r
        # op level: (softmax kernel + grid + loc-taking op) vs fused grid vs fused grid + softmax, fwd and bwd
        with torch.no_grad():
            boxes, _ = mod._boxes_and_angles(q, refw)
            boxes = boxes.contiguous()
            logits = torch.randn(1, S, 8, 4, K, K, device=dev)
            attn = torch.softmax(logits.view(1, S, 8, -1), -1).view_as(logits)
            kidx = mod.kernel_indices
            v4 = w.value
            go_m = torch.randn(1, S, 256, device=dev)
            o = {}
            o["fused_fwd_ms"] = time_call(lambda: ops.box_grid_attn_forward(v4, w.shapes, w.level_start, boxes, None, None, kidx, attn, 64))
            o["fused_softmax_fwd_ms"] = time_call(lambda: ops.box_grid_attn_forward(v4, w.shapes, w.level_start, boxes, None, None, kidx, logits, 64, softmax=True))
            o["fused_bwd_ms"] = time_call(lambda: ops.box_grid_attn_backward(v4, w.shapes, w.level_start, boxes, None, None, kidx, attn, go_m, 64))
            o["fused_softmax_bwd_ms"] = time_call(lambda: ops.box_grid_attn_backward(v4, w.shapes, w.level_start, boxes, None, None, kidx, attn, go_m, 64, softmax=True))
            o["torch_softmax_fwd_ms"] = time_call(lambda: torch.softmax(logits.view(1, S, 8, -1), -1))
        res[f"op_box_grid_K{K}"] = o

    res["model_configs2_BoxTransformer"] = model_leg(dev)

    # InstanceAttention module (mask head input path, rows f3 / f4): fwd+bwd, fp32 vs autocast with the reference's AMP
    # contract (op in fp32, two fp32 -> bf16 conversions in front of out_proj) vs set_amp_native (bf16 written once)
    for K in (14, 28):
        torch.manual_seed(0)
        im = boxer_b200.InstanceAttention(256, 4, 8, K).to(dev)
        im.inferencing = False
        with torch.no_grad():
            im.linear_box_weight.normal_(0, 0.02)
            im.linear_attn_weight.normal_(0, 0.05)
        qi = torch.randn(1, 300, 256, device=dev, requires_grad=True)
        mem = torch.randn(1, S, 256, device=dev, requires_grad=True)
        gen = torch.Generator(device=dev).manual_seed(5)
        refb = W.random_boxes(1, 300, gen, dev)
        r = {}
        for label, amp, native in (("fp32_ms", False, False), ("autocast_reference_amp_ms", True, False), ("autocast_native_ms", True, True)):
            boxer_b200.set_amp_native(native)

            def run_inst():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    o, mo, _ = im(qi, mem, w.shapes, None, w.level_start, None, refb)
                    loss = o.float().sum() + mo.float().sum()
                loss.backward()

            r[label] = time_call(run_inst, reps=10)
        boxer_b200.set_amp_native(False)
        r["native_vs_reference_amp"] = r["autocast_reference_amp_ms"] / r["autocast_native_ms"]
        res[f"module_InstanceAttention_K{K}_fwdbwd"] = r
        del im, qi, mem

    # warm vs L2-flushed: one launch at a time, a 512 MB write in between evicts value / loc from L2
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def cold(fn, reps=10):
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    a0 = (w.value, w.shapes, w.level_start, w.loc, w.weights[0])
    res["enc_K4_box_f32_L2_flushed"] = {
        "fwd_ms": cold(lambda: ops.box_attn_forward(*a0, 64)),
        "bwd_ms": cold(lambda: ops.box_attn_backward(*a0, go, 64)),
        "note": "median of 10 single launches, each after a 512 MB fill (L2 flushed); compare enc_K4_box_f32 (back-to-back)"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variants", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries the JSON line only: NCCL prints its version banner on fd 1 while the communicator comes up
        # (init is eager with device_id; the barrier covers a lazy one), so fd 1 points at stderr for that long
        sys.stdout.flush()
        saved = os.dup(1)
        try:
            os.dup2(2, 1)
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    n_gpus = world

    from boxer_b200 import _native
    if _native.is_stale() and rank == 0:      # fresh checkout: compile the CUDA library first (no fallback exists)
        _native.build()
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
    import boxer_b200
    from boxer_b200 import ops
    from boxer_b200 import workloads as W

    bw_peak, peak_src = peaks()
    sets = make_sets(dev, 2, seed0=rank_seed(rank), K=CFG["K"], dist=CFG["dist"], B=CFG["B_per_gpu"])
    n_samples = sets[0][0].n_samples
    launches = [0]

    def fn(i):
        w, go = sets[i % len(sets)]
        _, _, n = step_resident(ops, w, go)
        launches[0] += n

    for i in range(3):
        fn(i)
    launches[0] = 0
    e2e_steps = max(3, min(args.steps, 20))
    with ClockSampler(local_rank) as clk:
        # the sampler spans every measured region (the K-step region alone lasts ~30 ms, shorter than
        # nvidia-smi's sampling period); a short sustained loop first so that several samples see load
        t_end = time.time() + 0.6
        while time.time() < t_end:
            for i in range(20):
                fn(i)
            torch.cuda.synchronize()
        launches[0] = 0
        ms = timed(fn, args.steps, args.warmup, dist_on)
        n_launch = launches[0] - 2 * args.warmup      # kernels inside the timed region
        kf_ms, kb_ms = kernel_times(ops, sets, max(10, min(args.steps, 50)))
        all_cpus = os.sched_getaffinity(0)
        local = gpu_local_cpus(local_rank)
        if local:
            os.sched_setaffinity(0, local)
        try:
            ms_e2e, h2d, d2h = e2e_run(sets, e2e_steps, 5, dist_on)
        finally:
            os.sched_setaffinity(0, all_cpus)
    clocks = clk.summary()
    value = n_gpus * n_samples * args.steps / (ms * 1e-3) / 1e9
    e2e_val = n_gpus * n_samples * e2e_steps / (ms_e2e * 1e-3) / 1e9

    d = sets[0][0].dims
    bf = W.bytes_per_sample(False, False, d["D"], d["L"], d["P"], 4)
    bb = W.bytes_per_sample(False, True, d["D"], d["L"], d["P"], 4)
    ach_f = n_samples * bf / (kf_ms * 1e-3) / 1e9
    ach_b = n_samples * bb / (kb_ms * 1e-3) / 1e9
    ach_s = n_samples * (bf + bb) / ((kf_ms + kb_ms) * 1e-3) / 1e9

    tr_f, tr_b, tr_src = ncu_traffic()
    is_f, is_b = ncu_issue()
    # the second location distribution of SURVEY.md 8(d): uniform-random points (worst-case locality, what the
    # reference's own tests draw) -- reported beside the headline, same sizes
    uni = make_sets(dev, 1, seed0=rank_seed(rank) + 100, K=CFG["K"], dist="uniform", B=CFG["B_per_gpu"])
    uf_ms, ub_ms = kernel_times(ops, uni, 10)
    uf_ms, ub_ms = kernel_times(ops, uni, 20)
    red_uni = W.backward_reduction_bytes(uni[0][0])
    del uni
    # third distribution: trained-like encoder boxes (per-query offsets, sizes log-uniform in 2..64 px) -- between the
    # init-state headline (every window fits the 64-pixel footprint) and uniform points (none does)
    trn = make_sets(dev, 1, seed0=rank_seed(rank) + 200, K=CFG["K"], dist="trained", B=CFG["B_per_gpu"])
    tf_ms, tb_ms = kernel_times(ops, trn, 10)
    tf_ms, tb_ms = kernel_times(ops, trn, 20)
    win_frac = {"box": W.window_mode_fraction(sets[0][0]), "trained": W.window_mode_fraction(trn[0][0])}
    red_trn = W.backward_reduction_bytes(trn[0][0])
    red_box = W.backward_reduction_bytes(sets[0][0])
    del trn

    red_peak, red_peak_src = l2_reduction_peak() if rank == 0 else (None, None)

    def l2_red(ms, nbytes, full=False):
        """The backward against the L2 reduction rate: payload bytes of the `red.v4`s it issues (counted on the host from
        the workload, boxer_b200.workloads.backward_reduction_bytes) over the live launch time, against the microbenchmark."""
        if not red_peak:
            return None
        ach = nbytes / (ms * 1e-3) / 1e9
        r = {"achieved": ach, "frac": ach / red_peak, "red_bytes": nbytes, "ms_per_launch": ms}
        if full:
            ncu_bytes, bsrc = ncu_red_bytes()
            r = dict({"bound": "l2_reduction", "kernel": "box_bwd_win_kernel<float,G=8,SUB=8,PPL=2,atomic>", "peak": red_peak,
                      "unit": "GB/s", "peak_source": red_peak_src}, **r,
                     red_bytes_ncu=ncu_bytes, red_bytes_ncu_source=bsrc,
                     note="the scatter alone would take red_bytes / peak; an ablation build without it runs 0.239 ms "
                          "(profiles/r02xy_ablation.json): the backward sits between its two bounds, reductions and issue slots")
        return r

    def dram(tr, ms):      # measured DRAM bytes of the committed ncu capture over the live launch time
        return None if tr is None else {"GBs": tr / (ms * 1e-3) / 1e9, "frac_of_peak": tr / (ms * 1e-3) / 1e9 / bw_peak}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": dict(CFG, n_samples_per_gpu_step=n_samples,
                       l2="working set ~364 MB/step (> 126 MB L2); 2 rotating input sets; no explicit flush",
                       locations="box-structured (encoder reference windows + init-state offsets), see boxer_b200/workloads.py"),
        "roofline": {"bound": "hbm", "kernel": "box_bwd_win_kernel<float,G=8,SUB=8,PPL=2,atomic> (+ grad_value memset)",
                     "achieved": ach_b, "peak": bw_peak, "unit": "GB/s", "frac": ach_b / bw_peak, "traffic": tr_b,
                     "traffic_source": tr_src, "algorithmic_bytes": n_samples * bb,
                     "peak_source": peak_src, "bytes_per_sample": bb, "ms_per_launch": kb_ms, "dram": dram(tr_b, kb_ms),
                     "issue_active_pct_ncu": is_b,
                     "note": "frac > 1: the no-reuse byte model of SURVEY.md 8(d) counts every corner row as HBM traffic; "
                             "value (22.8 MB) lives in L1/L2, the measured DRAM traffic is `traffic` (see `dram`), and what "
                             "binds is instruction issue + gather latency (profiles/README.md)"},
        "roofline_fwd": {"bound": "hbm", "kernel": "box_fwd_win_kernel<float,G=8,SUB=8,PPL=2>", "achieved": ach_f, "peak": bw_peak,
                         "unit": "GB/s", "frac": ach_f / bw_peak, "traffic": tr_f, "algorithmic_bytes": n_samples * bf, "bytes_per_sample": bf,
                         "ms_per_launch": kf_ms, "Gsamples_per_s": n_samples / kf_ms / 1e6, "dram": dram(tr_f, kf_ms), "issue_active_pct_ncu": is_f},
        "uniform_locations": {"fwd_ms": uf_ms, "bwd_ms": ub_ms, "fwdbwd_Gsamples_per_s": n_samples / (uf_ms + ub_ms) / 1e6,
                              "bwd_l2_reduction": l2_red(ub_ms, red_uni),
                              "note": "same sizes, sampling points drawn uniformly in [0,1)^2 (no spatial structure)"},
        "trained_like_locations": {"fwd_ms": tf_ms, "bwd_ms": tb_ms, "fwdbwd_Gsamples_per_s": n_samples / (tf_ms + tb_ms) / 1e6,
                                   "bwd_l2_reduction": l2_red(tb_ms, red_trn),
                                   "window_mode_fraction": win_frac["trained"], "window_mode_fraction_headline": win_frac["box"],
                                   "note": "same sizes; every (query, head, level) has its own box: centre = pixel centre +- U(1/2) box, "
                                           "width / height log-uniform in 2..64 px of level 0; window_mode_fraction = share of (row, level) "
                                           "pairs whose footprint fits the 64-pixel window (the rest take the per-point walk)"},
        "roofline_step": {"achieved": ach_s, "frac": ach_s / bw_peak, "unit": "GB/s"},
        # what the kernels can physically be held against: the bytes that MUST cross HBM (compulsory: every input read
        # once, every output written once; grad_value also zero-filled) and the bytes that did (ncu, dram__bytes_*),
        # each over the live kernel time and the measured HBM peak
        "roofline_compulsory": {
            "bound": "hbm", "peak": bw_peak, "unit": "GB/s",
            "fwd": compulsory(n_samples, d, kf_ms, bw_peak, False, tr_f),
            "bwd": compulsory(n_samples, d, kb_ms, bw_peak, True, tr_b),
            "note": "frac = compulsory bytes / kernel time / peak; frac_measured_dram = ncu DRAM bytes of the committed capture / "
                    "live kernel time / peak.  These, not the no-reuse model above, say how far the kernels are from the HBM bound: "
                    "value (22.8 MB) is L2-resident, so the kernels are bound by instruction issue and gather latency"},
        # the resource the backward's scatter is bound by: one red.v4 per lane per unique pixel goes to the L2's reduction
        # units, whose payload rate is far below the L2's load bandwidth (microbenchmark: 6.4 TB/s against 17 TB/s of
        # 128-byte line gathers).  bytes = RED instructions executed (ncu, committed capture) x 16 B, identical inputs.
        "roofline_l2_reduction": l2_red(kb_ms, red_box, full=True),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                "host_cpus": ("NUMA-local to the GPU: %d cpus" % len(local)) if local else "unpinned",
                "path": "pinned host -> H2D -> BoxAttnFunction.apply + backward -> D2H(out, grad_value, grad_loc, grad_attn); "
                        "3 streams, double-buffered (upload i+1 / kernels i / download i-1 overlap)"},
        "gpu_launches": n_launch,
        "clocks": clocks,
    }

    line["per_rank_ms_per_step"] = {
        k: {"min": min(v) / n, "median": statistics.median(v) / n, "max": max(v) / n, "ranks": len(v)}
        for k, v, n in (("resident", PER_RANK_MS.get("resident", []), args.steps), ("e2e", PER_RANK_MS.get("e2e", []), e2e_steps)) if v}
    line["per_rank_note"] = ("every rank times its own images (seed 3 + 10 * rank); `value` uses the slowest rank.  Under torchrun the "
                             "timed region starts right after an NCCL barrier (idle gap + collective tail), which N = 1 does not have")
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        line["reference_cuda"] = reference_cuda_subprocess()
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cw, cgo = cpu_workload(CFG["K"])
        cpu_pass(cw, cgo, 0.25)                      # warm-up on a quarter of the queries
        runs = [cpu_pass(cw, cgo, 1.0) for _ in range(3)]
        t = statistics.median(r[0] for r in runs)
        line["cpu_baseline"] = {
            "value": runs[0][1] / t / 1e9, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"full config (all {CFG['Nq']} queries, K=4, fp32) fwd+bwd of the grid_sample formulation "
                      f"(oracle/plain.py), median of 3 after warm-up, {t * 1e3:.0f} ms/pass, {cpu_model()}"}

    if args.variants and rank == 0:
        v = variants(ops, dev, bw_peak)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w") as f:
            json.dump(v, f, indent=1)
        for k, r in v.items():
            print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in r.items()}, file=sys.stderr)

    if rank == 0:
        print(json.dumps(line))
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
