"""CPU: host-side logic, the module mirror, the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from tests import helpers, refinputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI surface
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "boxattn_b200.h")).read()
    names = set(re.findall(r"\b(bxr_[a-z0-9_]+)\s*\(", text))
    names -= {n for n in names if n.endswith("_")}           # macro stems: bxr_box_attn_fwd_##SUF
    for macro, ops in (("BXR_DECLARE_OPS", ("box_attn_fwd", "box_attn_bwd", "instance_attn_fwd", "instance_attn_bwd")),
                       ("BXR_DECLARE_FUSED", ("box_grid_attn_fwd", "box_grid_attn_bwd")),
                       ("BXR_DECLARE_SMAX", ("box_grid_softmax_attn_fwd", "box_grid_softmax_attn_bwd")),
                       ("BXR_DECLARE_INSTW", ("instance_weights_fwd", "instance_weights_bwd"))):
        for op in ops:
            for suf in re.findall(macro + r"\((\w+),", text):
                if suf != "SUF":
                    names.add(f"bxr_{op}_{suf}")
    return sorted(names)


def test_library_loads_and_exports_every_declared_symbol():
    from boxer_b200 import _native
    _native.build()
    lib = _native.load()
    declared = _header_symbols()
    assert len(declared) >= 24, declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/boxattn_b200.h but not exported"
    assert set(_native.EXPORTS) == set(declared)
    assert lib.bxr_abi_version() == 1
    assert lib.bxr_status_string(0) == b"ok"
    assert lib.bxr_status_string(3).startswith(b"workspace")


def test_library_is_sm100a_native_code_only():
    from boxer_b200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_argument_validation_without_touching_the_device():
    from boxer_b200 import _native
    lib = _native.load()
    # negative dim / too many levels are rejected before any CUDA call
    st = lib.bxr_box_attn_fwd_f32(None, None, None, None, None, 1, 10, 8, 32, 4, -1, 4, None, 0, None)
    assert st == 2 and b"negative" in lib.bxr_last_error_detail()
    st = lib.bxr_box_attn_fwd_f32(None, None, None, None, None, 1, 10, 8, 32, 33, 5, 4, None, 0, None)
    assert st == 2
    # empty problem: OK with null pointers, nothing launched
    st = lib.bxr_box_attn_fwd_f32(None, None, None, None, None, 1, 10, 8, 32, 4, 0, 4, None, 0, None)
    assert st == 0 and lib.bxr_last_launch_count() == 0
    # non-empty with null pointers
    st = lib.bxr_box_attn_fwd_f32(None, None, None, None, None, 1, 10, 8, 32, 4, 5, 4, None, 0, None)
    assert st == 1
    with pytest.raises(RuntimeError, match="null pointer"):
        _native.check(st, "probe")
    # workspace sizing: fp32 atomics need none, bf16 needs an fp32 accumulator, deterministic int64 + header
    n = 2 * 100 * 8 * 32
    assert lib.bxr_attn_bwd_workspace_bytes(4, 2, 100, 8, 32, 0) == 0
    assert lib.bxr_attn_bwd_workspace_bytes(2, 2, 100, 8, 32, 0) == 4 * n
    assert lib.bxr_attn_bwd_workspace_bytes(4, 2, 100, 8, 32, 1) == 256 + 8 * n


def test_fused_entry_points_validate_without_touching_the_device():
    """rows f1-f3 of the scope table: the same status-code behaviour as the four reference-shaped entry points."""
    from boxer_b200 import _native
    lib = _native.load()
    dims = (1, 10, 8, 32, 4, 5, 4)
    # softmax -> box -> grid -> attention: attn_out is required for a non-empty problem, bad dims are rejected first
    st = lib.bxr_box_grid_softmax_attn_fwd_f32(None, None, None, None, None, None, None, None, *dims, None, None, None, 0, 0, None)
    assert st == 1 and b"attn_out" in lib.bxr_last_error_detail()
    st = lib.bxr_box_grid_softmax_attn_fwd_f32(None, None, None, None, None, None, None, None, 1, 10, 8, 32, 4, -5, 4, None, None, None, 0, 0, None)
    assert st in (1, 2)
    st = lib.bxr_box_grid_softmax_attn_bwd_f32(None, None, None, None, None, None, None, None, None, 1, 10, 8, 32, 40, 5, 4,
                                               None, None, None, None, None, 0, 0, None)
    assert st == 2                                          # 40 levels > BXR_MAX_LEVELS
    # instance weights: K must be even, L <= 32, empty problems are fine
    assert lib.bxr_instance_weights_fwd_f32(None, 10, 4, 3, None, None, None) == 2
    assert lib.bxr_instance_weights_fwd_f32(None, 10, 33, 4, None, None, None) == 2
    assert lib.bxr_instance_weights_fwd_f32(None, 0, 4, 4, None, None, None) == 0 and lib.bxr_last_launch_count() == 0
    assert lib.bxr_instance_weights_fwd_f32(None, 10, 4, 4, None, None, None) == 1
    assert lib.bxr_instance_weights_bwd_f64(None, None, None, 10, 4, 14, None, None) == 1
    # value epilogue: element sizes 4 / 2 only
    assert lib.bxr_value_epilogue(None, 8, None, None, 4, 10, 256, None) == 5
    assert lib.bxr_value_epilogue(None, 4, None, None, 2, 0, 256, None) == 0
    assert lib.bxr_value_epilogue(None, 4, None, None, 2, 10, 256, None) == 1
    # the fused ops' workspace: nothing for the fused kernels' shapes in fp32 forward, the grid otherwise
    assert lib.bxr_box_grid_attn_workspace_bytes(4, 0, 1, 100, 8, 32, 4, 50, 16, 0) == 0
    assert lib.bxr_box_grid_attn_workspace_bytes(8, 0, 1, 100, 8, 32, 4, 50, 16, 0) >= 50 * 8 * 4 * 16 * 2 * 8


def test_fused_python_wrappers_reject_cpu_tensors_and_bad_shapes():
    import boxer_b200
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        boxer_b200.ops.instance_weights_forward(torch.zeros(1, 2, 2, 3, 2, 2), 4)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        boxer_b200.ops.value_epilogue(torch.zeros(1, 5, 8), None, torch.bfloat16)
    with pytest.raises(TypeError):
        boxer_b200.ops.instance_weights_forward([[1.0]], 4)


def test_ops_reject_cpu_tensors_like_the_reference():
    """box_attn.h:53 'Not implemented on the CPU' -- and there is no fallback here either."""
    import boxer_b200
    sh = torch.tensor([(4, 4)])
    start = torch.zeros(1, dtype=torch.long)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        boxer_b200.ops.box_attn_forward(torch.zeros(1, 16, 2, 8), sh, start, torch.zeros(1, 3, 2, 1, 4, 2),
                                        torch.zeros(1, 3, 2, 1, 4), 64)


def test_missing_extension_fails_loudly(tmp_path):
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from boxer_b200 import _native\n"
        "_native.LIB_PATH = %r\n"
        "try:\n"
        "    _native.load()\n"
        "except ImportError as e:\n"
        "    print('IMPORTERROR', e)\n" % (ROOT, str(tmp_path / "nope.so"))
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "IMPORTERROR" in out and "no CPU / PyTorch fallback" in out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "boxer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "kernel_ref" not in text and "grid_sample" not in text, f


# ------------------------------------------------------------------ host logic
def test_magic_division_is_exact():
    for P in list(range(2, 200)) + [256, 729, 784, 1024, 4096]:
        magic = (2 ** 32 + P - 1) // P
        assert magic < 2 ** 32
        for j in list(range(0, 3000)) + [65535, 65535 - P, 40000]:
            assert (j * magic) >> 32 == j // P, (P, j)


def test_fpn_levels_match_survey():
    from boxer_b200 import workloads as W
    assert W.fpn_levels(800, 1333) == [(100, 167), (50, 84), (25, 42), (13, 21)]
    assert sum(h * w for h, w in W.fpn_levels(800, 1333)) == 22223
    assert W.bytes_per_sample(False, False, 32, 4, 4) == 4 * (128 + 3 + 2)      # 532 B
    assert W.bytes_per_sample(False, False, 32, 4, 16) == 4 * (128 + 3 + 0.5)   # 526 B
    assert W.bytes_per_sample(False, True, 32, 4, 16) == 4 * (384 + 6 + 0.5)    # 1562 B


def test_workload_generators_cpu():
    from boxer_b200 import workloads as W
    w = W.coco_encoder(K=2, device="cpu", image=(64, 96))
    d = w.dims
    assert d["S"] == d["Nq"] and w.loc.shape == (1, d["S"], 8, 4, 4, 2)
    assert torch.allclose(w.weights[0].flatten(3).sum(-1), torch.ones(1, d["S"], 8))
    # level-0 query boxes are 4 px wide in their own level (box_transformer.py:101-111)
    x = w.loc[0, 0, 0, 0, :, 0] * 12
    assert float(x.max() - x.min()) == pytest.approx(4 * (1 + 0) / 2 * 1.0, rel=0.5)
    m = W.coco_mask_head(Nq=3, K=4, device="cpu", image=(64, 96))
    assert m.instance and torch.allclose(m.weights[1].sum(3), torch.ones(1, 3, 8, 4, 4))
    r = W.bev_rotated(Nq=5, size=32, device="cpu")
    assert r.loc.shape == (1, 5, 8, 1, 9, 2)


def test_trained_like_and_box3d_generators_cpu():
    from boxer_b200 import workloads as W
    t = W.coco_encoder(K=4, dist="trained", device="cpu", image=(64, 96))
    assert t.loc.shape == (1, t.dims["S"], 8, 4, 16, 2)
    # box extents are 2..64 px of level 0 (12 px wide here): per-(row, level) spread of x in level-0 pixels
    x = t.loc[..., 0] * 12
    ext = (x.amax(-1) - x.amin(-1)) * 4 / 3          # K=4 grid spans 3/4 of the box
    assert 1.9 <= float(ext.min()) and float(ext.max()) <= 64.1
    fb, ft, fu = (W.window_mode_fraction(W.coco_encoder(K=4, dist=d, device="cpu", image=(128, 192))) for d in ("box", "trained", "uniform"))
    assert fb > ft > fu
    e = W.box3d_encoder(device="cpu", levels=((12, 12), (6, 6)))
    assert e.dims == dict(B=1, S=180, H=8, D=32, L=2, Nq=180, P=4)
    # head 0: reference angle 0 -> normalised 0.5, used as radians (box_attention.py:318-327): the 2x2 grid is turned by 0.5 rad
    d = e.loc[0, 0, 0, 0]
    assert float(torch.atan2((d[1] - d[0])[1], (d[1] - d[0])[0])) == pytest.approx(0.5, abs=1e-4)


def test_compat_install_maps_reference_import_paths():
    import boxer_b200
    boxer_b200.compat.install()
    try:
        from e2edet.module.ops import BoxAttnFunction, InstanceAttnFunction
        from e2edet.module.box_attention import BoxAttention, InstanceAttention, Box3dAttention
        from e2edet import ops
        assert BoxAttnFunction is boxer_b200.BoxAttnFunction and InstanceAttnFunction is boxer_b200.InstanceAttnFunction
        assert BoxAttention is boxer_b200.BoxAttention and Box3dAttention is boxer_b200.Box3dAttention
        assert InstanceAttention is boxer_b200.InstanceAttention
        for f in ("box_attn_forward", "box_attn_backward", "instance_attn_forward", "instance_attn_backward"):
            assert callable(getattr(ops, f))      # vision.cpp:7-12
    finally:
        boxer_b200.compat.uninstall()


# ------------------------------------------------------------------ module mirror vs the reference modules
def _oracle_box_fn(value, v_shape, v_start, grid, weights, step):
    from oracle import plain
    b, s = value.shape[:2]
    return plain.plain_box_attn(value.reshape(b, s, -1), v_shape, 2 * grid - 1, weights)


def _oracle_inst_fn(value, v_shape, v_start, grid, sw, lw, k, step):
    from oracle import plain
    b, s = value.shape[:2]
    return plain.plain_instance_attn(value.reshape(b, s, -1), v_shape, 2 * grid - 1, sw, lw, k)


@pytest.mark.parametrize("case", list(refinputs.module_cases()))
def test_module_mirror_matches_reference_modules(case, monkeypatch):
    """State-dict keys / shapes, box->grid (+rotation), softmaxes and projections of the three
    nn.Modules against outputs of the reference's own classes (tests/golden/modules_golden.npz).
    The native op is replaced by the oracle here (CPU); the GPU test runs the real thing."""
    import boxer_b200
    from boxer_b200 import box_attention as BA
    monkeypatch.setattr(BA, "_box_attn", _oracle_box_fn)
    monkeypatch.setattr(BA, "_instance_attn", _oracle_inst_fn)
    spec = refinputs.module_cases()[case]
    gold = helpers.golden("modules_golden")[case]
    mod = getattr(boxer_b200, spec["cls"])(**spec["ctor"]).double()
    state = {k[len("param_"):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("param_")}
    assert set(mod.state_dict()) == set(state)
    mod.load_state_dict(state, strict=True)
    if spec["cls"] == "InstanceAttention":
        mod.inferencing = spec["inferencing"]
    outs = mod(*refinputs.module_inputs(spec))
    flat = []
    for o in outs:
        if o is None:
            continue
        flat.extend(o if isinstance(o, tuple) else [o])
    assert len(flat) == int(gold["n_out"])
    for i, o in enumerate(flat):
        assert helpers.max_err(o, gold[f"out{i}"]) <= 1e-12, (case, i)


def test_module_init_matches_reference_init():
    import boxer_b200
    torch.manual_seed(0)
    m = boxer_b200.BoxAttention(256, 4, 8, 2)
    assert float(m.linear_box_weight.abs().sum()) == 0 and float(m.linear_attn_weight.abs().sum()) == 0
    assert float(m.linear_attn_bias.abs().sum()) == 0 and float(m.out_proj.bias.abs().sum()) == 0
    assert 0 <= float(m.linear_box_bias.min()) and float(m.linear_box_bias.max()) < 1
    assert m.linear_box_weight.shape == (4 * 8 * 4, 256) and m.linear_attn_weight.shape == (8 * 4 * 4, 256)
    assert m.im2col_step == 64 and m.kernel_indices.shape == (4, 2)
    assert torch.allclose(m.kernel_indices, torch.tensor([[-.25, -.25], [.25, -.25], [-.25, .25], [.25, .25]]))
    i = boxer_b200.InstanceAttention(256, 4, 8, 14)
    assert i.linear_attn_weight.shape == (8 * 4 * 4, 256) and i.kernel_indices.shape == (196, 2)
    assert not hasattr(i, "inferencing")      # injected by the model (base_model.py:49-67), as in the reference
    r = boxer_b200.Box3dAttention(256, 2, 8, with_rotation=True, kernel_size=2)
    assert r.linear_box_weight.shape == (2 * 8 * 5, 256)
    assert torch.allclose(r.kernel_indices.abs(), torch.full((4, 2), 0.25))      # /2, not /K (box_attention.py:291)
    r3 = boxer_b200.Box3dAttention(256, 2, 8, with_rotation=False, kernel_size=3)
    assert torch.allclose(r3.kernel_indices.abs().max(), torch.tensor(0.5))


def test_slot_table_index_arithmetic():
    """The slot-table walks (boxattn_window.cuh, phase B') split a dense window slot s < 64 into (row, column) with a
    multiply-shift instead of a division: y = (s * rcp) >> 16 with rcp = floor(q) + 1, q ~ 65536 / nx from the fast
    (2 ulp) float division.  Host restatement of that identity, including the approximate quotient landing on
    either side of the exact one, and of the empty-range sentinel arithmetic (kNoPix = 2^30 - 1 stays inside int32)."""
    import numpy as np
    for nx in range(1, 65):
        q = np.float32(65536.0) / np.float32(nx)
        for qq in (np.nextafter(q, np.float32(0)), q, np.nextafter(q, np.float32(1e9)),
                   np.nextafter(np.nextafter(q, np.float32(1e9)), np.float32(1e9)),
                   np.nextafter(np.nextafter(q, np.float32(0)), np.float32(0))):
            rcp = int(np.floor(qq)) + 1
            for s in range(64):
                y = (s * rcp) >> 16
                assert y == s // nx and s - y * nx == s % nx, (nx, s, rcp)
    k = 0x3FFFFFFF
    for hi in (1, 10 ** 6):                      # extent of an empty range: max - min + 1 with the clamps applied
        assert -(2 ** 31) <= min(-k, hi - 1) - max(k, 0) + 1 < 0


def test_torch_shim_loads_and_defers_to_python_for_unusual_inputs():
    """boxer_b200/_C/_boxattn_torch.so (csrc/boxattn_torch.cpp): the pybind fast path over the same C ABI.  Without a GPU
    only its loading and its "not the plain case -> None" contract can be exercised."""
    from boxer_b200 import _native
    if not os.path.exists(_native.SHIM_PATH):
        pytest.skip("shim not built")
    shim = _native.load_shim()
    assert shim is not None and shim.abi_version() == 1
    for f in ("box_attn_forward", "box_attn_backward", "instance_attn_forward", "instance_attn_backward"):
        assert callable(getattr(shim, f))
    v = torch.zeros(1, 4, 2, 32)
    sh = torch.tensor([[2, 2]])
    ls = torch.tensor([0])
    loc = torch.zeros(1, 3, 2, 1, 4, 2)
    w = torch.zeros(1, 3, 2, 1, 4)
    assert shim.box_attn_forward(v, sh, ls, loc, w, 64, 0) is None          # CPU tensors: the Python route raises the error
    import boxer_b200
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        boxer_b200.ops.box_attn_forward(v, sh, ls, loc, w, 64)


def test_bench_compulsory_bytes_model():
    """bench.py's `roofline_compulsory`: inputs read once, outputs written once."""
    sys.path.insert(0, ROOT)
    import bench
    d = dict(B=1, S=22223, H=8, D=32, L=4, Nq=22223, P=16)
    n = 22223 * 8 * 4 * 16
    f = bench.compulsory(n, d, 0.176, 6555.5, False, 189_681_408)
    b = bench.compulsory(n, d, 0.347, 6555.5, True, None)
    assert f["bytes"] == 22223 * 256 * 4 * 2 + n * 12 == 182_050_816        # value + out + (loc, weight) per sample
    assert b["bytes"] == 2 * f["bytes"] and "frac_measured_dram" not in b     # + grad_out, three gradients, the zero fill
    assert f["frac"] == pytest.approx(182_050_816 / 0.176e-3 / 1e9 / 6555.5)
    assert f["frac_measured_dram"] > f["frac"]


def test_backward_reduction_bytes_matches_a_brute_force_count():
    """bench.py's roofline_l2_reduction numerator: one reduction per unique touched pixel where the touched range fits the
    64-slot window, one per in-range corner where it does not (boxattn_window.cuh phases B / C), restated with loops."""
    import math
    import torch
    from boxer_b200 import workloads as W
    for dist in ("box", "uniform"):
        w = W.coco_encoder(K=4, image=(64, 96), device="cpu", dist=dist, oob=0.1)
        loc = w.loc[0, ::23]                   # a few queries are enough for the loops
        sub = W.Workload(**{**w.__dict__, "loc": w.loc[:, ::23], "weights": tuple(t[:, ::23] for t in w.weights)})
        D = w.value.shape[-1]
        want = 0
        for q in range(loc.shape[0]):
            for h in range(loc.shape[1]):
                for l, (hh, ww) in enumerate(w.shapes.tolist()):
                    pix, corners, xs, ys = set(), 0, [], []
                    for p in range(loc.shape[3]):
                        x = float(loc[q, h, l, p, 0]) * ww - 0.5
                        y = float(loc[q, h, l, p, 1]) * hh - 0.5
                        x, y = float(torch.tensor(x, dtype=torch.float32)), float(torch.tensor(y, dtype=torch.float32))
                        if not (x > -1 and y > -1 and x < ww and y < hh):
                            continue
                        x0, y0 = math.floor(x), math.floor(y)
                        xs += [x0, x0 + 1]
                        ys += [y0, y0 + 1]
                        for dy in (0, 1):
                            for dx in (0, 1):
                                if 0 <= x0 + dx < ww and 0 <= y0 + dy < hh:
                                    corners += 1
                                    pix.add((y0 + dy, x0 + dx))
                    if not xs:
                        continue
                    nx = min(max(xs), ww - 1) - max(min(xs), 0) + 1
                    ny = min(max(ys), hh - 1) - max(min(ys), 0) + 1
                    want += len(pix) if nx * ny <= 64 else corners
        got = W.backward_reduction_bytes(sub)
        # (a corner whose bilinear weight is exactly zero is skipped by the kernel and by the function, not by the loops)
        assert abs(got - want * 4 * D) <= 0.01 * want * 4 * D, (dist, got, want * 4 * D)
