// Minimal torch binding of the C ABI (include/boxattn_b200.h) for the four reference entry points.
//
// Why it exists: a decoder-sized call (300 queries x 8 heads) is launch-latency bound, and the ctypes route through
// boxer_b200/ops.py costs ~15 us of Python per call (argument checks, output allocation, 16 ctypes conversions) against
// ~10 us for the reference's pybind function (vision.cpp:7-12).  This shim does the same work in C++ -- fast-path
// validation, at::empty for outputs / workspace, current stream, device guard -- and then calls EXACTLY the C-ABI
// functions the ctypes route calls (it links against libboxattn_b200.so; no kernel lives here).  Anything unusual
// (non-contiguous input, dtype / shape mismatch, CPU tensor ...) makes it return None, and ops.py runs its own
// validating path, so error messages and behaviour have one source.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <string>
#include <vector>

#include "../../include/boxattn_b200.h"

namespace {

struct Geo {
    int B, S, H, D, L, Nq, P;
    int kind;      // 0 f32, 1 f64, 2 bf16
};

bool cuda_contig(const at::Tensor& t, const at::Device& dev) { return t.is_cuda() && t.is_contiguous() && t.device() == dev; }

// the checks of ops._geometry / ops._dtypes; false -> let the Python path diagnose
bool geometry(const at::Tensor& value, const at::Tensor& shapes, const at::Tensor& lsi, const at::Tensor& loc,
              const std::vector<const at::Tensor*>& weights, int64_t im2col_step, Geo& g) {
    if (!value.is_cuda() || value.dim() != 4) return false;
    const at::Device dev = value.device();
    if (!cuda_contig(value, dev) || !cuda_contig(shapes, dev) || !cuda_contig(lsi, dev) || !cuda_contig(loc, dev)) return false;
    if (shapes.scalar_type() != at::kLong || lsi.scalar_type() != at::kLong || shapes.dim() != 2 || shapes.size(1) != 2) return false;
    const int64_t B = value.size(0), S = value.size(1), H = value.size(2), D = value.size(3), L = shapes.size(0);
    if (lsi.numel() != L || L > BXR_MAX_LEVELS) return false;
    if (loc.dim() != 6 || loc.size(0) != B || loc.size(2) != H || loc.size(3) != L || loc.size(5) != 2) return false;
    const int64_t Nq = loc.size(1), P = loc.size(4);
    at::ScalarType tw;
    switch (value.scalar_type()) {
        case at::kFloat: g.kind = 0; tw = at::kFloat; break;
        case at::kDouble: g.kind = 1; tw = at::kDouble; break;
        case at::kBFloat16: g.kind = 2; tw = at::kFloat; break;
        default: return false;
    }
    if (loc.scalar_type() != tw) return false;
    for (const at::Tensor* w : weights) {
        if (!cuda_contig(*w, dev) || w->scalar_type() != tw || w->numel() != B * Nq * H * L * P) return false;
    }
    const int64_t step = std::min<int64_t>(B, im2col_step);
    if (B > 0 && (step <= 0 || B % step != 0)) return false;
    if (B > INT32_MAX || S > INT32_MAX || H > INT32_MAX || D > INT32_MAX || Nq > INT32_MAX || P > INT32_MAX) return false;
    g.B = (int)B; g.S = (int)S; g.H = (int)H; g.D = (int)D; g.L = (int)L; g.Nq = (int)Nq; g.P = (int)P;
    return true;
}

void check(int status, const char* what) {
    if (status != BXR_OK) {
        const char* detail = bxr_last_error_detail();
        std::string msg = std::string(what) + " failed: " + bxr_status_string(status);
        if (detail && detail[0]) msg += std::string(" (") + detail + ")";
        TORCH_CHECK(false, msg);
    }
}

at::Tensor workspace(const at::Tensor& value, const Geo& g, unsigned flags, size_t& bytes) {
    bytes = bxr_attn_bwd_workspace_bytes((int)value.element_size(), g.B, g.S, g.H, g.D, flags);
    if (!bytes) return at::Tensor();
    return at::empty({(int64_t)bytes}, value.options().dtype(at::kByte));
}

#define BXR_BY_KIND(kind, CALL_F32, CALL_F64, CALL_BF16) ((kind) == 0 ? (CALL_F32) : ((kind) == 1 ? (CALL_F64) : (CALL_BF16)))

py::object box_attn_forward(const at::Tensor& value, const at::Tensor& shapes, const at::Tensor& lsi, const at::Tensor& loc,
                            const at::Tensor& attn, int64_t im2col_step, int64_t flags) {
    Geo g;
    if (!geometry(value, shapes, lsi, loc, {&attn}, im2col_step, g)) return py::none();
    c10::cuda::CUDAGuard guard(value.device());
    at::Tensor out = at::empty({g.B, g.Nq, (int64_t)g.H * g.D}, value.options());
    bxr_stream_t st = at::cuda::getCurrentCUDAStream().stream();
    const void *v = value.data_ptr(), *l = loc.data_ptr(), *a = attn.data_ptr();
    const int64_t *sh = shapes.data_ptr<int64_t>(), *ls = lsi.data_ptr<int64_t>();
    void* o = out.data_ptr();
    const unsigned f = (unsigned)flags;
    const int s = BXR_BY_KIND(g.kind,
        bxr_box_attn_fwd_f32((const float*)v, sh, ls, (const float*)l, (const float*)a, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P, (float*)o, f, st),
        bxr_box_attn_fwd_f64((const double*)v, sh, ls, (const double*)l, (const double*)a, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P, (double*)o, f, st),
        bxr_box_attn_fwd_bf16((const bxr_bf16*)v, sh, ls, (const float*)l, (const float*)a, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P, (bxr_bf16*)o, f, st));
    check(s, "box_attn_forward");
    return py::cast(out);
}

py::object box_attn_backward(const at::Tensor& value, const at::Tensor& shapes, const at::Tensor& lsi, const at::Tensor& loc,
                             const at::Tensor& attn, const at::Tensor& grad_out, int64_t im2col_step, int64_t flags) {
    Geo g;
    if (!geometry(value, shapes, lsi, loc, {&attn}, im2col_step, g)) return py::none();
    if (!cuda_contig(grad_out, value.device()) || grad_out.scalar_type() != value.scalar_type() ||
        grad_out.numel() != (int64_t)g.B * g.Nq * g.H * g.D)
        return py::none();
    c10::cuda::CUDAGuard guard(value.device());
    at::Tensor gv = at::empty_like(value), gl = at::empty_like(loc), ga = at::empty_like(attn);
    const unsigned f = (unsigned)flags;
    size_t wsb = 0;
    at::Tensor ws = workspace(value, g, f, wsb);
    void* wsp = wsb ? ws.data_ptr() : nullptr;
    bxr_stream_t st = at::cuda::getCurrentCUDAStream().stream();
    const void *v = value.data_ptr(), *l = loc.data_ptr(), *a = attn.data_ptr(), *go = grad_out.data_ptr();
    const int64_t *sh = shapes.data_ptr<int64_t>(), *ls = lsi.data_ptr<int64_t>();
    const int s = BXR_BY_KIND(g.kind,
        bxr_box_attn_bwd_f32((const float*)v, sh, ls, (const float*)l, (const float*)a, (const float*)go, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P,
                             (float*)gv.data_ptr(), (float*)gl.data_ptr(), (float*)ga.data_ptr(), wsp, wsb, f, st),
        bxr_box_attn_bwd_f64((const double*)v, sh, ls, (const double*)l, (const double*)a, (const double*)go, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P,
                             (double*)gv.data_ptr(), (double*)gl.data_ptr(), (double*)ga.data_ptr(), wsp, wsb, f, st),
        bxr_box_attn_bwd_bf16((const bxr_bf16*)v, sh, ls, (const float*)l, (const float*)a, (const bxr_bf16*)go, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P,
                              (bxr_bf16*)gv.data_ptr(), (float*)gl.data_ptr(), (float*)ga.data_ptr(), wsp, wsb, f, st));
    check(s, "box_attn_backward");
    return py::cast(std::vector<at::Tensor>{gv, gl, ga});
}

py::object instance_attn_forward(const at::Tensor& value, const at::Tensor& shapes, const at::Tensor& lsi, const at::Tensor& loc,
                                 const at::Tensor& sw, const at::Tensor& lw, int64_t im2col_step, int64_t flags) {
    Geo g;
    if (!geometry(value, shapes, lsi, loc, {&sw, &lw}, im2col_step, g)) return py::none();
    c10::cuda::CUDAGuard guard(value.device());
    at::Tensor out = at::empty({g.B, g.Nq, (int64_t)g.H * g.D}, value.options());
    at::Tensor mask = at::empty({g.B, g.Nq, g.P, (int64_t)g.H * g.D}, value.options());
    bxr_stream_t st = at::cuda::getCurrentCUDAStream().stream();
    const void *v = value.data_ptr(), *l = loc.data_ptr(), *a = sw.data_ptr(), *b = lw.data_ptr();
    const int64_t *sh = shapes.data_ptr<int64_t>(), *ls = lsi.data_ptr<int64_t>();
    const unsigned f = (unsigned)flags;
    const int s = BXR_BY_KIND(g.kind,
        bxr_instance_attn_fwd_f32((const float*)v, sh, ls, (const float*)l, (const float*)a, (const float*)b, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P,
                                  (float*)out.data_ptr(), (float*)mask.data_ptr(), f, st),
        bxr_instance_attn_fwd_f64((const double*)v, sh, ls, (const double*)l, (const double*)a, (const double*)b, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P,
                                  (double*)out.data_ptr(), (double*)mask.data_ptr(), f, st),
        bxr_instance_attn_fwd_bf16((const bxr_bf16*)v, sh, ls, (const float*)l, (const float*)a, (const float*)b, g.B, g.S, g.H, g.D, g.L, g.Nq, g.P,
                                   (bxr_bf16*)out.data_ptr(), (bxr_bf16*)mask.data_ptr(), f, st));
    check(s, "instance_attn_forward");
    return py::cast(std::vector<at::Tensor>{out, mask});
}

py::object instance_attn_backward(const at::Tensor& value, const at::Tensor& shapes, const at::Tensor& lsi, const at::Tensor& loc,
                                  const at::Tensor& sw, const at::Tensor& lw, const at::Tensor& grad_out, const at::Tensor& grad_mask,
                                  int64_t im2col_step, int64_t flags) {
    Geo g;
    if (!geometry(value, shapes, lsi, loc, {&sw, &lw}, im2col_step, g)) return py::none();
    const at::Device dev = value.device();
    if (!cuda_contig(grad_out, dev) || grad_out.scalar_type() != value.scalar_type() || grad_out.numel() != (int64_t)g.B * g.Nq * g.H * g.D)
        return py::none();
    if (!cuda_contig(grad_mask, dev) || grad_mask.scalar_type() != value.scalar_type() ||
        grad_mask.numel() != (int64_t)g.B * g.Nq * g.P * g.H * g.D)
        return py::none();
    c10::cuda::CUDAGuard guard(dev);
    at::Tensor gv = at::empty_like(value), gl = at::empty_like(loc), gs = at::empty_like(sw), gw = at::empty_like(lw);
    const unsigned f = (unsigned)flags;
    size_t wsb = 0;
    at::Tensor ws = workspace(value, g, f, wsb);
    void* wsp = wsb ? ws.data_ptr() : nullptr;
    bxr_stream_t st = at::cuda::getCurrentCUDAStream().stream();
    const void *v = value.data_ptr(), *l = loc.data_ptr(), *a = sw.data_ptr(), *b = lw.data_ptr(), *go = grad_out.data_ptr(), *gm = grad_mask.data_ptr();
    const int64_t *sh = shapes.data_ptr<int64_t>(), *ls = lsi.data_ptr<int64_t>();
    const int s = BXR_BY_KIND(g.kind,
        bxr_instance_attn_bwd_f32((const float*)v, sh, ls, (const float*)l, (const float*)a, (const float*)b, (const float*)go, (const float*)gm,
                                  g.B, g.S, g.H, g.D, g.L, g.Nq, g.P, (float*)gv.data_ptr(), (float*)gl.data_ptr(), (float*)gs.data_ptr(),
                                  (float*)gw.data_ptr(), wsp, wsb, f, st),
        bxr_instance_attn_bwd_f64((const double*)v, sh, ls, (const double*)l, (const double*)a, (const double*)b, (const double*)go, (const double*)gm,
                                  g.B, g.S, g.H, g.D, g.L, g.Nq, g.P, (double*)gv.data_ptr(), (double*)gl.data_ptr(), (double*)gs.data_ptr(),
                                  (double*)gw.data_ptr(), wsp, wsb, f, st),
        bxr_instance_attn_bwd_bf16((const bxr_bf16*)v, sh, ls, (const float*)l, (const float*)a, (const float*)b, (const bxr_bf16*)go, (const bxr_bf16*)gm,
                                   g.B, g.S, g.H, g.D, g.L, g.Nq, g.P, (bxr_bf16*)gv.data_ptr(), (float*)gl.data_ptr(), (float*)gs.data_ptr(),
                                   (float*)gw.data_ptr(), wsp, wsb, f, st));
    check(s, "instance_attn_backward");
    return py::cast(std::vector<at::Tensor>{gv, gl, gs, gw});
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "torch binding of libboxattn_b200.so's four reference entry points (fast path; None = use the Python route)";
    m.def("abi_version", []() { return bxr_abi_version(); });
    m.def("last_launch_count", []() { return bxr_last_launch_count(); });
    m.def("box_attn_forward", &box_attn_forward);
    m.def("box_attn_backward", &box_attn_backward);
    m.def("instance_attn_forward", &instance_attn_forward);
    m.def("instance_attn_backward", &instance_attn_backward);
}
