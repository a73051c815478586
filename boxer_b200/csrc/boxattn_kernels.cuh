// Box-attention / instance-attention device code for sm_100a (B200).
//
// The op (reference semantics: e2edet/module/ops/src/box_attn/box_attn_kernel.cuh:34-349,
// instance_attn/instance_attn_kernel.cuh:98-364) is an irregular multi-scale bilinear
// gather-reduce: no dense contraction, so no tensor cores.  What matters on B200 is
//   * 16-byte vector loads of the D-channel corner rows (a row of D=32 fp32 channels is one
//     128 B line = 8 lanes x float4; bf16: 4 lanes x 8 channels),
//   * many independent corner loads in flight per warp (a warp serves 32/G rows at once and
//     batches U sample points, i.e. 4*U*32/G 128-byte requests outstanding),
//   * consecutive (query, head) rows handled by the same CTA over time so that the
//     overlapping boxes of neighbouring queries hit in L1/L2 (value is ~23 MB at COCO scale
//     and lives in the 126 MB L2),
//   * warp-shuffle reductions over the G lanes of a row for the location / weight gradients
//     (the reference does a thread-0 serial sum behind two __syncthreads per point),
//   * vector reductions (red.global.add.v4.f32) for the grad_value scatter, or an
//     order-independent 64-bit fixed-point scatter when determinism is requested.
//
// Two families of kernels:
//   *_vec : G = D / VEC lanes per (b, q, head) row, VEC = 16 B / sizeof(TV); G in {1..32}, pow2.
//   *_gen : any D, any of float / double / bf16: one warp per row, lanes stride over channels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace bxr {

constexpr int kMaxLevels = 32;
constexpr int kThreads = 256;

// ----------------------------------------------------------------------------- parameters
struct AttnParams {
    // inputs
    const void* value;            // (B,S,H,D) TV
    const int64_t* shapes;        // (L,2) device
    const int64_t* level_start;   // (L,)  device
    const void* loc;              // (B,Nq,H,L,P,2) TW
    const void* w0;               // attn (box) / spatial_w (instance)   TW
    const void* w1;               // level_w (instance)                   TW
    const void* grad_out;         // (B,Nq,H*D) TV            (bwd)
    const void* grad_mask;        // (B,Nq,P,H*D) TV          (instance bwd)
    // outputs
    void* out;                    // (B,Nq,H*D) TV            (fwd)
    void* mask_out;               // (B,Nq,P,H*D) TV          (instance fwd)
    void* grad_value_acc;         // (B,S,H,D) accumulator: TC (float/double) or int64 fixed point (DET)
    void* grad_loc;               // TW
    void* grad_w0;                // TW
    void* grad_w1;                // TW
    const float* det_scale;       // device scalar: power-of-two scale of the fixed-point scatter (DET)
    // fused box -> grid entry points (boxattn_fused.cuh): locations are generated in-kernel
    const void* boxes;            // (B,Nq,H,L,4) TW  cx, cy, w, h (normalised)
    const void* angles;           // (B,Nq,H,L)   TW  radians, or nullptr
    const void* valid_ratios;     // (B,L,2)      TW  (x, y), or nullptr
    const void* kidx;             // (P,2)        TW  kernel_indices
    void* grad_boxes;             // (B,Nq,H,L,4) TW
    void* grad_angles;            // (B,Nq,H,L)   TW  or nullptr
    void* attn_out;               // (B,Nq,H,L,P) TW  softmax weights written by the fused-softmax forward
    // sizes
    int B, S, H, D, L, Nq, P;
    int LP;                       // L*P
    unsigned magicP;              // ceil(2^32 / P): j / P == __umulhi(j, magicP) for j < 65536
    long long rows;               // B*Nq*H
    int nsplit_log2;              // a row's points are split over 2^k lane groups of one CTA
    int chunk;                    // points per split (box: of L*P, instance: of P)
    int units;                    // ceil(rows / rows_per_unit)
};

// ----------------------------------------------------------------------------- small helpers
template <typename T> struct Compute { using type = float; };
template <> struct Compute<double> { using type = double; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ double to_f(double v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename TV, typename TC> __device__ __forceinline__ TV from_f(TC v);
template <> __device__ __forceinline__ float from_f<float, float>(float v) { return v; }
template <> __device__ __forceinline__ double from_f<double, double>(double v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16, float>(float v) { return __float2bfloat16_rn(v); }

// 16-byte vector <-> VEC floats
template <typename TV> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int VEC = 4;
    __device__ __forceinline__ static void load(const float* p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ static void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    // 16-byte-unit index off a uniform base: one IMAD.WIDE.U32 of address math per load
    __device__ __forceinline__ static void load16(const void* base, unsigned idx, float (&v)[4]) {
        const uint4 t = __ldg(static_cast<const uint4*>(base) + idx);
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    }
    // the same load kept in its storage form, so that many can be in flight before the first is unpacked
    using Raw = uint4;
    __device__ __forceinline__ static Raw zero_raw() { return make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ static Raw load_raw(const void* base, unsigned idx) { return __ldg(static_cast<const uint4*>(base) + idx); }
    __device__ __forceinline__ static void unpack_raw(const Raw& t, float (&v)[4]) {
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
        const unsigned u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // bf16 -> fp32 is a 16-bit shift
            v[2 * i] = __uint_as_float(u[i] << 16);
            v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ static void load16(const void* base, unsigned idx, float (&v)[8]) {
        const uint4 t = __ldg(static_cast<const uint4*>(base) + idx);
        const unsigned u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(u[i] << 16);
            v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
        }
    }
    using Raw = uint4;
    __device__ __forceinline__ static Raw zero_raw() { return make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ static Raw load_raw(const void* base, unsigned idx) { return __ldg(static_cast<const uint4*>(base) + idx); }
    __device__ __forceinline__ static void unpack_raw(const Raw& t, float (&v)[8]) {
        const unsigned u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(u[i] << 16);
            v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        unsigned u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            u[i] = *reinterpret_cast<const unsigned*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
    }
};

// bf16 storage with 4 channels (8 bytes) per lane: for head_dim 32 this keeps G = 8 lanes per row --
// the fp32 kernels' geometry (taps spread over 8 lanes, 4 rows per warp) at half the bytes per load.
struct bf16x4_t { __nv_bfloat16 v; };
template <> struct Vec16<bf16x4_t> {
    static constexpr int VEC = 4;
    __device__ __forceinline__ static void unpack(const uint2 t, float (&v)[4]) {
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    __device__ __forceinline__ static void load(const bf16x4_t* p, float (&v)[4]) {
        unpack(__ldg(reinterpret_cast<const uint2*>(p)), v);
    }
    __device__ __forceinline__ static void load16(const void* base, unsigned idx, float (&v)[4]) {   // idx in 8-byte units
        unpack(__ldg(static_cast<const uint2*>(base) + idx), v);
    }
    using Raw = uint2;
    __device__ __forceinline__ static Raw zero_raw() { return make_uint2(0u, 0u); }
    __device__ __forceinline__ static Raw load_raw(const void* base, unsigned idx) { return __ldg(static_cast<const uint2*>(base) + idx); }
    __device__ __forceinline__ static void unpack_raw(const Raw& t, float (&v)[4]) { unpack(t, v); }
    __device__ __forceinline__ static void store(bf16x4_t* p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const unsigned*>(&a), *reinterpret_cast<const unsigned*>(&b));
    }
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ void red_add(float* p, float a) { atomicAdd(p, a); }
__device__ __forceinline__ void red_add(double* p, double a) { atomicAdd(p, a); }
__device__ __forceinline__ void red_add_fixed(long long* p, float v, float scale) {
    // scale is a power of two: v*scale is exact, the integer sum is order independent
    atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(__float2ll_rn(v * scale)));
}
__device__ __forceinline__ void red_add_fixed(long long* p, double v, float scale) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(__double2ll_rn(v * (double)scale)));
}

// level table in shared memory (read from the device int64 tensors, like the reference does)
struct LevelTable {
    int h[kMaxLevels];
    int w[kMaxLevels];
    long long start[kMaxLevels];
};

__device__ __forceinline__ void load_levels(LevelTable& t, const AttnParams& p) {
    if (threadIdx.x < p.L) {
        t.h[threadIdx.x] = static_cast<int>(p.shapes[2 * threadIdx.x]);
        t.w[threadIdx.x] = static_cast<int>(p.shapes[2 * threadIdx.x + 1]);
        t.start[threadIdx.x] = p.level_start[threadIdx.x];
    }
    __syncthreads();
}

// One bilinear tap.  Corner order 0=(y0,x0) 1=(y0,x1) 2=(y1,x0) 3=(y1,x1).
// Follows box_attn_kernel.cuh:325-328 (pixel coords + window test) and :47-93 (corners).
template <typename TC>
struct Tap {
    bool ok[4];       // corner exists AND the sample passed the window test
    TC cw[4];         // bilinear corner weights (un-masked)
    TC lx, ly, hx, hy;
    long long pix;    // y0*w + x0 (may be negative when a corner is outside)
    int w;            // level width (row pitch in pixels)
};

template <typename TC>
__device__ __forceinline__ Tap<TC> make_tap(TC loc_x, TC loc_y, int h, int w, bool active) {
    Tap<TC> t;
    const TC x = loc_x * (TC)w - (TC)0.5;
    const TC y = loc_y * (TC)h - (TC)0.5;
    const bool inside = active && (y > (TC)-1) && (x > (TC)-1) && (y < (TC)h) && (x < (TC)w);
    // outside samples may carry NaN/inf: keep the integer conversion well defined
    const TC xs = inside ? x : (TC)0, ys = inside ? y : (TC)0;
    const TC xf = floor(xs), yf = floor(ys);
    const int x0 = (int)xf, y0 = (int)yf;
    t.lx = xs - xf; t.ly = ys - yf;
    t.hx = (TC)1 - t.lx; t.hy = (TC)1 - t.ly;
    const bool okx0 = x0 >= 0, okx1 = x0 + 1 <= w - 1, oky0 = y0 >= 0, oky1 = y0 + 1 <= h - 1;
    t.ok[0] = inside && oky0 && okx0;
    t.ok[1] = inside && oky0 && okx1;
    t.ok[2] = inside && oky1 && okx0;
    t.ok[3] = inside && oky1 && okx1;
    t.cw[0] = t.hy * t.hx; t.cw[1] = t.hy * t.lx; t.cw[2] = t.ly * t.hx; t.cw[3] = t.ly * t.lx;
    t.pix = (long long)y0 * w + x0;
    t.w = w;
    return t;
}

template <typename TC>
__device__ __forceinline__ long long corner_pix(const Tap<TC>& t, int k) {
    return t.pix + ((k & 1) ? 1 : 0) + ((k & 2) ? t.w : 0);
}

template <int G, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// contiguous slice of work units for this CTA (neighbouring rows stay on one SM -> L1 reuse)
__device__ __forceinline__ void unit_range(int units, int& u0, int& u1) {
    const long long n = units;
    u0 = (int)((n * blockIdx.x) / gridDim.x);
    u1 = (int)((n * (blockIdx.x + 1)) / gridDim.x);
}

// =============================================================================== forward, vector path
// One group of G lanes per (row, split); lane i owns channels [i*VEC, (i+1)*VEC).
template <typename TV, int G, bool INSTANCE, int U>
__global__ void __launch_bounds__(kThreads) attn_fwd_vec_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int GROUPS = kThreads / G;
    __shared__ LevelTable lv;
    __shared__ float s_red[kThreads * VEC];   // cross-split reduction of `out`
    load_levels(lv, p);

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const int nsplit = 1 << p.nsplit_log2;
    const int rows_per_unit = GROUPS >> p.nsplit_log2;
    const int r_local = gid >> p.nsplit_log2;
    const int split = gid & (nsplit - 1);
    const int HD = p.H * p.D;
    const TV* __restrict__ value = static_cast<const TV*>(p.value);
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    const float* __restrict__ w1 = static_cast<const float*>(p.w1);

    int u0, u1;
    unit_range(p.units, u0, u1);
    for (int u = u0; u < u1; ++u) {
        const long long row_raw = (long long)u * rows_per_unit + r_local;
        const bool row_ok = row_raw < p.rows;
        const long long row = row_ok ? row_raw : 0;
        const int head = (int)(row % p.H);
        const long long bq = row / p.H;
        const long long b = bq / p.Nq;
        const TV* vrow = value + (b * p.S * HD + head * p.D + lane * VEC);
        const float* loc_row = loc + row * p.LP * 2;
        const float* w0_row = w0 + row * p.LP;

        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

        if constexpr (!INSTANCE) {
            const int j0 = split * p.chunk;
            const int j1 = min(j0 + p.chunk, p.LP);
            for (int t = 0; t < p.chunk; t += U) {
                Tap<float> tap[U];
                float aw[U];
                long long off[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    const int j = j0 + t + k;
                    const bool act = row_ok && (j < j1);
                    const int jc = act ? j : 0;
                    const int l = p.magicP ? (int)__umulhi((unsigned)jc, p.magicP) : jc;   // magicP == 0 <=> P == 1
                    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc_row) + jc);
                    aw[k] = __ldg(w0_row + jc);
                    tap[k] = make_tap<float>(xy.x, xy.y, lv.h[l], lv.w[l], act);
                    off[k] = lv.start[l];
                }
                float v[U][4][VEC];
#pragma unroll
                for (int k = 0; k < U; ++k)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (tap[k].ok[c]) {
                            V::load(vrow + (off[k] + corner_pix(tap[k], c)) * HD, v[k][c]);
                        } else {
#pragma unroll
                            for (int i = 0; i < VEC; ++i) v[k][c][i] = 0.f;
                        }
                    }
#pragma unroll
                for (int k = 0; k < U; ++k)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float val = tap[k].cw[0] * v[k][0][i] + tap[k].cw[1] * v[k][1][i] +
                                          tap[k].cw[2] * v[k][2][i] + tap[k].cw[3] * v[k][3][i];
                        acc[i] += val * aw[k];
                    }
            }
        } else {
            const float* w1_row = w1 + row * p.LP;
            TV* mrow = static_cast<TV*>(p.mask_out) + (bq * p.P * HD + head * p.D + lane * VEC);
            const int p0 = split * p.chunk;
            const int p1 = min(p0 + p.chunk, p.P);
            for (int t = 0; t < p.chunk; ++t) {
                const int pt = p0 + t;
                const bool pact = row_ok && (pt < p1);
                float macc[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) macc[i] = 0.f;
                for (int l0 = 0; l0 < p.L; l0 += U) {
                    Tap<float> tap[U];
                    float sw[U], lw[U];
                    long long off[U];
#pragma unroll
                    for (int k = 0; k < U; ++k) {
                        const int l = l0 + k;
                        const bool act = pact && (l < p.L);
                        const int lc = act ? l : 0;
                        const int j = act ? lc * p.P + pt : 0;
                        const float2 xy = __ldg(reinterpret_cast<const float2*>(loc_row) + j);
                        sw[k] = __ldg(w0_row + j);
                        lw[k] = __ldg(w1_row + j);
                        tap[k] = make_tap<float>(xy.x, xy.y, lv.h[lc], lv.w[lc], act);
                        off[k] = lv.start[lc];
                    }
                    float v[U][4][VEC];
#pragma unroll
                    for (int k = 0; k < U; ++k)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (tap[k].ok[c]) {
                                V::load(vrow + (off[k] + corner_pix(tap[k], c)) * HD, v[k][c]);
                            } else {
#pragma unroll
                                for (int i = 0; i < VEC; ++i) v[k][c][i] = 0.f;
                            }
                        }
#pragma unroll
                    for (int k = 0; k < U; ++k)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const float val = tap[k].cw[0] * v[k][0][i] + tap[k].cw[1] * v[k][1][i] +
                                              tap[k].cw[2] * v[k][2][i] + tap[k].cw[3] * v[k][3][i];
                            acc[i] += val * sw[k];
                            macc[i] += val * lw[k];
                        }
                }
                if (pact) V::store(mrow + (long long)pt * HD, macc);
            }
        }

        TV* orow = static_cast<TV*>(p.out) + (row * p.D + lane * VEC);
        if (nsplit == 1) {
            if (row_ok) V::store(orow, acc);
        } else {
            // deterministic cross-split sum through shared memory
            __syncthreads();
#pragma unroll
            for (int i = 0; i < VEC; ++i) s_red[threadIdx.x * VEC + i] = acc[i];
            __syncthreads();
            if (split == 0 && row_ok) {
                for (int s = 1; s < nsplit; ++s)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[i] += s_red[(threadIdx.x + s * G) * VEC + i];
                V::store(orow, acc);
            }
        }
    }
}

// =============================================================================== backward, vector path
// ACC: float (red.v4.f32 straight into an fp32 grad_value accumulator) or long long (fixed point).
template <typename TV, int G, bool INSTANCE, typename ACC>
__global__ void __launch_bounds__(kThreads) attn_bwd_vec_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int GROUPS = kThreads / G;
    constexpr bool DET = sizeof(ACC) == 8;
    __shared__ LevelTable lv;
    load_levels(lv, p);

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const int nsplit = 1 << p.nsplit_log2;
    const int rows_per_unit = GROUPS >> p.nsplit_log2;
    const int r_local = gid >> p.nsplit_log2;
    const int split = gid & (nsplit - 1);
    const int HD = p.H * p.D;
    const TV* __restrict__ value = static_cast<const TV*>(p.value);
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    const float* __restrict__ w1 = static_cast<const float*>(p.w1);
    ACC* __restrict__ gacc = static_cast<ACC*>(p.grad_value_acc);
    float* __restrict__ grad_loc = static_cast<float*>(p.grad_loc);
    float* __restrict__ grad_w0 = static_cast<float*>(p.grad_w0);
    float* __restrict__ grad_w1 = static_cast<float*>(p.grad_w1);
    float dscale = 1.f;
    if constexpr (DET) dscale = *p.det_scale;

    int u0, u1;
    unit_range(p.units, u0, u1);
    for (int u = u0; u < u1; ++u) {
        const long long row_raw = (long long)u * rows_per_unit + r_local;
        const bool row_ok = row_raw < p.rows;
        const long long row = row_ok ? row_raw : 0;
        const int head = (int)(row % p.H);
        const long long bq = row / p.H;
        const long long b = bq / p.Nq;
        const long long vbase = b * p.S * HD + head * p.D + lane * VEC;
        const float* loc_row = loc + row * p.LP * 2;
        const float* w0_row = w0 + row * p.LP;
        const float* w1_row = INSTANCE ? w1 + row * p.LP : nullptr;

        float go[VEC];
        V::load(static_cast<const TV*>(p.grad_out) + (row * p.D + lane * VEC), go);

        const int lim = INSTANCE ? p.P : p.LP;
        const int i0 = split * p.chunk;
        const int i1 = min(i0 + p.chunk, lim);
        const int inner = INSTANCE ? p.L : 1;
        for (int t = 0; t < p.chunk; ++t) {
            const int it = i0 + t;
            const bool oact = row_ok && (it < i1);
            float gm[VEC];
            if constexpr (INSTANCE) {
                const int pc = oact ? it : 0;
                V::load(static_cast<const TV*>(p.grad_mask) + (bq * p.P * HD + (long long)pc * HD + head * p.D + lane * VEC), gm);
            }
            for (int l_in = 0; l_in < inner; ++l_in) {
                int j, l;
                if constexpr (INSTANCE) {
                    l = l_in;
                    j = oact ? l * p.P + it : 0;
                } else {
                    j = oact ? it : 0;
                    l = p.magicP ? (int)__umulhi((unsigned)j, p.magicP) : j;
                }
                const float2 xy = __ldg(reinterpret_cast<const float2*>(loc_row) + j);
                const float a0 = __ldg(w0_row + j);
                float a1 = 0.f;
                if constexpr (INSTANCE) a1 = __ldg(w1_row + j);
                const int lh = lv.h[l], lw = lv.w[l];
                const Tap<float> tap = make_tap<float>(xy.x, xy.y, lh, lw, oact);
                const long long lbase = vbase + lv.start[l] * HD;

                float v[4][VEC];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (tap.ok[c]) {
                        V::load(value + lbase + corner_pix(tap, c) * HD, v[c]);
                    } else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) v[c][i] = 0.f;
                    }
                }
                float s_a0 = 0.f, s_a1 = 0.f, s_x = 0.f, s_y = 0.f;
                float tg[VEC];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    // top_grad_value: box_attn_kernel.cuh:136, instance_attn_kernel.cuh:139
                    tg[i] = INSTANCE ? (go[i] * a0 + gm[i] * a1) : (go[i] * a0);
                    const float val = tap.cw[0] * v[0][i] + tap.cw[1] * v[1][i] + tap.cw[2] * v[2][i] + tap.cw[3] * v[3][i];
                    s_a0 += go[i] * val;
                    if constexpr (INSTANCE) s_a1 += gm[i] * val;
                    const float dx = tap.hy * (v[1][i] - v[0][i]) + tap.ly * (v[3][i] - v[2][i]);
                    const float dy = tap.hx * (v[2][i] - v[0][i]) + tap.lx * (v[3][i] - v[1][i]);
                    s_x += dx * tg[i];
                    s_y += dy * tg[i];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (tap.ok[c]) {
                        ACC* dst = gacc + lbase + corner_pix(tap, c) * HD;
                        if constexpr (DET) {
#pragma unroll
                            for (int i = 0; i < VEC; ++i) red_add_fixed(reinterpret_cast<long long*>(dst) + i, tap.cw[c] * tg[i], dscale);
                        } else {
#pragma unroll
                            for (int i = 0; i < VEC; i += 4)
                                red_add_v4(reinterpret_cast<float*>(dst) + i, tap.cw[c] * tg[i], tap.cw[c] * tg[i + 1],
                                           tap.cw[c] * tg[i + 2], tap.cw[c] * tg[i + 3]);
                        }
                    }
                }
                s_a0 = group_sum<G>(s_a0);
                s_x = group_sum<G>(s_x);
                s_y = group_sum<G>(s_y);
                if constexpr (INSTANCE) s_a1 = group_sum<G>(s_a1);
                if (lane == 0 && oact) {
                    const long long s = row * p.LP + j;
                    // box_attn_kernel.cuh:181-183
                    grad_w0[s] = s_a0;
                    if constexpr (INSTANCE) grad_w1[s] = s_a1;
                    reinterpret_cast<float2*>(grad_loc)[s] = make_float2((float)lw * s_x, (float)lh * s_y);
                }
            }
        }
    }
}

// =============================================================================== generic path
// One warp per row; lanes stride over the D channels.  Any D, TV in {float,double,bf16}.
// Slow-is-fine: serves fp64 gradcheck and the odd head dims of the reference tests
// (D = 30, 71, 1025, 2048, 3096: box_attn_test.py:194).
template <typename TV, bool INSTANCE>
__global__ void __launch_bounds__(kThreads) attn_fwd_gen_kernel(const AttnParams p) {
    using TC = typename Compute<TV>::type;
    __shared__ LevelTable lv;
    load_levels(lv, p);
    const int lane = threadIdx.x & 31;
    const int warps = kThreads / 32;
    const long long HD = (long long)p.H * p.D;
    const TV* __restrict__ value = static_cast<const TV*>(p.value);
    const TC* __restrict__ loc = static_cast<const TC*>(p.loc);
    const TC* __restrict__ w0 = static_cast<const TC*>(p.w0);
    const TC* __restrict__ w1 = static_cast<const TC*>(p.w1);
    for (long long row = (long long)blockIdx.x * warps + (threadIdx.x >> 5); row < p.rows; row += (long long)gridDim.x * warps) {
        const int head = (int)(row % p.H);
        const long long bq = row / p.H;
        const long long b = bq / p.Nq;
        const TV* vrow = value + (b * p.S * HD + (long long)head * p.D);
        for (int c = lane; c < p.D; c += 32) {
            TC acc = 0;
            if constexpr (!INSTANCE) {
                for (int l = 0; l < p.L; ++l)
                    for (int pt = 0; pt < p.P; ++pt) {
                        const long long s = row * p.LP + (long long)l * p.P + pt;
                        const Tap<TC> tap = make_tap<TC>(loc[2 * s], loc[2 * s + 1], lv.h[l], lv.w[l], true);
                        TC val = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (tap.ok[k]) val += tap.cw[k] * (TC)to_f(vrow[(lv.start[l] + corner_pix(tap, k)) * HD + c]);
                        acc += val * w0[s];
                    }
            } else {
                TV* mrow = static_cast<TV*>(p.mask_out) + (bq * p.P * HD + (long long)head * p.D + c);
                for (int pt = 0; pt < p.P; ++pt) {
                    TC macc = 0;
                    for (int l = 0; l < p.L; ++l) {
                        const long long s = row * p.LP + (long long)l * p.P + pt;
                        const Tap<TC> tap = make_tap<TC>(loc[2 * s], loc[2 * s + 1], lv.h[l], lv.w[l], true);
                        TC val = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (tap.ok[k]) val += tap.cw[k] * (TC)to_f(vrow[(lv.start[l] + corner_pix(tap, k)) * HD + c]);
                        acc += val * w0[s];
                        macc += val * w1[s];
                    }
                    mrow[(long long)pt * HD] = from_f<TV, TC>(macc);
                }
            }
            static_cast<TV*>(p.out)[row * p.D + c] = from_f<TV, TC>(acc);
        }
    }
}

template <typename TV, bool INSTANCE, typename ACC>
__global__ void __launch_bounds__(kThreads) attn_bwd_gen_kernel(const AttnParams p) {
    using TC = typename Compute<TV>::type;
    constexpr bool DET = std::is_same<ACC, long long>::value;
    __shared__ LevelTable lv;
    load_levels(lv, p);
    const int lane = threadIdx.x & 31;
    const int warps = kThreads / 32;
    const long long HD = (long long)p.H * p.D;
    const TV* __restrict__ value = static_cast<const TV*>(p.value);
    const TC* __restrict__ loc = static_cast<const TC*>(p.loc);
    const TC* __restrict__ w0 = static_cast<const TC*>(p.w0);
    const TC* __restrict__ w1 = static_cast<const TC*>(p.w1);
    const TV* __restrict__ grad_out = static_cast<const TV*>(p.grad_out);
    const TV* __restrict__ grad_mask = static_cast<const TV*>(p.grad_mask);
    ACC* __restrict__ gacc = static_cast<ACC*>(p.grad_value_acc);
    TC* __restrict__ grad_loc = static_cast<TC*>(p.grad_loc);
    TC* __restrict__ grad_w0 = static_cast<TC*>(p.grad_w0);
    TC* __restrict__ grad_w1 = static_cast<TC*>(p.grad_w1);
    float dscale = 1.f;
    if constexpr (DET) dscale = *p.det_scale;
    // the whole warp walks the same rows/points, so the shuffles below are convergent
    for (long long row = (long long)blockIdx.x * warps + (threadIdx.x >> 5); row < p.rows; row += (long long)gridDim.x * warps) {
        const int head = (int)(row % p.H);
        const long long bq = row / p.H;
        const long long b = bq / p.Nq;
        const long long vbase = b * p.S * HD + (long long)head * p.D;
        const TV* go = grad_out + row * p.D;
        for (int l = 0; l < p.L; ++l)
            for (int pt = 0; pt < p.P; ++pt) {
                const long long s = row * p.LP + (long long)l * p.P + pt;
                const Tap<TC> tap = make_tap<TC>(loc[2 * s], loc[2 * s + 1], lv.h[l], lv.w[l], true);
                const TC a0 = w0[s];
                TC a1 = 0;
                const TV* gm = nullptr;
                if constexpr (INSTANCE) {
                    a1 = w1[s];
                    gm = grad_mask + (bq * p.P * HD + (long long)pt * HD + (long long)head * p.D);
                }
                const long long lbase = vbase + lv.start[l] * HD;
                TC s_a0 = 0, s_a1 = 0, s_x = 0, s_y = 0;
                for (int c = lane; c < p.D; c += 32) {
                    const TC g = (TC)to_f(go[c]);
                    TC g1 = 0;
                    if constexpr (INSTANCE) g1 = (TC)to_f(gm[c]);
                    const TC tg = INSTANCE ? (g * a0 + g1 * a1) : (g * a0);
                    TC v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[k] = 0;
                        if (tap.ok[k]) {
                            const long long idx = lbase + corner_pix(tap, k) * HD + c;
                            v[k] = (TC)to_f(value[idx]);
                            if constexpr (DET) red_add_fixed(reinterpret_cast<long long*>(gacc) + idx, tap.cw[k] * tg, dscale);
                            else red_add(reinterpret_cast<TC*>(gacc) + idx, tap.cw[k] * tg);
                        }
                    }
                    const TC val = tap.cw[0] * v[0] + tap.cw[1] * v[1] + tap.cw[2] * v[2] + tap.cw[3] * v[3];
                    s_a0 += g * val;
                    if constexpr (INSTANCE) s_a1 += g1 * val;
                    s_x += (tap.hy * (v[1] - v[0]) + tap.ly * (v[3] - v[2])) * tg;
                    s_y += (tap.hx * (v[2] - v[0]) + tap.lx * (v[3] - v[1])) * tg;
                }
                s_a0 = group_sum<32>(s_a0);
                s_x = group_sum<32>(s_x);
                s_y = group_sum<32>(s_y);
                if constexpr (INSTANCE) s_a1 = group_sum<32>(s_a1);
                if (lane == 0) {
                    grad_w0[s] = s_a0;
                    if constexpr (INSTANCE) grad_w1[s] = s_a1;
                    grad_loc[2 * s] = (TC)lv.w[l] * s_x;
                    grad_loc[2 * s + 1] = (TC)lv.h[l] * s_y;
                }
            }
    }
}

// =============================================================================== helpers for the scatter
// max |x| over a TV array -> atomicMax on the float bit pattern (non-negative floats order as ints)
template <typename T>
__global__ void absmax_kernel(const T* __restrict__ x, long long n, unsigned* __restrict__ out_bits) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = fabsf((float)to_f(x[i]));
        if (a == a && a <= 3.0e38f) m = fmaxf(m, a);   // ignore NaN/inf: they poison the result either way
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}

// scale = 2^(40 - ceil(log2(bound))), bound = gmax0*wmax0 (+ gmax1*wmax1): every contribution
// |cw*tg| <= bound maps below 2^40, leaving 2^23 worst-case contributions of headroom in int64.
// (a template only so that the definition may appear in several translation units)
template <int = 0>
__global__ void det_scale_kernel(const unsigned* __restrict__ bits, float* __restrict__ scale) {
    const float g0 = __uint_as_float(bits[0]), a0 = __uint_as_float(bits[1]);
    const float g1 = __uint_as_float(bits[2]), a1 = __uint_as_float(bits[3]);
    const float bound = g0 * a0 + g1 * a1;
    int e = 0;
    if (bound > 0.f) { frexpf(bound, &e); }   // bound = m * 2^e, m in [0.5,1)  =>  bound <= 2^e
    int k = 40 - e;
    k = max(-100, min(100, k));
    *scale = ldexpf(1.f, k);
}

template <typename TV, typename ACC>
__global__ void finalize_grad_value_kernel(const ACC* __restrict__ acc, TV* __restrict__ out, long long n, const float* __restrict__ det_scale) {
    using TC = typename Compute<TV>::type;
    double inv = 1.0;
    if constexpr (std::is_same<ACC, long long>::value) inv = 1.0 / (double)(*det_scale);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if constexpr (std::is_same<ACC, long long>::value) out[i] = from_f<TV, TC>((TC)((double)acc[i] * inv));
        else out[i] = from_f<TV, TC>((TC)acc[i]);
    }
}

}  // namespace bxr
