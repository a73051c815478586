// Instance attention (the mask head: a K x K RoI grid per query, two weights per point and a
// per-point "mask" output -- reference: instance_attn_kernel.cuh:98-187, 282-364) with
// owner-computed taps.
//
// The point kernels of boxattn_kernels.cuh derive every tap on all G lanes of a row.  Here the
// G lanes of a group take G consecutive points of the row: each lane computes the taps of ITS point
// for all levels once (coalesced location / weight loads), then the group walks the G points and
// every lane receives the current point's tap by shuffle, fetches its 16-byte slice of the four
// corner rows and accumulates
//     out      += spatial_w * val                (reduced over the row's point splits through smem)
//     mask[p]  += level_w  * val   over levels   (one 16-byte store per lane per point)
// Backward: the same walk with the scatter to grad_value (red.v4 / fixed point), and the four
// per-point reductions (d spatial_w, d level_w, d x, d y) handed back to the owner lane, which
// writes its point's gradients coalesced after the walk.
#pragma once

#include "boxattn_kernels.cuh"
#include "boxattn_window.cuh"

namespace bxr {

// tap of one (point, level) as broadcast to the whole group
struct BTap {
    bool inside;
    int x0, y0;
    float lx, ly, sw, lw;
};

template <int G>
__device__ __forceinline__ BTap bcast_tap(const LanePoint& t, float lw, int src, unsigned gm) {
    BTap b;
    b.inside = __shfl_sync(gm, (int)t.inside, src, G) != 0;
    b.x0 = __shfl_sync(gm, t.x0, src, G);
    b.y0 = __shfl_sync(gm, t.y0, src, G);
    b.lx = __shfl_sync(gm, t.lx, src, G);
    b.ly = __shfl_sync(gm, t.ly, src, G);
    b.sw = __shfl_sync(gm, t.aw, src, G);
    b.lw = __shfl_sync(gm, lw, src, G);
    return b;
}

#ifndef BXR_INST_LG
#define BXR_INST_LG 2
#endif
// resident CTAs per SM (A/B r01q): the fp32 forward gains 8-11 % from a third CTA (85 registers), the bf16
// forward (8 channels per lane) and both backwards lose to the spills
#ifndef BXR_INST_FWD_MINB_F32
#define BXR_INST_FWD_MINB_F32 3
#endif
// resident CTAs asked of the bf16 (8 channels per lane) forward and of the backward kernels
#ifndef BXR_INST_FWD_MINB_BF16
#define BXR_INST_FWD_MINB_BF16 2
#endif
#ifndef BXR_INST_BWD_MINB
#define BXR_INST_BWD_MINB 2
#endif

// LB: levels held in registers per lane (L <= LB); LG: levels whose corner rows are in flight together
template <typename TV, int G, int LB>
__global__ void __launch_bounds__(kThreads, (Vec16<TV>::VEC > 4) ? BXR_INST_FWD_MINB_BF16 : BXR_INST_FWD_MINB_F32) inst_fwd_own_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int LG = BXR_INST_LG;
    static_assert(LB % LG == 0, "level batches must tile LB");
    constexpr int GROUPS = kThreads / G;
    __shared__ LevelTable lv;
    __shared__ float s_red[kThreads * VEC];
    load_levels(lv, p);

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const unsigned gm = group_mask<G>();
    const int nsplit = 1 << p.nsplit_log2;
    const int rows_per_unit = GROUPS >> p.nsplit_log2;
    const int r_local = gid >> p.nsplit_log2;
    const int split = gid & (nsplit - 1);
    const unsigned HDV = (unsigned)(p.H * p.D) / VEC;
    const long long HD = (long long)p.H * p.D;
    const void* __restrict__ value16 = p.value;   // indexed in lane-chunk units by V::load16
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    const float* __restrict__ w1 = static_cast<const float*>(p.w1);
    const int nchunks = (p.P + G - 1) / G;

    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const long long row = (long long)u * rows_per_unit + r_local;
        const bool row_ok = row < p.rows;
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

        if (row_ok) {
            const int head = (int)(row % p.H);
            const long long bq = row / p.H;
            const long long b = bq / p.Nq;
            const unsigned vrow = (unsigned)(b * p.S * HDV + head * G + lane);
            const float* loc_row = loc + row * p.LP * 2;
            const float* sw_row = w0 + row * p.LP;
            const float* lw_row = w1 + row * p.LP;
            TV* mrow = static_cast<TV*>(p.mask_out) + (bq * p.P * HD + (long long)head * p.D + lane * VEC);

            for (int c = split; c < nchunks; c += nsplit) {
                const int p0 = c * G;
                const int pm = p0 + lane;
                // ---- my point's taps for every level
                LanePoint tp[LB];
                float lwv[LB];
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    if (l < p.L) {
                        tp[l] = lane_point(loc_row + l * p.P * 2, sw_row + l * p.P, pm, p.P, lv.h[l], lv.w[l]);
                        lwv[l] = pm < p.P ? __ldg(lw_row + l * p.P + pm) : 0.f;
                    } else {
                        tp[l].inside = false; tp[l].x0 = tp[l].y0 = 0; tp[l].lx = tp[l].ly = tp[l].aw = 0.f;
                        lwv[l] = 0.f;
                    }
                }
                const int n_here = min(G, p.P - p0);
                // ---- walk the points of the chunk
#pragma unroll 1
                for (int o = 0; o < n_here; ++o) {
                    float macc[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) macc[i] = 0.f;
                    // levels in batches of LG: all 4*LG corner rows are requested (in storage form) before the
                    // first is used -- the walk is a chain of dependent gathers, memory-level parallelism is what
                    // it runs on (one level at a time left the kernel 5x above its issue bound)
#pragma unroll
                    for (int l0 = 0; l0 < LB; l0 += LG) {
                        if (l0 >= p.L) break;
                        typename V::Raw raw[LG][4];
                        float cw[LG][4], wsp[LG], wlv[LG];
#pragma unroll
                        for (int j = 0; j < LG; ++j) {
                            const int l = l0 + j;
                            const BTap t = bcast_tap<G>(tp[l], lwv[l], o, gm);      // (levels >= L carry inside = false)
                            const int lh = lv.h[l < p.L ? l : 0], lw = lv.w[l < p.L ? l : 0];
                            const float hx = 1.f - t.lx, hy = 1.f - t.ly;
                            const bool vx0 = t.x0 >= 0, vx1 = t.x0 + 1 <= lw - 1, vy0 = t.y0 >= 0, vy1 = t.y0 + 1 <= lh - 1;
                            const bool ok[4] = {t.inside && vy0 && vx0, t.inside && vy0 && vx1, t.inside && vy1 && vx0, t.inside && vy1 && vx1};
                            cw[j][0] = hy * hx; cw[j][1] = hy * t.lx; cw[j][2] = t.ly * hx; cw[j][3] = t.ly * t.lx;
                            wsp[j] = t.sw; wlv[j] = t.lw;
                            const unsigned c00 = vrow + ((unsigned)lv.start[l < p.L ? l : 0] + (unsigned)(t.y0 * lw + t.x0)) * HDV;
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                raw[j][k] = V::zero_raw();
                                if (ok[k]) raw[j][k] = V::load_raw(value16, c00 + ((k & 1) ? HDV : 0u) + ((k & 2) ? (unsigned)lw * HDV : 0u));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < LG; ++j) {
                            float v[4][VEC];
#pragma unroll
                            for (int k = 0; k < 4; ++k) V::unpack_raw(raw[j][k], v[k]);
#pragma unroll
                            for (int i = 0; i < VEC; ++i) {
                                const float val = cw[j][0] * v[0][i] + cw[j][1] * v[1][i] + cw[j][2] * v[2][i] + cw[j][3] * v[3][i];
                                acc[i] += val * wsp[j];       // instance_attn_kernel.cuh:354
                                macc[i] += val * wlv[j];      // :355
                            }
                        }
                    }
                    V::store(mrow + (long long)(p0 + o) * HD, macc);
                }
            }
        }

        TV* orow = static_cast<TV*>(p.out) + (row * p.D + lane * VEC);
        if (nsplit == 1) {
            if (row_ok) V::store(orow, acc);
        } else {
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

#ifndef BXR_INST_BWD_LG
#define BXR_INST_BWD_LG 2
#endif

template <typename TV, int G, int LB, typename ACC>
__global__ void __launch_bounds__(kThreads, BXR_INST_BWD_MINB) inst_bwd_own_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int LG = (VEC > 4) ? 1 : BXR_INST_BWD_LG;     // 8 channels per lane: no registers left for a second level (A/B, r01q)
    static_assert(LB % LG == 0, "level batches must tile LB");
    constexpr int GROUPS = kThreads / G;
    constexpr bool DET = sizeof(ACC) == 8;
    __shared__ LevelTable lv;
    load_levels(lv, p);

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const unsigned gm = group_mask<G>();
    const int nsplit = 1 << p.nsplit_log2;
    const int rows_per_unit = GROUPS >> p.nsplit_log2;
    const int r_local = gid >> p.nsplit_log2;
    const int split = gid & (nsplit - 1);
    const unsigned HDV = (unsigned)(p.H * p.D) / VEC;
    const long long HD = (long long)p.H * p.D;
    const void* __restrict__ value16 = p.value;   // indexed in lane-chunk units by V::load16
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    const float* __restrict__ w1 = static_cast<const float*>(p.w1);
    ACC* __restrict__ gacc = static_cast<ACC*>(p.grad_value_acc);
    float* __restrict__ grad_loc = static_cast<float*>(p.grad_loc);
    float* __restrict__ grad_w0 = static_cast<float*>(p.grad_w0);
    float* __restrict__ grad_w1 = static_cast<float*>(p.grad_w1);
    const int nchunks = (p.P + G - 1) / G;
    float dscale = 1.f;
    if constexpr (DET) dscale = *p.det_scale;

    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const long long row = (long long)u * rows_per_unit + r_local;
        if (row >= p.rows) continue;               // no block-wide barrier in this kernel
        const int head = (int)(row % p.H);
        const long long bq = row / p.H;
        const long long b = bq / p.Nq;
        const unsigned vrow = (unsigned)(b * p.S * HDV + head * G + lane);
        const float* loc_row = loc + row * p.LP * 2;
        const float* sw_row = w0 + row * p.LP;
        const float* lw_row = w1 + row * p.LP;
        const TV* gmrow = static_cast<const TV*>(p.grad_mask) + (bq * p.P * HD + (long long)head * p.D + lane * VEC);
        float go[VEC];
        V::load(static_cast<const TV*>(p.grad_out) + (row * p.D + lane * VEC), go);

        for (int c = split; c < nchunks; c += nsplit) {
            const int p0 = c * G;
            const int pm = p0 + lane;
            LanePoint tp[LB];
            float lwv[LB];
            float r_s[LB], r_l[LB], r_x[LB], r_y[LB];      // my point's gradients, per level
#pragma unroll
            for (int l = 0; l < LB; ++l) {
                r_s[l] = r_l[l] = r_x[l] = r_y[l] = 0.f;
                if (l < p.L) {
                    tp[l] = lane_point(loc_row + l * p.P * 2, sw_row + l * p.P, pm, p.P, lv.h[l], lv.w[l]);
                    lwv[l] = pm < p.P ? __ldg(lw_row + l * p.P + pm) : 0.f;
                } else {
                    tp[l].inside = false; tp[l].x0 = tp[l].y0 = 0; tp[l].lx = tp[l].ly = tp[l].aw = 0.f;
                    lwv[l] = 0.f;
                }
            }
            const int n_here = min(G, p.P - p0);
            // grad_mask is a DRAM stream read once: the next point's row is requested while this point is worked on
            const typename V::Raw* gmraw = reinterpret_cast<const typename V::Raw*>(gmrow);
            const long long gm_pitch = HD / VEC;
            typename V::Raw gm_next = __ldg(gmraw + (long long)p0 * gm_pitch);
#pragma unroll 1
            for (int o = 0; o < n_here; ++o) {
                float gmv[VEC];
                V::unpack_raw(gm_next, gmv);
                if (o + 1 < n_here) gm_next = __ldg(gmraw + (long long)(p0 + o + 1) * gm_pitch);
                // levels in batches of LG: the corner rows of the whole batch are requested before the first is used
#pragma unroll
                for (int l0 = 0; l0 < LB; l0 += LG) {
                    if (l0 >= p.L) break;
                    typename V::Raw raw[LG][4];
                    unsigned offs[LG];
                    BTap tb[LG];
#pragma unroll
                    for (int j = 0; j < LG; ++j) {
                        const int l = l0 + j;
                        tb[j] = bcast_tap<G>(tp[l], lwv[l], o, gm);             // (levels >= L carry inside = false)
                        const BTap& t = tb[j];
                        const int lc = l < p.L ? l : 0;
                        const int lh = lv.h[lc], lw = lv.w[lc];
                        const bool vx0 = t.x0 >= 0, vx1 = t.x0 + 1 <= lw - 1, vy0 = t.y0 >= 0, vy1 = t.y0 + 1 <= lh - 1;
                        const bool ok[4] = {t.inside && vy0 && vx0, t.inside && vy0 && vx1, t.inside && vy1 && vx0, t.inside && vy1 && vx1};
                        offs[j] = vrow + ((unsigned)lv.start[lc] + (unsigned)(t.y0 * lw + t.x0)) * HDV;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            raw[j][k] = V::zero_raw();
                            if (ok[k]) raw[j][k] = V::load_raw(value16, offs[j] + ((k & 1) ? HDV : 0u) + ((k & 2) ? (unsigned)lw * HDV : 0u));
                        }
                    }
#pragma unroll
                    for (int j = 0; j < LG; ++j) {
                        const int l = l0 + j;
                        const BTap& t = tb[j];
                        if (!t.inside) continue;                                  // uniform in the group
                        const int lh = lv.h[l], lw = lv.w[l];
                        const float hx = 1.f - t.lx, hy = 1.f - t.ly;
                        const bool vx0 = t.x0 >= 0, vx1 = t.x0 + 1 <= lw - 1, vy0 = t.y0 >= 0, vy1 = t.y0 + 1 <= lh - 1;
                        const bool ok[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
                        const float cw[4] = {hy * hx, hy * t.lx, t.ly * hx, t.ly * t.lx};
                        float tg[VEC];
#pragma unroll
                        for (int i = 0; i < VEC; ++i) tg[i] = go[i] * t.sw + gmv[i] * t.lw;    // instance_attn_kernel.cuh:139
                        float v[4][VEC];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            V::unpack_raw(raw[j][k], v[k]);
                            const unsigned off = offs[j] + ((k & 1) ? HDV : 0u) + ((k & 2) ? (unsigned)lw * HDV : 0u);
                            if (ok[k]) scatter_row<ACC, VEC>(gacc + (size_t)off * VEC, tg, cw[k], dscale);
                        }
                        float d[4] = {0.f, 0.f, 0.f, 0.f};    // d spatial_w, d level_w, d x, d y (partials over my channels)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            const float val = cw[0] * v[0][i] + cw[1] * v[1][i] + cw[2] * v[2][i] + cw[3] * v[3][i];
                            d[0] += go[i] * val;                                                  // :183
                            d[1] += gmv[i] * val;                                                 // :184
                            d[2] += (hy * (v[1][i] - v[0][i]) + t.ly * (v[3][i] - v[2][i])) * tg[i];
                            d[3] += (hx * (v[2][i] - v[0][i]) + t.lx * (v[3][i] - v[1][i])) * tg[i];
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) d[k] = gsum<G>(d[k], gm);
                        if (lane == o) {
                            r_s[l] = d[0];
                            r_l[l] = d[1];
                            r_x[l] = (float)lw * d[2];
                            r_y[l] = (float)lh * d[3];
                        }
                    }
                }
            }
            // my point's gradients, coalesced over the lanes of the group (zeros where the sample was outside)
            if (pm < p.P) {
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    if (l < p.L) {
                        const long long s = row * p.LP + (long long)l * p.P + pm;
                        grad_w0[s] = r_s[l];
                        grad_w1[s] = r_l[l];
                        reinterpret_cast<float2*>(grad_loc)[s] = make_float2(r_x[l], r_y[l]);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Round 2: the same two kernels with the taps in a shared-memory table instead of shuffles, and the backward's four
// per-point sums rebuilt from two dot products per corner.
//
// The owner-tap kernels above broadcast every (point, level) tap with 7 shuffles and then ALL G lanes redo the corner
// arithmetic (validity, offsets, weights) -- ~32 of the ~64 (forward) / ~150 (backward) instructions a (point, level)
// costs.  Here the owner lane does that arithmetic once and writes, per level, {4 corner offsets (kAbsent: no such
// corner), lx, ly, spatial_w, level_w} = 32 bytes to the group's table; the walk reads them back with two 16-byte
// shared loads.  The backward no longer forms val / dval-dx / dval-dy per channel (18 FMAs per channel and level): with
//     a_c = <grad_out, v_c>,  b_c = <grad_mask[p], v_c>          (two dot products per corner, 8 FMAs per channel and level)
// everything it needs is linear in them:
//     d spatial_w = sum_c cw_c a_c,   d level_w = sum_c cw_c b_c,   t_c = <top_grad, v_c> = sw a_c + lw b_c,
//     d x = W (hy (t_1 - t_0) + ly (t_3 - t_2)),   d y = H (hx (t_2 - t_0) + lx (t_3 - t_1))
// (instance_attn_kernel.cuh:139-186, re-associated).  The 8 partials are summed over the group's lanes with two 4-slot
// transpose reductions, parked in shared memory, and the owner lane of the point finishes them.
#ifndef BXR_INST_TAB
#define BXR_INST_TAB 1
#endif

struct InstTap {
    uint4 off;       // value offsets of the four corners in lane-chunk units, relative to the row's base; kAbsent = none
    float4 f;        // lx, ly, spatial_w, level_w
};

__device__ __forceinline__ void inst_write_tap(InstTap* dst, const LanePoint& t, float lw_, unsigned lbase, int lh, int lwid, unsigned HDV) {
    const bool vx0 = t.inside && t.x0 >= 0, vx1 = t.inside && t.x0 + 1 <= lwid - 1, vy0 = t.y0 >= 0, vy1 = t.y0 + 1 <= lh - 1;
    const unsigned c00 = lbase + (unsigned)(t.y0 * lwid + t.x0) * HDV;       // wraps for -1; such corners are absent
    InstTap e;
    e.off = make_uint4((vy0 && vx0) ? c00 : kAbsent, (vy0 && vx1) ? c00 + HDV : kAbsent,
                       (vy1 && vx0) ? c00 + (unsigned)lwid * HDV : kAbsent, (vy1 && vx1) ? c00 + (unsigned)lwid * HDV + HDV : kAbsent);
    e.f = make_float4(t.lx, t.ly, t.aw, lw_);
    *dst = e;
}

template <typename TV, int G, int LB>
__global__ void __launch_bounds__(kThreads, (Vec16<TV>::VEC > 4) ? BXR_INST_FWD_MINB_BF16 : BXR_INST_FWD_MINB_F32) inst_fwd_tab_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int LG = BXR_INST_LG;
    static_assert(LB % LG == 0, "level batches must tile LB");
    constexpr int GROUPS = kThreads / G;
    __shared__ LevelTable lv;
    __shared__ float s_red[kThreads * VEC];
    __shared__ __align__(16) InstTap s_tap[kThreads * LB];        // [group][point of the chunk][level]
    load_levels(lv, p);

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const unsigned gm = group_mask<G>();       // the table is per group: its barriers involve the group's lanes only
    const int nsplit = 1 << p.nsplit_log2;
    const int rows_per_unit = GROUPS >> p.nsplit_log2;
    const int r_local = gid >> p.nsplit_log2;
    const int split = gid & (nsplit - 1);
    const unsigned HDV = (unsigned)(p.H * p.D) / VEC;
    const long long HD = (long long)p.H * p.D;
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    const float* __restrict__ w1 = static_cast<const float*>(p.w1);
    const int nchunks = (p.P + G - 1) / G;
    InstTap* gtap = s_tap + gid * G * LB;

    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const long long row = (long long)u * rows_per_unit + r_local;
        const bool row_ok = row < p.rows;
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

        if (row_ok) {
            const int head = (int)(row % p.H);
            const long long bq = row / p.H;
            const long long b = bq / p.Nq;
            const typename V::Raw* vlane = static_cast<const typename V::Raw*>(p.value) + (size_t)(b * p.S * HDV + head * G + lane);
            const float* loc_row = loc + row * p.LP * 2;
            const float* sw_row = w0 + row * p.LP;
            const float* lw_row = w1 + row * p.LP;
            TV* mrow = static_cast<TV*>(p.mask_out) + (bq * p.P * HD + (long long)head * p.D + lane * VEC);

            for (int c = split; c < nchunks; c += nsplit) {
                const int p0 = c * G;
                const int pm = p0 + lane;
                // ---- my point's taps for every level, once, into the group's table
                __syncwarp(gm);                    // the previous chunk's entries have been read
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    if (l < p.L) {
                        const LanePoint t = lane_point(loc_row + l * p.P * 2, sw_row + l * p.P, pm, p.P, lv.h[l], lv.w[l]);
                        const float lwv = pm < p.P ? __ldg(lw_row + l * p.P + pm) : 0.f;
                        inst_write_tap(gtap + lane * LB + l, t, lwv, (unsigned)lv.start[l] * HDV, lv.h[l], lv.w[l], HDV);
                    }
                }
                __syncwarp(gm);
                const int n_here = min(G, p.P - p0);
#pragma unroll 1
                for (int o = 0; o < n_here; ++o) {
                    float macc[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) macc[i] = 0.f;
#pragma unroll
                    for (int l0 = 0; l0 < LB; l0 += LG) {
                        if (l0 >= p.L) break;
                        typename V::Raw raw[LG][4];
                        float4 f[LG];
#pragma unroll
                        for (int j = 0; j < LG; ++j) {
                            const int l = l0 + j;
                            InstTap t;
                            if (l < p.L) {
                                t = gtap[o * LB + l];
                            } else {
                                t.off = make_uint4(kAbsent, kAbsent, kAbsent, kAbsent);
                                t.f = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            f[j] = t.f;
                            const unsigned off[4] = {t.off.x, t.off.y, t.off.z, t.off.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                raw[j][k] = V::zero_raw();
                                if (off[k] != kAbsent) raw[j][k] = __ldg(vlane + off[k]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < LG; ++j) {
                            const float lx = f[j].x, ly = f[j].y, hx = 1.f - lx, hy = 1.f - ly;
                            const float cw[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
                            float v[4][VEC];
#pragma unroll
                            for (int k = 0; k < 4; ++k) V::unpack_raw(raw[j][k], v[k]);
#pragma unroll
                            for (int i = 0; i < VEC; ++i) {
                                const float val = cw[0] * v[0][i] + cw[1] * v[1][i] + cw[2] * v[2][i] + cw[3] * v[3][i];
                                acc[i] += val * f[j].z;       // instance_attn_kernel.cuh:354
                                macc[i] += val * f[j].w;      // :355
                            }
                        }
                    }
                    V::store(mrow + (long long)(p0 + o) * HD, macc);
                }
            }
        }

        TV* orow = static_cast<TV*>(p.out) + (row * p.D + lane * VEC);
        if (nsplit == 1) {
            if (row_ok) V::store(orow, acc);
        } else {
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

template <typename TV, int G, int LB, typename ACC>
__global__ void __launch_bounds__(kThreads, BXR_INST_BWD_MINB) inst_bwd_tab_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int GROUPS = kThreads / G;
    constexpr bool DET = sizeof(ACC) == 8;
    static_assert(G == 4 || G == 8 || G == 16, "transpose reduction");
    __shared__ LevelTable lv;
    __shared__ __align__(16) InstTap s_tap[kThreads * LB];
    __shared__ __align__(16) float s_sum[GROUPS * LB * 8];        // [group][level][a_0..a_3, b_0..b_3] of the current point
    load_levels(lv, p);

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const unsigned gm = group_mask<G>();
    const int nsplit = 1 << p.nsplit_log2;
    const int rows_per_unit = GROUPS >> p.nsplit_log2;
    const int r_local = gid >> p.nsplit_log2;
    const int split = gid & (nsplit - 1);
    const unsigned HDV = (unsigned)(p.H * p.D) / VEC;
    const long long HD = (long long)p.H * p.D;
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    const float* __restrict__ w1 = static_cast<const float*>(p.w1);
    ACC* __restrict__ gacc = static_cast<ACC*>(p.grad_value_acc);
    float* __restrict__ grad_loc = static_cast<float*>(p.grad_loc);
    float* __restrict__ grad_w0 = static_cast<float*>(p.grad_w0);
    float* __restrict__ grad_w1 = static_cast<float*>(p.grad_w1);
    const int nchunks = (p.P + G - 1) / G;
    InstTap* gtap = s_tap + gid * G * LB;
    float* gsum_ = s_sum + gid * LB * 8;
    float dscale = 1.f;
    if constexpr (DET) dscale = *p.det_scale;

    // (barriers and reductions use the group's mask: the groups of a warp walk different chunk ranges when a row's chunks
    // are split over groups; rows past the end are carried along inactive)
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const long long row_raw = (long long)u * rows_per_unit + r_local;
        const bool row_ok = row_raw < p.rows;
        const long long row = row_ok ? row_raw : 0;
        const int head = (int)(row % p.H);
        const long long bq = row / p.H;
        const long long b = bq / p.Nq;
        const size_t vrow = (size_t)(b * p.S * HDV + head * G + lane);
        const typename V::Raw* vlane = static_cast<const typename V::Raw*>(p.value) + vrow;
        ACC* glane = gacc + vrow * VEC;
        const float* loc_row = loc + row * p.LP * 2;
        const float* sw_row = w0 + row * p.LP;
        const float* lw_row = w1 + row * p.LP;
        const TV* gmrow = static_cast<const TV*>(p.grad_mask) + (bq * p.P * HD + (long long)head * p.D + lane * VEC);
        float go[VEC];
        V::load(static_cast<const TV*>(p.grad_out) + (row * p.D + lane * VEC), go);

        for (int c = split; c < nchunks; c += nsplit) {
            const int p0 = c * G;
            const int pm = p0 + lane;
            float r_s[LB], r_l[LB], r_x[LB], r_y[LB];      // my point's gradients, per level
            __syncwarp(gm);
#pragma unroll
            for (int l = 0; l < LB; ++l) {
                r_s[l] = r_l[l] = r_x[l] = r_y[l] = 0.f;
                if (l < p.L) {
                    const LanePoint t = lane_point(loc_row + l * p.P * 2, sw_row + l * p.P, row_ok ? pm : p.P, p.P, lv.h[l], lv.w[l]);
                    const float lwv = (row_ok && pm < p.P) ? __ldg(lw_row + l * p.P + pm) : 0.f;
                    inst_write_tap(gtap + lane * LB + l, t, lwv, (unsigned)lv.start[l] * HDV, lv.h[l], lv.w[l], HDV);
                }
            }
            __syncwarp(gm);
            const int n_here = min(G, p.P - p0);
            const typename V::Raw* gmraw = reinterpret_cast<const typename V::Raw*>(gmrow);
            const long long gm_pitch = HD / VEC;
            typename V::Raw gm_next = __ldg(gmraw + (long long)p0 * gm_pitch);
#pragma unroll 1
            for (int o = 0; o < n_here; ++o) {
                float gmv[VEC];
                V::unpack_raw(gm_next, gmv);
                if (o + 1 < n_here) gm_next = __ldg(gmraw + (long long)(p0 + o + 1) * gm_pitch);
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    if (l >= p.L) break;
                    const InstTap t = gtap[o * LB + l];
                    const unsigned off[4] = {t.off.x, t.off.y, t.off.z, t.off.w};
                    typename V::Raw raw[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        raw[k] = V::zero_raw();
                        if (off[k] != kAbsent) raw[k] = __ldg(vlane + off[k]);
                    }
                    const float lx = t.f.x, ly = t.f.y, hx = 1.f - lx, hy = 1.f - ly;
                    const float cw[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
                    float tg[VEC];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) tg[i] = go[i] * t.f.z + gmv[i] * t.f.w;    // instance_attn_kernel.cuh:139
                    float a[4], bb[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float v[VEC];
                        V::unpack_raw(raw[k], v);
                        float sa = 0.f, sb = 0.f;
#pragma unroll
                        for (int i = 0; i < VEC; ++i) { sa += go[i] * v[i]; sb += gmv[i] * v[i]; }
                        a[k] = sa; bb[k] = sb;
                        if (row_ok && off[k] != kAbsent) scatter_row<ACC, VEC>(glane + (size_t)off[k] * VEC, tg, cw[k], dscale);
                    }
                    float ta, tb;
                    const int ia = reduce4<G>(a, ta, lane, gm);
                    reduce4<G>(bb, tb, lane, gm);
                    gsum_[l * 8 + ia] = ta;             // lanes holding the same index hold the same total
                    gsum_[l * 8 + 4 + ia] = tb;
                }
                __syncwarp(gm);
                if (lane == o) {
#pragma unroll
                    for (int l = 0; l < LB; ++l) {
                        if (l < p.L) {
                            const float4 A = *reinterpret_cast<const float4*>(gsum_ + l * 8);
                            const float4 Bv = *reinterpret_cast<const float4*>(gsum_ + l * 8 + 4);
                            const InstTap t = gtap[o * LB + l];
                            const float lx = t.f.x, ly = t.f.y, hx = 1.f - lx, hy = 1.f - ly, sw = t.f.z, lw_ = t.f.w;
                            r_s[l] = hy * hx * A.x + hy * lx * A.y + ly * hx * A.z + ly * lx * A.w;      // :183
                            r_l[l] = hy * hx * Bv.x + hy * lx * Bv.y + ly * hx * Bv.z + ly * lx * Bv.w;  // :184
                            const float t0 = sw * A.x + lw_ * Bv.x, t1 = sw * A.y + lw_ * Bv.y, t2 = sw * A.z + lw_ * Bv.z, t3 = sw * A.w + lw_ * Bv.w;
                            r_x[l] = (float)lv.w[l] * (hy * (t1 - t0) + ly * (t3 - t2));
                            r_y[l] = (float)lv.h[l] * (hx * (t2 - t0) + lx * (t3 - t1));
                        }
                    }
                }
                __syncwarp(gm);      // the sums are overwritten by the next point
            }
            if (row_ok && pm < p.P) {
#pragma unroll
                for (int l = 0; l < LB; ++l) {
                    if (l < p.L) {
                        const long long s = row * p.LP + (long long)l * p.P + pm;
                        grad_w0[s] = r_s[l];
                        grad_w1[s] = r_l[l];
                        reinterpret_cast<float2*>(grad_loc)[s] = make_float2(r_x[l], r_y[l]);
                    }
                }
            }
        }
    }
}

}  // namespace bxr
