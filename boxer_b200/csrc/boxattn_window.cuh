// "Footprint window" kernels for box attention (the encoder / large-Nq hot path).
//
// Observation (ncu, profiles/r01a_*): the straightforward per-point kernel is issue- and
// L1-wavefront-bound, not DRAM- or L2-bound -- every one of the P points of a (query, head,
// level) re-derives its tap on all G lanes and fetches its own 4 corner rows, although the
// K x K points of a box land on a handful of shared pixels (a 4 px box with a 4 x 4 grid covers
// ~5 x 5 pixels at its own level and 2 x 2 .. 3 x 3 at the coarser ones: ~50 unique pixels for
// 256 corner fetches).
//
// So, per (row = (b, q, head), level):
//   A. the G lanes of the row's group each take P/G points (coalesced loc / weight loads),
//      compute their taps once, and min/max-reduce the touched pixel range with shuffles;
//   B. if the range fits an 8 x 8 window, the lanes scatter  attn * bilinear weight  of their
//      points' corners into a per-group window of pixel weights in shared memory;
//   C. the group walks the window: one 16-byte row load per *unique* pixel,
//         forward :  acc        += W[pix] * value[pix]
//         backward:  grad_value[pix] += W[pix] * grad_out            (one red.v4 per unique pixel)
//                    d[pix]      = <grad_out, value[pix]>           (one shuffle reduction per pixel)
//   D. (backward) every lane finishes its own points from the d window -- scalar math only:
//         grad_attn = sum_c cw_c d_c,   grad_x = W * attn * (hy (d01 - d00) + ly (d11 - d10)), ...
//   If the range does not fit (a coarse-level query looking at a fine level, or arbitrary
//   locations), the group falls back to the per-point path for that (row, level), still with the
//   taps computed once by their owner lane and broadcast by shuffle.
// The result is the same sum as the reference's (box_attn_kernel.cuh:311-346), re-associated.
#pragma once

#include "boxattn_kernels.cuh"

namespace bxr {

constexpr int kWinSide = 8;
constexpr int kWinSlots = kWinSide * kWinSide;

template <int G>
__device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return 0xffffffffu;
    const unsigned lane_w = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane_w & ~(unsigned)(G - 1));
}

template <int G>
__device__ __forceinline__ int gmin(int v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(m, v, o));
    return v;
}
template <int G>
__device__ __forceinline__ int gmax(int v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o));
    return v;
}
template <int G>
__device__ __forceinline__ float gsum(float v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}

// a lane's own sample point at one level
struct LanePoint {
    float lx, ly, aw;
    int x0, y0;       // floor of the pixel coordinates (meaningful only if inside)
    bool inside;      // window test of box_attn_kernel.cuh:328 (false for padding lanes)
};

__device__ __forceinline__ LanePoint lane_point(const float* __restrict__ loc_l, const float* __restrict__ w_l,
                                                int pt, int P, int h, int w) {
    LanePoint t;
    const bool act = pt < P;
    const int pc = act ? pt : 0;
    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc_l) + pc);
    t.aw = __ldg(w_l + pc);
    const float x = xy.x * (float)w - 0.5f;
    const float y = xy.y * (float)h - 0.5f;
    t.inside = act && (y > -1.f) && (x > -1.f) && (y < (float)h) && (x < (float)w);
    const float xs = t.inside ? x : 0.f, ys = t.inside ? y : 0.f;
    const float xf = floorf(xs), yf = floorf(ys);
    t.x0 = (int)xf;
    t.y0 = (int)yf;
    t.lx = xs - xf;
    t.ly = ys - yf;
    return t;
}

// ------------------------------------------------------------------------------------------------
// Forward.  One group of G lanes per row; grid-stride over contiguous blocks of rows.
template <typename TV, int G, int PPL>
__global__ void __launch_bounds__(kThreads) box_fwd_win_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int GROUPS = kThreads / G;
    __shared__ LevelTable lv;
    __shared__ __align__(16) float s_win[GROUPS][kWinSlots];
    load_levels(lv, p);

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const unsigned gm = group_mask<G>();
    float* win = s_win[gid];
    const int HD = p.H * p.D;
    const TV* __restrict__ value = static_cast<const TV*>(p.value);
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);

    int u0, u1;
    unit_range(p.units, u0, u1);
    for (int u = u0; u < u1; ++u) {
        const long long row = (long long)u * GROUPS + gid;
        if (row >= p.rows) continue;            // whole group leaves; everything below is group-scoped
        const int head = (int)(row % p.H);
        const long long b = row / ((long long)p.H * p.Nq);
        const TV* vrow = value + (b * p.S * HD + head * p.D + lane * VEC);
        const float* loc_row = loc + row * p.LP * 2;
        const float* w_row = w0 + row * p.LP;

        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

        for (int l = 0; l < p.L; ++l) {
            const int lh = lv.h[l], lw = lv.w[l];
            const TV* vlev = vrow + lv.start[l] * HD;
            // ---- A: own points, touched pixel range
            LanePoint pt[PPL];
            int bx0 = 0x7fffffff, bx1 = -0x7fffffff, by0 = 0x7fffffff, by1 = -0x7fffffff;
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                pt[k] = lane_point(loc_row + l * p.P * 2, w_row + l * p.P, lane + k * G, p.P, lh, lw);
                if (pt[k].inside) {
                    bx0 = min(bx0, pt[k].x0); bx1 = max(bx1, pt[k].x0 + 1);
                    by0 = min(by0, pt[k].y0); by1 = max(by1, pt[k].y0 + 1);
                }
            }
            const int X0 = max(gmin<G>(bx0, gm), 0), X1 = min(gmax<G>(bx1, gm), lw - 1);
            const int Y0 = max(gmin<G>(by0, gm), 0), Y1 = min(gmax<G>(by1, gm), lh - 1);
            const int nx = X1 - X0 + 1, ny = Y1 - Y0 + 1;
            if (nx <= 0 || ny <= 0) continue;   // no point of this level passed the window test

            if (nx <= kWinSide && ny <= kWinSide) {
                // ---- B: scatter pixel weights into the window
                for (int s = lane; s < ny * kWinSide; s += G) win[s] = 0.f;
                __syncwarp(gm);
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    if (pt[k].inside) {
                        const int sx = pt[k].x0 - X0, sy = pt[k].y0 - Y0;   // -1 .. n-1
                        const float hx = 1.f - pt[k].lx, hy = 1.f - pt[k].ly;
                        const bool vx0 = sx >= 0, vx1 = sx + 1 < nx, vy0 = sy >= 0, vy1 = sy + 1 < ny;
                        float* wp = win + sy * kWinSide + sx;
                        if (vy0 && vx0) atomicAdd(wp, hy * hx * pt[k].aw);
                        if (vy0 && vx1) atomicAdd(wp + 1, hy * pt[k].lx * pt[k].aw);
                        if (vy1 && vx0) atomicAdd(wp + kWinSide, pt[k].ly * hx * pt[k].aw);
                        if (vy1 && vx1) atomicAdd(wp + kWinSide + 1, pt[k].ly * pt[k].lx * pt[k].aw);
                    }
                }
                __syncwarp(gm);
                // ---- C: one row load per unique pixel
                const TV* wbase = vlev + ((long long)Y0 * lw + X0) * HD;
                for (int iy = 0; iy < ny; ++iy) {
                    const float4 wa = *reinterpret_cast<const float4*>(win + iy * kWinSide);
                    const float4 wb = *reinterpret_cast<const float4*>(win + iy * kWinSide + 4);
                    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                    const TV* rbase = wbase + (long long)iy * lw * HD;
                    float v[8][VEC];
#pragma unroll
                    for (int ix = 0; ix < 8; ++ix) {
                        if (ix < nx && wv[ix] != 0.f) {
                            V::load(rbase + ix * HD, v[ix]);
                        } else {
#pragma unroll
                            for (int i = 0; i < VEC; ++i) v[ix][i] = 0.f;
                        }
                    }
#pragma unroll
                    for (int ix = 0; ix < 8; ++ix)
#pragma unroll
                        for (int i = 0; i < VEC; ++i) acc[i] += wv[ix] * v[ix][i];
                }
                __syncwarp(gm);   // the window is re-zeroed by the next level
            } else {
                // ---- per-point fallback: owner lane broadcasts its tap
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    for (int o = 0; o < G; ++o) {
                        if (o + k * G >= p.P) break;            // uniform in the group
                        const bool inside = __shfl_sync(gm, (int)pt[k].inside, o, G) != 0;
                        const int x0 = __shfl_sync(gm, pt[k].x0, o, G), y0 = __shfl_sync(gm, pt[k].y0, o, G);
                        const float lx = __shfl_sync(gm, pt[k].lx, o, G), ly = __shfl_sync(gm, pt[k].ly, o, G);
                        const float aw = __shfl_sync(gm, pt[k].aw, o, G);
                        if (!inside) continue;
                        const float hx = 1.f - lx, hy = 1.f - ly;
                        const bool vx0 = x0 >= 0, vx1 = x0 + 1 <= lw - 1, vy0 = y0 >= 0, vy1 = y0 + 1 <= lh - 1;
                        const bool ok[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
                        const float cw[4] = {hy * hx * aw, hy * lx * aw, ly * hx * aw, ly * lx * aw};
                        const TV* c00 = vlev + ((long long)y0 * lw + x0) * HD;
                        float v[4][VEC];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (ok[c]) {
                                V::load(c00 + ((c & 1) ? HD : 0) + ((c & 2) ? (long long)lw * HD : 0), v[c]);
                            } else {
#pragma unroll
                                for (int i = 0; i < VEC; ++i) v[c][i] = 0.f;
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int i = 0; i < VEC; ++i) acc[i] += cw[c] * v[c][i];
                    }
                }
            }
        }
        V::store(static_cast<TV*>(p.out) + (row * p.D + lane * VEC), acc);
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.
template <typename ACC, int VEC>
__device__ __forceinline__ void scatter_row(ACC* dst, const float (&g)[VEC], float wgt, float dscale) {
    if constexpr (sizeof(ACC) == 8) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) red_add_fixed(reinterpret_cast<long long*>(dst) + i, wgt * g[i], dscale);
    } else {
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
            red_add_v4(reinterpret_cast<float*>(dst) + i, wgt * g[i], wgt * g[i + 1], wgt * g[i + 2], wgt * g[i + 3]);
    }
}

template <typename TV, int G, int PPL, typename ACC>
__global__ void __launch_bounds__(kThreads) box_bwd_win_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int GROUPS = kThreads / G;
    constexpr bool DET = sizeof(ACC) == 8;
    __shared__ LevelTable lv;
    __shared__ __align__(16) float s_win[GROUPS][kWinSlots];   // pixel weights  W[pix]
    __shared__ __align__(16) float s_dot[GROUPS][kWinSlots];   // d[pix] = <grad_out, value[pix]>
    load_levels(lv, p);

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const unsigned gm = group_mask<G>();
    float* win = s_win[gid];
    float* dot = s_dot[gid];
    const int HD = p.H * p.D;
    const TV* __restrict__ value = static_cast<const TV*>(p.value);
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    ACC* __restrict__ gacc = static_cast<ACC*>(p.grad_value_acc);
    float* __restrict__ grad_loc = static_cast<float*>(p.grad_loc);
    float* __restrict__ grad_w0 = static_cast<float*>(p.grad_w0);
    float dscale = 1.f;
    if constexpr (DET) dscale = *p.det_scale;

    int u0, u1;
    unit_range(p.units, u0, u1);
    for (int u = u0; u < u1; ++u) {
        const long long row = (long long)u * GROUPS + gid;
        if (row >= p.rows) continue;
        const int head = (int)(row % p.H);
        const long long b = row / ((long long)p.H * p.Nq);
        const long long vbase = b * p.S * HD + head * p.D + lane * VEC;
        const float* loc_row = loc + row * p.LP * 2;
        const float* w_row = w0 + row * p.LP;
        float go[VEC];
        V::load(static_cast<const TV*>(p.grad_out) + (row * p.D + lane * VEC), go);

        for (int l = 0; l < p.L; ++l) {
            const int lh = lv.h[l], lw = lv.w[l];
            const long long lbase = vbase + lv.start[l] * HD;
            LanePoint pt[PPL];
            float g_a[PPL], g_x[PPL], g_y[PPL];     // this lane's results for its own points
            int bx0 = 0x7fffffff, bx1 = -0x7fffffff, by0 = 0x7fffffff, by1 = -0x7fffffff;
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                pt[k] = lane_point(loc_row + l * p.P * 2, w_row + l * p.P, lane + k * G, p.P, lh, lw);
                g_a[k] = g_x[k] = g_y[k] = 0.f;
                if (pt[k].inside) {
                    bx0 = min(bx0, pt[k].x0); bx1 = max(bx1, pt[k].x0 + 1);
                    by0 = min(by0, pt[k].y0); by1 = max(by1, pt[k].y0 + 1);
                }
            }
            const int X0 = max(gmin<G>(bx0, gm), 0), X1 = min(gmax<G>(bx1, gm), lw - 1);
            const int Y0 = max(gmin<G>(by0, gm), 0), Y1 = min(gmax<G>(by1, gm), lh - 1);
            const int nx = X1 - X0 + 1, ny = Y1 - Y0 + 1;

            if (nx > 0 && ny > 0 && nx <= kWinSide && ny <= kWinSide) {
                // B: pixel weights.  The d window doubles as a "touched" flag (1.0) until C overwrites it
                //    with the dot products: a pixel touched with zero total weight still needs its d.
                for (int s = lane; s < ny * kWinSide; s += G) { win[s] = 0.f; dot[s] = 0.f; }
                __syncwarp(gm);
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    if (pt[k].inside) {
                        const int sx = pt[k].x0 - X0, sy = pt[k].y0 - Y0;
                        const float hx = 1.f - pt[k].lx, hy = 1.f - pt[k].ly;
                        const bool vx0 = sx >= 0, vx1 = sx + 1 < nx, vy0 = sy >= 0, vy1 = sy + 1 < ny;
                        const int s00 = sy * kWinSide + sx;
                        if (vy0 && vx0) { atomicAdd(win + s00, hy * hx * pt[k].aw); dot[s00] = 1.f; }
                        if (vy0 && vx1) { atomicAdd(win + s00 + 1, hy * pt[k].lx * pt[k].aw); dot[s00 + 1] = 1.f; }
                        if (vy1 && vx0) { atomicAdd(win + s00 + kWinSide, pt[k].ly * hx * pt[k].aw); dot[s00 + kWinSide] = 1.f; }
                        if (vy1 && vx1) { atomicAdd(win + s00 + kWinSide + 1, pt[k].ly * pt[k].lx * pt[k].aw); dot[s00 + kWinSide + 1] = 1.f; }
                    }
                }
                __syncwarp(gm);
                // C: per unique pixel: value row, scatter W*go, d = <go, v>
                const long long wbase = lbase + ((long long)Y0 * lw + X0) * HD;
                for (int iy = 0; iy < ny; ++iy) {
                    const float4 wa = *reinterpret_cast<const float4*>(win + iy * kWinSide);
                    const float4 wb = *reinterpret_cast<const float4*>(win + iy * kWinSide + 4);
                    const float4 ta = *reinterpret_cast<const float4*>(dot + iy * kWinSide);
                    const float4 tb = *reinterpret_cast<const float4*>(dot + iy * kWinSide + 4);
                    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                    const float tv[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
                    const long long rbase = wbase + (long long)iy * lw * HD;
                    float v[8][VEC];
#pragma unroll
                    for (int ix = 0; ix < 8; ++ix) {
                        if (ix < nx && tv[ix] != 0.f) {
                            V::load(value + rbase + ix * HD, v[ix]);
                        } else {
#pragma unroll
                            for (int i = 0; i < VEC; ++i) v[ix][i] = 0.f;
                        }
                    }
                    __syncwarp(gm);      // all lanes have read the touched flags of this window row
                    float dsum[8];
#pragma unroll
                    for (int ix = 0; ix < 8; ++ix) {
                        float s = 0.f;
#pragma unroll
                        for (int i = 0; i < VEC; ++i) s += go[i] * v[ix][i];
                        dsum[ix] = s;
                        if (ix < nx && wv[ix] != 0.f) scatter_row<ACC, VEC>(gacc + rbase + ix * HD, go, wv[ix], dscale);
                    }
                    // transpose-reduce the 8 partial dot products over the G lanes
#pragma unroll
                    for (int ix = 0; ix < 8; ++ix) dsum[ix] = gsum<G>(dsum[ix], gm);
                    if (lane == 0) {
                        *reinterpret_cast<float4*>(dot + iy * kWinSide) = make_float4(dsum[0], dsum[1], dsum[2], dsum[3]);
                        *reinterpret_cast<float4*>(dot + iy * kWinSide + 4) = make_float4(dsum[4], dsum[5], dsum[6], dsum[7]);
                    }
                }
                __syncwarp(gm);
                // D: finish own points from the d window
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    if (pt[k].inside) {
                        const int sx = pt[k].x0 - X0, sy = pt[k].y0 - Y0;
                        const float lx = pt[k].lx, ly = pt[k].ly, hx = 1.f - lx, hy = 1.f - ly;
                        const bool vx0 = sx >= 0, vx1 = sx + 1 < nx, vy0 = sy >= 0, vy1 = sy + 1 < ny;
                        const int s00 = sy * kWinSide + sx;
                        const float d00 = (vy0 && vx0) ? dot[s00] : 0.f;
                        const float d01 = (vy0 && vx1) ? dot[s00 + 1] : 0.f;
                        const float d10 = (vy1 && vx0) ? dot[s00 + kWinSide] : 0.f;
                        const float d11 = (vy1 && vx1) ? dot[s00 + kWinSide + 1] : 0.f;
                        g_a[k] = hy * hx * d00 + hy * lx * d01 + ly * hx * d10 + ly * lx * d11;
                        g_x[k] = (float)lw * pt[k].aw * (hy * (d01 - d00) + ly * (d11 - d10));
                        g_y[k] = (float)lh * pt[k].aw * (hx * (d10 - d00) + lx * (d11 - d01));
                    }
                }
                __syncwarp(gm);
            } else if (nx > 0 && ny > 0) {
                // per-point fallback (window too large)
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    for (int o = 0; o < G; ++o) {
                        if (o + k * G >= p.P) break;
                        const bool inside = __shfl_sync(gm, (int)pt[k].inside, o, G) != 0;
                        const int x0 = __shfl_sync(gm, pt[k].x0, o, G), y0 = __shfl_sync(gm, pt[k].y0, o, G);
                        const float lx = __shfl_sync(gm, pt[k].lx, o, G), ly = __shfl_sync(gm, pt[k].ly, o, G);
                        const float aw = __shfl_sync(gm, pt[k].aw, o, G);
                        if (!inside) continue;
                        const float hx = 1.f - lx, hy = 1.f - ly;
                        const bool vx0 = x0 >= 0, vx1 = x0 + 1 <= lw - 1, vy0 = y0 >= 0, vy1 = y0 + 1 <= lh - 1;
                        const bool ok[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
                        const float cw[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
                        const long long c00 = lbase + ((long long)y0 * lw + x0) * HD;
                        float d[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float s = 0.f;
                            if (ok[c]) {
                                const long long off = c00 + ((c & 1) ? HD : 0) + ((c & 2) ? (long long)lw * HD : 0);
                                float v[VEC];
                                V::load(value + off, v);
#pragma unroll
                                for (int i = 0; i < VEC; ++i) s += go[i] * v[i];
                                scatter_row<ACC, VEC>(gacc + off, go, cw[c] * aw, dscale);
                            }
                            d[c] = s;
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) d[c] = gsum<G>(d[c], gm);
                        if (lane == o) {
                            g_a[k] = cw[0] * d[0] + cw[1] * d[1] + cw[2] * d[2] + cw[3] * d[3];
                            g_x[k] = (float)lw * aw * (hy * (d[1] - d[0]) + ly * (d[3] - d[2]));
                            g_y[k] = (float)lh * aw * (hx * (d[2] - d[0]) + lx * (d[3] - d[1]));
                        }
                    }
                }
            }
            // coalesced stores of this level's gradients (zeros for points outside the window test)
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                const int ptn = lane + k * G;
                if (ptn < p.P) {
                    const long long s = row * p.LP + (long long)l * p.P + ptn;
                    grad_w0[s] = g_a[k];
                    reinterpret_cast<float2*>(grad_loc)[s] = make_float2(g_x[k], g_y[k]);
                }
            }
        }
    }
}

}  // namespace bxr
