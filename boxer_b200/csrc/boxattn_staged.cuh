// "Staged row" kernels for box attention: the footprint-window algorithm of boxattn_window.cuh with
//   * the per-row operands (sampling locations or boxes, attention weights) streamed into shared
//     memory by the TMA unit (cp.async.bulk + mbarrier, one private ring per warp): they are the
//     kernel's whole DRAM stream (loc + weights = 137 MB of the 160 MB a COCO-size forward reads
//     once) and arrive while the previous rows are being processed -- the ncu line profile of the
//     window kernels (profiles/r01m_*) had 10 % of all stall samples on the first use of a
//     location load;
//   * all levels of a row scattered first into ONE pool of window slots, then a single walk over
//     the pool: one exposed gather latency per row instead of one per level, L+1 group barriers
//     instead of 3L, and every slot carries its own pixel offset so the walk has no index
//     arithmetic (the (ix, iy) stepping was 12 % of the forward's instructions);
//   * one fixed-point scale per row (sum |attn| over all levels) instead of one per level.
// Arithmetic is the window kernels' (and the reference's, box_attn_kernel.cuh:34-184,311-346),
// re-associated.
#pragma once

#include "boxattn_window.cuh"

namespace bxr {

#ifndef BXR_STG_POOL
#define BXR_STG_POOL 128          // window slots per row, shared by all its levels
#endif
#ifndef BXR_STG_FWD_MINB
#define BXR_STG_FWD_MINB 3
#endif
#ifndef BXR_STG_BWD_MINB
#define BXR_STG_BWD_MINB 3
#endif

constexpr int kPool = BXR_STG_POOL;
constexpr int kPoolPitch = kPool + 4;                 // words; [weights | offsets] per group = 2 pitches (8 mod 32 banks apart)
constexpr unsigned kNoPixel = 0xffffffffu;
constexpr int kStgMaxPasses = 4;                       // passes (of LPP levels) scattered before one walk

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Shared-memory plan, computed identically by the host (launch size) and the device.
//   [ pools: GROUPS x 2 x kPoolPitch words ][ per warp, per stage: operands of the warp's RW rows ]
// MODE 0 stage: loc (RW*LP*2 floats) | weights (RW*LP floats)
// MODE 1 stage: boxes (RW*L*4 floats) | weights (RW*LP floats)      (angles: one float per level, read directly)
// Every piece is a multiple of 16 bytes when LP % 4 == 0 (host-side condition for these kernels).
struct StgPlan {
    unsigned pool_bytes, stage_bytes, total_bytes;
    unsigned off_w;               // byte offset of the weights inside a stage
};
__host__ __device__ inline StgPlan stg_plan(int G, int mode, int L, int LP, int stages) {
    const unsigned groups = kThreads / G, rw = 32 / G, warps = kThreads / 32;
    StgPlan s;
    s.pool_bytes = groups * 2u * kPoolPitch * 4u;
    s.off_w = mode == 0 ? rw * LP * 8u : rw * L * 16u;
    s.stage_bytes = (s.off_w + rw * LP * 4u + 15u) & ~15u;
    s.total_bytes = s.pool_bytes + warps * stages * s.stage_bytes;
    return s;
}

// one warp's bulk copies for unit u into `stage`; called by one lane
template <int G, int MODE>
__device__ __forceinline__ void stg_issue(const AttnParams& p, const StgPlan& pl, unsigned char* stage, unsigned long long* bar,
                                          long long row0) {
    constexpr int RW = 32 / G;
    const long long left = p.rows - row0;
    const unsigned n = left < RW ? (unsigned)left : (unsigned)RW;
    const unsigned wb = n * p.LP * 4u;
    if (MODE == 0) {
        mbar_expect_tx(bar, wb * 3u);
        bulk_g2s(stage, static_cast<const float*>(p.loc) + row0 * p.LP * 2, wb * 2u, bar);
        bulk_g2s(stage + pl.off_w, static_cast<const float*>(p.w0) + row0 * p.LP, wb, bar);
    } else {
        const unsigned bb = n * p.L * 16u;
        mbar_expect_tx(bar, wb + bb);
        bulk_g2s(stage, static_cast<const float*>(p.boxes) + row0 * p.L * 4, bb, bar);
        bulk_g2s(stage + pl.off_w, static_cast<const float*>(p.w0) + row0 * p.LP, wb, bar);
    }
}

// the box of (row, level) from the staged operands
__device__ __forceinline__ LevelBox stg_level_box(const AttnParams& p, const float* sbox, long long row, int l, long long b) {
    LevelBox q;
    const float4 bx = *reinterpret_cast<const float4*>(sbox + l * 4);
    q.cx = bx.x; q.cy = bx.y;
    q.pos_w = bx.z > 0.f; q.pos_h = bx.w > 0.f;
    q.sx = q.pos_w ? bx.z : 0.f; q.sy = q.pos_h ? bx.w : 0.f;
    q.cs = 1.f; q.sn = 0.f;
    if (p.angles) sincosf(__ldg(static_cast<const float*>(p.angles) + (row * p.L + l)), &q.sn, &q.cs);
    q.vx = q.vy = 1.f;
    if (p.valid_ratios) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p.valid_ratios) + (b * p.L + l));
        q.vx = v.x; q.vy = v.y;
    }
    return q;
}

// a lane's point ptn of level l of its row, from the staged operands
template <int MODE>
__device__ __forceinline__ LanePoint stg_point(const AttnParams& p, const float* srow, const float* swrow, const LevelBox& lbx,
                                               int l, int ptn, int h, int w) {
    const bool act = ptn < p.P;
    const int pc = act ? ptn : 0;
    const float aw = swrow[l * p.P + pc];
    float x, y;
    if (MODE == 0) {
        const float2 xy = *reinterpret_cast<const float2*>(srow + (l * p.P + pc) * 2);
        x = xy.x; y = xy.y;
    } else {
        const float2 k = __ldg(reinterpret_cast<const float2*>(p.kidx) + pc);
        box_point(lbx, k.x, k.y, x, y);
    }
    return lane_point_xy(x, y, aw, act, h, w);
}

// window of one (row, level) inside the row's pool
struct PoolWin {
    int X0, Y0, nx, ny;
    int base;             // first slot
    int mode;             // 0 skip, 1 window, 2 per-point
};

// ------------------------------------------------------------------------------------------------
// Forward.
template <typename TV, int G, int SUB, int PPL, int MODE>
__global__ void __launch_bounds__(kThreads, BXR_STG_FWD_MINB) box_fwd_stg_kernel(const AttnParams p) {
    using V = Vec16<TV>;
    constexpr int VEC = V::VEC;
    constexpr int GROUPS = kThreads / G, RW = 32 / G, LPP = G / SUB;
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ LevelTable lv;
    __shared__ __align__(8) unsigned long long s_bar[kThreads / 32];

    const int warp = threadIdx.x >> 5, lane_w = threadIdx.x & 31;
    if (lane_w == 0) mbar_init(&s_bar[warp], 1);
    mbar_fence_init();
    load_levels(lv, p);                       // (__syncthreads inside)

    const StgPlan pl = stg_plan(G, MODE, p.L, p.LP, 1);
    const int lane = threadIdx.x % G, gid = threadIdx.x / G, gi = gid % RW;
    const int sub = lane / SUB, slane = lane % SUB;
    const unsigned gm = group_mask<G>();
    int* pw = reinterpret_cast<int*>(dsm) + gid * (2 * kPoolPitch);          // pixel weights (fixed point)
    unsigned* po = reinterpret_cast<unsigned*>(pw) + kPoolPitch;               // pixel offsets (lane-chunk units)
    unsigned char* stage = dsm + pl.pool_bytes + warp * pl.stage_bytes;
    unsigned long long* bar = &s_bar[warp];
    const float* srow = reinterpret_cast<const float*>(stage) + (MODE == 0 ? gi * p.LP * 2 : gi * p.L * 4);
    const float* swrow = reinterpret_cast<const float*>(stage + pl.off_w) + gi * p.LP;
    const unsigned HDV = (unsigned)(p.H * p.D) / VEC;
    const void* __restrict__ value16 = p.value;

    unsigned phase = 0;
    {
        const long long r0 = (long long)blockIdx.x * GROUPS + warp * RW;
        if (lane_w == 0 && blockIdx.x < p.units && r0 < p.rows) stg_issue<G, MODE>(p, pl, stage, bar, r0);
    }
    for (int u = blockIdx.x; u < p.units; u += gridDim.x) {
        const long long row0 = (long long)u * GROUPS + warp * RW;
        if (row0 >= p.rows) break;                                  // warp-uniform (last unit only)
        const long long row = row0 + gi;
        const bool ract = row < p.rows;
        const long long rowc = ract ? row : row0;
        const int head = (int)(rowc % p.H);
        const long long b = rowc / ((long long)p.H * p.Nq);
        const unsigned vrow = (unsigned)(b * p.S * HDV + head * G + lane);
        mbar_wait(bar, phase);
        phase ^= 1u;

        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

        // one fixed-point scale for the whole row
        float S = 0.f;
        if (ract)
            for (int i = lane; i < p.LP; i += G) S += fabsf(swrow[i]);
        S = gsum<G>(S, gm);
        const bool finite = S <= 3.0e38f;                           // false for inf / NaN: float path, so they propagate
        const int ke = fixed_scale_exp(fmaxf(S, 1e-30f));
        const float scale = pow2f(ke), inv_scale = pow2f(-ke);
        const bool work = ract && S != 0.f;

        const int npass = (p.L + LPP - 1) / LPP;
        for (int pass0 = 0; pass0 < npass; pass0 += kStgMaxPasses) {
            const int pass1 = min(npass, pass0 + kStgMaxPasses);
            int pb = 0;                                                // slots handed out so far
            // ---- A + B: every level of this batch into the pool
            for (int ps = pass0; ps < pass1; ++ps) {
                const int lm = ps * LPP + sub;
                const bool lact = work && lm < p.L;
                const int lmc = lm < p.L ? lm : 0;
                const int mh = lv.h[lmc], mw = lv.w[lmc];
                LevelBox lbx;
                if (MODE == 1) lbx = stg_level_box(p, srow, rowc, lmc, b);
                LanePoint pt[PPL];
                int bx0 = kNoPix, bx1 = -kNoPix, by0 = kNoPix, by1 = -kNoPix;
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    const int ptn = lact ? slane + k * SUB : p.P;
                    pt[k] = stg_point<MODE>(p, srow, swrow, lbx, lmc, ptn, mh, mw);
                    if (pt[k].inside) {
                        bx0 = min(bx0, pt[k].x0); bx1 = max(bx1, pt[k].x0 + 1);
                        by0 = min(by0, pt[k].y0); by1 = max(by1, pt[k].y0 + 1);
                    }
                }
                PoolWin me;
                me.X0 = max(smin<SUB>(bx0, gm), 0); me.Y0 = max(smin<SUB>(by0, gm), 0);
                me.nx = min(smax<SUB>(bx1, gm), mw - 1) - me.X0 + 1;
                me.ny = min(smax<SUB>(by1, gm), mh - 1) - me.Y0 + 1;
                const bool empty = me.nx <= 0 || me.ny <= 0;
                const int nq4 = empty ? 0 : ((me.nx * me.ny + 3) & ~3);
                // hand out pool slots level by level (sub-groups of a pass in order)
                me.base = pb;
                me.mode = 0;
#pragma unroll
                for (int sl = 0; sl < LPP; ++sl) {
                    int take = 0;
                    if (sub == sl) {
                        me.base = pb;
                        me.mode = empty ? 0 : ((finite && nq4 <= kPool - pb) ? 1 : 2);
                        take = me.mode == 1 ? nq4 : 0;
                    }
                    if (LPP > 1) take = __shfl_sync(gm, take, sl * SUB, G);
                    pb += take;
                }
                if (me.mode == 1) {
                    for (int s = slane * 4; s < nq4; s += SUB * 4) {
                        *reinterpret_cast<int4*>(pw + me.base + s) = make_int4(0, 0, 0, 0);
                        *reinterpret_cast<uint4*>(po + me.base + s) = make_uint4(kNoPixel, kNoPixel, kNoPixel, kNoPixel);
                    }
                }
                __syncwarp(gm);
                if (me.mode == 1) {
                    const unsigned lev = (unsigned)lv.start[lmc] * HDV;
#pragma unroll
                    for (int k = 0; k < PPL; ++k) {
                        if (pt[k].inside) {
                            const int sx = pt[k].x0 - me.X0, sy = pt[k].y0 - me.Y0;   // -1 .. n-1
                            const float hx = 1.f - pt[k].lx, hy = 1.f - pt[k].ly;
                            const float a = pt[k].aw * scale;
                            const bool vx0 = sx >= 0, vx1 = sx + 1 < me.nx, vy0 = sy >= 0, vy1 = sy + 1 < me.ny;
                            const int s00 = me.base + sy * me.nx + sx;
                            const unsigned o00 = lev + (unsigned)(pt[k].y0 * mw + pt[k].x0) * HDV;   // wraps for -1; valid corners are right
                            const unsigned orow = (unsigned)mw * HDV;
                            if (vy0 && vx0) { atomicAdd(pw + s00, __float2int_rn(hy * hx * a)); po[s00] = o00; }
                            if (vy0 && vx1) { atomicAdd(pw + s00 + 1, __float2int_rn(hy * pt[k].lx * a)); po[s00 + 1] = o00 + HDV; }
                            if (vy1 && vx0) { atomicAdd(pw + s00 + me.nx, __float2int_rn(pt[k].ly * hx * a)); po[s00 + me.nx] = o00 + orow; }
                            if (vy1 && vx1) { atomicAdd(pw + s00 + me.nx + 1, __float2int_rn(pt[k].ly * pt[k].lx * a)); po[s00 + me.nx + 1] = o00 + orow + HDV; }
                        }
                    }
                }
                // levels that do not fit the pool (or non-finite weights): per-point gather, taps broadcast by their owners
#pragma unroll
                for (int sl = 0; sl < LPP; ++sl) {
                    const int md = (LPP == 1) ? me.mode : __shfl_sync(gm, me.mode, sl * SUB, G);
                    if (md != 2) continue;
                    const int l = ps * LPP + sl;
                    const int lh = lv.h[l], lw = lv.w[l];
                    const unsigned vlev = vrow + (unsigned)lv.start[l] * HDV;
#pragma unroll
                    for (int k = 0; k < PPL; ++k) {
#pragma unroll kFbUnroll
                        for (int o = 0; o < SUB; ++o) {
                            if (o + k * SUB >= p.P) break;            // uniform in the group
                            const int src = sl * SUB + o;
                            const bool inside = __shfl_sync(gm, (int)pt[k].inside, src, G) != 0;
                            const int x0 = __shfl_sync(gm, pt[k].x0, src, G), y0 = __shfl_sync(gm, pt[k].y0, src, G);
                            const float lx = __shfl_sync(gm, pt[k].lx, src, G), ly = __shfl_sync(gm, pt[k].ly, src, G);
                            const float aw = __shfl_sync(gm, pt[k].aw, src, G);
                            if (!inside) continue;
                            const float hx = 1.f - lx, hy = 1.f - ly;
                            const bool vx0 = x0 >= 0, vx1 = x0 + 1 <= lw - 1, vy0 = y0 >= 0, vy1 = y0 + 1 <= lh - 1;
                            const bool ok[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
                            const float cw[4] = {hy * hx * aw, hy * lx * aw, ly * hx * aw, ly * lx * aw};
                            const unsigned c00 = vlev + (unsigned)(y0 * lw + x0) * HDV;
                            float v[4][VEC];
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if (ok[c]) V::load16(value16, c00 + ((c & 1) ? HDV : 0u) + ((c & 2) ? (unsigned)lw * HDV : 0u), v[c]);
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if (ok[c]) {
#pragma unroll
                                    for (int i = 0; i < VEC; ++i) acc[i] += cw[c] * v[c][i];
                                }
                        }
                    }
                }
            }
            __syncwarp();          // whole warp: the pools are complete, and (last batch) the stage is no longer read
            if (pass1 == npass) {
                const int un = u + gridDim.x;
                const long long rn = (long long)un * GROUPS + warp * RW;
                if (lane_w == 0 && un < p.units && rn < p.rows) stg_issue<G, MODE>(p, pl, stage, bar, rn);
            }
            // ---- C: one walk over the pool -- a 16-byte row load and VEC FMAs per touched pixel
#pragma unroll 2
            for (int q = 0; q < pb; q += 4) {
                const int4 wq = *reinterpret_cast<const int4*>(pw + q);
                const uint4 oq = *reinterpret_cast<const uint4*>(po + q);
                const int wi[4] = {wq.x, wq.y, wq.z, wq.w};
                const unsigned oi[4] = {oq.x, oq.y, oq.z, oq.w};
                float v[4][VEC];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (wi[j] != 0) V::load16(value16, vrow + oi[j], v[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (wi[j] != 0) {
                        const float wv = (float)wi[j] * inv_scale;
#pragma unroll
                        for (int i = 0; i < VEC; ++i) acc[i] += wv * v[j][i];
                    }
                }
            }
            __syncwarp(gm);        // the pool is rewritten by the next batch / row
        }
        if (ract) V::store(static_cast<TV*>(p.out) + (row * p.D + lane * VEC), acc);
    }
}

}  // namespace bxr
