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
//   B. if the range fits a 64-pixel window, the lanes scatter  attn * bilinear weight  of their
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

// CTA size and resident CTAs per SM the window kernels are compiled for (A/B measured on B200, profiles/README.md).
// What matters is warps per SM against the register budget: 65536 / (warps * 32) registers per thread.
//   forward : 128-thread CTAs, 7 per SM = 28 warps at <= 72 registers (one 4-byte spill).  r01u: 0.1786 -> 0.1757 ms,
//             uniform locations 0.3627 -> 0.3465 ms against 256 x 3 = 24 warps at 80 registers; 32 warps (64
//             registers) spill and lose (0.198 ms).
//   backward: 256-thread CTAs, 4 per SM = 32 warps at 64 registers; 128-thread CTAs cost 3 % there (0.346 -> 0.356 ms),
//             36 warps (56 registers) spill.  Since the table walk the fp32 head_dim-32 one-level-per-pass kernels run 3 per
//             SM instead (bwd_min_blocks below).
// The 16-byte-lane bf16 instantiations (8 channels per lane) keep the equivalent of two 256-thread CTAs.
#ifndef BXR_FWD_THREADS
#define BXR_FWD_THREADS 128
#endif
#ifndef BXR_FWD_MINB
#define BXR_FWD_MINB 7
#endif
#ifndef BXR_BWD_THREADS
#define BXR_BWD_THREADS 256
#endif
#ifndef BXR_BWD_MINB
#define BXR_BWD_MINB 4
#endif
constexpr int kFwdThreads = BXR_FWD_THREADS, kFwdMinB = BXR_FWD_MINB;
// the 72-register budget holds for the fp32 location-taking kernels with up to 2 points per lane (ptxas: 4 bytes of
// spill); the fused-grid / softmax variants, 4 points per lane and the bf16 kernels (unpacking registers; r01x:
// bf16 forward 3-6 % slower at 72) keep the 80-register budget (24 warps)
// (r02l: 6 instead of 7 resident CTAs for the two-levels-per-pass kernels of 2 x 2 grids measured no faster on a second
// box -- 0.0885 vs 0.0874 ms -- and the hook was removed)
constexpr int fwd_min_blocks(int vec, int ppl, bool fused, bool fp32, bool /*two_levels*/) {
    return vec > 4 ? 2 * (kThreads / kFwdThreads)
                   : ((fused || ppl > 2 || !fp32) ? 3 * (kThreads / kFwdThreads)
                                                  : kFwdMinB);
}
constexpr int kBwdThreads = BXR_BWD_THREADS, kBwdMinB = BXR_BWD_MINB;
// The fp32 one-level-per-pass backward of head_dim 32 with the atomic scatter runs 3 CTAs per SM (24 warps, 78 registers, no spill) since
// the table walk: r02bw, 4 -> 3 CTAs: box 0.3395 -> 0.3301 ms, trained-like 0.558 -> 0.531, uniform 1.108 -> 1.047.  The
// two-levels-per-pass kernels of 2 x 2 grids (uniform +9 %) and bf16 (+2 %) lose with it; the fused-grid variants gain the
// same (r02bf: 0.378 -> 0.356 ms, with the softmax 0.401 -> 0.378); the deterministic ones do not care (2.06 -> 2.05) and
// other head dims were not measured: both keep 4.
#ifndef BXR_BWD_MINB_F32_LPP1
#define BXR_BWD_MINB_F32_LPP1 3
#endif
constexpr int bwd_min_blocks(int vec, bool fp32, bool one_level, bool det) {
    return vec > 4 ? 2 * (kThreads / kBwdThreads)
                   : ((fp32 && one_level && !det) ? BXR_BWD_MINB_F32_LPP1 : kBwdMinB);
}
// Forward window walk from a slot table: after the scatter, the lanes of the level turn the dense window into
// (pixel offset, float weight) pairs, one lane per slot, and the walk reads two pairs per 16-byte shared load --
// the per-slot index arithmetic (row wrap, int -> float, scaling: 9 of the 16 instructions a slot costs) is done
// once per slot instead of once per slot and lane.  Measured (r02a, K=4 encoder forward): bf16 0.1988 -> 0.1858 ms,
// fp32 at 28 warps / 72 registers 0.1786 -> 0.1865 ms at the time (20 bytes of spill).  Re-measured in round 2 (r02x), after
// the corner table had put the table walk into the fp32 kernels anyway and the in-register walk's row-wrap arithmetic had
// been identified as ~10 % of the forward by ablation: fp32 K=4 0.1728 -> 0.1548 ms, trained-like 0.303 -> 0.247, K=2
// 0.088 -> 0.076 -- adopted.  BXR_FWD_TAB: 0 off, 1 all types (default), 2 all but fp32.
#ifndef BXR_FWD_TAB
#define BXR_FWD_TAB 1
#endif
template <typename TV>
struct FwdSlotTable { static constexpr bool value = BXR_FWD_TAB == 1 || (BXR_FWD_TAB == 2 && !std::is_same<TV, float>::value); };
// BXR_FWD_CTAB: forward, wide footprints (the per-point mode).  Instead of broadcasting every point's tap from its owner
// lane (6 shuffles + the corner arithmetic on all G lanes, per point), each lane writes the four corners of its own points
// as (offset, weight) entries of the slot table and the group walks the 4 P entries exactly like a window's slots.
// 0 off, 1 on for P <= 16 (the table holds 64 entries per group).
#ifndef BXR_FWD_CTAB
#define BXR_FWD_CTAB 1
#endif
// BXR_BWD_TAB: the backward walk from a slot table like the forward's.  0 off, 1 all types, 2 (default) see TABB below.
// Measured r02l (K=4 encoder backward, before the corner table): bf16 0.4118 -> 0.3946 ms, fp32 0.3526 -> 0.3767 ms;
// r02x (with it): fp32 0.3452 -> 0.3418 ms, trained-like 0.636 -> 0.557.  One lane per slot writes (offset relative to the
// lane's base pointers, float weight) after the scatter.  The walk then
// needs no row-wrap arithmetic, no flag read, no int -> float scaling per lane, and no group barrier before the
// d totals overwrite the flag window (the flags are consumed when the table is built).  An untouched pixel is marked by
// the offset kAbsent (not by its weight: non-finite weights must reach grad_value).
#ifndef BXR_BWD_TAB
#define BXR_BWD_TAB 2
#endif
// BXR_BASE_REGPAIR: gather through a per-lane base pointer kept as an opaque register pair, offsets relative to it -- one
// IMAD.WIDE per load; otherwise the compiler re-loads the tensor base from the constant bank (LDC.64) in front of every
// load (seen in SASS, forward walk).  Measured r02l: forward K=4 0.1779 -> 0.1745 ms, uniform 0.346 -> 0.335, bf16
// 0.186 -> 0.1826; the two-levels-per-pass forward (2 x 2 grids) 0.0874 -> 0.0898 and the backward 0.3526 -> 0.3619 lose.
// 0 off, 1 = forward kernels with one level per pass (default), 2 = everywhere (the r02l A/B build).
#ifndef BXR_BASE_REGPAIR
#define BXR_BASE_REGPAIR 1
#endif
// unroll factors of the per-point fallback loops: the walk is a chain of dependent gathers, unrolling lets the
// compiler request the corner rows of several points before the first is used (A/B r01s: forward 1 -> 8:
// 0.1846 -> 0.1794 ms, uniform 0.382 -> 0.362 ms; backward 1 -> 4: 0.365 -> 0.350 ms, 8 is worse there)
#ifndef BXR_BWD_CTAB
#define BXR_BWD_CTAB 1
#endif
#ifndef BXR_DIAG
#define BXR_DIAG 0
#endif
#ifndef BXR_FB_UNROLL
#define BXR_FB_UNROLL 8
#endif
#ifndef BXR_FB_UNROLL_BWD
#define BXR_FB_UNROLL_BWD 4
#endif

// Work units are dealt round-robin from the LAST one down: in BoxeR's encoder the trailing queries belong to the
// coarse levels, whose rows cost several times more (wide footprints -> per-point walk); starting with them leaves
// the cheap units for the tail of the launch (A/B r01q: forward 0.1875 -> 0.1844 ms; harmless for other inputs).
#ifndef BXR_UNIT_REVERSE
#define BXR_UNIT_REVERSE 1
#endif
// 4-entry batches of the forward table walk in flight (r02un: fp32 K=4 box forward 0.1498 (2) -> 0.1469 (1) -> 0.158 (4) ms,
// K=2 0.0750 -> 0.0739, bf16 0.1579 -> 0.1556; uniform points 0.2938 (2) -> 0.2977 (1) -> 0.2871 (4)).  One loop for both
// table kinds: a separate loop per kind, each with its own unroll factor, cost 18 % (r02uw: 72 registers + spill).
// backward slot table in the deterministic kernels: 0 never (default), 1 like the atomic kernels, 2 bf16 only.  With the
// table the 64-bit reductions of a 4-slot batch leave back to back and the LSU-bound kernel loses 25 % (r02dt: fp32 2.08 ->
// 2.59 ms, bf16 2.12 -> 2.60, trained-like 3.5 -> 4.27)
#ifndef BXR_BWD_TAB_DET
#define BXR_BWD_TAB_DET 0
#endif
#ifndef BXR_FWD_TAB_UNROLL
#define BXR_FWD_TAB_UNROLL 1
#endif
constexpr int kFwdTabUnroll = BXR_FWD_TAB_UNROLL;
// Software-pipelined point loads in the fp32 one-level-per-pass forward: the next level's (at the last level: the next
// work unit's) locations / weights are requested before the current level is walked.  The first use of a level's points
// was the longest single stall of the kernel (ncu r02z: 12 % of the stall samples on 0.5 % of the instructions); r02pf:
// K=4 forward 0.1558 -> 0.1508 ms, trained-like 0.248 -> 0.2375, uniform 0.301 -> 0.295.  The backward (+4 %: it is bound
// by the L2 reduction rate, not by this latency), the two-levels-per-pass kernels (+4 %) and bf16 (+5 %) lose and keep the
// plain loads (the backward again at its final 3 CTAs per SM / 78 registers: 0.3301 -> 0.3375 ms); prefetch.global.L2 / .L1 of the
// next unit's rows instead: forward -1.6 %, backward +2.6 %, not adopted.
#ifndef BXR_FWD_PIPE
#define BXR_FWD_PIPE 1
#endif
constexpr int kFbUnroll = BXR_FB_UNROLL;
constexpr int kFbUnrollBwd = BXR_FB_UNROLL_BWD;
// table entry of a pixel nobody touched / a corner outside the level (valid offsets stay below it: use_window())
constexpr unsigned kAbsent = 0xffffffffu;
// "no pixel touched yet" sentinel of the range reductions: with +-kNoPix in both ends the extent
// max - min + 1 of an empty range is a large negative number that still fits an int (+-INT_MAX wrapped
// around to 3, which sent rows without any in-range point through a 3 x 3 window of zeros)
constexpr int kNoPix = 0x3fffffff;
constexpr int kWinSide = 8;
constexpr int kWinSlots = kWinSide * kWinSide;
// per-group pitch of a window in 32-bit words: 64 slots + 4 words of skew, so that the 16-byte
// row reads of the 4 (G=8) or 8 (G=4) groups of a warp fall into different banks
constexpr int kWinPitch = kWinSlots + 4;
// slot table: 8-byte entries, pitch 64 + 2 entries (528 bytes: a multiple of 16, and the groups of a warp are 4 banks apart)
constexpr int kTabPitch = kWinSlots + 2;

// Pixel weights are accumulated in the shared-memory window as 32-bit fixed point with a per
// (row, level) power-of-two scale: integer atomics are single instructions (a float atomicAdd on
// shared memory is a compare-and-swap loop on sm_100) and the sum is order independent.
// With S = sum_p |attn_p| <= 2^e every pixel weight is bounded by S, so scale = 2^(30-e) cannot
// overflow; the rounding step is 2^-31 of S, below fp32 resolution of the largest weight.
__device__ __forceinline__ int fixed_scale_exp(float S) {
    const int e = ((__float_as_int(S) >> 23) & 0xff) - 126;     // S < 2^e  (normal S)
    return max(-126, min(126, 30 - e));
}
__device__ __forceinline__ float pow2f(int k) { return __int_as_float((127 + k) << 23); }

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

__device__ __forceinline__ LanePoint lane_point_xy(float lx_n, float ly_n, float aw, bool act, int h, int w) {
    LanePoint t;
    t.aw = aw;
    const float x = lx_n * (float)w - 0.5f;
    const float y = ly_n * (float)h - 0.5f;
    t.inside = act && (y > -1.f) && (x > -1.f) && (y < (float)h) && (x < (float)w);
    const float xs = t.inside ? x : 0.f, ys = t.inside ? y : 0.f;
    const float xf = floorf(xs), yf = floorf(ys);
    t.x0 = (int)xf;
    t.y0 = (int)yf;
    t.lx = xs - xf;
    t.ly = ys - yf;
    return t;
}

__device__ __forceinline__ LanePoint lane_point(const float* __restrict__ loc_l, const float* __restrict__ w_l,
                                                int pt, int P, int h, int w) {
    const bool act = pt < P;
    const int pc = act ? pt : 0;
    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc_l) + pc);
    return lane_point_xy(xy.x, xy.y, __ldg(w_l + pc), act, h, w);
}

// the loads of lane_point alone, for the software-pipelined forward (BXR_FWD_PIPE): the next level's (at the last
// level: the next work unit's) points are requested before the current level is walked
struct RawPoint {
    float2 xy;
    float aw;
};
__device__ __forceinline__ RawPoint raw_point(const float* __restrict__ loc_l, const float* __restrict__ w_l, int pt, int P) {
    const int pc = pt < P ? pt : 0;
    RawPoint r;
    r.xy = __ldg(reinterpret_cast<const float2*>(loc_l) + pc);
    r.aw = __ldg(w_l + pc);
    return r;
}

// One (row, level) box of the fused entry points: the K x K grid of BoxAttention._where_to_attend
// (box_attention.py:196-214) / Box3dAttention (:304-338) generated in registers:
//   loc_p = (centre + R(angle) (kernel_index_p * relu(size))) * valid_ratio
struct LevelBox {
    float cx, cy, sx, sy;     // sx, sy = relu(w), relu(h)
    float cs, sn;             // cos / sin of the angle (1, 0 without rotation)
    float vx, vy;             // valid ratios (1, 1 without)
    bool pos_w, pos_h;        // w > 0, h > 0 (relu gradient)
};

__device__ __forceinline__ LevelBox load_level_box(const AttnParams& p, long long rl, long long b, int l) {
    LevelBox q;
    const float4 bx = __ldg(reinterpret_cast<const float4*>(p.boxes) + rl);
    q.cx = bx.x; q.cy = bx.y;
    q.pos_w = bx.z > 0.f; q.pos_h = bx.w > 0.f;
    q.sx = q.pos_w ? bx.z : 0.f; q.sy = q.pos_h ? bx.w : 0.f;
    q.cs = 1.f; q.sn = 0.f;
    if (p.angles) {
        const float a = __ldg(static_cast<const float*>(p.angles) + rl);
        sincosf(a, &q.sn, &q.cs);
    }
    q.vx = q.vy = 1.f;
    if (p.valid_ratios) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p.valid_ratios) + (b * p.L + l));
        q.vx = v.x; q.vy = v.y;
    }
    return q;
}

__device__ __forceinline__ void box_point(const LevelBox& q, float kx, float ky, float& x, float& y) {
    const float ux = kx * q.sx, uy = ky * q.sy;
    x = (q.cx + (ux * q.cs - uy * q.sn)) * q.vx;
    y = (q.cy + (ux * q.sn + uy * q.cs)) * q.vy;
}

// softmax over a row's L*P logits, fused into the op (box_attention.py:227-231):
// weight = exp(logit - row max) / row sum
struct RowSoftmax {
    float mx, inv;
};
template <int G>
__device__ __forceinline__ RowSoftmax row_softmax(const float* __restrict__ logits, int LP, int lane, unsigned gm) {
    float mx = -INFINITY;
    for (int i = lane; i < LP; i += G) mx = fmaxf(mx, __ldg(logits + i));
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(gm, mx, o));
    float sum = 0.f;
    for (int i = lane; i < LP; i += G) sum += __expf(__ldg(logits + i) - mx);     // ex2.approx: 2 ulp, inside the 1e-4 bar
    sum = gsum<G>(sum, gm);
    RowSoftmax r;
    r.mx = mx;
    r.inv = 1.f / sum;
    return r;
}

template <bool SMAX>
__device__ __forceinline__ LanePoint lane_point_box(const LevelBox& q, const float* __restrict__ kidx,
                                                    const float* __restrict__ w_l, int pt, int P, int h, int w,
                                                    const RowSoftmax& sm) {
    const bool act = pt < P;
    const int pc = act ? pt : 0;
    const float2 k = __ldg(reinterpret_cast<const float2*>(kidx) + pc);
    float x, y;
    box_point(q, k.x, k.y, x, y);
    float aw = __ldg(w_l + pc);
    if (SMAX) aw = __expf(aw - sm.mx) * sm.inv;
    return lane_point_xy(x, y, aw, act, h, w);
}

// ------------------------------------------------------------------------------------------------
// Lane geometry shared by the forward and backward kernels.
//   G    lanes per row (channel slices of 16 bytes)
//   SUB  lanes that share one level's points (SUB == G: one level per pass; SUB == G/2: two levels
//        per pass, used for 2x2 grids where P = 4 would leave half of the lanes idle)
//   PPL  points per lane per level: P <= SUB * PPL
template <int G, int SUB>
struct WinGeom {
    static constexpr int LPP = G / SUB;                  // levels per pass
    static constexpr int CAP = kWinSlots / LPP;          // window slots per level of a pass
};

template <int SUB>
__device__ __forceinline__ int smin(int v, unsigned m) {
#pragma unroll
    for (int o = SUB / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(m, v, o));
    return v;
}
template <int SUB>
__device__ __forceinline__ int smax(int v, unsigned m) {
#pragma unroll
    for (int o = SUB / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o));
    return v;
}
template <int SUB>
__device__ __forceinline__ float ssum(float v, unsigned m) {
#pragma unroll
    for (int o = SUB / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o);
    return v;
}

// All groups of a warp walk the level loop together (rows past the end are carried along as inactive), so the
// reductions and barriers that sit outside mode-dependent branches use the full warp mask: a partial-mask
// shuffle / __syncwarp compiles to a MATCH.ANY + REDUX + VOTE + BRA.DIV convergence check in front of it.
constexpr unsigned kFullMask = 0xffffffffu;

// zero the first n4 (multiple of 4) words of a CAP-word window: SUB lanes, 16-byte stores, no loop
template <int SUB, int CAP, typename T>
__device__ __forceinline__ void zero_window(T* win, int n4, int slane) {
    static_assert(sizeof(T) == 4, "window slots are 32-bit");
#pragma unroll
    for (int j = 0; j < (CAP + SUB * 4 - 1) / (SUB * 4); ++j) {
        const int s = (j * SUB + slane) * 4;
        if (s < n4) *reinterpret_cast<uint4*>(win + s) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ------------------------------------------------------------------------------------------------
// Tile-ordered work units (self-attention-shaped calls, Nq == S: the queries are the pixels of the value pyramid and
// every query's box sits on its own pixel, box_transformer.py:70-116).  With rows dealt in memory order a work unit is
// "2 (4) consecutive queries x all heads": the heads share nothing (each gathers its own 128-byte slice of a pixel) and
// the queries of the units a CTA sees next are far away, so the ~80 % overlap between the footprints of neighbouring
// queries never meets in one L1 (ncu r01x: sector hit rate 29 %).  Here a unit is a TW x TH tile of queries of one level
// for ONE head -- a warp takes a tile line -- so that the rows a CTA works on at the same time gather the same pixels
// (the tile's footprint: 131 pixels for a 4 x 4 tile against 720 window slots) and L1 serves them.  The order of the units
// is a schedule, not an assumption: any row is still processed by the general algorithm.
template <int TW, int TH>
struct TileOrder {
    int before[kMaxLevels + 1];      // tiles of the levels processed before level order k (coarse levels first)
};

template <int TW, int TH>
__device__ __forceinline__ void tile_order_init(TileOrder<TW, TH>& t, const LevelTable& lv, int L) {
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < L; ++k) {
            const int l = L - 1 - k;
            t.before[k] = acc;
            acc += ((lv.w[l] + TW - 1) / TW) * ((lv.h[l] + TH - 1) / TH);
        }
        t.before[L] = acc;
    }
    __syncthreads();
}

// unit -> this group's row (or -1): unit = ((image * tiles + tile) * H + head), group gid = (line, column) of the tile
template <int TW, int TH>
__device__ __forceinline__ long long tile_order_row(const TileOrder<TW, TH>& t, const LevelTable& lv, const AttnParams& p, unsigned u, int gid) {
    const unsigned H = (unsigned)p.H;
    const unsigned v = u / H;
    const int head = (int)(u - v * H);
    const unsigned tiles = (unsigned)t.before[p.L];
    const unsigned b = v / tiles;
    const int ti = (int)(v - b * tiles);
    int k = 0;
    while (k + 1 < p.L && t.before[k + 1] <= ti) ++k;
    const int l = p.L - 1 - k;
    const int tl = ti - t.before[k];
    const int ntx = (lv.w[l] + TW - 1) / TW;
    const int ty = tl / ntx, tx = tl - ty * ntx;
    const int x = tx * TW + gid % TW, y = ty * TH + gid / TW;
    if (x >= lv.w[l] || y >= lv.h[l]) return -1;
    const long long q = lv.start[l] + (long long)y * lv.w[l] + x;
    if (q >= p.Nq) return -1;
    return ((long long)b * p.Nq + q) * p.H + head;
}

// what every lane of the group needs to know about one level of the current pass
struct SubWin {
    int X0, Y0, nx, ny;   // touched pixel range (clamped to the level); nx <= 0: nothing inside
    int ke;               // fixed-point scale exponent
    int mode;             // 0 skip, 1 window, 2 per-point fallback
};

// ------------------------------------------------------------------------------------------------
// Forward.  One group of G lanes per row; work units (256/G rows) dealt round-robin to the CTAs.
// SMAX (with FUSED): `w0` holds logits; the softmax over the row's L*P points is taken here and written to attn_out.
template <typename TV, int G, int SUB, int PPL, bool FUSED, bool SMAX = false, bool TILED = false>
__global__ void __launch_bounds__(kFwdThreads, fwd_min_blocks(Vec16<TV>::VEC, PPL, FUSED, std::is_same<TV, float>::value, SUB < G)) box_fwd_win_kernel(const AttnParams p) {
    static_assert(FUSED || !SMAX, "the softmax prologue is built for the fused entry points only");
    using V = Vec16<TV>;
    using GEO = WinGeom<G, SUB>;
    constexpr int VEC = V::VEC;
    constexpr int GROUPS = kFwdThreads / G;
    constexpr int LPP = GEO::LPP, CAP = GEO::CAP;
    __shared__ LevelTable lv;
    __shared__ __align__(16) int s_win[GROUPS * kWinPitch];
    constexpr bool TAB = FwdSlotTable<TV>::value;
    // corner entries of one level (4 per point) must fit the level's share of the table: decided by the geometry alone, so
    // that an instantiation carries either the table walk or the shuffle walk for wide footprints, never both
    constexpr bool CTAB = BXR_FWD_CTAB != 0 && 4 * SUB * PPL <= CAP;
    __shared__ __align__(16) uint2 s_tab[(TAB || CTAB) ? GROUPS * kTabPitch : 2];
    load_levels(lv, p);
    constexpr bool ctab = CTAB;
    // tile-ordered units: a warp (32 / G rows) is a tile line, the CTA a tile of (32 / G) x (warps) queries of one head
    constexpr int TW = 32 / G, TH = kFwdThreads / 32;
    __shared__ TileOrder<TW, TH> s_order;
    int n_units = p.units;
    if constexpr (TILED) {
        tile_order_init(s_order, lv, p.L);
        n_units = (int)((unsigned)p.B * (unsigned)s_order.before[p.L] * (unsigned)p.H);
    }

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const int sub = lane / SUB, slane = lane % SUB;          // my level slot of a pass, my lane within it
    const unsigned gm = group_mask<G>();
    int* gwin = s_win + gid * kWinPitch;
    int* win = gwin + sub * CAP;
    uint2* gtab = s_tab + ((TAB || CTAB) ? gid * kTabPitch : 0);
    uint2* tab = gtab + ((TAB || CTAB) ? sub * CAP : 0);
    const unsigned HDV = (unsigned)(p.H * p.D) / VEC;      // pixel pitch in lane chunks (VEC elements)
    const void* __restrict__ value16 = p.value;   // indexed in lane-chunk units by V::load16
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);

    constexpr bool PIPE = BXR_FWD_PIPE != 0 && !FUSED && !TILED && LPP == 1 && std::is_same<TV, float>::value;
    RawPoint nxt[PPL];
    if constexpr (PIPE) {
        const long long r0 = (long long)(BXR_UNIT_REVERSE ? n_units - 1 - (int)blockIdx.x : (int)blockIdx.x) * GROUPS + gid;
        const long long r0c = (r0 >= 0 && r0 < p.rows) ? r0 : 0;
        const int lc0 = sub < p.L ? sub : 0;
#pragma unroll
        for (int k = 0; k < PPL; ++k) nxt[k] = raw_point(loc + r0c * p.LP * 2 + lc0 * p.P * 2, w0 + r0c * p.LP + lc0 * p.P, slane + k * SUB, p.P);
    }
    // units are dealt round-robin: rows of the coarse levels (wide windows -> per-point fallback) cost
    // several times more than level-0 rows, and a contiguous split leaves the CTAs that own them as a tail
#if BXR_UNIT_REVERSE
    for (int u = (TILED ? (int)blockIdx.x : n_units - 1 - (int)blockIdx.x); TILED ? (u < n_units) : (u >= 0); u += TILED ? (int)gridDim.x : -(int)gridDim.x) {
#else
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
#endif
        // (the tile order lists the coarse levels first by construction: no reversal needed there)
        const long long row_raw = TILED ? tile_order_row(s_order, lv, p, (unsigned)u, gid) : (long long)u * GROUPS + gid;
        const bool ract = TILED ? (row_raw >= 0) : (row_raw < p.rows);     // groups without a row stay with their warp, doing nothing
        const long long row = ract ? row_raw : 0;
        const int head = (int)(row % p.H);
        const long long b = row / ((long long)p.H * p.Nq);
        // value addressing in 16-byte units (one lane chunk) as 32-bit indices off the tensor base
        const unsigned vrow = (unsigned)(b * p.S * HDV + head * G + lane);
        const float* loc_row = FUSED ? nullptr : loc + row * p.LP * 2;
        const float* w_row = w0 + row * p.LP;
        RowSoftmax rsm;
        rsm.mx = 0.f; rsm.inv = 1.f;
        if constexpr (SMAX) rsm = row_softmax<G>(w_row, p.LP, lane, kFullMask);

        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

        for (int l0 = 0; l0 < p.L; l0 += LPP) {
            // ---- A: own points of my level, touched pixel range (reductions stay inside the SUB lanes)
            const int lm = l0 + sub;
            const bool lact = ract && lm < p.L;
            const int lmc = lm < p.L ? lm : 0;
            const int mh = lv.h[lmc], mw = lv.w[lmc];
            LanePoint pt[PPL];
            int bx0 = kNoPix, bx1 = -kNoPix, by0 = kNoPix, by1 = -kNoPix;
            float S = 0.f;
            LevelBox lbx;
            if constexpr (FUSED) lbx = load_level_box(p, row * p.L + lmc, b, lmc);
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                const int ptn = lact ? slane + k * SUB : p.P;
                if constexpr (FUSED) pt[k] = lane_point_box<SMAX>(lbx, static_cast<const float*>(p.kidx), w_row + lmc * p.P, ptn, p.P, mh, mw, rsm);
                else if constexpr (PIPE) pt[k] = lane_point_xy(nxt[k].xy.x, nxt[k].xy.y, nxt[k].aw, ptn < p.P, mh, mw);
                else pt[k] = lane_point(loc_row + lmc * p.P * 2, w_row + lmc * p.P, ptn, p.P, mh, mw);
                if constexpr (SMAX) {
                    if (ptn < p.P) static_cast<float*>(p.attn_out)[row * p.LP + lmc * p.P + ptn] = pt[k].aw;
                }
                if (pt[k].inside) {
                    bx0 = min(bx0, pt[k].x0); bx1 = max(bx1, pt[k].x0 + 1);
                    by0 = min(by0, pt[k].y0); by1 = max(by1, pt[k].y0 + 1);
                    S += fabsf(pt[k].aw);
                }
            }
            if constexpr (PIPE) {
                // request the next pass's points now: the next level of this row, or level 0 of the next unit's row
                const bool last = l0 + LPP >= p.L;
                const long long nrow = row_raw + (BXR_UNIT_REVERSE ? -1LL : 1LL) * (long long)gridDim.x * GROUPS;
                const long long rn = last ? ((nrow >= 0 && nrow < p.rows) ? nrow : row) : row;
                const int ln = (last ? 0 : l0 + LPP) + sub;
                const int lnc = ln < p.L ? ln : 0;
#pragma unroll
                for (int k = 0; k < PPL; ++k) nxt[k] = raw_point(loc + rn * p.LP * 2 + lnc * p.P * 2, w0 + rn * p.LP + lnc * p.P, slane + k * SUB, p.P);
            }
            SubWin me;
            me.X0 = max(smin<SUB>(bx0, kFullMask), 0); me.Y0 = max(smin<SUB>(by0, kFullMask), 0);
            me.nx = min(smax<SUB>(bx1, kFullMask), mw - 1) - me.X0 + 1;
            me.ny = min(smax<SUB>(by1, kFullMask), mh - 1) - me.Y0 + 1;
            S = ssum<SUB>(S, kFullMask);
            const int nq = (me.nx > 0 && me.ny > 0) ? me.nx * me.ny : 0;
            // non-finite weights (S is NaN/inf) take the float path so that they propagate
            me.mode = (me.nx <= 0 || me.ny <= 0 || S == 0.f) ? 0 : ((nq <= CAP && S <= 3.0e38f) ? 1 : 2);
            me.ke = fixed_scale_exp(S);

            // ---- B: scatter pixel weights into my level's dense nx x ny window (32-bit fixed point)
            if (me.mode == 1) zero_window<SUB, CAP>(win, (nq + 3) & ~3, slane);
            __syncwarp();
            if (me.mode == 1) {
                const float scale = pow2f(me.ke);
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    if (pt[k].inside) {
                        const int sx = pt[k].x0 - me.X0, sy = pt[k].y0 - me.Y0;   // -1 .. n-1
                        const float hx = 1.f - pt[k].lx, hy = 1.f - pt[k].ly;
                        const float ax = pt[k].aw * scale * hx, bx = pt[k].aw * scale * pt[k].lx;
                        // the four corner weights first, so that each guarded body is a single predicated atomic
                        const int w00 = __float2int_rn(hy * ax), w01 = __float2int_rn(hy * bx);
                        const int w10 = __float2int_rn(pt[k].ly * ax), w11 = __float2int_rn(pt[k].ly * bx);
                        const bool vx0 = sx >= 0, vx1 = sx + 1 < me.nx, vy0 = sy >= 0, vy1 = sy + 1 < me.ny;
                        int* wp = win + sy * me.nx + sx;
                        if (vy0 && vx0) atomicAdd(wp, w00);
                        if (vy0 && vx1) atomicAdd(wp + 1, w01);
                        if (vy1 && vx0) atomicAdd(wp + me.nx, w10);
                        if (vy1 && vx1) atomicAdd(wp + me.nx + 1, w11);
                    }
                }
            }
            __syncwarp();
            // ---- B': one lane per slot: dense slot s = (y, x) -> (value offset relative to the row's base, weight)
            if (TAB && me.mode == 1) {
                const float inv_scale = pow2f(-me.ke);
                // s / nx for s < 64, nx <= 64 as a multiply-shift: rcp = floor(65536 / nx) + 1
                const unsigned rcp = (unsigned)__float2int_rz(__fdividef(65536.f, (float)me.nx)) + 1u;
                const unsigned tbase = ((unsigned)lv.start[lmc] + (unsigned)(me.Y0 * mw + me.X0)) * HDV;
                const unsigned trow = (unsigned)mw * HDV;
                const int nq4 = (nq + 3) & ~3;
#pragma unroll 2
                for (int ts = slane; ts < nq4; ts += SUB) {
                    const unsigned y = ((unsigned)ts * rcp) >> 16;
                    const unsigned x = (unsigned)ts - y * (unsigned)me.nx;
                    const float wv = (float)win[ts] * inv_scale;            // slots past nq were zeroed: weight 0
                    tab[ts] = make_uint2(tbase + y * trow + x * HDV, __float_as_uint(wv));
                }
            }
            // ---- B'': wide footprint: my points' corners as table entries (offset from the row's base, attn * bilinear weight)
            if constexpr (CTAB) if (me.mode == 2) {
                const unsigned tlev = (unsigned)lv.start[lmc] * HDV;
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    const int ptn = slane + k * SUB;
                    if (ptn < p.P) {
                        const LanePoint& t = pt[k];
                        const float hx = 1.f - t.lx, hy = 1.f - t.ly;
                        const bool vx0 = t.x0 >= 0, vx1 = t.x0 + 1 <= mw - 1, vy0 = t.y0 >= 0, vy1 = t.y0 + 1 <= mh - 1;
                        const unsigned c00 = tlev + (unsigned)(t.y0 * mw + t.x0) * HDV;      // wraps for -1; such corners get weight 0
                        const float aw = t.inside ? t.aw : 0.f;
                        uint4* e = reinterpret_cast<uint4*>(tab + ptn * 4);
                        e[0] = make_uint4((vy0 && vx0) ? c00 : 0u, __float_as_uint((vy0 && vx0) ? hy * hx * aw : 0.f),
                                          (vy0 && vx1) ? c00 + HDV : 0u, __float_as_uint((vy0 && vx1) ? hy * t.lx * aw : 0.f));
                        e[1] = make_uint4((vy1 && vx0) ? c00 + (unsigned)mw * HDV : 0u, __float_as_uint((vy1 && vx0) ? t.ly * hx * aw : 0.f),
                                          (vy1 && vx1) ? c00 + (unsigned)mw * HDV + HDV : 0u, __float_as_uint((vy1 && vx1) ? t.ly * t.lx * aw : 0.f));
                    }
                }
            }
            if constexpr (TAB || CTAB) __syncwarp();

            // ---- C: all G lanes walk the window(s) of this pass
#pragma unroll
            for (int sl = 0; sl < LPP; ++sl) {
                SubWin w;
                if (LPP == 1) {
                    w = me;
                } else {
                    const int src = sl * SUB;
                    w.X0 = __shfl_sync(kFullMask, me.X0, src, G); w.Y0 = __shfl_sync(kFullMask, me.Y0, src, G);
                    w.nx = __shfl_sync(kFullMask, me.nx, src, G); w.ny = __shfl_sync(kFullMask, me.ny, src, G);
                    w.ke = __shfl_sync(kFullMask, me.ke, src, G); w.mode = __shfl_sync(kFullMask, me.mode, src, G);
                }
#if BXR_DIAG == 7
                if (w.mode >= 0) { acc[0] += (float)w.nx; continue; }      // DIAGNOSTIC: phases A + B only
#endif
                if (w.mode == 0) continue;
                const int l = l0 + sl;
                const int lh = lv.h[l], lw = lv.w[l];
                constexpr bool REGPAIR = BXR_BASE_REGPAIR == 2 || (BXR_BASE_REGPAIR == 1 && LPP == 1);
                const typename V::Raw* vbase = static_cast<const typename V::Raw*>(value16) + (REGPAIR ? vrow : 0u);
                if constexpr (REGPAIR) asm volatile("" : "+l"(vbase));
                const unsigned vlev = (REGPAIR ? 0u : vrow) + (unsigned)lv.start[l] * HDV;
                if ((TAB && w.mode == 1) || (ctab && w.mode == 2)) {
                    // one row load per unique pixel (window) or per corner (wide footprint), four table entries (two
                    // 16-byte shared loads) at a time
                    const uint2* ct = gtab + sl * CAP;
                    const int wq_n = w.mode == 1 ? w.nx * w.ny : 4 * p.P;
                    // per-lane base pointer in registers: the table offsets are the same for all lanes of the group
                    const typename V::Raw* vlane = static_cast<const typename V::Raw*>(value16) + vrow;
                    asm volatile("" : "+l"(vlane));      // keep it a register pair: one IMAD.WIDE per load, no re-derivation
#pragma unroll kFwdTabUnroll
                    for (int q = 0; q < wq_n; q += 4) {
                        const uint4 t0 = *reinterpret_cast<const uint4*>(ct + q);
                        const uint4 t1 = *reinterpret_cast<const uint4*>(ct + q + 2);
                        const unsigned to[4] = {t0.x, t0.z, t1.x, t1.z};
                        const float tw[4] = {__uint_as_float(t0.y), __uint_as_float(t0.w), __uint_as_float(t1.y), __uint_as_float(t1.w)};
                        float v[4][VEC];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#if BXR_DIAG == 11
                            if (tw[j] != 0.f) V::load16(vlane, (unsigned)j, v[j]);      // DIAGNOSTIC (wrong results): gathers hit the same lines
#elif BXR_DIAG == 13
                            if (tw[j] != 0.f) { v[j][0] = v[j][1] = v[j][2] = v[j][3] = __uint_as_float(to[j]); }   // DIAGNOSTIC: no gather
#else
                            if (tw[j] != 0.f) V::load16(vlane, to[j], v[j]);
#endif
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (tw[j] != 0.f) {
#pragma unroll
                                for (int i = 0; i < VEC; ++i) acc[i] += tw[j] * v[j][i];
                            }
                        }
                    }
                } else if (w.mode == 1) {
                    // one row load per unique pixel, four window slots at a time; the pixel index is
                    // advanced incrementally (no division by nx)
                    const int* cw = gwin + sl * CAP;
                    const float inv_scale = pow2f(-w.ke);
                    const int wq_n = w.nx * w.ny;
                    const unsigned row_skip = (unsigned)(lw - w.nx) * HDV;
                    unsigned off = vlev + (unsigned)(w.Y0 * lw + w.X0) * HDV;
                    int ix = 0;
#pragma unroll 2
                    for (int q = 0; q < wq_n; q += 4) {
                        const int4 wq = *reinterpret_cast<const int4*>(cw + q);
                        const int wi[4] = {wq.x, wq.y, wq.z, wq.w};
                        unsigned offs[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
#if BXR_DIAG == 4
                            offs[j] = off + (unsigned)(q + j) * HDV;      // DIAGNOSTIC (wrong pixels): no row wrap
#else
                            offs[j] = off;
                            off += HDV;
                            if (++ix == w.nx) { ix = 0; off += row_skip; }
#endif
                        }
                        float v[4][VEC];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#if BXR_DIAG == 1
                            if (wi[j] != 0) V::load16(vbase, vlev + (unsigned)j, v[j]);   // DIAGNOSTIC (wrong results): every gather hits the same lines
#elif BXR_DIAG == 3
                            if (wi[j] != 0) { v[j][0] = v[j][1] = v[j][2] = v[j][3] = __uint_as_float(offs[j]); }   // DIAGNOSTIC: no gather at all
#else
                            if (wi[j] != 0) V::load16(vbase, offs[j], v[j]);   // slots past nq were zeroed and never written
#endif
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (wi[j] != 0) {
#if BXR_DIAG == 5
                                const float wv = __int_as_float(wi[j]) * inv_scale;      // DIAGNOSTIC (wrong values): no int -> float conversion
#else
                                const float wv = (float)wi[j] * inv_scale;
#endif
#pragma unroll
                                for (int i = 0; i < VEC; ++i) acc[i] += wv * v[j][i];
                            }
                        }
                    }
                } else if constexpr (!CTAB) {
                    // per-point fallback: the owner lane broadcasts its tap
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
                            const unsigned c00 = vlev + (unsigned)(y0 * lw + x0) * HDV;   // wraps for x0/y0 = -1; valid corners are right
                            float v[4][VEC];
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if (ok[c]) V::load16(vbase, c00 + ((c & 1) ? HDV : 0u) + ((c & 2) ? (unsigned)lw * HDV : 0u), v[c]);
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
            __syncwarp();   // the windows are re-zeroed by the next pass
        }
        if (ract) V::store(static_cast<TV*>(p.out) + (row * p.D + lane * VEC), acc);
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

// Sum 4 per-lane partials over the G lanes of a group with a transpose reduction (4 shuffles for
// G = 8 instead of 12): afterwards every lane holds the full sum of ONE of the four values;
// returns its index.  Lanes that hold the same index hold the same total.
template <int G>
__device__ __forceinline__ int reduce4(float (&d)[4], float& total, int lane, unsigned gm) {
    static_assert(G == 4 || G == 8 || G == 16, "window kernels are built for G in {4, 8, 16}");
    if (G == 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] += __shfl_xor_sync(gm, d[i], 8);
    }
    constexpr int HI = (G == 4) ? 2 : 4;      // lane bit that picks the pair {0,1} or {2,3}
    constexpr int LO = HI / 2;                // lane bit that picks within the pair
    const bool hi = (lane & HI) != 0;
    float a = hi ? d[2] : d[0], b = hi ? d[3] : d[1];
    const float sa = hi ? d[0] : d[2], sb = hi ? d[1] : d[3];
    a += __shfl_xor_sync(gm, sa, HI);
    b += __shfl_xor_sync(gm, sb, HI);
    const bool lo = (lane & LO) != 0;
    float r = lo ? b : a;
    const float sr = lo ? a : b;
    r += __shfl_xor_sync(gm, sr, LO);
    if (G >= 8) r += __shfl_xor_sync(gm, r, 1);
    total = r;
    return (hi ? 2 : 0) + (lo ? 1 : 0);
}

// SMAX (with FUSED): `w0` holds the softmax weights the forward wrote; the weight gradients are chained through
// the softmax before they leave the kernel:  grad_logit = w * (grad_w - sum_row(w * grad_w)).
template <typename TV, int G, int SUB, int PPL, typename ACC, bool FUSED, bool SMAX = false, bool TILED = false>
__global__ void __launch_bounds__(kBwdThreads, bwd_min_blocks(Vec16<TV>::VEC, std::is_same<TV, float>::value, G == 8 && SUB == G, sizeof(ACC) == 8)) box_bwd_win_kernel(const AttnParams p) {
    static_assert(FUSED || !SMAX, "the softmax epilogue is built for the fused entry points only");
    using V = Vec16<TV>;
    using GEO = WinGeom<G, SUB>;
    constexpr int VEC = V::VEC;
    constexpr int GROUPS = kBwdThreads / G;
    constexpr int LPP = GEO::LPP, CAP = GEO::CAP;
    constexpr bool DET = sizeof(ACC) == 8;
    __shared__ LevelTable lv;
    __shared__ __align__(16) int s_win[GROUPS * kWinPitch];     // pixel weights W[pix], fixed point
    __shared__ __align__(16) float s_dot[GROUPS * kWinPitch];   // "touched" flag, then d[pix] = <grad_out, value[pix]>
    // 2 (default): every type except the fp32 two-levels-per-pass kernels of 2 x 2 grids (r02x: fp32 K=4 0.3452 -> 0.3418 ms,
    // trained-like 0.636 -> 0.557; K=2 0.1844 -> 0.1923 loses)
    // The deterministic scatter (64-bit integer reductions, LSU-bound) keeps the in-register walk (BXR_BWD_TAB_DET).
    constexpr bool TABB = (BXR_BWD_TAB == 1 || (BXR_BWD_TAB == 2 && (!std::is_same<TV, float>::value || LPP == 1))) && G >= 8     // G = 4: 64 groups per CTA, the table would not fit 48 KB
                          && (!DET || BXR_BWD_TAB_DET == 1 || (BXR_BWD_TAB_DET == 2 && !std::is_same<TV, float>::value));
    // BXR_BWD_CTAB: wide footprints (the per-point mode) as a window whose slots are the 4 P corners: each lane writes its
    // points' corners as table entries, the group walks them like a window's slots (d per corner by transpose reduction,
    // one scatter per corner) and every lane finishes its own points from the d entries -- instead of 6 broadcast shuffles
    // and 4 full group reductions per point.  Decided by the geometry, like the forward's.
    // Measured r02n (K=4 encoder backward, corner table vs shuffle walk): trained-like boxes 0.784 -> 0.636 ms, init-state
    // boxes 0.348 -> 0.346, uniform points 1.095 -> 1.108 (bound by the L2 atomics either way); the two-levels-per-pass
    // kernels of 2 x 2 grids lose 3 % (0.1855 -> 0.191) and keep the shuffle walk.
    constexpr bool CTABB = BXR_BWD_CTAB != 0 && G >= 8 && LPP == 1 && 4 * SUB * PPL <= CAP;
    __shared__ __align__(16) uint2 s_tab[(TABB || CTABB) ? GROUPS * kTabPitch : 2];
    load_levels(lv, p);
    constexpr int TW = 32 / G, TH = kBwdThreads / 32;
    __shared__ TileOrder<TW, TH> s_order;
    int n_units = p.units;
    if constexpr (TILED) {
        tile_order_init(s_order, lv, p.L);
        n_units = (int)((unsigned)p.B * (unsigned)s_order.before[p.L] * (unsigned)p.H);
    }

    const int lane = threadIdx.x % G;
    const int gid = threadIdx.x / G;
    const int sub = lane / SUB, slane = lane % SUB;
    const unsigned gm = group_mask<G>();
    int* gwin = s_win + gid * kWinPitch;
    float* gdot = s_dot + gid * kWinPitch;
    int* win = gwin + sub * CAP;
    float* dot = gdot + sub * CAP;
    uint2* gtab = s_tab + ((TABB || CTABB) ? gid * kTabPitch : 0);
    uint2* tab = gtab + ((TABB || CTABB) ? sub * CAP : 0);
    const unsigned HDV = (unsigned)(p.H * p.D) / VEC;      // pixel pitch in lane chunks (VEC elements)
    const void* __restrict__ value16 = p.value;   // indexed in lane-chunk units by V::load16
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    ACC* __restrict__ gacc = static_cast<ACC*>(p.grad_value_acc);
    float* __restrict__ grad_loc = static_cast<float*>(p.grad_loc);
    float* __restrict__ grad_w0 = static_cast<float*>(p.grad_w0);
    float dscale = 1.f;
    if constexpr (DET) dscale = *p.det_scale;

#if BXR_UNIT_REVERSE
    for (int u = (TILED ? (int)blockIdx.x : n_units - 1 - (int)blockIdx.x); TILED ? (u < n_units) : (u >= 0); u += TILED ? (int)gridDim.x : -(int)gridDim.x) {
#else
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
#endif
        const long long row_raw = TILED ? tile_order_row(s_order, lv, p, (unsigned)u, gid) : (long long)u * GROUPS + gid;
        const bool ract = TILED ? (row_raw >= 0) : (row_raw < p.rows);     // groups without a row stay with their warp, doing nothing
        const long long row = ract ? row_raw : 0;
        const int head = (int)(row % p.H);
        const long long b = row / ((long long)p.H * p.Nq);
        const unsigned vbase = (unsigned)(b * p.S * HDV + head * G + lane);
        const float* loc_row = FUSED ? nullptr : loc + row * p.LP * 2;
        const float* w_row = w0 + row * p.LP;
        float go[VEC];
        V::load(static_cast<const TV*>(p.grad_out) + (row * p.D + lane * VEC), go);

        for (int l0 = 0; l0 < p.L; l0 += LPP) {
            const int lm = l0 + sub;
            const bool lact = ract && lm < p.L;
            const int lmc = lm < p.L ? lm : 0;
            const int mh = lv.h[lmc], mw = lv.w[lmc];
            LanePoint pt[PPL];
            float g_a[PPL], g_x[PPL], g_y[PPL];     // this lane's results for its own points
            int bx0 = kNoPix, bx1 = -kNoPix, by0 = kNoPix, by1 = -kNoPix;
            float S = 0.f;
            LevelBox lbx;
            if constexpr (FUSED) lbx = load_level_box(p, row * p.L + lmc, b, lmc);
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                const int ptn = lact ? slane + k * SUB : p.P;
                if constexpr (FUSED) pt[k] = lane_point_box<false>(lbx, static_cast<const float*>(p.kidx), w_row + lmc * p.P, ptn, p.P, mh, mw, RowSoftmax{0.f, 1.f});
                else pt[k] = lane_point(loc_row + lmc * p.P * 2, w_row + lmc * p.P, ptn, p.P, mh, mw);
                g_a[k] = g_x[k] = g_y[k] = 0.f;
                if (pt[k].inside) {
                    bx0 = min(bx0, pt[k].x0); bx1 = max(bx1, pt[k].x0 + 1);
                    by0 = min(by0, pt[k].y0); by1 = max(by1, pt[k].y0 + 1);
                    S += fabsf(pt[k].aw);
                }
            }
            SubWin me;
            me.X0 = max(smin<SUB>(bx0, kFullMask), 0); me.Y0 = max(smin<SUB>(by0, kFullMask), 0);
            me.nx = min(smax<SUB>(bx1, kFullMask), mw - 1) - me.X0 + 1;
            me.ny = min(smax<SUB>(by1, kFullMask), mh - 1) - me.Y0 + 1;
            S = ssum<SUB>(S, kFullMask);
            const int nq = (me.nx > 0 && me.ny > 0) ? me.nx * me.ny : 0;
            me.mode = (me.nx <= 0 || me.ny <= 0) ? 0 : ((nq <= CAP && S <= 3.0e38f) ? 1 : 2);
            me.ke = fixed_scale_exp(fmaxf(S, 1e-30f));

            // B: pixel weights (fixed point) in my level's dense nx x ny window.  The d window doubles as a
            //    "touched" flag (1.0) until C overwrites it with the dot products: a pixel touched with zero
            //    total weight still needs its d.
            if (me.mode == 1) {
                zero_window<SUB, CAP>(win, (nq + 3) & ~3, slane);
                zero_window<SUB, CAP>(dot, (nq + 3) & ~3, slane);
            }
            __syncwarp();
            if (me.mode == 1) {
                const float scale = pow2f(me.ke);
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    if (pt[k].inside) {
                        const int sx = pt[k].x0 - me.X0, sy = pt[k].y0 - me.Y0;
                        const float hx = 1.f - pt[k].lx, hy = 1.f - pt[k].ly;
                        const float ax = pt[k].aw * scale * hx, bx = pt[k].aw * scale * pt[k].lx;
                        const int w00 = __float2int_rn(hy * ax), w01 = __float2int_rn(hy * bx);
                        const int w10 = __float2int_rn(pt[k].ly * ax), w11 = __float2int_rn(pt[k].ly * bx);
                        const bool vx0 = sx >= 0, vx1 = sx + 1 < me.nx, vy0 = sy >= 0, vy1 = sy + 1 < me.ny;
                        const int s00 = sy * me.nx + sx;
                        if (vy0 && vx0) { atomicAdd(win + s00, w00); dot[s00] = 1.f; }
                        if (vy0 && vx1) { atomicAdd(win + s00 + 1, w01); dot[s00 + 1] = 1.f; }
                        if (vy1 && vx0) { atomicAdd(win + s00 + me.nx, w10); dot[s00 + me.nx] = 1.f; }
                        if (vy1 && vx1) { atomicAdd(win + s00 + me.nx + 1, w11); dot[s00 + me.nx + 1] = 1.f; }
                    }
                }
            }
            __syncwarp();
            // B': one lane per slot: (offset from the lane's base pointers, weight; NaN = nobody touched this pixel)
            if (TABB && me.mode == 1) {
                const float inv_scale = pow2f(-me.ke);
                const unsigned rcp = (unsigned)__float2int_rz(__fdividef(65536.f, (float)me.nx)) + 1u;   // s / nx, s < 64
                const unsigned tbase = ((unsigned)lv.start[lmc] + (unsigned)(me.Y0 * mw + me.X0)) * HDV;
                const unsigned trow = (unsigned)mw * HDV;
                const int nq4 = (nq + 3) & ~3;
#pragma unroll 2
                for (int ts = slane; ts < nq4; ts += SUB) {
                    const unsigned y = ((unsigned)ts * rcp) >> 16;
                    const unsigned x = (unsigned)ts - y * (unsigned)me.nx;
                    const bool touched = dot[ts] != 0.f;
                    tab[ts] = make_uint2(touched ? tbase + y * trow + x * HDV : kAbsent, __float_as_uint(touched ? (float)win[ts] * inv_scale : 0.f));
                }
            }
            // B'': wide footprint: my points' corners as table entries (offset, attn * bilinear weight; NaN = no such corner)
            if constexpr (CTABB) if (me.mode == 2) {
                const unsigned tlev = (unsigned)lv.start[lmc] * HDV;
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    const int ptn = slane + k * SUB;
                    if (ptn < p.P) {
                        const LanePoint& t = pt[k];
                        const float hx = 1.f - t.lx, hy = 1.f - t.ly;
                        const bool vx0 = t.inside && t.x0 >= 0, vx1 = t.inside && t.x0 + 1 <= mw - 1, vy0 = t.y0 >= 0, vy1 = t.y0 + 1 <= mh - 1;
                        const unsigned c00 = tlev + (unsigned)(t.y0 * mw + t.x0) * HDV;      // wraps for -1; such corners are marked absent
                        uint4* e = reinterpret_cast<uint4*>(tab + ptn * 4);
                        e[0] = make_uint4((vy0 && vx0) ? c00 : kAbsent, __float_as_uint((vy0 && vx0) ? hy * hx * t.aw : 0.f),
                                          (vy0 && vx1) ? c00 + HDV : kAbsent, __float_as_uint((vy0 && vx1) ? hy * t.lx * t.aw : 0.f));
                        e[1] = make_uint4((vy1 && vx0) ? c00 + (unsigned)mw * HDV : kAbsent, __float_as_uint((vy1 && vx0) ? t.ly * hx * t.aw : 0.f),
                                          (vy1 && vx1) ? c00 + (unsigned)mw * HDV + HDV : kAbsent, __float_as_uint((vy1 && vx1) ? t.ly * t.lx * t.aw : 0.f));
                    }
                }
            }
            if constexpr (TABB || CTABB) __syncwarp();

            // C: all G lanes walk the window(s) of this pass
#pragma unroll
            for (int sl = 0; sl < LPP; ++sl) {
                SubWin w;
                if (LPP == 1) {
                    w = me;
                } else {
                    const int src = sl * SUB;
                    w.X0 = __shfl_sync(kFullMask, me.X0, src, G); w.Y0 = __shfl_sync(kFullMask, me.Y0, src, G);
                    w.nx = __shfl_sync(kFullMask, me.nx, src, G); w.ny = __shfl_sync(kFullMask, me.ny, src, G);
                    w.ke = __shfl_sync(kFullMask, me.ke, src, G); w.mode = __shfl_sync(kFullMask, me.mode, src, G);
                }
#if BXR_DIAG == 23
                if (w.mode >= 0) continue;      // DIAGNOSTIC: no walk (phases A, B, D only)
#endif
                if (w.mode == 0) continue;
                const int l = l0 + sl;
                const int lh = lv.h[l], lw = lv.w[l];
#if BXR_BASE_REGPAIR == 2
                const typename V::Raw* vptr = static_cast<const typename V::Raw*>(value16) + vbase;
                ACC* gptr = gacc + (size_t)vbase * VEC;
                asm volatile("" : "+l"(vptr), "+l"(gptr));
                const unsigned lbase = (unsigned)lv.start[l] * HDV;
#else
                const void* vptr = value16;
                ACC* gptr = gacc;
                const unsigned lbase = vbase + (unsigned)lv.start[l] * HDV;
#endif
                if ((TABB && w.mode == 1) || (CTABB && w.mode == 2)) {
                    const uint2* ct = gtab + sl * CAP;
                    float* cdot = gdot + sl * CAP;
                    const int wq_n = w.mode == 1 ? w.nx * w.ny : 4 * p.P;
                    // lane base pointers as opaque register pairs; the table offsets are shared by the group's lanes
                    const typename V::Raw* tvp = static_cast<const typename V::Raw*>(value16) + vbase;
                    ACC* tgp = gacc + (size_t)vbase * VEC;
                    asm volatile("" : "+l"(tvp), "+l"(tgp));
                    // four table entries: gather, d = <go, v> by transpose reduction, one scatter per entry
                    auto walk4 = [&](int q) {
                        const uint4 t0 = *reinterpret_cast<const uint4*>(ct + q);
                        const uint4 t1 = *reinterpret_cast<const uint4*>(ct + q + 2);
                        const unsigned to[4] = {t0.x, t0.z, t1.x, t1.z};
                        const float tw[4] = {__uint_as_float(t0.y), __uint_as_float(t0.w), __uint_as_float(t1.y), __uint_as_float(t1.w)};
                        float v[4][VEC];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (to[j] != kAbsent) {      // a touched pixel / an existing corner
#if BXR_DIAG == 22
                                v[j][0] = v[j][1] = v[j][2] = v[j][3] = __uint_as_float(to[j]);      // DIAGNOSTIC: no gather
#else
                                V::load16(tvp, to[j], v[j]);
#endif
                            } else {
#pragma unroll
                                for (int i = 0; i < VEC; ++i) v[j][i] = 0.f;
                            }
                        }
                        float dsum[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float t = 0.f;
#pragma unroll
                            for (int i = 0; i < VEC; ++i) t += go[i] * v[j][i];
                            dsum[j] = t;
#if BXR_DIAG == 24
                            if (to[j] != kAbsent && tw[j] != 0.f && ((q + j) % 5) >= 2) scatter_row<ACC, VEC>(tgp + (size_t)to[j] * VEC, go, tw[j], dscale);   // DIAGNOSTIC: 60 % of the reductions
#elif BXR_DIAG == 25
                            if (to[j] != kAbsent && tw[j] != 0.f && ((q + j) & 1)) scatter_row<ACC, VEC>(tgp + (size_t)to[j] * VEC, go, tw[j], dscale);   // DIAGNOSTIC: 50 % of the reductions
#elif BXR_DIAG != 21
                            if (to[j] != kAbsent && tw[j] != 0.f) scatter_row<ACC, VEC>(tgp + (size_t)to[j] * VEC, go, tw[j], dscale);   // NaN weights propagate
#endif
                        }
                        float total;
                        const int mine = reduce4<G>(dsum, total, lane, gm);
                        cdot[q + mine] = total;                                // lanes sharing an index write the same value
                    };
                    // two batches in flight for the 8-byte-lane bf16 kernels (r02u: bf16 0.402 -> 0.390 ms; the fp32 kernels,
                    // which come here for wide footprints only, measured 0.4 % slower with it)
                    if constexpr (std::is_same<TV, float>::value) {
                        for (int q = 0; q < wq_n; q += 4) walk4(q);
                    } else {
#pragma unroll 2
                        for (int q = 0; q < wq_n; q += 4) walk4(q);
                    }
                } else if (w.mode == 1) {
                    // per unique pixel, four window slots at a time:
                    // value row -> scatter W*go into grad_value, d = <go, v> by transpose reduction
                    const int* cwin = gwin + sl * CAP;
                    float* cdot = gdot + sl * CAP;
                    const float inv_scale = pow2f(-w.ke);
                    const int wq_n = w.nx * w.ny;
                    const unsigned row_skip = (unsigned)(lw - w.nx) * HDV;
                    unsigned off = lbase + (unsigned)(w.Y0 * lw + w.X0) * HDV;
                    int ix = 0;
                    for (int q = 0; q < wq_n; q += 4) {
                        const int4 wq = *reinterpret_cast<const int4*>(cwin + q);
                        const int wi[4] = {wq.x, wq.y, wq.z, wq.w};
                        const float4 tq = *reinterpret_cast<const float4*>(cdot + q);
                        const bool tv[4] = {tq.x != 0.f, tq.y != 0.f, tq.z != 0.f, tq.w != 0.f};
                        unsigned offs[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            offs[j] = off;
                            off += HDV;
                            if (++ix == w.nx) { ix = 0; off += row_skip; }
                        }
                        float v[4][VEC];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (tv[j]) {               // slots past nq were zeroed and never flagged
                                V::load16(vptr, offs[j], v[j]);
                            } else {
#pragma unroll
                                for (int i = 0; i < VEC; ++i) v[j][i] = 0.f;
                            }
                        }
                        float dsum[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float t = 0.f;
#pragma unroll
                            for (int i = 0; i < VEC; ++i) t += go[i] * v[j][i];
                            dsum[j] = t;
                            if (wi[j] != 0) scatter_row<ACC, VEC>(gptr + (size_t)offs[j] * VEC, go, (float)wi[j] * inv_scale, dscale);
                        }
                        float total;
                        const int mine = reduce4<G>(dsum, total, lane, gm);
                        __syncwarp(gm);                                        // every lane has consumed the flags of these slots
                        cdot[q + mine] = total;                                // lanes sharing an index write the same value
                    }
                } else if constexpr (!CTABB) {
                    // per-point fallback (window too large, or non-finite weights)
#pragma unroll
                    for (int k = 0; k < PPL; ++k) {
#pragma unroll kFbUnrollBwd
                        for (int o = 0; o < SUB; ++o) {
                            if (o + k * SUB >= p.P) break;
                            const int src = sl * SUB + o;
                            const bool inside = __shfl_sync(gm, (int)pt[k].inside, src, G) != 0;
                            const int x0 = __shfl_sync(gm, pt[k].x0, src, G), y0 = __shfl_sync(gm, pt[k].y0, src, G);
                            const float lx = __shfl_sync(gm, pt[k].lx, src, G), ly = __shfl_sync(gm, pt[k].ly, src, G);
                            const float aw = __shfl_sync(gm, pt[k].aw, src, G);
                            if (!inside) continue;
                            const float hx = 1.f - lx, hy = 1.f - ly;
                            const bool vx0 = x0 >= 0, vx1 = x0 + 1 <= lw - 1, vy0 = y0 >= 0, vy1 = y0 + 1 <= lh - 1;
                            const bool ok[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
                            const float cw[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
                            const unsigned c00 = lbase + (unsigned)(y0 * lw + x0) * HDV;
                            float d[4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                float t = 0.f;
                                if (ok[c]) {
                                    const unsigned off = c00 + ((c & 1) ? HDV : 0u) + ((c & 2) ? (unsigned)lw * HDV : 0u);
                                    float v[VEC];
                                    V::load16(vptr, off, v);
#pragma unroll
                                    for (int i = 0; i < VEC; ++i) t += go[i] * v[i];
                                    scatter_row<ACC, VEC>(gptr + (size_t)off * VEC, go, cw[c] * aw, dscale);
                                }
                                d[c] = t;
                            }
#pragma unroll
                            for (int c = 0; c < 4; ++c) d[c] = gsum<G>(d[c], gm);
                            if (lane == src) {
                                g_a[k] = cw[0] * d[0] + cw[1] * d[1] + cw[2] * d[2] + cw[3] * d[3];
                                g_x[k] = (float)lw * aw * (hy * (d[1] - d[0]) + ly * (d[3] - d[2]));
                                g_y[k] = (float)lh * aw * (hx * (d[2] - d[0]) + lx * (d[3] - d[1]));
                            }
                        }
                    }
                }
            }
            __syncwarp();
            // D: finish own points from the d window of my level
            if (me.mode == 1) {
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    if (pt[k].inside) {
                        const int sx = pt[k].x0 - me.X0, sy = pt[k].y0 - me.Y0;
                        const float lx = pt[k].lx, ly = pt[k].ly, hx = 1.f - lx, hy = 1.f - ly;
                        const bool vx0 = sx >= 0, vx1 = sx + 1 < me.nx, vy0 = sy >= 0, vy1 = sy + 1 < me.ny;
                        const int s00 = sy * me.nx + sx;
                        const float d00 = (vy0 && vx0) ? dot[s00] : 0.f;
                        const float d01 = (vy0 && vx1) ? dot[s00 + 1] : 0.f;
                        const float d10 = (vy1 && vx0) ? dot[s00 + me.nx] : 0.f;
                        const float d11 = (vy1 && vx1) ? dot[s00 + me.nx + 1] : 0.f;
                        g_a[k] = hy * hx * d00 + hy * lx * d01 + ly * hx * d10 + ly * lx * d11;
                        g_x[k] = (float)mw * pt[k].aw * (hy * (d01 - d00) + ly * (d11 - d10));
                        g_y[k] = (float)mh * pt[k].aw * (hx * (d10 - d00) + lx * (d11 - d01));
                    }
                }
            }
            if constexpr (CTABB) if (me.mode == 2) {
                // D': finish own points from the d entries of their four corners (absent corners hold 0)
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    const int ptn = slane + k * SUB;
                    if (ptn < p.P && pt[k].inside) {
                        const float4 d = *reinterpret_cast<const float4*>(dot + ptn * 4);
                        const float lx = pt[k].lx, ly = pt[k].ly, hx = 1.f - lx, hy = 1.f - ly;
                        g_a[k] = hy * hx * d.x + hy * lx * d.y + ly * hx * d.z + ly * lx * d.w;
                        g_x[k] = (float)mw * pt[k].aw * (hy * (d.y - d.x) + ly * (d.w - d.z));
                        g_y[k] = (float)mh * pt[k].aw * (hx * (d.z - d.x) + lx * (d.w - d.y));
                    }
                }
            }
            __syncwarp();
            // coalesced stores of this pass's gradients (zeros for points outside the window test)
            if (lact) {
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    const int ptn = slane + k * SUB;
                    if (ptn < p.P) {
                        const long long s = row * p.LP + (long long)lm * p.P + ptn;
                        grad_w0[s] = g_a[k];
                        if constexpr (!FUSED) reinterpret_cast<float2*>(grad_loc)[s] = make_float2(g_x[k], g_y[k]);
                    }
                }
            }
            if constexpr (FUSED) {
                // chain the per-point location gradients to the level's box (cx, cy, w, h) and angle:
                //   loc = (c + R u) * vr,  u = k * relu(size)   (box_attention.py:207-212, :321-336)
                float bcx = 0.f, bcy = 0.f, bw = 0.f, bh = 0.f, ba = 0.f;
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    const int ptn = slane + k * SUB;
                    if (lact && ptn < p.P) {
                        const float2 kk = __ldg(reinterpret_cast<const float2*>(p.kidx) + ptn);
                        const float gx = g_x[k] * lbx.vx, gy = g_y[k] * lbx.vy;
                        const float ux = kk.x * lbx.sx, uy = kk.y * lbx.sy;
                        bcx += gx;
                        bcy += gy;
                        bw += kk.x * (gx * lbx.cs + gy * lbx.sn);
                        bh += kk.y * (gy * lbx.cs - gx * lbx.sn);
                        ba += gx * (-ux * lbx.sn - uy * lbx.cs) + gy * (ux * lbx.cs - uy * lbx.sn);
                    }
                }
                bcx = ssum<SUB>(bcx, kFullMask); bcy = ssum<SUB>(bcy, kFullMask);
                bw = ssum<SUB>(bw, kFullMask); bh = ssum<SUB>(bh, kFullMask);
                if (p.grad_angles) ba = ssum<SUB>(ba, kFullMask);
                if (lact && slane == 0) {
                    const long long rl = row * p.L + lm;
                    reinterpret_cast<float4*>(p.grad_boxes)[rl] =
                        make_float4(bcx, bcy, lbx.pos_w ? bw : 0.f, lbx.pos_h ? bh : 0.f);
                    if (p.grad_angles) static_cast<float*>(p.grad_angles)[rl] = ba;
                }
            }
        }
        if constexpr (SMAX) {
            // the row's weight gradients were written by the lanes of this group: read them back (L2) after a
            // group barrier and chain them through the softmax, in place
            __syncwarp();
            float* gw = grad_w0 + row * p.LP;
            float dotp = 0.f;
            if (ract)
                for (int i = lane; i < p.LP; i += G) dotp += __ldg(w_row + i) * __ldcg(gw + i);
            dotp = gsum<G>(dotp, kFullMask);
            if (ract)
                for (int i = lane; i < p.LP; i += G) gw[i] = __ldg(w_row + i) * (__ldcg(gw + i) - dotp);
        }
    }
}

}  // namespace bxr
