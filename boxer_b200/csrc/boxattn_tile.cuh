// "Query-tile x value-tile" kernels for self-attention-shaped box attention (BoxeR's encoder: the queries ARE the
// pixels of the value pyramid, Nq == S, and every query's reference window is centred on its own pixel,
// e2edet/module/box_transformer.py:70-116).
//
// What the footprint-window kernels (boxattn_window.cuh) leave on the table there (ncu, profiles/r01x_*): work units
// are dealt round-robin over all SMs, so the ~80 % overlap between the windows of neighbouring queries never meets
// in one L1 (sector hit rate 29 %), every unique pixel of every window is a 128-byte gather that goes to L2
// (long-scoreboard is the top stall), and 42 % of all instructions are the window walk's address arithmetic.
//
// Here a CTA owns a 2-D tile of 8 x 8 queries of one level for one head, and
//   * TMA-stages the value pixels that tile can reach -- per level, the tile's footprint plus a halo sized for the
//     reference windows -- into shared memory with bulk copies (cp.async.bulk + mbarrier: one 128-byte copy per
//     (pixel, head) row, issued by all threads, landing while the first level's sample points are being loaded);
//   * gives each (query, head) row to FOUR lanes holding 8 channels each (a warp = the 8 queries of one tile line),
//     so the per-(row, level) bookkeeping of the window algorithm is shared by 8 rows per warp instead of 4;
//   * scatters the K x K bilinear weights of a (row, level) into a dense 8 x 8 window of shared memory exactly like
//     the window kernels (32-bit fixed point, integer ATOMS), and walks it with a fixed pitch: no row-wrap
//     arithmetic, value rows come from shared memory at immediate offsets;
//   * verifies per (row, level) that the window lies inside the staged region -- staging is a cache keyed on
//     geometry, never an assumption: rows whose boxes have wandered (learned offsets are unbounded), levels whose
//     footprint does not fit (a coarse-level tile looking at a fine level) and non-pixel queries take the same walk
//     over global memory, wide footprints the per-point walk.
// The sums are the reference's (box_attn_kernel.cuh:311-346), re-associated.
#pragma once

#include "boxattn_staged.cuh"

namespace bxr {

#ifndef BXR_TILE_MINB
#define BXR_TILE_MINB 3
#endif
#ifndef BXR_TILE_PIX
#define BXR_TILE_PIX 320          // staged (pixel, head) rows per tile, all levels together
#endif
// halo of a staged region, in units of the reference-window size seen from the target level: the init-state box
// reaches 1.69 box-quarters left and 2.19 right of the query's pixel centre (offset in [0, 1/2) px, size 4 .. 4.5 px)
#ifndef BXR_TILE_HALO_LO
#define BXR_TILE_HALO_LO 2.0f
#endif
#ifndef BXR_TILE_HALO_HI
#define BXR_TILE_HALO_HI 2.5f
#endif

constexpr int kTileW = 8, kTileH = 8;                 // queries per tile (x, y)
constexpr int kTileThreads = 256;                     // 8 warps = 8 tile lines; a warp = 8 rows x 4 lanes
constexpr int kTileRows = kTileW * kTileH;
constexpr int kTG = 4;                                // lanes per row
constexpr int kTWinPitch = 64 + 4;                    // ints per row window (8 x 8, pitch 8) + bank skew
constexpr int kTilePix = BXR_TILE_PIX;

struct TileRegion {
    int x0, y0, w, h;      // staged pixel rectangle of one level (w == 0: not staged)
    int off;               // first slot in the pool
};

// per-lane channel chunk: 8 channels = NV vectors of 16 bytes
template <typename TV> struct TileLane;
template <> struct TileLane<float> {
    static constexpr int NV = 2, ROWB = 128, LANEB = 32;
    __device__ __forceinline__ static void fma(const uint4& t, float w, float* a) {
        a[0] += w * __uint_as_float(t.x); a[1] += w * __uint_as_float(t.y);
        a[2] += w * __uint_as_float(t.z); a[3] += w * __uint_as_float(t.w);
    }
    __device__ __forceinline__ static uint4 pack(const float* a) {
        return make_uint4(__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3]));
    }
    // <t, g> over the vector's channels (g: the matching 4 entries of the lane's 8)
    __device__ __forceinline__ static float dot(const uint4& t, const float* g) {
        return __uint_as_float(t.x) * g[0] + __uint_as_float(t.y) * g[1] + __uint_as_float(t.z) * g[2] + __uint_as_float(t.w) * g[3];
    }
    __device__ __forceinline__ static void unpack(const uint4& t, float* a) {
        a[0] = __uint_as_float(t.x); a[1] = __uint_as_float(t.y); a[2] = __uint_as_float(t.z); a[3] = __uint_as_float(t.w);
    }
};
template <> struct TileLane<__nv_bfloat16> {
    static constexpr int NV = 1, ROWB = 64, LANEB = 16;
    __device__ __forceinline__ static void fma(const uint4& t, float w, float* a) {
        const unsigned u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[2 * i] += w * __uint_as_float(u[i] << 16);
            a[2 * i + 1] += w * __uint_as_float(u[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ static uint4 pack(const float* a) {
        unsigned u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
            u[i] = *reinterpret_cast<const unsigned*>(&h);
        }
        return make_uint4(u[0], u[1], u[2], u[3]);
    }
    __device__ __forceinline__ static float dot(const uint4& t, const float* g) {
        const unsigned u[4] = {t.x, t.y, t.z, t.w};
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) d += __uint_as_float(u[i] << 16) * g[2 * i] + __uint_as_float(u[i] & 0xffff0000u) * g[2 * i + 1];
        return d;
    }
    __device__ __forceinline__ static void unpack(const uint4& t, float* a) {
        const unsigned u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[2 * i] = __uint_as_float(u[i] << 16);
            a[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
        }
    }
};

// BXR_TILE_COPY: how the value pixels a tile can reach get close to the SM
//   0 = TMA bulk copies into a shared-memory pool (cp.async.bulk, one per (pixel, head) row, completion on an mbarrier)
//   1 = the same pool filled with 16-byte ld.global / st.shared by all threads, completion by __syncthreads (A/B)
//   2 = no pool: the tile ordering alone -- the 8 warps of a CTA walk the 8 lines of one (tile, head) at the same time
//       and gather through L1, which then holds the tile's footprint; no CTA-level synchronisation at all
#ifndef BXR_TILE_COPY
#define BXR_TILE_COPY 0
#endif
constexpr bool kTilePool = BXR_TILE_COPY != 2;
constexpr int kTilePoolPix = kTilePool ? kTilePix : 0;

__host__ __device__ inline unsigned tile_smem_bytes(int rowb, bool backward) {
    // [ value pool ][ weight windows ][ backward: d windows ]
    return (unsigned)kTilePoolPix * rowb + (unsigned)kTileRows * kTWinPitch * 4u * (backward ? 2u : 1u);
}

__device__ __forceinline__ int q4min(int v) {
    v = min(v, __shfl_xor_sync(kFullMask, v, 1));
    return min(v, __shfl_xor_sync(kFullMask, v, 2));
}
__device__ __forceinline__ int q4max(int v) {
    v = max(v, __shfl_xor_sync(kFullMask, v, 1));
    return max(v, __shfl_xor_sync(kFullMask, v, 2));
}
__device__ __forceinline__ float q4sum(float v) {
    v += __shfl_xor_sync(kFullMask, v, 1);
    return v + __shfl_xor_sync(kFullMask, v, 2);
}

// Tile bookkeeping shared by forward and backward.  Work item = (image, tile, head); 32-bit arithmetic throughout
// (the host checks that B * tiles * H fits).  Warp 0 prepares the descriptor of the CTA's NEXT item while the
// current one is being processed, so an item costs one CTA barrier.
struct TileDesc {
    int b, head, lq, tx0, ty0, total;      // total: staged (pixel, head) rows of all levels
    TileRegion reg[kMaxLevels];
};

// tiles of the coarse levels first (assuming, as in BoxeR, that later levels are coarser: their rows look at fine
// levels through wide footprints and cost several times more -- a scheduling order only, nothing depends on it)
__device__ __forceinline__ void tile_prefix(const LevelTable& lv, int L, int* before) {
    int acc = 0;
    for (int k = 0; k < L; ++k) {
        const int l = L - 1 - k;
        before[k] = acc;
        acc += ((lv.w[l] + kTileW - 1) / kTileW) * ((lv.h[l] + kTileH - 1) / kTileH);
    }
    before[L] = acc;
}

// staged rectangle of level l for a tile of level lq: the tile's query centres mapped to level l, plus the halo
__device__ __forceinline__ void tile_span(int t0, int t1, int size_q, int size_l, int& lo, int& hi) {
    const float r = __fdividef((float)size_l, (float)size_q);
    lo = (int)floorf(((float)t0 + 0.5f) * r - 0.5f - BXR_TILE_HALO_LO * r);
    hi = (int)floorf(((float)t1 - 0.5f) * r - 0.5f + BXR_TILE_HALO_HI * r) + 1;
    lo = max(lo, 0);
    hi = min(hi, size_l - 1);
}

// item -> (image, head, level, tile origin)
__device__ __forceinline__ void tile_decode(const LevelTable& lv, const int* before, int L, int H, unsigned item,
                                            int& b, int& head, int& lq, int& tx0, int& ty0) {
    const unsigned u = item / (unsigned)H;
    head = (int)(item - u * (unsigned)H);
    const unsigned tiles_img = (unsigned)before[L];
    b = (int)(u / tiles_img);
    const int ti = (int)(u - (unsigned)b * tiles_img);
    int k = 0;
    while (k + 1 < L && before[k + 1] <= ti) ++k;
    lq = L - 1 - k;
    const int tl = ti - before[k];
    const int ntx = (lv.w[lq] + kTileW - 1) / kTileW;
    const int tyi = tl / ntx;
    ty0 = tyi * kTileH;
    tx0 = (tl - tyi * ntx) * kTileW;
}

// warp 0, all lanes: decode the item, lane l sizes the region of level l, lane 0 packs them into the pool greedily
// (own level first, then the later levels, then the earlier ones; what does not fit is not staged)
__device__ __forceinline__ void tile_prepare(const LevelTable& lv, const int* before, int L, int H, unsigned item, TileDesc& d) {
    const int lane = threadIdx.x & 31;
    int b, head, lq, tx0, ty0;
    tile_decode(lv, before, L, H, item, b, head, lq, tx0, ty0);
    if (lane < L) {
        const int tx1 = min(tx0 + kTileW, lv.w[lq]), ty1 = min(ty0 + kTileH, lv.h[lq]);
        int x0, x1, y0, y1;
        tile_span(tx0, tx1, lv.w[lq], lv.w[lane], x0, x1);
        tile_span(ty0, ty1, lv.h[lq], lv.h[lane], y0, y1);
        TileRegion g;
        g.x0 = x0; g.y0 = y0; g.w = max(x1 - x0 + 1, 0); g.h = max(y1 - y0 + 1, 0); g.off = 0;
        d.reg[lane] = g;
    }
    __syncwarp();
    if (lane == 0) {
        int used = 0;
        for (int i = 0; i < L; ++i) {
            const int l = (i < L - lq) ? lq + i : L - 1 - i;
            const int n = d.reg[l].w * d.reg[l].h;
            if (n <= 0 || used + n > kTilePix) { d.reg[l].w = 0; d.reg[l].h = 0; }
            else { d.reg[l].off = used; used += n; }
        }
        d.b = b; d.head = head; d.lq = lq; d.tx0 = tx0; d.ty0 = ty0; d.total = used;
    }
}


// all threads: one bulk copy per staged (pixel, head) row
template <int ROWB>
__device__ __forceinline__ void tile_issue(const AttnParams& p, const LevelTable& lv, const TileDesc& d,
                                           unsigned char* s_val, unsigned long long* bar) {
    const unsigned char* vimg = static_cast<const unsigned char*>(p.value) + ((size_t)d.b * p.S * p.H + d.head) * ROWB;
    const unsigned ppitch = (unsigned)p.H * ROWB;
    for (int l = 0; l < p.L; ++l) {
        const TileRegion g = d.reg[l];
        const int n = g.w * g.h;
        if (n <= 0) continue;
        const unsigned char* vl = vimg + (size_t)lv.start[l] * ppitch;
        const int lw = lv.w[l];
#if BXR_TILE_COPY == 0
        for (int k = threadIdx.x; k < n; k += kTileThreads) {
            const int ry = k / g.w, rx = k - ry * g.w;
            const unsigned pix = (unsigned)((g.y0 + ry) * lw + g.x0 + rx);
            bulk_g2s(s_val + (unsigned)(g.off + k) * ROWB, vl + (size_t)pix * ppitch, ROWB, bar);
        }
#else
        constexpr int CH = ROWB / 16;       // 16-byte chunks per row
        for (int i = threadIdx.x; i < n * CH; i += kTileThreads) {
            const int k = i / CH, ch = i % CH;
            const int ry = k / g.w, rx = k - ry * g.w;
            const unsigned pix = (unsigned)((g.y0 + ry) * lw + g.x0 + rx);
            *reinterpret_cast<uint4*>(s_val + (unsigned)(g.off + k) * ROWB + ch * 16) =
                __ldg(reinterpret_cast<const uint4*>(vl + (size_t)pix * ppitch + ch * 16));
        }
        (void)bar;
#endif
    }
}

__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }
__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
// BXR_TILE_PREFETCH: 0 off, 1 = next level's locations / weights into L1 one level ahead, 2 = into L2
#ifndef BXR_TILE_PREFETCH
#define BXR_TILE_PREFETCH 1
#endif

// the sampling locations / weights of the warp's rows in its NEXT item start their trip from DRAM to L2 one item ahead
// (they are the kernel's only DRAM stream; the next level's lines are then pulled into L1 one level ahead)
__device__ __forceinline__ void tile_prefetch_next(const AttnParams& p, const LevelTable& lv, const int* before, unsigned item,
                                                   unsigned items, int wrp, int r, int j) {
#if BXR_TILE_PREFETCH
    if (item >= items) return;
    int b, head, lq, tx0, ty0;
    tile_decode(lv, before, p.L, p.H, item, b, head, lq, tx0, ty0);
    const int qx = tx0 + r, qy = ty0 + wrp;
    if (qx >= lv.w[lq] || qy >= lv.h[lq]) return;
    const long long q = lv.start[lq] + (long long)qy * lv.w[lq] + qx;
    if (q >= p.Nq) return;
    const long long row = ((long long)b * p.Nq + q) * p.H + head;
    const unsigned char* lp = reinterpret_cast<const unsigned char*>(static_cast<const float*>(p.loc) + row * p.LP * 2);
    const unsigned char* wp = reinterpret_cast<const unsigned char*>(static_cast<const float*>(p.w0) + row * p.LP);
    for (int o = j * 128; o < p.LP * 8; o += 4 * 128) prefetch_l2(lp + o);
    for (int o = j * 128; o < p.LP * 4; o += 4 * 128) prefetch_l2(wp + o);
#endif
}

// scatter one point's four bilinear corner weights (32-bit fixed point) into the row's pitch-8 window
__device__ __forceinline__ void tile_scatter(int* win, const LanePoint& t, float scale, int X0, int Y0, int nx, int ny) {
    const int sx = t.x0 - X0, sy = t.y0 - Y0;       // -1 .. n-1
    const float hx = 1.f - t.lx, hy = 1.f - t.ly;
    const float ax = t.aw * scale * hx, bx = t.aw * scale * t.lx;
    const int w00 = __float2int_rn(hy * ax), w01 = __float2int_rn(hy * bx);
    const int w10 = __float2int_rn(t.ly * ax), w11 = __float2int_rn(t.ly * bx);
    int* wp = win + sy * 8 + sx;
    if (sx >= 0 && sy >= 0 && sx + 1 < nx && sy + 1 < ny) {      // interior point: the common case
        atomicAdd(wp, w00); atomicAdd(wp + 1, w01); atomicAdd(wp + 8, w10); atomicAdd(wp + 9, w11);
    } else {
        const bool vx0 = sx >= 0, vx1 = sx + 1 < nx, vy0 = sy >= 0, vy1 = sy + 1 < ny;
        if (vy0 && vx0) atomicAdd(wp, w00);
        if (vy0 && vx1) atomicAdd(wp + 1, w01);
        if (vy1 && vx0) atomicAdd(wp + 8, w10);
        if (vy1 && vx1) atomicAdd(wp + 9, w11);
    }
}

// Window walk of the forward: the rows of a warp whose (row, level) is in window mode, value rows from the shared-memory
// pool (STAGED) or from global memory through L1.  (An explicitly batched variant -- all eight 16-byte loads of four
// slots requested before the first use -- measured slower for fp32: 0.215 vs 0.195 ms, 118 M vs 98 M instructions, r02h.)
template <typename TV, bool STAGED>
__device__ __forceinline__ void tile_fwd_walk(bool mine, int nx, int ny, const int* win, const unsigned char* vp, size_t vrow_pitch,
                                              size_t vpix_pitch, int rpar, float inv_scale, float* acc) {
    using TL = TileLane<TV>;
    constexpr int NV = TL::NV;
    const int nym = __reduce_max_sync(kFullMask, mine ? ny : 0);
    const int nxm = __reduce_max_sync(kFullMask, mine ? nx : 0);
    for (int wy = 0; wy < nym; ++wy) {
        const int4 wa = *reinterpret_cast<const int4*>(win + wy * 8);
        int4 wb = make_int4(0, 0, 0, 0);
        if (nxm > 4) wb = *reinterpret_cast<const int4*>(win + wy * 8 + 4);
        const int wi[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int wx = 0; wx < 8; ++wx) {
            if (wx < nxm && mine && wi[wx] != 0) {      // slots outside the row's nx x ny range were zeroed and never written
                const float wv = (float)wi[wx] * inv_scale;
                const unsigned char* a = vp + wx * vpix_pitch;
                if (STAGED) {
                    const uint4 v0 = *reinterpret_cast<const uint4*>(a + rpar * 16);
                    TL::fma(v0, wv, acc);
                    if (NV == 2) {
                        const uint4 v1 = *reinterpret_cast<const uint4*>(a + (rpar ^ 1) * 16);
                        TL::fma(v1, wv, acc + 4);
                    }
                } else {
                    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(a + rpar * 16));
                    TL::fma(v0, wv, acc);
                    if (NV == 2) {
                        const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(a + (rpar ^ 1) * 16));
                        TL::fma(v1, wv, acc + 4);
                    }
                }
            }
        }
        vp += vrow_pitch;
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <typename TV, int PPL>
__global__ void __launch_bounds__(kTileThreads, BXR_TILE_MINB) box_fwd_tile_kernel(const AttnParams p) {
    using TL = TileLane<TV>;
    constexpr int NV = TL::NV, ROWB = TL::ROWB, LANEB = TL::LANEB;
    extern __shared__ __align__(128) unsigned char t_smem[];
    unsigned char* s_val = t_smem;
    int* s_win = reinterpret_cast<int*>(t_smem + kTilePoolPix * ROWB);
    __shared__ LevelTable lv;
    __shared__ TileDesc s_desc[2];
    __shared__ int s_before[kMaxLevels + 1];
    __shared__ __align__(8) unsigned long long s_bar;

    load_levels(lv, p);
    if (threadIdx.x == 0) {
        tile_prefix(lv, p.L, s_before);
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int r = lane >> 2, j = lane & 3;              // row of the warp's tile line, channel chunk
    const int rpar = (NV == 2) ? (r & 1) : 0;           // odd rows read their two 16-byte halves in swapped order (banks)
    int* win = s_win + (wrp * 8 + r) * kTWinPitch;
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    const unsigned items = (unsigned)p.B * (unsigned)s_before[p.L] * (unsigned)p.H;
#if BXR_TILE_COPY != 2
    unsigned parity = 0;
    if (wrp == 0 && blockIdx.x < items) tile_prepare(lv, s_before, p.L, p.H, blockIdx.x, s_desc[0]);
    __syncthreads();
#endif

    int it = 0;
    for (unsigned item = blockIdx.x; item < items; item += gridDim.x, ++it) {
#if BXR_TILE_COPY == 2
        // no pool: every warp decodes for itself and runs free of the CTA's other warps
        TileDesc t;
        tile_decode(lv, s_before, p.L, p.H, item, t.b, t.head, t.lq, t.tx0, t.ty0);
        bool landed = true;
        tile_prefetch_next(p, lv, s_before, item + gridDim.x, items, wrp, r, j);
#else
        const TileDesc& t = s_desc[it & 1];
        // the pool is free and the descriptor complete: guaranteed by the barrier that ended the previous item
#if BXR_TILE_COPY == 0
        if (threadIdx.x == 0) mbar_expect_tx(&s_bar, (unsigned)t.total * ROWB);
#endif
        tile_issue<ROWB>(p, lv, t, s_val, &s_bar);
#if BXR_TILE_COPY == 0
        bool landed = false;
#else
        __syncthreads();
        bool landed = true;
#endif
        if (wrp == 0 && item + gridDim.x < items) tile_prepare(lv, s_before, p.L, p.H, item + gridDim.x, s_desc[(it + 1) & 1]);
        const TileRegion* s_reg = t.reg;
#endif

        const int qx = t.tx0 + r, qy = t.ty0 + wrp;
        const bool ract = qx < lv.w[t.lq] && qy < lv.h[t.lq];
        const long long q = ract ? (lv.start[t.lq] + (long long)qy * lv.w[t.lq] + qx) : 0;
        const bool qok = ract && q < p.Nq;
        const long long row = qok ? (((long long)t.b * p.Nq + q) * p.H + t.head) : 0;
        const float* loc_row = loc + row * p.LP * 2;
        const float* w_row = w0 + row * p.LP;
        const unsigned char* vimg = static_cast<const unsigned char*>(p.value) + ((size_t)t.b * p.S * p.H + t.head) * ROWB;

        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;

        for (int l = 0; l < p.L; ++l) {
            const int lh = lv.h[l], lw = lv.w[l];
#if BXR_TILE_PREFETCH
            if (l + 1 < p.L && qok && j < 2) {       // the next level's operands start their trip from DRAM now
                const void* nxt = j == 0 ? static_cast<const void*>(loc_row + (l + 1) * p.P * 2) : static_cast<const void*>(w_row + (l + 1) * p.P);
                if (BXR_TILE_PREFETCH == 1) prefetch_l1(nxt); else prefetch_l2(nxt);
            }
#endif
            // ---- A: own points, touched pixel range of the row
            LanePoint pt[PPL];
            int bx0 = kNoPix, bx1 = -kNoPix, by0 = kNoPix, by1 = -kNoPix;
            float S = 0.f;
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                const int ptn = qok ? j + k * kTG : p.P;
                pt[k] = lane_point(loc_row + l * p.P * 2, w_row + l * p.P, ptn, p.P, lh, lw);
                if (pt[k].inside) {
                    bx0 = min(bx0, pt[k].x0); bx1 = max(bx1, pt[k].x0 + 1);
                    by0 = min(by0, pt[k].y0); by1 = max(by1, pt[k].y0 + 1);
                    S += fabsf(pt[k].aw);
                }
            }
            const int X0 = max(q4min(bx0), 0), Y0 = max(q4min(by0), 0);
            const int nx = min(q4max(bx1), lw - 1) - X0 + 1, ny = min(q4max(by1), lh - 1) - Y0 + 1;
            S = q4sum(S);
            // 0 skip, 1 window (fits 8 x 8), 2 per-point walk (wide footprint or non-finite weights)
            const int mode = (nx <= 0 || ny <= 0 || S == 0.f) ? 0 : ((nx <= 8 && ny <= 8 && S <= 3.0e38f) ? 1 : 2);
            const int ke = fixed_scale_exp(S);
#if BXR_TILE_COPY == 2
            TileRegion g;
            g.x0 = g.y0 = g.w = g.h = g.off = 0;
            const bool staged = false;
#else
            const TileRegion g = s_reg[l];
            const bool staged = mode == 1 && g.w > 0 && X0 >= g.x0 && Y0 >= g.y0 && X0 + nx <= g.x0 + g.w && Y0 + ny <= g.y0 + g.h;
#endif

            // ---- B: pixel weights into the row's dense 8 x 8 window (32-bit fixed point, pitch 8)
            if (mode == 1) {
#pragma unroll
                for (int z = 0; z < 4; ++z) *reinterpret_cast<uint4*>(win + (z * kTG + j) * 4) = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            if (mode == 1) {
                const float scale = pow2f(ke);
#pragma unroll
                for (int k = 0; k < PPL; ++k)
                    if (pt[k].inside) tile_scatter(win, pt[k], scale, X0, Y0, nx, ny);
            }
            __syncwarp();

            // ---- C: walk.  Three warp-uniform passes: staged windows, windows over global memory, per-point rows.
            const unsigned m_st = __ballot_sync(kFullMask, staged);
            const unsigned m_gl = __ballot_sync(kFullMask, mode == 1 && !staged);
            const unsigned m_pp = __ballot_sync(kFullMask, mode == 2);
            const float inv_scale = pow2f(-ke);
            if (m_st) {
#if BXR_TILE_COPY == 0
                if (!landed) { mbar_wait(&s_bar, parity); landed = true; }
#endif
                const unsigned char* vp = s_val + ((size_t)g.off + (staged ? (Y0 - g.y0) * g.w + (X0 - g.x0) : 0)) * ROWB + j * LANEB;
                tile_fwd_walk<TV, true>(staged, nx, ny, win, vp, (size_t)g.w * ROWB, ROWB, rpar, inv_scale, acc);
            }
            if (m_gl) {
                const bool mine = mode == 1 && !staged;
                const size_t ppitch = (size_t)p.H * ROWB;
                const unsigned char* vp = vimg + ((size_t)lv.start[l] + (mine ? (size_t)Y0 * lw + X0 : 0)) * ppitch + j * LANEB;
                tile_fwd_walk<TV, false>(mine, nx, ny, win, vp, (size_t)lw * ppitch, ppitch, rpar, inv_scale, acc);
            }
            if (m_pp) {
                // per-point walk over global memory; the owner lane broadcasts its tap to the row's four lanes
                const unsigned char* vl = vimg + (size_t)lv.start[l] * p.H * ROWB + j * LANEB;
                const size_t ppitch = (size_t)p.H * ROWB;
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
#pragma unroll
                    for (int o = 0; o < kTG; ++o) {
                        if (o + k * kTG >= p.P) break;
                        const bool inside = __shfl_sync(kFullMask, (int)pt[k].inside, o, kTG) != 0;
                        const int x0 = __shfl_sync(kFullMask, pt[k].x0, o, kTG), y0 = __shfl_sync(kFullMask, pt[k].y0, o, kTG);
                        const float lx = __shfl_sync(kFullMask, pt[k].lx, o, kTG), ly = __shfl_sync(kFullMask, pt[k].ly, o, kTG);
                        const float aw = __shfl_sync(kFullMask, pt[k].aw, o, kTG);
                        if (mode == 2 && inside) {
                            const float hx = 1.f - lx, hy = 1.f - ly;
                            const bool vx0 = x0 >= 0, vx1 = x0 + 1 <= lw - 1, vy0 = y0 >= 0, vy1 = y0 + 1 <= lh - 1;
                            const bool ok[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
                            const float cw[4] = {hy * hx * aw, hy * lx * aw, ly * hx * aw, ly * lx * aw};
                            const unsigned char* c00 = vl + ((long long)y0 * lw + x0) * (long long)ppitch;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                if (ok[c]) {
                                    const unsigned char* cp = c00 + ((c & 1) ? ppitch : 0) + ((c & 2) ? (size_t)lw * ppitch : 0);
                                    const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(cp + rpar * 16));
                                    TL::fma(v0, cw[c], acc);
                                    if (NV == 2) {
                                        const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(cp + (rpar ^ 1) * 16));
                                        TL::fma(v1, cw[c], acc + 4);
                                    }
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();     // the windows are re-zeroed by the next level
        }
#if BXR_TILE_COPY == 0
        if (!landed) mbar_wait(&s_bar, parity);          // keep the barrier's phases in step with the items
        parity ^= 1u;
#endif
        if (qok) {
            unsigned char* op = static_cast<unsigned char*>(p.out) + (size_t)row * ROWB + j * LANEB;
            *reinterpret_cast<uint4*>(op + rpar * 16) = TL::pack(acc);
            if (NV == 2) *reinterpret_cast<uint4*>(op + (rpar ^ 1) * 16) = TL::pack(acc + 4);
        }
#if BXR_TILE_COPY != 2
        __syncthreads();      // every warp is done with the pool and the windows; warp 0 has written the next descriptor
#endif
        (void)landed;
    }
}

// ------------------------------------------------------------------------------------------------ backward
#ifndef BXR_TILE_BWD_MINB
#define BXR_TILE_BWD_MINB 2
#endif

// grad_value[pixel row] += wgt * go for the lane's 8 channels.  `dst` points at the lane's first-loaded vector's
// channels; `swap` says the lane's second vector lies 4 elements BEFORE it (odd rows read their halves swapped).
template <typename ACC, int NV>
__device__ __forceinline__ void tile_scatter_grad(ACC* dst, int rpar, const float* go, float wgt, float dscale) {
    if constexpr (sizeof(ACC) == 8) {
        constexpr int N0 = NV == 2 ? 4 : 8;
#pragma unroll
        for (int i = 0; i < N0; ++i) red_add_fixed(reinterpret_cast<long long*>(dst) + rpar * 4 + i, wgt * go[i], dscale);
        if (NV == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) red_add_fixed(reinterpret_cast<long long*>(dst) + (rpar ^ 1) * 4 + i, wgt * go[4 + i], dscale);
        }
    } else {
        float* d = reinterpret_cast<float*>(dst);
        if (NV == 2) {
            red_add_v4(d + rpar * 4, wgt * go[0], wgt * go[1], wgt * go[2], wgt * go[3]);
            red_add_v4(d + (rpar ^ 1) * 4, wgt * go[4], wgt * go[5], wgt * go[6], wgt * go[7]);
        } else {
            red_add_v4(d, wgt * go[0], wgt * go[1], wgt * go[2], wgt * go[3]);
            red_add_v4(d + 4, wgt * go[4], wgt * go[5], wgt * go[6], wgt * go[7]);
        }
    }
}

// Walk of the rows of a warp whose (row, level) is in window mode -- STAGED: value rows from the shared-memory pool,
// otherwise from global memory.  Per in-range slot of the dense nx x ny window: d[slot] = <go, value row> (four slots
// at a time, transpose-reduced over the row's four lanes) and grad_value[pixel] += W[slot] * go.
template <typename TV, typename ACC, bool STAGED>
__device__ __forceinline__ void tile_bwd_walk(bool mine, int nx, int ny, const int* win, float* dwin, const unsigned char* vp,
                                              size_t vrow_pitch, size_t vpix_pitch, ACC* gp, size_t grow_pitch, size_t gpix_pitch,
                                              int rpar, int j, const float* go, float inv_scale, float dscale) {
    using TL = TileLane<TV>;
    constexpr int NV = TL::NV;
    const int nym = __reduce_max_sync(kFullMask, mine ? ny : 0);
    const int nxm = __reduce_max_sync(kFullMask, mine ? nx : 0);
    for (int wy = 0; wy < nym; ++wy) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            if (hf * 4 < nxm) {
                const int4 wq = *reinterpret_cast<const int4*>(win + wy * 8 + hf * 4);
                const int wi[4] = {wq.x, wq.y, wq.z, wq.w};
                float ds[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int wx = hf * 4 + s;
                    float t = 0.f;
                    if (mine && wx < nx && wy < ny) {
                        const unsigned char* a = vp + wx * vpix_pitch;
                        uint4 v0, v1;
                        if (STAGED) {
                            v0 = *reinterpret_cast<const uint4*>(a + rpar * 16);
                            if (NV == 2) v1 = *reinterpret_cast<const uint4*>(a + (rpar ^ 1) * 16);
                        } else {
                            v0 = __ldg(reinterpret_cast<const uint4*>(a + rpar * 16));
                            if (NV == 2) v1 = __ldg(reinterpret_cast<const uint4*>(a + (rpar ^ 1) * 16));
                        }
                        t = TL::dot(v0, go);
                        if (NV == 2) t += TL::dot(v1, go + 4);
                        if (wi[s] != 0) tile_scatter_grad<ACC, NV>(gp + wx * gpix_pitch, rpar, go, (float)wi[s] * inv_scale, dscale);
                    }
                    ds[s] = t;
                }
                float total;
                const int idx = reduce4<4>(ds, total, j, kFullMask);
                if (mine) dwin[wy * 8 + hf * 4 + idx] = total;
            }
        }
        vp += vrow_pitch;
        gp += grow_pitch;
    }
}

template <typename TV, int PPL, typename ACC>
__global__ void __launch_bounds__(kTileThreads, BXR_TILE_BWD_MINB) box_bwd_tile_kernel(const AttnParams p) {
    using TL = TileLane<TV>;
    constexpr int NV = TL::NV, ROWB = TL::ROWB, LANEB = TL::LANEB;
    constexpr bool DET = sizeof(ACC) == 8;
    extern __shared__ __align__(128) unsigned char t_smem[];
    unsigned char* s_val = t_smem;
    int* s_win = reinterpret_cast<int*>(t_smem + kTilePoolPix * ROWB);
    float* s_dot = reinterpret_cast<float*>(s_win + kTileRows * kTWinPitch);
    __shared__ LevelTable lv;
    __shared__ TileDesc s_desc[2];
    __shared__ int s_before[kMaxLevels + 1];
    __shared__ __align__(8) unsigned long long s_bar;

    load_levels(lv, p);
    if (threadIdx.x == 0) {
        tile_prefix(lv, p.L, s_before);
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int r = lane >> 2, j = lane & 3;
    const int rpar = (NV == 2) ? (r & 1) : 0;
    int* win = s_win + (wrp * 8 + r) * kTWinPitch;
    float* dwin = s_dot + (wrp * 8 + r) * kTWinPitch;
    const float* __restrict__ loc = static_cast<const float*>(p.loc);
    const float* __restrict__ w0 = static_cast<const float*>(p.w0);
    float* __restrict__ grad_loc = static_cast<float*>(p.grad_loc);
    float* __restrict__ grad_w0 = static_cast<float*>(p.grad_w0);
    float dscale = 1.f;
    if constexpr (DET) dscale = *p.det_scale;
    const unsigned items = (unsigned)p.B * (unsigned)s_before[p.L] * (unsigned)p.H;
#if BXR_TILE_COPY != 2
    unsigned parity = 0;
    if (wrp == 0 && blockIdx.x < items) tile_prepare(lv, s_before, p.L, p.H, blockIdx.x, s_desc[0]);
    __syncthreads();
#endif

    int it = 0;
    for (unsigned item = blockIdx.x; item < items; item += gridDim.x, ++it) {
#if BXR_TILE_COPY == 2
        TileDesc t;
        tile_decode(lv, s_before, p.L, p.H, item, t.b, t.head, t.lq, t.tx0, t.ty0);
        bool landed = true;
        tile_prefetch_next(p, lv, s_before, item + gridDim.x, items, wrp, r, j);
#else
        const TileDesc& t = s_desc[it & 1];
#if BXR_TILE_COPY == 0
        if (threadIdx.x == 0) mbar_expect_tx(&s_bar, (unsigned)t.total * ROWB);
#endif
        tile_issue<ROWB>(p, lv, t, s_val, &s_bar);
#if BXR_TILE_COPY == 0
        bool landed = false;
#else
        __syncthreads();
        bool landed = true;
#endif
        if (wrp == 0 && item + gridDim.x < items) tile_prepare(lv, s_before, p.L, p.H, item + gridDim.x, s_desc[(it + 1) & 1]);
#endif

        const int qx = t.tx0 + r, qy = t.ty0 + wrp;
        const bool ract = qx < lv.w[t.lq] && qy < lv.h[t.lq];
        const long long q = ract ? (lv.start[t.lq] + (long long)qy * lv.w[t.lq] + qx) : 0;
        const bool qok = ract && q < p.Nq;
        const long long row = qok ? (((long long)t.b * p.Nq + q) * p.H + t.head) : 0;
        const float* loc_row = loc + row * p.LP * 2;
        const float* w_row = w0 + row * p.LP;
        const size_t ppitch = (size_t)p.H * ROWB;                       // value pixel pitch in bytes
        const size_t gppitch = (size_t)p.H * p.D;                       // grad_value pixel pitch in elements
        const size_t img_first = ((size_t)t.b * p.S * p.H + t.head);    // (pixel 0, head) row of the image
        const unsigned char* vimg = static_cast<const unsigned char*>(p.value) + img_first * ROWB;
        ACC* gimg = static_cast<ACC*>(p.grad_value_acc) + img_first * p.D + j * 8;

        float go[8];
        {
            const unsigned char* gp = static_cast<const unsigned char*>(p.grad_out) + (size_t)row * ROWB + j * LANEB;
            const uint4 g0 = __ldg(reinterpret_cast<const uint4*>(gp + rpar * 16));
            TL::unpack(g0, go);
            if (NV == 2) {
                const uint4 g1 = __ldg(reinterpret_cast<const uint4*>(gp + (rpar ^ 1) * 16));
                TL::unpack(g1, go + 4);
            }
            if (!qok) {
#pragma unroll
                for (int i = 0; i < 8; ++i) go[i] = 0.f;
            }
        }

        for (int l = 0; l < p.L; ++l) {
            const int lh = lv.h[l], lw = lv.w[l];
#if BXR_TILE_PREFETCH
            if (l + 1 < p.L && qok && j < 2) {
                const void* nxt = j == 0 ? static_cast<const void*>(loc_row + (l + 1) * p.P * 2) : static_cast<const void*>(w_row + (l + 1) * p.P);
                if (BXR_TILE_PREFETCH == 1) prefetch_l1(nxt); else prefetch_l2(nxt);
            }
#endif
            LanePoint pt[PPL];
            float g_a[PPL], g_x[PPL], g_y[PPL];
            int bx0 = kNoPix, bx1 = -kNoPix, by0 = kNoPix, by1 = -kNoPix;
            float S = 0.f;
#pragma unroll
            for (int k = 0; k < PPL; ++k) {
                const int ptn = qok ? j + k * kTG : p.P;
                pt[k] = lane_point(loc_row + l * p.P * 2, w_row + l * p.P, ptn, p.P, lh, lw);
                g_a[k] = g_x[k] = g_y[k] = 0.f;
                if (pt[k].inside) {
                    bx0 = min(bx0, pt[k].x0); bx1 = max(bx1, pt[k].x0 + 1);
                    by0 = min(by0, pt[k].y0); by1 = max(by1, pt[k].y0 + 1);
                    S += fabsf(pt[k].aw);
                }
            }
            const int X0 = max(q4min(bx0), 0), Y0 = max(q4min(by0), 0);
            const int nx = min(q4max(bx1), lw - 1) - X0 + 1, ny = min(q4max(by1), lh - 1) - Y0 + 1;
            S = q4sum(S);
            // a touched pixel needs its d even when every weight is zero: S == 0 stays in window mode here
            const int mode = (nx <= 0 || ny <= 0) ? 0 : ((nx <= 8 && ny <= 8 && S <= 3.0e38f) ? 1 : 2);
            const int ke = fixed_scale_exp(fmaxf(S, 1e-30f));
#if BXR_TILE_COPY == 2
            TileRegion g;
            g.x0 = g.y0 = g.w = g.h = g.off = 0;
            const bool staged = false;
#else
            const TileRegion g = t.reg[l];
            const bool staged = mode == 1 && g.w > 0 && X0 >= g.x0 && Y0 >= g.y0 && X0 + nx <= g.x0 + g.w && Y0 + ny <= g.y0 + g.h;
#endif

            if (mode == 1) {
#pragma unroll
                for (int z = 0; z < 4; ++z) *reinterpret_cast<uint4*>(win + (z * kTG + j) * 4) = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            if (mode == 1) {
                const float scale = pow2f(ke);
#pragma unroll
                for (int k = 0; k < PPL; ++k)
                    if (pt[k].inside) tile_scatter(win, pt[k], scale, X0, Y0, nx, ny);
            }
            __syncwarp();

            const unsigned m_st = __ballot_sync(kFullMask, staged);
            const unsigned m_gl = __ballot_sync(kFullMask, mode == 1 && !staged);
            const unsigned m_pp = __ballot_sync(kFullMask, mode == 2);
            const float inv_scale = pow2f(-ke);
            const size_t lpix0 = (size_t)lv.start[l] + (mode == 1 ? (size_t)Y0 * lw + X0 : 0);
            if (m_st) {
#if BXR_TILE_COPY == 0
                if (!landed) { mbar_wait(&s_bar, parity); landed = true; }
#endif
                const unsigned char* vp = s_val + ((size_t)g.off + (staged ? (Y0 - g.y0) * g.w + (X0 - g.x0) : 0)) * ROWB + j * LANEB;
                tile_bwd_walk<TV, ACC, true>(staged, nx, ny, win, dwin, vp, (size_t)g.w * ROWB, ROWB, gimg + lpix0 * gppitch,
                                             (size_t)lw * gppitch, gppitch, rpar, j, go, inv_scale, dscale);
            }
            if (m_gl) {
                const bool mine = mode == 1 && !staged;
                tile_bwd_walk<TV, ACC, false>(mine, nx, ny, win, dwin, vimg + lpix0 * ppitch + j * LANEB, (size_t)lw * ppitch, ppitch,
                                              gimg + lpix0 * gppitch, (size_t)lw * gppitch, gppitch, rpar, j, go, inv_scale, dscale);
            }
            if (m_pp) {
                const unsigned char* vl = vimg + (size_t)lv.start[l] * ppitch + j * LANEB;
                ACC* gl = gimg + (size_t)lv.start[l] * gppitch;
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
#pragma unroll
                    for (int o = 0; o < kTG; ++o) {
                        if (o + k * kTG >= p.P) break;
                        const bool inside = __shfl_sync(kFullMask, (int)pt[k].inside, o, kTG) != 0;
                        const int x0 = __shfl_sync(kFullMask, pt[k].x0, o, kTG), y0 = __shfl_sync(kFullMask, pt[k].y0, o, kTG);
                        const float lx = __shfl_sync(kFullMask, pt[k].lx, o, kTG), ly = __shfl_sync(kFullMask, pt[k].ly, o, kTG);
                        const float aw = __shfl_sync(kFullMask, pt[k].aw, o, kTG);
                        const bool act = mode == 2 && inside;
                        const float hx = 1.f - lx, hy = 1.f - ly;
                        const float cw[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
                        float d[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            float tt = 0.f;
                            const bool okc = act && ((c & 1) ? (x0 + 1 <= lw - 1) : (x0 >= 0)) && ((c & 2) ? (y0 + 1 <= lh - 1) : (y0 >= 0));
                            if (okc) {
                                const long long pix = (long long)(y0 + ((c & 2) ? 1 : 0)) * lw + x0 + ((c & 1) ? 1 : 0);
                                const unsigned char* cp = vl + pix * (long long)ppitch;
                                const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(cp + rpar * 16));
                                tt = TL::dot(v0, go);
                                if (NV == 2) {
                                    const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(cp + (rpar ^ 1) * 16));
                                    tt += TL::dot(v1, go + 4);
                                }
                                tile_scatter_grad<ACC, NV>(gl + pix * (long long)gppitch, rpar, go, cw[c] * aw, dscale);
                            }
                            d[c] = tt;
                        }
                        float total;
                        const int idx = reduce4<4>(d, total, j, kFullMask);
                        // every lane now holds the row total of corner `idx`; collect the four on the owner lane
                        const float d0 = __shfl_sync(kFullMask, total, 0, kTG), d1 = __shfl_sync(kFullMask, total, 1, kTG);
                        const float d2 = __shfl_sync(kFullMask, total, 2, kTG), d3 = __shfl_sync(kFullMask, total, 3, kTG);
                        (void)idx;
                        if (act && j == o) {
                            g_a[k] = cw[0] * d0 + cw[1] * d1 + cw[2] * d2 + cw[3] * d3;
                            g_x[k] = (float)lw * aw * (hy * (d1 - d0) + ly * (d3 - d2));
                            g_y[k] = (float)lh * aw * (hx * (d2 - d0) + lx * (d3 - d1));
                        }
                    }
                }
            }
            __syncwarp();
            // D: own points from the d window
            if (mode == 1) {
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    if (pt[k].inside) {
                        const int sx = pt[k].x0 - X0, sy = pt[k].y0 - Y0;
                        const float lx = pt[k].lx, ly = pt[k].ly, hx = 1.f - lx, hy = 1.f - ly;
                        const bool vx0 = sx >= 0, vx1 = sx + 1 < nx, vy0 = sy >= 0, vy1 = sy + 1 < ny;
                        const float* dp = dwin + sy * 8 + sx;
                        const float d00 = (vy0 && vx0) ? dp[0] : 0.f;
                        const float d01 = (vy0 && vx1) ? dp[1] : 0.f;
                        const float d10 = (vy1 && vx0) ? dp[8] : 0.f;
                        const float d11 = (vy1 && vx1) ? dp[9] : 0.f;
                        g_a[k] = hy * hx * d00 + hy * lx * d01 + ly * hx * d10 + ly * lx * d11;
                        g_x[k] = (float)lw * pt[k].aw * (hy * (d01 - d00) + ly * (d11 - d10));
                        g_y[k] = (float)lh * pt[k].aw * (hx * (d10 - d00) + lx * (d11 - d01));
                    }
                }
            }
            __syncwarp();
            if (qok) {
#pragma unroll
                for (int k = 0; k < PPL; ++k) {
                    const int ptn = j + k * kTG;
                    if (ptn < p.P) {
                        const long long sidx = row * p.LP + (long long)l * p.P + ptn;
                        grad_w0[sidx] = g_a[k];
                        reinterpret_cast<float2*>(grad_loc)[sidx] = make_float2(g_x[k], g_y[k]);
                    }
                }
            }
        }
#if BXR_TILE_COPY == 0
        if (!landed) mbar_wait(&s_bar, parity);
        parity ^= 1u;
#endif
#if BXR_TILE_COPY != 2
        __syncthreads();
#endif
        (void)landed;
    }
}

}  // namespace bxr
