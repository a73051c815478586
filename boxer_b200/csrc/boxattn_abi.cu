// Host side of the C ABI declared in include/boxattn_b200.h: argument checks, path selection,
// launch geometry.  Pure CUDA runtime -- no torch / ATen here.
//
// Replaces the reference's host functions (e2edet/module/ops/src/box_attn/box_attn.cu:15-135,
// instance_attn/instance_attn.cu:15-157): same inputs, but the caller owns all memory, there is
// no im2col_step chunk loop (every image goes in one launch), and failures are returned.
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "boxattn_kernels.cuh"
#include "boxattn_window.cuh"
#include "boxattn_instance.cuh"
#include "boxattn_fused.cuh"
#include "boxattn_staged.cuh"
#include "boxattn_tile.cuh"
#include "../../include/boxattn_b200.h"

// Build-time slicing: the same source can be compiled once per (dtype, direction) slice, in parallel,
// and the objects linked into one library (boxer_b200/_native.py); with no -D it is one complete TU.
#ifndef BXR_TU_DTYPES
#define BXR_TU_DTYPES 7     // bit set: 1 = f32, 2 = f64, 4 = bf16
#endif
#ifndef BXR_TU_DIRS
#define BXR_TU_DIRS 3       // bit set: 1 = forward entry points, 2 = backward entry points
#endif
#ifndef BXR_TU_COMMON
#define BXR_TU_COMMON 1     // the untyped entry points and the per-thread error state live in one slice
#endif

namespace bxr_host {
// per-thread state behind bxr_last_error_detail() / bxr_last_launch_count(); shared by all slices
extern thread_local char g_detail[256];
extern thread_local int g_launches;
#if BXR_TU_COMMON
thread_local char g_detail[256] = "";
thread_local int g_launches = 0;
#endif
}  // namespace bxr_host

// nvcc gives functions of an unnamed namespace names derived from the file name, which collide between
// slices of the same file: every slice gets its own named namespace instead
#ifndef BXR_TU_NAME
#define BXR_TU_NAME all
#endif
#define BXR_CAT2(a, b) a##b
#define BXR_CAT(a, b) BXR_CAT2(a, b)
#define BXR_SLICE_NS BXR_CAT(bxr_slice_, BXR_TU_NAME)

namespace BXR_SLICE_NS {

using namespace bxr;
using bxr_host::g_detail;
using bxr_host::g_launches;

int fail(int status, const char* what) {
    snprintf(g_detail, sizeof(g_detail), "%s", what);
    return status;
}

int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_detail, sizeof(g_detail), "%s: %s", where, cudaGetErrorString(e));
    return BXR_ERR_CUDA;
}

#define BXR_CUDA(expr)                                        \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
    } while (0)

int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

#ifndef BXR_INST_TAB_BWD
#define BXR_INST_TAB_BWD 0
#endif
#ifndef BXR_TILE_DEFAULT
#define BXR_TILE_DEFAULT 0
#endif
#ifndef BXR_TRIM_CARVEOUT
#define BXR_TRIM_CARVEOUT 1
#endif

template <void (*K)(const AttnParams), int THREADS = kThreads>
int ctas_per_sm() {
    // per device: the occupancy answer and the carve-out hint belong to the (kernel, device) pair -- a second GPU
    // driven from the same process gets its own query and its own hint.  Entries are immutable once set; a benign
    // race only recomputes the same number.
    static int occ[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (occ[dev] == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, K, THREADS, 0) != cudaSuccess || n <= 0) n = 1;
#if BXR_TRIM_CARVEOUT
        // The gathers live on L1 hits; left alone the driver carves out far more shared memory than the resident
        // CTAs use (ncu r01s: 135 KB configured for 4 x 19 KB in the window backward, L1 hit rate 22 %).  Ask for
        // just what the register-limited residency needs (a hint; the driver rounds up to a supported split).
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, K) == cudaSuccess) {
            const size_t need = (size_t)n * (fa.sharedSizeBytes + 1024);
            int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            if (pct > 100) pct = 100;
            cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
#endif
        occ[dev] = n;
    }
    return occ[dev];
}

template <void (*K)(const AttnParams), int THREADS = kThreads>
int launch(const AttnParams& p, int grid, cudaStream_t st, const char* name) {
    K<<<grid, THREADS, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    ++g_launches;
    return BXR_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

int check_dims(int B, int S, int H, int D, int L, int Nq, int P) {
    if (B < 0 || S < 0 || H < 0 || D < 0 || L < 0 || Nq < 0 || P < 0) return fail(BXR_ERR_BAD_DIM, "negative dimension");
    if (L > BXR_MAX_LEVELS) return fail(BXR_ERR_BAD_DIM, "L exceeds BXR_MAX_LEVELS");
    if ((long long)L * P > 0x7fffffffLL / 4) return fail(BXR_ERR_BAD_DIM, "L*P too large");
    return BXR_OK;
}

// lanes per row of the vector path, or 0 when it does not apply
template <typename TV>
int vec_group(int D, int LP) {
    if (std::is_same<TV, double>::value) return 0;
    const int vec = 16 / (int)sizeof(TV);
    if (D <= 0 || D % vec) return 0;
    const int g = D / vec;
    if (g > 32 || (g & (g - 1))) return 0;
    if (LP >= 65536) return 0;
    return g;
}

// Split a row's points over several lane groups when there are too few rows to fill the GPU
// (decoder calls: 300 queries x 8 heads; the mask head has 4*196 points per row).
void choose_split(AttnParams& p, int G, int lim, int min_chunk) {
    const int groups = kThreads / G;
    const long long want_units = 4LL * sm_count();
    int k = 0;
    while ((1 << (k + 1)) <= groups && (lim >> (k + 1)) >= min_chunk &&
           (p.rows << k) / groups < want_units)
        ++k;
    p.nsplit_log2 = k;
    const int ns = 1 << k;
    p.chunk = (lim + ns - 1) / ns;
    const int rows_per_unit = groups >> k;
    p.units = (int)((p.rows + rows_per_unit - 1) / rows_per_unit);
}

template <void (*K)(const AttnParams), int THREADS = kThreads>
int launch_units(const AttnParams& p, cudaStream_t st, const char* name) {
    int grid = sm_count() * ctas_per_sm<K, THREADS>();
    if (grid > p.units) grid = p.units;
    if (grid < 1) grid = 1;
    return launch<K, THREADS>(p, grid, st, name);
}

template <void (*K)(const AttnParams)>
int launch_rows(const AttnParams& p, cudaStream_t st, const char* name) {
    const long long warps = kThreads / 32;
    long long grid = (p.rows + warps - 1) / warps;
    const long long cap = (long long)sm_count() * ctas_per_sm<K>() * 4;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    return launch<K>(p, (int)grid, st, name);
}

template <typename TV, bool INSTANCE, int G>
int fwd_vec(AttnParams& p, cudaStream_t st) {
    constexpr int U = (sizeof(TV) == 2) ? 2 : 4;
    choose_split(p, G, INSTANCE ? p.P : p.LP, INSTANCE ? 2 : 2 * U);
    if (!INSTANCE) p.chunk = (p.chunk + U - 1) / U * U;   // whole batches; the tail is masked by j < j1
    return launch_units<attn_fwd_vec_kernel<TV, G, INSTANCE, U>>(p, st, "attn_fwd_vec_kernel");
}

template <typename TV, bool INSTANCE, typename ACC, int G>
int bwd_vec(AttnParams& p, cudaStream_t st) {
    choose_split(p, G, INSTANCE ? p.P : p.LP, INSTANCE ? 2 : 4);
    return launch_units<attn_bwd_vec_kernel<TV, G, INSTANCE, ACC>>(p, st, "attn_bwd_vec_kernel");
}

#define BXR_DISPATCH_G(G_, CALL)                                  \
    switch (G_) {                                                 \
        case 1: { constexpr int G = 1; return CALL; }             \
        case 2: { constexpr int G = 2; return CALL; }             \
        case 4: { constexpr int G = 4; return CALL; }             \
        case 8: { constexpr int G = 8; return CALL; }             \
        case 16: { constexpr int G = 16; return CALL; }           \
        default: { constexpr int G = 32; return CALL; }           \
    }

template <typename TV, bool INSTANCE>
int dispatch_fwd_vec(int g, AttnParams& p, cudaStream_t st) {
    BXR_DISPATCH_G(g, (fwd_vec<TV, INSTANCE, G>(p, st)))
}
template <typename TV, bool INSTANCE, typename ACC>
int dispatch_bwd_vec(int g, AttnParams& p, cudaStream_t st) {
    BXR_DISPATCH_G(g, (bwd_vec<TV, INSTANCE, ACC, G>(p, st)))
}

// bf16: lanes of 8 bytes (4 channels) when that gives the fp32 kernels' geometry (head_dim 32 -> G = 8)
int bf16_lane8_group(int D) {
    if (D % 8 == 0 && D / 8 == 8) return 0;        // head_dim 64: 16-byte lanes already give G = 8
    if (D % 4) return 0;
    const int g = D / 4;
    return (g == 8 || g == 4 || g == 16) ? g : 0;
}

// ---- footprint-window kernels (boxattn_window.cuh): box op, enough rows to fill the GPU, P <= 4*G
bool use_window(const AttnParams& p, int g, unsigned flags) {
    if (flags & BXR_FLAG_PATH_POINT) return false;
    if (g != 4 && g != 8 && g != 16) return false;
    if (p.P > 4 * g || p.P < 1) return false;
    // the window kernels address value in 16-byte units with 32-bit indices
    if ((long long)p.B * p.S * g * p.H >= 0xffffffffLL) return false;
    if (flags & BXR_FLAG_PATH_WINDOW) return true;
    if (p.P < 3) return false;
    const long long groups = kThreads / g;
    return p.rows >= 2LL * sm_count() * groups;     // small (decoder-sized) calls keep the point-split kernels
}

// Tile-ordered work units (boxattn_window.cuh, TileOrder) for self-attention-shaped calls: Nq == S, rows of 8 lanes
// (head_dim 32 fp32 / bf16 with 8-byte lanes), location-taking op.  The unit count then depends on the level shapes,
// which live on the device: the kernel derives it, the host only sizes the persistent grid.
// MEASURED AND NOT ADOPTED (B200, r02k, profiles/README.md): the L1 sector hit rate of the forward goes from 29 % to 72 %
// as intended, and nothing else moves -- long-scoreboard stalls 3.08 -> 2.72 cycles per issue, 137 M instead of 133 M
// instructions, forward 0.178 -> 0.191 ms, backward 0.350 -> 0.353 ms, uniform / trained-like 10-17 % slower (the rows of
// a unit are 4 KB apart in loc / weights instead of contiguous).  Where the gathered rows come from is not what these
// kernels wait for.  BXR_WIN_TILED=1 compiles the variant (selected unless BXR_FLAG_NO_TILE_ORDER is given).
#ifndef BXR_WIN_TILED
#define BXR_WIN_TILED 0
#endif
bool use_tiled_units(const AttnParams& p, int G, unsigned flags) {
#if BXR_WIN_TILED
    if (flags & BXR_FLAG_NO_TILE_ORDER) return false;
    if (G != 8 || p.Nq != p.S || (long long)p.B * p.S * p.H >= 0x7fffffffLL) return false;
    return (flags & BXR_FLAG_PATH_WINDOW) || p.rows >= 2LL * sm_count() * (kThreads / G);
#else
    (void)p; (void)G; (void)flags;
    return false;
#endif
}

template <typename TV, int G, int SUB, int PPL, bool FUSED, bool SMAX>
int fwd_win(AttnParams& p, cudaStream_t st, unsigned flags = 0) {
    p.units = (int)((p.rows + kFwdThreads / G - 1) / (kFwdThreads / G));
#if BXR_WIN_TILED
    if constexpr (G == 8 && !FUSED && !SMAX) {
        if (use_tiled_units(p, G, flags))
            return launch_units<box_fwd_win_kernel<TV, G, SUB, PPL, FUSED, SMAX, true>, kFwdThreads>(p, st, "box_fwd_win_kernel");
    }
#endif
    (void)flags;
    return launch_units<box_fwd_win_kernel<TV, G, SUB, PPL, FUSED, SMAX>, kFwdThreads>(p, st, "box_fwd_win_kernel");
}
template <typename TV, int G, int SUB, int PPL, typename ACC, bool FUSED, bool SMAX>
int bwd_win(AttnParams& p, cudaStream_t st, unsigned flags = 0) {
    p.units = (int)((p.rows + kBwdThreads / G - 1) / (kBwdThreads / G));
#if BXR_WIN_TILED
    if constexpr (G == 8 && !FUSED && !SMAX) {
        if (use_tiled_units(p, G, flags))
            return launch_units<box_bwd_win_kernel<TV, G, SUB, PPL, ACC, FUSED, SMAX, true>, kBwdThreads>(p, st, "box_bwd_win_kernel");
    }
#endif
    (void)flags;
    return launch_units<box_bwd_win_kernel<TV, G, SUB, PPL, ACC, FUSED, SMAX>, kBwdThreads>(p, st, "box_bwd_win_kernel");
}

// (G, SUB, PPL): SUB lanes share a level's points, PPL points per lane; P <= SUB * PPL.
// P <= G/2 packs two levels per pass (SUB = G/2) so that 2x2 grids keep all lanes busy.
#define BXR_WIN_CASE(G_, SUB_, PPL_, CALL) \
    case (G_) * 1000 + (SUB_) * 10 + (PPL_): { constexpr int G = G_, SUB = SUB_, PPL = PPL_; return CALL; }
#define BXR_DISPATCH_WIN(KEY, CALL)                                                        \
    switch (KEY) {                                                                         \
        BXR_WIN_CASE(4, 4, 1, CALL) BXR_WIN_CASE(4, 4, 2, CALL) BXR_WIN_CASE(4, 4, 4, CALL) \
        BXR_WIN_CASE(8, 4, 1, CALL)                                                        \
        BXR_WIN_CASE(8, 8, 1, CALL) BXR_WIN_CASE(8, 8, 2, CALL) BXR_WIN_CASE(8, 8, 4, CALL) \
        BXR_WIN_CASE(16, 8, 1, CALL)                                                       \
        BXR_WIN_CASE(16, 16, 1, CALL) BXR_WIN_CASE(16, 16, 2, CALL)                        \
        default: { constexpr int G = 16, SUB = 16, PPL = 4; return CALL; }                 \
    }

int win_key(int P, int g) {
    int sub = g;
    if (g >= 8 && P <= g / 2) sub = g / 2;
    const int n = (P + sub - 1) / sub;
    const int ppl = n <= 1 ? 1 : (n <= 2 ? 2 : 4);
    return g * 1000 + sub * 10 + ppl;
}

template <typename TV, bool FUSED = false, bool SMAX = false>
int dispatch_fwd_win(int g, AttnParams& p, cudaStream_t st, unsigned flags = 0) {
    BXR_DISPATCH_WIN(win_key(p.P, g), (fwd_win<TV, G, SUB, PPL, FUSED, SMAX>(p, st, flags)))
}
template <typename TV, typename ACC, bool FUSED = false, bool SMAX = false>
int dispatch_bwd_win(int g, AttnParams& p, cudaStream_t st, unsigned flags = 0) {
    BXR_DISPATCH_WIN(win_key(p.P, g), (bwd_win<TV, G, SUB, PPL, ACC, FUSED, SMAX>(p, st, flags)))
}

// ---- staged-row kernels (boxattn_staged.cuh): the window algorithm with TMA-staged row operands.
// Needs 16-byte granular rows (LP % 4 == 0) and operand pointers, and a shared-memory plan that
// leaves at least two CTAs per SM.
template <void (*K)(const AttnParams)>
int launch_units_smem(const AttnParams& p, size_t smem, cudaStream_t st, const char* name) {
    thread_local int cfg_dev = -1, cfg_occ = 0;
    thread_local size_t cfg_smem = 0;
    int dev = 0;
    BXR_CUDA(cudaGetDevice(&dev));
    if (dev != cfg_dev || smem != cfg_smem) {
        BXR_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n = 0;
        BXR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, K, kThreads, smem));
        if (n <= 0) return fail(BXR_ERR_UNSUPPORTED, "staged kernel does not fit an SM");
        cfg_dev = dev; cfg_smem = smem; cfg_occ = n;
    }
    int grid = sm_count() * cfg_occ;
    if (grid > p.units) grid = p.units;
    if (grid < 1) grid = 1;
    K<<<grid, kThreads, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    ++g_launches;
    return BXR_OK;
}

constexpr size_t kStgMaxSmem = 100 * 1024;

bool use_staged(const AttnParams& p, int g, int mode, int stages, unsigned flags) {
    // opt-in: measured on B200 (profiles/r01n_*) the staged forward equals the plain window kernel on the
    // box-structured encoder workload (0.193 vs 0.195 ms) and is slower on uniform locations / 2x2 grids
    if (!(flags & BXR_FLAG_STAGED) || (flags & BXR_FLAG_PATH_POINT)) return false;
    if (p.LP % 4) return false;
    if (!aligned16(p.w0) || !aligned16(mode == 0 ? p.loc : p.boxes)) return false;
    return stg_plan(g, mode, p.L, p.LP, stages).total_bytes <= kStgMaxSmem;
}

template <typename TV, int G, int SUB, int PPL, int MODE>
int fwd_stg(AttnParams& p, cudaStream_t st) {
    p.units = (int)((p.rows + kThreads / G - 1) / (kThreads / G));
    return launch_units_smem<box_fwd_stg_kernel<TV, G, SUB, PPL, MODE>>(p, stg_plan(G, MODE, p.L, p.LP, 1).total_bytes, st,
                                                                         "box_fwd_stg_kernel");
}
template <typename TV, int MODE>
int dispatch_fwd_stg(int g, AttnParams& p, cudaStream_t st) {
    BXR_DISPATCH_WIN(win_key(p.P, g), (fwd_stg<TV, G, SUB, PPL, MODE>(p, st)))
}

// ---- query-tile x value-tile kernels (boxattn_tile.cuh): self-attention-shaped calls (Nq == S: the queries are the
// pixels), head_dim 32, at most 16 points per level, enough rows to fill the GPU
template <typename TV>
bool use_tile(const AttnParams& p, unsigned flags) {
    if (std::is_same<TV, double>::value) return false;
    if (flags & (BXR_FLAG_PATH_POINT | BXR_FLAG_PATH_WINDOW | BXR_FLAG_STAGED)) return false;
    if (p.Nq != p.S || p.D != 32 || p.P < 1 || p.P > 16) return false;
    if ((long long)p.B * p.S * p.H >= 0x7fffffffLL) return false;      // the kernels count (image, tile, head) items in 32 bits
    if (flags & BXR_FLAG_PATH_TILE) return true;
#if BXR_TILE_DEFAULT
    return p.rows >= 16LL * sm_count() * kTileRows;
#else
    return false;
#endif
}

template <void (*K)(const AttnParams)>
int launch_tile(const AttnParams& p, size_t smem, cudaStream_t st, const char* name) {
    static int occ[64] = {0};      // per device; immutable once set
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (occ[dev] == 0) {
        BXR_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n = 0;
        BXR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, K, kTileThreads, smem));
        if (n <= 0) return fail(BXR_ERR_UNSUPPORTED, "tile kernel does not fit an SM");
        // what is not shared memory is L1, which the unstaged gathers live on: ask for no more than the resident CTAs use
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, K) == cudaSuccess) {
            const size_t need = (size_t)n * (smem + fa.sharedSizeBytes + 1024);
            int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
            cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
        }
        occ[dev] = n;
    }
    const int grid = sm_count() * occ[dev];
    K<<<grid, kTileThreads, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, name);
    ++g_launches;
    return BXR_OK;
}

template <typename TV, typename ACC>
int bwd_tile(AttnParams& p, cudaStream_t st) {
    const size_t smem = tile_smem_bytes(TileLane<TV>::ROWB, true);
    const int ppl = (p.P + kTG - 1) / kTG;
    if (ppl <= 1) return launch_tile<box_bwd_tile_kernel<TV, 1, ACC>>(p, smem, st, "box_bwd_tile_kernel");
    if (ppl <= 2) return launch_tile<box_bwd_tile_kernel<TV, 2, ACC>>(p, smem, st, "box_bwd_tile_kernel");
    return launch_tile<box_bwd_tile_kernel<TV, 4, ACC>>(p, smem, st, "box_bwd_tile_kernel");
}

template <typename TV>
int fwd_tile(AttnParams& p, cudaStream_t st) {
    const size_t smem = tile_smem_bytes(TileLane<TV>::ROWB, false);
    const int ppl = (p.P + kTG - 1) / kTG;
    if (ppl <= 1) return launch_tile<box_fwd_tile_kernel<TV, 1>>(p, smem, st, "box_fwd_tile_kernel");
    if (ppl <= 2) return launch_tile<box_fwd_tile_kernel<TV, 2>>(p, smem, st, "box_fwd_tile_kernel");
    return launch_tile<box_fwd_tile_kernel<TV, 4>>(p, smem, st, "box_fwd_tile_kernel");
}

// the fused (box -> grid) window kernels apply whenever the vector layout does; no row-count threshold
bool fused_applies(int dtype_bytes, int B, int S, int H, int D, int L, int P, unsigned flags) {
    if (flags & BXR_FLAG_PATH_POINT) return false;
    if (dtype_bytes != 4 && dtype_bytes != 2) return false;
    const int vec = 16 / dtype_bytes;
    int g = (D > 0 && D % vec == 0) ? D / vec : 0;
    if (dtype_bytes == 2 && bf16_lane8_group(D)) g = bf16_lane8_group(D);
    if (g != 4 && g != 8 && g != 16) return false;
    if (P < 1 || P > 4 * g || (long long)L * P >= 65536) return false;
    if ((long long)B * S * g * H >= 0xffffffffLL) return false;
    return true;
}

#ifndef BXR_INST_BWD_BF16X4
#define BXR_INST_BWD_BF16X4 1
#endif
#ifndef BXR_INST_FWD_BF16X4
#define BXR_INST_FWD_BF16X4 0
#endif

// ---- owner-tap instance kernels (boxattn_instance.cuh)
constexpr int kInstLevels = 4;
bool use_inst_own(const AttnParams& p, int g, unsigned flags) {
    if (flags & BXR_FLAG_PATH_POINT) return false;
    if (g != 4 && g != 8) return false;
    if (p.L > kInstLevels || p.P < 1) return false;
    if ((long long)p.B * p.S * g * p.H >= 0xffffffffLL) return false;
    return true;
}

// split a row's point chunks (G points each) over 2^k groups of a CTA until the GPU is filled
void choose_chunk_split(AttnParams& p, int G) {
    const int groups = kThreads / G;
    const int nchunks = (p.P + G - 1) / G;
    const long long want_units = 4LL * sm_count();
    int k = 0;
    while ((1 << (k + 1)) <= groups && (1 << (k + 1)) <= nchunks && (p.rows << k) / groups < want_units) ++k;
    p.nsplit_log2 = k;
    const int rows_per_unit = groups >> k;
    p.units = (int)((p.rows + rows_per_unit - 1) / rows_per_unit);
}

template <typename TV, int G>
int fwd_inst_own(AttnParams& p, cudaStream_t st) {
    choose_chunk_split(p, G);
    // taps through the shared-memory table for fp32 (r02q: K=14 0.118 -> 0.095 ms, K=28 0.427 -> 0.344); the bf16 forward
    // (8 channels per lane: FMA-bound) measured 2-5 % slower with it and keeps the shuffle broadcast
#if BXR_INST_TAB
    if constexpr (sizeof(TV) == 4) return launch_units<inst_fwd_tab_kernel<TV, G, kInstLevels>>(p, st, "inst_fwd_tab_kernel");
#endif
    return launch_units<inst_fwd_own_kernel<TV, G, kInstLevels>>(p, st, "inst_fwd_own_kernel");
}
template <typename TV, int G, typename ACC>
int bwd_inst_own(AttnParams& p, cudaStream_t st) {
    choose_chunk_split(p, G);
    // the table / two-dots-per-corner backward measured SLOWER than the owner-tap one (r02q: fp32 K=14 0.289 -> 0.297 ms,
    // bf16 0.284 -> 0.310): it is compiled only with BXR_INST_TAB_BWD=1
#if BXR_INST_TAB_BWD
    return launch_units<inst_bwd_tab_kernel<TV, G, kInstLevels, ACC>>(p, st, "inst_bwd_tab_kernel");
#else
    return launch_units<inst_bwd_own_kernel<TV, G, kInstLevels, ACC>>(p, st, "inst_bwd_own_kernel");
#endif
}

void fill_sizes(AttnParams& p, int B, int S, int H, int D, int L, int Nq, int P) {
    p.B = B; p.S = S; p.H = H; p.D = D; p.L = L; p.Nq = Nq; p.P = P;
    p.LP = L * P;
    p.magicP = P > 1 ? (unsigned)((0x100000000ULL + (unsigned)P - 1) / (unsigned)P) : 0u;
    p.rows = (long long)B * Nq * H;
    p.nsplit_log2 = 0;
    p.chunk = p.LP;
    p.units = 0;
}

template <typename TV, typename TW, bool INSTANCE>
int forward(const TV* value, const int64_t* shapes, const int64_t* level_start, const TW* loc, const TW* w0, const TW* w1,
            int B, int S, int H, int D, int L, int Nq, int P, TV* out, TV* mask_out, unsigned flags, bxr_stream_t stream) {
    g_launches = 0;
    g_detail[0] = 0;
    if (int s = check_dims(B, S, H, D, L, Nq, P)) return s;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    fill_sizes(p, B, S, H, D, L, Nq, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p.rows == 0 || D == 0) return BXR_OK;
    if (!out || (INSTANCE && !mask_out && P > 0)) return fail(BXR_ERR_NULL_POINTER, "output pointer is NULL");
    if (p.LP == 0 || S == 0) {   // nothing to sample: outputs are all zero
        BXR_CUDA(cudaMemsetAsync(out, 0, sizeof(TV) * (size_t)p.rows * D, st));
        if (INSTANCE && P > 0) BXR_CUDA(cudaMemsetAsync(mask_out, 0, sizeof(TV) * (size_t)p.rows * D * P, st));
        return BXR_OK;
    }
    if (!value || !shapes || !level_start || !loc || !w0 || (INSTANCE && !w1))
        return fail(BXR_ERR_NULL_POINTER, "input pointer is NULL");
    p.value = value; p.shapes = shapes; p.level_start = level_start;
    p.loc = loc; p.w0 = w0; p.w1 = w1; p.out = out; p.mask_out = mask_out;

    const int g = vec_group<TV>(D, p.LP);
    if (g && aligned16(value) && aligned16(out) && (!INSTANCE || aligned16(mask_out)) && aligned8(loc)) {
        if constexpr (!std::is_same<TV, double>::value) {
            if constexpr (!INSTANCE) {
                if (use_tile<TV>(p, flags)) return fwd_tile<TV>(p, st);
            }
            if constexpr (std::is_same<TV, __nv_bfloat16>::value) {
                // (measured, r01j: the instance kernels are faster with 16-byte lanes, so only the box op switches)
                if constexpr (!INSTANCE) {
                    if (const int g8 = bf16_lane8_group(D)) {
                        if (use_window(p, g8, flags))
                            return use_staged(p, g8, 0, 1, flags) ? dispatch_fwd_stg<bf16x4_t, 0>(g8, p, st)
                                                                  : dispatch_fwd_win<bf16x4_t>(g8, p, st, flags);
                    }
                }
            }
            if constexpr (!INSTANCE) {
                if (use_window(p, g, flags))
                    return use_staged(p, g, 0, 1, flags) ? dispatch_fwd_stg<TV, 0>(g, p, st) : dispatch_fwd_win<TV>(g, p, st, flags);
            } else {
#if BXR_INST_FWD_BF16X4
                if constexpr (std::is_same<TV, __nv_bfloat16>::value) {
                    if (bf16_lane8_group(D) == 8 && use_inst_own(p, 8, flags)) return fwd_inst_own<bf16x4_t, 8>(p, st);
                }
#endif
                if (use_inst_own(p, g, flags)) return g == 8 ? fwd_inst_own<TV, 8>(p, st) : fwd_inst_own<TV, 4>(p, st);
            }
            return dispatch_fwd_vec<TV, INSTANCE>(g, p, st);
        }
    }
    return launch_rows<attn_fwd_gen_kernel<TV, INSTANCE>>(p, st, "attn_fwd_gen_kernel");
}

constexpr size_t kWsHeader = 256;   // absmax bits[4] + scale, kept apart from the accumulator

size_t workspace_bytes(int dtype_bytes, long long n, unsigned flags) {
    if (n <= 0) return 0;
    if (flags & BXR_FLAG_DETERMINISTIC) return kWsHeader + 8 * (size_t)n;
    if (dtype_bytes == 2) return 4 * (size_t)n;
    return 0;
}

template <typename T>
int absmax(const T* x, long long n, unsigned* bits, cudaStream_t st) {
    if (n <= 0) return BXR_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
    absmax_kernel<T><<<(int)blocks, 256, 0, st>>>(x, n, bits);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "absmax_kernel");
    ++g_launches;
    return BXR_OK;
}

template <typename TV, typename ACC>
int finalize(const ACC* acc, TV* out, long long n, const float* scale, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    finalize_grad_value_kernel<TV, ACC><<<(int)blocks, 256, 0, st>>>(acc, out, n, scale);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "finalize_grad_value_kernel");
    ++g_launches;
    return BXR_OK;
}

// fused box->grid inputs handed through backward()/forward() (null for the location-taking ops)
struct FusedArgs {
    const void* boxes = nullptr;
    const void* angles = nullptr;
    const void* valid_ratios = nullptr;
    const void* kidx = nullptr;
    void* grad_boxes = nullptr;
    void* grad_angles = nullptr;
    bool softmax = false;     // grad_w0 receives the gradient of the logits (softmax chained in-kernel)
};

template <typename TV, typename TW, bool INSTANCE>
int backward(const TV* value, const int64_t* shapes, const int64_t* level_start, const TW* loc, const TW* w0, const TW* w1,
             const TV* grad_out, const TV* grad_mask, int B, int S, int H, int D, int L, int Nq, int P,
             TV* grad_value, TW* grad_loc, TW* grad_w0, TW* grad_w1,
             void* workspace, size_t workspace_bytes_given, unsigned flags, bxr_stream_t stream,
             const FusedArgs* fused = nullptr) {
    using TC = typename Compute<TV>::type;
    g_launches = 0;
    g_detail[0] = 0;
    if (int s = check_dims(B, S, H, D, L, Nq, P)) return s;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    fill_sizes(p, B, S, H, D, L, Nq, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n_value = (long long)B * S * H * D;
    const bool det = (flags & BXR_FLAG_DETERMINISTIC) != 0;
    if (n_value > 0) {
        if (!grad_value) return fail(BXR_ERR_NULL_POINTER, "grad_value is NULL");
        BXR_CUDA(cudaMemsetAsync(grad_value, 0, sizeof(TV) * (size_t)n_value, st));
    }
    if (p.rows == 0 || p.LP == 0) return BXR_OK;
    if ((!fused && !grad_loc) || !grad_w0 || (INSTANCE && !grad_w1)) return fail(BXR_ERR_NULL_POINTER, "gradient pointer is NULL");
    if (D == 0 || S == 0) {   // no channels / no pixels: all gradients are zero
        if (fused) {
            BXR_CUDA(cudaMemsetAsync(fused->grad_boxes, 0, sizeof(TW) * (size_t)p.rows * p.L * 4, st));
            if (fused->grad_angles) BXR_CUDA(cudaMemsetAsync(fused->grad_angles, 0, sizeof(TW) * (size_t)p.rows * p.L, st));
        } else
        BXR_CUDA(cudaMemsetAsync(grad_loc, 0, sizeof(TW) * (size_t)p.rows * p.LP * 2, st));
        BXR_CUDA(cudaMemsetAsync(grad_w0, 0, sizeof(TW) * (size_t)p.rows * p.LP, st));
        if (INSTANCE) BXR_CUDA(cudaMemsetAsync(grad_w1, 0, sizeof(TW) * (size_t)p.rows * p.LP, st));
        return BXR_OK;
    }
    if (!value || !shapes || !level_start || (!fused && !loc) || !w0 || !grad_out || (INSTANCE && (!w1 || !grad_mask)))
        return fail(BXR_ERR_NULL_POINTER, "input pointer is NULL");

    const size_t need = workspace_bytes((int)sizeof(TV), n_value, flags);
    if (need && (!workspace || workspace_bytes_given < need)) return fail(BXR_ERR_WORKSPACE, "workspace too small");
    if (need && !aligned16(workspace)) return fail(BXR_ERR_WORKSPACE, "workspace must be 16-byte aligned");

    p.value = value; p.shapes = shapes; p.level_start = level_start;
    p.loc = loc; p.w0 = w0; p.w1 = w1; p.grad_out = grad_out; p.grad_mask = grad_mask;
    p.grad_loc = grad_loc; p.grad_w0 = grad_w0; p.grad_w1 = grad_w1;
    if (fused) {
        p.boxes = fused->boxes; p.angles = fused->angles; p.valid_ratios = fused->valid_ratios; p.kidx = fused->kidx;
        p.grad_boxes = fused->grad_boxes; p.grad_angles = fused->grad_angles;
    }

    unsigned* bits = nullptr;
    float* scale = nullptr;
    void* acc = grad_value;
    if (det) {
        bits = static_cast<unsigned*>(workspace);
        scale = reinterpret_cast<float*>(bits + 4);
        acc = static_cast<char*>(workspace) + kWsHeader;
        BXR_CUDA(cudaMemsetAsync(workspace, 0, need, st));
        if (int s = absmax<TV>(grad_out, p.rows * D, bits + 0, st)) return s;
        if (int s = absmax<TW>(w0, p.rows * p.LP, bits + 1, st)) return s;
        if (INSTANCE) {
            if (int s = absmax<TV>(grad_mask, p.rows * D * P, bits + 2, st)) return s;
            if (int s = absmax<TW>(w1, p.rows * p.LP, bits + 3, st)) return s;
        }
        det_scale_kernel<0><<<1, 1, 0, st>>>(bits, scale);
        BXR_CUDA(cudaGetLastError());
        ++g_launches;
        p.det_scale = scale;
    } else if (sizeof(TV) == 2) {
        acc = workspace;
        BXR_CUDA(cudaMemsetAsync(workspace, 0, need, st));
    }
    p.grad_value_acc = acc;

    int status;
    const int g = vec_group<TV>(D, p.LP);
    const bool vec_ok = g && aligned16(value) && aligned16(grad_out) && aligned16(acc) && (!INSTANCE || aligned16(grad_mask)) &&
                        (fused ? (aligned16(fused->boxes) && aligned16(fused->grad_boxes) && aligned8(fused->kidx) &&
                                  (!fused->valid_ratios || aligned8(fused->valid_ratios)))
                               : (aligned8(loc) && aligned8(grad_loc)));
    if (fused && !vec_ok) return fail(BXR_ERR_UNSUPPORTED, "fused backward reached with a layout the fused kernels do not cover");
    if constexpr (!std::is_same<TV, double>::value && !INSTANCE) {
        if (vec_ok && !fused && use_tile<TV>(p, flags)) {
            status = det ? bwd_tile<TV, long long>(p, st) : bwd_tile<TV, float>(p, st);
            if (status) return status;
            if (det) return finalize<TV, long long>(static_cast<const long long*>(acc), grad_value, n_value, scale, st);
            if (sizeof(TV) == 2) return finalize<TV, float>(static_cast<const float*>(acc), grad_value, n_value, nullptr, st);
            return BXR_OK;
        }
    }
    if constexpr (!std::is_same<TV, double>::value) {
        bool win = false;
        if constexpr (!INSTANCE) win = vec_ok && (fused ? true : use_window(p, g, flags));
        bool own = false;
        if constexpr (INSTANCE) own = vec_ok && use_inst_own(p, g, flags);
        int g8 = 0;      // bf16 with 8-byte lanes
        if constexpr (std::is_same<TV, __nv_bfloat16>::value) {
            g8 = (vec_ok && !INSTANCE) ? bf16_lane8_group(D) : 0;
            if (g8) {
                if (fused || use_window(p, g8, flags)) win = true; else g8 = 0;
            }
        }
        if (win && g8) {
            if constexpr (!INSTANCE) {
                if (fused && fused->softmax) status = det ? dispatch_bwd_win<bf16x4_t, long long, true, true>(g8, p, st) : dispatch_bwd_win<bf16x4_t, float, true, true>(g8, p, st);
                else if (fused) status = det ? dispatch_bwd_win<bf16x4_t, long long, true>(g8, p, st) : dispatch_bwd_win<bf16x4_t, float, true>(g8, p, st);
                else status = det ? dispatch_bwd_win<bf16x4_t, long long>(g8, p, st, flags) : dispatch_bwd_win<bf16x4_t, float>(g8, p, st, flags);
            }
        } else if (win) {
            if constexpr (!INSTANCE) {
                if (fused && fused->softmax) status = det ? dispatch_bwd_win<TV, long long, true, true>(g, p, st) : dispatch_bwd_win<TV, float, true, true>(g, p, st);
                else if (fused) status = det ? dispatch_bwd_win<TV, long long, true>(g, p, st) : dispatch_bwd_win<TV, float, true>(g, p, st);
                else status = det ? dispatch_bwd_win<TV, long long>(g, p, st, flags) : dispatch_bwd_win<TV, float>(g, p, st, flags);
            }
        } else if (own) {
            if constexpr (INSTANCE) {
#if BXR_INST_BWD_BF16X4
                // bf16, head_dim 32: 8-byte lanes (G = 8, 4 channels per lane) halve the per-lane register load of the
                // 16-byte-lane kernel, which spills at 128 registers
                if constexpr (std::is_same<TV, __nv_bfloat16>::value) {
                    if (bf16_lane8_group(D) == 8 && use_inst_own(p, 8, flags))
                        return (status = det ? bwd_inst_own<bf16x4_t, 8, long long>(p, st) : bwd_inst_own<bf16x4_t, 8, float>(p, st))
                                   ? status
                                   : (det ? finalize<TV, long long>(static_cast<const long long*>(acc), grad_value, n_value, scale, st)
                                          : finalize<TV, float>(static_cast<const float*>(acc), grad_value, n_value, nullptr, st));
                }
#endif
                if (g == 8) status = det ? bwd_inst_own<TV, 8, long long>(p, st) : bwd_inst_own<TV, 8, float>(p, st);
                else status = det ? bwd_inst_own<TV, 4, long long>(p, st) : bwd_inst_own<TV, 4, float>(p, st);
            }
        } else if (vec_ok) {
            status = det ? dispatch_bwd_vec<TV, INSTANCE, long long>(g, p, st) : dispatch_bwd_vec<TV, INSTANCE, float>(g, p, st);
        } else {
            status = det ? launch_rows<attn_bwd_gen_kernel<TV, INSTANCE, long long>>(p, st, "attn_bwd_gen_kernel")
                         : launch_rows<attn_bwd_gen_kernel<TV, INSTANCE, TC>>(p, st, "attn_bwd_gen_kernel");
        }
    } else {
        status = det ? launch_rows<attn_bwd_gen_kernel<TV, INSTANCE, long long>>(p, st, "attn_bwd_gen_kernel")
                     : launch_rows<attn_bwd_gen_kernel<TV, INSTANCE, TC>>(p, st, "attn_bwd_gen_kernel");
    }
    if (status) return status;

    if (det) return finalize<TV, long long>(static_cast<const long long*>(acc), grad_value, n_value, scale, st);
    if (sizeof(TV) == 2) return finalize<TV, float>(static_cast<const float*>(acc), grad_value, n_value, nullptr, st);
    return BXR_OK;
}

// -------------------------------------------------------------------------------- fused box -> grid
size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

template <typename TW>
int launch_grid_gen(const AttnParams& p, TW* loc, cudaStream_t st) {
    const long long n = p.rows * p.LP;
    long long blocks = (n + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    grid_gen_kernel<TW><<<(int)blocks, 256, 0, st>>>(p, loc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "grid_gen_kernel");
    ++g_launches;
    return BXR_OK;
}

template <typename TW>
int launch_grid_bwd(const AttnParams& p, const TW* grad_loc, cudaStream_t st) {
    const long long n = p.rows * p.L;
    long long blocks = (n + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    grid_bwd_kernel<TW><<<(int)blocks, 256, 0, st>>>(p, grad_loc);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "grid_bwd_kernel");
    ++g_launches;
    return BXR_OK;
}

template <typename TW>
int launch_softmax_rows(const TW* logits, TW* out, long long rows, int n, cudaStream_t st) {
    if (rows <= 0 || n <= 0) return BXR_OK;
    long long blocks = (rows + 7) / 8;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    softmax_rows_kernel<TW><<<(int)blocks, 256, 0, st>>>(logits, out, rows, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "softmax_rows_kernel");
    ++g_launches;
    return BXR_OK;
}

template <typename TW>
int launch_softmax_bwd_rows(const TW* w, TW* g, long long rows, int n, cudaStream_t st) {
    if (rows <= 0 || n <= 0) return BXR_OK;
    long long blocks = (rows + 7) / 8;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    softmax_bwd_rows_kernel<TW><<<(int)blocks, 256, 0, st>>>(w, g, rows, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "softmax_bwd_rows_kernel");
    ++g_launches;
    return BXR_OK;
}

size_t fused_workspace_bytes(int dtype_bytes, int backward, int B, int S, int H, int D, int L, int Nq, int P, unsigned flags) {
    if (B < 0 || S < 0 || H < 0 || D < 0 || L < 0 || Nq < 0 || P < 0) return 0;
    const size_t tw = dtype_bytes == 8 ? 8 : 4;
    const size_t loc_bytes = align256((size_t)B * Nq * H * L * P * 2 * tw);
    const bool fused = fused_applies(dtype_bytes, B, S, H, D, L, P, flags);
    size_t n = 0;
    if (backward) n += align256(workspace_bytes(dtype_bytes, (long long)B * S * H * D, flags));
    if (!fused) n += loc_bytes * (backward ? 2 : 1);
    return n;
}

template <typename TV, typename TW>
int fused_forward(const TV* value, const int64_t* shapes, const int64_t* level_start, const TW* boxes, const TW* angles,
                  const TW* valid_ratios, const TW* kidx, const TW* attn, int B, int S, int H, int D, int L, int Nq, int P,
                  TV* out, void* ws, size_t ws_bytes, unsigned flags, bxr_stream_t stream, TW* attn_out = nullptr) {
    // attn_out != NULL: `attn` holds logits; their softmax over (L, P) is written to attn_out and used as the weights
    g_launches = 0;
    g_detail[0] = 0;
    if (int s = check_dims(B, S, H, D, L, Nq, P)) return s;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    fill_sizes(p, B, S, H, D, L, Nq, P);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p.rows == 0) return BXR_OK;
    if (D == 0 || S == 0 || p.LP == 0) {    // nothing to sample: zero output, but the weights are still defined
        if (attn_out && p.LP > 0) {
            if (!attn) return fail(BXR_ERR_NULL_POINTER, "input pointer is NULL");
            if (int s = launch_softmax_rows<TW>(attn, attn_out, p.rows, p.LP, st)) return s;
        }
        if (D == 0) return BXR_OK;
        if (!out) return fail(BXR_ERR_NULL_POINTER, "output pointer is NULL");
        BXR_CUDA(cudaMemsetAsync(out, 0, sizeof(TV) * (size_t)p.rows * D, st));
        return BXR_OK;
    }
    if (!out) return fail(BXR_ERR_NULL_POINTER, "output pointer is NULL");
    if (!value || !shapes || !level_start || !boxes || !kidx || !attn) return fail(BXR_ERR_NULL_POINTER, "input pointer is NULL");
    p.value = value; p.shapes = shapes; p.level_start = level_start; p.w0 = attn; p.out = out;
    p.boxes = boxes; p.angles = angles; p.valid_ratios = valid_ratios; p.kidx = kidx; p.attn_out = attn_out;

    const bool aligned = aligned16(value) && aligned16(out) && aligned16(boxes) && aligned8(kidx) && (!valid_ratios || aligned8(valid_ratios));
    if constexpr (!std::is_same<TV, double>::value) {
        if (attn_out && aligned && fused_applies((int)sizeof(TV), B, S, H, D, L, P, flags)) {
            if constexpr (std::is_same<TV, __nv_bfloat16>::value) {
                if (const int g8 = bf16_lane8_group(D)) return dispatch_fwd_win<bf16x4_t, true, true>(g8, p, st);
            }
            return dispatch_fwd_win<TV, true, true>(vec_group<TV>(D, p.LP), p, st);
        }
    }
    if (attn_out) {     // general path: softmax first, then as with given weights
        if (int s = launch_softmax_rows<TW>(attn, attn_out, p.rows, p.LP, st)) return s;
        attn = attn_out;
        p.w0 = attn_out;
        p.attn_out = nullptr;
    }
    if constexpr (!std::is_same<TV, double>::value) {
        if (aligned && fused_applies((int)sizeof(TV), B, S, H, D, L, P, flags)) {
            if constexpr (std::is_same<TV, __nv_bfloat16>::value) {
                if (const int g8 = bf16_lane8_group(D))
                    return use_staged(p, g8, 1, 1, flags) ? dispatch_fwd_stg<bf16x4_t, 1>(g8, p, st)
                                                          : dispatch_fwd_win<bf16x4_t, true>(g8, p, st);
            }
            const int g = vec_group<TV>(D, p.LP);
            return use_staged(p, g, 1, 1, flags) ? dispatch_fwd_stg<TV, 1>(g, p, st) : dispatch_fwd_win<TV, true>(g, p, st);
        }
    }
    // general path: materialise the grid, then the location-taking op
    const size_t loc_bytes = align256((size_t)p.rows * p.LP * 2 * sizeof(TW));
    if (!ws || ws_bytes < loc_bytes || !aligned16(ws)) return fail(BXR_ERR_WORKSPACE, "workspace too small for the sampling grid");
    TW* loc = static_cast<TW*>(ws);
    if (int s = launch_grid_gen<TW>(p, loc, st)) return s;
    const int n0 = g_launches;
    const int status = forward<TV, TW, false>(value, shapes, level_start, loc, attn, nullptr, B, S, H, D, L, Nq, P, out, nullptr,
                                              flags, stream);
    g_launches += n0;
    return status;
}

template <typename TV, typename TW>
int fused_backward(const TV* value, const int64_t* shapes, const int64_t* level_start, const TW* boxes, const TW* angles,
                   const TW* valid_ratios, const TW* kidx, const TW* attn, const TV* grad_out,
                   int B, int S, int H, int D, int L, int Nq, int P, TV* grad_value, TW* grad_boxes, TW* grad_angles,
                   TW* grad_attn, void* ws, size_t ws_bytes, unsigned flags, bxr_stream_t stream, bool smax = false) {
    // smax: `attn` are the softmax weights the forward wrote; grad_attn receives the gradient of the LOGITS
    g_detail[0] = 0;
    if (int s = check_dims(B, S, H, D, L, Nq, P)) return s;
    const long long rows = (long long)B * Nq * H;
    if (rows > 0 && L > 0 && (!boxes || !kidx || !grad_boxes || (angles && !grad_angles)))
        return fail(BXR_ERR_NULL_POINTER, "box pointer is NULL");
    FusedArgs fa;
    fa.boxes = boxes; fa.angles = angles; fa.valid_ratios = valid_ratios; fa.kidx = kidx;
    fa.grad_boxes = grad_boxes; fa.grad_angles = angles ? grad_angles : nullptr;
    fa.softmax = smax;
    const size_t bwd_ws = align256(workspace_bytes((int)sizeof(TV), (long long)B * S * H * D, flags));
    if (bwd_ws && (!ws || ws_bytes < bwd_ws)) return fail(BXR_ERR_WORKSPACE, "workspace too small");

    bool fused = false;
    if constexpr (!std::is_same<TV, double>::value) {
        fused = fused_applies((int)sizeof(TV), B, S, H, D, L, P, flags) && aligned16(value) && aligned16(grad_out) &&
                aligned16(grad_value) && aligned16(boxes) && aligned16(grad_boxes) && aligned8(kidx) &&
                (!valid_ratios || aligned8(valid_ratios)) && (!ws || aligned16(ws));
    }
    if (fused)
        return backward<TV, TW, false>(value, shapes, level_start, nullptr, attn, nullptr, grad_out, nullptr, B, S, H, D, L, Nq, P,
                                       grad_value, nullptr, grad_attn, nullptr, ws, ws_bytes, flags, stream, &fa);

    // general path: grid -> location-taking backward -> chain to the boxes
    const size_t loc_bytes = align256((size_t)rows * L * P * 2 * sizeof(TW));
    if (rows > 0 && (long long)L * P > 0) {
        if (!ws || ws_bytes < bwd_ws + 2 * loc_bytes || !aligned16(ws)) return fail(BXR_ERR_WORKSPACE, "workspace too small for the sampling grid");
    }
    TW* loc = reinterpret_cast<TW*>(static_cast<char*>(ws) + bwd_ws);
    TW* gloc = reinterpret_cast<TW*>(static_cast<char*>(ws) + bwd_ws + loc_bytes);
    AttnParams p;
    memset(&p, 0, sizeof(p));
    fill_sizes(p, B, S, H, D, L, Nq, P);
    p.boxes = boxes; p.angles = angles; p.valid_ratios = valid_ratios; p.kidx = kidx;
    p.grad_boxes = grad_boxes; p.grad_angles = fa.grad_angles;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int extra = 0;
    if (p.rows > 0 && p.LP > 0) {
        g_launches = 0;
        if (int s = launch_grid_gen<TW>(p, loc, st)) return s;
        extra = g_launches;
    }
    int status = backward<TV, TW, false>(value, shapes, level_start, loc, attn, nullptr, grad_out, nullptr, B, S, H, D, L, Nq, P,
                                         grad_value, gloc, grad_attn, nullptr, ws, bwd_ws, flags, stream);
    if (status) return status;
    extra += g_launches;
    if (p.rows > 0 && p.LP > 0) {
        if (int s = launch_grid_bwd<TW>(p, gloc, st)) return s;
        extra += 1;
    } else if (p.rows > 0 && p.L > 0) {
        BXR_CUDA(cudaMemsetAsync(grad_boxes, 0, sizeof(TW) * (size_t)p.rows * p.L * 4, st));
        if (fa.grad_angles) BXR_CUDA(cudaMemsetAsync(fa.grad_angles, 0, sizeof(TW) * (size_t)p.rows * p.L, st));
    }
    if (smax && p.rows > 0 && p.LP > 0) {
        if (int s = launch_softmax_bwd_rows<TW>(attn, grad_attn, p.rows, p.LP, st)) return s;
        extra += 1;
    }
    g_launches = extra;
    return BXR_OK;
}

// value_proj epilogue: mask fill + storage cast in one pass (boxattn_fused.cuh)
template <typename TI, typename TO>
int value_epilogue(const void* in, const unsigned char* mask, void* out, long long rows, int C, bxr_stream_t stream) {
    g_launches = 0;
    g_detail[0] = 0;
    if (rows < 0 || C < 0) return fail(BXR_ERR_BAD_DIM, "negative dimension");
    if (rows == 0 || C == 0) return BXR_OK;
    if (!in || !out) return fail(BXR_ERR_NULL_POINTER, "pointer is NULL");
    const int vec = (C % 8 == 0 && aligned16(in) && aligned16(out)) ? 1 : 0;
    const long long n = vec ? rows * (C / 8) : rows * C;
    long long blocks = (n + 255) / 256;
    if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
    value_epilogue_kernel<TI, TO><<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const TI*>(in), mask, static_cast<TO*>(out), rows, C, vec);
    BXR_CUDA(cudaGetLastError());
    ++g_launches;
    return BXR_OK;
}

}  // namespace BXR_SLICE_NS

using namespace BXR_SLICE_NS;

extern "C" {

#if BXR_TU_COMMON
int bxr_abi_version(void) { return BXR_ABI_VERSION; }

const char* bxr_status_string(int status) {
    switch (status) {
        case BXR_OK: return "ok";
        case BXR_ERR_NULL_POINTER: return "null pointer";
        case BXR_ERR_BAD_DIM: return "bad dimension";
        case BXR_ERR_WORKSPACE: return "workspace missing, too small or misaligned";
        case BXR_ERR_CUDA: return "CUDA error";
        case BXR_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

const char* bxr_last_error_detail(void) { return g_detail; }
int bxr_last_launch_count(void) { return g_launches; }

size_t bxr_attn_bwd_workspace_bytes(int dtype_bytes, int B, int S, int H, int D, unsigned flags) {
    if (B < 0 || S < 0 || H < 0 || D < 0) return 0;
    return workspace_bytes(dtype_bytes, (long long)B * S * H * D, flags);
}

size_t bxr_box_grid_attn_workspace_bytes(int dtype_bytes, int backward, int B, int S, int H, int D, int L, int Nq, int P,
                                         unsigned flags) {
    return fused_workspace_bytes(dtype_bytes, backward, B, S, H, D, L, Nq, P, flags);
}
int bxr_value_epilogue(const void* in, int in_bytes, const unsigned char* mask, void* out, int out_bytes, long long rows,
                       int C, bxr_stream_t stream) {
    if (in_bytes == 4 && out_bytes == 4) return value_epilogue<float, float>(in, mask, out, rows, C, stream);
    if (in_bytes == 4 && out_bytes == 2) return value_epilogue<float, __nv_bfloat16>(in, mask, out, rows, C, stream);
    if (in_bytes == 2 && out_bytes == 4) return value_epilogue<__nv_bfloat16, float>(in, mask, out, rows, C, stream);
    if (in_bytes == 2 && out_bytes == 2) return value_epilogue<__nv_bfloat16, __nv_bfloat16>(in, mask, out, rows, C, stream);
    return fail(BXR_ERR_UNSUPPORTED, "value epilogue: element sizes must be 4 (float) or 2 (bfloat16)");
}
#endif  // BXR_TU_COMMON

#define BXR_DEFINE_OPS_FWD(SUF, TVABI, TV, TW)                                                                       \
    int bxr_box_attn_fwd_##SUF(const TVABI* value, const int64_t* shapes, const int64_t* level_start, const TW* loc, \
                               const TW* attn, int B, int S, int H, int D, int L, int Nq, int P, TVABI* out,         \
                               unsigned flags, bxr_stream_t stream) {                                                \
        return forward<TV, TW, false>(reinterpret_cast<const TV*>(value), shapes, level_start, loc, attn, nullptr,   \
                                      B, S, H, D, L, Nq, P, reinterpret_cast<TV*>(out), nullptr, flags, stream);     \
    }                                                                                                                \
    int bxr_instance_attn_fwd_##SUF(const TVABI* value, const int64_t* shapes, const int64_t* level_start,           \
                                    const TW* loc, const TW* spatial_w, const TW* level_w, int B, int S, int H,      \
                                    int D, int L, int Nq, int P, TVABI* out, TVABI* mask_out, unsigned flags,        \
                                    bxr_stream_t stream) {                                                           \
        return forward<TV, TW, true>(reinterpret_cast<const TV*>(value), shapes, level_start, loc, spatial_w,        \
                                     level_w, B, S, H, D, L, Nq, P, reinterpret_cast<TV*>(out),                      \
                                     reinterpret_cast<TV*>(mask_out), flags, stream);                                \
    }                                                                                                                \
    int bxr_box_grid_attn_fwd_##SUF(const TVABI* value, const int64_t* shapes, const int64_t* level_start,           \
                                    const TW* boxes, const TW* angles, const TW* valid_ratios, const TW* kidx,       \
                                    const TW* attn, int B, int S, int H, int D, int L, int Nq, int P, TVABI* out,    \
                                    void* workspace, size_t workspace_bytes, unsigned flags, bxr_stream_t stream) {  \
        return fused_forward<TV, TW>(reinterpret_cast<const TV*>(value), shapes, level_start, boxes, angles,         \
                                     valid_ratios, kidx, attn, B, S, H, D, L, Nq, P, reinterpret_cast<TV*>(out),     \
                                     workspace, workspace_bytes, flags, stream);                                     \
    }

// InstanceAttention weights from the 2x2 logit maps (boxattn_fused.cuh); only float / double exist
#define BXR_DEFINE_INSTW_FWD(SUF, T)                                                                                 \
    int bxr_instance_weights_fwd_##SUF(const T* logits, long long rows, int L, int K, T* spatial_w, T* level_w,     \
                                       bxr_stream_t stream) {                                                        \
        g_launches = 0;                                                                                              \
        g_detail[0] = 0;                                                                                             \
        if (rows < 0 || L < 0 || K < 0 || L > BXR_MAX_LEVELS || (K & 1)) return fail(BXR_ERR_BAD_DIM, "rows, L >= 0, L <= 32, K even"); \
        if (rows == 0 || L == 0 || K == 0) return BXR_OK;                                                            \
        if (!logits || !spatial_w || !level_w) return fail(BXR_ERR_NULL_POINTER, "pointer is NULL");                 \
        long long blocks = (rows + 7) / 8;                                                                           \
        if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();                                                  \
        inst_weights_fwd_kernel<T><<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, spatial_w,    \
                                                                                               level_w, rows, L, K); \
        BXR_CUDA(cudaGetLastError());                                                                                \
        ++g_launches;                                                                                                \
        return BXR_OK;                                                                                               \
    }
#define BXR_DEFINE_INSTW_BWD(SUF, T)                                                                                 \
    int bxr_instance_weights_bwd_##SUF(const T* logits, const T* grad_spatial_w, const T* grad_level_w,             \
                                       long long rows, int L, int K, T* grad_logits, bxr_stream_t stream) {         \
        g_launches = 0;                                                                                              \
        g_detail[0] = 0;                                                                                             \
        if (rows < 0 || L < 0 || K < 0 || L > BXR_MAX_LEVELS || (K & 1)) return fail(BXR_ERR_BAD_DIM, "rows, L >= 0, L <= 32, K even"); \
        if (rows == 0 || L == 0) return BXR_OK;                                                                      \
        if (!logits || !grad_logits || (K > 0 && (!grad_spatial_w || !grad_level_w)))                                \
            return fail(BXR_ERR_NULL_POINTER, "pointer is NULL");                                                    \
        long long blocks = (rows + 7) / 8;                                                                           \
        if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();                                                  \
        inst_weights_bwd_kernel<T><<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(                      \
            logits, grad_spatial_w, grad_level_w, grad_logits, rows, L, K);                                          \
        BXR_CUDA(cudaGetLastError());                                                                                \
        ++g_launches;                                                                                                \
        return BXR_OK;                                                                                               \
    }

#define BXR_DEFINE_SMAX_FWD(SUF, TVABI, TV, TW)                                                                      \
    int bxr_box_grid_softmax_attn_fwd_##SUF(const TVABI* value, const int64_t* shapes, const int64_t* level_start,   \
                                            const TW* boxes, const TW* angles, const TW* valid_ratios,               \
                                            const TW* kidx, const TW* logits, int B, int S, int H, int D, int L,     \
                                            int Nq, int P, TVABI* out, TW* attn_out, void* workspace,                \
                                            size_t workspace_bytes, unsigned flags, bxr_stream_t stream) {           \
        if (!attn_out && (long long)B * Nq * H * L * P > 0) return fail(BXR_ERR_NULL_POINTER, "attn_out is NULL");   \
        return fused_forward<TV, TW>(reinterpret_cast<const TV*>(value), shapes, level_start, boxes, angles,         \
                                     valid_ratios, kidx, logits, B, S, H, D, L, Nq, P, reinterpret_cast<TV*>(out),   \
                                     workspace, workspace_bytes, flags, stream, attn_out);                           \
    }

#define BXR_DEFINE_SMAX_BWD(SUF, TVABI, TV, TW)                                                                      \
    int bxr_box_grid_softmax_attn_bwd_##SUF(const TVABI* value, const int64_t* shapes, const int64_t* level_start,   \
                                            const TW* boxes, const TW* angles, const TW* valid_ratios,               \
                                            const TW* kidx, const TW* attn, const TVABI* grad_out, int B, int S,     \
                                            int H, int D, int L, int Nq, int P, TVABI* grad_value, TW* grad_boxes,   \
                                            TW* grad_angles, TW* grad_logits, void* workspace,                       \
                                            size_t workspace_bytes, unsigned flags, bxr_stream_t stream) {           \
        return fused_backward<TV, TW>(reinterpret_cast<const TV*>(value), shapes, level_start, boxes, angles,        \
                                      valid_ratios, kidx, attn, reinterpret_cast<const TV*>(grad_out), B, S, H, D,   \
                                      L, Nq, P, reinterpret_cast<TV*>(grad_value), grad_boxes, grad_angles,          \
                                      grad_logits, workspace, workspace_bytes, flags, stream, true);                 \
    }

#define BXR_DEFINE_OPS_BWD(SUF, TVABI, TV, TW)                                                                       \
    int bxr_box_attn_bwd_##SUF(const TVABI* value, const int64_t* shapes, const int64_t* level_start, const TW* loc, \
                               const TW* attn, const TVABI* grad_out, int B, int S, int H, int D, int L, int Nq,     \
                               int P, TVABI* grad_value, TW* grad_loc, TW* grad_attn, void* workspace,               \
                               size_t workspace_bytes, unsigned flags, bxr_stream_t stream) {                        \
        return backward<TV, TW, false>(reinterpret_cast<const TV*>(value), shapes, level_start, loc, attn, nullptr,  \
                                       reinterpret_cast<const TV*>(grad_out), nullptr, B, S, H, D, L, Nq, P,         \
                                       reinterpret_cast<TV*>(grad_value), grad_loc, grad_attn, nullptr, workspace,   \
                                       workspace_bytes, flags, stream);                                              \
    }                                                                                                                \
    int bxr_instance_attn_bwd_##SUF(const TVABI* value, const int64_t* shapes, const int64_t* level_start,           \
                                    const TW* loc, const TW* spatial_w, const TW* level_w, const TVABI* grad_out,    \
                                    const TVABI* grad_mask, int B, int S, int H, int D, int L, int Nq, int P,        \
                                    TVABI* grad_value, TW* grad_loc, TW* grad_spatial_w, TW* grad_level_w,           \
                                    void* workspace, size_t workspace_bytes, unsigned flags, bxr_stream_t stream) {  \
        return backward<TV, TW, true>(reinterpret_cast<const TV*>(value), shapes, level_start, loc, spatial_w,       \
                                      level_w, reinterpret_cast<const TV*>(grad_out),                                \
                                      reinterpret_cast<const TV*>(grad_mask), B, S, H, D, L, Nq, P,                  \
                                      reinterpret_cast<TV*>(grad_value), grad_loc, grad_spatial_w, grad_level_w,     \
                                      workspace, workspace_bytes, flags, stream);                                    \
    }                                                                                                                \
    int bxr_box_grid_attn_bwd_##SUF(const TVABI* value, const int64_t* shapes, const int64_t* level_start,           \
                                    const TW* boxes, const TW* angles, const TW* valid_ratios, const TW* kidx,       \
                                    const TW* attn, const TVABI* grad_out, int B, int S, int H, int D, int L,        \
                                    int Nq, int P, TVABI* grad_value, TW* grad_boxes, TW* grad_angles,               \
                                    TW* grad_attn, void* workspace, size_t workspace_bytes, unsigned flags,          \
                                    bxr_stream_t stream) {                                                           \
        return fused_backward<TV, TW>(reinterpret_cast<const TV*>(value), shapes, level_start, boxes, angles,        \
                                      valid_ratios, kidx, attn, reinterpret_cast<const TV*>(grad_out), B, S, H, D,   \
                                      L, Nq, P, reinterpret_cast<TV*>(grad_value), grad_boxes, grad_angles,          \
                                      grad_attn, workspace, workspace_bytes, flags, stream);                         \
    }

#if BXR_TU_DIRS & 1
#define BXR_FWD(...) BXR_DEFINE_OPS_FWD(__VA_ARGS__) BXR_DEFINE_SMAX_FWD(__VA_ARGS__)
#else
#define BXR_FWD(...)
#endif
#if BXR_TU_DIRS & 2
#define BXR_BWD(...) BXR_DEFINE_OPS_BWD(__VA_ARGS__) BXR_DEFINE_SMAX_BWD(__VA_ARGS__)
#else
#define BXR_BWD(...)
#endif

#if BXR_TU_DTYPES & 1
BXR_FWD(f32, float, float, float)
BXR_BWD(f32, float, float, float)
#if BXR_TU_DIRS & 1
BXR_DEFINE_INSTW_FWD(f32, float)
#endif
#if BXR_TU_DIRS & 2
BXR_DEFINE_INSTW_BWD(f32, float)
#endif
#endif
#if BXR_TU_DTYPES & 2
BXR_FWD(f64, double, double, double)
BXR_BWD(f64, double, double, double)
#if BXR_TU_DIRS & 1
BXR_DEFINE_INSTW_FWD(f64, double)
#endif
#if BXR_TU_DIRS & 2
BXR_DEFINE_INSTW_BWD(f64, double)
#endif
#endif
#if BXR_TU_DTYPES & 4
BXR_FWD(bf16, bxr_bf16, __nv_bfloat16, float)
BXR_BWD(bf16, bxr_bf16, __nv_bfloat16, float)
#endif

}  // extern "C"
