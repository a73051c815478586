/*
 * boxattn_b200 -- C ABI of the B200-native box-attention operators.
 *
 * This is the drop-in boundary for the hot path of kienduynguyen/BoxeR:
 * the four native entry points the reference binds through pybind11
 *
 *   box_attn_forward / box_attn_backward           e2edet/module/ops/src/vision.cpp:8-9
 *     -> e2edet::box_attn_cuda_forward/backward     e2edet/module/ops/src/box_attn/box_attn.cu:15-71, :74-135
 *   instance_attn_forward / instance_attn_backward  e2edet/module/ops/src/vision.cpp:10-11
 *     -> e2edet::instance_attn_cuda_forward/backward e2edet/module/ops/src/instance_attn/instance_attn.cu:15-82, :85-157
 *
 * re-stated as plain C: raw device pointers, sizes, a stream, an int status.
 * No torch / ATen types cross this boundary.
 *
 * Contract
 *   - Ownership: the CALLER allocates every buffer (outputs, gradients, workspace).
 *     The library never allocates or frees device memory and never synchronises
 *     the device or the stream; all work is enqueued on `stream`.
 *   - Device: the current CUDA device of the calling thread (the host wrapper
 *     sets it from the tensors).  Re-entrant, no mutable global state apart
 *     from a per-device cache of immutable device attributes.
 *   - Layouts are the reference's, all contiguous:
 *       value        (B, S, H, D)          S = sum_l h_l*w_l, levels concatenated row-major (y, x)
 *       shapes       (L, 2) int64 DEVICE   (h_l, w_l)            [read on the device, as the reference does,
 *       level_start  (L,)   int64 DEVICE   first index of level l  box_attn_kernel.cuh:313-316 -- no D2H sync]
 *       loc          (B, Nq, H, L, P, 2)   (x, y) normalised; pixel = loc * size - 0.5
 *       attn / spatial_w / level_w (B, Nq, H, L, P)
 *       out          (B, Nq, H*D)
 *       mask_out     (B, Nq, P, H*D)       instance op only
 *     Gradients have the layout of what they are the gradient of.
 *   - Arithmetic: bilinear, zero padding, window test -1 < pixel < size
 *     (box_attn_kernel.cuh:325-328), fp32 accumulation (fp64 for _f64).
 *   - dtypes: _f32 (all tensors float), _f64 (all double; slow-is-fine, for
 *     gradcheck), _bf16 (value / out / mask_out / grad_out / grad_mask /
 *     grad_value are bfloat16 bits; loc, weights and their gradients float).
 *   - Backward needs `grad_value` zero-filled: the library enqueues that memset
 *     itself.  grad_loc / grad_* weights are fully overwritten.
 *   - Errors: a non-zero bxr_status is returned (nothing is printed); the
 *     reference only printf()s launch failures (box_attn_kernel.cuh:1118-1122).
 *     bxr_status_string() names it, bxr_last_error_detail() gives the CUDA
 *     error string of the calling thread's last failure.
 *   - L <= BXR_MAX_LEVELS.
 */
#ifndef BOXATTN_B200_H_
#define BOXATTN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BXR_ABI_VERSION 1
#define BXR_MAX_LEVELS 32

typedef enum bxr_status {
    BXR_OK = 0,
    BXR_ERR_NULL_POINTER = 1,   /* a required pointer is NULL */
    BXR_ERR_BAD_DIM = 2,        /* negative dim, L > BXR_MAX_LEVELS, index range overflow */
    BXR_ERR_WORKSPACE = 3,      /* workspace missing or smaller than *_workspace_bytes() */
    BXR_ERR_CUDA = 4,           /* a CUDA runtime call or kernel launch failed */
    BXR_ERR_UNSUPPORTED = 5     /* flag / dtype combination not available */
} bxr_status;

/* flags (bit set) */
#define BXR_FLAG_DETERMINISTIC 0x1u /* backward: order-independent (bit-reproducible) grad_value scatter via
                                       64-bit fixed-point accumulation instead of floating-point atomics */

#define BXR_FLAG_PATH_WINDOW 0x2u   /* tuning/testing: use the footprint-window kernels whenever they apply,
                                       even for row counts where the point-split kernels are the default */
#define BXR_FLAG_PATH_POINT 0x4u    /* tuning/testing: never use the footprint-window kernels */

#define BXR_FLAG_STAGED 0x8u        /* tuning/testing: footprint-window forward with TMA-staged row operands (cp.async.bulk
                                       + mbarrier per warp) and a pooled multi-level window; see boxattn_staged.cuh */

#define BXR_FLAG_PATH_TILE 0x10u     /* tuning/testing: query-tile x value-tile kernels (TMA-staged value halos in shared
                                       memory, boxattn_tile.cuh) whenever they apply (Nq == S, head_dim 32, P <= 16) */

#define BXR_FLAG_NO_TILE_ORDER 0x20u /* tuning/testing: footprint-window kernels with work units in memory order even for
                                       self-attention-shaped calls (Nq == S), where units are dealt as 2-D query tiles */

typedef void* bxr_stream_t;   /* a cudaStream_t */
typedef uint16_t bxr_bf16;    /* raw bfloat16 bits */

int bxr_abi_version(void);
const char* bxr_status_string(int status);
const char* bxr_last_error_detail(void);
/* number of kernels (not memsets) the calling thread's last successful call enqueued */
int bxr_last_launch_count(void);

/* Bytes of scratch the backward calls need for these sizes / flags (0 = none).
 * dtype_bytes: 4 (_f32), 8 (_f64), 2 (_bf16). */
size_t bxr_attn_bwd_workspace_bytes(int dtype_bytes, int B, int S, int H, int D, unsigned flags);

#define BXR_DECLARE_OPS(SUF, TV, TW)                                                                         \
    /* box_attn_forward  (box_attn.cu:15-71): out[b,q,h,:] = sum_l sum_p attn * bilinear(value_l, loc) */   \
    int bxr_box_attn_fwd_##SUF(const TV* value, const int64_t* shapes, const int64_t* level_start,           \
                               const TW* loc, const TW* attn,                                                \
                               int B, int S, int H, int D, int L, int Nq, int P,                             \
                               TV* out, unsigned flags, bxr_stream_t stream);                                \
    /* box_attn_backward (box_attn.cu:74-135) */                                                            \
    int bxr_box_attn_bwd_##SUF(const TV* value, const int64_t* shapes, const int64_t* level_start,           \
                               const TW* loc, const TW* attn, const TV* grad_out,                            \
                               int B, int S, int H, int D, int L, int Nq, int P,                             \
                               TV* grad_value, TW* grad_loc, TW* grad_attn,                                  \
                               void* workspace, size_t workspace_bytes, unsigned flags, bxr_stream_t stream); \
    /* instance_attn_forward (instance_attn.cu:15-82): out as above with spatial_w;                          \
       mask_out[b,q,p,h,:] = sum_l level_w * bilinear */                                                     \
    int bxr_instance_attn_fwd_##SUF(const TV* value, const int64_t* shapes, const int64_t* level_start,      \
                                    const TW* loc, const TW* spatial_w, const TW* level_w,                   \
                                    int B, int S, int H, int D, int L, int Nq, int P,                        \
                                    TV* out, TV* mask_out, unsigned flags, bxr_stream_t stream);             \
    /* instance_attn_backward (instance_attn.cu:85-157) */                                                   \
    int bxr_instance_attn_bwd_##SUF(const TV* value, const int64_t* shapes, const int64_t* level_start,      \
                                    const TW* loc, const TW* spatial_w, const TW* level_w,                   \
                                    const TV* grad_out, const TV* grad_mask,                                 \
                                    int B, int S, int H, int D, int L, int Nq, int P,                        \
                                    TV* grad_value, TW* grad_loc, TW* grad_spatial_w, TW* grad_level_w,      \
                                    void* workspace, size_t workspace_bytes, unsigned flags,                 \
                                    bxr_stream_t stream);

BXR_DECLARE_OPS(f32, float, float)
BXR_DECLARE_OPS(f64, double, double)
BXR_DECLARE_OPS(bf16, bxr_bf16, float)

/*
 * Fused box -> grid -> attention (beyond the reference's native surface).
 * Replaces BoxAttention._where_to_attend (e2edet/module/box_attention.py:196-214) and the rotated
 * variant of Box3dAttention (:304-338) *plus* box_attn_forward/backward: the K x K sampling grid
 *     loc[b,q,h,l,p] = (centre + R(angle) (kernel_index_p * relu(size))) * valid_ratio[b,l]
 * is generated inside the kernel, and the backward returns the gradient w.r.t. the boxes / angles
 * instead of a (B,Nq,H,L,P,2) location gradient.
 *   boxes          (B, Nq, H, L, 4)  cx, cy, w, h   (= ref_windows + offset/8 * ref_wh, computed by the caller)
 *   angles         (B, Nq, H, L)     radians, or NULL (no rotation)
 *   valid_ratios   (B, L, 2)         (x, y), or NULL
 *   kernel_indices (P, 2)            the module's buffer
 * Workspace: bxr_box_grid_attn_workspace_bytes() (used when a shape falls outside the fused kernels
 * and the grid has to be materialised internally, and by the _bf16 / deterministic backward).
 */
size_t bxr_box_grid_attn_workspace_bytes(int dtype_bytes, int backward, int B, int S, int H, int D, int L, int Nq, int P,
                                         unsigned flags);

#define BXR_DECLARE_FUSED(SUF, TV, TW)                                                                          \
    int bxr_box_grid_attn_fwd_##SUF(const TV* value, const int64_t* shapes, const int64_t* level_start,          \
                                    const TW* boxes, const TW* angles, const TW* valid_ratios,                   \
                                    const TW* kernel_indices, const TW* attn,                                    \
                                    int B, int S, int H, int D, int L, int Nq, int P, TV* out,                   \
                                    void* workspace, size_t workspace_bytes, unsigned flags, bxr_stream_t stream); \
    int bxr_box_grid_attn_bwd_##SUF(const TV* value, const int64_t* shapes, const int64_t* level_start,          \
                                    const TW* boxes, const TW* angles, const TW* valid_ratios,                   \
                                    const TW* kernel_indices, const TW* attn, const TV* grad_out,                \
                                    int B, int S, int H, int D, int L, int Nq, int P,                            \
                                    TV* grad_value, TW* grad_boxes, TW* grad_angles, TW* grad_attn,              \
                                    void* workspace, size_t workspace_bytes, unsigned flags, bxr_stream_t stream);

BXR_DECLARE_FUSED(f32, float, float)
BXR_DECLARE_FUSED(f64, double, double)
BXR_DECLARE_FUSED(bf16, bxr_bf16, float)

/*
 * Fused softmax -> box -> grid -> attention.  As bxr_box_grid_attn_*, but takes the
 * attention LOGITS (B, Nq, H, L*P) that BoxAttention.forward feeds to F.softmax(dim=-1)
 * (e2edet/module/box_attention.py:227-231; Box3dAttention :348-352): the softmax over a (b, q, head) row's
 * L*P points runs in the kernel prologue.
 *   fwd: writes out and attn_out (B, Nq, H, L, P) = softmax(logits) -- the module returns it, the backward needs it.
 *   bwd: takes attn (= the forward's attn_out) and returns grad_logits = attn * (grad_attn - sum_row(attn * grad_attn)).
 * Workspace: bxr_box_grid_attn_workspace_bytes().
 */
#define BXR_DECLARE_SMAX(SUF, TV, TW)                                                                               \
    int bxr_box_grid_softmax_attn_fwd_##SUF(const TV* value, const int64_t* shapes, const int64_t* level_start,     \
                                            const TW* boxes, const TW* angles, const TW* valid_ratios,              \
                                            const TW* kernel_indices, const TW* logits,                             \
                                            int B, int S, int H, int D, int L, int Nq, int P, TV* out, TW* attn_out, \
                                            void* workspace, size_t workspace_bytes, unsigned flags,                \
                                            bxr_stream_t stream);                                                   \
    int bxr_box_grid_softmax_attn_bwd_##SUF(const TV* value, const int64_t* shapes, const int64_t* level_start,     \
                                            const TW* boxes, const TW* angles, const TW* valid_ratios,              \
                                            const TW* kernel_indices, const TW* attn, const TV* grad_out,           \
                                            int B, int S, int H, int D, int L, int Nq, int P,                       \
                                            TV* grad_value, TW* grad_boxes, TW* grad_angles, TW* grad_logits,       \
                                            void* workspace, size_t workspace_bytes, unsigned flags,                \
                                            bxr_stream_t stream);

BXR_DECLARE_SMAX(f32, float, float)
BXR_DECLARE_SMAX(f64, double, double)
BXR_DECLARE_SMAX(bf16, bxr_bf16, float)

/*
 * InstanceAttention's two weight tensors from its 2 x 2 logit map per (head, level).
 * Replaces, in one launch each way, the chain in e2edet/module/box_attention.py:93-110:
 *   view(B,Nq,H,L,2,2) -> repeat_interleave(K/2) x2 -> softmax over (L,K,K) = spatial_w; softmax over L = level_w
 * and its autograd backward.
 *   logits (rows, L, 2, 2), rows = B*Nq*H;  spatial_w / level_w (rows, L, K, K);  K even.  fp32 / fp64 only
 *   (the weights are fp32 in every mode of the ops above).
 */
/*
 * value_proj epilogue (e2edet/module/box_attention.py:222-225): padding-mask fill + cast to the
 * storage type of the gather in one pass:  out[r, :] = mask[r] ? 0 : cast(in[r, :]),  r over rows = B*S pixels.
 *   in / out: float (element bytes 4) or bfloat16 (2), C channels per pixel;  mask: rows bytes (bool), or NULL.
 * Its backward is the same call with grad_out as `in` and the input's type as the output type.
 */
int bxr_value_epilogue(const void* in, int in_bytes, const unsigned char* mask, void* out, int out_bytes, long long rows,
                       int C, bxr_stream_t stream);

#define BXR_DECLARE_INSTW(SUF, T)                                                                                  \
    int bxr_instance_weights_fwd_##SUF(const T* logits, long long rows, int L, int K, T* spatial_w, T* level_w,   \
                                       bxr_stream_t stream);                                                       \
    int bxr_instance_weights_bwd_##SUF(const T* logits, const T* grad_spatial_w, const T* grad_level_w,           \
                                       long long rows, int L, int K, T* grad_logits, bxr_stream_t stream);

BXR_DECLARE_INSTW(f32, float)
BXR_DECLARE_INSTW(f64, double)

#ifdef __cplusplus
}
#endif
#endif /* BOXATTN_B200_H_ */
