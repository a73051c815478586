// Fused "box -> K x K grid -> attention" entry points.
//
// The reference builds the sampling grid in PyTorch (BoxAttention._where_to_attend,
// box_attention.py:196-214; Box3dAttention adds the rotation, :304-338), materialising a
// (B, Nq, H, L, P, 2) tensor that the op then re-reads, and autograd materialises its gradient.
// Here the op takes the boxes themselves:
//     loc[b,q,h,l,p] = (centre + R(angle) (kernel_index_p * relu(size))) * valid_ratio[b,l]
// The footprint-window kernels generate the points in registers (boxattn_window.cuh, FUSED) and
// reduce the per-point location gradients to the 4 (+1) box gradients in-kernel.  Shapes those
// kernels do not cover (odd head dims, fp64, tiny calls, huge grids) go through the two small
// kernels below around the location-taking op -- same interface, same results.
#pragma once

#include "boxattn_kernels.cuh"

namespace bxr {

template <typename TW>
__global__ void grid_gen_kernel(const AttnParams p, TW* __restrict__ loc_out) {
    const TW* __restrict__ boxes = static_cast<const TW*>(p.boxes);
    const TW* __restrict__ angles = static_cast<const TW*>(p.angles);
    const TW* __restrict__ vr = static_cast<const TW*>(p.valid_ratios);
    const TW* __restrict__ kidx = static_cast<const TW*>(p.kidx);
    const long long n = p.rows * p.LP;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (long long)gridDim.x * blockDim.x) {
        const int pt = (int)(s % p.P);
        const long long rl = s / p.P;
        const int l = (int)(rl % p.L);
        const long long b = rl / ((long long)p.L * p.H * p.Nq);
        const TW cx = boxes[rl * 4], cy = boxes[rl * 4 + 1];
        const TW sx = boxes[rl * 4 + 2] > (TW)0 ? boxes[rl * 4 + 2] : (TW)0;
        const TW sy = boxes[rl * 4 + 3] > (TW)0 ? boxes[rl * 4 + 3] : (TW)0;
        const TW ux = kidx[2 * pt] * sx, uy = kidx[2 * pt + 1] * sy;
        TW cs = 1, sn = 0;
        if (angles) { cs = cos(angles[rl]); sn = sin(angles[rl]); }
        TW x = cx + (ux * cs - uy * sn), y = cy + (ux * sn + uy * cs);
        if (vr) { x *= vr[(b * p.L + l) * 2]; y *= vr[(b * p.L + l) * 2 + 1]; }
        loc_out[2 * s] = x;
        loc_out[2 * s + 1] = y;
    }
}

// one thread per (row, level): reduce the P location gradients to the box / angle gradients
template <typename TW>
__global__ void grid_bwd_kernel(const AttnParams p, const TW* __restrict__ grad_loc) {
    const TW* __restrict__ boxes = static_cast<const TW*>(p.boxes);
    const TW* __restrict__ angles = static_cast<const TW*>(p.angles);
    const TW* __restrict__ vr = static_cast<const TW*>(p.valid_ratios);
    const TW* __restrict__ kidx = static_cast<const TW*>(p.kidx);
    TW* __restrict__ gb = static_cast<TW*>(p.grad_boxes);
    TW* __restrict__ ga = static_cast<TW*>(p.grad_angles);
    const long long n = p.rows * p.L;
    for (long long rl = (long long)blockIdx.x * blockDim.x + threadIdx.x; rl < n; rl += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(rl % p.L);
        const long long b = rl / ((long long)p.L * p.H * p.Nq);
        const bool pw = boxes[rl * 4 + 2] > (TW)0, ph = boxes[rl * 4 + 3] > (TW)0;
        const TW sx = pw ? boxes[rl * 4 + 2] : (TW)0, sy = ph ? boxes[rl * 4 + 3] : (TW)0;
        TW cs = 1, sn = 0;
        if (angles) { cs = cos(angles[rl]); sn = sin(angles[rl]); }
        TW vx = 1, vy = 1;
        if (vr) { vx = vr[(b * p.L + l) * 2]; vy = vr[(b * p.L + l) * 2 + 1]; }
        TW bcx = 0, bcy = 0, bw = 0, bh = 0, ba = 0;
        for (int pt = 0; pt < p.P; ++pt) {
            const long long s = rl * p.P + pt;
            const TW gx = grad_loc[2 * s] * vx, gy = grad_loc[2 * s + 1] * vy;
            const TW kx = kidx[2 * pt], ky = kidx[2 * pt + 1];
            const TW ux = kx * sx, uy = ky * sy;
            bcx += gx;
            bcy += gy;
            bw += kx * (gx * cs + gy * sn);
            bh += ky * (gy * cs - gx * sn);
            ba += gx * (-ux * sn - uy * cs) + gy * (ux * cs - uy * sn);
        }
        gb[rl * 4] = bcx;
        gb[rl * 4 + 1] = bcy;
        gb[rl * 4 + 2] = pw ? bw : (TW)0;
        gb[rl * 4 + 3] = ph ? bh : (TW)0;
        if (ga) ga[rl] = ba;
    }
}

// ---- softmax over the last dimension (n = L*P) of (rows, n), one warp per row: the general-path companions
// of the in-kernel softmax of the window kernels
template <typename TW>
__global__ void softmax_rows_kernel(const TW* __restrict__ logits, TW* __restrict__ out, long long rows, int n) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < rows; r += nwarps) {
        const TW* z = logits + r * n;
        TW mx = -INFINITY;
        for (int i = lane; i < n; i += 32) mx = z[i] > mx ? z[i] : mx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const TW t = __shfl_xor_sync(0xffffffffu, mx, o);
            mx = t > mx ? t : mx;
        }
        TW sum = 0;
        for (int i = lane; i < n; i += 32) sum += exp(z[i] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const TW inv = (TW)1 / sum;
        for (int i = lane; i < n; i += 32) out[r * n + i] = exp(z[i] - mx) * inv;
    }
}

// g <- w * (g - sum_row(w * g)), in place
template <typename TW>
__global__ void softmax_bwd_rows_kernel(const TW* __restrict__ w, TW* __restrict__ g, long long rows, int n) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < rows; r += nwarps) {
        TW dotp = 0;
        for (int i = lane; i < n; i += 32) dotp += w[r * n + i] * g[r * n + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dotp += __shfl_xor_sync(0xffffffffu, dotp, o);
        for (int i = lane; i < n; i += 32) g[r * n + i] = w[r * n + i] * (g[r * n + i] - dotp);
    }
}

// ---- InstanceAttention's weights from its 2 x 2 logit map per (head, level) (SURVEY.md 8 row f2;
// e2edet/module/box_attention.py:93-110): the map is nearest-upsampled to K x K (repeat_interleave), then
//   spatial_w = softmax over all (L, K, K) entries,    level_w = softmax over L at every (i, j).
// With r = (K/2)^2 copies of each logit:  spatial_w = exp(z - m) / (r * sum_{l,q} exp(z - m)),
// level_w = exp(z - m_q) / sum_l exp(z_lq - m_q)  per quadrant q.  One warp per (b, query, head) row.
constexpr int kInstWMaxLogits = 4 * kMaxLevels;

template <typename T>
__device__ __forceinline__ void inst_row_weights(const T* __restrict__ z, int L, int K, int lane, T* s_sw, T* s_lw) {
    // every lane derives the (tiny) statistics itself; lanes share the per-(l, q) results through shared memory
    const int n = 4 * L;
    T m = -INFINITY, mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int t = 0; t < n; ++t) {
        const T v = z[t];
        m = v > m ? v : m;
        mq[t & 3] = v > mq[t & 3] ? v : mq[t & 3];
    }
    T sum = 0, sq[4] = {0, 0, 0, 0};
    for (int t = 0; t < n; ++t) {
        sum += exp(z[t] - m);
        sq[t & 3] += exp(z[t] - mq[t & 3]);
    }
    const T r = (T)((K / 2) * (K / 2));
    for (int t = lane; t < n; t += 32) {
        s_sw[t] = exp(z[t] - m) / (r * sum);
        s_lw[t] = exp(z[t] - mq[t & 3]) / sq[t & 3];
    }
    __syncwarp();
}

template <typename T>
__global__ void inst_weights_fwd_kernel(const T* __restrict__ logits, T* __restrict__ spatial_w, T* __restrict__ level_w,
                                        long long rows, int L, int K) {
    __shared__ T s_sw[8][kInstWMaxLogits], s_lw[8][kInstWMaxLogits];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int KK = K * K, half = K / 2, n_out = L * KK;
    for (long long row = warp; row < rows; row += nwarps) {
        inst_row_weights<T>(logits + row * 4 * L, L, K, lane, s_sw[wid], s_lw[wid]);
        for (int idx = lane; idx < n_out; idx += 32) {
            const int l = idx / KK, rem = idx - l * KK, i = rem / K, j = rem - i * K;
            const int t = l * 4 + (i >= half ? 2 : 0) + (j >= half ? 1 : 0);
            spatial_w[row * n_out + idx] = s_sw[wid][t];
            level_w[row * n_out + idx] = s_lw[wid][t];
        }
        __syncwarp();
    }
}

// grad_logits[l, q] = s_lq (Gs_lq - r * sum_{l',q'} s_l'q' Gs_l'q') + w_lq (Gl_lq - sum_l' w_l'q Gl_l'q)
// with Gs / Gl the gradients of spatial_w / level_w summed over the quadrant's (K/2)^2 entries.
template <typename T>
__global__ void inst_weights_bwd_kernel(const T* __restrict__ logits, const T* __restrict__ grad_sw, const T* __restrict__ grad_lw,
                                        T* __restrict__ grad_logits, long long rows, int L, int K) {
    __shared__ T s_sw[8][kInstWMaxLogits], s_lw[8][kInstWMaxLogits], s_gs[8][kInstWMaxLogits], s_gl[8][kInstWMaxLogits];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int KK = K * K, half = K / 2, n_out = L * KK, r = half * half, n = 4 * L;
    for (long long row = warp; row < rows; row += nwarps) {
        inst_row_weights<T>(logits + row * n, L, K, lane, s_sw[wid], s_lw[wid]);
        for (int t = 0; t < n; ++t) {
            const int l = t >> 2, a = (t >> 1) & 1, b = t & 1;
            T ps = 0, pl = 0;
            for (int e = lane; e < r; e += 32) {
                const int i = a * half + e / half, j = b * half + e % half;
                const long long idx = row * n_out + l * KK + i * K + j;
                ps += grad_sw[idx];
                pl += grad_lw[idx];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ps += __shfl_xor_sync(0xffffffffu, ps, o);
                pl += __shfl_xor_sync(0xffffffffu, pl, o);
            }
            if (lane == 0) { s_gs[wid][t] = ps; s_gl[wid][t] = pl; }
        }
        __syncwarp();
        T ds = 0;
        for (int t = 0; t < n; ++t) ds += s_sw[wid][t] * s_gs[wid][t];
        for (int t = lane; t < n; t += 32) {
            T dl = 0;
            for (int l2 = 0; l2 < L; ++l2) dl += s_lw[wid][l2 * 4 + (t & 3)] * s_gl[wid][l2 * 4 + (t & 3)];
            grad_logits[row * n + t] = s_sw[wid][t] * (s_gs[wid][t] - (T)r * ds) + s_lw[wid][t] * (s_gl[wid][t] - dl);
        }
        __syncwarp();
    }
}

// ---- value_proj epilogue (e2edet/module/box_attention.py:222-225): padding-mask fill and the
// cast to the storage type the gather wants, in one pass over the projected value:
//   out[r, :] = mask[r] ? 0 : (TO) in[r, :]          r = (b, s) pixel, C = heads * head_dim channels
// (the (B,S,C) -> (B,S,H,D) "head-major" step is a view).  The backward is the same kernel with the types swapped.
template <typename T> struct Chan8;
template <> struct Chan8<float> {
    __device__ static void load(const float* p, float (&v)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ static void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Chan8<__nv_bfloat16> {
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) { Vec16<__nv_bfloat16>::load(p, v); }
    __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) { Vec16<__nv_bfloat16>::store(p, v); }
};

template <typename TI, typename TO>
__global__ void value_epilogue_kernel(const TI* __restrict__ in, const unsigned char* __restrict__ mask, TO* __restrict__ out,
                                      long long rows, int C, int vec) {
    if (vec) {       // C % 8 == 0 and 16-byte aligned rows: 8 channels per thread
        const int per_row = C / 8;
        const long long n = rows * per_row;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
            const long long r = t / per_row;
            float v[8];
            if (mask && mask[r]) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = 0.f;
            } else {
                Chan8<TI>::load(in + t * 8, v);
            }
            Chan8<TO>::store(out + t * 8, v);
        }
    } else {
        const long long n = rows * C;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
            out[t] = (mask && mask[t / C]) ? from_f<TO, float>(0.f) : from_f<TO, float>(to_f(in[t]));
    }
}

}  // namespace bxr
