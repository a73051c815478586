// L2 reduction throughput of `red.global.add.v4.f32` at the access shape of the backward scatter: 8 lanes x 16 B = one
// 128-byte line of grad_value per (pixel, head), lines picked pseudo-randomly in a buffer of the headline grad_value
// size (22 223 x 256 fp32 = 22.8 MB), in a 1 MB hot set, and sequentially.  Prints GB/s of reduction payload; the
// backward's scatter cost (profiles/r02xy_ablation.json) is read against this number.
// `ld_v4_*` is the forward's counterpart: independent random 128-byte line gathers served by L2 (22.8 MB) / L1+L2 (1 MB).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/microbench/red_throughput scripts/microbench/red_throughput.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int MODE>  // 0 red.v4, 1 st.v4, 2 scalar atomicAdd x4
__global__ void scatter(float* buf, uint32_t lines, int iters, int sequential) {
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, lane = threadIdx.x & 7;
    const uint32_t ngroups = (gridDim.x * blockDim.x) >> 3;
    uint32_t s = group * 2654435761u + 12345u;
    for (int i = 0; i < iters; ++i) {
        uint32_t line;
        if (sequential) line = (group + (uint32_t)i * ngroups) % lines;
        else { s = s * 1664525u + 1013904223u; line = (uint32_t)(((uint64_t)s * lines) >> 32); }
        float* p = buf + (size_t)line * 32 + lane * 4;
        const float v = 1.0f + lane;
        if (MODE == 0) red4(p, v, v, v, v);
        else if (MODE == 1) *reinterpret_cast<float4*>(p) = make_float4(v, v, v, v);
        else { atomicAdd(p, v); atomicAdd(p + 1, v); atomicAdd(p + 2, v); atomicAdd(p + 3, v); }
    }
}

// the forward's side of the same question: independent 128-byte line gathers (ld.global.nc.v4 per lane) from L2
__global__ void gather(const float* buf, float* out, uint32_t lines, int iters) {
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, lane = threadIdx.x & 7;
    uint32_t s = group * 2654435761u + 12345u;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < iters; i += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t line = (uint32_t)(((uint64_t)s * lines) >> 32);
            v[u] = __ldg(reinterpret_cast<const float4*>(buf + (size_t)line * 32 + lane * 4));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x == 123.456f) out[0] = acc.x + acc.y + acc.z + acc.w;
}

static double run_gather(const float* buf, float* out, uint32_t lines) {
    const int ctas = 148 * 8, threads = 256, iters = 512;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather<<<ctas, threads>>>(buf, out, lines, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) gather<<<ctas, threads>>>(buf, out, lines, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return 5.0 * ctas * threads * (double)iters * 16.0 / (ms * 1e-3) / 1e9;
}

template <int MODE>
static double run(float* buf, uint32_t lines, int sequential) {
    const int ctas = 148 * 8, threads = 256, iters = 512;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    scatter<MODE><<<ctas, threads>>>(buf, lines, iters, sequential);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) scatter<MODE><<<ctas, threads>>>(buf, lines, iters, sequential);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double bytes = 5.0 * ctas * threads * (double)iters * 16.0;
    return bytes / (ms * 1e-3) / 1e9;
}

int main() {
    const uint32_t big = 22223u * 8u, hot = 8192u;  // lines of 128 B
    float* buf;
    cudaMalloc(&buf, (size_t)big * 128);
    cudaMemset(buf, 0, (size_t)big * 128);
    printf("{\"unit\": \"GB/s of payload\", \"line_bytes\": 128,\n");
    printf(" \"red_v4_random_22.8MB\": %.0f, \"red_v4_random_1MB\": %.0f, \"red_v4_sequential_22.8MB\": %.0f,\n", run<0>(buf, big, 0), run<0>(buf, hot, 0), run<0>(buf, big, 1));
    printf(" \"atomicAdd_x4_random_22.8MB\": %.0f,\n", run<2>(buf, big, 0));
    printf(" \"ld_v4_random_22.8MB\": %.0f, \"ld_v4_random_1MB\": %.0f,\n", run_gather(buf, buf, big), run_gather(buf, buf, hot));
    printf(" \"st_v4_random_22.8MB\": %.0f, \"st_v4_sequential_22.8MB\": %.0f}\n", run<1>(buf, big, 0), run<1>(buf, big, 1));
    return cudaGetLastError() != cudaSuccess;
}
