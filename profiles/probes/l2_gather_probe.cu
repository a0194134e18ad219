// Developer probe: what can row gathers from an L2-resident table reach on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_probe l2_gather_probe.cu && ./l2_gather_probe
// Modes: rows of ROWB bytes gathered by whole warps (lane i reads 16 B at i*16 [+512*k]), random row index,
// table sizes 16 / 64 / 256 MB, U independent gathers in flight per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int U, int VEC4_PER_LANE>
__global__ void gather(const float4* __restrict__ table, uint32_t n_rows, uint32_t iters, float4* out, uint32_t seed) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t s = seed ^ (warp * 2654435761u);
    float4 acc = make_float4(0, 0, 0, 0);
    const uint32_t row_f4 = 32 * VEC4_PER_LANE;
    for (uint32_t it = 0; it < iters; ++it) {
        float4 v[U][VEC4_PER_LANE];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = (uint32_t)(((uint64_t)(s >> 4) * n_rows) >> 28);
#pragma unroll
            for (int k = 0; k < VEC4_PER_LANE; ++k) v[u][k] = __ldg(table + (size_t)r * row_f4 + k * 32 + lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < VEC4_PER_LANE; ++k) { acc.x += v[u][k].x; acc.y += v[u][k].y; acc.z += v[u][k].z; acc.w += v[u][k].w; }
    }
    if (acc.x == 123.456f) out[0] = acc;
}

template <int U, int VPL>
void run(const float4* t, size_t bytes, float4* out, int blocks_per_sm, int threads) {
    int sms = 148;
    uint32_t n_rows = (uint32_t)(bytes / (512 * VPL));
    uint32_t iters = 2000 / U;
    dim3 grid(sms * blocks_per_sm), block(threads);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather<U, VPL><<<grid, block>>>(t, n_rows, iters, out, 1);
    cudaEventRecord(a);
    gather<U, VPL><<<grid, block>>>(t, n_rows, iters, out, 2);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double total = (double)grid.x * (threads / 32) * iters * U * 512.0 * VPL;
    printf("table %4zu MB row %4d B U=%d warps/SM=%2d : %.2f TB/s (%.3f ms)\n", bytes >> 20, 512 * VPL, U, blocks_per_sm * threads / 32, total / ms / 1e9, ms);
}

int main() {
    float4 *t, *out;
    size_t maxb = 512ull << 20;
    cudaMalloc(&t, maxb);
    cudaMalloc(&out, 64);
    cudaMemset(t, 0, maxb);
    for (size_t mb : {16, 64, 256, 512}) {
        size_t bytes = mb << 20;
        run<4, 1>(t, bytes, out, 2, 256);
        run<4, 1>(t, bytes, out, 4, 256);
        run<8, 1>(t, bytes, out, 4, 256);
        run<8, 1>(t, bytes, out, 8, 256);
        run<4, 2>(t, bytes, out, 4, 256);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
