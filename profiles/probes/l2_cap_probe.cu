// Developer probe (round 2): is ~12.4 TB/s the L2 -> SM ceiling for the row gathers of the E / M passes, and what made
// l2_gather_probe.cu read 19-20 TB/s?  Same gather as that probe, varying ONE thing at a time:
//   fill      : the table holds zeros (as in l2_gather_probe) or random bits
//   regs      : the kernel is compiled lean (full occupancy) or padded to 80 registers / 24 warps per SM like estep_kernel<8>
//   row bytes : 512 or 1024 per warp gather
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_cap_probe l2_cap_probe.cu && ./l2_cap_probe
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

template <int U, int VPL>
__device__ __forceinline__ float4 body(const float4* __restrict__ table, uint32_t n_rows, uint32_t iters, uint32_t seed, uint32_t row_f4) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t s = seed ^ (warp * 2654435761u);
    float4 acc = make_float4(0, 0, 0, 0);
    for (uint32_t it = 0; it < iters; ++it) {
        float4 v[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            const uint32_t r = (uint32_t)(((uint64_t)(s >> 4) * n_rows) >> 28);
#pragma unroll
            for (int k = 0; k < VPL; ++k) v[u][k] = __ldg(table + (size_t)r * row_f4 + k * 32 + lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < VPL; ++k) { acc.x += v[u][k].x; acc.y += v[u][k].y; acc.z += v[u][k].z; acc.w += v[u][k].w; }
    }
    return acc;
}
template <int U, int VPL>
__global__ void __launch_bounds__(256) gather_lean(const float4* __restrict__ table, uint32_t n_rows, uint32_t iters, float4* out, uint32_t seed, uint32_t row_f4) {
    float4 acc = body<U, VPL>(table, n_rows, iters, seed, row_f4);
    if (acc.x == 123.456f) out[0] = acc;
}
// 3 CTAs of 256 threads per SM = 24 warps, like estep_kernel<8> (80 registers): dynamic shared memory caps the residency
template <int U, int VPL>
__global__ void __launch_bounds__(256) gather_occ(const float4* __restrict__ table, uint32_t n_rows, uint32_t iters, float4* out, uint32_t seed, uint32_t row_f4) {
    extern __shared__ float pad[];
    float4 acc = body<U, VPL>(table, n_rows, iters, seed, row_f4);
    if (acc.x == 123.456f) { out[0] = acc; pad[0] = 1; }
}

template <int U, int VPL>
void run(const char* what, const float4* t, size_t bytes, float4* out, int warps_per_sm, uint32_t row_f4) {
    const int sms = 148, threads = 256;
    const uint32_t n_rows = (uint32_t)(bytes / (16ull * row_f4));
    const uint32_t iters = 2000 / U;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms = 0;
    dim3 grid(sms * 32 / 8 * 4);   // many CTAs: the hardware keeps the SMs full
    if (warps_per_sm >= 64) {
        gather_lean<U, VPL><<<grid, threads>>>(t, n_rows, iters, out, 1, row_f4);
        cudaEventRecord(a);
        gather_lean<U, VPL><<<grid, threads>>>(t, n_rows, iters, out, 2, row_f4);
        cudaEventRecord(b);
    } else {
        const int ctas = warps_per_sm / 8;
        const size_t smem = (227 * 1024) / ctas - 2048;   // so that exactly `ctas` CTAs fit
        cudaFuncSetAttribute(gather_occ<U, VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gather_occ<U, VPL><<<grid, threads, smem>>>(t, n_rows, iters, out, 1, row_f4);
        cudaEventRecord(a);
        gather_occ<U, VPL><<<grid, threads, smem>>>(t, n_rows, iters, out, 2, row_f4);
        cudaEventRecord(b);
    }
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    const double total = (double)grid.x * (threads / 32) * iters * U * 512.0 * VPL;
    printf("%-7s table %4zu MB row %4d B (stride %5u B) U=%d warps/SM<=%2d : %6.2f TB/s\n", what, bytes >> 20, 512 * VPL, row_f4 * 16, U, warps_per_sm, total / ms / 1e9);
}

int main() {
    float4 *t, *out;
    const size_t maxb = 256ull << 20;
    cudaMalloc(&t, maxb);
    cudaMalloc(&out, 64);
    for (int fill = 0; fill < 2; ++fill) {
        if (fill == 0) cudaMemset(t, 0, maxb);
        else {
            std::vector<uint32_t> h(maxb / 4);
            uint32_t s = 12345;
            for (auto& x : h) { s = s * 1664525u + 1013904223u; x = (s >> 9) | 0x3f800000u; }   // floats in [1, 2)
            cudaMemcpy(t, h.data(), maxb, cudaMemcpyHostToDevice);
        }
        const char* what = fill ? "random" : "zeros";
        for (size_t mb : {16, 64, 128}) {
            const size_t bytes = mb << 20;
            run<4, 1>(what, t, bytes, out, 64, 32);
            run<4, 2>(what, t, bytes, out, 64, 64);
            run<4, 2>(what, t, bytes, out, 24, 64);
            run<4, 2>(what, t, bytes, out, 24, 104);   // 1 KB of a 1664-byte row, as with 416 padded columns
            run<8, 2>(what, t, bytes, out, 24, 64);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
