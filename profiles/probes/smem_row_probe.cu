// Developer probe for DESIGN.md §8 item 1: if the gathered operand of the E / M pass were staged in shared memory, what would the
// row reads reach?  Two mappings, no global traffic in the timed loops (row indices come from a per-lane / per-warp LCG):
//   lane_rows<LAYOUT> : lane = cell.  Every lane reads ITS OWN random 32-byte row (8 columns) per step with two LDS.128 and adds
//                       it to 8 accumulators — the "CTA = 512 cells x 8 columns, tile [R][8] x 2 tables" mapping.
//                       LAYOUT 0 = [row][8] contiguous, 1 = two planes [row][4] (halves apart), 2 = conflict-free control
//                       (row = step + lane: what the banks give when nothing collides).
//   warp_row          : warp = cell.  All lanes read ONE random 1 KB row (256 columns, lane i: 16 B at i*16 and 512 + i*16) per step
//                       — today's mapping with the row coming from shared memory instead of L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_row_probe smem_row_probe.cu && ./smem_row_probe
// Reported: aggregate TB/s of row bytes and bytes per clock per SM at the nominal 1965 MHz (128 B/clk/SM is the LDS ceiling).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

extern __shared__ float4 smem4[];

__device__ __forceinline__ void add4(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }

template <int LAYOUT>
__global__ void lane_rows(uint32_t rows, uint32_t iters, float4* out, uint32_t seed) {
    // tile: rows x 32 B, filled with something non-trivial
    for (uint32_t i = threadIdx.x; i < rows * 2; i += blockDim.x) smem4[i] = make_float4(i * 1e-6f, 1.0f, -1.0f, 0.5f);
    __syncthreads();
    const uint32_t mask = rows - 1;                       // rows is a power of two
    uint32_t s = seed ^ ((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u);
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll 4
    for (uint32_t it = 0; it < iters; ++it) {
        uint32_t r;
        if (LAYOUT == 2) r = (it * 32 + lane) & mask;
        else { s = s * 1664525u + 1013904223u; r = (s >> 9) & mask; }
        float4 v0, v1;
        if (LAYOUT == 1) { v0 = smem4[r]; v1 = smem4[rows + r]; }
        else if (LAYOUT == 2) { v0 = smem4[r]; v1 = smem4[rows + r]; }
        else { v0 = smem4[2 * r]; v1 = smem4[2 * r + 1]; }
        add4(a0, v0);
        add4(a1, v1);
    }
    add4(a0, a1);
    if (a0.x == 123.456f) out[0] = a0;
}

__global__ void warp_row(uint32_t rows, uint32_t iters, float4* out, uint32_t seed) {
    // tile: rows x 1 KB (64 float4 per row)
    for (uint32_t i = threadIdx.x; i < rows * 64; i += blockDim.x) smem4[i] = make_float4(i * 1e-6f, 1.0f, -1.0f, 0.5f);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t s = seed ^ (((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2654435761u);   // warp-uniform
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0;
#pragma unroll 4
    for (uint32_t it = 0; it < iters; ++it) {
        s = s * 1664525u + 1013904223u;
        const uint32_t r = (uint32_t)(((uint64_t)(s >> 4) * rows) >> 28);
        add4(a0, smem4[r * 64 + lane]);
        add4(a1, smem4[r * 64 + 32 + lane]);
    }
    add4(a0, a1);
    if (a0.x == 123.456f) out[0] = a0;
}

template <typename K>
int time_it(const char* what, K kernel, uint32_t rows, size_t smem_bytes, int ctas_per_sm, int threads, double bytes_per_thread_step, float4* out) {
    const int sms = 148;
    const uint32_t iters = 8192;
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem_bytes));
    if (occ < ctas_per_sm) { printf("%-44s : only %d CTAs/SM fit, wanted %d\n", what, occ, ctas_per_sm); return 0; }
    dim3 grid(sms * ctas_per_sm), block(threads);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    kernel<<<grid, block, smem_bytes>>>(rows, iters, out, 1);
    CK(cudaEventRecord(a));
    kernel<<<grid, block, smem_bytes>>>(rows, iters, out, 2);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double total = (double)grid.x * threads * iters * bytes_per_thread_step;
    printf("%-44s smem %3zu KB x %d CTAs/SM (%2d warps/SM): %6.2f TB/s  %6.1f B/clk/SM  (%.3f ms)\n", what, smem_bytes >> 10, ctas_per_sm,
           ctas_per_sm * threads / 32, total / ms / 1e9, total / (ms * 1e-3) / sms / 1.965e9, ms);
    return 0;
}

int main() {
    float4* out;
    CK(cudaMalloc(&out, 64));
    // lane = cell: 32 B per lane and step
    for (int c : {1, 3}) {
        const uint32_t rows = c == 1 ? 4096 : 2048;                     // 128 KB or 64 KB of 32-byte rows
        const size_t bytes = (size_t)rows * 32;
        if (time_it("lane rows, [row][8] contiguous", lane_rows<0>, rows, bytes, c, 512, 32.0, out)) return 1;
        if (time_it("lane rows, two [row][4] planes", lane_rows<1>, rows, bytes, c, 512, 32.0, out)) return 1;
        if (time_it("lane rows, conflict-free control", lane_rows<2>, rows, bytes, c, 512, 32.0, out)) return 1;
    }
    // warp = cell: 1 KB per warp and step = 32 B per lane and step
    for (int c : {1, 3}) {
        const uint32_t rows = c == 1 ? 128 : 64;                        // 128 KB or 64 KB of 1 KB rows
        if (time_it("warp row, one 1 KB row per warp", warp_row, rows, (size_t)rows * 1024, c, 512, 32.0, out)) return 1;
    }
    CK(cudaDeviceSynchronize());
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
