// Developer probe: issue rate of scalar FADD/FMUL chains vs the packed add.rn.f32x2 / mul.rn.f32x2 of sm_100,
// and a bit-exactness check of the packed forms against the scalar ones (round-to-nearest, no FTZ).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_probe f32x2_probe.cu && ./f32x2_probe
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ void add2(float& x0, float& x1, float a0, float a1) {
    uint64_t x, a;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(a));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
}
__device__ __forceinline__ void mul2(float& x0, float& x1, float a0, float a1) {
    uint64_t x, a;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(a));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
}

// mode 0: 8 scalar FADD chains; 1: 4 packed add chains; 2: 8 x (FMUL + FADD); 3: 4 x (mul2 + add2)
template <int MODE>
__global__ void __launch_bounds__(256) chains(const float* __restrict__ in, float* out, int iters) {
    float acc[8], t[8];
    for (int i = 0; i < 8; ++i) { acc[i] = in[threadIdx.x + i * 256]; t[i] = in[threadIdx.x + 2048 + i * 256]; }
    float s = in[5000];
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fadd_rn(acc[i], t[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) add2(acc[i], acc[i + 1], t[i], t[i + 1]);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __fadd_rn(acc[i], __fmul_rn(t[i], s));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float m0 = t[i], m1 = t[i + 1];
                mul2(m0, m1, s, s);
                add2(acc[i], acc[i + 1], m0, m1);
            }
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void exact(const float* __restrict__ a, const float* __restrict__ b, uint32_t* bad, int n) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 >= n) return;
    float x0 = a[i], x1 = a[i + 1];
    add2(x0, x1, b[i], b[i + 1]);
    float y0 = a[i], y1 = a[i + 1];
    mul2(y0, y1, b[i], b[i + 1]);
    const float r0 = __fadd_rn(a[i], b[i]), r1 = __fadd_rn(a[i + 1], b[i + 1]);
    const float m0 = __fmul_rn(a[i], b[i]), m1 = __fmul_rn(a[i + 1], b[i + 1]);
    if (__float_as_uint(x0) != __float_as_uint(r0) || __float_as_uint(x1) != __float_as_uint(r1)) atomicAdd(bad, 1);
    if (__float_as_uint(y0) != __float_as_uint(m0) || __float_as_uint(y1) != __float_as_uint(m1)) atomicAdd(bad + 1, 1);
}

template <int MODE>
void run(const float* in, float* out, const char* what, double flop_per_iter) {
    const int iters = 20000, blocks = 148 * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    chains<MODE><<<blocks, 256>>>(in, out, iters);
    cudaEventRecord(a);
    chains<MODE><<<blocks, 256>>>(in, out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("%-34s %.3f ms  %.2f T lane-op/s\n", what, ms, (double)blocks * 256 * iters * flop_per_iter / ms / 1e9);
}

int main() {
    const int n = 1 << 22;
    float *in, *out, *a, *b;
    uint32_t* bad;
    cudaMalloc(&in, 8192 * 4); cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&bad, 8);
    float* h = new float[n];
    for (int i = 0; i < 8192; ++i) h[i] = 1.0f + (i % 97) * 1e-3f;
    cudaMemcpy(in, h, 8192 * 4, cudaMemcpyHostToDevice);
    run<0>(in, out, "8 x FADD", 8);
    run<1>(in, out, "4 x add.rn.f32x2", 8);
    run<2>(in, out, "8 x (FMUL, FADD)", 16);
    run<3>(in, out, "4 x (mul.f32x2, add.f32x2)", 16);
    // exactness on random bit patterns (denormals, infinities and NaNs included)
    uint32_t s = 12345;
    uint32_t* hu = reinterpret_cast<uint32_t*>(h);
    for (int rep = 0; rep < 2; ++rep) {
        for (int i = 0; i < n; ++i) {
            s = s * 1664525u + 1013904223u;
            uint32_t v = s;
            if ((i & 7) == 0) v = (v & 0x807fffffu) | (((s >> 9) % 3 == 0 ? 0u : 0x7f000000u));   // denormals / huge
            if ((i & 7) == 1) v = (v & 0x80000000u) | (0x3f800000u + (s & 0xfffff));              // near 1
            hu[i] = v;
        }
        cudaMemcpy(rep ? b : a, h, n * 4, cudaMemcpyHostToDevice);
    }
    cudaMemset(bad, 0, 8);
    exact<<<n / 2 / 256, 256>>>(a, b, bad, n);
    uint32_t hb[2];
    cudaMemcpy(hb, bad, 8, cudaMemcpyDeviceToHost);
    printf("packed vs scalar mismatches (NaN payloads count): add %u mul %u of %d pairs\n", hb[0], hb[1], n / 2);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
