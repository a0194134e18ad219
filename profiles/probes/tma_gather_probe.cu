// Developer probe (round 2): row gathers of the E / M passes through TMA `tile::gather4` instead of per-lane LDG.128.
//
// The passes gather one 1 KB row (256 f32 columns) of a table per matrix entry and add it to 8 accumulators per lane.  Today every lane
// computes the row address (IMAD.WIDE) and issues two LDG.128; the data comes back through the L1 data pipe into registers.  With
// `cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4` ONE elected lane hands four row indices to the TMA unit, the four rows
// land in shared memory (mbarrier-tracked, no LSU address work, no registers held while in flight), and the warp reads them back with
// LDS.128.  Same access pattern for both paths: every warp walks its own list of random row indices of a [rows][256] f32 table,
// strictly in order, 4 rows per step, a ring of STAGES steps in flight.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_gather_probe tma_gather_probe.cu && ./tma_gather_probe
//
// Output: TB/s of gathered rows per (path, table size, warps per CTA, stages); the sums of both paths must agree (checked).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int ROW_FLOATS = 256;   // 1 KB rows: the widest box a tensor map without swizzle takes
constexpr int ROW_BYTES = ROW_FLOATS * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000ll) __trap();   // ~2 s: never hang the GPU
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

// ---- TMA path.  Each warp owns STAGES stages of 4 rows (4 KB each) and one mbarrier per stage.
template <int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32) gather_tma(const __grid_constant__ CUtensorMap map, const uint32_t* __restrict__ idx, uint32_t n_per_warp, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[WARPS][STAGES];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage0 = reinterpret_cast<float*>(smem_raw) + (size_t)warp * STAGES * 4 * ROW_FLOATS;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[warp][s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t gw = blockIdx.x * WARPS + warp;
    const uint32_t* p = idx + (size_t)gw * n_per_warp;
    const uint32_t n_steps = n_per_warp / 4;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // lane j of a chunk holds index j; the four indices of a step are broadcast to the issuing lane (the E walk's SHFL broadcast)
    uint32_t w = __ldg(p + lane), w_next = 0;
    auto issue = [&](uint32_t step, uint32_t word) {   // all lanes take part in the shuffles, lane 0 issues
        const uint32_t e = (step & 7u) * 4u;
        const int r0 = (int)__shfl_sync(0xffffffffu, word, e), r1 = (int)__shfl_sync(0xffffffffu, word, e + 1);
        const int r2 = (int)__shfl_sync(0xffffffffu, word, e + 2), r3 = (int)__shfl_sync(0xffffffffu, word, e + 3);
        if (lane == 0) {
            const uint32_t s = step % STAGES;
            mbar_expect_tx(&bars[warp][s], 4 * ROW_BYTES);
            tma_gather4(stage0 + (size_t)s * 4 * ROW_FLOATS, &map, 0, r0, r1, r2, r3, &bars[warp][s]);
        }
    };
    static_assert(STAGES <= 8, "the prologue stays inside the first chunk of 32 indices");
    for (uint32_t s = 0; s < STAGES && s < n_steps; ++s) issue(s, w);
    for (uint32_t step = 0; step < n_steps; ++step) {
        if ((step & 7u) == 0 && step + 8 < n_steps) w_next = __ldg(p + (step + 8) * 4 + lane);   // next chunk of 32 indices, one step-group ahead
        const uint32_t s = step % STAGES;
        mbar_wait(&bars[warp][s], (step / STAGES) & 1u);
        const float* rows = stage0 + (size_t)s * 4 * ROW_FLOATS;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(rows + r * ROW_FLOATS + lane * 4);
            const float4 b = *reinterpret_cast<const float4*>(rows + r * ROW_FLOATS + 128 + lane * 4);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        }
        __syncwarp();                                   // every lane has read the stage: it may be overwritten
        const uint32_t nxt = step + STAGES;
        if (nxt < n_steps) issue(nxt, (nxt >> 3) == (step >> 3) ? w : w_next);
        if ((step & 7u) == 7u) w = w_next;
    }
    float t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += acc[i];
    out[(size_t)gw * 32 + lane] = t;
}

// ---- LDG path: the E walk's mapping (SHFL-broadcast row index, IMAD.WIDE, two LDG.128 per lane, ring of U rows in registers)
template <int WARPS, int U>
__global__ void __launch_bounds__(WARPS * 32) gather_ldg(const float* __restrict__ table, const uint32_t* __restrict__ idx, uint32_t n_per_warp, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];   // only to pin the residency to the TMA variant's
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * WARPS + warp;
    const uint32_t* p = idx + (size_t)gw * n_per_warp;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float4 ra[U], rb[U];
    auto fetch = [&](int u, uint32_t row) {
        const float4* q = reinterpret_cast<const float4*>(table + (size_t)row * ROW_FLOATS) + lane;
        ra[u] = __ldg(q); rb[u] = __ldg(q + 32);
    };
    for (uint32_t base = 0; base < n_per_warp; base += 32) {
        const uint32_t w = __ldg(p + base + lane);
#pragma unroll
        for (int u = 0; u < U; ++u) fetch(u, __shfl_sync(0xffffffffu, w, u));
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int u = e % U;
            acc[0] += ra[u].x; acc[1] += ra[u].y; acc[2] += ra[u].z; acc[3] += ra[u].w;
            acc[4] += rb[u].x; acc[5] += rb[u].y; acc[6] += rb[u].z; acc[7] += rb[u].w;
            if (e + U < 32) fetch(u, __shfl_sync(0xffffffffu, w, e + U));
        }
    }
    float t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += acc[i];
    out[(size_t)gw * 32 + lane] = t;
    if (t == 123.456f) smem_raw[0] = 1;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int WARPS, int STAGES>
void run_pair(EncodeTiled encode, float* table, size_t bytes, const uint32_t* d_idx_unused, float* out_a, float* out_b, int sms) {
    const uint32_t n_rows = (uint32_t)(bytes / ROW_BYTES);
    const uint32_t n_per_warp = 1024;
    // residency: the TMA variant's shared memory decides how many CTAs fit; the LDG variant is launched with the same amount
    const size_t smem = (size_t)WARPS * STAGES * 4 * ROW_BYTES;
    const int ctas_per_sm = (int)std::min<size_t>(32 / (WARPS > 8 ? 2 : 1), (227 * 1024) / (smem + 1024));
    const uint32_t grid = (uint32_t)(sms * ctas_per_sm * 4);
    const size_t n_idx = (size_t)grid * WARPS * n_per_warp;
    std::vector<uint32_t> h(n_idx);
    uint32_t s = 777;
    for (auto& x : h) { s = s * 1664525u + 1013904223u; x = (uint32_t)(((uint64_t)(s >> 4) * n_rows) >> 28); }
    uint32_t* d_idx;
    CK(cudaMalloc(&d_idx, n_idx * 4));
    CK(cudaMemcpy(d_idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
    CUtensorMap map;
    const cuuint64_t gdim[2] = {ROW_FLOATS, n_rows};
    const cuuint64_t gstride[1] = {ROW_BYTES};
    const cuuint32_t box[2] = {ROW_FLOATS, 1};       // gather4: the box is ONE row; the instruction names four of them
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    CK(cudaFuncSetAttribute(gather_tma<WARPS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(gather_ldg<WARPS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms_t = 0, ms_l = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        gather_tma<WARPS, STAGES><<<grid, WARPS * 32, smem>>>(map, d_idx, n_per_warp, out_a);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms_t, a, b);
        cudaEventRecord(a);
        gather_ldg<WARPS, 4><<<grid, WARPS * 32, smem>>>(table, d_idx, n_per_warp, out_b);
        cudaEventRecord(b);
        CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms_l, a, b);
    }
    CK(cudaGetLastError());
    // both paths add the same rows in the same order
    const size_t n_out = (size_t)grid * WARPS * 32;
    std::vector<float> ha(n_out), hb(n_out);
    CK(cudaMemcpy(ha.data(), out_a, n_out * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), out_b, n_out * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n_out; ++i) bad += ha[i] != hb[i];
    const double total = (double)n_idx * ROW_BYTES;
    printf("table %4zu MB  %2d warps/CTA x %d CTAs/SM, %d stages (%2d rows in flight per warp): TMA gather4 %6.2f TB/s (%.3f ms)   LDG.128 U=4 %6.2f TB/s (%.3f ms)   sums %s\n",
           bytes >> 20, WARPS, ctas_per_sm, STAGES, STAGES * 4, total / ms_t / 1e9, ms_t, total / ms_l / 1e9, ms_l, bad ? "DIFFER" : "equal");
    fflush(stdout);
    cudaFree(d_idx);
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    EncodeTiled encode = (EncodeTiled)fn;
    const size_t maxb = 256ull << 20;
    float *table, *out_a, *out_b;
    CK(cudaMalloc(&table, maxb));
    CK(cudaMalloc(&out_a, 64 << 20));
    CK(cudaMalloc(&out_b, 64 << 20));
    {
        std::vector<uint32_t> h(maxb / 4);
        uint32_t s = 12345;
        for (auto& x : h) { s = s * 1664525u + 1013904223u; x = (s >> 9) | 0x3f800000u; }   // floats in [1, 2)
        CK(cudaMemcpy(table, h.data(), maxb, cudaMemcpyHostToDevice));
    }
    for (size_t mb : {16, 64, 256}) {
        const size_t bytes = mb << 20;
        run_pair<8, 1>(encode, table, bytes, nullptr, out_a, out_b, sms);
        run_pair<8, 2>(encode, table, bytes, nullptr, out_a, out_b, sms);
        run_pair<8, 4>(encode, table, bytes, nullptr, out_a, out_b, sms);
        run_pair<4, 4>(encode, table, bytes, nullptr, out_a, out_b, sms);
        run_pair<4, 8>(encode, table, bytes, nullptr, out_a, out_b, sms);
        run_pair<16, 2>(encode, table, bytes, nullptr, out_a, out_b, sms);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
