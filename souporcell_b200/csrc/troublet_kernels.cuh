// Device kernels of the doublet caller.  Reference: /root/reference/troublet/src/main.rs ("T:").
// f64 throughout, no FMA contraction (-fmad=false), IEEE division; `log` is the CUDA libm's (<= 1 ulp), so results agree
// with the CPU restatement to ~1e-13 relative, not bit for bit (tests/test_troublet.py states the tolerance).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace troublet_dev {

__constant__ double c_ln_fact[171];     // ln(n!) for n <= 170 with the HOST libm (statrs FCACHE[n].ln())
__constant__ double c_gamma_dk[11];     // statrs GAMMA_DK

struct Consts {
    double p_soup, ln_singlet_prior, ln_doublet_prior;   // T:63, T:95: ln(1 - prior), ln(prior) from the host libm
    uint32_t min_locus_lines;                             // T:66: loci with fewer lines than this are skipped (20)
};

__device__ __forceinline__ double ln_gamma(double x) {   // statrs gamma::ln_gamma, x >= 0.5 branch (every argument here is >= 1)
    double s = c_gamma_dk[0];
#pragma unroll
    for (int i = 1; i < 11; ++i) s = s + c_gamma_dk[i] / (x + (double)i - 1.0);
    return log(s) + 0.6207822376352452223455184457816472122518527279025978 + (x - 0.5) * log((x - 0.5 + 10.900511) / 2.718281828459045235360287471352662498);
}
__device__ __forceinline__ double ln_beta(double a, double b) { return ln_gamma(a) + ln_gamma(b) - ln_gamma(a + b); }
__device__ __forceinline__ double ln_factorial(uint64_t x) { return x <= 170 ? c_ln_fact[x] : ln_gamma((double)x + 1.0); }
__device__ __forceinline__ double ln_binomial(uint64_t n, uint64_t k) { return k > n ? -INFINITY : ln_factorial(n) - ln_factorial(k) - ln_factorial(n - k); }

// one term of T:77-81 / T:120-124
__device__ __forceinline__ double term(uint64_t ref, uint64_t alt, double af, double count) {
    return ln_binomial(ref + alt, ref) + ln_beta((double)ref + af * count + 1.0, (double)alt + (1.0 - af) * count + 1.0) -
           ln_beta(af * count + 1.0, (1.0 - af) * count + 1.0);
}
__device__ __forceinline__ double clamp_frac(double f) { return fmax(fmin(f, 0.99), 0.01); }

struct Args {
    const uint64_t* row_ptr; const uint32_t* locus; const uint32_t* ref; const uint32_t* alt;   // raw CSR, file order inside a cell
    const unsigned long long* S0;     // [L][K][2] (ref, alt) sums of every cluster's cells as loaded (T:305-313)
    const unsigned long long* A;      // [L][K][2] the same minus the cells removed as doublets so far (T:215-224)
    const double* soup;               // [L] T:335-342
    const uint32_t* locus_lines;      // [L] T:303-304
    uint32_t N; int K; int round;
    Consts c;
};

// A warp owns one (cell, cluster) or (cell, pair) sum.  Its lanes evaluate the terms of 32 consecutive entries in parallel (the
// expensive part: six Lanczos ln_gamma each), then the terms are added to the running sum one by one in entry order, so the f64
// sum is the reference's sequential `log_prob += ...` (T:82, T:125) — a skipped locus contributes +0.0, which changes nothing.
__device__ __forceinline__ double ordered_add(double lp, double term) {
#pragma unroll
    for (int i = 0; i < 32; ++i) lp += __shfl_sync(0xffffffffu, term, i);
    return lp;
}

// counts[(c1, c2)] of the reference's K x K tables, never materialised: (c, c) = A[c]; (a, b) = A[a] + S0[b]  (T:305-313, T:219-222)
__global__ void __launch_bounds__(128) singlet_kernel(Args a, double* __restrict__ out /*[N][K]*/) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp = (cell, cluster), cluster fastest
    if (i >= (uint64_t)a.N * a.K) return;
    const uint32_t cell = (uint32_t)(i / a.K), lane = threadIdx.x & 31;
    const int c = (int)(i - (uint64_t)cell * a.K);
    double lp = a.c.ln_singlet_prior;
    const uint64_t j1 = a.row_ptr[cell + 1];
    for (uint64_t j0 = a.row_ptr[cell]; j0 < j1; j0 += 32) {
        const uint64_t j = j0 + lane;
        double t = 0.0;
        if (j < j1) {
            const uint32_t l = a.locus[j];
            if (a.locus_lines[l] >= a.c.min_locus_lines) {
                const size_t o = ((size_t)l * a.K + c) * 2;
                const double r = (double)a.A[o], al = (double)a.A[o + 1], tot = r + al;
                double frac = 0.5;                               // T:375; with tot == 0 the value cannot matter (it multiplies count == 0)
                if (tot > 0.0) frac = a.round == 0 ? clamp_frac(0.5 * (r / tot) + 0.5 * (r / tot)) : clamp_frac(r / tot);   // T:359-363 / T:232
                const double af = (1.0 - a.c.p_soup) * frac + a.c.p_soup * a.soup[l];
                t = term(a.ref[j], a.alt[j], af, tot);
            }
        }
        lp = ordered_add(lp, t);
    }
    if (lane == 0) out[i] = lp;
}

constexpr int PAIRS = 6;   // pairs among the top min(K, 4) singlets (T:91-92)
__global__ void __launch_bounds__(128) doublet_kernel(Args a, const int32_t* __restrict__ pairs /*[N][6][2], -1 = unused*/, double* __restrict__ out /*[N][6]*/) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp = (cell, pair)
    if (i >= (uint64_t)a.N * PAIRS) return;
    const uint32_t cell = (uint32_t)(i / PAIRS), lane = threadIdx.x & 31;
    const int c1 = pairs[i * 2], c2 = pairs[i * 2 + 1];
    if (c1 < 0) { if (lane == 0) out[i] = -INFINITY; return; }
    double lp = a.c.ln_doublet_prior;
    const uint64_t j1 = a.row_ptr[cell + 1];
    for (uint64_t j0 = a.row_ptr[cell]; j0 < j1; j0 += 32) {
        const uint64_t j = j0 + lane;
        double t = 0.0;
        if (j < j1) {
            const uint32_t l = a.locus[j];
            if (a.locus_lines[l] >= a.c.min_locus_lines) {
                const size_t o1 = ((size_t)l * a.K + c1) * 2, o2 = ((size_t)l * a.K + c2) * 2;
                const double r1 = (double)a.A[o1], a1 = (double)a.A[o1 + 1], r3 = (double)a.A[o2], a3 = (double)a.A[o2 + 1];
                double frac = 0.5;
                if (a.round == 0) {                              // T:351-377 on the counts as loaded
                    const double t1 = r1 + a1, t2 = r3 + a3;
                    if (t1 > 0.0) frac = t2 > 0.0 ? clamp_frac(0.5 * (r1 / t1) + 0.5 * (r3 / t2)) : clamp_frac(r1 / t1);
                } else {                                         // T:226-235 on counts[(c1, c2)] = A[c1] + S0[c2]
                    const double r = r1 + (double)a.S0[o2], al = a1 + (double)a.S0[o2 + 1];
                    if (r + al > 0.0) frac = clamp_frac(r / (r + al));
                }
                const double af = (1.0 - a.c.p_soup) * frac + a.c.p_soup * a.soup[l];
                const double count = ((r1 + a1) + (r3 + a3)) / 2.0;   // T:117
                t = term(a.ref[j], a.alt[j], af, count);
            }
        }
        lp = ordered_add(lp, t);
    }
    if (lane == 0) out[i] = lp;
}

// S0[l][cluster(cell)] += (ref, alt); lines per locus; exact integer atomics (the reference adds integer-valued f64)
__global__ void tally_kernel(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ locus, const uint32_t* __restrict__ ref,
                             const uint32_t* __restrict__ alt, const int32_t* __restrict__ cluster, unsigned long long* S0, uint32_t* lines,
                             uint32_t N, int K) {
    const uint32_t cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= N) return;
    const int c = cluster[cell];
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t j = row_ptr[cell] + lane; j < row_ptr[cell + 1]; j += 32) {
        atomicAdd(lines + locus[j], 1u);
        unsigned long long* o = S0 + ((size_t)locus[j] * K + c) * 2;
        if (ref[j]) atomicAdd(o, (unsigned long long)ref[j]);
        if (alt[j]) atomicAdd(o + 1, (unsigned long long)alt[j]);
    }
}
__global__ void soup_kernel(const unsigned long long* __restrict__ S0, double* __restrict__ soup, uint32_t L, int K) {   // T:335-342
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    unsigned long long r = 0, a = 0;
    for (int c = 0; c < K; ++c) { r += S0[((size_t)l * K + c) * 2]; a += S0[((size_t)l * K + c) * 2 + 1]; }
    const double rd = (double)r, ad = (double)a;
    soup[l] = rd + ad > 0.0 ? rd / (rd + ad) : 0.0;
}
// T:215-224: take the removed cells out of their cluster's counts
__global__ void remove_kernel(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ locus, const uint32_t* __restrict__ ref,
                              const uint32_t* __restrict__ alt, const int32_t* __restrict__ cluster, const uint32_t* __restrict__ removed,
                              uint32_t n_removed, unsigned long long* A, int K) {
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_removed) return;
    const uint32_t cell = removed[w], lane = threadIdx.x & 31;
    const int c = cluster[cell];
    for (uint64_t j = row_ptr[cell] + lane; j < row_ptr[cell + 1]; j += 32) {
        unsigned long long* o = A + ((size_t)locus[j] * K + c) * 2;
        if (ref[j]) atomicAdd(o, 0ull - (unsigned long long)ref[j]);
        if (alt[j]) atomicAdd(o + 1, 0ull - (unsigned long long)alt[j]);
    }
}

}  // namespace troublet_dev
