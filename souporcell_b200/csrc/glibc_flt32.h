// Bit-exact clones of glibc 2.39's x86-64 `logf` / `expf` (the FMA ifunc variants,
// sysdeps/ieee754/flt-32/e_logf.c, e_expf.c — ARM optimized-routines algorithm), in
// double precision, callable from host and device code.
//
// Why: the reference computes ln(theta), ln(1-theta), exp() with Rust's f32::ln/exp,
// which lower to the system libm (souporcell/src/main.rs:417-418, 430-431, 451).  glibc's
// results differ from the correctly rounded value for ~0.15 % (logf) / ~0.04 % (expf) of
// inputs, so neither CUDA's logf nor (float)log((double)x) reproduces them.  Running the
// same double-precision polynomial with the same FMA contraction does, bit for bit
// (exhaustively checked on the CPU against libm in tests/test_glibc_math.py).
//
// Constants are glibc's published tables (__logf_data, __exp2f_data with N = 16 / 32).
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SOUP_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define SOUP_HD inline
#endif

namespace soupmath {

#define SOUP_LOGF_TAB                                                                                   \
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,        \
     0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,        \
     0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,        \
     0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,        \
     0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1.0000000000000p+0, 0x0.0p+0,                      \
     0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,         \
     0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,         \
     0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}

#define SOUP_EXP2F_TAB                                                                                  \
    {0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,       \
     0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,       \
     0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,       \
     0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,       \
     0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,       \
     0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,       \
     0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,       \
     0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL}

static const double h_logf_tab[32] = SOUP_LOGF_TAB;      // {invc, logc} x 16
static const uint64_t h_exp2f_tab[32] = SOUP_EXP2F_TAB;  // bits(2^(i/32)) - (i << 47)
#ifdef __CUDACC__
static __device__ const double d_logf_tab[32] = SOUP_LOGF_TAB;
static __device__ const uint64_t d_exp2f_tab[32] = SOUP_EXP2F_TAB;
#endif

SOUP_HD const double* logf_tab() {
#ifdef __CUDA_ARCH__
    return d_logf_tab;
#else
    return h_logf_tab;
#endif
}
SOUP_HD const uint64_t* exp2f_tab() {
#ifdef __CUDA_ARCH__
    return d_exp2f_tab;
#else
    return h_exp2f_tab;
#endif
}

SOUP_HD double fma_d(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
SOUP_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
SOUP_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
SOUP_HD uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
SOUP_HD double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

// glibc __logf_fma
SOUP_HD float glibc_logf(float x) {
    uint32_t ix = f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return u2f(0xff800000u);          // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                   // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return u2f(0x7fc00000u);   // x < 0 or NaN
        ix = f2u(x * 0x1p23f);                             // subnormal: normalise
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    const double* T = logf_tab();
    double invc = T[2 * i], logc = T[2 * i + 1];
    double z = (double)u2f(iz);
    double r = fma_d(z, invc, -1.0);
    double y0 = fma_d((double)k, 0x1.62e42fefa39efp-1, logc);
    double r2 = r * r;
    double y = fma_d(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
    y = fma_d(-0x1.00ea348b88334p-2, r2, y);
    y = fma_d(y, r2, y0 + r);
    return (float)y;
}

// glibc __expf_fma
SOUP_HD float glibc_expf(float x) {
    uint32_t ux = f2u(x);
    uint32_t abstop = (ux >> 20) & 0x7ffu;
    if (abstop >= 0x42bu) {                                // |x| >= 88 or NaN
        if (ux == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8u) return x + x;
        if (x > 0x1.62e42ep6f) return u2f(0x7f800000u);    // overflow
        if (x < -0x1.9fe368p6f) return 0.0f;               // underflow
        if (x < -0x1.9d1d9ep6f) return u2f(1u);            // may_uflow: 0x1.4p-75f^2 -> 2^-149
    }
    double xd = (double)x;
    const double invln2n = 0x1.71547652b82fep+5;           // 32 / ln 2
    const double shift = 0x1.8p+52;
    double kd = fma_d(invln2n, xd, shift);
    uint64_t ki = d2u(kd);
    kd -= shift;
    double r = fma_d(invln2n, xd, -kd);
    uint64_t t = exp2f_tab()[ki & 31u];
    t += ki << 47;
    double s = u2d(t);
    double z = fma_d(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13);   // C0/N^3, C1/N^2
    double r2 = r * r;
    double y = fma_d(0x1.62e42ff0c52d6p-6, r, 1.0);                      // C2/N
    y = fma_d(z, r2, y);
    y = y * s;
    return (float)y;
}

}  // namespace soupmath
