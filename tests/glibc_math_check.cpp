// Brute-force check: soupmath::glibc_logf / glibc_expf (souporcell_b200/csrc/glibc_flt32.h)
// against this machine's libm logf / expf over every `stride`-th float bit pattern.
// usage: glibc_math_check <stride> ; prints "logf <n> <mismatches>" and "expf <n> <mismatches>".
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../souporcell_b200/csrc/glibc_flt32.h"

static bool same(float a, float b) {
    if (std::isnan(a) && std::isnan(b)) return true;
    return soupmath::f2u(a) == soupmath::f2u(b);
}

int main(int argc, char** argv) {
    uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 4;
    std::atomic<uint64_t> n_log{0}, bad_log{0}, n_exp{0}, bad_exp{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&, t] {
            uint64_t nl = 0, bl = 0, ne = 0, be = 0;
            for (uint64_t u = t * stride; u <= 0xffffffffull; u += (uint64_t)nt * stride) {
                float x = soupmath::u2f((uint32_t)u);
                if (u <= 0x7f800000ull) { ++nl; if (!same(logf(x), soupmath::glibc_logf(x))) { if (bl < 3) fprintf(stderr, "logf mismatch x=%a\n", x); ++bl; } }
                ++ne;
                if (!same(expf(x), soupmath::glibc_expf(x))) { if (be < 3) fprintf(stderr, "expf mismatch x=%a libm=%a clone=%a\n", x, expf(x), soupmath::glibc_expf(x)); ++be; }
            }
            n_log += nl; bad_log += bl; n_exp += ne; bad_exp += be;
        });
    for (auto& th : pool) th.join();
    printf("logf %llu %llu\n", (unsigned long long)n_log.load(), (unsigned long long)bad_log.load());
    printf("expf %llu %llu\n", (unsigned long long)n_exp.load(), (unsigned long long)bad_exp.load());
    return (bad_log.load() || bad_exp.load()) ? 1 : 0;
}
