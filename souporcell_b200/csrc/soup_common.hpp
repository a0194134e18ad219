// Internal shared declarations of libsoupgpu (host side).  Not installed.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/soupgpu.h"

namespace soup {

// ---- error plumbing: message of the last failing ctx-less call on this thread
void set_thread_error(const char* fmt, ...);
const char* thread_error();
std::string vformat(const char* fmt, va_list ap);

// ---- ChaCha (rand_chacha 0.9.0 ChaChaXRng: key = seed, 64-bit block counter, stream 0)
#ifdef __CUDACC__
#define SOUP_HOSTDEV __host__ __device__ __forceinline__
#else
#define SOUP_HOSTDEV inline
#endif

SOUP_HOSTDEV uint32_t rotl32(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }

// One 16-word keystream block.  `rounds` = 12 for rand 0.9.3 StdRng.
SOUP_HOSTDEV void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
    uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
    uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3], x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
    uint32_t x12 = (uint32_t)counter, x13 = (uint32_t)(counter >> 32), x14 = 0u, x15 = 0u;
#define SOUP_QR(a, b, c, d)                 \
    a += b; d ^= a; d = rotl32(d, 16);      \
    c += d; b ^= c; b = rotl32(b, 12);      \
    a += b; d ^= a; d = rotl32(d, 8);       \
    c += d; b ^= c; b = rotl32(b, 7);
    for (int r = 0; r < rounds; r += 2) {
        SOUP_QR(x0, x4, x8, x12) SOUP_QR(x1, x5, x9, x13) SOUP_QR(x2, x6, x10, x14) SOUP_QR(x3, x7, x11, x15)
        SOUP_QR(x0, x5, x10, x15) SOUP_QR(x1, x6, x11, x12) SOUP_QR(x2, x7, x8, x13) SOUP_QR(x3, x4, x9, x14)
    }
#undef SOUP_QR
    out[0] = x0 + 0x61707865u; out[1] = x1 + 0x3320646eu; out[2] = x2 + 0x79622d32u; out[3] = x3 + 0x6b206574u;
    out[4] = x4 + key[0]; out[5] = x5 + key[1]; out[6] = x6 + key[2]; out[7] = x7 + key[3];
    out[8] = x8 + key[4]; out[9] = x9 + key[5]; out[10] = x10 + key[6]; out[11] = x11 + key[7];
    out[12] = x12 + (uint32_t)counter; out[13] = x13 + (uint32_t)(counter >> 32); out[14] = x14; out[15] = x15;
}

inline void seed_to_key(const uint8_t seed[32], uint32_t key[8]) {
    for (int i = 0; i < 8; ++i)
        key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
                 ((uint32_t)seed[4 * i + 3] << 24);
}

// Sequential word reader (host): the StdRng view of the stream.
struct StdRng {
    uint32_t key[8];
    uint64_t block = 0;
    uint32_t buf[16];
    int idx = 16;
    static StdRng from_seed(const uint8_t seed[32]) {
        StdRng r;
        seed_to_key(seed, r.key);
        return r;
    }
    uint32_t next_u32() {
        if (idx >= 16) { chacha_block(key, block++, 12, buf); idx = 0; }
        return buf[idx++];
    }
    uint8_t gen_u8() { return (uint8_t)next_u32(); }                                     // StandardUniform for u8
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }          // 24-bit mantissa draw
};

// ---- bad_cluster_detection S:150-187 in two halves (cell-sharded processes add their counts up in between)
void count_cluster_cells(const float* best_lp, uint64_t n_cells, int32_t k, uint64_t* counts);
int32_t bad_clusters_from_counts(int32_t run, int32_t k, const uint64_t* counts, int32_t* out_idx, void* log_file);

// ---- host math / formatting
float ln_binomial_f32(uint32_t n, uint32_t k);        // statrs 0.11.0 ln_binomial -> f32
std::string format_f32(float v);                      // Rust `{}`
std::string format_f32_w8(float v);                   // Rust `{:8}`

// Rust f32::total_cmp key
SOUP_HOSTDEV int32_t total_key(float f) {
    int32_t b;
    __builtin_memcpy(&b, &f, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}

}  // namespace soup
