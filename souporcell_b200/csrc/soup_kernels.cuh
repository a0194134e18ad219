// Device kernels of libsoupgpu (sm_100a).  Reference: /root/reference/souporcell/src/main.rs ("S:").
//
// Layout in HBM.  A *column* is one (slot, cluster) pair, col = slot*K + k; a
// *slot* is one resident solve (one em()/khm() call, S:189/S:280).  Columns are padded to a
// multiple of the group width GW = 32*V (V = columns per lane: 1, 2, 4, 8 or 16).
//   LA | LB    : [2L][Cpad] f32  ONE joint table: rows [0, L) ln(theta), rows [L, 2L) ln(1-theta)   (row = used locus)
//   TH         : [L][Cpad] f32   theta
//   ELL, W     : [N][Cpad] f32   per-cell log-likelihoods, M-step weights (row = cell)
//   LSE        : [slots][N] f32  per-cell log_sum_exp (S:230)
//   csr_idx    : [nnz] u32 entry words, rows = cells, loci ascending;  csr_aux [nnz] float4 {alt, ref, ln_binomial, 0}
//   csc_idx    : [nnz] u32 entry words, cols = loci, cells ascending;  csc_aux [nnz] float2 {alt, alt+ref}
//   (entry words: see "entry words" below; the side records are only read for counts that do not fit the packed fields)
// One owner per (row, column) chain — a warp lane in estep_kernel / mstep_kernel, a thread in mstep_long_kernel — walks the
// row's entries strictly in order, so every (cell,k) / (locus,k) sum is accumulated in the reference's own order (S:415-418,
// S:476-483) and the results are bit-identical to the scalar code; columns never exchange data.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "glibc_flt32.h"
#include "soup_common.hpp"

namespace soup {

// entries (row gathers) in flight per warp in the E walk, by columns per lane: a narrow body runs when few solves are
// still live, and then the walk is a latency chain per warp, so narrow bodies go deeper
#ifndef SOUP_E_U1
#define SOUP_E_U1 8
#endif
#ifndef SOUP_E_U2
#define SOUP_E_U2 8
#endif
#ifndef SOUP_E_U4
#define SOUP_E_U4 8
#endif
#ifndef SOUP_E_U8
#define SOUP_E_U8 3      // (with SOUP_E_MINB 4: 64 registers, 32 warps per SM.  Measured, profiles/r2g_variants.txt / r2h_variants.txt: a ring of 3 rows at
#endif                   //  4 CTAs per SM beats 4 rows at 3 CTAs by 7 %; 2 rows at 5 CTAs, 8 rows at 2 CTAs and the fully unrolled chunk are all slower)
#ifndef SOUP_E_U16
#define SOUP_E_U16 2
#endif
#ifndef SOUP_M_U8
#define SOUP_M_U8 3      // (with SOUP_M_MINB 4, like the E walk: -6 % against 4 rows at 3 CTAs per SM)
#endif
#ifndef SOUP_M_U16
#define SOUP_M_U16 2
#endif
#ifndef SOUP_M_VLONG
#define SOUP_M_VLONG 2   // columns per lane for the long loci of the M pass
#endif
#ifndef SOUP_M_ULONG
#define SOUP_M_ULONG 16  // row gathers in flight per warp in the long role
#endif
#ifndef SOUP_M_LONG_SHORTCUT
#define SOUP_M_LONG_SHORTCUT 0   // 1: the long role also takes the single-UMI shortcuts (sign test + predicated adds)
#endif
#ifndef SOUP_DEEP_VMAX
#define SOUP_DEEP_VMAX 4  // bodies of at most this many columns per lane run the cross-chunk ring of 32 / V rows (few live solves: latency chains)
#endif
#ifndef SOUP_E_DEEP_U8
#define SOUP_E_DEEP_U8 4  // ring slots of the unrolled E walk at 8 columns per lane (when SOUP_DEEP_VMAX >= 8)
#endif
#ifndef SOUP_E_VMAX
#define SOUP_E_VMAX 8    // widest E-pass column group, in columns per lane
#endif
#ifndef SOUP_M_VWIDE
#define SOUP_M_VWIDE 8   // columns per lane for all other loci
#endif
#ifndef SOUP_E_MINB
#define SOUP_E_MINB 4    // min resident CTAs per SM the E kernel is compiled for
#endif
#ifndef SOUP_M_U
#define SOUP_M_U 8
#endif
#ifndef SOUP_M_U2
#define SOUP_M_U2 8   // 1 / 2 columns per lane (few live solves: a latency chain per warp)
#endif
#ifndef SOUP_M_MINB
#define SOUP_M_MINB 4
#endif
#ifndef SOUP_EM_THREADS
#define SOUP_EM_THREADS 256
#endif
constexpr int EM_THREADS = SOUP_EM_THREADS;   // threads per CTA in the E / M kernels (one warp per row)
constexpr int EM_WARPS = EM_THREADS / 32;
constexpr int TEMP_STEPS = 9;              // S:215

struct SlotState {
    int solve;        // global solve index (stream * solves_per_stream + iteration), -1 = empty
    int local;        // index into this call's solve queue (trace / result rows)
    int active;       // takes part in this iteration's E / post / M
    int temp_step;    // S:217
    int iterations;   // S:204
    float last_loss;  // S:216
    float loss;       // total_log_loss S:245
    int fin;          // finished in this iteration (theta is final after this iteration's M pass)
    int pending;      // next queue position to load into the slot, -1 = none
    int pad[3];
};

struct Ctl {
    int queue_head, queue_len;
    int done_count;
    int copy_slot;      // slot whose log-probs / theta the harvest kernel copies this iteration
    int best_solve;
    int has_best;
    float best_loss;
    int nonfinite;      // a ln-table holds a non-finite value: exact zero-count shortcuts are off
    int n_active;
    int n_moves;        // slot migrations decided by finalize_kernel for migrate_kernel (tail compaction)
    uint32_t live_cols; // columns [0, live_cols) hold every active slot: (highest active slot + 1) * K
    unsigned long long slot_iterations;
};

struct IterRec { int temp_step; float loss; float change; int above; };

// ------------------------------------------------------------------ small helpers
// V = columns per lane.  V <= 4: one vector load per row, lane l owns columns [l*V, l*V+V) of its group.
// V = 4*H (8, 12, 16): the group is H adjacent 128-column sub-blocks and lane l owns columns
// [h*128 + 4*l, +4) of every sub-block, so each of the H loads is one fully coalesced 512-byte row piece.
//
// Live columns.  The resident solves occupy a PREFIX of the columns (Ctl::live_cols, kept by finalize_kernel's
// compaction); a group picks the narrowest body that covers its live columns at run time, and inside the body
//   * a lane whose first column is dead re-reads the group's first column (the same sectors lane 0 fetches: no
//     extra traffic, no predicate in the loop) and stores nothing,
//   * sub-blocks h >= 1 are loaded under a loop-invariant per-lane predicate (bit h of `hm`).
// Dead lanes compute on whatever their registers hold; lanes never exchange data, so this cannot reach a result.
template <int V> struct FVec { float v[V]; };
template <int V> __device__ __forceinline__ constexpr uint32_t lane_cols() { return V < 4 ? V : 4; }
template <int V> __device__ __forceinline__ void ld_row(FVec<V>& r, const float* p, uint32_t hm) {
    static_assert(V % 4 == 0, "wide rows are multiples of 4 columns per lane");
#pragma unroll
    for (int h = 0; h < V / 4; ++h) {
        if (h == 0 || ((hm >> h) & 1u)) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p + h * 128));
            r.v[4 * h] = t.x; r.v[4 * h + 1] = t.y; r.v[4 * h + 2] = t.z; r.v[4 * h + 3] = t.w;
        }
    }
}
template <> __device__ __forceinline__ void ld_row<1>(FVec<1>& r, const float* p, uint32_t) { r.v[0] = __ldg(p); }
template <> __device__ __forceinline__ void ld_row<2>(FVec<2>& r, const float* p, uint32_t) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p)); r.v[0] = t.x; r.v[1] = t.y;
}
template <int V> __device__ __forceinline__ void st_row(float* p, const FVec<V>& x, uint32_t hm) {
    static_assert(V % 4 == 0, "wide rows are multiples of 4 columns per lane");
#pragma unroll
    for (int h = 0; h < V / 4; ++h)
        if (h == 0 || ((hm >> h) & 1u))
            *reinterpret_cast<float4*>(p + h * 128) = make_float4(x.v[4 * h], x.v[4 * h + 1], x.v[4 * h + 2], x.v[4 * h + 3]);
}
template <> __device__ __forceinline__ void st_row<1>(float* p, const FVec<1>& x, uint32_t) { *p = x.v[0]; }
template <> __device__ __forceinline__ void st_row<2>(float* p, const FVec<2>& x, uint32_t) { *reinterpret_cast<float2*>(p) = make_float2(x.v[0], x.v[1]); }
// global column of slot v of lane `lane` in a group starting at col0
template <int V> __device__ __forceinline__ uint32_t col_of(uint32_t col0, uint32_t lane, int v) {
    return V < 4 ? col0 + lane * V + v : col0 + (v >> 2) * 128 + lane * 4 + (v & 3);
}
template <int V> struct LaneCols {
    uint32_t col;    // first column this lane reads (redirected to the group's first column when the lane is dead)
    uint32_t hm;     // bit h (h >= 1): sub-block h of this lane holds live columns
    bool live;       // the lane's first sub-block holds live columns
};
template <int V> __device__ __forceinline__ LaneCols<V> lane_setup(uint32_t col0, uint32_t live_cols) {
    const uint32_t lane = threadIdx.x & 31;
    LaneCols<V> lc;
    lc.col = col0 + lane * lane_cols<V>();
    lc.live = lc.col < live_cols;
    lc.hm = 0;
#pragma unroll
    for (int h = 1; h < V / 4; ++h)
        if (lc.col + h * 128 < live_cols) lc.hm |= 1u << h;
    if (!lc.live) lc.col = col0;
    return lc;
}

// threadIdx.x >> 5 in a form ptxas KNOWS to be warp-uniform (REDUX writes a uniform register).  Everything derived from it — the
// row a warp owns, its entry count, the words it loads from uniform addresses — then stays uniform in ptxas' divergence analysis, and
// the walks' loops and class branches compile without BSSY / BSYNC reconvergence pairs and without the BRA.DIV guard in front of every SHFL.
__device__ __forceinline__ uint32_t warp_id_uniform() { return __reduce_max_sync(0xffffffffu, threadIdx.x >> 5); }

// base + row * row_bytes in ONE instruction (IMAD.WIDE.U32); the compiler otherwise widens the
// index arithmetic to 64 bits and spends four instructions per gathered row.
__device__ __forceinline__ const float* row_at(const float* base, uint32_t row, uint32_t row_bytes) {
    const float* r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(row), "r"(row_bytes), "l"(base));
    return r;
}

// Packed f32x2 addition (sm_100: FADD2): each half is the same IEEE round-to-nearest operation as the
// scalar instruction (checked bit for bit on random patterns, profiles/probes/f32x2_probe.cu); the FP32 lanes run
// at the same rate, but one instruction carries two columns, which frees issue slots for the address / decode work.
#ifndef SOUP_PACKED_FP
#define SOUP_PACKED_FP 1
#endif
__device__ __forceinline__ void add2(float& x0, float& x1, float a0, float a1) {
#if SOUP_PACKED_FP
    uint64_t x, a;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(a));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
#else
    x0 = __fadd_rn(x0, a0); x1 = __fadd_rn(x1, a1);
#endif
}
// acc = fma(p, c, acc) on two columns (FFMA2).  Only used with c == 0 or a power of two, where p*c is exact and the fused
// operation therefore rounds exactly like the reference's separate multiply and add.
__device__ __forceinline__ void fma2(float& x0, float& x1, float p0, float p1, float c) {
#if SOUP_PACKED_FP
    uint64_t x, p, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(p0), "f"(p1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c), "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x) : "l"(p), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
#else
    x0 = __fmaf_rn(p0, c, x0); x1 = __fmaf_rn(p1, c, x1);
#endif
}
// acc += t
template <int V> __device__ __forceinline__ void vadd(FVec<V>& acc, const FVec<V>& t) {
    if constexpr (V == 1) acc.v[0] = __fadd_rn(acc.v[0], t.v[0]);
    else {
#pragma unroll
        for (int v = 0; v + 1 < V; v += 2) add2(acc.v[v], acc.v[v + 1], t.v[v], t.v[v + 1]);
    }
}
// r = t * c and r = t + c (c the same for every column): scalar on purpose — ptxas contracts a packed multiply that
// feeds a packed add into FFMA2 even with explicit .rn and -fmad=false, which would change the rounding
template <int V> __device__ __forceinline__ FVec<V> vscale(const FVec<V>& t, float c) {
    FVec<V> r;
#pragma unroll
    for (int v = 0; v < V; ++v) r.v[v] = __fmul_rn(t.v[v], c);
    return r;
}
template <int V> __device__ __forceinline__ FVec<V> vshift(const FVec<V>& t, float c) {
    FVec<V> r = t;
    if constexpr (V == 1) r.v[0] = __fadd_rn(c, r.v[0]);
    else {
#pragma unroll
        for (int v = 0; v + 1 < V; v += 2) add2(r.v[v], r.v[v + 1], c, c);
    }
    return r;
}

// r = t + t  (== 2*t exactly)
template <int V> __device__ __forceinline__ FVec<V> vdouble(const FVec<V>& t) {
    FVec<V> r = t;
    vadd<V>(r, t);
    return r;
}
// r = c * t for a small count c: 1 and 2 need no multiply (1*t == t, 2*t == t+t, both exact)
template <int V> __device__ __forceinline__ FVec<V> vcount(const FVec<V>& t, uint32_t ci, float c) {
    if (ci == 1) return t;
    if (ci == 2) return vdouble<V>(t);
    return vscale<V>(t, c);
}

// ------------------------------------------------------------------ entry words
// ONE 32-bit word per stored entry.  70 % of vartrix entries are single-UMI (alt + ref == 1), so those get the
// cheapest possible form — bit 31 clear and the word is directly usable as an index:
//   CSR stream (E pass): bits 0..29 = row of the joint ln table [ln th rows | ln(1-th) rows] this entry adds,
//                        i.e. locus (alt only) or plane_rows + locus (ref only); bit 30 = the row counts twice
//                        (two UMIs of one allele, another 15 % of the entries: 2*x == x + x exactly and lbc == +0);
//   CSC stream (M pass): word = cell | 1 << sh_r | alt << sh_a (the general layout with bit 31 clear).
// Every other entry has bit 31 set:  1 | c0 << sh_a | c1 << sh_r | index  with (c0, c1) = (alt, ref) in the CSR
// stream and (alt, alt+ref) in the CSC stream; counts that do not fit their field set both fields to the all-ones
// marker and live in a side record (float4 {alt, ref, lbc, 0} / float2 {alt, alt+ref}).
struct Packing {
    uint32_t sh_a, sh_r;   // shifts of the two count fields of a general word
    uint32_t cmask;        // (1 << count_bits) - 1: field mask and overflow marker
    uint32_t imask;        // index mask (general CSR words, and every CSC word)
    uint32_t cbits;
    uint32_t plane_rows;   // CSR stream: rows of one ln-table plane (max(L, 1))
    uint32_t alt_bit;      // CSC stream: 1 << sh_a, the alt bit of a simple word
};
__device__ __forceinline__ bool word_simple(uint32_t w) { return (int32_t)w >= 0; }
constexpr uint32_t CSR_DOUBLE = 0x40000000u, CSR_ROW_MASK = 0x3fffffffu;

__constant__ float c_lbc[4096];   // ln_binomial(alt + ref, alt) as f32 for in-field counts, index alt << 6 | ref (same for every context)

// ------------------------------------------------------------------ E pass  (S:409-426)
// acc_k = ln(1/K); for each entry j (ascending locus): acc_k += (lbc_j + alt_j*ln th) + ref_j*ln(1-th).
// Exact shortcuts, all bit-identical to the literal expression:
//   alt 1, ref 0 : acc += ln th          (1*x == x, lbc == +0, 0 + x == x, and "+ 0*ln(1-th)" adds -0.0)
//   alt 0, ref 1 : acc += ln(1-th)         -> both are "simple" words: one row gather, V adds
//   ref 0        : acc += lbc + alt*ln th
//   alt 0        : acc += lbc + ref*ln(1-th)
// Dropping a zero-count product is valid only while every table value is finite (0*inf = NaN), so
// Ctl::nonfinite switches every entry to the literal expression.
struct EDec { uint32_t locus, ai, ri; };
__device__ __forceinline__ EDec e_decode(uint32_t w, const Packing& pk) {
    EDec d;
    if (word_simple(w)) {
        const uint32_t row = w & pk.imask, n = (w & CSR_DOUBLE) ? 2u : 1u;
        const bool isref = row >= pk.plane_rows;
        d.locus = isref ? row - pk.plane_rows : row;
        d.ai = isref ? 0u : n;
        d.ri = isref ? n : 0u;
    } else {
        const uint32_t row = w & pk.imask;
        d.locus = row >= pk.plane_rows ? row - pk.plane_rows : row;
        d.ai = (w >> pk.sh_a) & pk.cmask;
        d.ri = (w >> pk.sh_r) & pk.cmask;
    }
    return d;
}

template <int V>
struct ETile { FVec<V> a, b; };

// ---- literal walk (Ctl::nonfinite: a ln table holds +-inf, so no zero-count product may be dropped): both rows of every
// entry, the reference's expression as written.  Rare; kept simple.
template <int V>
__device__ __forceinline__ void e_literal_entry(FVec<V>& acc, uint32_t w, const Packing& pk, const float4* __restrict__ aux,
                                                const float* __restrict__ la, const float* __restrict__ lb, uint32_t row_bytes, uint32_t hm) {
    const EDec d = e_decode(w, pk);
    ETile<V> t;
    ld_row<V>(t.a, row_at(la, d.locus, row_bytes), hm);
    ld_row<V>(t.b, row_at(lb, d.locus, row_bytes), hm);
    float alt, ref, lbc;
    if (d.ai == pk.cmask && d.ri == pk.cmask && !word_simple(w)) {          // counts outside the packed fields (rare)
        const float4 x = __ldg(aux);
        alt = x.x; ref = x.y; lbc = x.z;
    } else {
        alt = (float)d.ai; ref = (float)d.ri; lbc = c_lbc[(d.ai << 6) | d.ri];
    }
    FVec<V> x = vshift<V>(vscale<V>(t.a, alt), lbc);
    vadd<V>(x, vscale<V>(t.b, ref));
    vadd<V>(acc, x);
}
template <int V>
__device__ __forceinline__ void e_walk_literal(FVec<V>& acc, const uint32_t* __restrict__ p, const float4* __restrict__ aux, uint32_t n,
                                               const Packing& pk, const float* __restrict__ la, const float* __restrict__ lb, uint32_t row_bytes, uint32_t hm) {
#pragma unroll 1
    for (uint32_t j = 0; j < n; ++j) e_literal_entry<V>(acc, __ldg(p + j), pk, aux + j, la, lb, row_bytes, hm);
}

// ---- the walk (every table value finite).  A warp owns (cell, column group).  It takes the cell's entry words 32 at a time with
// ONE coalesced load (lane j holds word j), classifies them once per chunk with warp votes — whose results ptxas knows to be
// warp-uniform, so the per-entry class tests below are uniform branches without the BSSY / BSYNC reconvergence pairs that
// tests on a loaded word cost — and then serves the entries strictly in order: the owner lane's pre-masked row index and
// multiplier are broadcast with SHFL, every lane gathers its V columns of that ONE row, and
//   * count 1 / 2 / 4 ... of one allele (85 % of the entries are 1 or 2):   acc = fma(t, c, acc), c a power of two.  t*c is then
//     exact, so the fused operation rounds exactly like the reference's  acc + c*t  (and lbc == ln C(n,n) == +0, 0 + x == x);
//   * any other count of one allele:  acc += rn(c*t);
//   * both alleles (or counts beyond the packed fields):  the second row is fetched late and the literal expression runs.
// The general classes read {alt, ref, lbc} from the side record of the entry (owner lane, once per chunk).
template <int V> __device__ __forceinline__ void vfma(FVec<V>& acc, const FVec<V>& t, float c) {
    if constexpr (V == 1) acc.v[0] = __fmaf_rn(t.v[0], c, acc.v[0]);
    else {
#pragma unroll
        for (int v = 0; v + 1 < V; v += 2) fma2(acc.v[v], acc.v[v + 1], t.v[v], t.v[v + 1], c);
    }
}
__device__ __forceinline__ bool pow2_or_zero(float c) { return (__float_as_uint(c) & 0x007fffffu) == 0u; }   // (finite by construction)

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

template <int V>
__device__ __forceinline__ void e_general(FVec<V>& acc, const FVec<V>& t, bool both, uint32_t src, uint32_t row, const float4& x, float cf,
                                          const Packing& pk, const float* __restrict__ la, uint32_t row_bytes, uint32_t hm) {
    if (both) {
        FVec<V> tb;
        ld_row<V>(tb, row_at(la, __shfl_sync(0xffffffffu, row, src) + pk.plane_rows, row_bytes), hm);
        const float alt = __shfl_sync(0xffffffffu, x.x, src), ref = __shfl_sync(0xffffffffu, x.y, src), lbc = __shfl_sync(0xffffffffu, x.z, src);
        FVec<V> y = vshift<V>(vscale<V>(t, alt), lbc);
        vadd<V>(y, vscale<V>(tb, ref));
        vadd<V>(acc, y);
    } else {
        vadd<V>(acc, vscale<V>(t, __shfl_sync(0xffffffffu, cf, src)));
    }
}

// One chunk of 32 entry words, classified (lane j holds entry j; gm / bm are warp votes)
struct EChunk { uint32_t row; float cf; float4 x; uint32_t gm, bm; };
__device__ __forceinline__ EChunk e_chunk(const uint32_t* __restrict__ p, const float4* __restrict__ aux, uint32_t base, uint32_t n, const Packing& pk) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t w = base + lane < n ? __ldg(p + base + lane) : 0u;      // (beyond the end: row 0, a valid unused gather)
    EChunk c;
    c.row = w & pk.imask;
    c.cf = (w & CSR_DOUBLE) ? 2.0f : 1.0f;
    c.x = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    bool gen = false, both = false;
    if (!word_simple(w)) {
        c.x = __ldg(aux + base + lane);
        const bool marker = ((w >> pk.sh_a) & pk.cmask) == pk.cmask && ((w >> pk.sh_r) & pk.cmask) == pk.cmask;
        both = marker || (c.x.x != 0.0f && c.x.y != 0.0f);
        c.cf = c.x.x != 0.0f ? c.x.x : c.x.y;
        gen = both || !pow2_or_zero(c.cf);
    }
    c.gm = __ballot_sync(0xffffffffu, gen);
    c.bm = __ballot_sync(0xffffffffu, both);
    return c;
}
// entry `e` (warp-uniform) of chunk c; t is its gathered row
template <int V>
__device__ __forceinline__ void e_consume(FVec<V>& acc, const FVec<V>& t, uint32_t e, const EChunk& c,
                                          const Packing& pk, const float* __restrict__ la, uint32_t row_bytes, uint32_t hm) {
    if (!((c.gm >> e) & 1u)) vfma<V>(acc, t, __shfl_sync(0xffffffffu, c.cf, e));
    else e_general<V>(acc, t, (c.bm >> e) & 1u, e, c.row, c.x, c.cf, pk, la, row_bytes, hm);
}
template <int V>
__device__ __forceinline__ void e_fetch(FVec<V>& t, uint32_t e, const EChunk& c, const Packing& pk, const float* __restrict__ la, uint32_t row_bytes, uint32_t hm, bool pf) {
    const uint32_t r = __shfl_sync(0xffffffffu, c.row, e);
    ld_row<V>(t, row_at(la, r, row_bytes), hm);
    if (pf && ((c.bm >> e) & 1u)) prefetch_l1(row_at(la, r + pk.plane_rows, row_bytes));   // the second row of a two-allele entry: into L1 ahead of its late load
}

// The rows are gathered through a ring of U row registers: entry e lives in slot e % U, and as soon as it has been consumed the slot
// takes the row of entry e + U.  So U - 1 .. U row gathers (U KB at 256 columns) stay in flight per warp the whole time, instead of a
// batch that is issued, awaited and consumed before the next one starts (the pass is bound by the latency of these gathers: ncu shows
// long-scoreboard stalls dominating with the L1 data pipe at 63 %, L2 at 52 % and issue at 45 %).
// Wide bodies (V = 8, 16; the full-occupancy passes, bound by throughput) refill the ring inside a chunk and prefetch the next
// chunk's words.  Narrow bodies (V <= 4: few live solves, the tail of a job, where a pass lasts as long as its longest cell) take
// U = 32 / V slots and refill ACROSS the chunk boundary from the already classified next chunk, so the chain never drains.
template <int V, int U>
__device__ __forceinline__ void e_walk(FVec<V>& acc, const uint32_t* __restrict__ p, const float4* __restrict__ aux, uint32_t n,
                                       const Packing& pk, const float* __restrict__ la, uint32_t row_bytes, uint32_t hm) {
#pragma unroll 1
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t cnt = min(32u, n - base);
        const EChunk c = e_chunk(p, aux, base, n, pk);
        FVec<V> t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) e_fetch<V>(t[u], u, c, pk, la, row_bytes, hm, false);
        uint32_t e = 0;
#pragma unroll 1
        for (; e + 2 * U <= cnt; e += U) {                                // steady state: consume entry e + u, refill its slot with e + U + u
#pragma unroll
            for (int u = 0; u < U; ++u) {
                e_consume<V>(acc, t[u], e + u, c, pk, la, row_bytes, hm);
                e_fetch<V>(t[u], e + U + u, c, pk, la, row_bytes, hm, false);
            }
        }
        // fewer than 2U entries left in the chunk: the ring holds entries [e, e + U); refills only where an entry exists
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e + u < cnt) {
                e_consume<V>(acc, t[u], e + u, c, pk, la, row_bytes, hm);
                if (e + U + u < cnt) e_fetch<V>(t[u], e + U + u, c, pk, la, row_bytes, hm, false);
            }
        }
        e += U;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (e + u < cnt) e_consume<V>(acc, t[u], e + u, c, pk, la, row_bytes, hm);
    }
}

template <int V, int U, bool GUARD>
__device__ __forceinline__ void e_deep_chunk(FVec<V>& acc, FVec<V> (&t)[U], uint32_t cnt, uint32_t left_after, const EChunk& c, const EChunk& c2,
                                             const Packing& pk, const float* __restrict__ la, uint32_t row_bytes, uint32_t hm) {
    static_assert(32 % U == 0, "the ring must divide a chunk");
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        FVec<V>& slot = t[i % U];
        if (!GUARD || (uint32_t)i < cnt) e_consume<V>(acc, slot, i, c, pk, la, row_bytes, hm);
        const int j = i + U;                    // the stream entry that takes the slot
        if (j < 32) {
            if (!GUARD || (uint32_t)j < cnt) e_fetch<V>(slot, j, c, pk, la, row_bytes, hm, true);
        } else {
            if (!GUARD || (uint32_t)(j - 32) < left_after) e_fetch<V>(slot, j - 32, c2, pk, la, row_bytes, hm, true);
        }
    }
}
template <int V, int U>
__device__ __forceinline__ void e_walk_deep(FVec<V>& acc, const uint32_t* __restrict__ p, const float4* __restrict__ aux, uint32_t n,
                                            const Packing& pk, const float* __restrict__ la, uint32_t row_bytes, uint32_t hm) {
    EChunk c = e_chunk(p, aux, 0, n, pk);
    FVec<V> t[U];
#pragma unroll
    for (int u = 0; u < U; ++u) e_fetch<V>(t[u], u, c, pk, la, row_bytes, hm, true);
#pragma unroll 1
    for (uint32_t base = 0; base < n; base += 32) {
        const EChunk c2 = e_chunk(p, aux, base + 32, n, pk);
        const uint32_t left = n - base;
        if (left >= 64) e_deep_chunk<V, U, false>(acc, t, 32u, 32u, c, c2, pk, la, row_bytes, hm);
        else e_deep_chunk<V, U, true>(acc, t, min(left, 32u), left > 32u ? left - 32u : 0u, c, c2, pk, la, row_bytes, hm);
        c = c2;
    }
}

struct EArgs {
    const uint32_t* csr_idx; const float4* csr_aux; const uint64_t* row_ptr; const uint32_t* cell_order;
    const float* LA; const float* LB; float* ELL; const int* group_active; const Ctl* ctl;
    Packing pk;
    uint32_t N, Cpad, blocks_per_group;
    uint32_t no_narrow;    // tests: every group runs the full-width body whatever its live columns
    float log_prior;
};

template <int V>
__device__ __forceinline__ void estep_body(const EArgs& a, uint32_t cell, uint32_t col0, uint32_t live_cols) {
    const LaneCols<V> lc = lane_setup<V>(col0, live_cols);
    const uint64_t j0 = a.row_ptr[cell];
    const uint32_t n = (uint32_t)(a.row_ptr[cell + 1] - j0);
    FVec<V> acc;
#pragma unroll
    for (int v = 0; v < V; ++v) acc.v[v] = a.log_prior;
    constexpr int U = V == 1 ? SOUP_E_U1 : (V == 2 ? SOUP_E_U2 : (V == 4 ? SOUP_E_U4 : (V == 8 ? SOUP_E_U8 : SOUP_E_U16)));
    if (!a.ctl->nonfinite) {
        if constexpr (V <= SOUP_DEEP_VMAX) e_walk_deep<V, (V == 8 ? SOUP_E_DEEP_U8 : 32 / V)>(acc, a.csr_idx + j0, a.csr_aux + j0, n, a.pk, a.LA + lc.col, a.Cpad * 4u, lc.hm);
        else e_walk<V, U>(acc, a.csr_idx + j0, a.csr_aux + j0, n, a.pk, a.LA + lc.col, a.Cpad * 4u, lc.hm);
    }
    else e_walk_literal<V>(acc, a.csr_idx + j0, a.csr_aux + j0, n, a.pk, a.LA + lc.col, a.LB + lc.col, a.Cpad * 4u, lc.hm);
    if (lc.live) st_row<V>(a.ELL + (size_t)cell * a.Cpad + lc.col, acc, lc.hm);
}

// the narrowest body (columns per lane) whose 32*V columns cover `rem` live columns, at most VM
template <int VM> __device__ __forceinline__ int fit_lanes(uint32_t rem, uint32_t no_narrow) {
    if (no_narrow) return VM;
    if (VM >= 16 && rem > 256) return 16;
    if (VM >= 8 && rem > 128) return 8;
    if (VM >= 4 && rem > 64) return 4;
    if (VM >= 2 && rem > 32) return 2;
    return 1;
}

// grid: group-major (all cells of group 0, then group 1, ...) so one group's table slice stays hot in L2;
// inside a group the cells are visited longest-first.  Groups are 32*VM columns wide; a group beyond the live
// prefix exits, a partly live one runs the narrowest body that covers what is live (the tail of a job, when
// a handful of solves are still iterating, then costs a fraction of a full pass).
template <int VM>
__global__ void __launch_bounds__(EM_THREADS, (VM <= 8 ? SOUP_E_MINB : 2) * (256 / EM_THREADS)) estep_kernel(const EArgs a) {
    const uint32_t group = blockIdx.x / a.blocks_per_group;
    const uint32_t live = a.ctl->live_cols, col0 = group * (32 * VM);
    if (col0 >= live || !a.group_active[group]) return;
    const uint32_t chunk = blockIdx.x - group * a.blocks_per_group;
    const uint32_t ci = chunk * EM_WARPS + (EM_WARPS > 1 ? warp_id_uniform() : 0);
    if (ci >= a.N) return;
    const uint32_t cell = a.cell_order[ci];
    const int v = fit_lanes<VM>(live - col0, a.no_narrow);
    if (VM >= 16 && v == 16) estep_body<(VM >= 16 ? 16 : VM)>(a, cell, col0, live);
    else if (VM >= 8 && v == 8) estep_body<(VM >= 8 ? 8 : VM)>(a, cell, col0, live);
    else if (VM >= 4 && v == 4) estep_body<(VM >= 4 ? 4 : VM)>(a, cell, col0, live);
    else if (VM >= 2 && v == 2) estep_body<(VM >= 2 ? 2 : VM)>(a, cell, col0, live);
    else estep_body<1>(a, cell, col0, live);
}

// ------------------------------------------------------------------ per-cell pass (S:230-236 / S:318-330)
struct PostArgs {
    const float* ELL;
    float* W;
    float* LSE;
    const float* total_alleles;
    const SlotState* st;
    const Ctl* ctl;
    int* counts;          // [slots][K] cells per cluster (LAST max, S:254), or nullptr
    uint32_t N, Cpad;
    int K, slots, method; // K = every cluster of the model (the loops below run over all of them, in index order)
    // Locked clusters of the souporcell3 re-runs (S:92-97).  Their centers never change during the run, so
    //  * their log-likelihood columns are computed ONCE per run (ELLfix [N][Kf]) and the engine's tables / E / M passes hold only the
    //    Kx re-drawn clusters per slot: cluster k reads ELL[cell][slot*Kx + src_of_k[k]] if src_of_k[k] >= 0, else ELLfix[cell][-src_of_k[k]-1];
    //  * they take no part in the M pass: the weight of cluster k goes to W[cell][slot*nd + w_of_k[k]], or nowhere if w_of_k[k] < 0.
    // nullptr = identity (every cluster has its own engine column / weight column).
    int Kx, nd;
    const int* src_of_k;
    const int* w_of_k;
    const float* ELLfix;
    int Kf;
};

// thread = (cell, slot), slot fastest: a warp covers `sw` neighbouring slots of 32 / sw cells, so its ELL / W accesses —
// nearly all of this kernel's traffic — are contiguous runs of sw*K*4 bytes (the LSE store, 4 bytes per thread, is the
// strided one).  sw = 32 while many solves are live, the next power of two above the live slots in the tail of a job.
constexpr int POST_WARPS = 4;

template <typename Fetch, typename Store>
__device__ __forceinline__ void post_body(const PostArgs& a, uint32_t cell, int slot, const SlotState& s, Fetch X, Store put) {
    const int K = a.K;
    {   // log_sum_exp S:428-432
        float m = -INFINITY;
        for (int k = 0; k < K; ++k) m = fmaxf(m, X(k));
        float sum = 0.0f;
        for (int k = 0; k < K; ++k) sum = sum + soupmath::glibc_expf(X(k) - m);
        a.LSE[(size_t)slot * a.N + cell] = m + soupmath::glibc_logf(sum);
    }
    const float total = a.total_alleles[cell];
    float temp;
    if (a.method == SOUP_METHOD_EM) temp = fmaxf(total / ldexpf(20.0f, s.temp_step), 1.0f);   // S:233
    else temp = fmaxf(total / ldexpf(0.5f, s.temp_step), 1.0f);                               // S:327
    if (s.temp_step == TEMP_STEPS - 1) temp = 1.0f;                                            // S:234 / S:328
    if (a.method == SOUP_METHOD_EM) {
        // normalize_in_log_with_temp S:443-454 on the log-likelihoods themselves
        float m = -INFINITY;
        for (int k = 0; k < K; ++k) m = fmaxf(m, X(k) / temp);
        float sum = 0.0f;
        for (int k = 0; k < K; ++k) sum = sum + soupmath::glibc_expf(X(k) / temp - m);
        const float lse = m + soupmath::glibc_logf(sum);
        for (int k = 0; k < K; ++k) put(k, soupmath::glibc_expf(X(k) / temp - lse));
    } else {
        // calculate_q_for_current_cell S:360-384 in the reference's literal operation order
        int win = 0;
        float best = -3.40282347e+38f;   // f32::MIN S:411
        for (int k = 0; k < K; ++k) { const float xk = X(k); if (xk > best) { best = xk; win = k; } }
        const float wl = -X(win);
        float m = -INFINITY;
        for (int k = 0; k < K; ++k) m = fmaxf(m, k != win ? 25.0f * (wl + X(k)) : 0.0f);
        float sum = 0.0f;
        for (int k = 0; k < K; ++k) sum = sum + soupmath::glibc_expf((k != win ? 25.0f * (wl + X(k)) : 0.0f) - m);
        const float q_denom = m + soupmath::glibc_logf(sum);
        const float c0 = 50.0f * wl, c2 = 2.0f * q_denom;
#define SOUP_Q(k) ((c0 - (27.0f * -X(k))) - c2)
        m = -INFINITY;
        for (int k = 0; k < K; ++k) m = fmaxf(m, SOUP_Q(k));
        sum = 0.0f;
        for (int k = 0; k < K; ++k) sum = sum + soupmath::glibc_expf(SOUP_Q(k) - m);
        const float q_sum = m + soupmath::glibc_logf(sum);
        // khm_prob = q - q_sum (S:322-325), then normalize_in_log_with_temp (S:330)
        m = -INFINITY;
        for (int k = 0; k < K; ++k) m = fmaxf(m, (SOUP_Q(k) - q_sum) / temp);
        sum = 0.0f;
        for (int k = 0; k < K; ++k) sum = sum + soupmath::glibc_expf((SOUP_Q(k) - q_sum) / temp - m);
        const float lse = m + soupmath::glibc_logf(sum);
        for (int k = 0; k < K; ++k) put(k, soupmath::glibc_expf((SOUP_Q(k) - q_sum) / temp - lse));
#undef SOUP_Q
    }
    if (a.counts) {   // max_by(total_cmp) keeps the LAST maximum (S:254)
        int bi = 0;
        int32_t bk = total_key(X(0));
        for (int k = 1; k < K; ++k) {
            int32_t kk = total_key(X(k));
            if (kk >= bk) { bk = kk; bi = k; }
        }
        atomicAdd(a.counts + (size_t)slot * K + bi, 1);
    }
}

__global__ void __launch_bounds__(32 * POST_WARPS) post_kernel(PostArgs a) {
    const uint32_t live_slots = min((a.ctl->live_cols + (uint32_t)a.Kx - 1) / (uint32_t)a.Kx, (uint32_t)a.slots);
    uint32_t sw = 32;
    while (sw > 1 && (sw >> 1) >= live_slots) sw >>= 1;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cw = 32 / sw;
    const uint32_t cell = (blockIdx.x * POST_WARPS + warp) * cw + lane / sw;
    const int slot = (int)(blockIdx.y * 32 + (lane & (sw - 1)));
    if (cell >= a.N || slot >= a.slots || (uint32_t)slot >= live_slots) return;
    const SlotState s = a.st[slot];
    if (!s.active) return;
    const float* x = a.ELL + (size_t)cell * a.Cpad + (size_t)slot * a.Kx;
    float* w = a.W + (size_t)cell * a.Cpad + (size_t)slot * a.nd;
    if (!a.src_of_k && !a.w_of_k) {
        post_body(a, cell, slot, s, [&](int k) { return x[k]; }, [&](int k, float v) { w[k] = v; });
    } else {
        const float* fx = a.ELLfix + (size_t)cell * a.Kf;
        post_body(a, cell, slot, s,
                  [&](int k) { const int i = a.src_of_k ? a.src_of_k[k] : k; return i >= 0 ? x[i] : fx[-i - 1]; },
                  [&](int k, float v) { const int d = a.w_of_k ? a.w_of_k[k] : k; if (d >= 0) w[d] = v; });
    }
}

// ------------------------------------------------------------------ loss + control (S:217-222, 245-250)
// log_binom_loss (S:225-230) is a strictly sequential f32 sum over cells in index order; it decides
// convergence (S:222), so it is reproduced literally: one block per slot streams LSE[slot][:] through
// shared memory (all warps fetch, double-buffered) and a single thread adds in order.
constexpr int LOSS_TILE = 2048;
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ LSE, const SlotState* __restrict__ st,
                                                   float* __restrict__ loss_out, uint32_t N) {
    const int slot = blockIdx.x;
    if (!st[slot].active) return;
    __shared__ __align__(16) float buf[2][LOSS_TILE];
    const float* p = LSE + (size_t)slot * N;
    const uint32_t tiles = (N + LOSS_TILE - 1) / LOSS_TILE;
    for (uint32_t i = threadIdx.x; i < LOSS_TILE; i += blockDim.x) buf[0][i] = i < N ? p[i] : 0.0f;
    __syncthreads();
    float acc = 0.0f;
    for (uint32_t t = 0; t < tiles; ++t) {
        const uint32_t base = (t + 1) * LOSS_TILE;
        if (t + 1 < tiles)
            for (uint32_t i = threadIdx.x; i < LOSS_TILE; i += blockDim.x) buf[(t + 1) & 1][i] = base + i < N ? p[base + i] : 0.0f;
        if (threadIdx.x == 0) {
            const uint32_t cnt = min((uint32_t)LOSS_TILE, N - t * LOSS_TILE);
            const float4* b4 = reinterpret_cast<const float4*>(buf[t & 1]);
            uint32_t i = 0;
            for (; i + 4 <= cnt; i += 4) {
                const float4 v = b4[i >> 2];
                acc = acc + v.x; acc = acc + v.y; acc = acc + v.z; acc = acc + v.w;
            }
            for (; i < cnt; ++i) acc = acc + buf[t & 1][i];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_out[slot] = acc;
}

struct ControlArgs {
    const float* loss_in;     // [slots] from loss_kernel
    SlotState* st;
    Ctl* ctl;
    int* counts;              // [slots][K] or nullptr
    IterRec* trace;           // [queue_len][iter_cap]
    float* solve_loss;        // [queue_len]
    int* solve_iters;         // [queue_len]
    uint32_t N;
    int K, slots, iter_cap;
    float limit;              // log_loss_change_limit S:214
};

// advance to the first temp step whose while-condition holds with log_loss_change = 10000 (S:219-222)
__device__ __forceinline__ int next_temp_step(int t, int iterations, int cap, float limit) {
    while (t < TEMP_STEPS && !(10000.0f > limit && iterations < cap)) ++t;
    return t;
}

__global__ void __launch_bounds__(1024) control_kernel(ControlArgs a) {
    for (int slot = threadIdx.x; slot < a.slots; slot += blockDim.x) {
        SlotState s = a.st[slot];
        if (!s.active) continue;
        const float loss = a.loss_in[slot];
        const float change = loss - s.last_loss;     // S:246
        s.loss = loss;
        s.last_loss = loss;
        s.iterations += 1;                           // S:250
        int above = -1;
        if (a.counts) {
            above = 0;
            for (int k = 0; k < a.K; ++k) {
                if (a.counts[(size_t)slot * a.K + k] > 200) ++above;   // S:260
                a.counts[(size_t)slot * a.K + k] = 0;
            }
        }
        if (s.iterations <= a.iter_cap) a.trace[(size_t)s.local * a.iter_cap + (s.iterations - 1)] = IterRec{s.temp_step, loss, change, above};
        if (!(change > a.limit && s.iterations < a.iter_cap))   // leave this temp step's while loop (S:222)
            s.temp_step = next_temp_step(s.temp_step + 1, s.iterations, a.iter_cap, a.limit);
        if (s.temp_step >= TEMP_STEPS) s.fin = 1;
        a.st[slot] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        Ctl c = *a.ctl;
        c.copy_slot = -1;
        for (int slot = 0; slot < a.slots; ++slot) {
            SlotState s = a.st[slot];
            if (!s.active) continue;
            c.slot_iterations += 1;
            if (!s.fin) continue;
            a.solve_loss[s.local] = s.loss;
            a.solve_iters[s.local] = s.iterations;
            // best-of (S:110, S:120: strict >, threads in index order) == max loss, ties -> lowest solve index
            if (s.loss > c.best_loss || (c.has_best && s.loss == c.best_loss && s.solve < c.best_solve)) {
                c.best_loss = s.loss;
                c.best_solve = s.solve;
                c.has_best = 1;
                c.copy_slot = slot;
            }
            c.done_count += 1;
            s.pending = c.queue_head < c.queue_len ? c.queue_head++ : -1;
            a.st[slot] = s;
        }
        *a.ctl = c;
    }
}

// ------------------------------------------------------------------ M pass (S:476-483, S:386-396)
// s = 1 + sum_cells p*alt ; d = 2 + sum_cells p*(alt+ref), cells ascending; theta = clamp(s/d).
// Entry stream: see "entry words" above (simple: cell | alt << 30; general: 1 | alt | alt+ref | cell).
__device__ __forceinline__ float small_u2f(uint32_t i) {   // exact (float)i for i < 2^23 without the conversion pipe
    return __fadd_rn(__uint_as_float(0x4B000000u | i), -8388608.0f);
}
// The walk mirrors the E pass: a warp owns (locus, column group), takes the locus's entry words 32 at a time with one coalesced
// load, the owner lane decodes cell / alt / alt+ref ONCE, and the entries are then served strictly in order — cell row index and
// the two multipliers broadcast with SHFL, every lane gathers its V weights of that cell's row:
//   * both counts 0 or a power of two (91 % of the entries have alt+ref <= 2):  s = fma(p, alt, s), d = fma(p, alt+ref, d).  p*c is
//     exact for such c, so the fused operation rounds exactly like the reference's  s + p*alt  — for every p, finite or not
//     (fma(NaN, 0, s) and s + NaN*0 are both NaN), which is why this pass needs no Ctl::nonfinite variant;
//   * anything else:  s += rn(p*alt); d += rn(p*(alt+ref))  as written (warp-uniform branch on a vote mask).
template <int V>
__device__ __forceinline__ void m_general(FVec<V>& s, FVec<V>& d, const FVec<V>& p, float af, float tf) {
    vadd<V>(s, vscale<V>(p, af));
    vadd<V>(d, vscale<V>(p, tf));
}
template <int V>
__device__ __forceinline__ void m_consume(FVec<V>& s, FVec<V>& d, const FVec<V>& p, uint32_t e, float af, float tf, uint32_t gm) {
    const float a = __shfl_sync(0xffffffffu, af, e), t = __shfl_sync(0xffffffffu, tf, e);
    if (!((gm >> e) & 1u)) { vfma<V>(s, p, a); vfma<V>(d, p, t); }
    else m_general<V>(s, d, p, a, t);
}
// (the same ring of U row registers as the E walk: entry e lives in slot e % U and the slot is refilled as soon as it is consumed)
template <int V, int U>
__device__ __forceinline__ void m_walk(FVec<V>& s, FVec<V>& d, const uint32_t* __restrict__ q, const float2* __restrict__ aux, uint32_t n,
                                       const Packing& pk, const float* __restrict__ w, uint32_t row_bytes, uint32_t hm) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t w_next = lane < n ? __ldg(q + lane) : 0u;
#pragma unroll 1
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t cnt = min(32u, n - base);
        const uint32_t wd = w_next;
        w_next = base + 32 + lane < n ? __ldg(q + base + 32 + lane) : 0u;
        // every CSC word carries both count fields (a simple word: alt 0/1, alt+ref 1), so the decode is uniform
        const uint32_t ai = (wd >> pk.sh_a) & pk.cmask, ti = (wd >> pk.sh_r) & pk.cmask;
        float af = small_u2f(ai), tf = small_u2f(ti);
        if (ai == pk.cmask && ti == pk.cmask && lane < cnt) {    // counts outside the packed fields (rare)
            const float2 x = __ldg(aux + base + lane);
            af = x.x; tf = x.y;
        }
        const uint32_t cell = wd & pk.imask;
        const uint32_t gm = __ballot_sync(0xffffffffu, !(pow2_or_zero(af) && pow2_or_zero(tf)));
        FVec<V> p[U];
#pragma unroll
        for (int u = 0; u < U; ++u) ld_row<V>(p[u], row_at(w, __shfl_sync(0xffffffffu, cell, u), row_bytes), hm);   // (lanes >= cnt hold cell 0: a valid, unused gather)
        uint32_t e = 0;
#pragma unroll 1
        for (; e + 2 * U <= cnt; e += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                m_consume<V>(s, d, p[u], e + u, af, tf, gm);
                ld_row<V>(p[u], row_at(w, __shfl_sync(0xffffffffu, cell, e + U + u), row_bytes), hm);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e + u < cnt) {
                m_consume<V>(s, d, p[u], e + u, af, tf, gm);
                if (e + U + u < cnt) ld_row<V>(p[u], row_at(w, __shfl_sync(0xffffffffu, cell, e + U + u), row_bytes), hm);
            }
        }
        e += U;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (e + u < cnt) m_consume<V>(s, d, p[u], e + u, af, tf, gm);
    }
}

// Long loci.  A locus covered by most cells is a 10^4..10^5-long strictly sequential chain per column, and the M pass cannot end
// before its longest chain does: with U gathers in flight a chain advances one entry per (gather latency / U) cycles (measured: at 25
// resident solves the pass took 85 % of the time it takes at 50 — the chains, not the throughput, set its duration).  So the long loci
// get narrow column groups (many warps per locus, few instructions per entry and warp) and a ring of U = 16 rows that is refilled
// ACROSS chunk boundaries — the next chunk is decoded while the current one is served, so the chain never drains — and the literal
// expression for every entry (at 1-2 columns per lane the multiplies cost less than a class test).
template <int V, int U, bool GUARD>
__device__ __forceinline__ void m_deep_chunk(FVec<V>& s, FVec<V>& d, FVec<V> (&p)[U], uint32_t cnt, uint32_t left_after,
                                             uint32_t cell, float af, float tf, uint32_t cell2,
                                             const float* __restrict__ w, uint32_t row_bytes, uint32_t hm) {
    static_assert(32 % U == 0, "the ring must divide a chunk");
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        FVec<V>& slot = p[i % U];
        if (!GUARD || (uint32_t)i < cnt) m_general<V>(s, d, slot, __shfl_sync(0xffffffffu, af, i), __shfl_sync(0xffffffffu, tf, i));
        const int j = i + U;                    // the stream entry that takes the slot
        if (j < 32) {
            if (!GUARD || (uint32_t)j < cnt) ld_row<V>(slot, row_at(w, __shfl_sync(0xffffffffu, cell, j), row_bytes), hm);
        } else {
            if (!GUARD || (uint32_t)(j - 32) < left_after) ld_row<V>(slot, row_at(w, __shfl_sync(0xffffffffu, cell2, j - 32), row_bytes), hm);
        }
    }
}
template <int V, int U>
__device__ __forceinline__ void m_walk_deep(FVec<V>& s, FVec<V>& d, const uint32_t* __restrict__ q, const float2* __restrict__ aux, uint32_t n,
                                            const Packing& pk, const float* __restrict__ w, uint32_t row_bytes, uint32_t hm) {
    const uint32_t lane = threadIdx.x & 31;
    auto decode = [&](uint32_t base, uint32_t& cell, float& af, float& tf) {       // chunk at `base` (may lie beyond the end: zeros)
        const uint32_t wd = base + lane < n ? __ldg(q + base + lane) : 0u;
        const uint32_t ai = (wd >> pk.sh_a) & pk.cmask, ti = (wd >> pk.sh_r) & pk.cmask;
        af = small_u2f(ai); tf = small_u2f(ti);
        if (ai == pk.cmask && ti == pk.cmask && base + lane < n) {                 // counts outside the packed fields (rare)
            const float2 x = __ldg(aux + base + lane);
            af = x.x; tf = x.y;
        }
        cell = wd & pk.imask;
    };
    uint32_t cell, cell2;
    float af, tf, af2, tf2;
    decode(0, cell, af, tf);
    FVec<V> p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) ld_row<V>(p[u], row_at(w, __shfl_sync(0xffffffffu, cell, u), row_bytes), hm);   // (beyond the end: cell 0, a valid unused gather)
#pragma unroll 1
    for (uint32_t base = 0; base < n; base += 32) {
        decode(base + 32, cell2, af2, tf2);
        const uint32_t left = n - base;
        if (left >= 64) m_deep_chunk<V, U, false>(s, d, p, 32u, 32u, cell, af, tf, cell2, w, row_bytes, hm);
        else m_deep_chunk<V, U, true>(s, d, p, min(left, 32u), left > 32u ? left - 32u : 0u, cell, af, tf, cell2, w, row_bytes, hm);
        cell = cell2; af = af2; tf = tf2;
    }
}

struct MArgs {
    const uint32_t* csc_idx; const float2* csc_aux; const uint64_t* col_ptr; const uint32_t* locus_order;
    const float* W; float* TH; float* LA; float* LB; const uint8_t* col_write;
    const int* group_active;        // wide role: column groups of 32*VM
    const int* long_group_active;   // long role: column groups of 32*SOUP_M_VLONG
    const Ctl* ctl;
    // cell-sharded mode: partial sums of this rank's cells, [L][2][scols]: row 2*locus holds sum p*alt, row 2*locus + 1 holds
    // sum p*(alt+ref), column = table column (else nullptr).  scols (a multiple of 4) covers every live column.  Rows are locus ids,
    // the same on every rank, so a range of loci is one contiguous block for the all-reduce.
    float* SP;
    uint32_t scols;
    uint32_t li_begin, li_end;      // wide role: this launch covers the positions [li_begin, li_end) of locus_order
    Packing pk;
    uint32_t L, Cpad, n_groups;
    uint32_t n_long, long_groups, long_blocks, wide_chunks, group_major;   // loci [0, n_long) of the LPT order take the long role
    uint32_t no_narrow;
    // long loci by the cooperative kernel (mstep_long_kernel): 0 never, 1 while at most tail_cols columns are live, 2 always.
    // With few live columns it also takes the shorter loci [n_long, n_long_tail) — their one-warp chains would be the tail's critical path.
    uint32_t coop_mode, tail_cols, n_long_tail;
    // compact weight columns when clusters are locked (PostArgs): column cc of W is (slot cc / nd, re-drawn cluster k_of_draw[cc % nd])
    const int* k_of_draw;
    uint32_t K, nd;
};
// the prefix of W's columns that holds live weights, and the table column a weight column belongs to
__device__ __forceinline__ uint32_t m_live_cols(const MArgs& a) { return a.nd == a.K ? a.ctl->live_cols : a.ctl->live_cols / a.K * a.nd; }
// (kept out of line: it runs once per output column and only when clusters are locked, and inlining it costs the hot loops registers)
__device__ __noinline__ uint32_t m_table_col_mapped(const int* __restrict__ k_of_draw, uint32_t K, uint32_t nd, uint32_t cc) {
    const uint32_t slot = cc / nd;
    return slot * K + (uint32_t)k_of_draw[cc - slot * nd];
}
__device__ __forceinline__ uint32_t m_table_col(const MArgs& a, uint32_t cc) {
    return a.nd == a.K ? cc : m_table_col_mapped(a.k_of_draw, a.K, a.nd, cc);
}
__device__ __forceinline__ bool m_coop(const MArgs& a, uint32_t live) { return a.coop_mode == 2 || (a.coop_mode == 1 && live <= a.tail_cols); }
__device__ __forceinline__ uint32_t m_coop_loci(const MArgs& a, uint32_t live) { return live <= a.tail_cols ? a.n_long_tail : a.n_long; }

template <int V, bool DEEP>
__device__ __forceinline__ void mstep_body(const MArgs& a, uint32_t li, uint32_t col0, uint32_t live_cols) {
    const uint32_t locus = a.locus_order[li];
    const uint32_t lane = threadIdx.x & 31;
    const LaneCols<V> lc = lane_setup<V>(col0, live_cols);
    const uint64_t j0 = a.col_ptr[locus];
    const uint32_t n = (uint32_t)(a.col_ptr[locus + 1] - j0);
    const bool partial = a.SP != nullptr;
    FVec<V> s, d;
#pragma unroll
    for (int v = 0; v < V; ++v) { s.v[v] = partial ? 0.0f : 1.0f; d.v[v] = partial ? 0.0f : 2.0f; }   // pseudocounts S:197-198, S:456-464
    constexpr int U = V <= 2 ? SOUP_M_U2 : (V == 4 ? SOUP_M_U : (V == 8 ? SOUP_M_U8 : SOUP_M_U16));
    if constexpr (DEEP) m_walk_deep<V, SOUP_M_ULONG>(s, d, a.csc_idx + j0, a.csc_aux + j0, n, a.pk, a.W + lc.col, a.Cpad * 4u, lc.hm);
    else if constexpr (V <= 2 && V <= SOUP_DEEP_VMAX) m_walk_deep<V, 32 / V>(s, d, a.csc_idx + j0, a.csc_aux + j0, n, a.pk, a.W + lc.col, a.Cpad * 4u, lc.hm);
    else m_walk<V, U>(s, d, a.csc_idx + j0, a.csc_aux + j0, n, a.pk, a.W + lc.col, a.Cpad * 4u, lc.hm);
    if (!lc.live) return;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if ((v >> 2) >= 1 && !((lc.hm >> (v >> 2)) & 1u)) continue;   // dead sub-block
        uint32_t c = col_of<V>(col0, lane, v);
        if (a.nd != a.K) {                              // compact weight columns: skip what lies beyond the live ones, map the rest back
            if (c >= live_cols) continue;
            c = m_table_col(a, c);
        }
        if (partial) {   // cell-sharded: hand the partial statistics to the all-reduce; theta_kernel finishes the M pass
            if (c >= a.scols) continue;
            const size_t so = (size_t)(2u * locus) * a.scols + c;
            a.SP[so] = s.v[v];
            a.SP[so + a.scols] = d.v[v];
            continue;
        }
        const size_t o = (size_t)locus * a.Cpad + c;
        if (!a.col_write[c]) continue;                // locked cluster (S:389) or idle column
        const float th = fmaxf(fminf(__fdiv_rn(s.v[v], d.v[v]), 0.99f), 0.01f);   // S:392-393
        a.TH[o] = th;
        a.LA[o] = soupmath::glibc_logf(th);
        a.LB[o] = soupmath::glibc_logf(1.0f - th);
    }
}

// Two roles in one grid.  Blocks [0, long_blocks): the n_long longest loci (LPT order), each split into
// narrow column groups so that many warps share a long locus and each warp's sequential chain is
// short per entry; they launch first.  Remaining blocks: all other loci with wide column groups
// where the per-entry cost is amortised over more columns (narrowed to the live columns like the E pass).
template <int VM>
__global__ void __launch_bounds__(EM_THREADS, (VM <= 8 ? SOUP_M_MINB : 2) * (256 / EM_THREADS)) mstep_kernel(const MArgs a) {
    const uint32_t warp = EM_WARPS > 1 ? warp_id_uniform() : 0;
    const uint32_t live = m_live_cols(a);
    const bool coop = m_coop(a, live);
    if (blockIdx.x < a.long_blocks) {
        if (coop) return;                       // mstep_long_kernel owns the long loci
        const uint32_t item = blockIdx.x * EM_WARPS + warp;
        if (item >= a.n_long * a.long_groups) return;
        const uint32_t li = item / a.long_groups, group = item - li * a.long_groups;
        const uint32_t col0 = group * (32 * SOUP_M_VLONG);
        if (col0 >= live || (a.nd == a.K && !a.long_group_active[group])) return;   // (the activity masks are in table columns)
        mstep_body<SOUP_M_VLONG, true>(a, li, col0, live);
        return;
    }
    // wide role: locus-major (group fastest) while the whole weight plane fits L2 — every group's longer loci then start
    // early; group-major (all loci of group 0, then group 1, ...) once it does not, so one group's slice stays in L2
    const uint32_t bi = blockIdx.x - a.long_blocks;
    uint32_t group, chunk;
    if (a.group_major) { group = bi / a.wide_chunks; chunk = bi - group * a.wide_chunks; }
    else { chunk = bi / a.n_groups; group = bi - chunk * a.n_groups; }
    const uint32_t col0 = group * (32 * VM);
    if (col0 >= live || (a.nd == a.K && !a.group_active[group])) return;
    // (the loci below `first` belong to the long role / the cooperative kernel; which of the two is decided on the device, so the grid
    // always starts at li_begin and the warps in front of `first` leave)
    const uint32_t first = coop ? m_coop_loci(a, live) : a.n_long;
    const uint32_t li = a.li_begin + chunk * EM_WARPS + warp;
    if (li < first || li >= a.li_end) return;
    const int v = fit_lanes<VM>(live - col0, a.no_narrow);
    if (VM >= 16 && v == 16) mstep_body<(VM >= 16 ? 16 : VM), false>(a, li, col0, live);
    else if (VM >= 8 && v == 8) mstep_body<(VM >= 8 ? 8 : VM), false>(a, li, col0, live);
    else if (VM >= 4 && v == 4) mstep_body<(VM >= 4 ? 4 : VM), false>(a, li, col0, live);
    else if (VM >= 2 && v == 2) mstep_body<(VM >= 2 ? 2 : VM), false>(a, li, col0, live);
    else mstep_body<1, false>(a, li, col0, live);
}

// ------------------------------------------------------------------ M pass, long loci: one CTA per (locus, 256 columns)
// A locus covered by most cells is a 10^4..10^5-long strictly sequential chain per column.  Walked by one warp
// (mstep_kernel's long role) every entry costs that warp its whole decode + address + gather sequence, and the row
// gathers of a batch cannot overlap the adds of the previous one.  Here the chain is split by function instead:
//   * all 256 threads stream the weight rows of the next chunks of entries into shared memory with cp.async
//     (16-byte pieces, L2 -> smem, no registers held while in flight) and pre-decode the chunk's counts to floats;
//   * thread t then owns column col0 + t and adds the staged rows strictly in entry order with the reference's
//     literal expression  s += p*alt; d += p*(alt+ref)  — 6 instructions per entry on the critical path.
// Results are bit-identical to the one-warp walk (same operations in the same order per column).
constexpr int ML_THREADS = 256;
constexpr int ML_MAX_STAGES = 8;
constexpr int ML_MAX_CHUNK = 64;
constexpr uint32_t ML_ROWS = 8;                 // rows of a chunk one thread copies, at most
#ifndef SOUP_ML_SMEM_KB
#define SOUP_ML_SMEM_KB 96                     // dynamic shared memory of the row stages (opt-in above 48 KB)
#endif
constexpr int ML_SMEM_FLOATS = SOUP_ML_SMEM_KB * 256;
#ifndef SOUP_ML_FULL_STAGES
#define SOUP_ML_FULL_STAGES 4                  // stages when all 256 columns are live (chunk = smem / stages / 1 KB)
#endif

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(uint32_t pending) {   // the immediate form only: dispatch on the (CTA-uniform) count
    switch (pending) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        default: cp_async_wait<7>(); break;
    }
}

__global__ void __launch_bounds__(ML_THREADS) mstep_long_kernel(const MArgs a) {
    extern __shared__ __align__(16) float rows[];                 // ML_SMEM_FLOATS floats of row stages
    __shared__ float2 counts[ML_MAX_STAGES][ML_MAX_CHUNK];
    const uint32_t live = min(m_live_cols(a), a.Cpad);
    if (!m_coop(a, live)) return;
    // few live columns: one CTA per locus (the live columns fit one CTA); otherwise one per (locus, 256 columns)
    const uint32_t n_groups = live <= a.tail_cols ? 1 : (a.Cpad + ML_THREADS - 1) / ML_THREADS;
    const uint32_t li = blockIdx.x / n_groups, group = blockIdx.x - li * n_groups;
    const uint32_t col0 = group * ML_THREADS;
    if (li >= m_coop_loci(a, live) || col0 >= live) return;
    const uint32_t locus = a.locus_order[li];
    const uint32_t rem = min(live - col0, (uint32_t)ML_THREADS);
    const uint32_t pieces = (rem + 3) >> 2;              // 16-byte pieces per staged row
    const uint32_t stride = pieces << 2;                 // floats per staged row
    // entries per chunk and chunks in flight: SOUP_ML_FULL_STAGES stages when every column is live; narrower rows
    // (few live solves: the chain is then pure latency) first lengthen the chunk to 64 entries, then add stages
    const uint32_t rows_per_pass = ML_THREADS / pieces;  // rows the CTA copies per pass (thread -> one 16-byte piece)
    const uint32_t chunk = min(min((uint32_t)ML_SMEM_FLOATS / (SOUP_ML_FULL_STAGES * stride), (uint32_t)ML_MAX_CHUNK), ML_ROWS * rows_per_pass);
    const uint32_t stages = min((uint32_t)ML_SMEM_FLOATS / (chunk * stride), (uint32_t)ML_MAX_STAGES);
    const uint64_t j0 = a.col_ptr[locus];
    const uint32_t n = (uint32_t)(a.col_ptr[locus + 1] - j0);
    const uint32_t* __restrict__ q = a.csc_idx + j0;
    const float2* __restrict__ aux = a.csc_aux + j0;
    const uint32_t n_chunks = (n + chunk - 1) / chunk;
    const uint32_t t = threadIdx.x;
    const bool partial = a.SP != nullptr;
    float s = partial ? 0.0f : 1.0f, d = partial ? 0.0f : 2.0f;   // pseudocounts S:197-198, S:456-464
    const uint32_t r0 = t / pieces, pc = t - r0 * pieces;   // thread -> (row, piece) of a pass

    // A thread copies piece `pc` of rows r0, r0 + rows_per_pass, ... of a chunk: at most ML_ROWS of them
    // (the chunk length is clamped to that).  Their entry words are fetched one chunk ahead (`wn`), so the
    // address of a copy never waits on the index stream.
    uint32_t wn[ML_ROWS];
    const bool copier = t < rows_per_pass * pieces;
    auto fetch_words = [&](uint32_t c) {
        const uint32_t base = c * chunk, cnt = c < n_chunks ? min(chunk, n - base) : 0;
#pragma unroll
        for (int k = 0; k < (int)ML_ROWS; ++k) {
            const uint32_t r = r0 + k * rows_per_pass;
            if (copier && r < cnt) wn[k] = __ldg(q + base + r);
        }
    };
    auto issue = [&](uint32_t c) {   // wn holds the words of chunk c; leaves those of chunk c + 1 behind
        const uint32_t base = c * chunk, cnt = min(chunk, n - base), stage = c % stages;
        float* dst = rows + stage * chunk * stride;
#pragma unroll
        for (int k = 0; k < (int)ML_ROWS; ++k) {
            const uint32_t r = r0 + k * rows_per_pass;
            if (copier && r < cnt)
                cp_async16(dst + r * stride + (pc << 2), a.W + (size_t)(wn[k] & a.pk.imask) * a.Cpad + col0 + (pc << 2));
        }
        if (t < cnt) {
            const uint32_t w = __ldg(q + base + t);
            const uint32_t ai = (w >> a.pk.sh_a) & a.pk.cmask, ti = (w >> a.pk.sh_r) & a.pk.cmask;
            float2 x = make_float2(small_u2f(ai), small_u2f(ti));
            if (ai == a.pk.cmask && ti == a.pk.cmask) x = __ldg(aux + base + t);   // counts outside the packed fields (rare)
            counts[stage][t] = x;
        }
        fetch_words(c + 1);
    };

    fetch_words(0);
    for (uint32_t c = 0; c < stages - 1; ++c) {
        if (c < n_chunks) issue(c);
        cp_async_commit();
    }
    for (uint32_t c = 0; c < n_chunks; ++c) {
        if (c + stages - 1 < n_chunks) issue(c + stages - 1);          // its stage was consumed in iteration c - 1 (barrier below)
        cp_async_commit();
        cp_async_wait_dyn(stages - 1);                                 // this thread's pieces of chunk c have landed
        __syncthreads();                                               // ... and everybody else's, and the decoded counts
        if (t < rem) {
            const uint32_t base = c * chunk, cnt = min(chunk, n - base), stage = c % stages;
            const float* src = rows + stage * chunk * stride + t;
#pragma unroll 4
            for (uint32_t r = 0; r < cnt; ++r) {
                const float p = src[r * stride];
                const float2 x = counts[stage][r];
                s = __fadd_rn(s, __fmul_rn(p, x.x));                   // S:478-481, literal
                d = __fadd_rn(d, __fmul_rn(p, x.y));
            }
        }
        __syncthreads();
    }
    if (t >= rem) return;
    const uint32_t col = m_table_col(a, col0 + t);
    if (partial) {
        if (col < a.scols) { const size_t so = (size_t)(2u * locus) * a.scols + col; a.SP[so] = s; a.SP[so + a.scols] = d; }
        return;
    }
    const size_t o = (size_t)locus * a.Cpad + col;
    if (!a.col_write[col]) return;                                     // locked cluster (S:389) or idle column
    const float th = fmaxf(fminf(__fdiv_rn(s, d), 0.99f), 0.01f);      // S:392-393
    a.TH[o] = th;
    a.LA[o] = soupmath::glibc_logf(th);
    a.LB[o] = soupmath::glibc_logf(1.0f - th);
}

// cell-sharded mode, after the all-reduce of (SP, DP): theta = clamp((1 + S) / (2 + D)) and the ln tables.
// The pseudocount joins the globally reduced sum here, so the rounding differs from the sequential reference
// (and from the single-GPU path) in the last bits: this mode agrees to tolerance, not bit for bit.
__global__ void __launch_bounds__(256) theta_kernel(const float* __restrict__ stats, uint32_t scols, uint32_t row0, uint32_t rows,
                                                    const uint8_t* __restrict__ col_write,
                                                    float* __restrict__ TH, float* __restrict__ LA, float* __restrict__ LB, uint32_t Cpad) {
    // the loci [row0, row0 + rows) of the reduced statistics (MArgs::SP layout), table columns [0, scols)
    const size_t n = (size_t)rows * scols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / scols), c = (uint32_t)(i - (size_t)r * scols);
        if (!col_write[c]) continue;
        const size_t so = (size_t)(2u * (row0 + r)) * scols + c;
        const float th = fmaxf(fminf(__fdiv_rn(__fadd_rn(1.0f, stats[so]), __fadd_rn(2.0f, stats[so + scols])), 0.99f), 0.01f);
        const size_t o = (size_t)(row0 + r) * Cpad + c;
        TH[o] = th;
        LA[o] = soupmath::glibc_logf(th);
        LB[o] = soupmath::glibc_logf(1.0f - th);
    }
}

// ------------------------------------------------------------------ harvest: keep the best solve (S:110-114)
// `src_of_k` (or nullptr = identity): where cluster k of the model lives — engine column slot*Kx + src_of_k[k], or, for a cluster locked
// during a souporcell3 re-run (src < 0), the per-run ELLfix column / its unchanged row of base_theta (PostArgs)
__global__ void harvest_kernel(const Ctl* __restrict__ ctl, const float* __restrict__ ELL, const float* __restrict__ TH,
                               float* __restrict__ best_lp, float* __restrict__ best_theta, uint32_t N, uint32_t L,
                               uint32_t Cpad, int K, int Kx, const int* __restrict__ src_of_k, const float* __restrict__ ELLfix, int Kf,
                               const float* __restrict__ base_theta) {
    const int slot = ctl->copy_slot;
    if (slot < 0) return;
    const size_t nlp = (size_t)N * K, nth = (size_t)L * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nlp + nth; i += (size_t)gridDim.x * blockDim.x) {
        if (i < nlp) {
            const size_t cell = i / K, k = i - cell * K;
            const int src = src_of_k ? src_of_k[k] : (int)k;
            best_lp[i] = src >= 0 ? ELL[cell * Cpad + (size_t)slot * Kx + src] : ELLfix[cell * Kf + (-src - 1)];
        } else {
            const size_t e = i - nlp, l = e / K, k = e - l * K;
            const int src = src_of_k ? src_of_k[k] : (int)k;
            best_theta[k * L + l] = src >= 0 ? TH[l * Cpad + (size_t)slot * Kx + src] : base_theta[k * L + l];
        }
    }
}

// ELLfix[cell][j] = ELL[cell][j] for j < Kf (the one-off E pass over the locked clusters of a re-run)
__global__ void copy_fixed_kernel(const float* __restrict__ ELL, float* __restrict__ ELLfix, uint32_t N, uint32_t Cpad, int Kf) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * Kf) return;
    const size_t cell = i / Kf, j = i - cell * Kf;
    ELLfix[i] = ELL[cell * Cpad + j];
}
// ln tables of the locked clusters in columns [0, Kf): theta = base_theta[fix_k[j]][l]
__global__ void fixed_tables_kernel(const float* __restrict__ base_theta, const int* __restrict__ fix_k, float* LA, float* LB, Ctl* ctl,
                                    uint32_t L, uint32_t Cpad, int Kf) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // e = l*Kf + j
    if (e >= (size_t)L * Kf) return;
    const size_t l = e / Kf, j = e - l * Kf;
    const float th = base_theta[(size_t)fix_k[j] * L + l];
    const float la = soupmath::glibc_logf(th), lb = soupmath::glibc_logf(1.0f - th);
    LA[l * Cpad + j] = la;
    LB[l * Cpad + j] = lb;
    if (!isfinite(la) || !isfinite(lb)) ctl->nonfinite = 1;
}

// ------------------------------------------------------------------ refill: theta0 of the next solve (S:485-498, 554-563)
struct RefillArgs {
    SlotState* st;
    Ctl* ctl;
    const int* queue;             // [queue_len] global solve indices
    const uint32_t* stream_keys;  // [n_streams][8] ChaCha keys of the ThreadData rngs (S:52)
    const float* theta0;          // [n_solves][K][L] or nullptr
    const float* base_theta;      // [K][L] or nullptr (S:93)
    const float* init_theta;      // [K][L] or nullptr: deterministic initial centers (known genotypes, S:487)
    const int* draw_of_k;         // [K]: position of cluster k in this run's draw order, -1 = locked (S:92-97)
    float *TH, *LA, *LB;
    uint32_t L, Cpad;
    int K, n_draw, solves_per_stream;
    // init_cluster_centers_random_assignment (S:565-594): the host walks each seed stream once to place its three
    // phases — n_draw*L f32 draws, one gen_range per cell (its word count is data dependent), n_draw*L f32 draws —
    // and hands over, per queued solve, the cells' clusters and the word offsets of the two f32 phases
    const uint16_t* assign;       // [queue_len][N] or nullptr
    const uint64_t* word_sums;    // [queue_len]: stream word of `rng.gen::<f32>()*0.01` for (draw 0, locus 0)   S:572
    const uint64_t* word_noise;   // [queue_len]: stream word of `rng.gen::<f32>()/2.0 - 0.25` for (draw 0, locus 0)   S:588
    const uint32_t* csc_idx; const float2* csc_aux; const uint64_t* col_ptr;
    Packing pk;
    uint32_t N;
};

__device__ __forceinline__ float chacha_f32(const uint32_t* key, uint64_t word) {   // rand 0.9 StandardUniform f32 at an absolute stream word
    uint32_t blk[16];
    chacha_block(key, word >> 4, 12, blk);
    uint32_t wv = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) if ((int)(word & 15) == i) wv = blk[i];
    return (float)(wv >> 8) * (1.0f / 16777216.0f);
}

// theta0[d][l] of the random-assignment init: sums = gen*0.01 then += alt of every cell assigned to d, cells ascending
// (the order `for cell in cell_data` visits them); denoms = 0.01 then += alt + ref (f32 adds of integer-valued floats)
__device__ __forceinline__ float random_assignment_theta(const RefillArgs& a, const uint32_t* key, int pending, int d, uint32_t l) {
    float sum = __fmul_rn(chacha_f32(key, a.word_sums[pending] + (uint64_t)d * a.L + l), 0.01f), den = 0.01f;
    const uint16_t* as = a.assign + (size_t)pending * a.N;
    for (uint64_t j = a.col_ptr[l]; j < a.col_ptr[l + 1]; ++j) {
        const uint32_t w = a.csc_idx[j];
        if (as[w & a.pk.imask] != (uint16_t)d) continue;
        const uint32_t ai = (w >> a.pk.sh_a) & a.pk.cmask, ti = (w >> a.pk.sh_r) & a.pk.cmask;
        float alt = (float)ai, ref = (float)(ti - ai);
        if (ai == a.pk.cmask && ti == a.pk.cmask) { const float2 x = a.csc_aux[j]; alt = x.x; ref = __fsub_rn(x.y, x.x); }
        sum = __fadd_rn(sum, alt);                       // S:582
        den = __fadd_rn(den, __fadd_rn(alt, ref));       // S:580, S:583
    }
    const float noise = __fsub_rn(__fdiv_rn(chacha_f32(key, a.word_noise[pending] + (uint64_t)d * a.L + l), 2.0f), 0.25f);
    return fmaxf(fminf(__fadd_rn(__fdiv_rn(sum, den), noise), 0.9999f), 0.0001f);   // S:588-589
}

__global__ void __launch_bounds__(256) refill_kernel(RefillArgs a) {
    const int slot = blockIdx.y;
    const SlotState s = a.st[slot];
    if (!s.fin || s.pending < 0) return;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)a.L * a.K; e += (size_t)gridDim.x * blockDim.x) {
    // e = l*K + k (k fastest: coalesced table writes)
    const uint32_t l = (uint32_t)(e / a.K);
    const int k = (int)(e - (size_t)l * a.K);
    const int solve = a.queue[s.pending];
    float th;
    const int d = a.draw_of_k[k];
    if (a.theta0) {
        th = a.theta0[((size_t)solve * a.K + k) * a.L + l];
    } else if (d < 0) {
        th = a.base_theta[(size_t)k * a.L + l];
    } else if (a.init_theta) {
        th = a.init_theta[(size_t)d * a.L + l];       // run 0: d == k; later runs: re-drawn cluster i takes row i (S:92-96)
    } else if (a.assign) {
        th = random_assignment_theta(a, a.stream_keys + (size_t)(solve / a.solves_per_stream) * 8, s.pending, d, l);
    } else {
        // init_cluster_centers_uniform S:554-563: cluster-major then locus draw order, one u32 word per f32
        const int stream = solve / a.solves_per_stream, it = solve - stream * a.solves_per_stream;
        const uint64_t word = ((uint64_t)it * a.n_draw + d) * a.L + l;
        uint32_t blk[16];
        chacha_block(a.stream_keys + (size_t)stream * 8, word >> 4, 12, blk);
        uint32_t wv = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) if ((int)(word & 15) == i) wv = blk[i];
        th = fmaxf(fminf((float)(wv >> 8) * (1.0f / 16777216.0f), 0.9999f), 0.0001f);
    }
    const size_t o = (size_t)l * a.Cpad + (size_t)slot * a.K + k;
    const float la = soupmath::glibc_logf(th), lb = soupmath::glibc_logf(1.0f - th);
    a.TH[o] = th;
    a.LA[o] = la;
    a.LB[o] = lb;
    if (!isfinite(la) || !isfinite(lb)) a.ctl->nonfinite = 1;
    }
}

// after refill: install pending solves, retire finished slots, rebuild the column / group masks
struct FinalizeArgs {
    SlotState* st;
    Ctl* ctl;
    const int* queue;
    const uint8_t* k_locked;   // [K]
    uint8_t* col_write;        // [Cpad]
    // column-group tables (E kernel, M wide role, M long role): groups [0, n_main) are WM columns wide;
    // an optional last group starts at n_main*WM and is WT wide
    struct Plan { int* active; int WM, WT, n_main, n_groups; } plan[3];
    int K, slots, iter_cap;
    uint32_t Cpad;
    float limit;
    int2* moves;               // [slots] (from, to) pairs for migrate_kernel
    int compact;               // 1: pack the surviving slots into the fewest column groups once the queue is empty
};

__global__ void __launch_bounds__(1024) finalize_kernel(FinalizeArgs a) {
    __shared__ int n_active;
    if (threadIdx.x == 0) n_active = 0;
    __syncthreads();
    for (int slot = threadIdx.x; slot < a.slots; slot += blockDim.x) {
        SlotState s = a.st[slot];
        if (s.fin) {
            s.fin = 0;
            if (s.pending >= 0) {
                s.local = s.pending;
                s.solve = a.queue[s.pending];
                s.active = 1;
                s.temp_step = next_temp_step(0, 0, a.iter_cap, a.limit);
                s.iterations = 0;
                s.last_loss = -INFINITY;
                s.loss = -INFINITY;
            } else {
                s.active = 0;
                s.solve = -1;
            }
            s.pending = -1;
            a.st[slot] = s;
        }
        if (s.active) atomicAdd(&n_active, 1);
    }
    __syncthreads();
    // Tail compaction.  Solves converge at different iterations; once the queue is empty the survivors would sit
    // scattered over all column groups, and a group with one live slot costs a full pass.  So the highest live slots
    // move into the lowest holes (disjoint source / destination sets: migrate_kernel copies all columns in parallel)
    // and the live slots always form a prefix; the E / M kernels size their work by Ctl::live_cols.  Only theta /
    // ln-table columns and the slot state move; everything else is recomputed every iteration.  Per-solve results do
    // not depend on the slot a solve sits in.
    if (threadIdx.x == 0) {
        int moves = 0;
        if (a.compact && a.ctl->queue_head >= a.ctl->queue_len && n_active > 0) {
            int hole = 0, top = a.slots - 1;
            for (;;) {
                while (hole < a.slots && a.st[hole].active) ++hole;
                while (top >= 0 && !a.st[top].active) --top;
                if (hole >= top) break;
                a.moves[moves++] = make_int2(top, hole);
                SlotState s = a.st[top];
                a.st[hole] = s;
                s.active = 0; s.solve = -1;
                a.st[top] = s;
            }
        }
        a.ctl->n_moves = moves;
        int top = a.slots - 1;
        while (top >= 0 && !a.st[top].active) --top;
        a.ctl->live_cols = (uint32_t)((top + 1) * a.K);
    }
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < a.Cpad; c += blockDim.x) {
        const int slot = (int)(c / a.K), k = (int)(c - (uint32_t)slot * a.K);
        a.col_write[c] = (slot < a.slots && a.st[slot].active && !a.k_locked[k]) ? 1 : 0;
    }
    for (int t = 0; t < 3; ++t) {
        const FinalizeArgs::Plan pl = a.plan[t];
        for (int g = threadIdx.x; g < pl.n_groups; g += blockDim.x) {
            const int c0 = g * pl.WM, c1 = g < pl.n_main ? c0 + pl.WM : c0 + pl.WT;
            const int first = c0 / a.K;
            int last = (c1 - 1) / a.K;
            if (last >= a.slots) last = a.slots - 1;
            int any = 0;
            for (int s = first; s <= last; ++s) any |= a.st[s].active;
            pl.active[g] = any;
        }
    }
    if (threadIdx.x == 0) a.ctl->n_active = n_active;
}

// copies the theta / ln-table columns of the migrated slots (blockIdx.y = move index)
__global__ void __launch_bounds__(256) migrate_kernel(const Ctl* __restrict__ ctl, const int2* __restrict__ moves, float* TH, float* LA, float* LB,
                                                      uint32_t L, uint32_t Cpad, int K) {
    if ((int)blockIdx.y >= ctl->n_moves) return;
    const int2 mv = moves[blockIdx.y];
    const size_t n = (size_t)L * K;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const size_t l = e / K, k = e - l * K;
        const size_t src = l * Cpad + (size_t)mv.x * K + k, dst = l * Cpad + (size_t)mv.y * K + k;
        TH[dst] = TH[src]; LA[dst] = LA[src]; LB[dst] = LB[src];
    }
}

// ------------------------------------------------------------------ load-time kernels
__constant__ double c_ln_factorial[171];   // statrs FCACHE: ln(n!) for n <= 170 (S:669-670)

struct LoadFlags { uint32_t bad_range, bad_order, bad_cell, n_zero, n_big, n_long, n_long_tail; };

__global__ void validate_csr_kernel(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ locus, const uint32_t* __restrict__ alt,
                                    const uint32_t* __restrict__ ref, LoadFlags* flags, uint32_t N, uint32_t L) {
    const uint32_t cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= N) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t j0 = row_ptr[cell], j1 = row_ptr[cell + 1];
    uint32_t zeros = 0, big = 0;
    for (uint64_t j = j0 + lane; j < j1; j += 32) {
        const uint32_t l = locus[j];
        if (l >= L) { flags->bad_range = 1; flags->bad_cell = cell; }
        if (j > j0 && l <= locus[j - 1]) { flags->bad_order = 1; flags->bad_cell = cell; }
        const uint32_t n = alt[j] + ref[j];
        zeros += n == 0;
        big += n > 170u;
    }
    if (zeros) atomicAdd(&flags->n_zero, zeros);
    if (big) atomicAdd(&flags->n_big, big);
}
// number of leading elements >= thr in a descending array
__global__ void count_ge_kernel(const uint32_t* __restrict__ sorted_desc, uint32_t n, uint32_t thr, uint32_t* out) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sorted_desc[mid] >= thr) lo = mid + 1; else hi = mid;
    }
    *out = lo;
}

// entry words (see "entry words" above); `simple` = exactly one UMI in total
__device__ __forceinline__ uint32_t pack_general(uint32_t idx, uint32_t c0, uint32_t c1, const Packing& pk) {
    if (c0 >= pk.cmask || c1 >= pk.cmask) c0 = c1 = pk.cmask;   // out-of-field: marker, the side record holds the counts
    return 0x80000000u | (c0 << pk.sh_a) | (c1 << pk.sh_r) | idx;
}
// CSR words: bits [0, sh_r) of EVERY word are the first row of the joint ln table the entry reads — locus (alt != 0, or counts beyond
// the fields) or plane_rows + locus (ref only); an entry with both alleles reads row + plane_rows as well.
__device__ __forceinline__ uint32_t pack_csr_word(uint32_t locus, uint32_t alt, uint32_t ref, const Packing& pk) {
    if (ref == 0 && (alt == 1 || alt == 2)) return locus | (alt == 2 ? CSR_DOUBLE : 0u);
    if (alt == 0 && (ref == 1 || ref == 2)) return (pk.plane_rows + locus) | (ref == 2 ? CSR_DOUBLE : 0u);
    const bool overflow = alt >= pk.cmask || ref >= pk.cmask;
    return pack_general(alt != 0 || overflow ? locus : pk.plane_rows + locus, alt, ref, pk);
}
__device__ __forceinline__ uint32_t pack_csc_word(uint32_t cell, uint32_t alt, uint32_t ref, const Packing& pk) {
    if (alt == 1 && ref == 0) return cell | (1u << pk.sh_r) | (1u << pk.sh_a);
    if (alt == 0 && ref == 1) return cell | (1u << pk.sh_r);
    return pack_general(cell, alt, alt + ref, pk);   // alt + ref: the reference's wrapping integer add, then cast (S:480)
}

__global__ void build_csr_kernel(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ locus,
                                 const uint32_t* __restrict__ alt, const uint32_t* __restrict__ ref, uint32_t* __restrict__ csr_idx,
                                 float4* __restrict__ csr_aux, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                                 float* __restrict__ total_alleles, Packing pk, uint32_t N) {
    const uint32_t cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= N) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t j0 = row_ptr[cell], j1 = row_ptr[cell + 1];
    for (uint64_t j = j0 + lane; j < j1; j += 32) {
        const uint32_t a = alt[j], r = ref[j], n = a + r;   // wrapping u32 add as in release-mode Rust (S:664, S:670)
        float lbc = 0.0f;
        if (n <= 170u) lbc = (float)(c_ln_factorial[n] - c_ln_factorial[a] - c_ln_factorial[n - a]);   // n > 170 patched from the host
        csr_idx[j] = pack_csr_word(locus[j], a, r, pk);
        csr_aux[j] = make_float4((float)a, (float)r, lbc, 0.0f);
        keys[j] = ((unsigned long long)locus[j] << 32) | cell;
        vals[j] = (uint32_t)j;
    }
    if (lane == 0) {
        float t = 0.0f;                                      // total_alleles S:671: sequential f32, ascending locus
        for (uint64_t j = j0; j < j1; ++j) t = t + (float)(alt[j] + ref[j]);
        total_alleles[cell] = t;
    }
}

__global__ void patch_lbc_kernel(float4* csr_aux, const uint32_t* __restrict__ idx, const float* __restrict__ val, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) csr_aux[idx[i]].z = val[i];
}

// csc stream in (locus, cell) order; alt+ref is the reference's integer add, then cast (S:480)
__global__ void build_csc_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ vals,
                                 const uint32_t* __restrict__ alt, const uint32_t* __restrict__ ref, uint32_t* __restrict__ csc_idx,
                                 float2* __restrict__ csc_aux, Packing pk, uint64_t nnz) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const uint32_t j = vals[i];
    const uint32_t a = alt[j], t = a + ref[j];
    csc_idx[i] = pack_csc_word((uint32_t)(keys[i] & 0xffffffffu), a, t - a, pk);
    csc_aux[i] = make_float2((float)a, (float)t);
}

__global__ void col_ptr_kernel(const unsigned long long* __restrict__ keys, uint64_t nnz, uint64_t* __restrict__ col_ptr, uint32_t L) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l > L) return;
    const unsigned long long target = (unsigned long long)l << 32;   // first key with locus >= l
    uint64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1; else hi = mid;
    }
    col_ptr[l] = lo;
}
__global__ void seg_len_kernel(const uint64_t* __restrict__ ptr, uint32_t* __restrict__ len, uint32_t* __restrict__ idx, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    len[i] = (uint32_t)(ptr[i + 1] - ptr[i]);
    idx[i] = i;
}

// ------------------------------------------------------------------ downstream sufficient statistics
// cluster_allele_counts[locus][cluster] = {sum ref, sum alt} over the cells assigned to the cluster (consensus.py:296-331):
// one warp per cell over the raw CSR arrays, exact 64-bit integer atomics (order cannot matter)
__global__ void cluster_counts_kernel(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ locus, const uint32_t* __restrict__ alt,
                                      const uint32_t* __restrict__ ref, const int32_t* __restrict__ cell_cluster, unsigned long long* counts,
                                      uint32_t N, int K) {
    const uint32_t cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= N) return;
    const int c = cell_cluster[cell];
    if (c < 0 || c >= K) return;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t j = row_ptr[cell] + lane; j < row_ptr[cell + 1]; j += 32) {
        unsigned long long* o = counts + ((size_t)locus[j] * K + c) * 2;
        if (ref[j]) atomicAdd(o, (unsigned long long)ref[j]);
        if (alt[j]) atomicAdd(o + 1, (unsigned long long)alt[j]);
    }
}

// ------------------------------------------------------------------ test hooks
__global__ void math_kernel(int op, const float* __restrict__ in, float* __restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = op == 0 ? soupmath::glibc_logf(in[i]) : soupmath::glibc_expf(in[i]);
}

// theta [n_slots][K][L] (host layout) -> TH / LA / LB planes
__global__ void tables_from_theta_kernel(const float* __restrict__ theta, float* TH, float* LA, float* LB, Ctl* ctl,
                                         uint32_t L, uint32_t Cpad, int K, int slots) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // e = (l*slots + slot)*K + k
    if (e >= (size_t)L * slots * K) return;
    const int k = (int)(e % K);
    const size_t t = e / K;
    const int slot = (int)(t % slots);
    const size_t l = t / slots;
    const float th = theta[((size_t)slot * K + k) * L + l];
    const size_t o = l * Cpad + (size_t)slot * K + k;
    const float la = soupmath::glibc_logf(th), lb = soupmath::glibc_logf(1.0f - th);
    TH[o] = th; LA[o] = la; LB[o] = lb;
    if (!isfinite(la) || !isfinite(lb)) ctl->nonfinite = 1;
}
// [rows][Cpad] plane -> host layout [slots][rows][K] (row_major_k = 1) or [slots][K][rows] (0)
__global__ void plane_to_host_kernel(const float* __restrict__ plane, float* __restrict__ out, uint32_t rows, uint32_t Cpad, int K,
                                     int slots, int row_major_k) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)rows * slots * K) return;
    const int k = (int)(e % K);
    const size_t t = e / K;
    const int slot = (int)(t % slots);
    const size_t r = t / slots;
    const float v = plane[r * Cpad + (size_t)slot * K + k];
    if (row_major_k) out[((size_t)slot * rows + r) * K + k] = v;
    else out[((size_t)slot * K + k) * rows + r] = v;
}
__global__ void host_to_plane_kernel(const float* __restrict__ in, float* __restrict__ plane, uint32_t rows, uint32_t Cpad, int K, int slots) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // in: [slots][rows][K]
    if (e >= (size_t)rows * slots * K) return;
    const int k = (int)(e % K);
    const size_t t = e / K;
    const int slot = (int)(t % slots);
    const size_t r = t / slots;
    plane[r * Cpad + (size_t)slot * K + k] = in[((size_t)slot * rows + r) * K + k];
}

}  // namespace soup
