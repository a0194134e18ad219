// Matrix Market loader + locus filter -> CSR over cells.  Replaces load_cell_data
// (/root/reference/souporcell/src/main.rs:601-681, "S:") with a hash-free, multi-threaded
// mmap parser that keeps the reference's observable semantics:
//   * both files are walked in lockstep by line (izip!, S:616) and stop at the shorter one;
//   * lines 0 and 1 are ignored, line 2 of the ALT file is "loci cells nnz" (S:635-639);
//   * locus and cell come from the alt line, only the value from the ref line (S:620-626);
//   * per-locus tallies count every line, duplicates included (S:629-632), in wrapping u32
//     (release-mode Rust arithmetic); the stored counts of a duplicate are last-wins (S:634);
//   * keep locus iff cells(ref>0) >= min_ref && cells(alt>0) >= min_alt && sum(ref) >=
//     min_ref_umis && sum(alt) >= min_alt_umis (S:659); kept loci are numbered densely in
//     ascending original id (S:648-675); entries with ref+alt == 0 are dropped (S:664).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include "host_parallel.hpp"
#include "soup_common.hpp"

namespace soup {
void* pinned_alloc(size_t bytes);   // defined in soup_device.cu (cudaHostAlloc)
void pinned_free(void* p);
}

using namespace soup;

namespace {

struct Mapped {
    const char* p = nullptr;
    size_t n = 0;
    int fd = -1;
    std::vector<char> inflated;   // ".gz" input (an extension: the reference opens the matrices uncompressed, S:602-607)
    bool open(const char* path) {
        const size_t plen = strlen(path);
        if (plen >= 3 && !strcmp(path + plen - 3, ".gz")) {
            gzFile f = gzopen(path, "rb");
            if (!f) return false;
            gzbuffer(f, 1 << 20);
            char buf[1 << 16];
            int got;
            while ((got = gzread(f, buf, sizeof(buf))) > 0) inflated.insert(inflated.end(), buf, buf + got);
            const bool ok = got == 0;
            gzclose(f);
            if (!ok) return false;
            n = inflated.size();
            p = n ? inflated.data() : "";
            return true;
        }
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = (size_t)st.st_size;
        if (n == 0) { p = ""; return true; }
        void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        madvise(m, n, MADV_SEQUENTIAL);
        p = (const char*)m;
        return true;
    }
    ~Mapped() {
        if (p && n && inflated.empty()) munmap((void*)p, n);
        if (fd >= 0) close(fd);
    }
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// Next whitespace-delimited token inside [p, end) (one line, '\n' excluded).  false at end of line.
inline bool next_token(const char*& p, const char* end, const char*& tb, const char*& te) {
    while (p < end && is_ws(*p)) ++p;
    if (p >= end) return false;
    tb = p;
    while (p < end && !is_ws(*p)) ++p;
    te = p;
    return true;
}
inline bool parse_u64(const char* b, const char* e, uint64_t* out) {
    if (b < e && *b == '+') ++b;   // Rust FromStr for unsigned ints accepts a leading '+'
    if (b >= e) return false;
    uint64_t v = 0;
    for (; b < e; ++b) {
        unsigned d = (unsigned)(*b - '0');
        if (d > 9) return false;
        if (v > (UINT64_MAX - d) / 10) return false;
        v = v * 10 + d;
    }
    *out = v;
    return true;
}

// offset just past the `k`-th '\n' at or after `from` (k >= 1), or n if the file ends first
size_t skip_lines(const Mapped& m, size_t from, size_t k) {
    size_t pos = from;
    while (k > 0 && pos < m.n) {
        const void* q = memchr(m.p + pos, '\n', m.n - pos);
        if (!q) return m.n;
        pos = (size_t)((const char*)q - m.p) + 1;
        --k;
    }
    return pos;
}
// Newline counts of a file body [b, e) per byte block, built in parallel: how many lines BufRead::lines() yields, and where
// line k starts, without a serial pass over the text.
struct NewlineIndex {
    size_t b = 0, e = 0, blk = 1;
    std::vector<uint64_t> before;   // before[i] = '\n' bytes in blocks [0, i); size nblk + 1
    uint64_t lines = 0;
};
NewlineIndex index_newlines(const char* base, size_t b, size_t e, unsigned nthreads) {
    NewlineIndex ix;
    ix.b = b; ix.e = e;
    const size_t len = e - b;
    ix.blk = 256 << 10;
    const size_t nblk = (len + ix.blk - 1) / ix.blk;
    ix.before.assign(nblk + 1, 0);
    parallel_for(nthreads, nblk, [&](size_t i) {
        const char* p = base + b + i * ix.blk;
        const char* f = base + std::min(e, b + (i + 1) * ix.blk);
        uint64_t n = 0;
        for (; p < f; ++p) n += (*p == '\n');   // vectorises
        ix.before[i + 1] = n;
    });
    for (size_t i = 0; i < nblk; ++i) ix.before[i + 1] += ix.before[i];
    ix.lines = ix.before[nblk] + ((len > 0 && base[e - 1] != '\n') ? 1 : 0);   // an unterminated last line counts
    return ix;
}
// Byte offset of the start of line `target` (0-based, relative to the body start): just past newline number `target`.
size_t line_offset(const char* base, const NewlineIndex& ix, uint64_t target) {
    if (target == 0) return ix.b;
    if (target > ix.before.back()) return ix.e;
    const size_t i = (size_t)(std::lower_bound(ix.before.begin(), ix.before.end(), target) - ix.before.begin()) - 1;   // before[i] < target <= before[i + 1]
    uint64_t need = target - ix.before[i];
    const char* p = base + ix.b + i * ix.blk;
    const char* f = base + std::min(ix.e, ix.b + (i + 1) * ix.blk);
    while (need > 0) {
        p = (const char*)memchr(p, '\n', (size_t)(f - p)) + 1;   // exists: the block holds at least `need` newlines
        --need;
    }
    return (size_t)(p - base);
}

struct Chunk { size_t a_begin, r_begin; uint64_t first_line, n_lines; };

struct ParseError { std::atomic<int> code{0}; char msg[256]; void set(int c, const char* m) { int z = 0; if (code.compare_exchange_strong(z, c)) { strncpy(msg, m, sizeof(msg) - 1); msg[sizeof(msg) - 1] = 0; } } };

// locus tallies of S:629-632, shared by the parse workers: relaxed atomic adds (wrapping u32 sums commute, so the totals equal
// the sequential ones).  One set for the whole matrix — the common-variants VCFs the pipeline feeds vartrix have tens of millions of
// loci, so a copy per thread is not an option.
struct Tally {
    std::unique_ptr<std::atomic<uint32_t>[]> cells_ref, cells_alt, umis_ref, umis_alt;
    std::unique_ptr<std::atomic<uint8_t>[]> seen;
    void init(uint64_t n, unsigned nthreads) {
        cells_ref.reset(new std::atomic<uint32_t>[n + 1]); cells_alt.reset(new std::atomic<uint32_t>[n + 1]);
        umis_ref.reset(new std::atomic<uint32_t>[n + 1]); umis_alt.reset(new std::atomic<uint32_t>[n + 1]);
        seen.reset(new std::atomic<uint8_t>[n + 1]);
        const uint64_t span = 1 << 16;
        parallel_for(nthreads, (size_t)((n + span - 1) / span), [&](size_t b) {
            for (uint64_t l = b * span; l < std::min<uint64_t>(n, (b + 1) * span); ++l) {
                cells_ref[l].store(0, std::memory_order_relaxed); cells_alt[l].store(0, std::memory_order_relaxed);
                umis_ref[l].store(0, std::memory_order_relaxed); umis_alt[l].store(0, std::memory_order_relaxed);
                seen[l].store(0, std::memory_order_relaxed);
            }
        });
    }
};

// ---- one line of each file.  The fast forms cover what vartrix writes ("l c v\n": 1-9 digit fields, single blanks or tabs,
// optional CR); anything else — signs, padding, extra tokens, long numbers — goes through the general tokenizer, which decides alone
// what is an error.  Both advance `p` past the line.
inline bool fast_uint(const char*& p, const char* end, uint64_t* out) {
    const char* s = p;
    uint32_t v = 0;
    while (p < end) {
        const unsigned d = (unsigned)(unsigned char)*p - '0';
        if (d > 9) break;
        v = v * 10 + d;
        ++p;
    }
    const size_t n = (size_t)(p - s);
    if (n == 0 || n > 9) return false;
    *out = v;
    return true;
}
inline bool fast_blanks(const char*& p, const char* end) {
    const char* s = p;
    while (p < end && (*p == ' ' || *p == '\t')) ++p;
    return p > s;
}
inline bool fast_token(const char*& p, const char* end) {   // a token whose content is not looked at
    const char* s = p;
    while (p < end && !is_ws(*p) && *p != '\n') ++p;
    return p > s;
}
inline bool fast_eol(const char*& p, const char* end) {
    if (p < end && *p == '\r') ++p;
    if (p == end) return true;
    if (*p != '\n') return false;
    ++p;
    return true;
}
inline bool alt_line(const char*& p, const char* end, uint64_t* l, uint64_t* c, uint64_t* v) {
    const char* q = p;
    if (fast_uint(q, end, l) && fast_blanks(q, end) && fast_uint(q, end, c) && fast_blanks(q, end) && fast_uint(q, end, v) && fast_eol(q, end)) {
        p = q;
        return true;
    }
    const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!e) e = end;
    const char *tb, *te;
    q = p;
    p = e < end ? e + 1 : end;
    return next_token(q, e, tb, te) && parse_u64(tb, te, l) && next_token(q, e, tb, te) && parse_u64(tb, te, c) &&
           next_token(q, e, tb, te) && parse_u64(tb, te, v) && *v <= UINT32_MAX;
}
inline bool ref_line(const char*& p, const char* end, uint64_t* v) {   // only the third token is ever parsed (S:625)
    const char* q = p;
    if (fast_token(q, end) && fast_blanks(q, end) && fast_token(q, end) && fast_blanks(q, end) && fast_uint(q, end, v) && fast_eol(q, end)) {
        p = q;
        return true;
    }
    const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!e) e = end;
    const char *tb, *te;
    q = p;
    p = e < end ? e + 1 : end;
    return next_token(q, e, tb, te) && next_token(q, e, tb, te) && next_token(q, e, tb, te) && parse_u64(tb, te, v) && *v <= UINT32_MAX;
}

// `raw` (troublet's view, T:312-320): locus and cell are read from the REF line too and must equal the alt line's — the reference
// asserts "allele matrices do not match"; souporcell itself never looks at them (S:622-626).
void parse_chunk(const Mapped& A, const Mapped& R, const Chunk& ch, uint64_t total_loci, uint64_t total_cells,
                 uint32_t* locus, uint32_t* cell, uint32_t* alt, uint32_t* ref, Tally& t, ParseError* err, bool raw) {
    const char* ap = A.p + ch.a_begin;
    const char* rp = R.p + ch.r_begin;
    const char* aend = A.p + A.n;
    const char* rend = R.p + R.n;
    for (uint64_t i = 0; i < ch.n_lines; ++i) {
        uint64_t l, c, av, rv;
        if (!alt_line(ap, aend, &l, &c, &av)) { err->set(SOUP_ERR_INVALID, "alt mtx: malformed entry line (reference panics on unwrap)"); return; }
        if (!raw) {
            if (!ref_line(rp, rend, &rv)) { err->set(SOUP_ERR_INVALID, "ref mtx: malformed entry line (reference panics on unwrap)"); return; }
        } else {
            uint64_t rl, rc;
            if (!alt_line(rp, rend, &rl, &rc, &rv)) { err->set(SOUP_ERR_INVALID, "ref mtx: malformed entry line (reference panics on unwrap)"); return; }
            if (rl != l || rc != c) { err->set(SOUP_ERR_INVALID, "allele matrices do not match"); return; }
        }
        if (l == 0 || l - 1 >= total_loci) { err->set(SOUP_ERR_INVALID, "assertion failed: locus < total_loci"); return; }
        if (c == 0 || c - 1 >= total_cells) { err->set(SOUP_ERR_INVALID, "assertion failed: cell < total_cells"); return; }
        const uint64_t o = ch.first_line + i;
        const uint32_t li = (uint32_t)(l - 1);
        locus[o] = li;
        cell[o] = (uint32_t)(c - 1);
        alt[o] = (uint32_t)av;
        ref[o] = (uint32_t)rv;
        if (!t.seen[li].load(std::memory_order_relaxed)) t.seen[li].store(1, std::memory_order_relaxed);
        if (rv > 0) { t.cells_ref[li].fetch_add(1, std::memory_order_relaxed); t.umis_ref[li].fetch_add((uint32_t)rv, std::memory_order_relaxed); }
        if (av > 0) { t.cells_alt[li].fetch_add(1, std::memory_order_relaxed); t.umis_alt[li].fetch_add((uint32_t)av, std::memory_order_relaxed); }
    }
}

// SOUP_DEBUG_LOAD=1: wall time of each loader phase on stderr
struct LoadTimer {
    const bool on = getenv("SOUP_DEBUG_LOAD") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "soup_debug load %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

void* alloc_out(size_t bytes, bool pinned) {
    if (bytes == 0) bytes = 8;
    return pinned ? pinned_alloc(bytes) : malloc(bytes);
}

}  // namespace

extern "C" void soup_mtx_free(soup_mtx* m) {
    if (!m) return;
    void* ptrs[] = {m->row_ptr, m->locus, m->alt, m->ref, m->index_to_locus};
    for (void* p : ptrs)
        if (p) { if (m->pinned) pinned_free(p); else free(p); }
    free(m);
}

extern "C" int32_t soup_mtx_load(const char* alt_path, const char* ref_path, uint32_t min_alt, uint32_t min_ref,
                                 uint32_t min_alt_umis, uint32_t min_ref_umis, uint32_t flags, soup_mtx** out) try {
    if (!alt_path || !ref_path || !out) { set_thread_error("soup_mtx_load: bad argument"); return SOUP_ERR_INVALID; }
    *out = nullptr;
    Mapped A, R;
    {   // both at once: a ".gz" input is inflated while it is opened
        bool ok_a = false;
        std::thread ta([&] { ok_a = A.open(alt_path); });
        const bool ok_r = R.open(ref_path);
        ta.join();
        if (!ok_a) { set_thread_error("cannot open alt mtx file"); return SOUP_ERR_IO; }   // S:602
        if (!ok_r) { set_thread_error("cannot open ref mtx file"); return SOUP_ERR_IO; }   // S:605
    }
    unsigned nthreads = host_threads();   // at most 64
    if (const char* e = getenv("SOUP_LOADER_THREADS")) nthreads = (unsigned)std::min(256, std::max(1, atoi(e)));   // tests / tuning
    LoadTimer lt;

    // header: the zip only reaches line 2 if both files have it
    size_t a2 = skip_lines(A, 0, 2), r2 = skip_lines(R, 0, 2);
    uint64_t total_loci = 0, total_cells = 0;
    size_t a3 = a2, r3 = r2;
    bool have_header = a2 < A.n && r2 < R.n;
    if (have_header) {
        const char* e = (const char*)memchr(A.p + a2, '\n', A.n - a2);
        if (!e) e = A.p + A.n;
        const char *p = A.p + a2, *tb, *te;
        if (!next_token(p, e, tb, te) || !parse_u64(tb, te, &total_loci) || !next_token(p, e, tb, te) || !parse_u64(tb, te, &total_cells)) {
            set_thread_error("alt mtx: malformed size line (line 3)");
            return SOUP_ERR_INVALID;
        }
        a3 = skip_lines(A, a2, 1);
        r3 = skip_lines(R, r2, 1);
    }
    if (total_loci > UINT32_MAX || total_cells > UINT32_MAX) { set_thread_error("matrix dimensions above 2^32 are not supported"); return SOUP_ERR_UNSUPPORTED; }
    uint64_t n_lines = 0;
    NewlineIndex ia, ir;
    if (have_header) {
        ia = index_newlines(A.p, a3, A.n, nthreads);
        ir = index_newlines(R.p, r3, R.n, nthreads);
        n_lines = std::min(ia.lines, ir.lines);   // izip! stops at the shorter file (S:616)
    }
    lt.lap("open + index lines");

    // ---- parallel lockstep parse into COO (file order) + per-thread locus tallies (S:629-632: every line counts, wrapping u32)
    std::unique_ptr<uint32_t[]> c_locus(new uint32_t[n_lines + 1]), c_cell(new uint32_t[n_lines + 1]), c_alt(new uint32_t[n_lines + 1]),
        c_ref(new uint32_t[n_lines + 1]);
    const size_t nchunks = std::max<size_t>(1, std::min<size_t>(nthreads * 4, n_lines / 65536 + 1));
    const unsigned nworkers = (unsigned)std::min<size_t>(nthreads, nchunks);
    Tally tally;
    tally.init(total_loci, nthreads);
    if (n_lines) {
        std::vector<Chunk> chunks(nchunks);
        parallel_for(nthreads, nchunks, [&](size_t i) {
            const uint64_t first = n_lines * i / nchunks, last = n_lines * (i + 1) / nchunks;
            chunks[i] = {line_offset(A.p, ia, first), line_offset(R.p, ir, first), first, last - first};
        });
        ParseError err;
        std::atomic<size_t> next{0};
        parallel_for(nworkers, nworkers, [&](size_t) {   // the workers share the chunk queue
            for (size_t i; (i = next.fetch_add(1)) < nchunks && !err.code.load();)
                parse_chunk(A, R, chunks[i], total_loci, total_cells, c_locus.get(), c_cell.get(), c_alt.get(), c_ref.get(), tally, &err, (flags & SOUP_MTX_RAW) != 0);
        });
        if (err.code.load()) { set_thread_error("%s", err.msg); return err.code.load(); }
    }
    lt.lap("parse");

    // ---- the filter  S:659 over the merged tallies; kept loci numbered densely in ascending original id (S:648-675)
    std::vector<uint8_t> keep(total_loci, 0);
    const bool raw = (flags & SOUP_MTX_RAW) != 0;
    {
        const size_t span = 1 << 14, nspan = (size_t)((total_loci + span - 1) / span);
        parallel_for(nthreads, nspan, [&](size_t b) {
            for (uint64_t l = b * span; l < std::min<uint64_t>(total_loci, (b + 1) * span); ++l) {
                const uint32_t cr = tally.cells_ref[l].load(std::memory_order_relaxed), ca = tally.cells_alt[l].load(std::memory_order_relaxed);
                const uint32_t ur = tally.umis_ref[l].load(std::memory_order_relaxed), ua = tally.umis_alt[l].load(std::memory_order_relaxed);
                const bool seen = tally.seen[l].load(std::memory_order_relaxed) != 0;
                keep[l] = raw || (seen && cr >= min_ref && ca >= min_alt && ur >= min_ref_umis && ua >= min_alt_umis);
            }
        });
    }
    tally = Tally();
    std::vector<uint32_t> remap(total_loci, UINT32_MAX);
    std::vector<uint64_t> index_to_locus;
    for (uint64_t l = 0; l < total_loci; ++l)
        if (keep[l]) {
            remap[l] = (uint32_t)index_to_locus.size();
            index_to_locus.push_back(l);
        }
    lt.lap("tallies + filter");

    // ---- stable counting sort of the kept lines by cell (file order preserved inside a cell), in parallel: every part owns a
    // contiguous range of lines and its own histogram over the cells; part p's slots of cell c follow those of parts < p.
    const size_t parts = (size_t)std::max<uint64_t>(1, std::min<uint64_t>(nthreads, n_lines / 65536 + 1));
    std::vector<std::vector<uint64_t>> hist(parts);
    parallel_for(nthreads, parts, [&](size_t p) {
        std::vector<uint64_t>& h = hist[p];
        h.assign(total_cells, 0);
        for (uint64_t i = n_lines * p / parts; i < n_lines * (p + 1) / parts; ++i)
            if (remap[c_locus[i]] != UINT32_MAX) h[c_cell[i]] += 1;
    });
    std::vector<uint64_t> start(total_cells + 1, 0);
    {
        uint64_t run = 0;
        for (uint64_t c = 0; c < total_cells; ++c) {
            start[c] = run;
            for (size_t p = 0; p < parts; ++p) { const uint64_t n = hist[p][c]; hist[p][c] = run; run += n; }
        }
        start[total_cells] = run;
    }
    const uint64_t kept = start[total_cells];
    struct Ent { uint32_t locus, alt, ref; };
    std::unique_ptr<Ent[]> ents(new Ent[kept + 1]);
    parallel_for(nthreads, parts, [&](size_t p) {
        uint64_t* cur = hist[p].data();
        for (uint64_t i = n_lines * p / parts; i < n_lines * (p + 1) / parts; ++i) {
            const uint32_t nl = remap[c_locus[i]];
            if (nl != UINT32_MAX) ents[cur[c_cell[i]]++] = {nl, c_alt[i], c_ref[i]};
        }
    });
    std::vector<std::vector<uint64_t>>().swap(hist);
    c_locus.reset(); c_cell.reset(); c_alt.reset(); c_ref.reset();
    lt.lap("sort by cell");

    // ---- per cell: ascending locus, duplicates last-wins (S:634), drop ref+alt == 0 (S:664)
    std::vector<uint64_t> row_len(total_cells, 0);
    {
        const uint64_t step = 256;
        parallel_for(nthreads, (size_t)((total_cells + step - 1) / step), [&](size_t blk) {
            for (uint64_t c = blk * step; c < std::min<uint64_t>(total_cells, (blk + 1) * step); ++c) {
                Ent* b = ents.get() + start[c];
                Ent* e = ents.get() + start[c + 1];
                if (raw) { row_len[c] = (uint64_t)(e - b); continue; }   // T:334: every line, in file order
                bool sorted = true;
                for (Ent* q = b; q + 1 < e; ++q)
                    if (q[0].locus >= q[1].locus) { sorted = false; break; }
                if (!sorted) std::stable_sort(b, e, [](const Ent& x, const Ent& y) { return x.locus < y.locus; });
                Ent* w = b;
                for (Ent* q = b; q < e; ++q) {
                    if (q + 1 < e && q[1].locus == q[0].locus) continue;   // a later duplicate overrides
                    if (q->alt + q->ref == 0) continue;                    // wrapping u32 add, as S:664
                    *w++ = *q;
                }
                row_len[c] = (uint64_t)(w - b);
            }
        });
    }
    uint64_t nnz = 0;
    for (uint64_t c = 0; c < total_cells; ++c) nnz += row_len[c];
    lt.lap("per-cell order");

    bool pinned = (flags & SOUP_MTX_PINNED) != 0;
    soup_mtx* m = (soup_mtx*)calloc(1, sizeof(soup_mtx));
    if (!m) { set_thread_error("out of memory"); return SOUP_ERR_NOMEM; }
    m->pinned = pinned ? 1 : 0;
    m->n_cells = total_cells;
    m->n_loci_total = total_loci;
    m->n_loci = index_to_locus.size();
    m->nnz = nnz;
    m->row_ptr = (uint64_t*)alloc_out((total_cells + 1) * sizeof(uint64_t), pinned);
    m->locus = (uint32_t*)alloc_out(nnz * sizeof(uint32_t), pinned);
    m->alt = (uint32_t*)alloc_out(nnz * sizeof(uint32_t), pinned);
    m->ref = (uint32_t*)alloc_out(nnz * sizeof(uint32_t), pinned);
    m->index_to_locus = (uint64_t*)alloc_out(index_to_locus.size() * sizeof(uint64_t), pinned);
    if (!m->row_ptr || !m->locus || !m->alt || !m->ref || !m->index_to_locus) {
        soup_mtx_free(m);
        set_thread_error(pinned ? "pinned host allocation failed (no CUDA driver?)" : "out of memory");
        return pinned ? SOUP_ERR_CUDA : SOUP_ERR_NOMEM;
    }
    {
        uint64_t pos = 0;
        for (uint64_t c = 0; c < total_cells; ++c) { m->row_ptr[c] = pos; pos += row_len[c]; }
        m->row_ptr[total_cells] = pos;
        const uint64_t step = 256;
        parallel_for(nthreads, (size_t)((total_cells + step - 1) / step), [&](size_t blk) {
            for (uint64_t c = blk * step; c < std::min<uint64_t>(total_cells, (blk + 1) * step); ++c) {
                const Ent* b = ents.get() + start[c];
                uint64_t o = m->row_ptr[c];
                for (uint64_t j = 0; j < row_len[c]; ++j, ++o) {
                    m->locus[o] = b[j].locus;
                    m->alt[o] = b[j].alt;
                    m->ref[o] = b[j].ref;
                }
            }
        });
    }
    if (!index_to_locus.empty()) memcpy(m->index_to_locus, index_to_locus.data(), index_to_locus.size() * sizeof(uint64_t));
    lt.lap("write csr");
    *out = m;
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_mtx_load")


// ------------------------------------------------------------------ binary CSR cache
// `.soupcsr`: what soup_mtx_load produced, so that re-runs of the pipeline (and troublet / consensus, which parse the same two
// text matrices again) skip the text parse.  Little-endian, 64-byte header, then the five arrays back to back.
namespace {
struct CacheHeader {
    char magic[8];            // "SOUPCSR2"
    uint64_t n_cells, n_loci_total, n_loci, nnz;
    uint32_t min_alt, min_ref, min_alt_umis, min_ref_umis;   // the filter the matrix was built with (S:659)
    uint32_t flags;           // SOUP_MTX_RAW or 0
    uint32_t reserved;
    // the two text matrices the cache was parsed from (alt, ref): size, modification time, and a hash of their ends — a cache file
    // reused for OTHER matrices would otherwise print a clusters TSV for the wrong data without a word
    uint64_t src_size[2], src_mtime_ns[2], src_hash[2];
    uint64_t pad[2];
};
static_assert(sizeof(CacheHeader) == 128, "cache header layout");

// size / mtime / FNV-1a of the first and last 64 KB (+ the size): cheap, and enough to tell two vartrix outputs apart
bool fingerprint(const char* path, uint64_t* size, uint64_t* mtime_ns, uint64_t* hash) {
    struct stat st;
    if (stat(path, &st) != 0) return false;
    *size = (uint64_t)st.st_size;
    *mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const unsigned char* b, size_t n) { for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    std::vector<unsigned char> buf(64 * 1024);
    size_t got = fread(buf.data(), 1, buf.size(), f);
    mix(buf.data(), got);
    if (*size > buf.size()) {
        const uint64_t off = *size > 2 * buf.size() ? *size - buf.size() : buf.size();
        if (fseeko(f, (off_t)off, SEEK_SET) == 0) { got = fread(buf.data(), 1, buf.size(), f); mix(buf.data(), got); }
    }
    fclose(f);
    mix((const unsigned char*)size, sizeof(*size));
    *hash = h;
    return true;
}
}  // namespace

extern "C" int32_t soup_mtx_save(const soup_mtx* m, const char* path, const char* alt_path, const char* ref_path, uint32_t min_alt, uint32_t min_ref,
                                 uint32_t min_alt_umis, uint32_t min_ref_umis, uint32_t flags) try {
    if (!m || !path || !alt_path || !ref_path) { set_thread_error("soup_mtx_save: bad argument"); return SOUP_ERR_INVALID; }
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { set_thread_error("soup_mtx_save: cannot create %s", tmp.c_str()); return SOUP_ERR_IO; }
    CacheHeader h{};
    memcpy(h.magic, "SOUPCSR2", 8);
    const char* src[2] = {alt_path, ref_path};
    for (int i = 0; i < 2; ++i)
        if (!fingerprint(src[i], &h.src_size[i], &h.src_mtime_ns[i], &h.src_hash[i])) { fclose(f); remove(tmp.c_str()); set_thread_error("soup_mtx_save: cannot read %s", src[i]); return SOUP_ERR_IO; }
    h.n_cells = m->n_cells; h.n_loci_total = m->n_loci_total; h.n_loci = m->n_loci; h.nnz = m->nnz;
    h.min_alt = min_alt; h.min_ref = min_ref; h.min_alt_umis = min_alt_umis; h.min_ref_umis = min_ref_umis; h.flags = flags & SOUP_MTX_RAW;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    auto put = [&](const void* p, size_t bytes) { if (ok && bytes) ok = fwrite(p, 1, bytes, f) == bytes; };
    put(m->row_ptr, (m->n_cells + 1) * 8); put(m->locus, m->nnz * 4); put(m->alt, m->nnz * 4); put(m->ref, m->nnz * 4); put(m->index_to_locus, m->n_loci * 8);
    ok = (fclose(f) == 0) && ok;
    if (ok) ok = rename(tmp.c_str(), path) == 0;     // readers never see a half-written cache
    if (!ok) { remove(tmp.c_str()); set_thread_error("soup_mtx_save: write to %s failed", path); return SOUP_ERR_IO; }
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_mtx_save")

extern "C" int32_t soup_mtx_load_cache(const char* path, const char* alt_path, const char* ref_path, uint32_t min_alt, uint32_t min_ref,
                                       uint32_t min_alt_umis, uint32_t min_ref_umis, uint32_t flags, soup_mtx** out) try {
    if (!path || !out || !alt_path || !ref_path) { set_thread_error("soup_mtx_load_cache: bad argument"); return SOUP_ERR_INVALID; }
    *out = nullptr;
    Mapped F;
    if (!F.open(path)) { set_thread_error("cannot open %s", path); return SOUP_ERR_IO; }
    if (F.n < sizeof(CacheHeader)) { set_thread_error("%s: not a .soupcsr file", path); return SOUP_ERR_INVALID; }
    CacheHeader h;
    memcpy(&h, F.p, sizeof(h));
    if (memcmp(h.magic, "SOUPCSR2", 8) != 0) { set_thread_error("%s: not a .soupcsr file (or an older layout)", path); return SOUP_ERR_INVALID; }
    {   // the cache must have been parsed from THESE two matrices
        const char* src[2] = {alt_path, ref_path};
        for (int i = 0; i < 2; ++i) {
            uint64_t sz = 0, mt = 0, hs = 0;
            if (!fingerprint(src[i], &sz, &mt, &hs)) { set_thread_error("cannot open %s", src[i]); return SOUP_ERR_IO; }
            if (sz != h.src_size[i] || mt != h.src_mtime_ns[i] || hs != h.src_hash[i]) {
                set_thread_error("%s was built from another %s matrix than %s (size / mtime / content differ)", path, i == 0 ? "alt" : "ref", src[i]);
                return SOUP_ERR_INVALID;
            }
        }
    }
    if (h.min_alt != min_alt || h.min_ref != min_ref || h.min_alt_umis != min_alt_umis || h.min_ref_umis != min_ref_umis || h.flags != (flags & SOUP_MTX_RAW)) {
        set_thread_error("%s was built with another filter (min_alt %u min_ref %u min_alt_umis %u min_ref_umis %u raw %u)", path, h.min_alt, h.min_ref,
                         h.min_alt_umis, h.min_ref_umis, h.flags);
        return SOUP_ERR_INVALID;
    }
    const uint64_t need = sizeof(h) + (h.n_cells + 1) * 8 + h.nnz * 12 + h.n_loci * 8;
    if (h.n_cells > (1ull << 32) || h.nnz > (1ull << 40) || h.n_loci > h.n_loci_total || F.n != need) { set_thread_error("%s: truncated or corrupt", path); return SOUP_ERR_INVALID; }
    const bool pinned = (flags & SOUP_MTX_PINNED) != 0;
    soup_mtx* m = (soup_mtx*)calloc(1, sizeof(soup_mtx));
    if (!m) { set_thread_error("out of memory"); return SOUP_ERR_NOMEM; }
    m->pinned = pinned ? 1 : 0;
    m->n_cells = h.n_cells; m->n_loci_total = h.n_loci_total; m->n_loci = h.n_loci; m->nnz = h.nnz;
    m->row_ptr = (uint64_t*)alloc_out((h.n_cells + 1) * 8, pinned);
    m->locus = (uint32_t*)alloc_out(h.nnz * 4, pinned);
    m->alt = (uint32_t*)alloc_out(h.nnz * 4, pinned);
    m->ref = (uint32_t*)alloc_out(h.nnz * 4, pinned);
    m->index_to_locus = (uint64_t*)alloc_out(h.n_loci * 8, pinned);
    if (!m->row_ptr || !m->locus || !m->alt || !m->ref || !m->index_to_locus) { soup_mtx_free(m); set_thread_error("out of memory"); return pinned ? SOUP_ERR_CUDA : SOUP_ERR_NOMEM; }
    const char* q = F.p + sizeof(h);
    auto get = [&](void* dst, size_t bytes) { if (bytes) memcpy(dst, q, bytes); q += bytes; };
    get(m->row_ptr, (h.n_cells + 1) * 8); get(m->locus, h.nnz * 4); get(m->alt, h.nnz * 4); get(m->ref, h.nnz * 4); get(m->index_to_locus, h.n_loci * 8);
    // structural checks: a cache is an input like any other
    bool ok = m->row_ptr[0] == 0 && m->row_ptr[h.n_cells] == h.nnz;
    for (uint64_t c = 0; ok && c < h.n_cells; ++c) ok = m->row_ptr[c] <= m->row_ptr[c + 1];
    for (uint64_t j = 0; ok && j < h.nnz; ++j) ok = m->locus[j] < h.n_loci;
    for (uint64_t l = 0; ok && l < h.n_loci; ++l) ok = m->index_to_locus[l] < h.n_loci_total && (l == 0 || m->index_to_locus[l] > m->index_to_locus[l - 1]);
    if (!ok) { soup_mtx_free(m); set_thread_error("%s: inconsistent contents", path); return SOUP_ERR_INVALID; }
    *out = m;
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_mtx_load_cache")
