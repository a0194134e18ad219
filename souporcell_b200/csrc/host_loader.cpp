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
#include <cstdlib>
#include <cstring>
#include <thread>

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
uint64_t count_lines(const char* b, const char* e) {   // number of lines BufRead::lines() yields for [b, e)
    uint64_t n = 0;
    const char* p = b;
    while (p < e) {
        const void* q = memchr(p, '\n', (size_t)(e - p));
        if (!q) { ++n; break; }
        ++n;
        p = (const char*)q + 1;
    }
    return n;
}

struct Chunk { size_t a_begin, r_begin; uint64_t first_line, n_lines; };

struct ParseError { std::atomic<int> code{0}; char msg[256]; void set(int c, const char* m) { int z = 0; if (code.compare_exchange_strong(z, c)) { strncpy(msg, m, sizeof(msg) - 1); msg[sizeof(msg) - 1] = 0; } } };

void parse_chunk(const Mapped& A, const Mapped& R, const Chunk& ch, uint64_t total_loci, uint64_t total_cells,
                 uint32_t* locus, uint32_t* cell, uint32_t* alt, uint32_t* ref, ParseError* err) {
    const char* ap = A.p + ch.a_begin;
    const char* rp = R.p + ch.r_begin;
    const char* aend = A.p + A.n;
    const char* rend = R.p + R.n;
    for (uint64_t i = 0; i < ch.n_lines; ++i) {
        const char* ae = (const char*)memchr(ap, '\n', (size_t)(aend - ap));
        if (!ae) ae = aend;
        const char* re = (const char*)memchr(rp, '\n', (size_t)(rend - rp));
        if (!re) re = rend;
        const char *tb, *te;
        uint64_t l, c, av, rv;
        const char* p = ap;
        if (!next_token(p, ae, tb, te) || !parse_u64(tb, te, &l) || !next_token(p, ae, tb, te) || !parse_u64(tb, te, &c) ||
            !next_token(p, ae, tb, te) || !parse_u64(tb, te, &av) || av > UINT32_MAX) {
            err->set(SOUP_ERR_INVALID, "alt mtx: malformed entry line (reference panics on unwrap)");
            return;
        }
        p = rp;
        if (!next_token(p, re, tb, te) || !next_token(p, re, tb, te) || !next_token(p, re, tb, te) || !parse_u64(tb, te, &rv) ||
            rv > UINT32_MAX) {
            err->set(SOUP_ERR_INVALID, "ref mtx: malformed entry line (reference panics on unwrap)");
            return;
        }
        if (l == 0 || l - 1 >= total_loci) { err->set(SOUP_ERR_INVALID, "assertion failed: locus < total_loci"); return; }
        if (c == 0 || c - 1 >= total_cells) { err->set(SOUP_ERR_INVALID, "assertion failed: cell < total_cells"); return; }
        uint64_t o = ch.first_line + i;
        locus[o] = (uint32_t)(l - 1);
        cell[o] = (uint32_t)(c - 1);
        alt[o] = (uint32_t)av;
        ref[o] = (uint32_t)rv;
        ap = ae < aend ? ae + 1 : aend;
        rp = re < rend ? re + 1 : rend;
    }
}

// Byte offsets of `want.size()` ascending line indices (relative to body start) inside body [b, e).
std::vector<size_t> locate_lines(const char* base, size_t b, size_t e, const std::vector<uint64_t>& want, unsigned nthreads) {
    // pass 1: newline counts per byte block, in parallel
    size_t len = e - b;
    size_t nblk = std::max<size_t>(1, std::min<size_t>(nthreads * 4, len / (1 << 20) + 1));
    size_t blk = (len + nblk - 1) / nblk;
    std::vector<uint64_t> cnt(nblk, 0);
    std::vector<std::thread> pool;
    std::atomic<size_t> next{0};
    for (unsigned t = 0; t < nthreads; ++t)
        pool.emplace_back([&] {
            for (size_t i; (i = next.fetch_add(1)) < nblk;) {
                size_t s = b + i * blk, f = std::min(e, s + blk);
                uint64_t n = 0;
                for (const char* p = base + s; p < base + f;) {
                    const void* q = memchr(p, '\n', (size_t)(base + f - p));
                    if (!q) break;
                    ++n;
                    p = (const char*)q + 1;
                }
                cnt[i] = n;
            }
        });
    for (auto& th : pool) th.join();
    std::vector<size_t> out(want.size());
    uint64_t lines_before = 0;
    size_t bi = 0;
    for (size_t w = 0; w < want.size(); ++w) {
        uint64_t target = want[w];   // offset of the start of line `target` = just past newline #target
        if (target == 0) { out[w] = b; continue; }
        while (bi < nblk && lines_before + cnt[bi] < target) { lines_before += cnt[bi]; ++bi; }
        if (bi >= nblk) { out[w] = e; continue; }
        size_t s = b + bi * blk, f = std::min(e, s + blk);
        uint64_t need = target - lines_before;
        const char* p = base + s;
        while (need > 0) {
            const void* q = memchr(p, '\n', (size_t)(base + f - p));
            p = (const char*)q + 1;   // guaranteed to exist: cnt[bi] >= need
            --need;
        }
        out[w] = (size_t)(p - base);
    }
    return out;
}

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
                                 uint32_t min_alt_umis, uint32_t min_ref_umis, uint32_t flags, soup_mtx** out) {
    if (!alt_path || !ref_path || !out) { set_thread_error("soup_mtx_load: bad argument"); return SOUP_ERR_INVALID; }
    *out = nullptr;
    Mapped A, R;
    if (!A.open(alt_path)) { set_thread_error("cannot open alt mtx file"); return SOUP_ERR_IO; }   // S:602
    if (!R.open(ref_path)) { set_thread_error("cannot open ref mtx file"); return SOUP_ERR_IO; }   // S:605
    unsigned nthreads = std::max(1u, std::thread::hardware_concurrency());

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
    if (have_header) n_lines = std::min(count_lines(A.p + a3, A.p + A.n), count_lines(R.p + r3, R.p + R.n));

    // ---- parallel lockstep parse into COO (file order)
    std::vector<uint32_t> c_locus(n_lines), c_cell(n_lines), c_alt(n_lines), c_ref(n_lines);
    if (n_lines) {
        size_t nchunks = std::max<size_t>(1, std::min<size_t>(nthreads * 4, n_lines / 65536 + 1));
        std::vector<uint64_t> first(nchunks + 1);
        for (size_t i = 0; i <= nchunks; ++i) first[i] = n_lines * i / nchunks;
        std::vector<uint64_t> want(first.begin(), first.end() - 1);
        std::vector<size_t> aoff = locate_lines(A.p, a3, A.n, want, nthreads);
        std::vector<size_t> roff = locate_lines(R.p, r3, R.n, want, nthreads);
        std::vector<Chunk> chunks(nchunks);
        for (size_t i = 0; i < nchunks; ++i) chunks[i] = {aoff[i], roff[i], first[i], first[i + 1] - first[i]};
        ParseError err;
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads; ++t)
            pool.emplace_back([&] {
                for (size_t i; (i = next.fetch_add(1)) < nchunks && !err.code.load();)
                    parse_chunk(A, R, chunks[i], total_loci, total_cells, c_locus.data(), c_cell.data(), c_alt.data(), c_ref.data(), &err);
            });
        for (auto& th : pool) th.join();
        if (err.code.load()) { set_thread_error("%s", err.msg); return err.code.load(); }
    }

    // ---- per-locus tallies (wrapping u32, every line counted) and the filter  S:629-632, S:659
    std::vector<uint32_t> cells_ref(total_loci, 0), cells_alt(total_loci, 0), umis_ref(total_loci, 0), umis_alt(total_loci, 0);
    std::vector<uint8_t> seen(total_loci, 0);
    for (uint64_t i = 0; i < n_lines; ++i) {
        uint32_t l = c_locus[i];
        seen[l] = 1;
        if (c_ref[i] > 0) { cells_ref[l] += 1; umis_ref[l] += c_ref[i]; }
        if (c_alt[i] > 0) { cells_alt[l] += 1; umis_alt[l] += c_alt[i]; }
    }
    std::vector<uint32_t> remap(total_loci, UINT32_MAX);
    std::vector<uint64_t> index_to_locus;
    const bool raw = (flags & SOUP_MTX_RAW) != 0;
    for (uint64_t l = 0; l < total_loci; ++l)
        if (raw || (seen[l] && cells_ref[l] >= min_ref && cells_alt[l] >= min_alt && umis_ref[l] >= min_ref_umis && umis_alt[l] >= min_alt_umis)) {
            remap[l] = (uint32_t)index_to_locus.size();
            index_to_locus.push_back(l);
        }

    // ---- stable counting sort of the kept lines by cell (file order preserved inside a cell)
    std::vector<uint64_t> start(total_cells + 1, 0);
    for (uint64_t i = 0; i < n_lines; ++i)
        if (remap[c_locus[i]] != UINT32_MAX) start[c_cell[i] + 1] += 1;
    for (uint64_t c = 0; c < total_cells; ++c) start[c + 1] += start[c];
    uint64_t kept = start[total_cells];
    struct Ent { uint32_t locus, alt, ref; };
    std::vector<Ent> ents(kept);
    {
        std::vector<uint64_t> cur(start.begin(), start.end() - 1);
        for (uint64_t i = 0; i < n_lines; ++i) {
            uint32_t nl = remap[c_locus[i]];
            if (nl != UINT32_MAX) ents[cur[c_cell[i]]++] = {nl, c_alt[i], c_ref[i]};
        }
    }
    std::vector<uint32_t>().swap(c_locus); std::vector<uint32_t>().swap(c_cell);
    std::vector<uint32_t>().swap(c_alt); std::vector<uint32_t>().swap(c_ref);

    // ---- per cell: ascending locus, duplicates last-wins (S:634), drop ref+alt == 0 (S:664)
    std::vector<uint64_t> row_len(total_cells, 0);
    {
        std::atomic<uint64_t> next{0};
        std::vector<std::thread> pool;
        const uint64_t step = 256;
        for (unsigned t = 0; t < nthreads; ++t)
            pool.emplace_back([&] {
                for (uint64_t c0; (c0 = next.fetch_add(step)) < total_cells;)
                    for (uint64_t c = c0; c < std::min(total_cells, c0 + step); ++c) {
                        Ent* b = ents.data() + start[c];
                        Ent* e = ents.data() + start[c + 1];
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
        for (auto& th : pool) th.join();
    }
    uint64_t nnz = 0;
    for (uint64_t c = 0; c < total_cells; ++c) nnz += row_len[c];

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
    uint64_t pos = 0;
    for (uint64_t c = 0; c < total_cells; ++c) {
        m->row_ptr[c] = pos;
        const Ent* b = ents.data() + start[c];
        for (uint64_t j = 0; j < row_len[c]; ++j, ++pos) {
            m->locus[pos] = b[j].locus;
            m->alt[pos] = b[j].alt;
            m->ref[pos] = b[j].ref;
        }
    }
    m->row_ptr[total_cells] = pos;
    if (!index_to_locus.empty()) memcpy(m->index_to_locus, index_to_locus.data(), index_to_locus.size() * sizeof(uint64_t));
    *out = m;
    return SOUP_OK;
}


// ------------------------------------------------------------------ binary CSR cache
// `.soupcsr`: what soup_mtx_load produced, so that re-runs of the pipeline (and troublet / consensus, which parse the same two
// text matrices again) skip the text parse.  Little-endian, 64-byte header, then the five arrays back to back.
namespace {
struct CacheHeader {
    char magic[8];            // "SOUPCSR1"
    uint64_t n_cells, n_loci_total, n_loci, nnz;
    uint32_t min_alt, min_ref, min_alt_umis, min_ref_umis;   // the filter the matrix was built with (S:659)
    uint32_t flags;           // SOUP_MTX_RAW or 0
    uint32_t reserved;
};
static_assert(sizeof(CacheHeader) == 64, "cache header layout");
}  // namespace

extern "C" int32_t soup_mtx_save(const soup_mtx* m, const char* path, uint32_t min_alt, uint32_t min_ref, uint32_t min_alt_umis,
                                 uint32_t min_ref_umis, uint32_t flags) {
    if (!m || !path) { set_thread_error("soup_mtx_save: bad argument"); return SOUP_ERR_INVALID; }
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { set_thread_error("soup_mtx_save: cannot create %s", tmp.c_str()); return SOUP_ERR_IO; }
    CacheHeader h{};
    memcpy(h.magic, "SOUPCSR1", 8);
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

extern "C" int32_t soup_mtx_load_cache(const char* path, uint32_t min_alt, uint32_t min_ref, uint32_t min_alt_umis, uint32_t min_ref_umis,
                                       uint32_t flags, soup_mtx** out) {
    if (!path || !out) { set_thread_error("soup_mtx_load_cache: bad argument"); return SOUP_ERR_INVALID; }
    *out = nullptr;
    Mapped F;
    if (!F.open(path)) { set_thread_error("cannot open %s", path); return SOUP_ERR_IO; }
    if (F.n < sizeof(CacheHeader)) { set_thread_error("%s: not a .soupcsr file", path); return SOUP_ERR_INVALID; }
    CacheHeader h;
    memcpy(&h, F.p, sizeof(h));
    if (memcmp(h.magic, "SOUPCSR1", 8) != 0) { set_thread_error("%s: not a .soupcsr file", path); return SOUP_ERR_INVALID; }
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
