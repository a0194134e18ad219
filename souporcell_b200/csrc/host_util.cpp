// Host-side pieces of libsoupgpu that need no GPU: RNG streams, Rust-compatible number
// formatting, the ln_binomial table, barcodes, the TSV writer, souporcell3 outlier logic.
// Reference: /root/reference/souporcell/src/main.rs ("S:").
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <memory>
#include <mutex>
#include <zlib.h>

#include "host_parallel.hpp"
#include "soup_common.hpp"

namespace soup {

static thread_local std::string g_thread_error;

std::string vformat(const char* fmt, va_list ap) {
    char buf[1024];
    vsnprintf(buf, sizeof(buf), fmt, ap);
    return std::string(buf);
}
void set_thread_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    g_thread_error = vformat(fmt, ap);
    va_end(ap);
}
const char* thread_error() { return g_thread_error.c_str(); }

// statrs 0.11.0 function::factorial — FCACHE[i] = FCACHE[i-1] * i for i <= 170, ln_factorial =
// ln(FCACHE[x]) (x <= 170) else ln_gamma(x + 1); ln_binomial(n, k) = lf(n) - lf(k) - lf(n - k).
// For n > 170 statrs evaluates its own Lanczos ln_gamma (function::gamma, r = 10.900511, 11 terms; arguments here are
// >= 172, so only the x >= 0.5 branch exists); libm's lgamma differs from it in the last bits, which can flip the f32 rounding.
static double lanczos_ln_gamma(double x) {
    constexpr double kR = 10.900511;
    constexpr double kLn2SqrtEOverPi = 0.6207822376352452223455184457816472122518527279025978;
    constexpr double kD[11] = {2.48574089138753565546e-5,  1.05142378581721974210,   -3.45687097222016235469,  4.51227709466894823700,
                               -2.98285225323576655721,    1.05639711577126713077,   -1.95428773191645869583e-1, 1.70970543404441224307e-2,
                               -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9};
    double series = kD[0];
    for (int i = 1; i < 11; ++i) series += kD[i] / (x + (double)i - 1.0);
    return std::log(series) + kLn2SqrtEOverPi + (x - 0.5) * std::log((x - 0.5 + kR) / M_E);
}
static double ln_factorial(uint64_t x) {
    static double cache[171];
    static std::once_flag once;
    std::call_once(once, [] {
        cache[0] = 1.0;
        for (int i = 1; i <= 170; ++i) cache[i] = cache[i - 1] * (double)i;
    });
    return x <= 170 ? std::log(cache[x]) : lanczos_ln_gamma((double)x + 1.0);
}
float ln_binomial_f32(uint32_t n, uint32_t k) {
    if (k > n) return -INFINITY;
    return (float)(ln_factorial(n) - ln_factorial(k) - ln_factorial((uint64_t)n - k));
}

// Rust's Display for f32: shortest digits that round-trip, positional notation, no exponent.
// Rust's `{}` for floats: the shortest digits that round-trip, positional notation, never an exponent.  Writes at dst and returns the
// end; at most 64 bytes for f32 ("-0." + 44 zeros + 9 digits), 352 for f64 ("-0." + 323 zeros + 17 digits).
template <typename T>
char* format_float_into(char* dst, T v) {
    auto put = [&dst](const char* lit) { while (*lit) *dst++ = *lit++; return dst; };
    if (std::isnan(v)) return put("NaN");
    if (std::isinf(v)) return put(v < 0 ? "-inf" : "inf");
    if (v == 0) return put(std::signbit(v) ? "-0" : "0");
    char buf[48];
    const auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);   // d[.ddd]e[+-]xx
    const char* p = buf;
    if (*p == '-') { *dst++ = '-'; ++p; }
    char digits[24];
    int nd = 0;
    for (; p < res.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    ++p;                                   // 'e'
    const bool neg = *p == '-';
    ++p;                                   // sign (always present)
    int exp10 = 0;
    for (; p < res.ptr; ++p) exp10 = exp10 * 10 + (*p - '0');
    if (!neg) {
        if (nd <= exp10 + 1) {
            memcpy(dst, digits, (size_t)nd); dst += nd;
            memset(dst, '0', (size_t)(exp10 + 1 - nd)); dst += exp10 + 1 - nd;
        } else {
            memcpy(dst, digits, (size_t)exp10 + 1); dst += exp10 + 1;
            *dst++ = '.';
            memcpy(dst, digits + exp10 + 1, (size_t)(nd - exp10 - 1)); dst += nd - exp10 - 1;
        }
    } else {
        *dst++ = '0'; *dst++ = '.';
        memset(dst, '0', (size_t)(exp10 - 1)); dst += exp10 - 1;
        memcpy(dst, digits, (size_t)nd); dst += nd;
    }
    return dst;
}
char* format_f32_into(char* dst, float v) { return format_float_into(dst, v); }
constexpr size_t kMaxF32 = 64, kMaxF64 = 352;

// Formats rows [0, n) in parallel (blocks of 512 rows, a wave of blocks at a time) and writes them to `out` in order.
// row_max(i) bounds the bytes of row i; emit(i, w) writes the row at w and returns its end.  Each block goes through a local pointer
// into its own cache-line-aligned buffer (neighbouring std::string headers would false-share on every append).
template <typename Max, typename Emit>
int32_t write_rows(FILE* out, uint64_t n, size_t typical_row, Max row_max, Emit emit) {
    struct alignas(128) Buf {
        std::unique_ptr<char[]> p;
        size_t cap = 0, len = 0;
    };
    const uint64_t block = 512;
    const unsigned nthreads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    const uint64_t n_blocks = (n + block - 1) / block, wave = (uint64_t)nthreads * 4;
    std::vector<Buf> text(std::min(wave, std::max<uint64_t>(n_blocks, 1)));
    for (uint64_t b0 = 0; b0 < n_blocks; b0 += wave) {
        const uint64_t nb = std::min(wave, n_blocks - b0);
        std::atomic<uint64_t> next{0};
        auto work = [&] {
            for (uint64_t j; (j = next.fetch_add(1)) < nb;) {
                Buf& buf = text[j];
                char* base = buf.p.get();
                size_t cap = buf.cap, len = 0;
                for (uint64_t i = (b0 + j) * block; i < std::min(n, (b0 + j + 1) * block); ++i) {
                    const size_t need = row_max(i);
                    if (len + need > cap) {
                        const size_t ncap = std::max(cap * 2, len + need + typical_row * (size_t)block);
                        std::unique_ptr<char[]> np(new char[ncap]);
                        if (len) memcpy(np.get(), base, len);
                        buf.p.swap(np);
                        base = buf.p.get();
                        cap = ncap;
                    }
                    len = (size_t)(emit(i, base + len) - base);
                }
                buf.cap = cap;
                buf.len = len;
            }
        };
        const size_t workers = (size_t)std::min<uint64_t>(nthreads, nb);
        parallel_for((unsigned)workers, workers, [&](size_t) { work(); });
        for (uint64_t j = 0; j < nb; ++j)
            if (fwrite(text[j].p.get(), 1, text[j].len, out) != text[j].len) { set_thread_error("write failed"); return SOUP_ERR_IO; }
    }
    return SOUP_OK;
}
std::string format_f32(float v) {
    char b[64];
    return std::string(b, format_f32_into(b, v));
}
std::string format_f32_w8(float v) {
    std::string s = format_f32(v);
    if (s.size() < 8) s.insert(0, 8 - s.size(), ' ');
    return s;
}

// bad_cluster_detection S:150-187 in two halves, so that cell-sharded processes can add their counts up in between
void count_cluster_cells(const float* best_lp, uint64_t n_cells, int32_t k, uint64_t* counts) {
    for (uint64_t i = 0; i < n_cells; ++i) {             // max_by(total_cmp) = LAST maximum (S:159)
        const float* row = best_lp + i * (uint64_t)k;
        int32_t best = 0;
        for (int32_t c = 1; c < k; ++c)
            if (total_key(row[c]) >= total_key(row[best])) best = c;
        counts[(size_t)best] += 1;
    }
}
int32_t bad_clusters_from_counts(int32_t run, int32_t k, const uint64_t* counts, int32_t* out_idx, void* log_file) {
    FILE* log = (FILE*)log_file;
    if (k < 16 || run == 0) return 0;
    std::vector<std::pair<int32_t, uint64_t>> assigned((size_t)k);
    for (int32_t c = 0; c < k; ++c) assigned[(size_t)c] = {c, counts[c]};
    std::stable_sort(assigned.begin(), assigned.end(),
                     [](const std::pair<int32_t, uint64_t>& a, const std::pair<int32_t, uint64_t>& b) { return a.second < b.second; });
    size_t run_value = (size_t)(3 - run);
    size_t low = run_value * ((size_t)k / 16), high = (16 - run_value) * ((size_t)k / 16);
    int32_t n_out = 0;
    if (log) fprintf(log, "Cluster Analysis\n");
    for (size_t index = 0; index < assigned.size(); ++index) {
        int32_t cluster = assigned[index].first;
        bool outlier = index < low || index > high;
        if (log) fprintf(log, "%zu:\tcluster\t%d\tloss\t%llu%s\n", index, cluster, (unsigned long long)assigned[index].second, outlier ? " outlier" : "");
        if (outlier && std::find(out_idx, out_idx + n_out, cluster) == out_idx + n_out) out_idx[n_out++] = cluster;
    }
    return n_out;
}

}  // namespace soup

using namespace soup;

extern "C" {

int32_t soup_version(void) { return 3000; }
void soup_free(void* p) { free(p); }

int32_t soup_chacha_words(const uint8_t seed[32], int32_t rounds, uint64_t word_offset, uint32_t* out, uint64_t n) {
    if (!seed || !out || (rounds != 8 && rounds != 12 && rounds != 20)) { set_thread_error("soup_chacha_words: bad argument"); return SOUP_ERR_INVALID; }
    uint32_t key[8], buf[16];
    seed_to_key(seed, key);
    uint64_t blk = word_offset / 16;
    int idx = (int)(word_offset % 16);
    chacha_block(key, blk++, rounds, buf);
    for (uint64_t i = 0; i < n; ++i) {
        if (idx == 16) { chacha_block(key, blk++, rounds, buf); idx = 0; }
        out[i] = buf[idx++];
    }
    return SOUP_OK;
}

int32_t soup_thread_seeds(uint8_t seed, uint32_t threads, uint32_t run, uint8_t* out) {
    if (!out || run > 2) { set_thread_error("soup_thread_seeds: bad argument"); return SOUP_ERR_INVALID; }
    uint8_t s[32];
    memset(s, seed, 32);                                   // S:61
    StdRng rng = StdRng::from_seed(s);                     // S:62
    for (uint32_t r = 0; r <= run; ++r)                    // each run draws a fresh set (S:77-80)
        for (uint32_t t = 0; t < threads; ++t)
            for (int i = 0; i < 32; ++i) {                 // new_seed S:855-861
                uint8_t b = rng.gen_u8();
                if (r == run) out[(size_t)t * 32 + i] = b;
            }
    return SOUP_OK;
}

// ---- vartrix text writer: the inverse of soup_mtx_load, for generators and benchmarks (coordinates 1-based, `locus cell count` per line)
int32_t soup_mtx_write_text(const char* path, uint64_t n_loci, uint64_t n_cells, uint64_t nnz, const int64_t* locus, const int64_t* cell,
                            const uint32_t* count, const char* comment) try {
    if (!path || (nnz && (!locus || !cell || !count))) { set_thread_error("soup_mtx_write_text: bad argument"); return SOUP_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { set_thread_error("cannot create %s", path); return SOUP_ERR_IO; }
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%% %s\n%llu %llu %llu\n", comment ? comment : "written by libsoupgpu",
            (unsigned long long)n_loci, (unsigned long long)n_cells, (unsigned long long)nnz);
    const uint64_t block = 1u << 18, n_blocks = (nnz + block - 1) / block;
    const unsigned T = host_threads();
    std::vector<std::string> bufs(T);
    auto put = [](char*& o, uint64_t v) {
        char tmp[24];
        int n = 0;
        do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (n) *o++ = tmp[--n];
    };
    bool ok = true;
    for (uint64_t b0 = 0; b0 < n_blocks && ok; b0 += T) {           // waves of T blocks: formatted in parallel, written in order
        const size_t wave = (size_t)std::min<uint64_t>(T, n_blocks - b0);
        parallel_for(T, wave, [&](size_t i) {
            const uint64_t s0 = (b0 + i) * block, s1 = std::min(nnz, s0 + block);
            std::string& buf = bufs[i];
            buf.resize((size_t)(s1 - s0) * 64);
            char* o = &buf[0];
            for (uint64_t j = s0; j < s1; ++j) {
                put(o, (uint64_t)locus[j] + 1); *o++ = ' ';
                put(o, (uint64_t)cell[j] + 1); *o++ = ' ';
                put(o, count[j]); *o++ = '\n';
            }
            buf.resize((size_t)(o - &buf[0]));
        });
        for (size_t i = 0; i < wave && ok; ++i) ok = fwrite(bufs[i].data(), 1, bufs[i].size(), f) == bufs[i].size();
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) { set_thread_error("write to %s failed", path); return SOUP_ERR_IO; }
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_mtx_write_text")

// ---- troublet host side: load_clusters T:414-449 and the TSV writer T:36-41, T:233-243
static std::string format_f64(double v) {   // Rust `{}` for f64
    char b[kMaxF64];
    return std::string(b, format_float_into(b, v));
}

int32_t soup_clusters_load(const char* path, char** barcode_blob, double** losses, uint64_t* n_cells, int32_t* k, int32_t* n_clusters) try {
    if (!path || !barcode_blob || !losses || !n_cells || !k || !n_clusters) { set_thread_error("soup_clusters_load: bad argument"); return SOUP_ERR_INVALID; }
    // the file is mapped (or, for something that is not a regular file, read) — a 100k x 64 clusters file is 70 MB
    struct Text {
        const char* p = nullptr;
        size_t n = 0;
        bool mapped = false;
        std::string owned;
        ~Text() { if (mapped) munmap((void*)p, n); }
        const char* data() const { return p; }
        size_t size() const { return n; }
    } content;
    {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) { set_thread_error("Unable to open file %s", path); return SOUP_ERR_IO; }
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) { content.p = (const char*)m; content.n = (size_t)st.st_size; content.mapped = true; }
        }
        if (!content.mapped) {
            char buf[1 << 16];
            ssize_t got;
            while ((got = ::read(fd, buf, sizeof(buf))) > 0) content.owned.append(buf, (size_t)got);
            content.p = content.owned.data();
            content.n = content.owned.size();
        }
        ::close(fd);
    }
    // lines as BufRead::lines() yields them, each trimmed like str::trim (T:422)
    struct Line { const char* b; const char* e; };
    std::vector<Line> lines;
    for (size_t pos = 0; pos < content.size();) {
        const char* b = content.data() + pos;
        const char* e = (const char*)memchr(b, '\n', content.size() - pos);
        if (!e) e = content.data() + content.size();
        pos = (size_t)(e - content.data()) + 1;
        auto blank = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n'; };
        while (b < e && blank(*b)) ++b;
        while (e > b && blank(e[-1])) --e;
        lines.push_back({b, e});
    }
    const uint64_t cells = lines.size();
    int32_t cols = 0;
    if (cells) cols = (int32_t)std::count(lines[0].b, lines[0].e, '\t') - 1;    // tokens = tabs + 1, minus barcode and cluster
    const size_t n_vals = cols > 0 ? (size_t)cells * (size_t)cols : 0;
    struct Free { void operator()(void* q) const { free(q); } };
    std::unique_ptr<double[], Free> vals((double*)malloc(std::max<size_t>(n_vals, 1) * sizeof(double)));   // handed to the caller at the end
    if (!vals) { set_thread_error("out of memory"); return SOUP_ERR_NOMEM; }
    std::vector<Line> names(cells);
    // rows are parsed in parallel; the error reported is the one of the earliest bad line, as a sequential reader would hit it
    std::mutex err_mu;
    uint64_t err_line = UINT64_MAX;
    std::string err_msg;
    auto fail = [&](uint64_t line, const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        std::string m = vformat(fmt, ap);
        va_end(ap);
        std::lock_guard<std::mutex> g(err_mu);
        if (line < err_line) { err_line = line; err_msg.swap(m); }
    };
    const uint64_t step = 256, n_blocks = (cells + step - 1) / step;
    std::vector<uint64_t> block_max(n_blocks, 0);
    parallel_for(host_threads(), (size_t)n_blocks, [&](size_t blk) {
        std::string tmp;
        for (uint64_t i = blk * step; i < std::min(cells, (blk + 1) * step); ++i) {
            const char* p = lines[i].b;
            const char* const end = lines[i].e;
            auto token = [&](const char*& tb, const char*& te) {   // str::split("\t"): empty tokens count, the last one ends the line
                if (p > end) return false;
                tb = p;
                te = (const char*)memchr(p, '\t', (size_t)(end - p));
                if (!te) te = end;
                p = te + 1;
                return true;
            };
            const char *tb, *te;
            token(tb, te);
            names[i] = {tb, te};
            if (!token(tb, te)) { fail(i, "clusters file line %llu: index out of bounds: the len is 1 but the index is 1", (unsigned long long)i + 1); return; }
            {   // parse::<usize>(): optional '+', then digits only
                const char* q = tb;
                if (q < te && *q == '+') ++q;
                uint64_t c1 = 0;
                bool ok = q < te;
                for (; ok && q < te; ++q) {
                    const unsigned d = (unsigned)(unsigned char)*q - '0';
                    if (d > 9 || c1 > (UINT64_MAX - d) / 10) ok = false; else c1 = c1 * 10 + d;
                }
                if (!ok) { fail(i, "clusters file line %llu: cannot parse the cluster column", (unsigned long long)i + 1); return; }
                block_max[blk] = std::max(block_max[blk], c1);
            }
            int32_t here = 0;
            while (token(tb, te)) {
                if (here >= cols) { ++here; continue; }
                double v = 0;
                // digits, '.', exponent and '-' only: from_chars (correctly rounded, like strtod); anything else — '+', inf, nan, blanks —
                // takes the strtod route that has always defined what this reader accepts
                bool plain = tb < te;
                for (const char* q = tb; plain && q < te; ++q) plain = (*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '-';
                bool ok;
                if (plain) {
                    const auto r = std::from_chars(tb, te, v, std::chars_format::general);
                    ok = r.ec == std::errc() && r.ptr == te;
                    if (r.ec == std::errc::result_out_of_range) plain = false;   // strtod decides (inf / 0 / subnormal)
                }
                if (!plain) {
                    tmp.assign(tb, te);
                    char* endp = nullptr;
                    v = strtod(tmp.c_str(), &endp);
                    ok = !tmp.empty() && *endp == 0;
                }
                if (!ok) { fail(i, "clusters file line %llu: cannot parse '%s' as f64", (unsigned long long)i + 1, std::string(tb, te).c_str()); return; }
                vals[(size_t)i * (size_t)cols + (size_t)here] = v;
                ++here;
            }
            if (here != cols) { fail(i, "clusters file line %llu: %d log-probability columns, expected %d", (unsigned long long)i + 1, here, cols < 0 ? 0 : cols); return; }
        }
    });
    if (err_line != UINT64_MAX) { set_thread_error("%s", err_msg.c_str()); return SOUP_ERR_INVALID; }
    uint64_t max_cluster = 0;
    for (uint64_t m : block_max) max_cluster = std::max(max_cluster, m);
    std::string blob;
    for (const Line& nm : names) { blob.append(nm.b, (size_t)(nm.e - nm.b)); blob.push_back('\0'); }
    *barcode_blob = (char*)malloc(std::max<size_t>(blob.size(), 1));
    if (!*barcode_blob) { set_thread_error("out of memory"); return SOUP_ERR_NOMEM; }
    memcpy(*barcode_blob, blob.data(), blob.size());
    *losses = vals.release();
    *n_cells = cells; *k = cols < 0 ? 0 : cols; *n_clusters = (int32_t)max_cluster + 1;
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_clusters_load")

int32_t soup_troublet_write(void* out_file, const char* barcode_blob, uint64_t n_cells, int32_t k, const soup_troublet_result* r) try {
    FILE* f = (FILE*)out_file;
    if (!f || !barcode_blob || !r || k <= 0) { set_thread_error("soup_troublet_write: bad argument"); return SOUP_ERR_INVALID; }
    static const char* NAME[4] = {"singlet", "unassigned", "doublet", "unassigned"};
    std::string head = "barcode\tstatus\tassignment\tsinglet_posterior\tdoublet_posterior\tlog_prob_singleton\tlog_prob_doublet\t";
    for (int32_t c = 0; c < k; ++c) { head += "cluster" + std::to_string(c); head += c + 1 == k ? "\n" : "\t"; }
    if (fwrite(head.data(), 1, head.size(), f) != head.size()) { set_thread_error("write failed"); return SOUP_ERR_IO; }
    std::vector<const char*> bc(n_cells);
    size_t longest = 0;
    {
        const char* q = barcode_blob;
        for (uint64_t i = 0; i < n_cells; ++i) { bc[i] = q; const size_t l = strlen(q); longest = std::max(longest, l); q += l + 1; }
    }
    auto field = [](char* w, double v) { w = format_float_into(w, v); *w++ = '\t'; return w; };
    const int32_t rc = write_rows(
        f, n_cells, longest + 48 + ((size_t)k + 4) * 20,
        [&](uint64_t cell) { return strlen(bc[cell]) + 64 + ((size_t)k + 4) * (kMaxF64 + 1); },
        [&](uint64_t cell, char* w) {
            const size_t bl = strlen(bc[cell]);
            memcpy(w, bc[cell], bl); w += bl;
            *w++ = '\t';
            const int st = r->status[cell];
            for (const char* nm = NAME[st & 3]; *nm;) *w++ = *nm++;
            *w++ = '\t';
            if (st == SOUP_TROUBLET_SINGLET || st == SOUP_TROUBLET_UNASSIGNED) w = std::to_chars(w, w + 12, r->best_singlet[cell]).ptr;
            else { w = std::to_chars(w, w + 12, r->doublet1[cell]).ptr; *w++ = '/'; w = std::to_chars(w, w + 12, r->doublet2[cell]).ptr; }
            *w++ = '\t';
            w = field(w, r->singlet_posterior[cell]);
            w = field(w, r->doublet_posterior[cell]);
            w = field(w, r->log_prob_singleton[cell]);
            w = field(w, r->log_prob_doublet[cell]);
            for (int32_t c = 0; c < k; ++c) {
                w = format_float_into(w, r->cluster_logprobs[cell * (uint64_t)k + c]);
                *w++ = c + 1 < k ? '\t' : '\n';
            }
            return w;
        });
    if (rc != SOUP_OK) return rc;
    return ferror(f) ? SOUP_ERR_IO : SOUP_OK;
}
SOUP_CATCH_RETURN("soup_troublet_write")

int32_t soup_format_f32(float v, int32_t width8, char* buf, int32_t buflen) try {
    std::string s = width8 ? format_f32_w8(v) : format_f32(v);
    if (!buf || (int32_t)s.size() + 1 > buflen) { set_thread_error("soup_format_f32: buffer too small"); return SOUP_ERR_INVALID; }
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int32_t)s.size();
}
SOUP_CATCH_RETURN("soup_format_f32")

float soup_ln_binomial_f32(uint32_t n, uint32_t k) { return ln_binomial_f32(n, k); }

// load_barcodes S:707-718 / reader S:500-511
int32_t soup_barcodes_load(const char* path, char** blob, uint64_t* n_barcodes) try {
    if (!path || !blob || !n_barcodes) { set_thread_error("soup_barcodes_load: bad argument"); return SOUP_ERR_INVALID; }
    std::string content;
    size_t plen = strlen(path);
    if (plen >= 3 && !strcmp(path + plen - 3, ".gz")) {
        gzFile f = gzopen(path, "rb");
        if (!f) { set_thread_error("couldn't open file %s", path); return SOUP_ERR_IO; }
        char buf[1 << 16];
        int n;
        while ((n = gzread(f, buf, sizeof(buf))) > 0) content.append(buf, (size_t)n);
        gzclose(f);
    } else {
        std::ifstream f(path, std::ios::binary);
        if (!f) { set_thread_error("couldn't open file %s", path); return SOUP_ERR_IO; }
        content.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    }
    std::string packed;
    uint64_t n = 0;
    size_t i = 0;
    while (i < content.size()) {   // BufRead::lines(): split on '\n', strip one trailing '\r'
        size_t j = content.find('\n', i);
        if (j == std::string::npos) j = content.size();
        size_t e = j;
        if (e > i && content[e - 1] == '\r') --e;
        packed.append(content, i, e - i);
        packed.push_back('\0');
        ++n;
        i = j + 1;
    }
    char* out = (char*)malloc(packed.size() + 1);
    if (!out) { set_thread_error("out of memory"); return SOUP_ERR_NOMEM; }
    memcpy(out, packed.data(), packed.size());
    out[packed.size()] = 0;
    *blob = out;
    *n_barcodes = n;
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_barcodes_load")

// init_cluster_centers_known_genotypes S:514-542: theta[k][l] = 0.5, then for every VCF record whose index is a used
// locus and every named sample: (min(hap0,1) + min(hap1,1)) / 2 clamped to [0.01, 0.99]; a GT starting with '.' keeps 0.5.
int32_t soup_known_genotypes_load(const char* vcf_path, const char* sample_names, int32_t n_samples, int32_t k,
                                  const uint64_t* index_to_locus, uint64_t n_loci, float* theta) try {
    if (!vcf_path || !theta || k <= 0 || n_samples < 0 || (n_samples && !sample_names) || (n_loci && !index_to_locus)) {
        set_thread_error("soup_known_genotypes_load: bad argument");
        return SOUP_ERR_INVALID;
    }
    if (n_samples == 0) { set_thread_error("currently requiring known_genotypes_sample_names if known_genotypes set"); return SOUP_ERR_INVALID; }   // S:537
    if (n_samples > k) { set_thread_error("known_genotypes: more sample names than clusters (the reference indexes out of bounds)"); return SOUP_ERR_INVALID; }
    std::vector<std::string> names;
    for (const char* q = sample_names; (int32_t)names.size() < n_samples; q += strlen(q) + 1) names.emplace_back(q);
    gzFile f = gzopen(vcf_path, "rb");   // plain text is read transparently; the reference's reader() keys on ".gz" (S:506)
    if (!f) { set_thread_error("couldn't open file %s", vcf_path); return SOUP_ERR_IO; }
    std::string text;
    {
        char buf[1 << 16];
        int n;
        while ((n = gzread(f, buf, sizeof(buf))) > 0) text.append(buf, (size_t)n);
        gzclose(f);
    }
    for (size_t i = 0; i < (size_t)k * n_loci; ++i) theta[i] = 0.5f;
    std::vector<int> sample_col(names.size(), -1);
    uint64_t record = 0, next_used = 0;          // index_to_locus is ascending: walk it alongside the records
    size_t pos = 0;
    auto field = [](const std::string& line, size_t idx, size_t* b, size_t* e) {   // idx-th tab-separated field
        size_t s0 = 0;
        for (size_t i = 0; i < idx; ++i) {
            size_t t = line.find('\t', s0);
            if (t == std::string::npos) return false;
            s0 = t + 1;
        }
        size_t t = line.find('\t', s0);
        *b = s0;
        *e = t == std::string::npos ? line.size() : t;
        return true;
    };
    while (pos < text.size()) {
        size_t e = text.find('\n', pos);
        if (e == std::string::npos) e = text.size();
        std::string line(text, pos, e - pos);
        pos = e + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '#') {
            if (line.compare(0, 6, "#CHROM") == 0) {
                size_t b, en;
                for (size_t c = 9; field(line, c, &b, &en); ++c)
                    for (size_t s = 0; s < names.size(); ++s)
                        if (line.compare(b, en - b, names[s]) == 0) sample_col[s] = (int)c;
            }
            continue;
        }
        while (next_used < n_loci && index_to_locus[next_used] < record) ++next_used;
        if (next_used < n_loci && index_to_locus[next_used] == record) {
            size_t fb, fe;
            if (!field(line, 8, &fb, &fe)) { set_thread_error("known_genotypes: record %llu has no FORMAT column", (unsigned long long)record); return SOUP_ERR_INVALID; }
            int gt_pos = -1, cur = 0;
            for (size_t i = fb; i <= fe; ++cur) {
                size_t c = line.find(':', i);
                if (c == std::string::npos || c > fe) c = fe;
                if (line.compare(i, c - i, "GT") == 0) gt_pos = cur;
                i = c + 1;
            }
            for (size_t s = 0; s < names.size(); ++s) {
                size_t b, en;
                if (sample_col[s] < 0 || gt_pos < 0 || !field(line, (size_t)sample_col[s], &b, &en)) {
                    set_thread_error("known_genotypes: sample %s or its GT is missing at record %llu (the reference panics on the lookup)", names[s].c_str(), (unsigned long long)record);
                    return SOUP_ERR_INVALID;
                }
                size_t gb = b;
                for (int i = 0; i < gt_pos; ++i) {
                    size_t c = line.find(':', gb);
                    if (c == std::string::npos || c >= en) { gb = en; break; }
                    gb = c + 1;
                }
                size_t ge = line.find(':', gb);
                if (ge == std::string::npos || ge > en) ge = en;
                if (gb >= ge) { set_thread_error("known_genotypes: empty GT at record %llu", (unsigned long long)record); return SOUP_ERR_INVALID; }
                if (line[gb] == '.') continue;                                             // S:532
                auto hap = [&](size_t i, uint32_t* out) {
                    if (gb + i >= ge || line[gb + i] < '0' || line[gb + i] > '9') return false;
                    *out = std::min<uint32_t>((uint32_t)(line[gb + i] - '0'), 1u);        // parse::<u32>().min(1)  S:533-534
                    return true;
                };
                uint32_t h0, h1;
                if (!hap(0, &h0) || !hap(2, &h1)) { set_thread_error("known_genotypes: cannot parse GT '%s' at record %llu", line.substr(gb, ge - gb).c_str(), (unsigned long long)record); return SOUP_ERR_INVALID; }
                theta[s * n_loci + next_used] = fmaxf(fminf((float)(h0 + h1) / 2.0f, 0.99f), 0.01f);   // S:535
            }
        }
        ++record;
    }
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_known_genotypes_load")

// S:133-147: zip(barcodes, best_log_probabilities), FIRST max (strict >), Rust {} floats.
int32_t soup_write_clusters(void* out_file, const char* barcode_blob, uint64_t n_barcodes, const float* lp,
                            uint64_t n_cells, int32_t k) try {
    FILE* out = (FILE*)out_file;
    if (!out || (!barcode_blob && n_barcodes) || (!lp && n_cells) || k <= 0) { set_thread_error("soup_write_clusters: bad argument"); return SOUP_ERR_INVALID; }
    const uint64_t n = std::min(n_barcodes, n_cells);   // zip() stops at the shorter (S:133)
    std::vector<const char*> bc(n);
    {
        const char* q = barcode_blob;
        for (uint64_t i = 0; i < n; ++i) { bc[i] = q; q += strlen(q) + 1; }
    }
    size_t longest = 0;
    for (uint64_t i = 0; i < n; ++i) longest = std::max(longest, strlen(bc[i]));
    return write_rows(
        out, n, longest + 16 + (size_t)k * 14,
        [&](uint64_t i) { return strlen(bc[i]) + 16 + (size_t)k * (kMaxF32 + 1); },
        [&](uint64_t i, char* w) {
            const float* row = lp + i * (uint64_t)k;
            int best = 0;
            float best_lp = -INFINITY;                              // first maximum (S:134-141)
            for (int c = 0; c < k; ++c)
                if (row[c] > best_lp) { best = c; best_lp = row[c]; }
            const size_t bl = strlen(bc[i]);
            memcpy(w, bc[i], bl); w += bl;
            *w++ = '\t';
            w = std::to_chars(w, w + 12, best).ptr;
            *w++ = '\t';
            for (int c = 0; c < k; ++c) {
                w = format_f32_into(w, row[c]);
                *w++ = c < k - 1 ? '\t' : '\n';
            }
            return w;
        });
}
SOUP_CATCH_RETURN("soup_write_clusters")

// S:150-187
int32_t soup_bad_cluster_detection(int32_t run, int32_t k, const float* best_lp, uint64_t n_cells, int32_t* out_idx,
                                   void* log_file) try {
    if (k <= 0 || run < 0 || !out_idx) { set_thread_error("soup_bad_cluster_detection: bad argument"); return SOUP_ERR_INVALID; }
    if (k < 16 || run == 0) return 0;
    std::vector<uint64_t> counts((size_t)k, 0);
    soup::count_cluster_cells(best_lp, n_cells, k, counts.data());
    return soup::bad_clusters_from_counts(run, k, counts.data(), out_idx, log_file);
}
SOUP_CATCH_RETURN("soup_bad_cluster_detection")

}  // extern "C"
