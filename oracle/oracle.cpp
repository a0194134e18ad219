// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp header).  PARITY UNPINNED.
// CPU restatement of /root/reference/souporcell/src/main.rs ("S:").
// Build: g++ -std=c++17 -O2 -ffp-contract=off (no fast-math); plain libm logf/expf,
// the same calls Rust's f32::ln / f32::exp lower to on linux-gnu.
#include "oracle.hpp"

#include <algorithm>
#include <array>
#include <charconv>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <mutex>
#include <stdexcept>
#include <atomic>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <zlib.h>

namespace orc {

static const float NEG_INF = -std::numeric_limits<float>::infinity();

// ------------------------------------------------------------------ ChaCha
static inline uint32_t rotl(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
#define ORC_QR(a, b, c, d)            \
    a += b; d ^= a; d = rotl(d, 16);  \
    c += d; b ^= c; b = rotl(b, 12);  \
    a += b; d ^= a; d = rotl(d, 8);   \
    c += d; b ^= c; b = rotl(b, 7);

void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
    uint32_t x[16];
    memcpy(x, s, sizeof(x));
    for (int r = 0; r < rounds; r += 2) {
        ORC_QR(x[0], x[4], x[8], x[12]) ORC_QR(x[1], x[5], x[9], x[13])
        ORC_QR(x[2], x[6], x[10], x[14]) ORC_QR(x[3], x[7], x[11], x[15])
        ORC_QR(x[0], x[5], x[10], x[15]) ORC_QR(x[1], x[6], x[11], x[12])
        ORC_QR(x[2], x[7], x[8], x[13]) ORC_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + s[i];
}

ChaChaRng ChaChaRng::from_seed(const uint8_t seed[32], int rounds) {
    ChaChaRng r;
    for (int i = 0; i < 8; ++i)
        r.key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) |
                   ((uint32_t)seed[4 * i + 3] << 24);
    r.rounds = rounds;
    r.block = 0;
    r.idx = 16;
    return r;
}
uint32_t ChaChaRng::next_u32() {
    if (idx >= 16) {
        chacha_block(key, block, rounds, buf);
        ++block;
        idx = 0;
    }
    return buf[idx++];
}
void ChaChaRng::seek_word(uint64_t word) {
    block = word / 16;
    chacha_block(key, block, rounds, buf);
    ++block;
    idx = (int)(word % 16);
}

// ------------------------------------------------------------------ statrs 0.11.0 ln_binomial
// function::factorial: FCACHE[i] = FCACHE[i-1] * i (f64, i <= 170); ln_factorial(x) =
// FCACHE[x].ln() for x <= 170 else ln_gamma(x+1); ln_binomial(n,k) = lf(n) - lf(k) - lf(n-k).
// function::gamma::ln_gamma (Lanczos approximation, r = 10.900511, 11 coefficients), the x >= 0.5 branch: the only caller
// passes n + 1 with n > 170.  Restated from the published statrs source (not under /root/reference): unpinned.
static double statrs_ln_gamma(double x) {
    static const double r = 10.900511;
    static const double dk[11] = {2.48574089138753565546e-5, 1.05142378581721974210,  -3.45687097222016235469,
                                  4.51227709466894823700,    -2.98285225323576655721, 1.05639711577126713077,
                                  -1.95428773191645869583e-1, 1.70970543404441224307e-2, -5.71926117404305781283e-4,
                                  4.63399473359905636708e-6, -2.71994908488607703910e-9};
    static const double ln_2_sqrt_e_over_pi = 0.6207822376352452223455184457816472122518527279025978;
    double s = dk[0];
    for (int i = 1; i < 11; ++i) s += dk[i] / (x + (double)i - 1.0);
    return std::log(s) + ln_2_sqrt_e_over_pi + (x - 0.5) * std::log((x - 0.5 + r) / M_E);
}
static double ln_factorial(uint64_t x) {
    static double fcache[171];
    static std::once_flag once;
    std::call_once(once, [] {
        fcache[0] = 1.0;
        for (int i = 1; i <= 170; ++i) fcache[i] = fcache[i - 1] * (double)i;
    });
    if (x <= 170) return std::log(fcache[x]);
    return statrs_ln_gamma((double)x + 1.0);
}
double ln_binomial(uint64_t n, uint64_t k) {
    if (k > n) return -std::numeric_limits<double>::infinity();
    return ln_factorial(n) - ln_factorial(k) - ln_factorial(n - k);
}

// ------------------------------------------------------------------ loader  S:601-681
static bool parse_unsigned(const std::string& tok, uint64_t* out) {
    size_t i = 0;
    if (!tok.empty() && tok[0] == '+') i = 1;   // Rust's integer FromStr accepts a leading '+'
    if (i >= tok.size()) return false;
    uint64_t v = 0;
    for (; i < tok.size(); ++i) {
        if (tok[i] < '0' || tok[i] > '9') return false;
        uint64_t nv = v * 10 + (uint64_t)(tok[i] - '0');
        if (nv < v) return false;
        v = nv;
    }
    *out = v;
    return true;
}
static std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> t;
    size_t i = 0, n = s.size();
    while (i < n) {
        while (i < n && isspace((unsigned char)s[i])) ++i;
        size_t j = i;
        while (j < n && !isspace((unsigned char)s[j])) ++j;
        if (j > i) t.emplace_back(s, i, j - i);
        i = j;
    }
    return t;
}
static uint64_t must_parse(const std::vector<std::string>& toks, size_t i, uint64_t maxv) {
    uint64_t v;
    if (i >= toks.size() || !parse_unsigned(toks[i], &v) || v > maxv)
        throw std::runtime_error("mtx parse error (reference: unwrap on ParseIntError / index out of bounds)");
    return v;
}

static void push_cell_entry(CellData& cd, size_t locus_index, uint32_t ref_c, uint32_t alt_c) {
    // S:665-671
    cd.alt_counts.push_back(alt_c);
    cd.ref_counts.push_back(ref_c);
    cd.loci.push_back(locus_index);
    cd.allele_fractions.push_back((float)alt_c / (float)(ref_c + alt_c));
    cd.log_binomial_coefficient.push_back((float)ln_binomial((uint64_t)alt_c + ref_c, alt_c));
    cd.total_alleles += (float)(ref_c + alt_c);
}

LoadedData load_cell_data(const Params& params, FILE* log) {
    std::ifstream alt_reader(params.alt_mtx);
    if (!alt_reader) throw std::runtime_error("cannot open alt mtx file");
    std::ifstream ref_reader(params.ref_mtx);
    if (!ref_reader) throw std::runtime_error("cannot open ref mtx file");
    size_t line_number = 0, total_loci = 0, total_cells = 0;
    std::unordered_set<size_t> all_loci;
    std::unordered_map<size_t, std::array<uint32_t, 2>> locus_cell_counts, locus_umi_counts;
    std::unordered_map<size_t, std::unordered_map<size_t, std::array<uint32_t, 2>>> locus_counts;
    std::string alt_line, ref_line;
    while (std::getline(alt_reader, alt_line) && std::getline(ref_reader, ref_line)) {   // izip!  S:616
        if (line_number > 2) {
            auto alt_tokens = split_ws(alt_line);
            auto ref_tokens = split_ws(ref_line);
            size_t locus = (size_t)must_parse(alt_tokens, 0, UINT64_MAX) - 1;     // S:622
            all_loci.insert(locus);
            size_t cell = (size_t)must_parse(alt_tokens, 1, UINT64_MAX) - 1;      // S:624
            uint32_t ref_count = (uint32_t)must_parse(ref_tokens, 2, UINT32_MAX); // S:625
            uint32_t alt_count = (uint32_t)must_parse(alt_tokens, 2, UINT32_MAX); // S:626
            if (!(locus < total_loci)) throw std::runtime_error("assertion failed: locus < total_loci");
            if (!(cell < total_cells)) throw std::runtime_error("assertion failed: cell < total_cells");
            auto& cc = locus_cell_counts.try_emplace(locus, std::array<uint32_t, 2>{0, 0}).first->second;
            auto& uc = locus_umi_counts.try_emplace(locus, std::array<uint32_t, 2>{0, 0}).first->second;
            if (ref_count > 0) { cc[0] += 1; uc[0] += ref_count; }                 // S:631
            if (alt_count > 0) { cc[1] += 1; uc[1] += alt_count; }                 // S:632
            locus_counts[locus][cell] = {ref_count, alt_count};                   // S:633-634 (last wins)
        } else if (line_number == 2) {
            auto tokens = split_ws(alt_line);
            total_loci = (size_t)must_parse(tokens, 0, UINT64_MAX);
            total_cells = (size_t)must_parse(tokens, 1, UINT64_MAX);
        }
        ++line_number;
    }
    std::vector<size_t> sorted_loci(all_loci.begin(), all_loci.end());
    std::sort(sorted_loci.begin(), sorted_loci.end());                            // S:648
    LoadedData out;
    out.total_cells = total_cells;
    out.cell_data.resize(total_cells);
    out.locus_to_index.assign(total_loci, SIZE_MAX);
    size_t locus_index = 0;
    for (size_t locus : sorted_loci) {
        const auto& cc = locus_cell_counts.at(locus);
        const auto& uc = locus_umi_counts.at(locus);
        if (cc[0] >= params.min_ref && cc[1] >= params.min_alt && uc[0] >= params.min_ref_umis &&
            uc[1] >= params.min_alt_umis) {                                       // S:659
            out.index_to_locus.push_back(locus);
            out.locus_to_index[locus] = locus_index;
            for (const auto& kv : locus_counts.at(locus)) {                       // hash order: irrelevant per cell
                const auto& counts = kv.second;
                if (counts[0] + counts[1] == 0) continue;                         // S:664
                push_cell_entry(out.cell_data[kv.first], locus_index, counts[0], counts[1]);
            }
            ++locus_index;
        }
    }
    out.loci_used = locus_index;
    if (log) fprintf(log, "total loci used %zu\n", out.loci_used);                // S:678
    return out;
}

LoadedData cell_data_from_csr(size_t n_cells, size_t n_loci, const uint64_t* row_ptr, const uint32_t* locus,
                              const uint32_t* alt, const uint32_t* ref) {
    // Test helper: builds the same Vec<CellData> S:663-674 would, from an already
    // filtered CSR (rows = cells, columns ascending = used-locus index).
    LoadedData out;
    out.loci_used = n_loci;
    out.total_cells = n_cells;
    out.cell_data.resize(n_cells);
    out.index_to_locus.resize(n_loci);
    out.locus_to_index.resize(n_loci);
    for (size_t l = 0; l < n_loci; ++l) out.index_to_locus[l] = out.locus_to_index[l] = l;
    for (size_t c = 0; c < n_cells; ++c)
        for (uint64_t j = row_ptr[c]; j < row_ptr[c + 1]; ++j) {
            if (ref[j] + alt[j] == 0) continue;
            push_cell_entry(out.cell_data[c], locus[j], ref[j], alt[j]);
        }
    return out;
}

// S:500-511, S:707-718
static bool ends_with(const std::string& s, const char* suf) {
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
std::vector<std::string> load_barcodes(const Params& p, FILE* log) {
    std::vector<std::string> out;
    std::string content;
    if (ends_with(p.barcodes, ".gz")) {
        gzFile f = gzopen(p.barcodes.c_str(), "rb");
        if (!f) throw std::runtime_error("couldn't open file " + p.barcodes);
        char buf[1 << 16];
        int n;
        while ((n = gzread(f, buf, sizeof(buf))) > 0) content.append(buf, n);
        gzclose(f);
    } else {
        std::ifstream f(p.barcodes, std::ios::binary);
        if (!f) throw std::runtime_error("couldn't open file " + p.barcodes);
        content.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    }
    size_t i = 0;
    while (i < content.size()) {                         // BufRead::lines: strips "\n" or "\r\n"
        size_t j = content.find('\n', i);
        if (j == std::string::npos) j = content.size();
        size_t e = j;
        if (e > i && content[e - 1] == '\r') --e;
        out.emplace_back(content, i, e - i);
        i = j + 1;
    }
    if (out.size() > 200000 && log)
        fprintf(log, "More than 200_000 barcodes in barcodes file? is this the filtered barcode file? (do not use raw "
                     "barcode file)\n");
    return out;
}

// ------------------------------------------------------------------ solver primitives
std::vector<float> binomial_loss_with_min_index(const CellData& c, const std::vector<std::vector<float>>& centers,
                                                float log_prior, size_t* min_index_out) {
    std::vector<float> lp;
    float min_log = std::numeric_limits<float>::lowest();   // f32::MIN
    size_t min_index = 0;
    for (size_t cluster = 0; cluster < centers.size(); ++cluster) {
        const std::vector<float>& center = centers[cluster];
        lp.push_back(log_prior);
        for (size_t j = 0; j < c.loci.size(); ++j) {
            // S:416-418: acc += (lbc + alt*ln(theta)) + ref*ln(1-theta), un-hoisted logs, no contraction
            float t = c.log_binomial_coefficient[j] + (float)c.alt_counts[j] * logf(center[c.loci[j]]);
            t = t + (float)c.ref_counts[j] * logf(1.0f - center[c.loci[j]]);
            lp[cluster] += t;
        }
        if (lp[cluster] > min_log) { min_log = lp[cluster]; min_index = cluster; }
    }
    if (min_index_out) *min_index_out = min_index;
    return lp;
}

float log_sum_exp(const std::vector<float>& p) {
    float max_p = NEG_INF;
    for (float x : p) max_p = (x > max_p || std::isnan(max_p)) ? x : max_p;   // f32::max fold (NaN-ignoring)
    float sum_rst = 0.0f;
    for (float x : p) sum_rst += expf(x - max_p);
    return max_p + logf(sum_rst);
}

std::vector<float> normalize_in_log_with_temp(const std::vector<float>& lp, float temp) {
    std::vector<float> scaled;
    for (float x : lp) scaled.push_back(x / temp);
    float sum = log_sum_exp(scaled);
    std::vector<float> out;
    for (size_t i = 0; i < lp.size(); ++i) out.push_back(expf(scaled[i] - sum));
    return out;
}

void calculate_q_for_current_cell(const std::vector<float>& ll, size_t min_clus, std::vector<float>* q_vec,
                                  float* q_sum) {
    float log_winner = -ll[min_clus];                                  // S:364
    std::vector<float> d;
    for (size_t i = 0; i < ll.size(); ++i) d.push_back(i != min_clus ? 25.0f * (log_winner + ll[i]) : 0.0f);
    float q_denom = log_sum_exp(d);                                    // S:375
    q_vec->clear();
    for (size_t i = 0; i < ll.size(); ++i)                             // S:378
        q_vec->push_back(((2.0f * 25.0f * log_winner) - ((25.0f + 2.0f) * -ll[i])) - (2.0f * q_denom));
    *q_sum = log_sum_exp(*q_vec);
}

static void reset_sums_denoms(size_t loci, std::vector<std::vector<float>>& sums,
                              std::vector<std::vector<float>>& denoms, size_t k) {   // S:456-464
    for (size_t c = 0; c < k; ++c)
        for (size_t i = 0; i < loci; ++i) { sums[c][i] = 1.0f; denoms[c][i] = 2.0f; }
}
static void update_centers_average(std::vector<std::vector<float>>& sums, std::vector<std::vector<float>>& denoms,
                                   const CellData& cell, const std::vector<float>& prob) {   // S:476-483
    for (size_t j = 0; j < cell.loci.size(); ++j)
        for (size_t c = 0; c < prob.size(); ++c) {
            sums[c][cell.loci[j]] += prob[c] * (float)cell.alt_counts[j];
            denoms[c][cell.loci[j]] += prob[c] * (float)(cell.alt_counts[j] + cell.ref_counts[j]);
        }
}
static void update_final_with_lock(size_t loci, const std::vector<std::vector<float>>& sums,
                                   const std::vector<std::vector<float>>& denoms,
                                   std::vector<std::vector<float>>& centers, const std::vector<size_t>& locked) {
    for (size_t locus = 0; locus < loci; ++locus)                      // S:386-396
        for (size_t c = 0; c < sums.size(); ++c) {
            if (std::find(locked.begin(), locked.end(), c) != locked.end()) continue;
            float update = sums[c][locus] / denoms[c][locus];
            centers[c][locus] = fmaxf(fminf(update, 0.99f), 0.01f);   // Rust f32::min / max ignore a NaN operand
        }
}

// Rust f32::total_cmp
static inline int32_t total_key(float f) {
    int32_t b;
    memcpy(&b, &f, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}
static size_t argmax_last(const std::vector<float>& v) {              // Iterator::max_by => LAST maximum
    size_t best = 0;
    for (size_t i = 1; i < v.size(); ++i)
        if (total_key(v[i]) >= total_key(v[best])) best = i;
    return best;
}

static float cell_temp(ClusterMethod m, float total_alleles, int temp_step, int temp_steps) {
    float t;
    if (m == ClusterMethod::EM) t = std::max(total_alleles / (20.0f * powf(2.0f, (float)temp_step)), 1.0f);  // S:233
    else t = std::max(total_alleles / (0.5f * powf(2.0f, (float)temp_step)), 1.0f);                          // S:327
    if (temp_step == temp_steps - 1) t = 1.0f;
    return t;
}

// The body of the while loop for one cell (S:229-242 / S:315-333)
static float cell_pass(ClusterMethod m, const CellData& cell, const std::vector<std::vector<float>>& centers,
                       float log_prior, int temp_step, std::vector<float>* log_binoms, std::vector<float>* prob) {
    size_t min_clus;
    *log_binoms = binomial_loss_with_min_index(cell, centers, log_prior, &min_clus);
    float lse = log_sum_exp(*log_binoms);
    float temp = cell_temp(m, cell.total_alleles, temp_step, 9);
    if (m == ClusterMethod::EM) {
        *prob = normalize_in_log_with_temp(*log_binoms, temp);
    } else {
        std::vector<float> q_vec;
        float q_sum;
        calculate_q_for_current_cell(*log_binoms, min_clus, &q_vec, &q_sum);
        std::vector<float> khm_prob;
        for (float q : q_vec) khm_prob.push_back(q - q_sum);
        *prob = normalize_in_log_with_temp(khm_prob, temp);
    }
    return lse;
}

StepOut em_khm_step(size_t loci, std::vector<std::vector<float>>& centers, const std::vector<CellData>& cells,
                    ClusterMethod method, int temp_step, const std::vector<size_t>& locked) {
    size_t k = centers.size();
    std::vector<std::vector<float>> sums(k, std::vector<float>(loci, 1.0f)), denoms(k, std::vector<float>(loci, 2.0f));
    float log_prior = logf(1.0f / (float)k);
    StepOut out;
    out.log_probs.resize(cells.size());
    out.weights.resize(cells.size());
    out.lse.resize(cells.size());
    float log_binom_loss = 0.0f;
    for (size_t i = 0; i < cells.size(); ++i) {
        float lse = cell_pass(method, cells[i], centers, log_prior, temp_step, &out.log_probs[i], &out.weights[i]);
        out.lse[i] = lse;
        log_binom_loss += lse;
        update_centers_average(sums, denoms, cells[i], out.weights[i]);
    }
    out.loss = log_binom_loss;
    update_final_with_lock(loci, sums, denoms, centers, locked);
    return out;
}

// Oracle-only (Params::inner_threads > 1): the SAME per-cell and per-(cluster, locus) arithmetic chains as the loop of S:225-243,
// evaluated concurrently so the full-size parity tests finish in seconds.  Bit-identical to the serial loop by construction (and by
// test): a cell's pass (S:229-236) reads only the centers; the loss is then summed over cells in index order by one thread
// (S:230); and the accumulation S:476-483 is split by CLUSTER — every (cluster, locus) sum still takes the cells in ascending order.
static void iteration_in_parallel(ClusterMethod method, const std::vector<CellData>& cells, const std::vector<std::vector<float>>& centers,
                                  float log_prior, int temp_step, int n_threads, std::vector<std::vector<float>>* final_lp,
                                  std::vector<std::vector<float>>* sums, std::vector<std::vector<float>>* denoms, float* loss_out) {
    const size_t n = cells.size(), k = centers.size();
    std::vector<std::vector<float>> probs(n);
    std::vector<float> lse(n);
    std::vector<std::thread> pool;
    std::atomic<size_t> next{0};
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                const size_t i0 = next.fetch_add(64);
                if (i0 >= n) break;
                for (size_t i = i0; i < std::min(n, i0 + 64); ++i)
                    lse[i] = cell_pass(method, cells[i], centers, log_prior, temp_step, &(*final_lp)[i], &probs[i]);
            }
        });
    for (auto& th : pool) th.join();
    pool.clear();
    float loss = 0.0f;
    for (size_t i = 0; i < n; ++i) loss += lse[i];
    *loss_out = loss;
    std::atomic<size_t> next_k{0};
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&] {
            for (;;) {
                const size_t c = next_k.fetch_add(1);
                if (c >= k) break;
                std::vector<float>& s = (*sums)[c];
                std::vector<float>& d = (*denoms)[c];
                for (size_t i = 0; i < n; ++i) {
                    const CellData& cell = cells[i];
                    const float p = probs[i][c];
                    for (size_t j = 0; j < cell.loci.size(); ++j) {
                        s[cell.loci[j]] += p * (float)cell.alt_counts[j];
                        d[cell.loci[j]] += p * (float)(cell.alt_counts[j] + cell.ref_counts[j]);
                    }
                }
            }
        });
    for (auto& th : pool) th.join();
}

static std::mutex g_trace_mutex;

static float em_khm(ClusterMethod method, size_t loci, std::vector<std::vector<float>>& centers,
                    const std::vector<CellData>& cells, const Params& params, int epoch, int thread_num,
                    const std::vector<size_t>& locked, std::vector<std::vector<float>>* final_lp, Trace* trace,
                    int* iterations_out) {
    size_t k = params.num_clusters;
    std::vector<std::vector<float>> sums(k, std::vector<float>(loci, 1.0f)), denoms(k, std::vector<float>(loci, 2.0f));
    float log_prior = logf(1.0f / (float)k);                                    // S:202
    int iterations = 0;
    float total_log_loss = NEG_INF;
    final_lp->assign(cells.size(), std::vector<float>());
    float log_loss_change_limit = 0.01f * (float)cells.size();                  // S:214
    const int temp_steps = 9;
    float last_log_loss = NEG_INF;
    std::vector<float> prob;
    for (int temp_step = 0; temp_step < temp_steps; ++temp_step) {
        float log_loss_change = 10000.0f;                                       // S:219
        while (log_loss_change > log_loss_change_limit && iterations < params.iteration_cap) {   // S:222 (cap=1000)
            float log_binom_loss = 0.0f;
            reset_sums_denoms(loci, sums, denoms, k);
            if (params.inner_threads <= 1) {
                for (size_t i = 0; i < cells.size(); ++i) {
                    log_binom_loss += cell_pass(method, cells[i], centers, log_prior, temp_step, &(*final_lp)[i], &prob);
                    update_centers_average(sums, denoms, cells[i], prob);
                }
            } else {
                iteration_in_parallel(method, cells, centers, log_prior, temp_step, params.inner_threads, final_lp, &sums, &denoms, &log_binom_loss);
            }
            total_log_loss = log_binom_loss;
            log_loss_change = log_binom_loss - last_log_loss;
            last_log_loss = log_binom_loss;
            update_final_with_lock(loci, sums, denoms, centers, locked);
            iterations += 1;
            std::vector<size_t> cells_per_cluster(k, 0);                        // S:252-263
            for (const auto& lp : *final_lp) cells_per_cluster[argmax_last(lp)] += 1;
            int above = 0;
            for (size_t v : cells_per_cluster) if (v > 200) ++above;
            if (trace) {
                std::lock_guard<std::mutex> g(g_trace_mutex);
                trace->iters.push_back({thread_num, epoch, iterations, temp_step, log_binom_loss, log_loss_change, above});
                if (trace->text) {
                    if (method == ClusterMethod::EM)                            // S:264 ({:8} on the two floats)
                        fprintf(trace->text, "binomial\t%d\t%d\t%d\t%d\t%s\t%s\t%d\n", thread_num, epoch, iterations,
                                temp_step, format_f32_width8(log_binom_loss).c_str(),
                                format_f32_width8(log_loss_change).c_str(), above);
                    else                                                        // S:354
                        fprintf(trace->text, "binomial\t%d\t%d\t%d\t%d\t%s\t%s\t%d\n", thread_num, epoch, iterations,
                                temp_step, format_f32_display(log_binom_loss).c_str(),
                                format_f32_display(log_loss_change).c_str(), above);
                }
            }
        }
    }
    if (iterations_out) *iterations_out = iterations;
    return total_log_loss;
}

float em(size_t loci, std::vector<std::vector<float>>& centers, const std::vector<CellData>& cells, const Params& p,
         int epoch, int thread_num, const std::vector<size_t>& locked, std::vector<std::vector<float>>* final_lp,
         Trace* trace, int* iterations_out) {
    return em_khm(ClusterMethod::EM, loci, centers, cells, p, epoch, thread_num, locked, final_lp, trace, iterations_out);
}
float khm(size_t loci, std::vector<std::vector<float>>& centers, const std::vector<CellData>& cells, const Params& p,
          int epoch, int thread_num, const std::vector<size_t>& locked, std::vector<std::vector<float>>* final_lp,
          Trace* trace, int* iterations_out) {
    return em_khm(ClusterMethod::KHM, loci, centers, cells, p, epoch, thread_num, locked, final_lp, trace, iterations_out);
}

// ------------------------------------------------------------------ souporcell3  S:150-187
std::vector<size_t> bad_cluster_detection(size_t run, size_t num_clusters,
                                          const std::vector<std::vector<float>>& best_lp, FILE* log) {
    if (num_clusters < 16 || run == 0) return {};
    std::vector<std::pair<size_t, size_t>> assigned(num_clusters);
    for (size_t i = 0; i < num_clusters; ++i) assigned[i] = {i, 0};
    for (const auto& lp : best_lp) assigned[argmax_last(lp)].second += 1;
    std::stable_sort(assigned.begin(), assigned.end(),
                     [](const std::pair<size_t, size_t>& a, const std::pair<size_t, size_t>& b) { return a.second < b.second; });
    size_t run_value = 3 - run;
    size_t cut_low = run_value * (assigned.size() / 16);
    size_t cut_high = (16 - run_value) * (assigned.size() / 16);
    std::vector<size_t> replace;
    if (log) fprintf(log, "Cluster Analysis\n");
    for (size_t index = 0; index < assigned.size(); ++index) {
        size_t cluster = assigned[index].first;
        if (log) fprintf(log, "%zu:\tcluster\t%zu\tloss\t%zu", index, cluster, assigned[index].second);
        if (index < cut_low || index > cut_high) {
            if (std::find(replace.begin(), replace.end(), cluster) == replace.end()) replace.push_back(cluster);
            if (log) fprintf(log, " outlier");
        }
        if (log) fprintf(log, "\n");
    }
    return replace;
}

// ------------------------------------------------------------------ init  S:554-563
std::vector<std::vector<float>> init_cluster_centers_uniform(size_t loci, size_t k, ChaChaRng& rng) {
    std::vector<std::vector<float>> centers;
    for (size_t c = 0; c < k; ++c) {
        centers.emplace_back();
        for (size_t l = 0; l < loci; ++l) centers[c].push_back(std::max(std::min(rng.gen_f32(), 0.9999f), 0.0001f));
    }
    return centers;
}

// rand 0.9.3 `rng.gen_range(0..n)` for usize (third-party arithmetic, NOT under /root/reference, unpinned: restated
// from the published rand 0.9 sources).  `UniformUsize::sample_single` samples as u32 whenever the upper bound fits
// in u32 (portability fix of 0.9.0), and `UniformInt::<u32>::sample_single_inclusive` is Canon's method with one
// bias-reducing extra draw:  (hi, lo) = wmul(next_u32, range); if lo > 2^32 - range { hi += carry(lo + wmul(next_u32,
// range).hi) }.  For the cluster counts souporcell uses the extra draw has probability ~ n / 2^32 per call.
uint32_t gen_range_u32(ChaChaRng& rng, uint32_t n) {
    const uint32_t range = n;                                   // low = 0, high = n - 1, range = high - low + 1
    if (range == 0) return rng.next_u32();
    const uint64_t m = (uint64_t)rng.next_u32() * range;
    uint32_t result = (uint32_t)(m >> 32);
    const uint32_t lo_order = (uint32_t)m;
    if (lo_order > (uint32_t)(0u - range)) {
        const uint32_t new_hi_order = (uint32_t)(((uint64_t)rng.next_u32() * range) >> 32);
        if ((uint64_t)lo_order + new_hi_order > 0xffffffffull) result += 1;
    }
    return result;
}

// S:565-594
std::vector<std::vector<float>> init_cluster_centers_random_assignment(size_t loci, const std::vector<CellData>& cell_data,
                                                                       size_t num_clusters, ChaChaRng& rng) {
    std::vector<std::vector<float>> sums, denoms;
    for (size_t cluster = 0; cluster < num_clusters; ++cluster) {
        sums.emplace_back();
        denoms.emplace_back();
        for (size_t l = 0; l < loci; ++l) {
            sums[cluster].push_back(rng.gen_f32() * 0.01f);     // S:572
            denoms[cluster].push_back(0.01f);
        }
    }
    for (const CellData& cell : cell_data) {
        const size_t cluster = gen_range_u32(rng, (uint32_t)num_clusters);   // S:577
        for (size_t j = 0; j < cell.loci.size(); ++j) {
            const float alt_c = (float)cell.alt_counts[j];
            const float total = alt_c + (float)cell.ref_counts[j];
            sums[cluster][cell.loci[j]] += alt_c;
            denoms[cluster][cell.loci[j]] += total;
        }
    }
    for (size_t cluster = 0; cluster < num_clusters; ++cluster)
        for (size_t l = 0; l < loci; ++l) {
            float v = sums[cluster][l] / denoms[cluster][l] + (rng.gen_f32() / 2.0f - 0.25f);   // S:588
            v = fmaxf(fminf(v, 0.9999f), 0.0001f);                                                // S:589 (Rust f32::min / max ignore a NaN operand)
            sums[cluster][l] = v;
        }
    return sums;
}

// S:514-542.  The reference reads the VCF with the `vcf` crate (0.6, not vendored): records in file order, record
// index = original locus id, record.call[sample]["GT"][0] = the sample's GT string.  Restated here for plain or
// gzipped text VCF: FORMAT column -> position of GT, sample column from the #CHROM header line.
static std::string read_text_maybe_gz(const std::string& path) {
    std::string content;
    gzFile f = gzopen(path.c_str(), "rb");   // zlib reads plain files transparently
    if (!f) throw std::runtime_error("couldn't open file " + path);
    char buf[1 << 16];
    int n;
    while ((n = gzread(f, buf, sizeof(buf))) > 0) content.append(buf, n);
    gzclose(f);
    return content;
}
static std::vector<std::string> split_char(const std::string& s, char c) {
    std::vector<std::string> out;
    size_t i = 0;
    for (;;) {
        size_t j = s.find(c, i);
        if (j == std::string::npos) { out.emplace_back(s, i); break; }
        out.emplace_back(s, i, j - i);
        i = j + 1;
    }
    return out;
}
std::vector<std::vector<float>> init_cluster_centers_known_genotypes(size_t loci, const Params& params,
                                                                      const std::vector<size_t>& locus_to_index) {
    std::vector<std::vector<float>> centers(params.num_clusters, std::vector<float>(loci, 0.5f));   // S:515-521
    const std::string text = read_text_maybe_gz(params.known_genotypes);
    std::vector<std::string> header_samples;
    size_t locus_id = 0, pos = 0;
    while (pos < text.size()) {
        size_t e = text.find('\n', pos);
        if (e == std::string::npos) e = text.size();
        std::string line(text, pos, e - pos);
        pos = e + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '#') {
            if (line.rfind("#CHROM", 0) == 0) {
                auto cols = split_char(line, '\t');
                for (size_t i = 9; i < cols.size(); ++i) header_samples.push_back(cols[i]);
            }
            continue;
        }
        if (locus_id < locus_to_index.size() && locus_to_index[locus_id] != SIZE_MAX) {   // S:526
            const size_t loci_index = locus_to_index[locus_id];
            if (params.known_genotypes_sample_names.empty())
                throw std::runtime_error("currently requiring known_genotypes_sample_names if known_genotypes set");   // S:537
            auto cols = split_char(line, '\t');
            if (cols.size() < 10) throw std::runtime_error("known_genotypes: record without sample columns");
            auto fmt = split_char(cols[8], ':');
            size_t gt_pos = SIZE_MAX;
            for (size_t i = 0; i < fmt.size(); ++i) if (fmt[i] == "GT") gt_pos = i;
            for (size_t sample_index = 0; sample_index < params.known_genotypes_sample_names.size(); ++sample_index) {
                const std::string& sample = params.known_genotypes_sample_names[sample_index];
                size_t col = SIZE_MAX;
                for (size_t i = 0; i < header_samples.size(); ++i) if (header_samples[i] == sample) col = i;
                if (col == SIZE_MAX || 9 + col >= cols.size() || gt_pos == SIZE_MAX)
                    throw std::runtime_error("known_genotypes: sample or GT missing (reference panics on the map lookup)");
                auto fields = split_char(cols[9 + col], ':');
                if (gt_pos >= fields.size()) throw std::runtime_error("known_genotypes: GT missing in the sample column");
                const std::string& gt = fields[gt_pos];
                if (gt.empty()) throw std::runtime_error("known_genotypes: empty GT (reference unwraps None)");
                if (gt[0] == '.') continue;                                            // S:532
                auto hap = [&](size_t i) -> uint32_t {                                   // single-char parse::<u32>().unwrap().min(1)
                    if (i >= gt.size() || gt[i] < '0' || gt[i] > '9') throw std::runtime_error("known_genotypes: cannot parse GT '" + gt + "'");
                    return std::min<uint32_t>((uint32_t)(gt[i] - '0'), 1u);
                };
                const uint32_t hap0 = hap(0), hap1 = hap(2);                             // S:533-534
                if (sample_index >= centers.size()) throw std::runtime_error("known_genotypes: more sample names than clusters (index out of bounds)");
                centers[sample_index][loci_index] = fmaxf(fminf((float)(hap0 + hap1) / 2.0f, 0.99f), 0.01f);   // S:535
            }
        }
        ++locus_id;
    }
    return centers;
}

static std::vector<std::vector<float>> init_cluster_centers(size_t loci_used, const std::vector<CellData>& cell_data, const Params& params, ChaChaRng& rng,
                                                            size_t num_clusters, const std::vector<size_t>& locus_to_index) {   // S:485-498
    if (params.has_known_genotypes) return init_cluster_centers_known_genotypes(loci_used, params, locus_to_index);   // S:487
    if (params.has_known_cell_assignments) throw std::runtime_error("known cell assignments not yet implemented");
    switch (params.initialization_strategy) {
        case ClusterInit::KmeansPP: throw std::runtime_error("kmeans++ not yet implemented");
        case ClusterInit::RandomUniform: return init_cluster_centers_uniform(loci_used, num_clusters, rng);
        case ClusterInit::RandomAssignment: return init_cluster_centers_random_assignment(loci_used, cell_data, num_clusters, rng);
        case ClusterInit::MiddleVariance: throw std::runtime_error("middle variance not yet implemented");
    }
    return {};
}

// ------------------------------------------------------------------ driver  S:60-131
struct ThreadData {   // S:38-58
    std::vector<std::vector<float>> best_log_probabilities;
    float best_total_log_probability = NEG_INF;
    ChaChaRng rng;
    size_t solves_per_thread = 0, thread_num = 0;
    std::vector<std::vector<float>> cluster_centers;
    std::vector<float> solve_loss;
    std::vector<int> solve_iters;
    std::string error;
};

static void new_seed(ChaChaRng& rng, uint8_t seed[32]) {   // S:855-861
    for (int i = 0; i < 32; ++i) seed[i] = rng.gen_u8();
}

MainResult souporcell_main(size_t loci_used, const std::vector<CellData>& cells, const Params& params,
                           const std::vector<size_t>& locus_to_index, Trace* trace, FILE* log,
                           const float* theta0_override) {
    uint8_t seed[32];
    memset(seed, params.seed, 32);
    ChaChaRng rng = ChaChaRng::from_seed(seed);
    size_t solves_per_thread = (size_t)std::ceil((float)params.restarts / (float)params.threads);   // S:63
    MainResult res;
    res.best_log_probability = NEG_INF;
    size_t nnz = 0;
    for (const auto& c : cells) nnz += c.loci.size();
    const size_t K = params.num_clusters;
    for (size_t run = 0; run < 3; ++run) {
        std::vector<size_t> bad = bad_cluster_detection(run, K, res.best_log_probabilities, log);
        if (run != 0 && bad.empty()) {
            if (log) fprintf(log, "No outliers detected on run %zu!, Exiting.\n", run + 1);
            break;
        }
        res.runs = (int)run + 1;
        std::vector<ThreadData> threads(params.threads);
        for (size_t i = 0; i < params.threads; ++i) {
            uint8_t s[32];
            new_seed(rng, s);
            threads[i].rng = ChaChaRng::from_seed(s);
            threads[i].solves_per_thread = solves_per_thread;
            threads[i].thread_num = i;
        }
        auto work = [&](ThreadData& td) {
            try {
                for (size_t iteration = 0; iteration < td.solves_per_thread; ++iteration) {
                    std::vector<std::vector<float>> centers;
                    std::vector<size_t> locked;
                    if (run == 0) {
                        centers = init_cluster_centers(loci_used, cells, params, td.rng, K, locus_to_index);
                        if (theta0_override) {
                            const float* t0 = theta0_override + (td.thread_num * td.solves_per_thread + iteration) * K * loci_used;
                            for (size_t c = 0; c < K; ++c)
                                for (size_t l = 0; l < loci_used; ++l) centers[c][l] = t0[c * loci_used + l];
                        }
                    } else {                                                   // S:91-98
                        auto re = init_cluster_centers(loci_used, cells, params, td.rng, bad.size(), locus_to_index);
                        centers = res.best_cluster_centers;
                        for (size_t i = 0; i < bad.size(); ++i) centers[bad[i]] = re[i];
                        for (size_t c = 0; c < K; ++c)
                            if (std::find(bad.begin(), bad.end(), c) == bad.end()) locked.push_back(c);
                    }
                    std::vector<std::vector<float>> lp;
                    int iters = 0;
                    float loss = em_khm(params.clustering_method, loci_used, centers, cells, params, (int)iteration,
                                        (int)td.thread_num, locked, &lp, trace, &iters);
                    td.solve_loss.push_back(loss);
                    td.solve_iters.push_back(iters);
                    if (loss > td.best_total_log_probability) {                // S:110 strict >
                        td.best_total_log_probability = loss;
                        td.best_log_probabilities = std::move(lp);
                        td.cluster_centers = std::move(centers);
                    }
                    if (log) {
                        std::lock_guard<std::mutex> g(g_trace_mutex);
                        fprintf(log, "thread %zu iteration %zu done with %s, best so far %s\n", td.thread_num, iteration,
                                format_f32_display(loss).c_str(), format_f32_display(td.best_total_log_probability).c_str());
                    }
                }
            } catch (const std::exception& e) { td.error = e.what(); }
        };
        std::vector<std::thread> pool;
        for (auto& td : threads) pool.emplace_back(work, std::ref(td));
        for (auto& t : pool) t.join();
        for (auto& td : threads) {
            if (!td.error.empty()) throw std::runtime_error(td.error);
            for (size_t i = 0; i < td.solve_loss.size(); ++i) {
                res.solve_loss.push_back(td.solve_loss[i]);
                res.solve_iters.push_back(td.solve_iters[i]);
                res.nnz_updates += (uint64_t)nnz * (uint64_t)td.solve_iters[i];
            }
            if (td.best_total_log_probability > res.best_log_probability) {   // S:120 strict >
                res.best_log_probability = td.best_total_log_probability;
                res.best_log_probabilities = std::move(td.best_log_probabilities);
                res.best_cluster_centers = std::move(td.cluster_centers);
            }
        }
        if (!params.souporcell3) break;                                       // S:127
    }
    if (log) fprintf(log, "best total log probability = %s\n", format_f32_display(res.best_log_probability).c_str());
    return res;
}

// ------------------------------------------------------------------ Rust float Display
std::string format_f32_display(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    if (v == 0.0f) return std::signbit(v) ? "-0" : "0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);   // shortest round-trip digits
    std::string s(buf, r.ptr);
    bool neg = false;
    size_t i = 0;
    if (s[0] == '-') { neg = true; i = 1; }
    size_t epos = s.find('e');
    std::string mant = s.substr(i, epos - i);
    int exp10 = std::stoi(s.substr(epos + 1));
    std::string digits;
    for (char ch : mant) if (ch != '.') digits.push_back(ch);
    std::string out = neg ? "-" : "";
    int nd = (int)digits.size();
    if (exp10 >= 0) {
        if (nd <= exp10 + 1) {
            out += digits + std::string(exp10 + 1 - nd, '0');
        } else {
            out += digits.substr(0, exp10 + 1) + "." + digits.substr(exp10 + 1);
        }
    } else {
        out += "0." + std::string(-exp10 - 1, '0') + digits;
    }
    return out;
}
std::string format_f32_width8(float v) {
    std::string s = format_f32_display(v);
    if (s.size() < 8) s = std::string(8 - s.size(), ' ') + s;   // numerics right-align by default
    return s;
}

void write_clusters_tsv(FILE* out, const std::vector<std::string>& barcodes,
                        const std::vector<std::vector<float>>& best_lp) {   // S:133-147
    size_t n = std::min(barcodes.size(), best_lp.size());
    for (size_t i = 0; i < n; ++i) {
        const auto& lp = best_lp[i];
        size_t best = 0;
        float best_v = NEG_INF;
        for (size_t k = 0; k < lp.size(); ++k)
            if (lp[k] > best_v) { best = k; best_v = lp[k]; }
        fprintf(out, "%s\t%zu\t", barcodes[i].c_str(), best);
        for (size_t k = 0; k < lp.size(); ++k) {
            fputs(format_f32_display(lp[k]).c_str(), out);
            if (k + 1 < lp.size()) fputc('\t', out);
        }
        fputc('\n', out);
    }
}

// ------------------------------------------------------------------ CLI  S:756-853 + params.yml
bool parse_cli(int argc, const char* const* argv, Params* p, std::string* err) {
    struct Opt { const char* name; char shrt; bool multiple; };
    static const Opt opts[] = {
        {"ref_matrix", 'r', false}, {"alt_matrix", 'a', false}, {"barcodes", 'b', false}, {"num_clusters", 'k', false},
        {"min_alt", 0, false}, {"min_ref", 0, false}, {"restarts", 0, false}, {"known_genotypes", 'g', false},
        {"known_genotypes_sample_names", 0, true}, {"known_cell_assignments", 0, false},
        {"initialization_strategy", 0, false}, {"threads", 't', false}, {"seed", 0, false}, {"min_alt_umis", 0, false},
        {"min_ref_umis", 0, false}, {"souporcell3", 's', false}, {"clustering_method", 'm', false},
        {"iteration_cap", 0, false},   // oracle-only extension
    };
    std::unordered_map<std::string, std::vector<std::string>> vals;
    auto find_long = [&](const std::string& n) -> const Opt* {
        for (const auto& o : opts) if (n == o.name) return &o;
        return nullptr;
    };
    auto find_short = [&](char c) -> const Opt* {
        for (const auto& o : opts) if (o.shrt && o.shrt == c) return &o;
        return nullptr;
    };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        const Opt* o = nullptr;
        std::string inline_val;
        bool has_inline = false;
        if (a.rfind("--", 0) == 0) {
            std::string name = a.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) { inline_val = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
            o = find_long(name);
        } else if (a.size() >= 2 && a[0] == '-') {
            o = find_short(a[1]);
            if (a.size() > 2) { inline_val = a.substr(a[2] == '=' ? 3 : 2); has_inline = true; }
        }
        if (!o) { *err = "error: Found argument '" + a + "' which wasn't expected, or isn't valid in this context"; return false; }
        if (has_inline) { vals[o->name].push_back(inline_val); continue; }
        if (i + 1 >= argc) { *err = std::string("error: The argument '--") + o->name + "' requires a value but none was supplied"; return false; }
        vals[o->name].push_back(argv[++i]);
        if (o->multiple)
            while (i + 1 < argc && argv[i + 1][0] != '-') vals[o->name].push_back(argv[++i]);
    }
    for (const char* req : {"ref_matrix", "alt_matrix", "barcodes", "num_clusters"})
        if (!vals.count(req)) { *err = std::string("error: The following required arguments were not provided: --") + req; return false; }
    auto get = [&](const char* n, const char* dflt) { return vals.count(n) ? vals[n].back() : std::string(dflt); };
    auto num = [&](const char* n, const char* dflt, uint64_t maxv) {
        uint64_t v;
        std::string s = get(n, dflt);
        if (!parse_unsigned(s, &v) || v > maxv) throw std::runtime_error(std::string("invalid value for --") + n + ": " + s);
        return v;
    };
    try {
        p->ref_mtx = get("ref_matrix", "");
        p->alt_mtx = get("alt_matrix", "");
        p->barcodes = get("barcodes", "");
        p->num_clusters = (size_t)num("num_clusters", "0", UINT64_MAX);
        p->min_alt = (uint32_t)num("min_alt", "4", UINT32_MAX);
        p->min_ref = (uint32_t)num("min_ref", "4", UINT32_MAX);
        p->restarts = (uint32_t)num("restarts", "100", UINT32_MAX);
        p->has_known_cell_assignments = vals.count("known_cell_assignments") > 0;
        p->known_cell_assignments = get("known_cell_assignments", "");
        p->has_known_genotypes = vals.count("known_genotypes") > 0;
        p->known_genotypes = get("known_genotypes", "");
        if (p->has_known_genotypes && p->has_known_cell_assignments)
            throw std::runtime_error("Cannot set both known_genotypes and known_cell_assignments");
        if (vals.count("known_genotypes_sample_names")) p->known_genotypes_sample_names = vals["known_genotypes_sample_names"];
        std::string init = get("initialization_strategy", "random_uniform");
        if (init == "kmeans++") p->initialization_strategy = ClusterInit::KmeansPP;
        else if (init == "random_uniform") p->initialization_strategy = ClusterInit::RandomUniform;
        else if (init == "random_cell_assignment") p->initialization_strategy = ClusterInit::RandomAssignment;
        else if (init == "middle_variance") p->initialization_strategy = ClusterInit::MiddleVariance;
        else throw std::runtime_error("initialization strategy must be one of kmeans++, random_uniform, random_cell_assignment, middle_variance");
        p->threads = (size_t)num("threads", "1", UINT64_MAX);
        p->seed = (uint8_t)num("seed", "4", 255);
        p->min_ref_umis = (uint32_t)num("min_ref_umis", "0", UINT32_MAX);
        p->min_alt_umis = (uint32_t)num("min_alt_umis", "0", UINT32_MAX);
        std::string method = get("clustering_method", "em");
        if (method == "em") p->clustering_method = ClusterMethod::EM;
        else if (method == "khm") p->clustering_method = ClusterMethod::KHM;
        else throw std::runtime_error("Clustering method must be one of em, khm");
        std::string s3 = get("souporcell3", "false");
        if (s3 == "true") p->souporcell3 = true;
        else if (s3 == "false") p->souporcell3 = false;
        else throw std::runtime_error("provided string was not `true` or `false`");
        p->iteration_cap = (int)num("iteration_cap", "1000", 1000000);
    } catch (const std::exception& e) { *err = e.what(); return false; }
    return true;
}

}  // namespace orc
