// ORACLE — TEST INFRASTRUCTURE ONLY.  C ABI over oracle.cpp for ctypes (tests/, bench.py
// cpu_baseline / --impl reference, __graft_entry__.smoke()).  PARITY UNPINNED (see oracle.hpp).
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "oracle.hpp"

using namespace orc;

static void set_err(char* err, int errlen, const char* msg) {
    if (err && errlen > 0) { strncpy(err, msg, errlen - 1); err[errlen - 1] = 0; }
}
static std::vector<std::vector<float>> to_vv(const float* p, size_t rows, size_t cols) {
    std::vector<std::vector<float>> v(rows, std::vector<float>(cols));
    for (size_t r = 0; r < rows; ++r) memcpy(v[r].data(), p + r * cols, cols * sizeof(float));
    return v;
}
static void from_vv(const std::vector<std::vector<float>>& v, float* p, size_t cols) {
    for (size_t r = 0; r < v.size(); ++r) memcpy(p + r * cols, v[r].data(), cols * sizeof(float));
}

extern "C" {

void* orc_data_from_csr(uint64_t n_cells, uint64_t n_loci, const uint64_t* row_ptr, const uint32_t* locus,
                        const uint32_t* alt, const uint32_t* ref) {
    return new LoadedData(cell_data_from_csr(n_cells, n_loci, row_ptr, locus, alt, ref));
}
void* orc_data_load_mtx(const char* alt_mtx, const char* ref_mtx, uint32_t min_alt, uint32_t min_ref,
                        uint32_t min_alt_umis, uint32_t min_ref_umis, char* err, int errlen) {
    try {
        Params p;
        p.alt_mtx = alt_mtx; p.ref_mtx = ref_mtx;
        p.min_alt = min_alt; p.min_ref = min_ref; p.min_alt_umis = min_alt_umis; p.min_ref_umis = min_ref_umis;
        return new LoadedData(load_cell_data(p, nullptr));
    } catch (const std::exception& e) { set_err(err, errlen, e.what()); return nullptr; }
}
void orc_data_free(void* d) { delete (LoadedData*)d; }
uint64_t orc_data_n_cells(void* d) { return ((LoadedData*)d)->total_cells; }
uint64_t orc_data_n_loci(void* d) { return ((LoadedData*)d)->loci_used; }
uint64_t orc_data_nnz(void* d) {
    uint64_t n = 0;
    for (auto& c : ((LoadedData*)d)->cell_data) n += c.loci.size();
    return n;
}
// any output pointer may be NULL
void orc_data_export(void* dv, uint64_t* row_ptr, uint32_t* locus, uint32_t* alt, uint32_t* ref, float* lbc,
                     float* total_alleles, uint64_t* index_to_locus) {
    LoadedData* d = (LoadedData*)dv;
    uint64_t pos = 0;
    for (size_t c = 0; c < d->cell_data.size(); ++c) {
        const CellData& cd = d->cell_data[c];
        if (row_ptr) row_ptr[c] = pos;
        for (size_t j = 0; j < cd.loci.size(); ++j, ++pos) {
            if (locus) locus[pos] = (uint32_t)cd.loci[j];
            if (alt) alt[pos] = cd.alt_counts[j];
            if (ref) ref[pos] = cd.ref_counts[j];
            if (lbc) lbc[pos] = cd.log_binomial_coefficient[j];
        }
        if (total_alleles) total_alleles[c] = cd.total_alleles;
    }
    if (row_ptr) row_ptr[d->cell_data.size()] = pos;
    if (index_to_locus)
        for (size_t i = 0; i < d->index_to_locus.size(); ++i) index_to_locus[i] = d->index_to_locus[i];
}

// one E pass + M pass; theta [K][L] updated in place
int orc_step(void* dv, int K, int method, int temp_step, float* theta, const int* locked, int n_locked,
             float* log_probs, float* weights, float* lse, float* loss) {
    LoadedData* d = (LoadedData*)dv;
    auto centers = to_vv(theta, K, d->loci_used);
    std::vector<size_t> lk(locked, locked + n_locked);
    StepOut o = em_khm_step(d->loci_used, centers, d->cell_data, method ? ClusterMethod::KHM : ClusterMethod::EM,
                            temp_step, lk);
    from_vv(centers, theta, d->loci_used);
    if (log_probs) from_vv(o.log_probs, log_probs, K);
    if (weights) from_vv(o.weights, weights, K);
    if (lse) memcpy(lse, o.lse.data(), o.lse.size() * sizeof(float));
    if (loss) *loss = o.loss;
    return 0;
}

// one full em()/khm() solve from theta [K][L] (updated in place).  trace_* arrays hold
// `trace_cap` entries (per-iteration loss / change / temp_step / clusters_above_threshold).
static int orc_solve_impl(void* dv, int K, int method, float* theta, const int* locked, int n_locked, int iteration_cap, int inner_threads,
              float* log_probs, float* loss, int* iterations, int trace_cap, float* trace_loss, float* trace_change,
              int* trace_temp_step, int* trace_above) {
    LoadedData* d = (LoadedData*)dv;
    auto centers = to_vv(theta, K, d->loci_used);
    std::vector<size_t> lk(locked, locked + n_locked);
    Params p;
    p.num_clusters = K;
    p.iteration_cap = iteration_cap;
    p.inner_threads = inner_threads;
    Trace tr;
    std::vector<std::vector<float>> lp;
    int iters = 0;
    float l = method ? khm(d->loci_used, centers, d->cell_data, p, 0, 0, lk, &lp, &tr, &iters)
                     : em(d->loci_used, centers, d->cell_data, p, 0, 0, lk, &lp, &tr, &iters);
    from_vv(centers, theta, d->loci_used);
    if (log_probs) from_vv(lp, log_probs, K);
    if (loss) *loss = l;
    if (iterations) *iterations = iters;
    for (int i = 0; i < (int)tr.iters.size() && i < trace_cap; ++i) {
        if (trace_loss) trace_loss[i] = tr.iters[i].loss;
        if (trace_change) trace_change[i] = tr.iters[i].change;
        if (trace_temp_step) trace_temp_step[i] = tr.iters[i].temp_step;
        if (trace_above) trace_above[i] = tr.iters[i].clusters_above_threshold;
    }
    return 0;
}

int orc_solve(void* dv, int K, int method, float* theta, const int* locked, int n_locked, int iteration_cap,
              float* log_probs, float* loss, int* iterations, int trace_cap, float* trace_loss, float* trace_change,
              int* trace_temp_step, int* trace_above) {
    return orc_solve_impl(dv, K, method, theta, locked, n_locked, iteration_cap, 1, log_probs, loss, iterations, trace_cap, trace_loss,
                          trace_change, trace_temp_step, trace_above);
}
// the same with the iteration's independent chains spread over `inner_threads` threads (bit-identical; Params::inner_threads)
int orc_solve_mt(void* dv, int K, int method, float* theta, const int* locked, int n_locked, int iteration_cap, int inner_threads,
                 float* log_probs, float* loss, int* iterations, int trace_cap, float* trace_loss, float* trace_change,
                 int* trace_temp_step, int* trace_above) {
    return orc_solve_impl(dv, K, method, theta, locked, n_locked, iteration_cap, inner_threads, log_probs, loss, iterations, trace_cap,
                          trace_loss, trace_change, trace_temp_step, trace_above);
}

// souporcell_main without the TSV print.  Returns 0 ok, 1 if nothing beat -inf (no best), <0 error.
int orc_main(void* dv, int K, int method, uint32_t restarts, uint32_t threads, uint8_t seed, int souporcell3,
             int iteration_cap, const float* theta0_override, float* best_loss, float* best_log_probs,
             float* best_theta, float* solve_loss, int* solve_iters, int solve_cap, int* n_solves, int* runs,
             uint64_t* nnz_updates, char* err, int errlen) {
    try {
        LoadedData* d = (LoadedData*)dv;
        Params p;
        p.num_clusters = K;
        p.clustering_method = method ? ClusterMethod::KHM : ClusterMethod::EM;
        p.restarts = restarts; p.threads = threads; p.seed = seed; p.souporcell3 = souporcell3 != 0;
        p.iteration_cap = iteration_cap;
        MainResult r = souporcell_main(d->loci_used, d->cell_data, p, d->locus_to_index, nullptr, nullptr, theta0_override);
        if (best_loss) *best_loss = r.best_log_probability;
        if (best_log_probs && !r.best_log_probabilities.empty()) from_vv(r.best_log_probabilities, best_log_probs, K);
        if (best_theta && !r.best_cluster_centers.empty()) from_vv(r.best_cluster_centers, best_theta, d->loci_used);
        int n = (int)r.solve_loss.size();
        for (int i = 0; i < n && i < solve_cap; ++i) {
            if (solve_loss) solve_loss[i] = r.solve_loss[i];
            if (solve_iters) solve_iters[i] = r.solve_iters[i];
        }
        if (n_solves) *n_solves = n;
        if (runs) *runs = r.runs;
        if (nnz_updates) *nnz_updates = r.nnz_updates;
        return r.best_log_probabilities.empty() ? 1 : 0;
    } catch (const std::exception& e) { set_err(err, errlen, e.what()); return -1; }
}

// init_cluster_centers_known_genotypes S:514-542.  sample_names: n NUL-terminated strings back to back.  theta [K][L].
int orc_known_genotypes(void* dv, const char* vcf_path, const char* sample_names, int n_samples, int K, float* theta, char* err, int errlen) {
    try {
        LoadedData* d = (LoadedData*)dv;
        Params p;
        p.num_clusters = K;
        p.has_known_genotypes = true;
        p.known_genotypes = vcf_path;
        const char* q = sample_names;
        for (int i = 0; i < n_samples; ++i) { p.known_genotypes_sample_names.emplace_back(q); q += strlen(q) + 1; }
        auto c = init_cluster_centers_known_genotypes(d->loci_used, p, d->locus_to_index);
        from_vv(c, theta, d->loci_used);
        return 0;
    } catch (const std::exception& e) { set_err(err, errlen, e.what()); return -1; }
}

void orc_chacha_words(const uint8_t* seed, int rounds, uint64_t word_offset, uint32_t* out, uint64_t n) {
    ChaChaRng r = ChaChaRng::from_seed(seed, rounds);
    if (word_offset) r.seek_word(word_offset);
    for (uint64_t i = 0; i < n; ++i) out[i] = r.next_u32();
}
// thread seeds of run `run` (S:61-62, 77-80, 855-861): out[threads][32]
void orc_thread_seeds(uint8_t seed, uint32_t threads, uint32_t run, uint8_t* out) {
    uint8_t s[32];
    memset(s, seed, 32);
    ChaChaRng r = ChaChaRng::from_seed(s);
    for (uint32_t rr = 0; rr <= run; ++rr)
        for (uint32_t t = 0; t < threads; ++t)
            for (int i = 0; i < 32; ++i) {
                uint8_t b = r.gen_u8();
                if (rr == run) out[t * 32 + i] = b;
            }
}
void orc_init_uniform(const uint8_t* seed, uint64_t word_offset, int K, uint64_t L, float* theta) {
    ChaChaRng r = ChaChaRng::from_seed(seed);
    if (word_offset) r.seek_word(word_offset);
    auto c = init_cluster_centers_uniform(L, K, r);
    from_vv(c, theta, L);
}
double orc_ln_binomial(uint64_t n, uint64_t k) { return ln_binomial(n, k); }
int orc_format_f32(float v, int width8, char* buf, int buflen) {
    std::string s = width8 ? format_f32_width8(v) : format_f32_display(v);
    if ((int)s.size() + 1 > buflen) return -1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
int orc_bad_cluster_detection(int run, int K, const float* best_lp, uint64_t n_cells, int* out_idx) {
    auto lp = to_vv(best_lp, n_cells, K);
    auto bad = bad_cluster_detection(run, K, lp, nullptr);
    for (size_t i = 0; i < bad.size(); ++i) out_idx[i] = (int)bad[i];
    return (int)bad.size();
}
float orc_logf(float x) { return logf(x); }
float orc_expf(float x) { return expf(x); }

}  // extern "C"
