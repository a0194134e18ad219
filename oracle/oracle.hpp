// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17, f32, -ffp-contract=off, no fast-math) of the
// genotype-clustering hot path of wheaton5/souporcell,
//   /root/reference/souporcell/src/main.rs        (abbreviated S: below)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library or run its CLI.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
//  and its Rust toolchain is absent from this image (F4), so the
// restatement is pinned only by (i) line-by-line review against S:, (ii) the
// hand-computed micro fixtures under tests/golden/, (iii) published ChaCha
// known-answer vectors.  Third-party arithmetic restated from its published
// algorithm:  rand 0.9.3 / rand_chacha 0.9.0 (StdRng = ChaCha12, one u32 word
// per f32/u8 draw), statrs 0.11.0 ln_binomial (factorial table n<=170, else
// ln_gamma — here libm lgamma, unpinned for n>170).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace orc {

// ---- rand_chacha 0.9.0 ChaChaXRng (key = seed, 64-bit block counter from 0,
// stream 0); words are consumed strictly sequentially (S:52,62,559,858).
struct ChaChaRng {
    uint32_t key[8];
    int rounds;            // 12 for StdRng (rand 0.9.3)
    uint64_t block;        // next block counter to generate
    uint32_t buf[16];
    int idx;               // next unread word in buf (16 = empty)
    ChaChaRng() : rounds(12), block(0), idx(16) {}
    static ChaChaRng from_seed(const uint8_t seed[32], int rounds = 12);
    uint32_t next_u32();
    // rand 0.9 StandardUniform: f32 = (u32 >> 8) * 2^-24 ; u8 = u32 as u8
    float gen_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    uint8_t gen_u8() { return (uint8_t)next_u32(); }
    void seek_word(uint64_t word);   // test helper: position at absolute word offset
};
void chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]);

// ---- S:683-703
struct CellData {
    std::vector<float> allele_fractions;
    std::vector<float> log_binomial_coefficient;
    std::vector<uint32_t> alt_counts;
    std::vector<uint32_t> ref_counts;
    std::vector<size_t> loci;
    float total_alleles = 0.0f;
};

enum class ClusterInit { KmeansPP, RandomUniform, RandomAssignment, MiddleVariance };
enum class ClusterMethod { EM, KHM };

// ---- S:721-740
struct Params {
    std::string ref_mtx, alt_mtx, barcodes;
    size_t num_clusters = 0;
    uint32_t min_alt = 4, min_ref = 4, min_alt_umis = 0, min_ref_umis = 0;
    uint32_t restarts = 100;
    bool has_known_cell_assignments = false;
    std::string known_cell_assignments;
    bool has_known_genotypes = false;
    std::string known_genotypes;
    std::vector<std::string> known_genotypes_sample_names;
    ClusterInit initialization_strategy = ClusterInit::RandomUniform;
    size_t threads = 1;
    uint8_t seed = 4;
    ClusterMethod clustering_method = ClusterMethod::EM;
    bool souporcell3 = false;
    // oracle-only knob (not in the reference): cap on iterations per solve so a
    // bounded CPU-baseline sample can be timed.  1000 = reference behaviour.
    int iteration_cap = 1000;
    // oracle-only knob: > 1 evaluates one iteration's independent chains on that many threads (bit-identical to the serial loop; see
    // iteration_in_parallel in oracle.cpp).  1 = the reference's loop as written.
    int inner_threads = 1;
};

struct LoadedData {
    size_t loci_used = 0;
    size_t total_cells = 0;
    std::vector<CellData> cell_data;
    std::vector<size_t> index_to_locus;
    // locus_to_index as a dense vector over header loci (SIZE_MAX = unused)
    std::vector<size_t> locus_to_index;
};

// one line of the reference's per-iteration stderr trace (S:264 / S:354)
struct IterRecord {
    int thread_num, epoch, iterations, temp_step;
    float loss, change;
    int clusters_above_threshold;
};

struct Trace {
    std::vector<IterRecord> iters;      // appended under a mutex by all threads
    FILE* text = nullptr;               // if non-null, reference-format lines go here
};

double ln_binomial(uint64_t n, uint64_t k);                       // statrs 0.11.0
LoadedData load_cell_data(const Params& p, FILE* log);            // S:601-681
LoadedData cell_data_from_csr(size_t n_cells, size_t n_loci, const uint64_t* row_ptr,
                              const uint32_t* locus, const uint32_t* alt, const uint32_t* ref);
std::vector<std::string> load_barcodes(const Params& p, FILE* log);   // S:707-718

std::vector<float> binomial_loss_with_min_index(const CellData& c, const std::vector<std::vector<float>>& centers,
                                                float log_prior, size_t* min_index);   // S:409-426
float log_sum_exp(const std::vector<float>& p);                                        // S:428-432
std::vector<float> normalize_in_log_with_temp(const std::vector<float>& lp, float temp);  // S:443-454
void calculate_q_for_current_cell(const std::vector<float>& ll, size_t min_clus,
                                  std::vector<float>* q_vec, float* q_sum);            // S:360-384

// One full E pass + M pass (one trip through the body of the while loop at
// S:222-265 / S:308-355) exposed for fine-grained parity tests.
struct StepOut {
    float loss;                                   // log_binom_loss (sequential f32 sum)
    std::vector<std::vector<float>> log_probs;    // [N][K]
    std::vector<std::vector<float>> weights;      // [N][K]
    std::vector<float> lse;                       // [N]
};
StepOut em_khm_step(size_t loci, std::vector<std::vector<float>>& centers, const std::vector<CellData>& cells,
                    ClusterMethod method, int temp_step, const std::vector<size_t>& locked);

// S:189-278 / S:280-358.  Returns (total_log_loss, final_log_probabilities).
float em(size_t loci, std::vector<std::vector<float>>& centers, const std::vector<CellData>& cells, const Params& p,
         int epoch, int thread_num, const std::vector<size_t>& locked, std::vector<std::vector<float>>* final_lp,
         Trace* trace, int* iterations_out);
float khm(size_t loci, std::vector<std::vector<float>>& centers, const std::vector<CellData>& cells, const Params& p,
          int epoch, int thread_num, const std::vector<size_t>& locked, std::vector<std::vector<float>>* final_lp,
          Trace* trace, int* iterations_out);

std::vector<size_t> bad_cluster_detection(size_t run, size_t num_clusters,
                                          const std::vector<std::vector<float>>& best_lp, FILE* log);  // S:150-187

std::vector<std::vector<float>> init_cluster_centers_uniform(size_t loci, size_t k, ChaChaRng& rng);   // S:554-563
struct CellData;
uint32_t gen_range_u32(ChaChaRng& rng, uint32_t n);   // rand 0.9.3 gen_range(0..n) for usize (restated, unpinned)
std::vector<std::vector<float>> init_cluster_centers_random_assignment(size_t loci, const std::vector<CellData>& cell_data, size_t num_clusters, ChaChaRng& rng);   // S:565-594
std::vector<std::vector<float>> init_cluster_centers_known_genotypes(size_t loci, const Params& params,
                                                                      const std::vector<size_t>& locus_to_index);   // S:514-542

struct MainResult {
    float best_log_probability;
    std::vector<std::vector<float>> best_log_probabilities;   // [N][K] (empty if nothing beat -inf)
    std::vector<std::vector<float>> best_cluster_centers;     // [K][L]
    std::vector<float> solve_loss;     // indexed run*T*spt + thread*spt + iteration
    std::vector<int> solve_iters;
    int runs = 0;
    uint64_t nnz_updates = 0;          // nnz * sum of iterations over all solves
};
// S:60-131 (everything except the TSV printing).  `theta0_override`, if non-null,
// supplies run-0 initial centers [solve][K][L] instead of the RNG (test hook).
MainResult souporcell_main(size_t loci_used, const std::vector<CellData>& cells, const Params& p,
                           const std::vector<size_t>& locus_to_index, Trace* trace, FILE* log,
                           const float* theta0_override = nullptr);

// S:133-147: assignment = FIRST max; Rust `{}` f32 formatting.
std::string format_f32_display(float v);                        // Rust Display for f32
std::string format_f32_width8(float v);                         // Rust {:8}
void write_clusters_tsv(FILE* out, const std::vector<std::string>& barcodes,
                        const std::vector<std::vector<float>>& best_lp);

bool parse_cli(int argc, const char* const* argv, Params* p, std::string* err);   // S:756-853 + params.yml

}  // namespace orc
