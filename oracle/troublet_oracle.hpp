// ORACLE — TEST INFRASTRUCTURE ONLY (see troublet_oracle.cpp).  Restatement of /root/reference/troublet/src/main.rs ("T:").
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

namespace troublet {

struct Params { double p_soup, doublet_prior, doublet_threshold, singlet_threshold; };   // T:462-471 (p_soup is not a CLI flag: 0.0)

struct Clusters {   // load_clusters T:414-449
    std::vector<std::string> barcodes;
    std::vector<std::vector<double>> losses;                 // per cell: the log-probability columns of clusters_tmp.tsv
    std::vector<std::pair<size_t, size_t>> cell_clusters;    // (best, second best) column
    size_t num_clusters = 0;                                 // max of column 2, plus one
};

struct Entry { size_t locus; uint64_t ref, alt; };

struct AlleleData {   // load_allele_data T:261-412
    size_t K = 0;
    std::vector<std::vector<std::pair<double, double>>> counts;   // [(c1, c2)][locus] = (ref, alt)
    std::vector<std::vector<double>> fractions;                   // [(c1, c2)][locus]
    std::vector<std::pair<double, double>> soup_counts;
    std::vector<double> soup_fractions;
    std::vector<std::vector<Entry>> cells;                        // per cell, file order
    std::vector<size_t> locus_counts;                             // lines per locus
    size_t num_0_counts = 0;
};

enum { STATUS_SINGLET = 0, STATUS_UNASSIGNED_SINGLET = 1, STATUS_DOUBLET = 2, STATUS_UNASSIGNED_DOUBLET = 3 };
struct CellCall {
    int status = 0, best_singlet = 0, doublet1 = 0, doublet2 = 0;
    double singlet_posterior = 0, doublet_posterior = 0, log_prob_singleton = 0, log_prob_doublet = 0;
};
struct Result {
    size_t k = 0;
    int rounds = 0;
    size_t removed = 0;
    std::vector<CellCall> cells;
    std::vector<std::vector<double>> cluster_logprobs;
};

double ln_gamma(double x);
double ln_beta(double a, double b);
double ln_binomial(uint64_t n, uint64_t k);
double log_sum_exp(const std::vector<double>& p);
std::string format_f64(double v);
Clusters load_clusters(const std::string& path);
AlleleData load_allele_data(const std::string& alts, const std::string& refs, const Clusters& cl);
Result call_doublets(const Params& params, AlleleData& d, const Clusters& cl, FILE* log);
void print_result(FILE* f, const Clusters& cl, const Result& r);

}  // namespace troublet
