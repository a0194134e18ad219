// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (the reference ships no fixtures and cannot be built here).
// CPU restatement of troublet, the doublet caller that consumes clusters_tmp.tsv and the two vartrix matrices:
// /root/reference/troublet/src/main.rs ("T:").  Only tests/, __graft_entry__.smoke() and bench.py may use it.
//
// Third-party arithmetic not under /root/reference (statrs, version pinned by troublet/Cargo.lock), restated from the
// published source, unpinned: factorial::ln_binomial (ln of the cached factorials up to 170!, ln_gamma above) and
// gamma::ln_gamma (Lanczos, g = 10.900511, 11 coefficients), beta::ln_beta = ln_gamma(a) + ln_gamma(b) - ln_gamma(a+b).
// f64 `ln`/`exp` are the host libm's, which is what Rust's f64::ln/exp call on linux-gnu.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "troublet_oracle.hpp"

namespace troublet {

// ---- statrs
static const double GAMMA_R = 10.900511;
static const double GAMMA_DK[11] = {
    2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469, 4.51227709466894823700,
    -2.98285225323576655721, 1.05639711577126713077, -1.95428773191645869583e-1, 1.70970543404441224307e-2,
    -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9};
static const double LN_2_SQRT_E_OVER_PI = 0.6207822376352452223455184457816472122518527279025978;
static const double LN_PI = 1.1447298858494001741434273513530587116472948129153;

double ln_gamma(double x) {   // statrs::function::gamma::ln_gamma
    if (x < 0.5) {
        double s = GAMMA_DK[0];
        for (int i = 1; i < 11; ++i) s = s + GAMMA_DK[i] / ((double)i - x);
        return LN_PI - std::log(std::sin(M_PI * x)) - std::log(s) - LN_2_SQRT_E_OVER_PI - (0.5 - x) * std::log((0.5 - x + GAMMA_R) / M_E);
    }
    double s = GAMMA_DK[0];
    for (int i = 1; i < 11; ++i) s = s + GAMMA_DK[i] / (x + (double)i - 1.0);
    return std::log(s) + LN_2_SQRT_E_OVER_PI + (x - 0.5) * std::log((x - 0.5 + GAMMA_R) / M_E);
}
double ln_beta(double a, double b) { return ln_gamma(a) + ln_gamma(b) - ln_gamma(a + b); }   // statrs::function::beta::ln_beta
static double ln_factorial(uint64_t x) {   // statrs::function::factorial::ln_factorial
    static double fcache[171];
    static bool init = false;
    if (!init) {
        fcache[0] = 1.0;
        for (int i = 1; i <= 170; ++i) fcache[i] = fcache[i - 1] * (double)i;
        init = true;
    }
    if (x <= 170) return std::log(fcache[x]);
    return ln_gamma((double)x + 1.0);
}
double ln_binomial(uint64_t n, uint64_t k) {   // statrs::function::factorial::ln_binomial
    if (k > n) return -INFINITY;
    return ln_factorial(n) - ln_factorial(k) - ln_factorial(n - k);
}

double log_sum_exp(const std::vector<double>& p) {   // T:246-250
    double max_p = -INFINITY;
    for (double x : p) max_p = (x > max_p || std::isnan(max_p)) ? x : max_p;   // f64::max fold
    double sum = 0.0;
    for (double x : p) sum += std::exp(x - max_p);
    return max_p + std::log(sum);
}

std::string format_f64(double v) {   // Rust `{}` for f64: shortest round-trip digits, never scientific
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    if (v == 0.0) return std::signbit(v) ? "-0" : "0";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf) - 1, v, std::chars_format::scientific);
    *res.ptr = '\0';
    const char* p = buf;
    std::string out;
    if (*p == '-') { out.push_back('-'); ++p; }
    std::string digits;
    const char* e = p;
    while (e < res.ptr && *e != 'e') { if (*e != '.') digits.push_back(*e); ++e; }
    const int exp10 = atoi(e + 1), nd = (int)digits.size();
    if (exp10 >= 0) {
        if (nd <= exp10 + 1) { out += digits; out.append((size_t)(exp10 + 1 - nd), '0'); }
        else { out.append(digits, 0, (size_t)exp10 + 1); out.push_back('.'); out.append(digits, (size_t)exp10 + 1, std::string::npos); }
    } else {
        out += "0.";
        out.append((size_t)(-exp10 - 1), '0');
        out += digits;
    }
    return out;
}

// ---- T:414-449
Clusters load_clusters(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("Unable to open file");
    Clusters c;
    size_t num_clusters = 0;
    std::string line;
    while (std::getline(f, line)) {
        // line.trim()
        size_t b = line.find_first_not_of(" \t\r\n"), e = line.find_last_not_of(" \t\r\n");
        line = b == std::string::npos ? std::string() : line.substr(b, e - b + 1);
        std::vector<std::string> tokens;
        size_t pos = 0;
        for (;;) {
            size_t t = line.find('\t', pos);
            tokens.push_back(line.substr(pos, t == std::string::npos ? std::string::npos : t - pos));
            if (t == std::string::npos) break;
            pos = t + 1;
        }
        if (tokens.size() < 2) throw std::runtime_error("clusters file: index out of bounds");
        const size_t cluster1 = (size_t)std::stoull(tokens[1]);
        num_clusters = std::max(num_clusters, cluster1);
        std::vector<double> losses;
        double best = -10000000000.0, second_best = -10000000000.0;
        size_t best_index = 0, second_best_index = 0;
        for (size_t i = 2; i < tokens.size(); ++i) {
            const double value = std::stod(tokens[i]);
            losses.push_back(value);
            if (value > best) { second_best = best; second_best_index = best_index; best = value; best_index = i - 2; }
            else if (value > second_best) { second_best = value; second_best_index = i - 2; }
        }
        c.barcodes.push_back(tokens[0]);
        c.losses.push_back(losses);
        c.cell_clusters.push_back({best_index, second_best_index});
    }
    c.num_clusters = num_clusters + 1;
    return c;
}

// ---- T:261-412 (the K x K tables of the reference are kept literally: this is the oracle)
AlleleData load_allele_data(const std::string& alts_path, const std::string& refs_path, const Clusters& cl) {
    const size_t K = cl.num_clusters;
    AlleleData d;
    d.K = K;
    std::ifstream alts(alts_path), refs(refs_path);
    if (!alts) throw std::runtime_error("unable to open alt file");
    if (!refs) throw std::runtime_error("unable to open ref file");
    d.counts.assign(K * K, {});
    size_t loci = 0, index = 0;
    std::string refline, altline;
    while (std::getline(refs, refline) && std::getline(alts, altline)) {
        const size_t idx = index++;
        if (!refline.empty() && refline[0] == '%') continue;
        // split_whitespace + parse::<usize>/<u64>, with the C library (no iostream extraction: this file is also loaded into
        // processes that already hold another libstdc++)
        auto three = [](const std::string& line, uint64_t out[3]) {
            const char* q = line.c_str();
            for (int i = 0; i < 3; ++i) {
                while (*q == ' ' || *q == '\t' || *q == '\r') ++q;
                if (*q < '0' || *q > '9') return i;
                char* e = nullptr;
                out[i] = strtoull(q, &e, 10);
                q = e;
            }
            return 3;
        };
        uint64_t rt[3] = {0, 0, 0}, at[3] = {0, 0, 0};
        if (idx == 2) {
            if (three(refline, rt) < 2) throw std::runtime_error("called `Result::unwrap()` on an `Err` value: ParseIntError (size line)");
            loci = rt[0];
            const size_t cells = rt[1];
            for (auto& v : d.counts) v.assign(loci, {0.0, 0.0});
            d.soup_counts.assign(loci, {0.0, 0.0});
            d.locus_counts.assign(loci, 0);
            d.cells.assign(cells, {});
            continue;
        }
        if (three(refline, rt) < 3 || three(altline, at) < 3) throw std::runtime_error("called `Option::unwrap()` on a `None` value (malformed matrix line)");
        const uint64_t rl = rt[0], rc = rt[1], rv = rt[2], al = at[0], ac = at[1], av = at[2];
        if (rl == 0 || rc == 0 || al == 0 || ac == 0) throw std::runtime_error("attempt to subtract with overflow");
        const size_t locus = rl - 1, cell = rc - 1;
        if (locus != al - 1 || cell != ac - 1) throw std::runtime_error("allele matrices do not match");   // T:297,300
        if (locus >= loci || cell >= d.cells.size() || cell >= cl.cell_clusters.size()) throw std::runtime_error("index out of bounds");
        d.locus_counts[locus] += 1;
        const size_t clust1 = cl.cell_clusters[cell].first;
        d.soup_counts[locus].first += (double)rv;
        d.soup_counts[locus].second += (double)av;
        for (size_t clust2 = 0; clust2 < K; ++clust2) {
            auto& v = d.counts[clust1 * K + clust2];
            v[locus].first += (double)rv;
            v[locus].second += (double)av;
            if (clust1 != clust2) {
                auto& w = d.counts[clust2 * K + clust1];
                w[locus].first += (double)rv;
                w[locus].second += (double)av;
            }
        }
        d.cells[cell].push_back({locus, rv, av});
    }
    d.soup_fractions.resize(loci);
    for (size_t l = 0; l < loci; ++l) {
        const double r = d.soup_counts[l].first, a = d.soup_counts[l].second;
        d.soup_fractions[l] = r + a > 0.0 ? r / (r + a) : 0.0;
    }
    const double mn = 0.01, mx = 0.99;
    d.fractions.assign(K * K, {});
    size_t num_0_counts = 0;
    for (size_t c1 = 0; c1 < K; ++c1)
        for (size_t c2 = 0; c2 < K; ++c2) {
            auto& nv = d.fractions[c1 * K + c2];
            nv.reserve(loci);
            const auto& k1 = d.counts[c1 * K + c1];
            const auto& k2 = d.counts[c2 * K + c2];
            for (size_t l = 0; l < loci; ++l) {
                const auto a = k1[l], b = k2[l];
                if (a.first + a.second > 0.0) {
                    double frac;
                    if (b.first + b.second > 0.0) frac = 0.5 * (a.first / (a.first + a.second)) + 0.5 * (b.first / (b.first + b.second));
                    else frac = a.first / (a.first + a.second);
                    if (frac > mx) frac = mx; else if (frac < mn) frac = mn;
                    nv.push_back(frac);
                } else {
                    num_0_counts += 1;
                    nv.push_back(0.5);
                }
            }
        }
    d.num_0_counts = num_0_counts;
    return d;
}

// ---- T:32-244
Result call_doublets(const Params& params, AlleleData& d, const Clusters& cl, FILE* log) {
    const size_t K = cl.losses.empty() ? 0 : cl.losses[0].size();
    if (K == 0) throw std::runtime_error("index out of bounds: the len is 0 but the index is 0");
    const size_t KK = d.K;
    if (K > KK) throw std::runtime_error("called `Option::unwrap()` on a `None` value");   // a (c, c) key the tables do not hold
    const size_t N = d.cells.size();
    if (cl.losses.size() < N) throw std::runtime_error("index out of bounds");
    Result out;
    out.k = K;
    std::vector<uint8_t> all_removed(N, 0);
    std::vector<size_t> best_singlets;
    size_t any_removed = 1;
    while (any_removed > 0) {
        out.rounds += 1;
        out.cells.assign(N, {});
        out.cluster_logprobs.assign(N, {});
        best_singlets.assign(N, 0);
        std::vector<size_t> new_removed;
        for (size_t cell = 0; cell < N; ++cell) {
            const auto& loci = d.cells[cell];
            size_t best_doublet1 = 0, best_doublet2 = 0, best_singlet = 0;
            double best_doublet_log_prob = -INFINITY, best_singleton_log_prob = -INFINITY;
            std::vector<std::pair<double, size_t>> singletons;
            for (size_t c1 = 0; c1 < K; ++c1) {
                const auto& saf = d.fractions[c1 * KK + c1];
                const auto& cv = d.counts[c1 * KK + c1];
                double lp = std::log(1.0 - params.doublet_prior);
                for (const Entry& en : loci) {
                    if (d.locus_counts[en.locus] < 20) continue;
                    const auto lc = cv[en.locus];
                    const double af = (1.0 - params.p_soup) * saf[en.locus] + params.p_soup * d.soup_fractions[en.locus];
                    const double count = lc.first + lc.second;
                    const double v = ln_binomial(en.ref + en.alt, en.ref) +
                                     ln_beta((double)en.ref + af * count + 1.0, (double)en.alt + (1.0 - af) * count + 1.0) -
                                     ln_beta(af * count + 1.0, (1.0 - af) * count + 1.0);
                    lp += v;
                }
                singletons.push_back({lp, c1});
                if (lp > best_singleton_log_prob) { best_singlet = c1; best_singleton_log_prob = lp; }
            }
            std::vector<double> cluster_logprobs;
            for (auto& s : singletons) cluster_logprobs.push_back(s.first);
            // sort_by(|(a,_),(b,_)| (-a).partial_cmp(&-b).unwrap()): stable, descending; a NaN would panic
            for (auto& s : singletons) if (std::isnan(s.first)) throw std::runtime_error("called `Option::unwrap()` on a `None` value");
            std::stable_sort(singletons.begin(), singletons.end(), [](const auto& a, const auto& b) { return -a.first < -b.first; });
            const size_t top = std::min<size_t>(K, 4);
            for (size_t i = 0; i < top; ++i)
                for (size_t j = i + 1; j < top; ++j) {
                    const size_t c1 = singletons[i].second, c2 = singletons[j].second;
                    const auto& daf = d.fractions[c1 * KK + c2];
                    const auto& v1 = d.counts[c1 * KK + c1];
                    const auto& v3 = d.counts[c2 * KK + c2];
                    double lp = std::log(params.doublet_prior);
                    for (const Entry& en : loci) {
                        if (d.locus_counts[en.locus] < 20) continue;
                        const auto a = v1[en.locus], b = v3[en.locus];
                        const double af = (1.0 - params.p_soup) * daf[en.locus] + params.p_soup * d.soup_fractions[en.locus];
                        const double count = ((a.first + a.second) + (b.first + b.second)) / 2.0;
                        const double v = ln_binomial(en.ref + en.alt, en.ref) +
                                         ln_beta((double)en.ref + af * count + 1.0, (double)en.alt + (1.0 - af) * count + 1.0) -
                                         ln_beta(af * count + 1.0, (1.0 - af) * count + 1.0);
                        lp += v;
                    }
                    if (lp > best_doublet_log_prob) { best_doublet1 = c1; best_doublet2 = c2; best_doublet_log_prob = lp; }
                }
            const double dq_denom = log_sum_exp({best_singleton_log_prob, best_doublet_log_prob});
            double top_singlet = -INFINITY;
            size_t singlet_assignment = 0;
            for (size_t c = 0; c < cl.losses[cell].size(); ++c)
                if (cl.losses[cell][c] > top_singlet) { top_singlet = cl.losses[cell][c]; singlet_assignment = c; }
            best_singlets[cell] = singlet_assignment;
            const double sq_denom = log_sum_exp(cl.losses[cell]);
            CellCall cc;
            cc.singlet_posterior = std::exp(top_singlet - sq_denom);
            cc.doublet_posterior = std::exp(best_doublet_log_prob - dq_denom);
            cc.best_singlet = (int)best_singlet; cc.doublet1 = (int)best_doublet1; cc.doublet2 = (int)best_doublet2;
            cc.log_prob_singleton = best_singleton_log_prob; cc.log_prob_doublet = best_doublet_log_prob;
            if (best_singleton_log_prob > best_doublet_log_prob) {
                cc.status = cc.singlet_posterior > params.singlet_threshold ? STATUS_SINGLET : STATUS_UNASSIGNED_SINGLET;
            } else if (cc.doublet_posterior >= params.doublet_threshold) {
                if (!all_removed[cell]) {
                    all_removed[cell] = 1;
                    new_removed.push_back(cell);
                    if (log) fprintf(log, "removing %s as doublet\n", cl.barcodes[cell].c_str());
                }
                cc.status = STATUS_DOUBLET;
            } else {
                cc.status = STATUS_UNASSIGNED_DOUBLET;
            }
            out.cells[cell] = cc;
            if (singlet_assignment != best_singlet && !all_removed[cell] && log)
                fprintf(log, "error\t%s\t%zu\t%zu\n", cl.barcodes[cell].c_str(), singlet_assignment, best_singlet);
            out.cluster_logprobs[cell] = cluster_logprobs;
        }
        if (log) fprintf(log, "%zu cells removed this round\n", new_removed.size());
        any_removed = new_removed.size();
        out.removed += new_removed.size();
        for (size_t cell : new_removed) {
            const size_t c1 = best_singlets[cell];
            for (size_t c2 = 0; c2 < K; ++c2) {
                auto& counts = d.counts[c1 * KK + c2];
                for (const Entry& en : d.cells[cell]) {
                    counts[en.locus].first = std::max(counts[en.locus].first - (double)en.ref, 0.0);
                    counts[en.locus].second = std::max(counts[en.locus].second - (double)en.alt, 0.0);
                }
            }
        }
        for (size_t c1 = 0; c1 < K; ++c1)
            for (size_t c2 = 0; c2 < K; ++c2) {
                auto& fr = d.fractions[c1 * KK + c2];
                const auto& cc = d.counts[c1 * KK + c2];
                for (size_t l = 0; l < fr.size(); ++l)
                    if (cc[l].first + cc[l].second > 0.0) fr[l] = std::max(std::min(cc[l].first / (cc[l].first + cc[l].second), 0.99), 0.01);
            }
    }
    return out;
}

static const char* STATUS_NAME[4] = {"singlet", "unassigned", "doublet", "unassigned"};

void print_result(FILE* f, const Clusters& cl, const Result& r) {   // T:36-41, T:233-243 and the format! calls T:176-201
    fprintf(f, "barcode\tstatus\tassignment\tsinglet_posterior\tdoublet_posterior\tlog_prob_singleton\tlog_prob_doublet\t");
    for (size_t c = 0; c < r.k; ++c) fprintf(f, "cluster%zu%s", c, c + 1 == r.k ? "\n" : "\t");
    for (size_t cell = 0; cell < r.cells.size(); ++cell) {
        const CellCall& cc = r.cells[cell];
        std::string assign = cc.status == STATUS_SINGLET || cc.status == STATUS_UNASSIGNED_SINGLET
                                 ? std::to_string(cc.best_singlet)
                                 : std::to_string(cc.doublet1) + "/" + std::to_string(cc.doublet2);
        fprintf(f, "%s\t%s\t%s\t%s\t%s\t%s\t%s\t", cl.barcodes[cell].c_str(), STATUS_NAME[cc.status], assign.c_str(),
                format_f64(cc.singlet_posterior).c_str(), format_f64(cc.doublet_posterior).c_str(),
                format_f64(cc.log_prob_singleton).c_str(), format_f64(cc.log_prob_doublet).c_str());
        const auto& lp = r.cluster_logprobs[cell];
        for (size_t i = 0; i < lp.size(); ++i) fprintf(f, "%s%s", format_f64(lp[i]).c_str(), i + 1 < lp.size() ? "\t" : "");
        fprintf(f, "\n");
    }
}

}  // namespace troublet

// ------------------------------------------------------------------ C API (ctypes) and CLI
extern "C" {
// returns the number of cells, or -1 (message in err).  Arrays are caller-allocated with n_cells_cap / k_cap capacity.
int64_t trb_run(const char* alts, const char* refs, const char* clusters, double p_soup, double doublet_prior, double doublet_threshold,
                double singlet_threshold, int64_t n_cells_cap, int32_t k_cap, int32_t* status, int32_t* best_singlet, int32_t* d1, int32_t* d2,
                double* singlet_post, double* doublet_post, double* lp_singleton, double* lp_doublet, double* cluster_logprobs,
                int32_t* k_out, int32_t* rounds, int64_t* removed, char* err, int32_t errlen) {
    try {
        using namespace troublet;
        Clusters cl = load_clusters(clusters);
        AlleleData d = load_allele_data(alts, refs, cl);
        Params p{p_soup, doublet_prior, doublet_threshold, singlet_threshold};
        Result r = call_doublets(p, d, cl, nullptr);
        if ((int64_t)r.cells.size() > n_cells_cap || (int32_t)r.k > k_cap) throw std::runtime_error("output capacity too small");
        for (size_t c = 0; c < r.cells.size(); ++c) {
            const CellCall& cc = r.cells[c];
            status[c] = cc.status; best_singlet[c] = cc.best_singlet; d1[c] = cc.doublet1; d2[c] = cc.doublet2;
            singlet_post[c] = cc.singlet_posterior; doublet_post[c] = cc.doublet_posterior;
            lp_singleton[c] = cc.log_prob_singleton; lp_doublet[c] = cc.log_prob_doublet;
            for (size_t k = 0; k < r.k; ++k) cluster_logprobs[c * r.k + k] = r.cluster_logprobs[c][k];
        }
        *k_out = (int32_t)r.k; *rounds = r.rounds; *removed = (int64_t)r.removed;
        return (int64_t)r.cells.size();
    } catch (const std::exception& e) {
        if (err && errlen > 0) { strncpy(err, e.what(), (size_t)errlen - 1); err[errlen - 1] = 0; }
        return -1;
    }
}
double trb_ln_gamma(double x) { return troublet::ln_gamma(x); }
double trb_ln_binomial(uint64_t n, uint64_t k) { return troublet::ln_binomial(n, k); }
int32_t trb_format_f64(double v, char* buf, int32_t buflen) {
    std::string s = troublet::format_f64(v);
    if ((int32_t)s.size() + 1 > buflen) return -1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int32_t)s.size();
}
}

#ifdef TROUBLET_ORACLE_MAIN
int main(int argc, char** argv) {
    using namespace troublet;
    std::string alts, refs, clusters;
    Params p{0.0, 0.5, 0.9, 0.9};
    auto need = [&](int& i) -> std::string { if (i + 1 >= argc) { fprintf(stderr, "error: missing value for %s\n", argv[i]); exit(1); } return argv[++i]; };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], v;
        size_t eq = a.find('=');
        bool has = a.rfind("--", 0) == 0 && eq != std::string::npos;
        if (has) { v = a.substr(eq + 1); a = a.substr(0, eq); }
        auto val = [&]() { return has ? v : need(i); };
        if (a == "-a" || a == "--alts") alts = val();
        else if (a == "-r" || a == "--refs") refs = val();
        else if (a == "-c" || a == "--clusters") clusters = val();
        else if (a == "-d" || a == "--doublet_prior") p.doublet_prior = std::stod(val());
        else if (a == "--doublet_threshold") p.doublet_threshold = std::stod(val());
        else if (a == "--singlet_threshold") p.singlet_threshold = std::stod(val());
        else if (a == "-b" || a == "--debug") val();
        else { fprintf(stderr, "error: Found argument '%s' which wasn't expected\n", argv[i]); return 1; }
    }
    if (alts.empty() || clusters.empty()) { fprintf(stderr, "error: The following required arguments were not provided: --alts --clusters\n"); return 1; }
    if (refs.empty()) { fprintf(stderr, "thread 'main' panicked at 'called `Option::unwrap()` on a `None` value'\n"); return 101; }   // T:472
    try {
        Clusters cl = load_clusters(clusters);
        AlleleData d = load_allele_data(alts, refs, cl);
        fprintf(stderr, "%zu loaded 0 counts, is this a problem?\n", d.num_0_counts);
        // the header line is printed before the loop (T:36-41); print_result writes it first as well
        Result r = call_doublets(p, d, cl, stderr);
        print_result(stdout, cl, r);
    } catch (const std::exception& e) {
        fprintf(stderr, "thread 'main' panicked at '%s'\n", e.what());
        return 101;
    }
    return 0;
}
#endif
