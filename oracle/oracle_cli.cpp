// ORACLE — TEST INFRASTRUCTURE ONLY.  `souporcell`-compatible CLI over the CPU restatement
// (main() S:31-36): same argv, clusters_tmp.tsv on stdout, the reference's log on stderr.
// Used by tests (CLI parity) and bench.py (cpu_baseline / --impl reference).  PARITY UNPINNED.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "oracle.hpp"

int main(int argc, char** argv) {
    using namespace orc;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-V") || !strcmp(argv[i], "--version")) { printf("souporcell 3.0\n"); return 0; }
        if (!strcmp(argv[i], "-h") || !strcmp(argv[i], "--help")) {
            printf("souporcell 3.0 (CPU oracle restatement)\nclustering scRNAseq cells by genotype\n");
            return 0;
        }
    }
    Params p;
    std::string err;
    if (!parse_cli(argc, argv, &p, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    try {
        if (p.num_clusters > 16 && (p.clustering_method == ClusterMethod::EM || !p.souporcell3))
            fprintf(stderr, "For k > 16, using souporcell3 (with '-s true' flag) is recommended with khm clustering "
                            "method (with '-m khm' flag)\n");
        auto barcodes = load_barcodes(p, stderr);
        LoadedData d = load_cell_data(p, stderr);
        Trace tr;
        tr.text = stderr;
        auto t0 = std::chrono::steady_clock::now();
        MainResult r = souporcell_main(d.loci_used, d.cell_data, p, d.locus_to_index, &tr, stderr);
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "oracle_timing\tsolve_seconds\t%.6f\tnnz_updates\t%llu\n", secs,
                (unsigned long long)r.nnz_updates);
        write_clusters_tsv(stdout, barcodes, r.best_log_probabilities);
    } catch (const std::exception& e) {
        fprintf(stderr, "thread 'main' panicked: %s\n", e.what());
        return 101;
    }
    return 0;
}
