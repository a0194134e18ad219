// `souporcell` — drop-in executable for the reference clustering binary
// (/root/reference/souporcell/src/main.rs:31-36 main(), argv schema souporcell/src/params.yml:1-98).
// Same flags, clusters_tmp.tsv on stdout, the reference's log lines on stderr, non-zero exit on
// any failure with nothing written to stdout.  All numerics run on the GPU through libsoupgpu.
//
// Extra, optional flags (absent from the reference; the pipeline's argv works unchanged):
//   --gpus N           use the first N visible GPUs (restart-sharded); default: as many as the job can keep busy
//   --max_slots N      resident solves per GPU (0 = auto)
//   --quiet_iterations do not print the per-iteration "binomial" lines
//   --csr_cache FILE   binary cache of the parsed + filtered matrix (.soupcsr): read when it exists and was built
//                      with the same filter, written after the text parse otherwise
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <algorithm>
#include <thread>
#include <vector>

#include <chrono>
#include <future>

#include <unistd.h>

#include <cuda_runtime_api.h>

#include "soup_common.hpp"

namespace {

struct Opt { const char* name; char shrt; bool multiple; bool flag; const char* help; };
const Opt kOpts[] = {
    {"ref_matrix", 'r', false, false, "ref matrix from vartrix"},
    {"alt_matrix", 'a', false, false, "alt matrix from vartrix"},
    {"barcodes", 'b', false, false, "cell barcodes"},
    {"num_clusters", 'k', false, false, "number of clusters"},
    {"min_alt", 0, false, false, "minimum number of cells containing the alt allele for the variant to be used for clustering"},
    {"min_ref", 0, false, false, "minimum number of cells containing the ref allele for the variant to be used for clustering"},
    {"restarts", 0, false, false, "number of random seedings"},
    {"known_genotypes", 'g', false, false, "population vcf/bcf of known genotypes if available."},
    {"known_genotypes_sample_names", 0, true, false, "sample names, must be samples from the known_genotypes vcf"},
    {"known_cell_assignments", 0, false, false, "tsv with barcodes and their known assignments"},
    {"initialization_strategy", 0, false, false, "cluster initialization strategy, defaults to random_uniform, other methods not yet implemented"},
    {"threads", 't', false, false, "number of threads to use"},
    {"seed", 0, false, false, "optional random seed"},
    {"min_alt_umis", 0, false, false, "min alt umis to use locus for clustering"},
    {"min_ref_umis", 0, false, false, "min ref umis to use locus for clustering"},
    {"souporcell3", 's', false, false, "activate bad cluster detection and multiple runs"},
    {"clustering_method", 'm', false, false, "choose between expectation maximization (em) and k harmonic means (khm)"},
    {"gpus", 0, false, false, "[soupgpu] number of GPUs to shard the job over (default: as many of the visible ones as the job can keep busy)"},
    {"shard", 0, false, false, "[soupgpu] what the GPUs share out: restarts, cells (NCCL all-reduce of the M-step sums), or auto (default)"},
    {"max_slots", 0, false, false, "[soupgpu] resident solves per GPU (default: auto)"},
    {"quiet_iterations", 0, false, true, "[soupgpu] do not log per-iteration lines"},
    {"csr_cache", 0, false, false, "[soupgpu] .soupcsr cache of the parsed matrix (read if present, else written)"},
};

// clap 2's help layout (what `souporcell -h` prints, README.md:252-297 of the reference): options sorted by long name, the help text
// in one column 4 blanks after the longest `--name <name>` — unless that leaves it less room than it needs on a 120-column line, in
// which case it moves to the next line (12 blanks) followed by an empty line.  The soupgpu extras follow in a section of their own.
void usage(FILE* f) {
    fprintf(f, "souporcell 3.0\nHaynes Heaton <whheaton@gmail.com>\nclustering scRNAseq cells by genotype\n\nUSAGE:\n"
               "    souporcell [OPTIONS] --alt_matrix <alt_matrix> --barcodes <barcodes> --num_clusters <num_clusters> --ref_matrix <ref_matrix>\n\n"
               "FLAGS:\n    -h, --help       Prints help information\n    -V, --version    Prints version information\n\nOPTIONS:\n");
    auto left = [](const Opt& o) {
        std::string s = std::string("--") + o.name;
        if (!o.flag) s += std::string(" <") + o.name + ">" + (o.multiple ? "..." : "");
        return s;
    };
    auto extra = [](const Opt& o) { return strncmp(o.help, "[soupgpu]", 9) == 0; };
    for (int section = 0; section < 2; ++section) {
        std::vector<const Opt*> opts;
        size_t longest = 0;
        for (const Opt& o : kOpts)
            if (extra(o) == (section == 1)) { opts.push_back(&o); longest = std::max(longest, left(o).size()); }
        std::sort(opts.begin(), opts.end(), [](const Opt* a, const Opt* b) { return strcmp(a->name, b->name) < 0; });
        if (section == 1) fprintf(f, "\nSOUPGPU OPTIONS:\n");
        const size_t term = 120, taken = longest + 12;
        for (const Opt* o : opts) {
            const std::string l = left(*o);
            if (o->shrt) fprintf(f, "    -%c, %s", o->shrt, l.c_str());
            else fprintf(f, "        %s", l.c_str());
            const bool next_line = term >= taken && (float)taken / (float)term > 0.40f && strlen(o->help) > term - taken;
            if (next_line) fprintf(f, "\n            %s\n\n", o->help);
            else fprintf(f, "%*s%s\n", (int)(longest - l.size() + 4), "", o->help);
        }
    }
}

[[noreturn]] void die(const std::string& msg, int code = 1) {
    fprintf(stderr, "%s\n", msg.c_str());
    exit(code);
}
// mirrors `value.parse::<T>().unwrap()` panics (exit code 101)
uint64_t parse_num(const std::string& s, const char* what, uint64_t maxv) {
    size_t i = (!s.empty() && s[0] == '+') ? 1 : 0;
    if (i >= s.size()) die(std::string("thread 'main' panicked: invalid value for ") + what + ": '" + s + "'", 101);
    uint64_t v = 0;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9' || v > (maxv - (uint64_t)(s[i] - '0')) / 10) die(std::string("thread 'main' panicked: invalid value for ") + what + ": '" + s + "'", 101);
        v = v * 10 + (uint64_t)(s[i] - '0');
    }
    return v;
}

struct Stopwatch {
    bool on = getenv("SOUP_DEBUG_ITERS") != nullptr || getenv("SOUP_DEBUG_CLI") != nullptr;   // SOUP_DEBUG_CLI: stage times only, no device syncs
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        auto t1 = std::chrono::steady_clock::now();
        if (on) fprintf(stderr, "soup_debug cli %-24s %.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

}  // namespace

int main(int argc, char** argv) {
    Stopwatch sw;
    std::map<std::string, std::vector<std::string>> vals;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") { usage(stdout); return 0; }
        if (a == "-V" || a == "--version") { printf("souporcell 3.0\n"); return 0; }
        const Opt* o = nullptr;
        std::string inl;
        bool has_inl = false;
        if (a.rfind("--", 0) == 0) {
            std::string name = a.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) { inl = name.substr(eq + 1); name = name.substr(0, eq); has_inl = true; }
            for (const Opt& c : kOpts) if (name == c.name) o = &c;
        } else if (a.size() >= 2 && a[0] == '-') {
            for (const Opt& c : kOpts) if (c.shrt && c.shrt == a[1]) o = &c;
            if (a.size() > 2) { inl = a.substr(a[2] == '=' ? 3 : 2); has_inl = true; }
        }
        if (!o) { fprintf(stderr, "error: Found argument '%s' which wasn't expected, or isn't valid in this context\n\n", a.c_str()); usage(stderr); return 1; }
        if (o->flag) { vals[o->name].push_back("true"); continue; }
        if (has_inl) { vals[o->name].push_back(inl); continue; }
        if (i + 1 >= argc) { fprintf(stderr, "error: The argument '--%s <%s>' requires a value but none was supplied\n\n", o->name, o->name); usage(stderr); return 1; }
        vals[o->name].push_back(argv[++i]);
        if (o->multiple) while (i + 1 < argc && argv[i + 1][0] != '-') vals[o->name].push_back(argv[++i]);
    }
    {   // clap lists every missing required argument, then the usage
        std::string missing;
        for (const char* req : {"alt_matrix", "barcodes", "num_clusters", "ref_matrix"})
            if (!vals.count(req)) missing += std::string("    --") + req + " <" + req + ">\n";
        if (!missing.empty()) { fprintf(stderr, "error: The following required arguments were not provided:\n%s\n", missing.c_str()); usage(stderr); return 1; }
    }
    auto get = [&](const char* n, const char* d) { return vals.count(n) ? vals[n].back() : std::string(d); };
    // stdout carries the TSV (S:133-147) and nothing else: the table is written through a private duplicate of it, and descriptor 1 is
    // pointed at stderr for the rest of the process, so whatever a library prints there (NCCL's version / warning lines) cannot reach it
    fflush(stdout);
    FILE* tsv_out = nullptr;
    {
        const int fd = dup(STDOUT_FILENO);
        if (fd >= 0) tsv_out = fdopen(fd, "w");
        if (!tsv_out) die("thread 'main' panicked: failed printing to stdout", 101);
        dup2(STDERR_FILENO, STDOUT_FILENO);
    }

    // ---- load_params S:756-853
    const std::string ref_mtx = get("ref_matrix", ""), alt_mtx = get("alt_matrix", ""), barcodes_path = get("barcodes", "");
    const uint64_t k = parse_num(get("num_clusters", ""), "num_clusters", INT32_MAX);
    const uint32_t min_alt = (uint32_t)parse_num(get("min_alt", "4"), "min_alt", UINT32_MAX);
    const uint32_t min_ref = (uint32_t)parse_num(get("min_ref", "4"), "min_ref", UINT32_MAX);
    const uint32_t restarts = (uint32_t)parse_num(get("restarts", "100"), "restarts", UINT32_MAX);
    const bool known_cells = vals.count("known_cell_assignments") > 0, known_gt = vals.count("known_genotypes") > 0;
    if (known_gt && known_cells) die("thread 'main' panicked: Cannot set both known_genotypes and known_cell_assignments", 101);
    const std::string init = get("initialization_strategy", "random_uniform");
    if (init != "kmeans++" && init != "random_uniform" && init != "random_cell_assignment" && init != "middle_variance")
        die("thread 'main' panicked: initialization strategy must be one of kmeans++, random_uniform, random_cell_assignment, middle_variance", 101);
    const uint64_t threads = parse_num(get("threads", "1"), "threads", 1u << 20);
    const uint8_t seed = (uint8_t)parse_num(get("seed", "4"), "seed", 255);
    const uint32_t min_ref_umis = (uint32_t)parse_num(get("min_ref_umis", "0"), "min_ref_umis", UINT32_MAX);
    const uint32_t min_alt_umis = (uint32_t)parse_num(get("min_alt_umis", "0"), "min_alt_umis", UINT32_MAX);
    const std::string method_s = get("clustering_method", "em");
    if (method_s != "em" && method_s != "khm") die("thread 'main' panicked: Clustering method must be one of em, khm", 101);
    const std::string s3_s = get("souporcell3", "false");
    if (s3_s != "true" && s3_s != "false") die("thread 'main' panicked: provided string was not `true` or `false`", 101);
    const bool s3 = s3_s == "true";
    if (k > 16 && (method_s == "em" || !s3))   // S:831-833
        fprintf(stderr, "For k > 16, using souporcell3 (with '-s true' flag) is recommended with khm clustering method (with '-m khm' flag)\n");
    if (k == 0 || threads == 0) die("thread 'main' panicked: num_clusters and threads must be positive (the reference divides by them)", 101);
    // initialisation paths: S:485-498
    if (known_cells) die("thread 'main' panicked: known cell assignments not yet implemented", 101);   // S:545
    if (init == "kmeans++") die("thread 'main' panicked: kmeans++ not yet implemented", 101);           // S:550
    if (init == "middle_variance") die("thread 'main' panicked: middle variance not yet implemented", 101);   // S:597

    // The CUDA driver context takes a few hundred ms to come up: start it now, in the background, while the text
    // matrices are parsed on the CPU.
    // Only device 0 here: most jobs need one GPU (below), and every further context costs its own start-up.
    std::future<int> cuda_up = std::async(std::launch::async, [] {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
        if (n > 0) { cudaSetDevice(0); cudaFree(nullptr); }
        return n;
    });

    // ---- load_barcodes S:707-718, load_cell_data S:601-681
    char* bc_blob = nullptr;
    uint64_t n_bc = 0;
    if (soup_barcodes_load(barcodes_path.c_str(), &bc_blob, &n_bc) != SOUP_OK) die(std::string("thread 'main' panicked: ") + soup_last_error(nullptr), 101);
    if (n_bc > 200000)
        fprintf(stderr, "More than 200_000 barcodes in barcodes file? is this the filtered barcode file? (do not use raw barcode file)\n");
    soup_mtx* m = nullptr;
    const std::string cache = get("csr_cache", "");
    bool from_cache = false;
    if (!cache.empty()) {
        FILE* probe = fopen(cache.c_str(), "rb");
        if (probe) {
            fclose(probe);
            if (soup_mtx_load_cache(cache.c_str(), alt_mtx.c_str(), ref_mtx.c_str(), min_alt, min_ref, min_alt_umis, min_ref_umis, 0, &m) == SOUP_OK) from_cache = true;
            else fprintf(stderr, "csr_cache ignored: %s\n", soup_last_error(nullptr));
        }
    }
    if (!from_cache) {
        if (soup_mtx_load(alt_mtx.c_str(), ref_mtx.c_str(), min_alt, min_ref, min_alt_umis, min_ref_umis, 0, &m) != SOUP_OK)
            die(std::string("thread 'main' panicked: ") + soup_last_error(nullptr), 101);
        if (!cache.empty() && soup_mtx_save(m, cache.c_str(), alt_mtx.c_str(), ref_mtx.c_str(), min_alt, min_ref, min_alt_umis, min_ref_umis, 0) != SOUP_OK)
            fprintf(stderr, "csr_cache not written: %s\n", soup_last_error(nullptr));
    }
    fprintf(stderr, "total loci used %llu\n", (unsigned long long)m->n_loci);   // S:678
    sw.lap("parse mtx + barcodes");

    // ---- GPUs
    int visible = cuda_up.get();
    sw.lap("wait for CUDA context");
    if (visible == 0) die("souporcell (soupgpu): no CUDA device available; this build has no CPU fallback", 101);
    const uint64_t spt = (restarts + threads - 1) / threads;
    int n_gpu;
    if (vals.count("gpus")) {
        n_gpu = (int)parse_num(get("gpus", "1"), "gpus", 1024);
        if (n_gpu < 1 || n_gpu > visible) die("souporcell (soupgpu): --gpus out of range", 1);
    } else {
        // default: as many of the visible GPUs as the job can keep busy.  A context takes 1-2 s to come up and a replica of the
        // matrix; one GPU sustains ~1.3e12 (entry x cluster) updates per second, a solve runs ~17 iterations, souporcell3 up to 3 runs
        const double est_s = (double)m->nnz * (double)(threads * spt) * (double)k * 17.0 * (s3 ? 3.0 : 1.0) / 1.3e12;
        n_gpu = (int)std::min<double>((double)visible, std::max(1.0, std::ceil(est_s / 0.5)));
    }
    // restarts: solve s runs on GPU s mod n (every GPU holds the whole matrix, no collective).  cells: GPU g holds the g-th contiguous
    // range of the cells (balanced by entries), every GPU runs every restart and NCCL all-reduces the M-step sums each iteration.
    // auto: restarts while there are at least as many solves as GPUs, cells otherwise (and for random_cell_assignment never: it needs
    // every cell on one GPU).
    const std::string shard_s = get("shard", "auto");
    if (shard_s != "auto" && shard_s != "restarts" && shard_s != "cells") die("souporcell (soupgpu): --shard must be one of auto, restarts, cells", 1);
    bool shard_cells = shard_s == "cells" || (shard_s == "auto" && n_gpu > 1 && (uint64_t)n_gpu > threads * spt && init != "random_cell_assignment");
    if (shard_cells && init == "random_cell_assignment") die("souporcell (soupgpu): --shard cells is not available with random_cell_assignment", 1);
    if (!shard_cells && (uint64_t)n_gpu > threads * spt) n_gpu = (int)std::max<uint64_t>(1, threads * spt);
    if (shard_cells && (uint64_t)n_gpu > std::max<uint64_t>(m->n_cells, 1)) n_gpu = (int)std::max<uint64_t>(m->n_cells, 1);
    if (n_gpu == 1) shard_cells = false;
    std::vector<soup_ctx*> ctxs(n_gpu, nullptr);
    std::vector<soup_comm*> comms;
    {   // one thread per GPU: context creation, upload and the GPU transpose of the replicas / shards run side by side
        std::vector<std::string> errs(n_gpu);
        std::vector<uint64_t> cut(n_gpu + 1, 0);
        std::vector<std::vector<uint64_t>> shard_ptr(n_gpu);
        if (shard_cells) {
            for (int g = 1; g < n_gpu; ++g)
                cut[g] = std::max<uint64_t>(cut[g - 1], (uint64_t)(std::lower_bound(m->row_ptr, m->row_ptr + m->n_cells + 1, m->nnz * (uint64_t)g / (uint64_t)n_gpu) - m->row_ptr));
            cut[n_gpu] = m->n_cells;
            for (int g = 0; g < n_gpu; ++g) {
                cut[g] = std::min(cut[g], m->n_cells);
                shard_ptr[g].resize(cut[g + 1] - cut[g] + 1);
                for (uint64_t c = cut[g]; c <= cut[g + 1]; ++c) shard_ptr[g][c - cut[g]] = m->row_ptr[c] - m->row_ptr[cut[g]];
            }
        }
        auto bring_up = [&](int g) {
            if (soup_ctx_create(g, nullptr, &ctxs[g]) != SOUP_OK) { errs[g] = soup_last_error(nullptr); return; }
            int32_t rc;
            if (shard_cells) {
                const uint64_t e0 = m->row_ptr[cut[g]];
                rc = soup_load_csr(ctxs[g], cut[g + 1] - cut[g], m->n_loci, m->row_ptr[cut[g + 1]] - e0, shard_ptr[g].data(), m->locus + e0, m->alt + e0, m->ref + e0);
            } else {
                rc = soup_load_csr(ctxs[g], m->n_cells, m->n_loci, m->nnz, m->row_ptr, m->locus, m->alt, m->ref);
            }
            if (rc != SOUP_OK) errs[g] = soup_last_error(ctxs[g]);
        };
        std::vector<std::thread> pool;
        for (int g = 1; g < n_gpu; ++g) pool.emplace_back(bring_up, g);
        bring_up(0);
        for (auto& t : pool) t.join();
        for (int g = 0; g < n_gpu; ++g)
            if (!errs[g].empty()) die("souporcell (soupgpu): " + errs[g], 101);
        if (shard_cells) {
            comms.assign(n_gpu, nullptr);
            if (soup_comm_create_all(ctxs.data(), n_gpu, comms.data()) != SOUP_OK) die(std::string("souporcell (soupgpu): ") + soup_last_error(nullptr), 101);
        }
    }

    sw.lap("upload + GPU transpose");
    // ---- init_cluster_centers_known_genotypes S:514-542
    std::vector<float> known_theta;
    if (known_gt) {
        std::string blob;
        int n_names = 0;
        if (vals.count("known_genotypes_sample_names"))
            for (const std::string& nm : vals["known_genotypes_sample_names"]) { blob += nm; blob.push_back('\0'); ++n_names; }
        known_theta.resize((size_t)k * m->n_loci);
        if (soup_known_genotypes_load(get("known_genotypes", "").c_str(), blob.c_str(), n_names, (int32_t)k, m->index_to_locus, m->n_loci, known_theta.data()) != SOUP_OK)
            die(std::string("thread 'main' panicked: ") + soup_last_error(nullptr), 101);
    }

    // ---- souporcell_main S:60-148
    soup_main_params mp{};
    mp.known_theta = known_gt ? known_theta.data() : nullptr;
    mp.init_strategy = init == "random_cell_assignment" ? SOUP_INIT_RANDOM_ASSIGNMENT : SOUP_INIT_RANDOM_UNIFORM;   // S:491-496
    mp.k = (int32_t)k; mp.method = method_s == "khm" ? SOUP_METHOD_KHM : SOUP_METHOD_EM;
    mp.restarts = restarts; mp.threads = (uint32_t)threads; mp.seed = seed; mp.souporcell3 = s3 ? 1 : 0;
    mp.max_slots = (int32_t)parse_num(get("max_slots", "0"), "max_slots", 1024);
    mp.cell_comms = shard_cells ? comms.data() : nullptr;
    const bool quiet = vals.count("quiet_iterations") > 0;
    mp.record_counts = quiet ? 0 : 1; mp.verbose_trace = quiet ? 0 : 1;
    std::vector<float> best_lp((size_t)m->n_cells * k);
    soup_main_result res{};
    res.best_log_probs = best_lp.data();
    if (soup_main(ctxs.data(), n_gpu, &mp, &res, stderr) != SOUP_OK) die(std::string("thread 'main' panicked: ") + soup_last_error(nullptr), 101);
    sw.lap("soup_main");
    fprintf(stderr, "soupgpu\tgpus\t%d\tshard\t%s\tsolve_ms\t%.3f\tnnz_updates\t%llu\n", n_gpu, shard_cells ? "cells" : "restarts", res.solve_ms, (unsigned long long)res.nnz_updates);
    // zip(barcodes, best_log_probabilities): an empty best (nothing beat -inf) prints nothing (S:133)
    if (res.has_best && soup_write_clusters(tsv_out, bc_blob, n_bc, best_lp.data(), m->n_cells, (int32_t)k) != SOUP_OK)
        die(std::string("thread 'main' panicked: ") + soup_last_error(nullptr), 101);
    if (fflush(tsv_out) != 0) die("thread 'main' panicked: failed printing to stdout", 101);
    sw.lap("write TSV");
    for (soup_comm* c : comms) soup_comm_destroy(c);
    for (soup_ctx* c : ctxs) soup_ctx_destroy(c);
    soup_mtx_free(m);
    soup_free(bc_blob);
    return 0;
}
