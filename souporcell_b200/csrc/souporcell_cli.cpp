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

#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <new>

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

// ---- GPUs visible to this process, found WITHOUT touching the CUDA driver: the tokens CUDA_VISIBLE_DEVICES would take.
// (cuInit brings up every visible GPU — 5.7 s on an 8 x B200 box against 0.6 s for one device — so a process that needs one GPU, or one
// rank per GPU, narrows CUDA_VISIBLE_DEVICES before its first CUDA call.)
std::vector<std::string> visible_gpu_tokens() {
    std::vector<std::string> out;
    if (const char* e = getenv("CUDA_VISIBLE_DEVICES")) {
        std::string cur;
        for (const char* p = e;; ++p) {
            if (*p == ',' || *p == 0) {
                if (cur.empty() || cur == "-1") break;          // CUDA stops at the first invalid entry
                out.push_back(cur);
                cur.clear();
                if (*p == 0) break;
            } else if (*p != ' ') cur.push_back(*p);
        }
        return out;
    }
    int n = 0;
    for (int i = 0; i < 1024; ++i) {
        char path[64];
        snprintf(path, sizeof path, "/dev/nvidia%d", i);
        if (access(path, F_OK) == 0) ++n;
    }
    for (int i = 0; i < n; ++i) out.push_back(std::to_string(i));
    return out;
}

// ---- one rank per GPU as PROCESSES (restart sharding): the ranks meet once per run through this block of shared memory
struct ShmCtl {
    std::atomic<uint32_t> count, gen, abort;
    uint32_t world;
    alignas(64) unsigned char slots[64][64];   // soup_allgather_fn: <= 64 bytes per rank
    double solve_ms[64];
    unsigned long long nnz_updates[64];
    unsigned char nccl_id[4][SOUP_COMM_ID_BYTES];   // cell sharding: rank 0's ncclUniqueId of every cell group
};
struct ProcGroup {
    ShmCtl* ctl = nullptr;
    unsigned char* xfer = nullptr;             // soup_bcast_fn: the winner's log-probabilities / centers
    size_t xfer_bytes = 0;
    int rank = 0, world = 1;
    std::vector<pid_t> children;               // rank 0 only
    bool wait() {                              // barrier; false = a peer failed
        ShmCtl& c = *ctl;
        const uint32_t g = c.gen.load();
        if (c.count.fetch_add(1) + 1 == (uint32_t)world) { c.count.store(0); c.gen.fetch_add(1); return !c.abort.load(); }
        for (uint64_t spins = 0; c.gen.load() == g; ++spins) {
            if (c.abort.load()) return false;
            if ((spins & 1023) == 1023) {      // every ~50 ms: is everybody still alive?
                if (rank != 0 && getppid() == 1) { c.abort.store(1); return false; }
                if (rank == 0)
                    for (pid_t pid : children) {
                        int st = 0;
                        if (waitpid(pid, &st, WNOHANG) == pid && !(WIFEXITED(st) && WEXITSTATUS(st) == 0)) { c.abort.store(1); return false; }
                    }
            }
            usleep(50);
        }
        return !c.abort.load();
    }
};
int32_t shm_allgather(void* user, const void* send, void* recv, uint64_t bytes) {
    ProcGroup& g = *(ProcGroup*)user;
    if (bytes > 64) return SOUP_ERR_INVALID;
    memcpy(g.ctl->slots[g.rank], send, bytes);
    if (!g.wait()) return SOUP_ERR_INVALID;
    for (int q = 0; q < g.world; ++q) memcpy((char*)recv + (size_t)q * bytes, g.ctl->slots[q], bytes);
    return g.wait() ? SOUP_OK : SOUP_ERR_INVALID;
}
int32_t shm_bcast(void* user, void* buf, uint64_t bytes, int32_t root) {
    ProcGroup& g = *(ProcGroup*)user;
    if (bytes > g.xfer_bytes) return SOUP_ERR_INVALID;
    if (g.rank == root) memcpy(g.xfer, buf, bytes);
    if (!g.wait()) return SOUP_ERR_INVALID;
    if (g.rank != root) memcpy(buf, g.xfer, bytes);
    return g.wait() ? SOUP_OK : SOUP_ERR_INVALID;
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

    // The CUDA driver takes 0.6 s to bring up one GPU and 5.7 s to bring up eight (cuInit initialises every visible device), so
    //  * a process that will use ONE GPU narrows CUDA_VISIBLE_DEVICES to it first, and starts the driver now, in the background, while
    //    the text matrices are parsed on the CPU;
    //  * when several GPUs are visible and the job may want them, the driver stays untouched until the matrix is parsed and the number of
    //    GPUs is decided: restart sharding then runs one PROCESS per GPU (each sees one device), cell sharding one thread per GPU.
    const std::vector<std::string> gpu_tokens = visible_gpu_tokens();
    const bool gpus_given = vals.count("gpus") > 0;
    const int gpus_arg = gpus_given ? (int)parse_num(get("gpus", "1"), "gpus", 1024) : 0;
    // ("for sure" = the GPUs this process will use are known before the parse: one visible device, or --gpus given — then exactly those
    //  are made visible and the driver starts right away)
    // Without --gpus the number of GPUs follows from the size of the job (below), and the size of the job from the size of the alt matrix:
    // a vartrix line is ~12-13 bytes, so the file length gives the entry count well enough for a sqrt rule.
    const uint64_t spt0 = (restarts + threads - 1) / threads;
    auto gpus_for = [&](double nnz_est) {
        // The number of GPUs that minimises the wall time of the PROCESS.  One GPU sustains ~1.3e12 (entry x cluster) updates per second, a
        // solve runs ~17 iterations, the souporcell3 re-runs add about as much again (they hold only the re-drawn clusters); every GPU the
        // process touches costs about a second of CUDA driver bring-up, upload and teardown (measured on 2- and 8-GPU B200 boxes).
        // wall(n) = est / n + n * 1 s  is smallest at n = sqrt(est): config 4 (14 s of solves) -> 4 GPUs, config 2 (0.03 s) -> 1.
        const double est_s = nnz_est * (double)(threads * spt0) * (double)k * 17.0 * (s3 ? 2.0 : 1.0) / 1.3e12;
        return (int)std::max(1.0, std::floor(std::sqrt(est_s) + 0.5));
    };
    int gpus_guess = 0;
    if (!gpus_given && gpu_tokens.size() > 1 && !getenv("SOUP_CLI_PROCESSES")) {
        double bytes = 0;
        if (FILE* f = fopen(alt_mtx.c_str(), "rb")) { if (fseek(f, 0, SEEK_END) == 0) bytes = (double)ftell(f); fclose(f); }
        const bool gz = alt_mtx.size() > 3 && alt_mtx.compare(alt_mtx.size() - 3, 3, ".gz") == 0;
        gpus_guess = std::min<int>((int)gpu_tokens.size(), gpus_for(bytes / 13.0 * (gz ? 5.0 : 1.0)));
    }
    const int gpus_early = gpus_given ? gpus_arg : gpus_guess;     // 0 = not known yet
    const bool one_gpu_for_sure = gpu_tokens.size() <= 1 || (gpus_early >= 1 && !getenv("SOUP_CLI_PROCESSES"));
    if (one_gpu_for_sure && gpus_early >= 1 && (size_t)gpus_early < gpu_tokens.size()) {
        std::string use;
        for (int g = 0; g < gpus_early; ++g) use += (g ? "," : "") + gpu_tokens[g];
        setenv("CUDA_VISIBLE_DEVICES", use.c_str(), 1);
    }
    std::future<int> cuda_up;
    auto start_cuda = [&] {
        cuda_up = std::async(std::launch::async, [] {
            int n = 0;
            if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
            if (n > 0) { cudaSetDevice(0); cudaFree(nullptr); }
            return n;
        });
    };
    if (one_gpu_for_sure) start_cuda();

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
    int visible = one_gpu_for_sure ? cuda_up.get() : (int)gpu_tokens.size();
    if (one_gpu_for_sure) sw.lap("wait for CUDA context");
    if (visible == 0) die("souporcell (soupgpu): no CUDA device available; this build has no CPU fallback", 101);
    const uint64_t spt = (restarts + threads - 1) / threads;
    int n_gpu;
    if (gpus_given) {
        n_gpu = gpus_arg;
        if (n_gpu < 1 || n_gpu > visible) die("souporcell (soupgpu): --gpus out of range", 1);
    } else {
        n_gpu = std::min(visible, gpus_for((double)m->nnz));   // (`visible` is what the pre-estimate above left visible)
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

    // ---- init_cluster_centers_known_genotypes S:514-542 (host only; before the ranks split)
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

    // ---- several GPUs: one process per GPU.  The parsed matrix is inherited by fork() (copy-on-write, never written), each rank sees
    // ONE device — its driver comes up in the time one GPU takes, all ranks in parallel.  Restart sharding: the ranks meet once per run
    // in shared memory (soup_main's allgather / bcast callbacks).  Cell sharding: rank 0's NCCL id travels through the same block, the
    // ranks form one communicator (soup_comm_create) and every rank's rows of the result are collected there at the end.
    // Rank 0 is this process and prints the TSV.  (SOUP_CLI_ONE_PROCESS keeps everything in one process: one thread per GPU.)
    ProcGroup pg;
    // Measured (8 x B200, config 4): one process with a thread per GPU comes up in 5.7 s (cuInit over the 8 devices) + 1.3 s (upload);
    // eight processes that each see one GPU take 9 s for the same step — the driver serialises their bring-up — so the process-per-GPU
    // model is opt-in (SOUP_CLI_PROCESSES=1) and the default is one process that sees exactly the GPUs it uses.
    const bool multi_proc = !one_gpu_for_sure && n_gpu > 1 && n_gpu <= 64 && getenv("SOUP_CLI_PROCESSES") != nullptr;
    if (multi_proc) {
        pg.world = n_gpu;
        pg.xfer_bytes = std::max<size_t>((size_t)m->n_cells * k, (size_t)k * m->n_loci) * sizeof(float);
        void* map = mmap(nullptr, sizeof(ShmCtl) + pg.xfer_bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
        if (map == MAP_FAILED) die("souporcell (soupgpu): cannot map the ranks' shared memory", 101);
        pg.ctl = new (map) ShmCtl();
        pg.ctl->count.store(0); pg.ctl->gen.store(0); pg.ctl->abort.store(0);
        pg.ctl->world = (uint32_t)n_gpu;
        pg.xfer = (unsigned char*)map + sizeof(ShmCtl);
        fflush(stderr);
        for (int g = 1; g < n_gpu; ++g) {
            const pid_t pid = fork();
            if (pid < 0) { pg.ctl->abort.store(1); die("souporcell (soupgpu): fork failed", 101); }
            if (pid == 0) { pg.rank = g; pg.children.clear(); sw.on = false; break; }
            pg.children.push_back(pid);
        }
        setenv("CUDA_VISIBLE_DEVICES", gpu_tokens[pg.rank].c_str(), 1);
    } else if (!one_gpu_for_sure && (size_t)n_gpu < gpu_tokens.size()) {
        std::string use;                           // the driver brings up every VISIBLE device: show it the ones this job uses
        for (int g = 0; g < n_gpu; ++g) use += (g ? "," : "") + gpu_tokens[g];
        setenv("CUDA_VISIBLE_DEVICES", use.c_str(), 1);
    }
    auto fail = [&](const std::string& msg) {     // a rank that fails takes the others down with it
        if (pg.ctl) pg.ctl->abort.store(1);
        die(msg, 101);
    };
    // Cell sharding runs the restarts as `cell_groups` groups on the SAME GPUs (each group every GPU, its own communicator, its share of
    // the restarts): while one group's M-step sums are on the NVLinks, the other's E / M passes have the SMs.  SOUP_CELL_GROUPS overrides.
    int cell_groups = 1;
    if (shard_cells) {
        cell_groups = threads * spt >= 2 ? 2 : 1;
        if (const char* e = getenv("SOUP_CELL_GROUPS")) cell_groups = std::max(1, std::min(4, atoi(e)));
        cell_groups = (int)std::min<uint64_t>((uint64_t)cell_groups, std::max<uint64_t>(1, threads * spt));
    }
    const int gpus_here = multi_proc ? 1 : n_gpu;  // GPUs of THIS process
    const int n_ctx = gpus_here * cell_groups;     // its contexts: group-major, context g * gpus_here + d on device d
    uint64_t cell_lo = 0, cell_hi = m->n_cells;    // cells whose rows this process's soup_main returns
    std::vector<soup_ctx*> ctxs(n_ctx, nullptr);
    std::vector<soup_comm*> comms;
    {   // one thread per GPU: context creation, upload and the GPU transpose of the replicas / shards run side by side
        std::vector<std::string> errs(n_ctx);
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
        auto bring_up = [&](int g) {               // context g of this process: device g % gpus_here, and (cell sharding) the shard of that GPU
            const int dev = g % gpus_here, sh = multi_proc ? pg.rank : dev;
            if (soup_ctx_create(dev, nullptr, &ctxs[g]) != SOUP_OK) { errs[g] = soup_last_error(nullptr); return; }
            int32_t rc;
            if (shard_cells) {
                const uint64_t e0 = m->row_ptr[cut[sh]];
                rc = soup_load_csr(ctxs[g], cut[sh + 1] - cut[sh], m->n_loci, m->row_ptr[cut[sh + 1]] - e0, shard_ptr[sh].data(), m->locus + e0, m->alt + e0, m->ref + e0);
            } else {
                rc = soup_load_csr(ctxs[g], m->n_cells, m->n_loci, m->nnz, m->row_ptr, m->locus, m->alt, m->ref);
            }
            if (rc != SOUP_OK) errs[g] = soup_last_error(ctxs[g]);
        };
        std::vector<std::thread> pool;
        for (int g = 1; g < n_ctx; ++g) pool.emplace_back(bring_up, g);
        bring_up(0);
        for (auto& t : pool) t.join();
        for (int g = 0; g < n_ctx; ++g)
            if (!errs[g].empty()) fail("souporcell (soupgpu): " + errs[g]);
        if (shard_cells && multi_proc) {           // one rank per process: ncclCommInitRank over the ids rank 0 publishes, one per cell group
            if (pg.rank == 0)
                for (int g = 0; g < cell_groups; ++g)
                    if (soup_comm_unique_id(pg.ctl->nccl_id[g]) != SOUP_OK) fail(std::string("souporcell (soupgpu): ") + soup_last_error(nullptr));
            if (!pg.wait()) fail("souporcell (soupgpu): another rank failed");
            comms.assign(cell_groups, nullptr);
            for (int g = 0; g < cell_groups; ++g)
                if (soup_comm_create(ctxs[g], pg.ctl->nccl_id[g], pg.rank, pg.world, &comms[g]) != SOUP_OK) fail(std::string("souporcell (soupgpu): ") + soup_last_error(nullptr));
        } else if (shard_cells) {
            comms.assign(n_ctx, nullptr);           // group 0: ncclCommInitAll; every further group: a split of it (no second bootstrap)
            if (soup_comm_create_all(ctxs.data(), n_gpu, comms.data()) != SOUP_OK) die(std::string("souporcell (soupgpu): ") + soup_last_error(nullptr), 101);
            for (int g = 1; g < cell_groups; ++g)
                if (soup_comm_dup_all(comms.data(), n_gpu, comms.data() + (size_t)g * n_gpu) != SOUP_OK &&
                    soup_comm_create_all(ctxs.data() + (size_t)g * n_gpu, n_gpu, comms.data() + (size_t)g * n_gpu) != SOUP_OK)
                    die(std::string("souporcell (soupgpu): ") + soup_last_error(nullptr), 101);
        }
        if (shard_cells && multi_proc) {           // where this rank's rows of the result go
            cell_lo = cut[pg.rank]; cell_hi = cut[pg.rank + 1];
        }
    }

    sw.lap("upload + GPU transpose");
    // ---- souporcell_main S:60-148
    soup_main_params mp{};
    mp.known_theta = known_gt ? known_theta.data() : nullptr;
    mp.init_strategy = init == "random_cell_assignment" ? SOUP_INIT_RANDOM_ASSIGNMENT : SOUP_INIT_RANDOM_UNIFORM;   // S:491-496
    mp.k = (int32_t)k; mp.method = method_s == "khm" ? SOUP_METHOD_KHM : SOUP_METHOD_EM;
    mp.restarts = restarts; mp.threads = (uint32_t)threads; mp.seed = seed; mp.souporcell3 = s3 ? 1 : 0;
    mp.max_slots = (int32_t)parse_num(get("max_slots", "0"), "max_slots", 1024);
    mp.cell_comms = shard_cells ? comms.data() : nullptr;
    mp.cell_groups = cell_groups;
    const bool quiet = vals.count("quiet_iterations") > 0;
    mp.record_counts = quiet ? 0 : 1; mp.verbose_trace = quiet ? 0 : 1;
    std::vector<float> best_lp((size_t)m->n_cells * k);
    soup_main_result res{};
    res.best_log_probs = best_lp.data();
    // every rank logs its own solves; the run-level lines (cluster analysis, best total) come from rank 0 alone
    char* log_buf = nullptr;
    size_t log_len = 0;
    FILE* log = pg.rank == 0 ? stderr : open_memstream(&log_buf, &log_len);
    if (multi_proc && !shard_cells) { mp.proc_rank = pg.rank; mp.proc_world = pg.world; mp.allgather = shm_allgather; mp.bcast = shm_bcast; mp.comm_user = &pg; }
    if (multi_proc && shard_cells) mp.n_cells_global = m->n_cells;
    if (soup_main(ctxs.data(), n_ctx, &mp, &res, log) != SOUP_OK) fail(std::string("thread 'main' panicked: ") + soup_last_error(nullptr));
    sw.lap("soup_main");
    double solve_ms = res.solve_ms;
    unsigned long long nnz_updates = res.nnz_updates;
    if (multi_proc) {
        pg.ctl->solve_ms[pg.rank] = res.solve_ms;
        pg.ctl->nnz_updates[pg.rank] = res.nnz_updates;
        if (shard_cells) {                         // every rank ran every solve on its cells: collect the rows, in cell order
            if (res.has_best) memcpy(pg.xfer + (size_t)cell_lo * k * sizeof(float), best_lp.data(), (size_t)(cell_hi - cell_lo) * k * sizeof(float));
            pg.ctl->nnz_updates[pg.rank] = res.nnz_updates;
        }
        if (pg.rank != 0 && !shard_cells) {
            fclose(log);
            std::string own;
            for (const char* p = log_buf; p && *p;) {
                const char* e = strchr(p, '\n');
                const size_t len = e ? (size_t)(e - p) + 1 : strlen(p);
                if (!strncmp(p, "thread ", 7) || !strncmp(p, "binomial", 8)) own.append(p, len);
                p += len;
            }
            fwrite(own.data(), 1, own.size(), stderr);
            fflush(stderr);
        }
        if (!pg.wait()) fail("souporcell (soupgpu): another rank failed");
        if (pg.rank != 0) _exit(0);               // (no teardown: the process ends here)
        solve_ms = 0; nnz_updates = 0;
        for (int q = 0; q < pg.world; ++q) { solve_ms = std::max(solve_ms, pg.ctl->solve_ms[q]); nnz_updates += pg.ctl->nnz_updates[q]; }
        if (shard_cells && res.has_best) memcpy(best_lp.data(), pg.xfer, (size_t)m->n_cells * k * sizeof(float));
    }
    fprintf(stderr, "soupgpu\tgpus\t%d\tshard\t%s\tsolve_ms\t%.3f\tnnz_updates\t%llu\n", n_gpu,
            shard_cells ? (cell_groups > 1 ? "cells (2+ groups)" : "cells") : "restarts", solve_ms, nnz_updates);
    // zip(barcodes, best_log_probabilities): an empty best (nothing beat -inf) prints nothing (S:133)
    if (res.has_best && soup_write_clusters(tsv_out, bc_blob, n_bc, best_lp.data(), m->n_cells, (int32_t)k) != SOUP_OK)
        fail(std::string("thread 'main' panicked: ") + soup_last_error(nullptr));
    if (fflush(tsv_out) != 0) fail("thread 'main' panicked: failed printing to stdout");
    sw.lap("write TSV");
    int rc = 0;
    for (pid_t pid : pg.children) {
        int st = 0;
        if (waitpid(pid, &st, 0) == pid && !(WIFEXITED(st) && WEXITSTATUS(st) == 0)) rc = 101;
    }
    if (getenv("SOUP_CLI_TEARDOWN")) {            // orderly release of every device object (leak checkers); the default is to end the
        for (soup_comm* c : comms) soup_comm_destroy(c);   // process right here: nothing is pending, and tearing down eight contexts with
        for (soup_ctx* c : ctxs) soup_ctx_destroy(c);      // gigabytes of buffers costs a second or two of wall time that buys nothing
        soup_mtx_free(m);
        soup_free(bc_blob);
        return rc;
    }
    fflush(stderr);
    _exit(rc);
}
