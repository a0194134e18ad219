/* soupgpu.h — C ABI of libsoupgpu, the B200-native (sm_100a) replacement for the
 * genotype-clustering hot path of wheaton5/souporcell.
 *
 * The reference has no FFI: its boundary is the `souporcell` process
 * (/root/reference/souporcell/src/main.rs, "S:" below; argv contract in
 * souporcell/src/params.yml).  Each entry point here names the reference function
 * it replaces so a Rust `-sys` crate can bind it 1:1.
 *
 * Conventions: every function returns int32_t status (SOUP_OK = 0, < 0 = error);
 * soup_last_error() gives the message.  No C++ exceptions cross the boundary.
 * The caller owns every host buffer; the library owns device memory behind the
 * opaque soup_ctx.  A ctx is single-threaded; distinct ctxs may be used from
 * distinct threads.  There is NO CPU fallback: device entry points fail with
 * SOUP_ERR_CUDA when no sm_100 GPU is usable.
 */
#ifndef SOUPGPU_H
#define SOUPGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOUP_OK 0
#define SOUP_ERR_INVALID (-1)     /* bad argument / malformed input (reference: panic!/assert!) */
#define SOUP_ERR_CUDA (-2)        /* CUDA runtime failure or no usable GPU */
#define SOUP_ERR_IO (-3)          /* cannot open / read a file */
#define SOUP_ERR_UNSUPPORTED (-4) /* valid in the reference but outside this build's limits */
#define SOUP_ERR_NOMEM (-5)

#define SOUP_METHOD_EM 0  /* S:189-278 */
#define SOUP_METHOD_KHM 1 /* S:280-358 */

typedef struct soup_ctx soup_ctx;

int32_t soup_version(void); /* 3000 = "souporcell 3.0" contract */
/* message of the last failure on `ctx` (or, with ctx == NULL, of the last failing
 * call on this thread that had no ctx).  Valid until the next call. */
const char* soup_last_error(const soup_ctx* ctx);

/* ------------------------------------------------------------------ host side (no GPU needed) */

/* Loader: replaces load_cell_data S:601-681 (parse both .mtx, per-locus tallies, the
 * min_ref/min_alt/min_*_umis filter S:659, dense re-index S:648-675, per-cell ascending-
 * locus vectors).  Output is a CSR over cells (columns = used-locus index, ascending),
 * entries with ref+alt == 0 dropped (S:664), duplicates resolved last-wins (S:633-634). */
typedef struct soup_mtx {
    uint64_t n_cells;          /* header column count (S:638) */
    uint64_t n_loci_total;     /* header row count (S:637) */
    uint64_t n_loci;           /* loci passing the filter = "total loci used" (S:678) */
    uint64_t nnz;              /* stored entries after filter */
    uint64_t* row_ptr;         /* [n_cells + 1] */
    uint32_t* locus;           /* [nnz] used-locus index */
    uint32_t* alt;             /* [nnz] */
    uint32_t* ref;             /* [nnz] */
    uint64_t* index_to_locus;  /* [n_loci] original 0-based locus id (S:649,661) */
    int32_t pinned;            /* buffers are cudaHostAlloc'ed */
} soup_mtx;
#define SOUP_MTX_PINNED 1 /* allocate the CSR in page-locked memory (needs a CUDA driver) */
#define SOUP_MTX_RAW 2    /* troublet's view of the files (T:283-334): every line kept (zeros and duplicates too), no locus filter or
                          * re-index (n_loci = header rows), file order inside a cell; the min_* arguments are ignored */
int32_t soup_mtx_load(const char* alt_path, const char* ref_path, uint32_t min_alt, uint32_t min_ref,
                      uint32_t min_alt_umis, uint32_t min_ref_umis, uint32_t flags, soup_mtx** out);
void soup_mtx_free(soup_mtx* m);
/* Binary cache of a loaded matrix (`.soupcsr`): soup_mtx_save writes what soup_mtx_load returned together with the
 * filter it was built with and a fingerprint (size, mtime, hash of both ends) of the two text matrices it was parsed from;
 * soup_mtx_load_cache reads it back (and fails with SOUP_ERR_INVALID if the filter / RAW flag differ, the cache was built from
 * other alt / ref files, or the file is not consistent), so re-runs, troublet and consensus skip the text parse.  The wire format of record stays the
 * vartrix text mtx; ".mtx.gz" inputs are accepted by soup_mtx_load as an extension (the reference opens them uncompressed, S:602-607). */
int32_t soup_mtx_save(const soup_mtx* m, const char* path, const char* alt_path, const char* ref_path, uint32_t min_alt, uint32_t min_ref,
                      uint32_t min_alt_umis, uint32_t min_ref_umis, uint32_t flags);
int32_t soup_mtx_load_cache(const char* path, const char* alt_path, const char* ref_path, uint32_t min_alt, uint32_t min_ref,
                            uint32_t min_alt_umis, uint32_t min_ref_umis, uint32_t flags, soup_mtx** out);

/* The inverse of soup_mtx_load for generators and benchmarks: one vartrix-style Matrix Market text file (header, `% comment`, `rows cols
 * nnz`, then `locus cell count` per entry, 1-based) from 0-based coordinates, formatted on all host cores. */
int32_t soup_mtx_write_text(const char* path, uint64_t n_loci, uint64_t n_cells, uint64_t nnz, const int64_t* locus, const int64_t* cell,
                            const uint32_t* count, const char* comment);

/* load_barcodes S:707-718 (+ reader S:500-511: ".gz" via zlib).  Returns one malloc'ed
 * buffer of NUL-separated barcodes; free with soup_free(). */
int32_t soup_barcodes_load(const char* path, char** blob, uint64_t* n_barcodes);
void soup_free(void* p);

/* init_cluster_centers_known_genotypes S:514-542: centers [k][n_loci] from a (plain or gzipped) VCF; record i is
 * original locus i (S:523-539).  `sample_names`: n_samples NUL-terminated strings back to back, sample s -> cluster s. */
int32_t soup_known_genotypes_load(const char* vcf_path, const char* sample_names, int32_t n_samples, int32_t k,
                                  const uint64_t* index_to_locus, uint64_t n_loci, float* theta /*[k][n_loci]*/);

/* rand 0.9.3 StdRng (= rand_chacha 0.9.0 ChaCha12, key = seed, 64-bit counter from 0):
 * `n` consecutive u32 words starting at absolute word `word_offset` (S:52,62,559,858). */
int32_t soup_chacha_words(const uint8_t seed[32], int32_t rounds, uint64_t word_offset, uint32_t* out, uint64_t n);
/* new_seed S:855-861 applied as souporcell_main does (S:61-62, 77-80): the 32-byte seeds of
 * the `threads` ThreadData of run `run` (0..2) for CLI seed `seed`.  out[threads][32]. */
int32_t soup_thread_seeds(uint8_t seed, uint32_t threads, uint32_t run, uint8_t* out);
/* Rust `{}` (width8 = 0) / `{:8}` (width8 = 1) formatting of an f32 (S:144, S:264). */
int32_t soup_format_f32(float v, int32_t width8, char* buf, int32_t buflen);
/* statrs 0.11.0 ln_binomial as used at S:669-670, already cast to f32. */
float soup_ln_binomial_f32(uint32_t n, uint32_t k);
/* bad_cluster_detection S:150-187.  best_log_probs [n_cells][k]; out_idx holds up to k
 * indices; returns the count (>= 0) or an error (< 0).  Writes the "Cluster Analysis" table
 * to `log_fd`'s FILE if log != NULL (a `FILE*`). */
int32_t soup_bad_cluster_detection(int32_t run, int32_t k, const float* best_log_probs, uint64_t n_cells,
                                   int32_t* out_idx, void* log_file);

/* ------------------------------------------------------------------ device side */

/* `cuda_stream` may be NULL (the ctx creates its own) or a cudaStream_t to enqueue on. */
int32_t soup_ctx_create(int32_t device, void* cuda_stream, soup_ctx** out);
void soup_ctx_destroy(soup_ctx* ctx);

/* Upload the filtered CSR (host pointers; what load_cell_data returns as Vec<CellData>,
 * S:663-674) and build on the GPU: one packed 32-bit word per entry (+ side records for large counts), the CSC transpose (cells
 * ascending inside a locus = the order update_centers_average S:476-483 accumulates in),
 * per-cell total_alleles (S:671, f32, ascending locus) and the ln_binomial table (S:669). */
int32_t soup_load_csr(soup_ctx* ctx, uint64_t n_cells, uint64_t n_loci, uint64_t nnz, const uint64_t* row_ptr,
                      const uint32_t* locus, const uint32_t* alt, const uint32_t* ref);
int32_t soup_get_dims(soup_ctx* ctx, uint64_t* n_cells, uint64_t* n_loci, uint64_t* nnz /* after dropping ref+alt == 0 */);
/* Test/inspection hooks: copy the device-built structures back (any pointer may be NULL). */
int32_t soup_get_csc(soup_ctx* ctx, uint64_t* col_ptr /*[L+1]*/, uint32_t* cell /*[nnz]*/, uint32_t* alt,
                     uint32_t* ref);
int32_t soup_get_cell_totals(soup_ctx* ctx, float* total_alleles /*[N]*/);
int32_t soup_get_csr_lbc(soup_ctx* ctx, float* lbc /*[nnz]: log_binomial_coefficient per stored entry, S:669*/);

/* Per-cluster, per-locus allele counts of hard cell assignments over the loaded matrix: the sufficient statistics
 * the downstream stages rebuild from the text matrices today — consensus.py:296-331 (`cluster_allele_counts[locus][cluster]
 * = [ref, alt]`, cells of `doublets` skipped) and troublet's load_allele_data (/root/reference/troublet/src/main.rs:261-412).
 * `cell_cluster[c]` < 0 (or >= k) skips cell c.  Exact integer sums (u64), one atomic add per stored entry; loci are the
 * used-locus indices of the loaded matrix (soup_mtx::index_to_locus maps them back to file rows). */
int32_t soup_cluster_allele_counts(soup_ctx* ctx, const int32_t* cell_cluster /*[n_cells]*/, int32_t k,
                                   uint64_t* counts /*[n_loci][k][2] = {ref, alt}*/);

/* Cell-sharded mode: sum `n_floats` f32 at device pointer `device_buf` in place over all ranks, enqueued on
 * `cuda_stream` (e.g. ncclAllReduce, or torch.distributed.all_reduce on a tensor aliasing the buffer).  Called once
 * per EM iteration with the M pass's sufficient statistics [sum p*alt | sum p*(alt+ref)] of all resident restarts and
 * their partial losses. */
typedef int32_t (*soup_allreduce_fn)(void* user, void* device_buf, uint64_t n_floats, void* cuda_stream);

/* The same all-reduce done by the library itself over NCCL (linked into libsoupgpu): a soup_comm is one rank's NCCL communicator bound to
 * a context's device.  The ranks of a communicator are either the contexts of ONE process (soup_comm_create_all = ncclCommInitAll; the
 * CLI's `--shard cells`) or one context per process (soup_comm_unique_id on one rank, the 128 bytes handed to every rank by the
 * launcher's own channel, then soup_comm_create = ncclCommInitRank).  With a communicator the statistics are reduced in locus
 * ranges while the M pass is still producing the next range, and only the columns of the live restarts travel.
 * The two host helpers move small host buffers over the communicator (staged through device memory): the once-per-run winner exchange
 * of restart sharding (S:110-125 across processes) and the per-cluster cell counts of souporcell3 (S:156-162). */
typedef struct soup_comm soup_comm;
#define SOUP_COMM_ID_BYTES 128
int32_t soup_comm_unique_id(uint8_t id[SOUP_COMM_ID_BYTES]);
int32_t soup_comm_create(soup_ctx* ctx, const uint8_t id[SOUP_COMM_ID_BYTES], int32_t rank, int32_t world, soup_comm** out);
int32_t soup_comm_create_all(soup_ctx* const* ctxs, int32_t n_ctx, soup_comm** out /*[n_ctx]*/);
/* a second communicator over the same ranks of this process (ncclCommSplit: no new bootstrap) — one per further cell group */
int32_t soup_comm_dup_all(soup_comm* const* comms, int32_t n, soup_comm** out /*[n]*/);
void soup_comm_destroy(soup_comm* comm);
int32_t soup_comm_rank(const soup_comm* comm, int32_t* rank, int32_t* world);
int32_t soup_comm_allgather(soup_comm* comm, const void* send, void* recv, uint64_t bytes_per_rank);   /* host buffers */
int32_t soup_comm_bcast(soup_comm* comm, void* buf, uint64_t bytes, int32_t root);                      /* host buffer */

/* cluster-center initialisation (ClusterInit S:749-754).  kmeans++ / middle_variance / known cell assignments are
 * `assert!(false)` in the reference (S:545,550,597) and are rejected by the CLI with the same message. */
#define SOUP_INIT_RANDOM_UNIFORM 0      /* init_cluster_centers_uniform S:554-563 (the default) */
#define SOUP_INIT_RANDOM_ASSIGNMENT 1   /* init_cluster_centers_random_assignment S:565-594 */

typedef struct soup_solve_params {
    int32_t k;                 /* num_clusters (S:726) */
    int32_t method;            /* SOUP_METHOD_EM | SOUP_METHOD_KHM (S:101-108) */
    int32_t n_streams;         /* `threads`: number of seed streams / ThreadData (S:77-80) */
    int32_t solves_per_stream; /* solves_per_thread (S:63); total solves = n_streams * this */
    const uint8_t* stream_seeds; /* [n_streams][32]: ThreadData rng seeds; used when theta0 == NULL */
    const float* theta0;       /* optional host [n_solves][k][L] initial centers, solve = stream*sps + iter */
    const float* init_theta;   /* optional host [k][L]: deterministic initial centers (known genotypes, S:487): draw d of a solve
                                * takes row d of it instead of the RNG stream, in every solve */
    const float* base_theta;   /* optional host [k][L]: best_cluster_centers of the previous run (S:93) */
    const int32_t* reinit_clusters; /* bad clusters re-drawn this run (S:92-96); the rest are locked (S:97) */
    int32_t n_reinit;          /* 0 = run 0: draw all k, lock nothing */
    int32_t iteration_cap;     /* 0 -> 1000 (S:222) */
    int32_t max_slots;         /* solves resident on the GPU at once; 0 = auto */
    int32_t lanes;             /* tests: force this many columns per lane (1, 2, 4, 8, 16) in every group and switch the run-time
                                * narrowing off; 0 = auto */
    int32_t record_counts;     /* also record clusters-with->200-cells per iteration (S:252-263) */
    int32_t keep_trace;        /* copy the per-iteration trace back (soup_get_trace) */
    int32_t poll_chunk;        /* iterations enqueued between host polls of the done counter; 0 = 4 */
    int32_t time_kernels;      /* bracket each iteration's E and M launches with CUDA events (soup_get_stats) */
    /* multi-GPU restart sharding: this process runs solves with (solve % world) == rank */
    int32_t shard_rank, shard_world; /* world 0 or 1 = no sharding */
    /* multi-GPU cell sharding: the loaded matrix holds a contiguous range of the cells (all loci); every rank runs
     * every restart and the ranks all-reduce the M-pass statistics each iteration.  best_log_probs then covers the
     * local cells only; results agree with the single-GPU path to tolerance (different summation order), not bitwise. */
    soup_allreduce_fn cell_allreduce; /* NULL = off */
    void* comm_user;
    uint64_t n_cells_global;          /* total cells over all ranks (the convergence limit S:214 uses it) */
    int32_t init_strategy;            /* SOUP_INIT_* : how a solve's centers are drawn from its seed stream when neither theta0
                                       * nor init_theta is given (--initialization_strategy, S:491-496) */
    soup_comm* cell_comm;             /* cell sharding over NCCL inside the library (instead of the cell_allreduce callback); NULL = off */
    int32_t cell_ranges;              /* cell sharding: locus ranges the M pass / all-reduce / theta pass are pipelined over; 0 = auto (4; the
                                       * SOUP_CELL_RANGES environment variable overrides both).  Must be the same on every rank. */
} soup_solve_params;

typedef struct soup_solve_result {
    int32_t has_best;      /* 0 if no solve beat -inf (reference prints an empty TSV) */
    int32_t best_solve;    /* index stream*sps + iteration of the winner (ties: lowest, S:110,120) */
    float best_loss;       /* best_log_probability (S:65,121) */
    float* best_log_probs; /* optional host [n_cells][k]  (S:122) */
    float* best_theta;     /* optional host [k][n_loci]   (S:123), after the winner's last M-step */
    float* solve_loss;     /* optional host [n_solves]; NaN for solves of other shards */
    int32_t* solve_iters;  /* optional host [n_solves] */
    uint64_t nnz_updates;  /* nnz * sum(iterations) over the solves this call ran */
} soup_solve_result;

/* Runs every solve of one `run` of souporcell_main (S:77-125): init (S:485-498, 554-563),
 * em()/khm() to convergence with all control flow on the device, best-of selection. */
int32_t soup_solve(soup_ctx* ctx, const soup_solve_params* p, soup_solve_result* r);

/* per-iteration stderr trace of one solve (S:264 / S:354) from the last soup_solve */
typedef struct soup_iter_record {
    int32_t temp_step;
    float loss;    /* log_binom_loss */
    float change;  /* log_loss_change */
    int32_t clusters_above_threshold; /* -1 unless record_counts */
} soup_iter_record;
int32_t soup_get_trace(soup_ctx* ctx, int32_t solve, soup_iter_record* out, int32_t cap, int32_t* n);

typedef struct soup_stats {
    uint64_t kernel_launches;   /* kernels launched by the last soup_solve / hook call */
    uint64_t estep_launches, mstep_launches;
    double solve_ms;            /* CUDA-event duration of the whole solve loop */
    double estep_ms, mstep_ms;  /* summed CUDA-event durations of the E / M launches (time_kernels = 1), else 0 */
    uint64_t timed_iterations;  /* launches of each kind covered by estep_ms / mstep_ms */
    uint64_t slot_iterations;   /* sum over launches of active solves (= sum of iterations) */
    int32_t slots, lanes, groups, padded_cols;
    uint64_t h2d_bytes, d2h_bytes; /* copies made by the last load / solve call */
    uint64_t allreduce_bytes;      /* cell sharding: bytes this rank handed to the all-reduce over the whole call */
    uint64_t allreduce_calls;      /* ... in this many calls (locus ranges x iterations) */
} soup_stats;
int32_t soup_get_stats(soup_ctx* ctx, soup_stats* out);

/* Fine-grained parity hooks: one E pass (binomial_loss_with_min_index S:409-426, log_sum_exp
 * S:428-432, normalize_in_log_with_temp S:443-454, KHM q S:360-384) and one M pass
 * (update_centers_average S:476-483 + update_final_with_lock S:386-396) for `n_slots`
 * independent center sets at once.  All pointers are host memory. */
int32_t soup_estep(soup_ctx* ctx, int32_t k, int32_t method, int32_t n_slots, int32_t lanes,
                   const float* theta /*[n_slots][k][L]*/, const int32_t* temp_step /*[n_slots]*/,
                   float* log_probs /*[n_slots][N][k]*/, float* weights /*[n_slots][N][k]*/,
                   float* lse /*[n_slots][N]*/, float* loss /*[n_slots]*/, int32_t* cells_per_cluster /*[n_slots][k]*/);
int32_t soup_mstep(soup_ctx* ctx, int32_t k, int32_t n_slots, int32_t lanes, const float* weights /*[n_slots][N][k]*/,
                   const float* theta_in /*[n_slots][k][L]: kept for locked clusters*/,
                   const int32_t* locked, int32_t n_locked, float* theta_out /*[n_slots][k][L]*/);

/* Device math self-test hook: applies the on-device ln / exp used by the kernels to n
 * host floats (parity check against the host libm the oracle uses). op 0 = ln, 1 = exp. */
int32_t soup_device_math(soup_ctx* ctx, int32_t op, const float* in, float* out, uint64_t n);

/* ------------------------------------------------------------------ troublet
 * The doublet caller that consumes clusters_tmp.tsv and the same two matrices: /root/reference/troublet/src/main.rs ("T:"). */

/* load_clusters T:414-449: `barcode \t cluster \t lp_0 .. lp_{k-1}` per line.  losses [n_cells][k] (f64), barcodes as one
 * NUL-separated blob; free both with soup_free().  n_clusters = max of column 2, plus one (T:426,448). */
int32_t soup_clusters_load(const char* path, char** barcode_blob, double** losses, uint64_t* n_cells, int32_t* k, int32_t* n_clusters);

typedef struct soup_troublet_params {
    double p_soup;             /* T:475-476: not a CLI flag of the reference, always 0.0 there */
    double doublet_prior;      /* T:477-478, default 0.5 */
    double doublet_threshold;  /* T:489-490, default 0.9 */
    double singlet_threshold;  /* T:491-492, default 0.9 */
} soup_troublet_params;
#define SOUP_TROUBLET_SINGLET 0
#define SOUP_TROUBLET_UNASSIGNED 1           /* best singlet beats best doublet, posterior below the threshold (T:183) */
#define SOUP_TROUBLET_DOUBLET 2
#define SOUP_TROUBLET_UNASSIGNED_DOUBLET 3   /* T:201 */
typedef struct soup_troublet_result {   /* caller-allocated, [n_cells] each unless noted */
    int32_t* status;
    int32_t* best_singlet;     /* T:83-86 */
    int32_t* doublet1;         /* T:127-131 */
    int32_t* doublet2;
    double* singlet_posterior; /* T:166 */
    double* doublet_posterior; /* T:167 */
    double* log_prob_singleton;
    double* log_prob_doublet;
    double* cluster_logprobs;  /* [n_cells][k] T:88-90 */
    int32_t rounds;            /* passes of the `while any_removed > 0` loop T:46 */
    uint64_t removed;          /* cells removed as doublets over all rounds */
    double device_ms;          /* CUDA-event time of the singlet + doublet kernels, all rounds */
} soup_troublet_result;
/* call_doublets T:32-244 + the count tables of load_allele_data T:261-412 on the GPU.  The matrix is the RAW one
 * (soup_mtx_load with SOUP_MTX_RAW: every line, no locus filter, file order inside a cell), `losses` the [n_cells][k]
 * log-probability columns of the clusters file.  Per-cell sums are accumulated in file order in f64 like the reference;
 * `ln` on the device is within 1 ulp of the host's, so log-probabilities agree to ~1e-13 relative, not bit for bit. */
int32_t soup_troublet(soup_ctx* ctx, uint64_t n_cells, uint64_t n_loci, uint64_t nnz, const uint64_t* row_ptr, const uint32_t* locus,
                      const uint32_t* alt, const uint32_t* ref, const double* losses, int32_t k, int32_t n_clusters,
                      const soup_troublet_params* p, soup_troublet_result* r, const char* barcode_blob, void* log_file);
/* the TSV troublet prints (T:36-41, T:233-243), Rust `{}` formatting of f64 */
int32_t soup_troublet_write(void* out_file, const char* barcode_blob, uint64_t n_cells, int32_t k, const soup_troublet_result* r);

/* ------------------------------------------------------------------ driver */

/* souporcell_main S:60-131 on top of soup_solve: seeds (S:61-63,77-80), up to 3 souporcell3
 * runs (S:70-76,127-129), bad-cluster detection and locking (S:72,92-97), global best-of
 * (S:119-125).  Restarts are independent (S:81-118), so they shard with no data-path
 * collective: solve s runs on shard (s mod world).  Inside one process the shards are the
 * `n_ctx` contexts (one per GPU, each holding the same matrix); across processes (one rank per
 * GPU, e.g. torchrun / MPI) the caller supplies two host-memory exchange callbacks that are
 * used once per run to agree on the winner (a few bytes all-gathered) and to fetch its
 * log-probabilities and centers from the owning rank.  `log_file` is a FILE* or NULL. */
typedef int32_t (*soup_allgather_fn)(void* user, const void* send, void* recv, uint64_t bytes_per_rank);
typedef int32_t (*soup_bcast_fn)(void* user, void* buf, uint64_t bytes, int32_t root);
typedef struct soup_main_params {
    int32_t k, method;
    uint32_t restarts, threads;
    uint8_t seed;
    int32_t souporcell3;
    int32_t iteration_cap; /* 0 -> 1000 */
    int32_t max_slots, lanes, record_counts, verbose_trace, time_kernels;
    const float* known_theta;        /* optional host [k][n_loci] from soup_known_genotypes_load (--known_genotypes) */
    int32_t proc_rank, proc_world;   /* process-level shard; world 0/1 = single process */
    soup_allgather_fn allgather;     /* required when proc_world > 1 */
    soup_bcast_fn bcast;
    void* comm_user;
    int32_t init_strategy;           /* SOUP_INIT_* (--initialization_strategy) */
    /* Cell sharding under the driver: this process holds a contiguous range of the cells in its ONE context;
     * every process of the cell group runs the same restarts and all-reduces the M-pass statistics each iteration (soup_solve_params).
     * `cell_allgather` (host memory, over the cell group) is used once per souporcell3 run to add up the per-cluster cell counts that
     * bad_cluster_detection (S:150-187) ranks.  best_log_probs then covers the local cells only.  Combined with proc_world > 1 this is the
     * hybrid mode: proc_rank / proc_world number the cell GROUPS (restarts shard over groups), and allgather / bcast run over the
     * processes that hold the SAME cell range in every group. */
    soup_allreduce_fn cell_allreduce; /* NULL = off */
    void* cell_comm_user;
    uint64_t n_cells_global;
    soup_allgather_fn cell_allgather;
    int32_t cell_group_size;          /* processes in the cell group (what cell_allgather gathers over) */
    /* The same two modes over NCCL inside the library (no callbacks):
     *  - cell_comms: one communicator per context.  n_ctx > 1: the contexts of this process are the cell group (context i holds the
     *    i-th contiguous cell range; best_log_probs covers all of them, in order).  n_ctx == 1: one process per cell shard, as above
     *    (the per-cluster cell counts travel through soup_comm_allgather).
     *  - proc_comm: restart sharding over processes (proc_rank / proc_world): the winner exchange runs over it instead of allgather / bcast. */
    soup_comm* const* cell_comms;     /* [n_ctx] or NULL */
    soup_comm* proc_comm;             /* or NULL */
    /* Several cell groups on the same GPUs (0 / 1 = one): the n_ctx contexts are cell_groups groups of n_ctx / cell_groups contexts,
     * group-major; every group holds the same cell ranges, has its own communicator and runs its share of the restarts (solve mod
     * groups), so one group's all-reduce overlaps the other groups' E / M passes. */
    int32_t cell_groups;
} soup_main_params;
typedef struct soup_main_result {
    int32_t has_best, runs;
    float best_loss;
    float* best_log_probs; /* host [n_cells][k], caller-allocated */
    float* best_theta;     /* host [k][n_loci], caller-allocated, may be NULL */
    uint64_t nnz_updates;  /* over the solves this process ran */
    uint64_t kernel_launches;
    uint64_t slot_iterations; /* solve-iterations executed (sum over E launches of the active slots) */
    uint64_t timed_iterations;
    double solve_ms;       /* sum over runs of the slowest context's device time */
    double estep_ms, mstep_ms; /* context 0, time_kernels = 1 */
} soup_main_result;
int32_t soup_main(soup_ctx* const* ctxs, int32_t n_ctx, const soup_main_params* p, soup_main_result* r, void* log_file);
/* The per-run cross-process step of soup_main on its own (host only): every rank passes its best
 * (loss, solve, has_best) and buffers; on return all ranks hold the winner's (higher loss; ties ->
 * lowest solve index, S:110/S:120).  Buffers may be NULL. */
int32_t soup_exchange_best(int32_t rank, int32_t world, soup_allgather_fn allgather, soup_bcast_fn bcast, void* user,
                           float* loss, int32_t* solve, int32_t* has_best, float* log_probs, uint64_t n_log_probs,
                           float* theta, uint64_t n_theta, int32_t* winner_rank);

/* TSV writer S:133-147 into `out_file` (a FILE*): barcode, FIRST-max cluster, k log-probs. */
int32_t soup_write_clusters(void* out_file, const char* barcode_blob, uint64_t n_barcodes,
                            const float* best_log_probs, uint64_t n_cells, int32_t k);

#ifdef __cplusplus
}
#endif
#endif /* SOUPGPU_H */
