// libsoupgpu device side: context, CSR upload + on-GPU CSC transpose, the resident-slot EM/KHM
// engine (soup_solve) and the fine-grained parity hooks.  sm_100a only; there is no CPU fallback.
// Reference: /root/reference/souporcell/src/main.rs ("S:").
#include <cuda_runtime.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>   // header-only; ranges cost nothing unless a profiler is attached

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <cub/cub.cuh>

#include "soup_kernels.cuh"
#include "host_parallel.hpp"

using namespace soup;

namespace soup {
void* pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void pinned_free(void* p) { if (p) cudaFreeHost(p); }
}  // namespace soup

// ------------------------------------------------------------------ workspace of one (K, slots, V) shape
struct Workspace {
    int K = 0, slots = 0, iter_cap = 0, queue_cap = 0, forced_lanes = 0;
    struct GroupPlan { int VM = 0, VT = 0, n_main = 0, n_groups = 0; uint32_t padded = 0; } ep, mp, lp;   // E / M wide / M long column groups
    uint32_t Cpad = 0;
    float *LA = nullptr, *LB = nullptr, *TH = nullptr, *ELL = nullptr, *W = nullptr, *LSE = nullptr, *loss = nullptr;
    SlotState* st = nullptr;
    Ctl* ctl = nullptr;
    uint8_t *col_write = nullptr, *k_locked = nullptr;
    int *e_group_active = nullptr, *m_group_active = nullptr, *l_group_active = nullptr, *counts = nullptr, *queue = nullptr, *draw_of_k = nullptr, *solve_iters = nullptr;
    float *solve_loss = nullptr, *best_lp = nullptr, *best_theta = nullptr;
    IterRec* trace = nullptr;
    float* stats = nullptr;     // cell-sharded mode: [SP | DP | per-slot loss], one all-reduce per iteration
    size_t stats_floats = 0;
    int2* moves = nullptr;
    int* k_of_draw = nullptr;   // [K] cluster of compact weight column d (locked clusters keep no weight column)
    int nd = 0;                 // weight columns per slot: K, or the re-drawn clusters of a souporcell3 re-run
    // compact re-run (PostArgs): the engine holds Kx = K columns per slot for the re-drawn clusters of a model with K_out clusters
    int K_out = 0, Kf = 0;
    uint32_t min_cols = 0;
    int *src_of_k = nullptr, *w_of_k = nullptr, *fix_k = nullptr;   // [K_out]
    float* ELLfix = nullptr;    // [N][K_out] (first Kf columns of a row used)
    bool mapped = false;        // src_of_k / w_of_k are not the identity
    uint32_t* stream_keys = nullptr;
    int n_streams_cap = 0;
    std::vector<void*> allocs;
    size_t bytes = 0;
};

struct DevBuf {   // grow-only device allocation
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DevBuf() {}
};
struct LoadBufs {
    DevBuf row_ptr, col_ptr, csr_idx, csc_idx, csr_aux, csc_aux, cell_order, locus_order, total_alleles;   // the matrix
    DevBuf cell_order_m;   // cell-sharded mode: the M pass's range-major visiting order of the loci
    DevBuf locus, alt, ref, vals, vals2, keys, keys2, len, len2, idx, idx2, tmp, flags;                        // load-time scratch
    void release_all() {
        for (DevBuf* b : {&row_ptr, &col_ptr, &csr_idx, &csc_idx, &csr_aux, &csc_aux, &cell_order, &locus_order, &total_alleles, &cell_order_m,
                          &locus, &alt, &ref, &vals, &vals2, &keys, &keys2, &len, &len2, &idx, &idx2, &tmp, &flags}) b->release();
    }
};

struct soup_ctx {
    int device = 0;
    LoadBufs lb;
    cudaStream_t stream = nullptr, side = nullptr;   // side: loss + control, concurrent with the M pass
    cudaStream_t side_long = nullptr;                // the M pass's long-locus kernel, concurrent with its wide role
    cudaEvent_t ev_mfork = nullptr, ev_mjoin = nullptr;
    // cell-sharded mode: the all-reduce of each locus range and the theta pass behind it run here while the M pass produces the next range
    cudaStream_t comm_stream = nullptr;
    static constexpr int MAX_CELL_CHUNKS = 16;
    cudaEvent_t ev_chunk[MAX_CELL_CHUNKS] = {}, ev_comm = nullptr, ev_loss = nullptr;
    int cell_ranges = 0;                                  // locus ranges of the cell-sharded M pass (0 = order not built for this matrix)
    uint32_t cell_range_row[MAX_CELL_CHUNKS + 1] = {};    // range c = locus ids [row[c], row[c + 1])
    uint32_t cell_range_pos[MAX_CELL_CHUNKS + 1] = {};    // ... visited at the positions [pos[c], pos[c + 1]) of lb.cell_order_m (pos[0] = n_long)
    int long_coop = 1;          // 0 never, 1 while few columns are live (the tail of a job), 2 always (SOUP_LONG_COOP)
    uint32_t tail_cols = 128;   // "few": at most this many live columns (SOUP_TAIL_COLS, <= 256)
    bool own_stream = false;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    int sm_count = 0;
    // matrix
    uint64_t N = 0, L = 0, nnz = 0;
    uint64_t *d_row_ptr = nullptr, *d_col_ptr = nullptr;
    uint32_t *d_csr_idx = nullptr, *d_csc_idx = nullptr;
    float4* d_csr_aux = nullptr;
    float2* d_csc_aux = nullptr;
    Packing pk_csr{}, pk_csc{};
    uint32_t *d_cell_order = nullptr, *d_locus_order = nullptr;
    float* d_total_alleles = nullptr;
    bool loaded = false;
    uint32_t n_long = 0;   // loci (in LPT order) whose M-pass chain is long enough for the narrow deep-walk role
    uint32_t n_long_tail = 0;   // ... for the cooperative kernel while few columns are live (>= n_long)
    Workspace ws;
    // pinned staging for the polling loop
    Ctl* h_ctl = nullptr;
    // last-call bookkeeping
    soup_stats stats{};
    std::vector<IterRec> h_trace;
    std::vector<int> h_trace_solve, h_trace_iters;
    int h_trace_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_poll[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_k;   // per-iteration E / M brackets (time_kernels)

    int32_t fail(int32_t code, const char* fmt, ...) {
        va_list ap;
        va_start(ap, fmt);
        err = vformat(fmt, ap);
        va_end(ap);
        return code;
    }
};

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return ctx->fail(SOUP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

static void ws_free(Workspace& w) {
    for (void* p : w.allocs) cudaFree(p);
    w = Workspace();
}
// NB: the zero fill must be enqueued on the context's own (non-blocking) stream: a plain cudaMemset runs on the
// legacy default stream, which does not order against non-blocking streams, and could land after later uploads.
template <typename T>
static cudaError_t ws_alloc(Workspace& w, cudaStream_t st, T** p, size_t count, bool zero) {
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e != cudaSuccess) return e;
    w.allocs.push_back(*p);
    w.bytes += bytes;
    if (zero) e = cudaMemsetAsync(*p, 0, bytes, st);
    return e;
}

static void free_matrix(soup_ctx* ctx) {
    ctx->lb.release_all();
    ctx->d_row_ptr = ctx->d_col_ptr = nullptr;
    ctx->d_csr_idx = ctx->d_csc_idx = nullptr;
    ctx->d_csr_aux = nullptr; ctx->d_csc_aux = nullptr;
    ctx->d_cell_order = ctx->d_locus_order = nullptr;
    ctx->d_total_alleles = nullptr;
    ctx->loaded = false;
}

extern "C" const char* soup_last_error(const soup_ctx* ctx) { return ctx ? ctx->err.c_str() : thread_error(); }

extern "C" int32_t soup_ctx_create(int32_t device, void* cuda_stream, soup_ctx** out) try {
    if (!out) { set_thread_error("soup_ctx_create: bad argument"); return SOUP_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_thread_error("no CUDA device available (%s); libsoupgpu has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return SOUP_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_thread_error("soup_ctx_create: device %d out of range (%d visible)", device, n); return SOUP_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { set_thread_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return SOUP_ERR_CUDA; }
    if (prop.major != 10) {
        set_thread_error("device %d is sm_%d%d; libsoupgpu is built for sm_100a (B200) only and has no fallback", device, prop.major, prop.minor);
        return SOUP_ERR_CUDA;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { set_thread_error("cudaSetDevice: %s", cudaGetErrorString(e)); return SOUP_ERR_CUDA; }
    soup_ctx* ctx = new soup_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cuda_stream) ctx->stream = (cudaStream_t)cuda_stream;
    else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            set_thread_error("cudaStreamCreate: %s", cudaGetErrorString(e));
            delete ctx;
            return SOUP_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking);
    {   // the long-locus chains are the M pass's critical path: their CTAs go first
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&ctx->side_long, cudaStreamNonBlocking, hi);
        cudaEventCreateWithFlags(&ctx->ev_mfork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_mjoin, cudaEventDisableTiming);
        cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi);
        for (cudaEvent_t& e : ctx->ev_chunk) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_loss, cudaEventDisableTiming);
        if (const char* e = getenv("SOUP_LONG_COOP")) ctx->long_coop = std::min(2, std::max(0, atoi(e)));
        if (const char* e = getenv("SOUP_TAIL_COLS")) ctx->tail_cols = (uint32_t)std::min(256, std::max(1, atoi(e)));
        if (cudaFuncSetAttribute(mstep_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ML_SMEM_FLOATS * sizeof(float))) != cudaSuccess) {
            cudaGetLastError();
            ctx->long_coop = 0;
        }
    }
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreateWithFlags(&ctx->ev_poll[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_poll[1], cudaEventDisableTiming);
    if (cudaHostAlloc((void**)&ctx->h_ctl, sizeof(Ctl) * 4, cudaHostAllocDefault) != cudaSuccess) {
        set_thread_error("cudaHostAlloc failed");
        delete ctx;
        return SOUP_ERR_CUDA;
    }
    // statrs FCACHE -> ln(n!) table (S:669-670)
    double lf[171], f = 1.0;
    lf[0] = std::log(1.0);
    for (int i = 1; i <= 170; ++i) { f *= (double)i; lf[i] = std::log(f); }
    if ((e = cudaMemcpyToSymbol(c_ln_factorial, lf, sizeof(lf))) != cudaSuccess) {
        set_thread_error("cudaMemcpyToSymbol: %s", cudaGetErrorString(e));
        delete ctx;
        return SOUP_ERR_CUDA;
    }
    {   // ln_binomial table for in-field counts (statrs FCACHE arithmetic, S:669-670); identical for every context
        std::vector<float> tab(4096, 0.0f);
        for (uint32_t a = 0; a < 64; ++a)
            for (uint32_t r = 0; r < 64; ++r) tab[(a << 6) | r] = ln_binomial_f32(a + r, a);
        if ((e = cudaMemcpyToSymbol(c_lbc, tab.data(), tab.size() * sizeof(float))) != cudaSuccess) {
            set_thread_error("cudaMemcpyToSymbol: %s", cudaGetErrorString(e));
            delete ctx;
            return SOUP_ERR_CUDA;
        }
    }
    *out = ctx;
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_ctx_create")

extern "C" void soup_ctx_destroy(soup_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ws_free(ctx->ws);
    free_matrix(ctx);
    if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->ev_poll) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_k) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->side_long) cudaStreamDestroy(ctx->side_long);
    if (ctx->ev_mfork) cudaEventDestroy(ctx->ev_mfork);
    if (ctx->ev_mjoin) cudaEventDestroy(ctx->ev_mjoin);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    for (cudaEvent_t e : ctx->ev_chunk) if (e) cudaEventDestroy(e);
    if (ctx->ev_comm) cudaEventDestroy(ctx->ev_comm);
    if (ctx->ev_loss) cudaEventDestroy(ctx->ev_loss);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ------------------------------------------------------------------ soup_comm: one rank of an NCCL communicator, bound to a context's device
struct soup_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;       // the helpers' own stream (host-buffer all-gather / broadcast)
    DevBuf stage;                        // device staging of the host helpers
};
static int32_t nccl_fail(const char* what, ncclResult_t r) {
    set_thread_error("%s failed: %s", what, ncclGetErrorString(r));
    return SOUP_ERR_CUDA;
}
extern "C" int32_t soup_comm_unique_id(uint8_t id[SOUP_COMM_ID_BYTES]) try {
    static_assert(sizeof(ncclUniqueId) == SOUP_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!id) { set_thread_error("soup_comm_unique_id: null argument"); return SOUP_ERR_INVALID; }
    ncclUniqueId u;
    ncclResult_t r = ncclGetUniqueId(&u);
    if (r != ncclSuccess) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id, &u, sizeof(u));
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_comm_unique_id")
static soup_comm* comm_wrap(ncclComm_t c, int rank, int world, int device) {
    soup_comm* m = new soup_comm();
    m->comm = c; m->rank = rank; m->world = world; m->device = device;
    cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    return m;
}
extern "C" int32_t soup_comm_create(soup_ctx* ctx, const uint8_t id[SOUP_COMM_ID_BYTES], int32_t rank, int32_t world, soup_comm** out) try {
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) { set_thread_error("soup_comm_create: bad argument"); return SOUP_ERR_INVALID; }
    *out = nullptr;
    if (cudaSetDevice(ctx->device) != cudaSuccess) { set_thread_error("soup_comm_create: cudaSetDevice failed"); return SOUP_ERR_CUDA; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    ncclResult_t r = ncclCommInitRank(&c, world, u, rank);
    if (r != ncclSuccess) return nccl_fail("ncclCommInitRank", r);
    *out = comm_wrap(c, rank, world, ctx->device);
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_comm_create")
extern "C" int32_t soup_comm_create_all(soup_ctx* const* ctxs, int32_t n_ctx, soup_comm** out) try {
    if (!ctxs || !out || n_ctx < 1) { set_thread_error("soup_comm_create_all: bad argument"); return SOUP_ERR_INVALID; }
    std::vector<int> devs(n_ctx);
    for (int i = 0; i < n_ctx; ++i) {
        if (!ctxs[i]) { set_thread_error("soup_comm_create_all: null context"); return SOUP_ERR_INVALID; }
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) { set_thread_error("soup_comm_create_all: contexts %d and %d share device %d (NCCL needs one GPU per rank)", j, i, devs[i]); return SOUP_ERR_INVALID; }
        out[i] = nullptr;
    }
    std::vector<ncclComm_t> cs(n_ctx, nullptr);
    ncclResult_t r = ncclCommInitAll(cs.data(), n_ctx, devs.data());
    if (r != ncclSuccess) return nccl_fail("ncclCommInitAll", r);
    for (int i = 0; i < n_ctx; ++i) { cudaSetDevice(devs[i]); out[i] = comm_wrap(cs[i], i, n_ctx, devs[i]); }
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_comm_create_all")
// A second communicator over the same ranks (ncclCommSplit with one colour: no new bootstrap), for another cell group on the same GPUs
extern "C" int32_t soup_comm_dup_all(soup_comm* const* in, int32_t n, soup_comm** out) try {
    if (!in || !out || n < 1) { set_thread_error("soup_comm_dup_all: bad argument"); return SOUP_ERR_INVALID; }
    std::vector<ncclComm_t> cs(n, nullptr);
    ncclResult_t r = ncclGroupStart();
    for (int i = 0; i < n && r == ncclSuccess; ++i) {
        if (!in[i]) { ncclGroupEnd(); set_thread_error("soup_comm_dup_all: null communicator"); return SOUP_ERR_INVALID; }
        cudaSetDevice(in[i]->device);
        r = ncclCommSplit(in[i]->comm, 0, in[i]->rank, &cs[i], nullptr);
    }
    ncclResult_t e = ncclGroupEnd();
    if (r == ncclSuccess) r = e;
    if (r != ncclSuccess) return nccl_fail("ncclCommSplit", r);
    for (int i = 0; i < n; ++i) { cudaSetDevice(in[i]->device); out[i] = comm_wrap(cs[i], in[i]->rank, in[i]->world, in[i]->device); }
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_comm_dup_all")
extern "C" void soup_comm_destroy(soup_comm* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    m->stage.release();
    if (m->comm) ncclCommDestroy(m->comm);
    delete m;
}
extern "C" int32_t soup_comm_rank(const soup_comm* m, int32_t* rank, int32_t* world) {
    if (!m) { set_thread_error("soup_comm_rank: null communicator"); return SOUP_ERR_INVALID; }
    if (rank) *rank = m->rank;
    if (world) *world = m->world;
    return SOUP_OK;
}
// host buffers over the communicator, staged through device memory (a few bytes to a few MB, once per run)
extern "C" int32_t soup_comm_allgather(soup_comm* m, const void* send, void* recv, uint64_t bytes) try {
    if (!m || !send || !recv) { set_thread_error("soup_comm_allgather: bad argument"); return SOUP_ERR_INVALID; }
    if (bytes == 0) return SOUP_OK;
    if (cudaSetDevice(m->device) != cudaSuccess || m->stage.ensure((size_t)bytes * (m->world + 1)) != cudaSuccess) { cudaGetLastError(); set_thread_error("soup_comm_allgather: staging allocation failed"); return SOUP_ERR_NOMEM; }
    char* d_send = (char*)m->stage.p;
    char* d_recv = d_send + bytes;
    cudaError_t e = cudaMemcpyAsync(d_send, send, bytes, cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess) {
        ncclResult_t r = ncclAllGather(d_send, d_recv, bytes, ncclChar, m->comm, m->stream);
        if (r != ncclSuccess) return nccl_fail("ncclAllGather", r);
        e = cudaMemcpyAsync(recv, d_recv, bytes * m->world, cudaMemcpyDeviceToHost, m->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) { set_thread_error("soup_comm_allgather: %s", cudaGetErrorString(e)); return SOUP_ERR_CUDA; }
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_comm_allgather")
extern "C" int32_t soup_comm_bcast(soup_comm* m, void* buf, uint64_t bytes, int32_t root) try {
    if (!m || !buf || root < 0 || root >= m->world) { set_thread_error("soup_comm_bcast: bad argument"); return SOUP_ERR_INVALID; }
    if (bytes == 0) return SOUP_OK;
    if (cudaSetDevice(m->device) != cudaSuccess || m->stage.ensure((size_t)bytes) != cudaSuccess) { cudaGetLastError(); set_thread_error("soup_comm_bcast: staging allocation failed"); return SOUP_ERR_NOMEM; }
    cudaError_t e = cudaSuccess;
    if (m->rank == root) e = cudaMemcpyAsync(m->stage.p, buf, bytes, cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess) {
        ncclResult_t r = ncclBroadcast(m->stage.p, m->stage.p, bytes, ncclChar, root, m->comm, m->stream);
        if (r != ncclSuccess) return nccl_fail("ncclBroadcast", r);
        if (m->rank != root) e = cudaMemcpyAsync(buf, m->stage.p, bytes, cudaMemcpyDeviceToHost, m->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) { set_thread_error("soup_comm_bcast: %s", cudaGetErrorString(e)); return SOUP_ERR_CUDA; }
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_comm_bcast")

// ------------------------------------------------------------------ soup_load_csr
#include <chrono>
struct PhaseTimer {
    bool on = getenv("SOUP_DEBUG_ITERS") != nullptr;
    bool slow_only = getenv("SOUP_DEBUG_SLOW") != nullptr;   // no extra syncs; report only phases above 5 ms
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (slow_only) {
            auto t1 = std::chrono::steady_clock::now();
            double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
            if (ms > 5.0 && strncmp(what, "solve: iterate", 14) != 0) fprintf(stderr, "soup_slow phase %-28s %.3f ms\n", what, ms);
            t0 = t1;
            return;
        }
        if (!on) return;
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "soup_debug phase %-28s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

extern "C" int32_t soup_load_csr(soup_ctx* ctx, uint64_t n_cells, uint64_t n_loci, uint64_t nnz_in, const uint64_t* row_ptr,
                                 const uint32_t* locus, const uint32_t* alt, const uint32_t* ref) try {
    if (!ctx) return SOUP_ERR_INVALID;
    PhaseTimer pt;
    if (!row_ptr || (nnz_in && (!locus || !alt || !ref))) return ctx->fail(SOUP_ERR_INVALID, "soup_load_csr: null pointer");
    if (n_cells >= (1u << 29) || n_loci >= (1u << 29) || nnz_in >= INT32_MAX)
        return ctx->fail(SOUP_ERR_UNSUPPORTED, "soup_load_csr: more than 2^29 cells / loci or 2^31 entries are not supported");
    if (row_ptr[0] != 0 || row_ptr[n_cells] != nnz_in) return ctx->fail(SOUP_ERR_INVALID, "soup_load_csr: row_ptr does not span [0, nnz]");
    for (uint64_t c = 0; c < n_cells; ++c)
        if (row_ptr[c + 1] < row_ptr[c]) return ctx->fail(SOUP_ERR_INVALID, "soup_load_csr: row_ptr not monotone at cell %llu", (unsigned long long)c);
    CK(cudaSetDevice(ctx->device));
    ctx->stats = soup_stats{};
    ctx->loaded = false;
    ctx->cell_ranges = 0;
    cudaStream_t st = ctx->stream;
    // Device buffers are grow-only and cached in the context: reloading a matrix of the same size allocates nothing.
    LoadBufs& B = ctx->lb;
    const bool same_shape = ctx->N == n_cells && ctx->L == n_loci;
    if (!same_shape) ws_free(ctx->ws);
#define CKL(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) { cudaGetLastError(); return ctx->fail(e_ == cudaErrorMemoryAllocation ? SOUP_ERR_NOMEM : SOUP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    std::vector<uint64_t> rp2;
    std::vector<uint32_t> l2, a2, r2;
    uint64_t nnz = nnz_in;
    LoadFlags flags{};
    for (int attempt = 0; attempt < 2; ++attempt) {
        const size_t nz = std::max<uint64_t>(nnz, 1);
        CKL(B.row_ptr.ensure((n_cells + 1) * 8)); CKL(B.locus.ensure(nz * 4)); CKL(B.alt.ensure(nz * 4)); CKL(B.ref.ensure(nz * 4));
        CKL(B.flags.ensure(sizeof(LoadFlags)));
        CKL(cudaMemcpyAsync(B.row_ptr.p, row_ptr, (n_cells + 1) * 8, cudaMemcpyHostToDevice, st));
        if (nnz) {
            CKL(cudaMemcpyAsync(B.locus.p, locus, nnz * 4, cudaMemcpyHostToDevice, st));
            CKL(cudaMemcpyAsync(B.alt.p, alt, nnz * 4, cudaMemcpyHostToDevice, st));
            CKL(cudaMemcpyAsync(B.ref.p, ref, nnz * 4, cudaMemcpyHostToDevice, st));
        }
        ctx->stats.h2d_bytes += (n_cells + 1) * 8 + nnz * 12;
        // validate on the device: loci strictly ascending inside a cell (the order S:663-674 appends in) and in range;
        // count ref+alt == 0 entries (dropped, S:664) and counts beyond the ln-factorial table
        CKL(cudaMemsetAsync(B.flags.p, 0, sizeof(LoadFlags), st));
        if (n_cells) validate_csr_kernel<<<(unsigned)((n_cells + 7) / 8), 256, 0, st>>>((const uint64_t*)B.row_ptr.p, (const uint32_t*)B.locus.p, (const uint32_t*)B.alt.p,
                                                                                        (const uint32_t*)B.ref.p, (LoadFlags*)B.flags.p, (uint32_t)n_cells, (uint32_t)n_loci);
        CKL(cudaMemcpyAsync(&flags, B.flags.p, sizeof(LoadFlags), cudaMemcpyDeviceToHost, st));
        CKL(cudaStreamSynchronize(st));
        if (flags.bad_range) return ctx->fail(SOUP_ERR_INVALID, "soup_load_csr: locus index out of range (cell %u)", flags.bad_cell);
        if (flags.bad_order) return ctx->fail(SOUP_ERR_INVALID, "soup_load_csr: loci must be strictly ascending inside a cell (cell %u)", flags.bad_cell);
        if (flags.n_zero == 0 || attempt == 1) break;
        // rare: explicit all-zero entries -> compact on the host (S:664) and upload again
        rp2.resize(n_cells + 1);
        l2.reserve(nnz_in); a2.reserve(nnz_in); r2.reserve(nnz_in);
        for (uint64_t c = 0; c < n_cells; ++c) {
            rp2[c] = l2.size();
            for (uint64_t j = row_ptr[c]; j < row_ptr[c + 1]; ++j)
                if ((uint32_t)(alt[j] + ref[j]) != 0) { l2.push_back(locus[j]); a2.push_back(alt[j]); r2.push_back(ref[j]); }
        }
        rp2[n_cells] = l2.size();
        nnz = l2.size();
        row_ptr = rp2.data(); locus = l2.data(); alt = a2.data(); ref = r2.data();
    }
    pt.lap("load: h2d + validate");
    ctx->N = n_cells; ctx->L = n_loci; ctx->nnz = nnz;
    // entry words (soup_kernels.cuh "entry words"): bit 31 = general; index bits as needed, the rest split between
    // the two count fields (<= 6 bits each)
    auto make_packing = [](uint64_t n_index, uint32_t top) {   // count fields below bit `top`
        uint32_t ib = 1;
        while ((1ull << ib) < std::max<uint64_t>(n_index, 2)) ++ib;
        Packing pk;
        pk.cbits = std::min<uint32_t>(6, (top - ib) / 2);
        pk.cmask = (1u << pk.cbits) - 1;
        pk.sh_a = top - pk.cbits;
        pk.sh_r = top - 2 * pk.cbits;
        pk.imask = (1u << pk.sh_r) - 1;
        pk.plane_rows = 0;
        pk.alt_bit = 1u << pk.sh_a;
        return pk;
    };
    // CSR words index the joint ln table (2 planes of max(L, 1) rows) and keep bit 30 for the "row counts twice" flag of simple words
    ctx->pk_csr = make_packing(2 * std::max<uint64_t>(n_loci, 1), 30);
    ctx->pk_csr.plane_rows = (uint32_t)std::max<uint64_t>(n_loci, 1);   // = the row count ensure_workspace gives each ln-table plane
    ctx->pk_csc = make_packing(n_cells, 31);
    const size_t nz = std::max<uint64_t>(nnz, 1);
    const size_t nmax = std::max<uint64_t>(std::max(n_cells, n_loci), 1);
    CKL(B.col_ptr.ensure((n_loci + 1) * 8)); CKL(B.csr_idx.ensure(nz * 4)); CKL(B.csc_idx.ensure(nz * 4));
    CKL(B.csr_aux.ensure(nz * sizeof(float4))); CKL(B.csc_aux.ensure(nz * sizeof(float2)));
    CKL(B.cell_order.ensure(std::max<uint64_t>(n_cells, 1) * 4)); CKL(B.locus_order.ensure(std::max<uint64_t>(n_loci, 1) * 4));
    CKL(B.total_alleles.ensure(std::max<uint64_t>(n_cells, 1) * 4));
    CKL(B.vals.ensure(nz * 4)); CKL(B.vals2.ensure(nz * 4)); CKL(B.keys.ensure(nz * 8)); CKL(B.keys2.ensure(nz * 8));
    CKL(B.len.ensure(nmax * 4)); CKL(B.len2.ensure(nmax * 4)); CKL(B.idx.ensure(nmax * 4)); CKL(B.idx2.ensure(nmax * 4));
    ctx->d_row_ptr = (uint64_t*)B.row_ptr.p; ctx->d_col_ptr = (uint64_t*)B.col_ptr.p;
    ctx->d_csr_idx = (uint32_t*)B.csr_idx.p; ctx->d_csc_idx = (uint32_t*)B.csc_idx.p;
    ctx->d_csr_aux = (float4*)B.csr_aux.p; ctx->d_csc_aux = (float2*)B.csc_aux.p;
    ctx->d_cell_order = (uint32_t*)B.cell_order.p; ctx->d_locus_order = (uint32_t*)B.locus_order.p;
    ctx->d_total_alleles = (float*)B.total_alleles.p;
    uint32_t *d_locus = (uint32_t*)B.locus.p, *d_alt = (uint32_t*)B.alt.p, *d_ref = (uint32_t*)B.ref.p;
    uint32_t *d_vals = (uint32_t*)B.vals.p, *d_vals2 = (uint32_t*)B.vals2.p, *d_len = (uint32_t*)B.len.p, *d_len2 = (uint32_t*)B.len2.p;
    uint32_t *d_idx = (uint32_t*)B.idx.p, *d_idx2 = (uint32_t*)B.idx2.p;
    unsigned long long *d_keys = (unsigned long long*)B.keys.p, *d_keys2 = (unsigned long long*)B.keys2.p;
    if (n_cells) {
        build_csr_kernel<<<(unsigned)((n_cells + 7) / 8), 256, 0, st>>>(ctx->d_row_ptr, d_locus, d_alt, d_ref, ctx->d_csr_idx, ctx->d_csr_aux, d_keys,
                                                                        d_vals, ctx->d_total_alleles, ctx->pk_csr, (uint32_t)n_cells);
        CKL(cudaGetLastError());
    }
    if (flags.n_big) {   // rare: ln_binomial for n > 170 comes from the host (statrs ln_gamma branch)
        std::vector<uint32_t> big_idx;
        std::vector<float> big_val;
        for (uint64_t j = 0; j < nnz; ++j)
            if ((uint32_t)(alt[j] + ref[j]) > 170u) { big_idx.push_back((uint32_t)j); big_val.push_back(ln_binomial_f32(alt[j] + ref[j], alt[j])); }
        DevBuf bi, bv;
        CKL(bi.ensure(big_idx.size() * 4)); CKL(bv.ensure(big_val.size() * 4));
        CKL(cudaMemcpyAsync(bi.p, big_idx.data(), big_idx.size() * 4, cudaMemcpyHostToDevice, st));
        CKL(cudaMemcpyAsync(bv.p, big_val.data(), big_val.size() * 4, cudaMemcpyHostToDevice, st));
        patch_lbc_kernel<<<(unsigned)((big_idx.size() + 255) / 256), 256, 0, st>>>(ctx->d_csr_aux, (const uint32_t*)bi.p, (const float*)bv.p, (uint32_t)big_idx.size());
        CKL(cudaStreamSynchronize(st));
        ctx->stats.h2d_bytes += big_idx.size() * 8;
    }
    // CSC transpose: stable radix sort of (locus, cell) keys -> cells ascending inside a locus (S:227 order)
    int bits_l = 1, bits_c = 1;
    while ((1ull << bits_l) < std::max<uint64_t>(n_loci, 2)) ++bits_l;
    while ((1ull << bits_c) < std::max<uint64_t>(n_cells, 2)) ++bits_c;
    size_t tmp_bytes = 0, tmp2 = 0;
    cub::DoubleBuffer<unsigned long long> kb(d_keys, d_keys2);
    cub::DoubleBuffer<uint32_t> vb(d_vals, d_vals2);
    {
        cub::DoubleBuffer<uint32_t> lq(d_len, d_len2), iq(d_idx, d_idx2);
        CKL(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)nnz, 0, 32 + bits_l, st));
        CKL(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp2, lq, iq, (int)nmax, 0, 32, st));
    }
    tmp_bytes = std::max<size_t>(std::max(tmp_bytes, tmp2), 16);
    CKL(B.tmp.ensure(tmp_bytes));
    if (nnz) {
        // the CSR already is sorted by cell, and LSD radix passes are stable: sorting on the locus bits alone is enough
        CKL(cub::DeviceRadixSort::SortPairs(B.tmp.p, tmp_bytes, kb, vb, (int)nnz, 32, 32 + bits_l, st));
        build_csc_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(kb.Current(), vb.Current(), d_alt, d_ref, ctx->d_csc_idx, ctx->d_csc_aux, ctx->pk_csc, nnz);
        CKL(cudaGetLastError());
    }
    (void)bits_c;
    col_ptr_kernel<<<(unsigned)((n_loci + 1 + 255) / 256), 256, 0, st>>>(kb.Current(), nnz, ctx->d_col_ptr, (uint32_t)n_loci);
    CKL(cudaGetLastError());
    // longest-first (LPT) visiting orders for the E and M kernels (ties -> ascending index; stable sort)
    auto lpt = [&](const uint64_t* ptr, uint32_t n, uint32_t* order) -> cudaError_t {
        if (n == 0) return cudaSuccess;
        seg_len_kernel<<<(n + 255) / 256, 256, 0, st>>>(ptr, d_len, d_idx, n);
        cub::DoubleBuffer<uint32_t> k(d_len, d_len2), v(d_idx, d_idx2);
        size_t tb = tmp_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(B.tmp.p, tb, k, v, (int)n, 0, 32, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(order, v.Current(), (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess && k.Current() != d_len) e = cudaMemcpyAsync(d_len, k.Current(), (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
        return e;
    };
    CKL(lpt(ctx->d_row_ptr, (uint32_t)n_cells, ctx->d_cell_order));
    CKL(lpt(ctx->d_col_ptr, (uint32_t)n_loci, ctx->d_locus_order));
    {   // long loci: covered by >= 40 % of the cells (and >= 512 entries); d_len now holds the locus lengths, descending
        uint64_t thr = std::max<uint64_t>(512, n_cells * 2 / 5);
        if (const char* e = getenv("SOUP_LONG_THRESHOLD")) thr = std::max<uint64_t>(1, strtoull(e, nullptr, 10));   // tests / tuning
        count_ge_kernel<<<1, 1, 0, st>>>(d_len, (uint32_t)n_loci, (uint32_t)std::min<uint64_t>(thr, UINT32_MAX), &((LoadFlags*)B.flags.p)->n_long);
        uint64_t thr_tail = std::min<uint64_t>(thr, 2048);   // with few live columns every chain longer than this goes to the cooperative kernel
        if (const char* e = getenv("SOUP_TAIL_THRESHOLD")) thr_tail = std::max<uint64_t>(1, strtoull(e, nullptr, 10));
        thr_tail = std::min(thr_tail, thr);
        count_ge_kernel<<<1, 1, 0, st>>>(d_len, (uint32_t)n_loci, (uint32_t)thr_tail, &((LoadFlags*)B.flags.p)->n_long_tail);
        CKL(cudaMemcpyAsync(&flags, B.flags.p, sizeof(LoadFlags), cudaMemcpyDeviceToHost, st));
        CKL(cudaStreamSynchronize(st));
        ctx->n_long = getenv("SOUP_NO_LONG") ? 0 : flags.n_long;
        ctx->n_long_tail = getenv("SOUP_NO_LONG") ? 0 : std::max(flags.n_long, flags.n_long_tail);
    }
    CKL(cudaGetLastError());
    pt.lap("load: build csr/csc/lpt");
#undef CKL
    ctx->stats.kernel_launches = 8;
    ctx->loaded = true;
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_load_csr")

extern "C" int32_t soup_get_dims(soup_ctx* ctx, uint64_t* n_cells, uint64_t* n_loci, uint64_t* nnz) {
    if (!ctx) return SOUP_ERR_INVALID;
    if (!ctx->loaded) return ctx->fail(SOUP_ERR_INVALID, "no matrix loaded (call soup_load_csr first)");
    if (n_cells) *n_cells = ctx->N;
    if (n_loci) *n_loci = ctx->L;
    if (nnz) *nnz = ctx->nnz;
    return SOUP_OK;
}
extern "C" int32_t soup_get_csc(soup_ctx* ctx, uint64_t* col_ptr, uint32_t* cell, uint32_t* alt, uint32_t* ref) try {
    if (!ctx || !ctx->loaded) return ctx ? ctx->fail(SOUP_ERR_INVALID, "no matrix loaded") : SOUP_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (col_ptr) CK(cudaMemcpy(col_ptr, ctx->d_col_ptr, (ctx->L + 1) * 8, cudaMemcpyDeviceToHost));
    if (cell || alt || ref) {
        std::vector<uint32_t> hi(ctx->nnz);
        std::vector<float2> hx(ctx->nnz);
        if (ctx->nnz) {
            CK(cudaMemcpy(hi.data(), ctx->d_csc_idx, ctx->nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hx.data(), ctx->d_csc_aux, ctx->nnz * sizeof(float2), cudaMemcpyDeviceToHost));
        }
        for (uint64_t i = 0; i < ctx->nnz; ++i) {
            if (cell) cell[i] = hi[i] & ctx->pk_csc.imask;
            if (alt) alt[i] = (uint32_t)hx[i].x;
            if (ref) ref[i] = (uint32_t)hx[i].y - (uint32_t)hx[i].x;
        }
    }
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_get_csc")
extern "C" int32_t soup_get_cell_totals(soup_ctx* ctx, float* total_alleles) try {
    if (!ctx || !ctx->loaded || !total_alleles) return ctx ? ctx->fail(SOUP_ERR_INVALID, "no matrix loaded") : SOUP_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->N) CK(cudaMemcpy(total_alleles, ctx->d_total_alleles, ctx->N * 4, cudaMemcpyDeviceToHost));
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_get_cell_totals")
extern "C" int32_t soup_get_csr_lbc(soup_ctx* ctx, float* lbc) try {
    if (!ctx || !ctx->loaded || !lbc) return ctx ? ctx->fail(SOUP_ERR_INVALID, "no matrix loaded") : SOUP_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    std::vector<float4> h(ctx->nnz);
    if (ctx->nnz) CK(cudaMemcpy(h.data(), ctx->d_csr_aux, ctx->nnz * sizeof(float4), cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < ctx->nnz; ++i) lbc[i] = h[i].z;
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_get_csr_lbc")

extern "C" int32_t soup_cluster_allele_counts(soup_ctx* ctx, const int32_t* cell_cluster, int32_t k, uint64_t* counts) try {
    if (!ctx) return SOUP_ERR_INVALID;
    if (!ctx->loaded) return ctx->fail(SOUP_ERR_INVALID, "no matrix loaded (call soup_load_csr first)");
    if (!cell_cluster || !counts || k <= 0) return ctx->fail(SOUP_ERR_INVALID, "soup_cluster_allele_counts: bad argument");
    CK(cudaSetDevice(ctx->device));
    const size_t n_out = (size_t)ctx->L * (size_t)k * 2;
    if (n_out == 0) return SOUP_OK;
    int32_t* d_cl = nullptr;
    unsigned long long* d_out = nullptr;
    cudaError_t e = cudaMalloc(&d_cl, std::max<uint64_t>(ctx->N, 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_out, n_out * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cl, cell_cluster, ctx->N * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_out, 0, n_out * 8, ctx->stream);
    if (e == cudaSuccess && ctx->N)
        cluster_counts_kernel<<<(unsigned)((ctx->N + 7) / 8), 256, 0, ctx->stream>>>(ctx->d_row_ptr, (const uint32_t*)ctx->lb.locus.p, (const uint32_t*)ctx->lb.alt.p,
                                                                                    (const uint32_t*)ctx->lb.ref.p, d_cl, d_out, (uint32_t)ctx->N, k);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, d_out, n_out * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(d_cl); cudaFree(d_out);
    if (e != cudaSuccess) { cudaGetLastError(); return ctx->fail(e == cudaErrorMemoryAllocation ? SOUP_ERR_NOMEM : SOUP_ERR_CUDA, "soup_cluster_allele_counts: %s", cudaGetErrorString(e)); }
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_cluster_allele_counts")

// ------------------------------------------------------------------ workspace
// Column groups of one kernel: `n_main` groups of 32*VM columns and, if the column count is not a
// multiple of that, one narrower last group (the smallest of 1/2/4/8 columns per lane that holds the
// rest).  The E pass is per-entry-latency bound, so it takes wide groups (8 columns per lane: two
// 512-byte row pieces per gather); the M pass owns a strictly sequential chain per (locus, column), so
// it takes narrower groups (4 per lane) that put more warps on the long loci.  `forced_lanes` (tests)
// makes every group of both kernels that many columns per lane wide.
static Workspace::GroupPlan plan_groups(int cols, int vmax, int forced_lanes) {
    Workspace::GroupPlan g;
    if (forced_lanes) {
        const int gw = 32 * forced_lanes;
        g.VM = g.VT = forced_lanes;
        g.n_main = g.n_groups = (cols + gw - 1) / gw;
        g.padded = (uint32_t)(g.n_groups * gw);
        return g;
    }
    auto fit = [](int c) { return c <= 32 ? 1 : (c <= 64 ? 2 : (c <= 128 ? 4 : (c <= 256 ? 8 : 16))); };
    const int wm = 32 * vmax;
    if (cols <= wm) {                       // one group, as narrow as fits
        g.VM = g.VT = std::min(vmax, fit(cols));
        g.n_main = g.n_groups = 1;
        g.padded = (uint32_t)(32 * g.VM);
        return g;
    }
    g.VM = vmax;
    g.n_main = cols / wm;
    const int rem = cols - g.n_main * wm;
    if (rem == 0) { g.VT = vmax; g.n_groups = g.n_main; g.padded = (uint32_t)cols; return g; }
    g.VT = std::min(vmax, fit(rem));
    g.n_groups = g.n_main + 1;
    g.padded = (uint32_t)(g.n_main * wm + 32 * g.VT);
    return g;
}

// K = clusters per slot in the engine's tables; K_out = clusters of the model (results, counts); min_cols = columns the planes must have at least
static int32_t ensure_workspace(soup_ctx* ctx, int K, int slots, int forced_lanes, int iter_cap, int queue_cap, int n_streams, int K_out = 0, uint32_t min_cols = 0) {
    Workspace& w = ctx->ws;
    if (K_out <= 0) K_out = K;
    if (w.K == K && w.slots == slots && w.forced_lanes == forced_lanes && w.iter_cap == iter_cap && w.queue_cap >= queue_cap && w.n_streams_cap >= n_streams &&
        w.K_out == K_out && w.min_cols == min_cols) return SOUP_OK;
    ws_free(w);
    w.K = K; w.slots = slots; w.forced_lanes = forced_lanes; w.iter_cap = iter_cap; w.queue_cap = queue_cap; w.n_streams_cap = n_streams;
    w.K_out = K_out; w.min_cols = min_cols;
    const size_t cols = std::max<size_t>((size_t)slots * K, min_cols);
    w.ep = plan_groups((int)cols, SOUP_E_VMAX, forced_lanes);
    w.mp = plan_groups((int)cols, SOUP_M_VWIDE, forced_lanes);
    w.lp = plan_groups((int)cols, SOUP_M_VLONG, SOUP_M_VLONG);
    w.Cpad = std::max(std::max(w.ep.padded, w.mp.padded), w.lp.padded);
    const size_t L = std::max<uint64_t>(ctx->L, 1), N = std::max<uint64_t>(ctx->N, 1);
    cudaError_t e = cudaSuccess;
    auto A = [&](auto** p, size_t n, bool zero) { if (e == cudaSuccess) e = ws_alloc(w, ctx->stream, p, n, zero); };
    A(&w.LA, 2 * L * w.Cpad, true); A(&w.TH, L * w.Cpad, true);   // LA | LB: one joint table (a simple CSR word is a row of it)
    if (e == cudaSuccess) w.LB = w.LA + L * w.Cpad;
    A(&w.ELL, N * w.Cpad, true); A(&w.W, N * w.Cpad, true); A(&w.LSE, N * slots, true);
    A(&w.st, (size_t)slots, true); A(&w.ctl, 1, true); A(&w.loss, (size_t)slots, true);
    A(&w.col_write, w.Cpad, true); A(&w.k_locked, (size_t)K, true);
    A(&w.e_group_active, (size_t)w.ep.n_groups, true); A(&w.m_group_active, (size_t)w.mp.n_groups, true);
    A(&w.l_group_active, (size_t)w.lp.n_groups, true); A(&w.counts, (size_t)slots * K_out, true);
    A(&w.src_of_k, (size_t)K_out, true); A(&w.w_of_k, (size_t)K_out, true); A(&w.fix_k, (size_t)K_out, true);
    if (K_out != K) A(&w.ELLfix, N * (size_t)K_out, true);
    A(&w.queue, (size_t)queue_cap, true); A(&w.draw_of_k, (size_t)K, true); A(&w.k_of_draw, (size_t)K, true);
    A(&w.solve_iters, (size_t)queue_cap, true); A(&w.solve_loss, (size_t)queue_cap, true);
    A(&w.best_lp, N * K_out, true); A(&w.best_theta, L * K_out, true);
    A(&w.trace, (size_t)queue_cap * iter_cap, true);
    A(&w.stream_keys, (size_t)std::max(n_streams, 1) * 8, true);
    A(&w.moves, (size_t)slots, true);
    if (e != cudaSuccess) {
        ws_free(w);
        cudaGetLastError();
        return ctx->fail(e == cudaErrorMemoryAllocation ? SOUP_ERR_NOMEM : SOUP_ERR_CUDA, "workspace allocation failed: %s", cudaGetErrorString(e));
    }
    return SOUP_OK;
}

static size_t bytes_per_slot(const soup_ctx* ctx, int K) {
    return (size_t)K * (3 * ctx->L + 2 * ctx->N) * 4 + ctx->N * 4 + 64;
}

// ------------------------------------------------------------------ one EM/KHM iteration for every resident slot
struct IterLaunch {
    soup_ctx* ctx;
    int method;
    float log_prior, limit;
    bool record_counts;
    int timed_iter = -1;   // >= 0: bracket this iteration's E and M launches with events ev_k[4*i .. 4*i+3]
    soup_allreduce_fn allreduce = nullptr;   // cell-sharded mode: the caller's all-reduce ...
    void* comm_user = nullptr;
    soup_comm* comm = nullptr;               // ... or NCCL inside the library
    int32_t comm_error = 0;
    bool cell_mode = false;
    // Columns the statistics rows carry: a multiple of 4 that covers every live column.  Every rank takes the value from the same
    // completed poll (two chunks of iterations back), and Ctl::live_cols never grows, so the ranks always agree on the reduced sizes.
    uint32_t scols = 0;
    int n_ranges = 1;                        // locus ranges the M pass / all-reduce / theta pass are pipelined over
    uint64_t allreduce_bytes = 0, allreduce_calls = 0;
    RefillArgs refill;
    FinalizeArgs fin;
    const float* base_theta_dev = nullptr;   // [K_out][L] centers of the previous run (compact re-run: the locked clusters' rows)
    bool queue_drained = false;   // the host has SEEN queue_head >= queue_len: no slot can get a pending solve any more, so refill launches stop
};

#define SOUP_DISPATCH(KERNEL, THREADS, PLAN, ARGS)                                              \
    do {                                                                                        \
        switch ((PLAN).VM) {                                                                    \
            case 16: KERNEL<16><<<grid, THREADS, 0, ctx->stream>>>(ARGS); break;                \
            case 8: KERNEL<8><<<grid, THREADS, 0, ctx->stream>>>(ARGS); break;                  \
            case 4: KERNEL<4><<<grid, THREADS, 0, ctx->stream>>>(ARGS); break;                  \
            case 2: KERNEL<2><<<grid, THREADS, 0, ctx->stream>>>(ARGS); break;                  \
            default: KERNEL<1><<<grid, THREADS, 0, ctx->stream>>>(ARGS); break;                 \
        }                                                                                       \
    } while (0)

static void launch_e(soup_ctx* ctx, float log_prior) {
    Workspace& w = ctx->ws;
    const uint32_t bpg = (uint32_t)((ctx->N + EM_WARPS - 1) / EM_WARPS);
    if (bpg == 0) return;
    const EArgs a{ctx->d_csr_idx, ctx->d_csr_aux, ctx->d_row_ptr, ctx->d_cell_order, w.LA, w.LB, w.ELL, w.e_group_active, w.ctl,
                  ctx->pk_csr, (uint32_t)ctx->N, w.Cpad, bpg, (uint32_t)(w.forced_lanes != 0), log_prior};
    const dim3 grid(bpg * w.ep.n_groups);
    SOUP_DISPATCH(estep_kernel, EM_THREADS, w.ep, a);
}
// One launch of the M pass over the loci [li_begin, li_end) of the LPT order (nullptr = the whole pass).  `with_long`: this launch
// also carries the long role's blocks and the cooperative long-locus kernel (both only ever touch the first loci of the order).
struct MRange {
    uint32_t li_begin, li_end;
    bool with_long;
    float* SP;                 // cell-sharded mode: the statistics buffer (MArgs::SP)
    uint32_t scols;
    const uint32_t* order;     // visiting order of the loci (cell-sharded mode: range-major), nullptr = the LPT order
};
static void launch_m(soup_ctx* ctx, const MRange* range = nullptr) {
    Workspace& w = ctx->ws;
    if (ctx->L == 0) return;
    const uint32_t n_long = ctx->n_long;
    const uint32_t n_tail = ctx->n_long_tail;
    MRange all{n_long, (uint32_t)ctx->L, true, nullptr, 0, nullptr};
    const MRange& r = range ? *range : all;
    const uint32_t li_begin = std::max(r.li_begin, n_long);     // (loci in front of n_long are never the wide role's)
    const uint32_t long_items = r.with_long ? n_long * (uint32_t)w.lp.n_groups : 0u;
    const bool coop = r.with_long && ctx->long_coop != 0 && n_tail > 0;   // long loci by mstep_long_kernel on its own high-priority stream (MArgs::coop_mode)
    const uint32_t long_blocks = (long_items + EM_WARPS - 1) / EM_WARPS;
    const uint32_t wide_chunks = r.li_end > li_begin ? (r.li_end - li_begin + EM_WARPS - 1) / EM_WARPS : 0u;
    const MArgs a{ctx->d_csc_idx, ctx->d_csc_aux, ctx->d_col_ptr, r.order ? r.order : ctx->d_locus_order, w.W, w.TH, w.LA, w.LB, w.col_write,
                  w.m_group_active, w.l_group_active, w.ctl, r.SP, r.scols, li_begin, r.li_end,
                  ctx->pk_csc, (uint32_t)ctx->L, w.Cpad, (uint32_t)w.mp.n_groups,
                  n_long, (uint32_t)w.lp.n_groups, long_blocks, wide_chunks,
                  (uint32_t)((uint64_t)ctx->N * w.Cpad * 4 > (48ull << 20)), (uint32_t)(w.forced_lanes != 0),
                  ctx->long_coop != 0 && n_tail > 0 ? (uint32_t)ctx->long_coop : 0u, ctx->tail_cols, n_tail, w.k_of_draw, (uint32_t)w.K, (uint32_t)(w.nd > 0 ? w.nd : w.K)};
    if (coop) {
        const uint32_t groups256 = (w.Cpad + ML_THREADS - 1) / ML_THREADS;
        const uint32_t blocks = ctx->long_coop == 2 ? std::max(n_tail, n_long * groups256) : n_tail;
        cudaEventRecord(ctx->ev_mfork, ctx->stream);
        cudaStreamWaitEvent(ctx->side_long, ctx->ev_mfork, 0);
        mstep_long_kernel<<<blocks, ML_THREADS, ML_SMEM_FLOATS * sizeof(float), ctx->side_long>>>(a);
        cudaEventRecord(ctx->ev_mjoin, ctx->side_long);
    }
    const dim3 grid(long_blocks + wide_chunks * w.mp.n_groups);
    if (grid.x != 0) SOUP_DISPATCH(mstep_kernel, EM_THREADS, w.mp, a);
    if (coop) cudaStreamWaitEvent(ctx->stream, ctx->ev_mjoin, 0);
}
static void launch_post(soup_ctx* ctx, int method, bool counts) {
    Workspace& w = ctx->ws;
    PostArgs a{w.ELL, w.W, w.LSE, ctx->d_total_alleles, w.st, w.ctl, counts ? w.counts : nullptr, (uint32_t)ctx->N, w.Cpad, w.K_out, w.slots, method,
               w.K, w.nd > 0 ? w.nd : w.K, w.mapped ? w.src_of_k : nullptr, w.mapped ? w.w_of_k : nullptr, w.ELLfix, w.Kf};
    if (ctx->N) post_kernel<<<dim3((unsigned)((ctx->N + POST_WARPS - 1) / POST_WARPS), (unsigned)((w.slots + 31) / 32)), 32 * POST_WARPS, 0, ctx->stream>>>(a);
}
static void launch_refill(soup_ctx* ctx, const RefillArgs& a) {
    Workspace& w = ctx->ws;
    const size_t n = (size_t)ctx->L * w.K;
    // a fixed, small number of CTAs per slot (they loop): the launch is a no-op on most iterations
    if (n) refill_kernel<<<dim3((unsigned)std::min<size_t>((n + 255) / 256, 48), (unsigned)w.slots), 256, 0, ctx->stream>>>(a);
}

// kernels one iteration enqueues: E, post, loss, control, M (+ the long-locus kernel), harvest, finalize, migrate
// (+ refill while the queue is not known to be drained, + theta in cell-sharded mode)
static uint64_t launches_per_iteration(const soup_ctx* ctx, bool refill, bool cell_mode) {
    return 8 + (ctx->long_coop != 0 && ctx->n_long_tail > 0 ? 1 : 0) + (refill ? 1 : 0) + (cell_mode ? 2 * (uint64_t)std::max(ctx->cell_ranges, 1) - 1 : 0);
}

static void launch_loss_control(soup_ctx* ctx, cudaStream_t st, bool counts, float limit) {
    Workspace& w = ctx->ws;
    loss_kernel<<<w.slots, 256, 0, st>>>(w.LSE, w.st, w.loss, (uint32_t)ctx->N);
    ControlArgs c{w.loss, w.st, w.ctl, counts ? w.counts : nullptr, w.trace, w.solve_loss, w.solve_iters,
                  (uint32_t)ctx->N, w.K_out, w.slots, w.iter_cap, limit};
    control_kernel<<<1, std::min(1024, (w.slots + 31) / 32 * 32), 0, st>>>(c);
}

// E -> post -> { loss + control on the side stream  ||  M on the main stream } -> harvest -> refill -> finalize.
// The M pass does not depend on the loss, so the strictly sequential loss sum hides behind it.
// Cell-sharded mode (each rank holds a range of cells): E -> post -> partial loss -> partial M -> ONE all-reduce of
// [S | D | loss] over the ranks (callback, enqueued on this stream) -> theta -> control -> ...; every rank then takes
// the same control decisions from the same reduced numbers.
// NVTX ranges around the phases of an iteration as the host enqueues them (nsys projects them onto the kernels they launched)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

static void enqueue_iteration(IterLaunch& it) {
    soup_ctx* ctx = it.ctx;
    Workspace& w = ctx->ws;
    NvtxRange r_iter("soup iteration");
    cudaEvent_t* ev = it.timed_iter >= 0 ? ctx->ev_k.data() + (size_t)it.timed_iter * 4 : nullptr;
    if (ev) cudaEventRecord(ev[0], ctx->stream);
    { NvtxRange r("E pass (binomial_loss)"); launch_e(ctx, it.log_prior); }
    if (ev) cudaEventRecord(ev[1], ctx->stream);
    { NvtxRange r("post (log_sum_exp, weights)"); launch_post(ctx, it.method, it.record_counts); }
    if (it.cell_mode) {
        NvtxRange r("M pass + all-reduce + theta (cell-sharded)");
        // statistics buffer [L][2][scols] (MArgs::SP), the slots' partial losses behind it.  The M pass runs as n_ranges launches over
        // ranges of locus ids (cell_order: the long loci first, then range by range, longest first inside a range); as soon as a
        // range is complete its all-reduce and theta pass start on comm_stream while the next range is being produced.
        const uint32_t L = (uint32_t)ctx->L, scols = it.scols;
        const int n_ranges = ctx->cell_ranges;
        float* loss = w.stats + 2 * (size_t)L * scols;
        cudaEventRecord(ctx->ev_fork, ctx->stream);                 // the partial loss sums hide behind the M pass
        cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0);
        loss_kernel<<<w.slots, 256, 0, ctx->side>>>(w.LSE, w.st, loss, (uint32_t)ctx->N);
        cudaEventRecord(ctx->ev_loss, ctx->side);
        if (ev) cudaEventRecord(ev[2], ctx->stream);
        for (int c = 0; c < n_ranges; ++c) {
            const uint32_t row0 = ctx->cell_range_row[c], rows = ctx->cell_range_row[c + 1] - row0;
            const MRange range{ctx->cell_range_pos[c], ctx->cell_range_pos[c + 1], c == 0, w.stats, scols, (const uint32_t*)ctx->lb.cell_order_m.p};
            launch_m(ctx, &range);
            cudaEventRecord(ctx->ev_chunk[c], ctx->stream);
            cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_chunk[c], 0);
            const bool last = c == n_ranges - 1;
            if (last) cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_loss, 0);
            float* sp = w.stats + 2 * (size_t)row0 * scols;
            const uint64_t n_floats = 2 * (uint64_t)rows * scols + (last ? (uint64_t)w.slots : 0u);
            int32_t e = SOUP_OK;
            if (it.comm) {
                if (ncclAllReduce(sp, sp, n_floats, ncclFloat, ncclSum, it.comm->comm, ctx->comm_stream) != ncclSuccess) e = SOUP_ERR_CUDA;
            } else {
                e = it.allreduce(it.comm_user, sp, n_floats, (void*)ctx->comm_stream);
            }
            if (e != SOUP_OK && it.comm_error == 0) it.comm_error = e;
            it.allreduce_bytes += n_floats * 4; it.allreduce_calls += 1;
            if (rows) theta_kernel<<<ctx->sm_count * 4, 256, 0, ctx->comm_stream>>>(w.stats, scols, row0, rows, w.col_write, w.TH, w.LA, w.LB, w.Cpad);
        }
        if (ev) cudaEventRecord(ev[3], ctx->stream);
        cudaEventRecord(ctx->ev_comm, ctx->comm_stream);
        cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0);
        ControlArgs c{loss, w.st, w.ctl, nullptr, w.trace, w.solve_loss, w.solve_iters, (uint32_t)ctx->N, w.K, w.slots, w.iter_cap, it.limit};
        control_kernel<<<1, std::min(1024, (w.slots + 31) / 32 * 32), 0, ctx->stream>>>(c);
    } else {
        cudaEventRecord(ctx->ev_fork, ctx->stream);
        cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0);
        { NvtxRange r("loss + control (side stream)"); launch_loss_control(ctx, ctx->side, it.record_counts, it.limit); }
        cudaEventRecord(ctx->ev_join, ctx->side);
        if (ev) cudaEventRecord(ev[2], ctx->stream);
        { NvtxRange r("M pass (update_centers)"); launch_m(ctx); }
        if (ev) cudaEventRecord(ev[3], ctx->stream);
        cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
    }
    NvtxRange r_tail("harvest + refill + finalize");
    harvest_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(w.ctl, w.ELL, w.TH, w.best_lp, w.best_theta, (uint32_t)ctx->N, (uint32_t)ctx->L, w.Cpad, w.K_out, w.K,
                                                               w.mapped ? w.src_of_k : nullptr, w.ELLfix, w.Kf, it.base_theta_dev);
    if (!it.queue_drained) launch_refill(ctx, it.refill);
    finalize_kernel<<<1, 1024, 0, ctx->stream>>>(it.fin);
    migrate_kernel<<<dim3(96, (unsigned)w.slots), 256, 0, ctx->stream>>>(w.ctl, w.moves, w.TH, w.LA, w.LB, (uint32_t)ctx->L, w.Cpad, w.K);
}

// Cell-sharded mode: the M pass's visiting order of the loci.  The statistics travel in ranges of locus ids (the same on every rank),
// so the order is: the long loci (the first n_long_tail of this rank's LPT order, wherever their ids lie: the first launch finishes
// them), then range by range, longest first inside a range.
static int32_t build_cell_order(soup_ctx* ctx, int requested) {
    const uint32_t L = (uint32_t)ctx->L;
    int P = requested > 0 ? requested : (L < 4096 ? 1 : 4);       // (a function of the caller's choice and L alone: every rank must cut the same ranges)
    if (const char* e = getenv("SOUP_CELL_RANGES")) P = atoi(e);  // tuning / tests; the same on every rank
    P = std::max(1, std::min(P, (int)soup_ctx::MAX_CELL_CHUNKS));
    if ((uint32_t)P > std::max(L, 1u)) P = (int)std::max(L, 1u);
    if (ctx->cell_ranges == P) return SOUP_OK;                    // (reset by soup_load_csr)
    std::vector<uint32_t> lpt(std::max<uint32_t>(L, 1)), order(std::max<uint32_t>(L, 1));
    CK(cudaSetDevice(ctx->device));
    if (L) CK(cudaMemcpy(lpt.data(), ctx->d_locus_order, (size_t)L * 4, cudaMemcpyDeviceToHost));
    const uint32_t rows = (L + (uint32_t)P - 1) / (uint32_t)P, prefix = std::min(L, std::max(ctx->n_long, ctx->n_long_tail));
    std::vector<uint32_t> count(P + 1, 0);
    for (uint32_t i = prefix; i < L; ++i) count[lpt[i] / std::max(rows, 1u) + 1] += 1;
    for (int c = 0; c <= P; ++c) ctx->cell_range_row[c] = std::min(L, (uint32_t)c * rows);
    std::vector<uint32_t> at(P + 1, 0);
    at[0] = prefix;
    for (int c = 0; c < P; ++c) at[c + 1] = at[c] + count[c + 1];
    ctx->cell_range_pos[0] = std::min(ctx->n_long, prefix);
    for (int c = 1; c <= P; ++c) ctx->cell_range_pos[c] = at[c];
    for (uint32_t i = 0; i < prefix; ++i) order[i] = lpt[i];
    for (uint32_t i = prefix; i < L; ++i) order[at[lpt[i] / std::max(rows, 1u)]++] = lpt[i];
    if (ctx->lb.cell_order_m.ensure(std::max<size_t>(L, 1) * 4) != cudaSuccess) { cudaGetLastError(); return ctx->fail(SOUP_ERR_NOMEM, "soup_solve: cannot allocate the cell-sharded visiting order"); }
    if (L) CK(cudaMemcpy(ctx->lb.cell_order_m.p, order.data(), (size_t)L * 4, cudaMemcpyHostToDevice));
    ctx->cell_ranges = P;
    return SOUP_OK;
}

// ------------------------------------------------------------------ soup_solve
extern "C" int32_t soup_solve(soup_ctx* ctx, const soup_solve_params* p, soup_solve_result* r) try {
    if (!ctx) return SOUP_ERR_INVALID;
    if (!p || !r) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: null argument");
    if (!ctx->loaded) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: no matrix loaded (call soup_load_csr first)");
    if (p->k <= 0) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: k must be >= 1");
    if (p->method != SOUP_METHOD_EM && p->method != SOUP_METHOD_KHM) return ctx->fail(SOUP_ERR_INVALID, "Clustering method must be one of em, khm");
    if (p->n_streams <= 0 || p->solves_per_stream <= 0) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: n_streams and solves_per_stream must be >= 1");
    if (!p->theta0 && !p->stream_seeds && !p->init_theta) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: need stream_seeds, theta0 or init_theta");
    if (p->n_reinit < 0 || p->n_reinit > p->k || (p->n_reinit > 0 && (!p->reinit_clusters || !p->base_theta)))
        return ctx->fail(SOUP_ERR_INVALID, "soup_solve: bad reinit_clusters / base_theta");
    if ((uint64_t)p->n_streams * (uint64_t)p->solves_per_stream > (1u << 30)) return ctx->fail(SOUP_ERR_UNSUPPORTED, "soup_solve: too many solves");
    CK(cudaSetDevice(ctx->device));
    const int Kt = p->k;   // clusters of the model
    // Compact re-run (souporcell3 runs 1 and 2, S:92-97): every solve of the run shares the locked clusters' centers (base_theta), so
    // their log-likelihood columns are computed once (ELLfix) and the engine — tables, E pass, M pass, memory per slot — holds only the
    // re-drawn clusters.  Bit-identical: per-cluster arithmetic does not depend on the other clusters.  (With explicit theta0 the locked
    // clusters may differ per solve: those calls keep all K columns and only drop the locked ones from the M pass.)
    const bool compact_run = p->n_reinit > 0 && p->n_reinit < Kt && !p->theta0 && p->base_theta && !getenv("SOUP_NO_COMPACT_RUN");
    const int K = compact_run ? p->n_reinit : Kt;   // clusters per slot in the engine
    const int Kf = compact_run ? Kt - K : 0;
    const int n_solves = p->n_streams * p->solves_per_stream;
    const int world = p->shard_world > 1 ? p->shard_world : 1, rank = world > 1 ? p->shard_rank : 0;
    if (rank < 0 || rank >= world) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: shard_rank out of range");
    const int iter_cap = p->iteration_cap > 0 ? p->iteration_cap : 1000;
    std::vector<int> queue;
    for (int s = rank; s < n_solves; s += world) queue.push_back(s);
    const int n_local = (int)queue.size();

    ctx->stats = soup_stats{};
    r->has_best = 0; r->best_solve = -1; r->best_loss = -INFINITY; r->nnz_updates = 0;
    if (r->solve_loss) for (int i = 0; i < n_solves; ++i) r->solve_loss[i] = NAN;
    if (r->solve_iters) for (int i = 0; i < n_solves; ++i) r->solve_iters[i] = 0;
    ctx->h_trace.clear(); ctx->h_trace_solve.clear(); ctx->h_trace_iters.clear(); ctx->h_trace_cap = iter_cap;

    const bool cell_mode = p->cell_allreduce != nullptr || p->cell_comm != nullptr;
    if (p->cell_comm && p->cell_comm->device != ctx->device) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: cell_comm belongs to another device");
    // cell_mode with world > 1 is the hybrid mode: every process of a cell group passes the same shard_rank (its
    // group), runs the same subset of the solves, and the all-reduce callback spans that group only
    if (cell_mode && p->n_cells_global < ctx->N) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: n_cells_global must be the total cell count over all ranks");
    const float limit = 0.01f * (float)(cell_mode ? p->n_cells_global : ctx->N);   // S:214 (all cells)
    const float log_prior = logf(1.0f / (float)Kt);            // S:202 (host libm = the reference's f32::ln)
    const bool any_iteration = (10000.0f > limit) && iter_cap > 0;
    if (n_local == 0 || !any_iteration) {
        // em()/khm() with no iteration returns (-inf, empty): nothing can beat -inf (S:65, S:110)
        if (r->solve_loss) for (int s : queue) r->solve_loss[s] = -INFINITY;
        return SOUP_OK;
    }

    // ---- shape: resident slots and vector width
    int slots = p->max_slots > 0 ? p->max_slots : std::max(1, 2048 / K);
    slots = std::min(slots, n_local);
    PhaseTimer pt;
    // a workspace of exactly this shape already exists (a repeated job): nothing to size, and cudaMemGetInfo — a driver
    // call that was seen to take tens of milliseconds now and then — stays off the repeated path
    const bool reuse = !p->theta0 && ctx->ws.K == K && ctx->ws.K_out == Kt && ctx->ws.min_cols == (uint32_t)Kf && ctx->ws.slots == std::min(slots, 1024) &&
                       ctx->ws.iter_cap == iter_cap && ctx->ws.queue_cap >= n_local && ctx->ws.n_streams_cap >= p->n_streams;
    if (!reuse) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        free_b += ctx->ws.bytes;
        size_t fixed = (size_t)Kt * (ctx->L + 2 * ctx->N) * 4 + (size_t)n_local * iter_cap * sizeof(IterRec) + (64u << 20);
        if (p->theta0) fixed += (size_t)n_solves * Kt * ctx->L * 4;
        size_t per = bytes_per_slot(ctx, K);
        if (free_b < fixed + per) return ctx->fail(SOUP_ERR_NOMEM, "soup_solve: not enough device memory for one resident solve");
        size_t cap = (size_t)((double)(free_b - fixed) * 0.85 / (double)per);
        slots = (int)std::min<size_t>((size_t)slots, std::max<size_t>(cap, 1));
    }
    slots = std::min(slots, 1024);
    // table / weight rows are addressed with 32-bit element offsets
    while (slots > 1 && (uint64_t)std::max(ctx->L, ctx->N) * ((uint64_t)slots * K + 128) >= (1ull << 32)) --slots;
    if ((uint64_t)std::max(ctx->L, ctx->N) * ((uint64_t)slots * K + 128) >= (1ull << 32))
        return ctx->fail(SOUP_ERR_UNSUPPORTED, "soup_solve: max(cells, loci) * k exceeds the 32-bit row-offset range");
    pt.lap("solve: shape");
    const int forced_lanes = (p->lanes == 1 || p->lanes == 2 || p->lanes == 4 || p->lanes == 8 || p->lanes == 16) ? p->lanes : 0;
    int32_t rc = ensure_workspace(ctx, K, slots, forced_lanes, iter_cap, n_local, p->n_streams, Kt, (uint32_t)Kf);
    if (rc != SOUP_OK) return rc;
    pt.lap("solve: workspace");
    Workspace& w = ctx->ws;
    cudaStream_t st = ctx->stream;
    if (cell_mode) {
        rc = build_cell_order(ctx, p->cell_ranges);
        if (rc != SOUP_OK) return rc;
        const size_t need = 2 * (size_t)ctx->L * w.Cpad + (size_t)slots;
        if (w.stats_floats != need) {
            float* sp = nullptr;
            cudaError_t e = ws_alloc(w, st, &sp, need, true);
            if (e != cudaSuccess) { cudaGetLastError(); return ctx->fail(SOUP_ERR_NOMEM, "soup_solve: cannot allocate the all-reduce buffer: %s", cudaGetErrorString(e)); }
            w.stats = sp;
            w.stats_floats = need;
        }
    } else {
        w.stats = nullptr;      // (kept allocated inside the workspace; not used by the M pass in this mode)
        w.stats_floats = 0;
    }

    // ---- per-call device inputs
    float *d_theta0 = nullptr, *d_base = nullptr, *d_init = nullptr;
    uint16_t* d_assign_free = nullptr;
    uint64_t* d_words_free = nullptr;
    auto cleanup = [&] { cudaFree(d_theta0); cudaFree(d_base); cudaFree(d_init); cudaFree(d_assign_free); cudaFree(d_words_free); };
#define CKS(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { cleanup(); return ctx->fail(e_ == cudaErrorMemoryAllocation ? SOUP_ERR_NOMEM : SOUP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    std::vector<int> draw_of_k(K);
    std::vector<uint8_t> k_locked(K, 0);
    std::vector<int> src_of_k(Kt), w_of_k(Kt), fix_k(Kt, 0);
    for (int k = 0; k < Kt; ++k) src_of_k[k] = w_of_k[k] = k;
    int n_draw = K;
    bool mapped = false;
    if (p->n_reinit > 0) {   // S:92-97: re-draw the flagged clusters in list order, lock the rest
        std::vector<int> pos(Kt, -1);
        for (int i = 0; i < p->n_reinit; ++i) {
            int c = p->reinit_clusters[i];
            if (c < 0 || c >= Kt || pos[c] >= 0) return ctx->fail(SOUP_ERR_INVALID, "soup_solve: bad reinit_clusters entry");
            pos[c] = i;
        }
        n_draw = p->n_reinit;
        if (compact_run) {   // the engine's cluster d is the model's cluster reinit_clusters[d]; nothing is locked inside the engine
            for (int d = 0; d < K; ++d) draw_of_k[d] = d;
            int j = 0;
            for (int k = 0; k < Kt; ++k) {
                if (pos[k] >= 0) { src_of_k[k] = w_of_k[k] = pos[k]; }
                else { src_of_k[k] = -(j + 1); w_of_k[k] = -1; fix_k[j++] = k; }
            }
            mapped = true;
        } else {
            for (int k = 0; k < Kt; ++k) { draw_of_k[k] = pos[k]; k_locked[k] = pos[k] < 0; }
            if (!getenv("SOUP_NO_COMPACT_W")) { for (int k = 0; k < Kt; ++k) w_of_k[k] = pos[k]; mapped = true; }
        }
    } else {
        for (int k = 0; k < K; ++k) draw_of_k[k] = k;
    }
    uint64_t h2d = 0;
    if (p->theta0) {
        size_t n = (size_t)n_solves * Kt * ctx->L;
        CKS(cudaMalloc(&d_theta0, std::max<size_t>(n, 1) * 4));
        CKS(cudaMemcpyAsync(d_theta0, p->theta0, n * 4, cudaMemcpyHostToDevice, st));
        h2d += n * 4;
    }
    if (p->base_theta && p->n_reinit > 0) {
        size_t n = (size_t)Kt * ctx->L;
        CKS(cudaMalloc(&d_base, std::max<size_t>(n, 1) * 4));
        CKS(cudaMemcpyAsync(d_base, p->base_theta, n * 4, cudaMemcpyHostToDevice, st));
        h2d += n * 4;
    }
    if (p->init_theta) {
        size_t n = (size_t)Kt * ctx->L;
        CKS(cudaMalloc(&d_init, std::max<size_t>(n, 1) * 4));
        CKS(cudaMemcpyAsync(d_init, p->init_theta, n * 4, cudaMemcpyHostToDevice, st));
        h2d += n * 4;
    }
    std::vector<uint32_t> keys((size_t)p->n_streams * 8, 0);
    if (p->stream_seeds)
        for (int t = 0; t < p->n_streams; ++t) seed_to_key(p->stream_seeds + (size_t)t * 32, keys.data() + (size_t)t * 8);
    CKS(cudaMemcpyAsync(w.stream_keys, keys.data(), keys.size() * 4, cudaMemcpyHostToDevice, st));
    CKS(cudaMemcpyAsync(w.queue, queue.data(), (size_t)n_local * 4, cudaMemcpyHostToDevice, st));
    CKS(cudaMemcpyAsync(w.draw_of_k, draw_of_k.data(), (size_t)K * 4, cudaMemcpyHostToDevice, st));
    {   // weight columns: only the clusters this run draws (S:92-97 locks the rest); in a compact re-run the engine holds nothing else anyway
        std::vector<int> k_of_draw(K, 0);
        for (int k = 0; k < K; ++k) if (draw_of_k[k] >= 0) k_of_draw[draw_of_k[k]] = k;
        w.nd = (!compact_run && mapped) ? n_draw : K;
        w.mapped = mapped; w.Kf = Kf;
        CKS(cudaMemcpyAsync(w.k_of_draw, k_of_draw.data(), (size_t)K * 4, cudaMemcpyHostToDevice, st));
        CKS(cudaMemcpyAsync(w.src_of_k, src_of_k.data(), (size_t)Kt * 4, cudaMemcpyHostToDevice, st));
        CKS(cudaMemcpyAsync(w.w_of_k, w_of_k.data(), (size_t)Kt * 4, cudaMemcpyHostToDevice, st));
        CKS(cudaMemcpyAsync(w.fix_k, fix_k.data(), (size_t)Kt * 4, cudaMemcpyHostToDevice, st));
    }
    if (compact_run) {   // ---- the locked clusters' log-likelihood columns, once for the whole run
        Ctl hc0{};
        hc0.copy_slot = -1; hc0.best_solve = -1; hc0.best_loss = -INFINITY; hc0.live_cols = (uint32_t)Kf;
        CKS(cudaMemcpyAsync(w.ctl, &hc0, sizeof(Ctl), cudaMemcpyHostToDevice, st));
        std::vector<int> ones(w.ep.n_groups, 1);
        CKS(cudaMemcpyAsync(w.e_group_active, ones.data(), (size_t)w.ep.n_groups * 4, cudaMemcpyHostToDevice, st));
        const size_t ne = (size_t)ctx->L * Kf;
        if (ne) fixed_tables_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(d_base, w.fix_k, w.LA, w.LB, w.ctl, (uint32_t)ctx->L, w.Cpad, Kf);
        launch_e(ctx, log_prior);
        const size_t nc = (size_t)ctx->N * Kf;
        if (nc) copy_fixed_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(w.ELL, w.ELLfix, (uint32_t)ctx->N, w.Cpad, Kf);
        CKS(cudaStreamSynchronize(st));   // `ones` goes out of scope
        CKS(cudaGetLastError());
    }
    CKS(cudaMemcpyAsync(w.k_locked, k_locked.data(), (size_t)K, cudaMemcpyHostToDevice, st));
    h2d += keys.size() * 4 + (size_t)n_local * 4 + (size_t)K * 5;
    // ---- init_cluster_centers_random_assignment S:565-594: place the three phases of every solve in its seed stream.
    // Solve `it` of a stream starts where solve it-1 ended; the middle phase (one gen_range per cell, S:577) consumes a
    // data-dependent number of words, so the host walks it (N words per solve) and the device does the rest.
    uint16_t* d_assign = nullptr;
    uint64_t* d_words = nullptr;
    const bool random_assignment = p->init_strategy == SOUP_INIT_RANDOM_ASSIGNMENT && !p->theta0 && !p->init_theta;
    if (p->init_strategy != SOUP_INIT_RANDOM_UNIFORM && p->init_strategy != SOUP_INIT_RANDOM_ASSIGNMENT)
    { cleanup(); return ctx->fail(SOUP_ERR_INVALID, "soup_solve: unknown init_strategy %d", p->init_strategy); }
    if (random_assignment) {
        if (cell_mode) { cleanup(); return ctx->fail(SOUP_ERR_UNSUPPORTED, "soup_solve: random_cell_assignment needs every cell on one GPU (not available with cell sharding)"); }
        if (!p->stream_seeds) { cleanup(); return ctx->fail(SOUP_ERR_INVALID, "soup_solve: random_cell_assignment needs stream_seeds"); }
        if (n_draw > 65535) { cleanup(); return ctx->fail(SOUP_ERR_UNSUPPORTED, "soup_solve: random_cell_assignment supports at most 65535 clusters"); }
        const uint64_t N = ctx->N, phase = (uint64_t)n_draw * ctx->L;
        std::vector<uint16_t> assign((size_t)n_local * N);
        std::vector<uint64_t> words((size_t)n_local * 2);
        std::vector<int> local_of(n_solves, -1);
        for (int i = 0; i < n_local; ++i) local_of[queue[i]] = i;
        for (int t = 0; t < p->n_streams; ++t) {
            uint64_t pos = 0;
            uint32_t blk[16];
            uint64_t have = UINT64_MAX;
            auto word_at = [&](uint64_t wpos) {
                if ((wpos >> 4) != have) { have = wpos >> 4; chacha_block(keys.data() + (size_t)t * 8, have, 12, blk); }
                return blk[wpos & 15];
            };
            for (int it2 = 0; it2 < p->solves_per_stream; ++it2) {
                const int li = local_of[t * p->solves_per_stream + it2];
                const uint64_t w_sums = pos;
                pos += phase;
                for (uint64_t c = 0; c < N; ++c) {                       // rand 0.9.3 gen_range(0..n_draw): Canon's method on u32
                    const uint64_t m = (uint64_t)word_at(pos++) * (uint32_t)n_draw;
                    uint32_t res = (uint32_t)(m >> 32);
                    const uint32_t lo = (uint32_t)m;
                    if (lo > (uint32_t)(0u - (uint32_t)n_draw)) {
                        const uint32_t hi2 = (uint32_t)(((uint64_t)word_at(pos++) * (uint32_t)n_draw) >> 32);
                        if ((uint64_t)lo + hi2 > 0xffffffffull) res += 1;
                    }
                    if (li >= 0) assign[(size_t)li * N + c] = (uint16_t)res;
                }
                if (li >= 0) { words[(size_t)li] = w_sums; words[(size_t)n_local + li] = pos; }
                pos += phase;
            }
        }
        CKS(cudaMalloc(&d_assign, std::max<size_t>(assign.size(), 1) * 2));
        d_assign_free = d_assign;
        CKS(cudaMalloc(&d_words, std::max<size_t>(words.size(), 1) * 8));
        d_words_free = d_words;
        CKS(cudaMemcpyAsync(d_assign, assign.data(), assign.size() * 2, cudaMemcpyHostToDevice, st));
        CKS(cudaMemcpyAsync(d_words, words.data(), words.size() * 8, cudaMemcpyHostToDevice, st));
        CKS(cudaStreamSynchronize(st));                                   // the staging vectors go out of scope here
        h2d += assign.size() * 2 + words.size() * 8;
    }
    // initial slot table: every slot "finished" with its first queue entry pending -> refill + finalize install them
    std::vector<SlotState> hst(slots);
    for (int s = 0; s < slots; ++s) {
        hst[s] = SlotState{};
        hst[s].solve = -1; hst[s].fin = 1; hst[s].pending = s < n_local ? s : -1;
        hst[s].last_loss = -INFINITY; hst[s].loss = -INFINITY;
    }
    Ctl hc{};
    hc.queue_head = std::min(slots, n_local); hc.queue_len = n_local; hc.copy_slot = -1; hc.best_solve = -1; hc.best_loss = -INFINITY;
    CKS(cudaMemcpyAsync(w.st, hst.data(), sizeof(SlotState) * slots, cudaMemcpyHostToDevice, st));
    CKS(cudaMemcpyAsync(w.ctl, &hc, sizeof(Ctl), cudaMemcpyHostToDevice, st));
    CKS(cudaMemsetAsync(w.counts, 0, (size_t)slots * Kt * 4, st));

    IterLaunch it;
    it.ctx = ctx; it.method = p->method; it.log_prior = log_prior; it.limit = limit; it.record_counts = p->record_counts != 0 && !cell_mode;
    it.allreduce = p->cell_allreduce; it.comm_user = p->comm_user; it.comm = p->cell_comm; it.cell_mode = cell_mode;
    auto cover4 = [&](uint32_t live) { return std::min(w.Cpad, (std::max(live, 1u) + 3u) & ~3u); };
    it.scols = cover4((uint32_t)(std::min(slots, n_local) * K));
    it.base_theta_dev = d_base;
    it.refill = RefillArgs{w.st, w.ctl, w.queue, w.stream_keys, d_theta0, d_base, d_init, w.draw_of_k, w.TH, w.LA, w.LB,
                           (uint32_t)ctx->L, w.Cpad, K, n_draw, p->solves_per_stream,
                           d_assign, d_words, d_words ? d_words + n_local : nullptr, ctx->d_csc_idx, ctx->d_csc_aux, ctx->d_col_ptr, ctx->pk_csc, (uint32_t)ctx->N};
    it.fin = FinalizeArgs{w.st, w.ctl, w.queue, w.k_locked, w.col_write,
                          {{w.e_group_active, 32 * w.ep.VM, 32 * w.ep.VT, w.ep.n_main, w.ep.n_groups},
                           {w.m_group_active, 32 * w.mp.VM, 32 * w.mp.VT, w.mp.n_main, w.mp.n_groups},
                           {w.l_group_active, 32 * w.lp.VM, 32 * w.lp.VT, w.lp.n_main, w.lp.n_groups}},
                          K, slots, iter_cap, w.Cpad, limit, w.moves, getenv("SOUP_NO_COMPACT") ? 0 : 1};

    pt.lap("solve: setup uploads");
    CKS(cudaEventRecord(ctx->ev0, st));
    launch_refill(ctx, it.refill);
    finalize_kernel<<<1, 1024, 0, st>>>(it.fin);
    CKS(cudaGetLastError());
    uint64_t launches = 2;

    // ---- iterate: chunk i+1 is enqueued before the host waits for chunk i's done counter, so the
    // GPU never idles on the poll; kernels of chunks past the end find no active slot and exit.
    int chunk = p->poll_chunk > 0 ? p->poll_chunk : 4;
    if (const char* e = getenv("SOUP_POLL_CHUNK")) chunk = std::max(1, atoi(e));   // tuning knob
    uint64_t iters_enqueued = 0;
    int polls = 0;
    const size_t max_timed = 4096;
    if (p->time_kernels) {
        while (ctx->ev_k.size() < max_timed * 4) {
            cudaEvent_t e = nullptr;
            CKS(cudaEventCreate(&e));
            ctx->ev_k.push_back(e);
        }
    }
    size_t n_timed = 0;
    auto enqueue_chunk = [&](int which) -> cudaError_t {
        for (int i = 0; i < chunk; ++i) {
            it.timed_iter = (p->time_kernels && n_timed < max_timed) ? (int)n_timed++ : -1;
            enqueue_iteration(it);
        }
        iters_enqueued += chunk;
        launches += (uint64_t)chunk * launches_per_iteration(ctx, !it.queue_drained, it.cell_mode);
        cudaError_t e = cudaMemcpyAsync(ctx->h_ctl + which, w.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_poll[which], st);
        return e;
    };
    const bool slow_dbg = getenv("SOUP_DEBUG_SLOW") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_loop0 = now_ms();
    double t_enq = 0, t_wait = 0, worst_enq = 0, worst_wait = 0;
    CKS(enqueue_chunk(0));
    for (;;) {
        const int cur = polls & 1;
        const double ta = slow_dbg ? now_ms() : 0;
        CKS(enqueue_chunk(cur ^ 1));
        const double tb = slow_dbg ? now_ms() : 0;
        CKS(cudaEventSynchronize(ctx->ev_poll[cur]));
        if (slow_dbg) { const double tc = now_ms(); t_enq += tb - ta; t_wait += tc - tb; worst_enq = std::max(worst_enq, tb - ta); worst_wait = std::max(worst_wait, tc - tb); }
        ++polls;
        if (it.comm_error) { cleanup(); return ctx->fail(it.comm_error, "soup_solve: the all-reduce callback failed (%d)", it.comm_error); }
        if (ctx->h_ctl[cur].done_count >= n_local) break;
        if (ctx->h_ctl[cur].queue_head >= ctx->h_ctl[cur].queue_len) it.queue_drained = true;
        if (cell_mode) it.scols = std::min(it.scols, cover4(ctx->h_ctl[cur].live_cols));   // (the same poll on every rank; live_cols never grows)
        if (iters_enqueued > (uint64_t)n_local * (uint64_t)(iter_cap + TEMP_STEPS) + 64) { cleanup(); return ctx->fail(SOUP_ERR_CUDA, "soup_solve: device control loop did not converge"); }
    }
    const double t_loop1 = now_ms();
    CKS(cudaEventRecord(ctx->ev1, st));
    CKS(cudaStreamSynchronize(st));
    CKS(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    if (slow_dbg && now_ms() - t_loop0 > ms + 3.0)
        fprintf(stderr, "soup_slow iterate: host %.2f ms (loop %.2f, final sync %.2f) vs device %.2f ms; polls %d, enqueue total %.2f worst %.2f, wait total %.2f worst %.2f\n",
                now_ms() - t_loop0, t_loop1 - t_loop0, now_ms() - t_loop1, ms, polls, t_enq, worst_enq, t_wait, worst_wait);

    pt.lap("solve: iterate");
    // ---- results
    Ctl fc;
    CKS(cudaMemcpy(&fc, w.ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    std::vector<float> h_loss(n_local);
    std::vector<int> h_iters(n_local);
    CKS(cudaMemcpy(h_loss.data(), w.solve_loss, (size_t)n_local * 4, cudaMemcpyDeviceToHost));
    CKS(cudaMemcpy(h_iters.data(), w.solve_iters, (size_t)n_local * 4, cudaMemcpyDeviceToHost));
    uint64_t d2h = sizeof(Ctl) * (polls + 1) + (size_t)n_local * 8;
    uint64_t total_iters = 0;
    for (int i = 0; i < n_local; ++i) {
        total_iters += (uint64_t)h_iters[i];
        if (r->solve_loss) r->solve_loss[queue[i]] = h_loss[i];
        if (r->solve_iters) r->solve_iters[queue[i]] = h_iters[i];
    }
    r->has_best = fc.has_best;
    r->best_solve = fc.best_solve;
    r->best_loss = fc.best_loss;
    r->nnz_updates = ctx->nnz * total_iters;
    if (fc.has_best) {
        if (r->best_log_probs) { CKS(cudaMemcpy(r->best_log_probs, w.best_lp, (size_t)ctx->N * Kt * 4, cudaMemcpyDeviceToHost)); d2h += (size_t)ctx->N * Kt * 4; }
        if (r->best_theta) { CKS(cudaMemcpy(r->best_theta, w.best_theta, (size_t)ctx->L * Kt * 4, cudaMemcpyDeviceToHost)); d2h += (size_t)ctx->L * Kt * 4; }
    }
    // per-iteration trace (S:264 / S:354)
    if (p->keep_trace) {
        ctx->h_trace.resize((size_t)n_local * iter_cap);
        CKS(cudaMemcpy(ctx->h_trace.data(), w.trace, ctx->h_trace.size() * sizeof(IterRec), cudaMemcpyDeviceToHost));
        d2h += ctx->h_trace.size() * sizeof(IterRec);
        ctx->h_trace_solve = queue;
        ctx->h_trace_iters = h_iters;
    }
    cleanup();
    pt.lap("solve: results");
#undef CKS
    ctx->stats.kernel_launches = launches;
    ctx->stats.estep_launches = iters_enqueued;
    ctx->stats.mstep_launches = iters_enqueued;
    ctx->stats.solve_ms = ms;
    for (size_t i = 0; i < n_timed; ++i) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ctx->ev_k[i * 4], ctx->ev_k[i * 4 + 1]);
        cudaEventElapsedTime(&b, ctx->ev_k[i * 4 + 2], ctx->ev_k[i * 4 + 3]);
        ctx->stats.estep_ms += a;
        ctx->stats.mstep_ms += b;
    }
    ctx->stats.timed_iterations = n_timed;
    if (getenv("SOUP_DEBUG_ITERS")) {
        for (size_t i = 0; i < n_timed; ++i) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, ctx->ev_k[i * 4], ctx->ev_k[i * 4 + 1]);
            cudaEventElapsedTime(&b, ctx->ev_k[i * 4 + 2], ctx->ev_k[i * 4 + 3]);
            if (i + 1 < n_timed) cudaEventElapsedTime(&c, ctx->ev_k[i * 4], ctx->ev_k[i * 4 + 4]);
            fprintf(stderr, "soup_debug iter %zu E %.3f ms M %.3f ms total %.3f ms\n", i, a, b, c);
        }
    }
    ctx->stats.slot_iterations = fc.slot_iterations;
    ctx->stats.slots = slots; ctx->stats.lanes = w.ep.VM; ctx->stats.groups = w.ep.n_groups; ctx->stats.padded_cols = (int32_t)w.Cpad;
    ctx->stats.h2d_bytes = h2d; ctx->stats.d2h_bytes = d2h;
    ctx->stats.allreduce_bytes = it.allreduce_bytes; ctx->stats.allreduce_calls = it.allreduce_calls;
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_solve")

extern "C" int32_t soup_get_trace(soup_ctx* ctx, int32_t solve, soup_iter_record* out, int32_t cap, int32_t* n) try {
    if (!ctx || !n) return SOUP_ERR_INVALID;
    for (size_t i = 0; i < ctx->h_trace_solve.size(); ++i)
        if (ctx->h_trace_solve[i] == solve) {
            int m = ctx->h_trace_iters[i];
            *n = m;
            for (int j = 0; j < m && j < cap && out; ++j) {
                const IterRec& t = ctx->h_trace[i * (size_t)ctx->h_trace_cap + j];
                out[j] = soup_iter_record{t.temp_step, t.loss, t.change, t.above};
            }
            return SOUP_OK;
        }
    return ctx->fail(SOUP_ERR_INVALID, "soup_get_trace: solve %d was not run by the last soup_solve (or keep_trace was 0)", solve);
}
SOUP_CATCH_RETURN_CTX("soup_get_trace")

extern "C" int32_t soup_get_stats(soup_ctx* ctx, soup_stats* out) {
    if (!ctx || !out) return SOUP_ERR_INVALID;
    *out = ctx->stats;
    return SOUP_OK;
}

// ------------------------------------------------------------------ parity hooks
static int32_t hook_setup(soup_ctx* ctx, int k, int n_slots, int lanes) {
    if (!ctx->loaded) return ctx->fail(SOUP_ERR_INVALID, "no matrix loaded (call soup_load_csr first)");
    if (k <= 0 || n_slots <= 0 || n_slots > 1024) return ctx->fail(SOUP_ERR_INVALID, "bad k / n_slots");
    if ((uint64_t)std::max(ctx->L, ctx->N) * ((uint64_t)n_slots * k + 128) >= (1ull << 32))
        return ctx->fail(SOUP_ERR_UNSUPPORTED, "max(cells, loci) * columns exceeds the 32-bit row-offset range");
    const int forced = (lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8) ? lanes : 0;
    int32_t rc = ensure_workspace(ctx, k, n_slots, forced, 1, n_slots, 1);
    if (rc != SOUP_OK) return rc;
    Workspace& w = ctx->ws;
    w.stats = nullptr;          // the hooks always run the single-GPU M pass
    w.nd = k;                   // ... with a weight column for every cluster
    w.mapped = false; w.Kf = 0;
    w.stats_floats = 0;
    std::vector<int> ga(std::max(std::max(w.ep.n_groups, w.mp.n_groups), w.lp.n_groups), 1);
    cudaMemcpyAsync(w.l_group_active, ga.data(), (size_t)w.lp.n_groups * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(w.e_group_active, ga.data(), (size_t)w.ep.n_groups * 4, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(w.m_group_active, ga.data(), (size_t)w.mp.n_groups * 4, cudaMemcpyHostToDevice, ctx->stream);
    Ctl hc{};
    hc.copy_slot = -1; hc.best_solve = -1; hc.best_loss = -INFINITY; hc.live_cols = (uint32_t)(n_slots * k);
    cudaMemcpyAsync(w.ctl, &hc, sizeof(Ctl), cudaMemcpyHostToDevice, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return SOUP_OK;
}

extern "C" int32_t soup_estep(soup_ctx* ctx, int32_t k, int32_t method, int32_t n_slots, int32_t lanes, const float* theta,
                              const int32_t* temp_step, float* log_probs, float* weights, float* lse, float* loss,
                              int32_t* cells_per_cluster) try {
    if (!ctx) return SOUP_ERR_INVALID;
    if (!theta || !temp_step) return ctx->fail(SOUP_ERR_INVALID, "soup_estep: null argument");
    CK(cudaSetDevice(ctx->device));
    int32_t rc = hook_setup(ctx, k, n_slots, lanes);
    if (rc != SOUP_OK) return rc;
    Workspace& w = ctx->ws;
    cudaStream_t st = ctx->stream;
    const size_t nth = (size_t)n_slots * k * ctx->L, nlp = (size_t)n_slots * ctx->N * k;
    float *d_theta = nullptr, *d_out = nullptr;
    auto cleanup = [&] { cudaFree(d_theta); cudaFree(d_out); };
#define CKH(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { cleanup(); return ctx->fail(SOUP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    CKH(cudaMalloc(&d_theta, std::max<size_t>(nth, 1) * 4));
    CKH(cudaMalloc(&d_out, std::max<size_t>(nlp, 1) * 4));
    CKH(cudaMemcpyAsync(d_theta, theta, nth * 4, cudaMemcpyHostToDevice, st));
    std::vector<SlotState> hst(n_slots);
    for (int s = 0; s < n_slots; ++s) {
        hst[s] = SlotState{};
        hst[s].solve = s; hst[s].local = s; hst[s].active = 1; hst[s].temp_step = temp_step[s];
        hst[s].last_loss = -INFINITY; hst[s].loss = -INFINITY; hst[s].pending = -1;
        if (temp_step[s] < 0 || temp_step[s] >= TEMP_STEPS) { cleanup(); return ctx->fail(SOUP_ERR_INVALID, "soup_estep: temp_step must be in [0, 9)"); }
    }
    CKH(cudaMemcpyAsync(w.st, hst.data(), sizeof(SlotState) * n_slots, cudaMemcpyHostToDevice, st));
    CKH(cudaMemsetAsync(w.counts, 0, (size_t)n_slots * k * 4, st));
    if (nth) tables_from_theta_kernel<<<(unsigned)((nth + 255) / 256), 256, 0, st>>>(d_theta, w.TH, w.LA, w.LB, w.ctl, (uint32_t)ctx->L, w.Cpad, k, n_slots);
    const float log_prior = logf(1.0f / (float)k);
    launch_e(ctx, log_prior);
    launch_post(ctx, method, true);
    std::vector<int> h_counts((size_t)n_slots * k);
    CKH(cudaMemcpyAsync(h_counts.data(), w.counts, h_counts.size() * 4, cudaMemcpyDeviceToHost, st));
    launch_loss_control(ctx, st, false, 0.01f * (float)ctx->N);
    CKH(cudaGetLastError());
    if (log_probs && nlp) {
        plane_to_host_kernel<<<(unsigned)((nlp + 255) / 256), 256, 0, st>>>(w.ELL, d_out, (uint32_t)ctx->N, w.Cpad, k, n_slots, 1);
        CKH(cudaMemcpyAsync(log_probs, d_out, nlp * 4, cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
    }
    if (weights && nlp) {
        plane_to_host_kernel<<<(unsigned)((nlp + 255) / 256), 256, 0, st>>>(w.W, d_out, (uint32_t)ctx->N, w.Cpad, k, n_slots, 1);
        CKH(cudaMemcpyAsync(weights, d_out, nlp * 4, cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
    }
    if (lse && ctx->N) {
        std::vector<float> h((size_t)ctx->N * n_slots);
        CKH(cudaMemcpyAsync(h.data(), w.LSE, h.size() * 4, cudaMemcpyDeviceToHost, st));
        CKH(cudaStreamSynchronize(st));
        memcpy(lse, h.data(), h.size() * 4);
    }
    CKH(cudaMemcpyAsync(hst.data(), w.st, sizeof(SlotState) * n_slots, cudaMemcpyDeviceToHost, st));
    CKH(cudaStreamSynchronize(st));
    CKH(cudaGetLastError());
    if (loss) for (int s = 0; s < n_slots; ++s) loss[s] = hst[s].loss;
    if (cells_per_cluster) for (size_t i = 0; i < h_counts.size(); ++i) cells_per_cluster[i] = h_counts[i];
    cleanup();
    ctx->stats = soup_stats{};
    ctx->stats.kernel_launches = 4 + (log_probs ? 1 : 0) + (weights ? 1 : 0);
    ctx->stats.estep_launches = 1;
    ctx->stats.slots = n_slots; ctx->stats.lanes = w.ep.VM; ctx->stats.groups = w.ep.n_groups; ctx->stats.padded_cols = (int32_t)w.Cpad;
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_estep")

extern "C" int32_t soup_mstep(soup_ctx* ctx, int32_t k, int32_t n_slots, int32_t lanes, const float* weights, const float* theta_in,
                              const int32_t* locked, int32_t n_locked, float* theta_out) try {
    if (!ctx) return SOUP_ERR_INVALID;
    if (!weights || !theta_in || !theta_out || n_locked < 0 || (n_locked && !locked)) return ctx->fail(SOUP_ERR_INVALID, "soup_mstep: null argument");
    CK(cudaSetDevice(ctx->device));
    int32_t rc = hook_setup(ctx, k, n_slots, lanes);
    if (rc != SOUP_OK) return rc;
    Workspace& w = ctx->ws;
    cudaStream_t st = ctx->stream;
    const size_t nth = (size_t)n_slots * k * ctx->L, nw = (size_t)n_slots * ctx->N * k;
    float *d_theta = nullptr, *d_w = nullptr;
    auto cleanup = [&] { cudaFree(d_theta); cudaFree(d_w); };
    CKH(cudaMalloc(&d_theta, std::max<size_t>(nth, 1) * 4));
    CKH(cudaMalloc(&d_w, std::max<size_t>(nw, 1) * 4));
    CKH(cudaMemcpyAsync(d_theta, theta_in, nth * 4, cudaMemcpyHostToDevice, st));
    CKH(cudaMemcpyAsync(d_w, weights, nw * 4, cudaMemcpyHostToDevice, st));
    std::vector<uint8_t> cw(w.Cpad, 0);
    for (int s = 0; s < n_slots; ++s)
        for (int c = 0; c < k; ++c) cw[(size_t)s * k + c] = 1;
    for (int i = 0; i < n_locked; ++i) {
        if (locked[i] < 0 || locked[i] >= k) { cleanup(); return ctx->fail(SOUP_ERR_INVALID, "soup_mstep: locked cluster out of range"); }
        for (int s = 0; s < n_slots; ++s) cw[(size_t)s * k + locked[i]] = 0;
    }
    CKH(cudaMemcpyAsync(w.col_write, cw.data(), cw.size(), cudaMemcpyHostToDevice, st));
    if (nth) tables_from_theta_kernel<<<(unsigned)((nth + 255) / 256), 256, 0, st>>>(d_theta, w.TH, w.LA, w.LB, w.ctl, (uint32_t)ctx->L, w.Cpad, k, n_slots);
    if (nw) host_to_plane_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(d_w, w.W, (uint32_t)ctx->N, w.Cpad, k, n_slots);
    // arbitrary caller weights may be non-finite: use the literal (no-shortcut) path
    Ctl hc{};
    hc.copy_slot = -1; hc.nonfinite = 1; hc.best_loss = -INFINITY; hc.live_cols = (uint32_t)(n_slots * k);
    CKH(cudaMemcpyAsync(w.ctl, &hc, sizeof(Ctl), cudaMemcpyHostToDevice, st));
    launch_m(ctx);
    CKH(cudaGetLastError());
    if (nth) {
        plane_to_host_kernel<<<(unsigned)((nth + 255) / 256), 256, 0, st>>>(w.TH, d_theta, (uint32_t)ctx->L, w.Cpad, k, n_slots, 0);
        CKH(cudaMemcpyAsync(theta_out, d_theta, nth * 4, cudaMemcpyDeviceToHost, st));
    }
    CKH(cudaStreamSynchronize(st));
    CKH(cudaGetLastError());
    cleanup();
#undef CKH
    ctx->stats = soup_stats{};
    ctx->stats.kernel_launches = 4;
    ctx->stats.mstep_launches = 1;
    ctx->stats.slots = n_slots; ctx->stats.lanes = w.ep.VM; ctx->stats.groups = w.ep.n_groups; ctx->stats.padded_cols = (int32_t)w.Cpad;
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_mstep")

extern "C" int32_t soup_device_math(soup_ctx* ctx, int32_t op, const float* in, float* out, uint64_t n) try {
    if (!ctx) return SOUP_ERR_INVALID;
    if (!in || !out || (op != 0 && op != 1)) return ctx->fail(SOUP_ERR_INVALID, "soup_device_math: bad argument");
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return SOUP_OK;
    float *d_in = nullptr, *d_out = nullptr;
    CK(cudaMalloc(&d_in, n * 4));
    if (cudaMalloc(&d_out, n * 4) != cudaSuccess) { cudaFree(d_in); return ctx->fail(SOUP_ERR_NOMEM, "soup_device_math: out of device memory"); }
    cudaMemcpyAsync(d_in, in, n * 4, cudaMemcpyHostToDevice, ctx->stream);
    math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(op, d_in, d_out, n);
    cudaMemcpyAsync(out, d_out, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return ctx->fail(SOUP_ERR_CUDA, "soup_device_math: %s", cudaGetErrorString(e));
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_device_math")

// ------------------------------------------------------------------ troublet: call_doublets T:32-244 on the GPU
#include "troublet_kernels.cuh"

extern "C" int32_t soup_troublet(soup_ctx* ctx, uint64_t n_cells, uint64_t n_loci, uint64_t nnz, const uint64_t* row_ptr, const uint32_t* locus,
                                 const uint32_t* alt, const uint32_t* ref, const double* losses, int32_t k, int32_t n_clusters,
                                 const soup_troublet_params* p, soup_troublet_result* r, const char* barcode_blob, void* log_file) try {
    using namespace troublet_dev;
    if (!ctx) return SOUP_ERR_INVALID;
    if (!row_ptr || !losses || !p || !r || k <= 0 || (nnz && (!locus || !alt || !ref))) return ctx->fail(SOUP_ERR_INVALID, "soup_troublet: bad argument");
    if (!r->status || !r->best_singlet || !r->doublet1 || !r->doublet2 || !r->singlet_posterior || !r->doublet_posterior || !r->log_prob_singleton ||
        !r->log_prob_doublet || !r->cluster_logprobs)
        return ctx->fail(SOUP_ERR_INVALID, "soup_troublet: every result array must be allocated by the caller");
    // the reference keys its count tables by clusters 0..=max(column 2) and indexes them with the loss columns (T:33, T:61): a loss column
    // beyond them is `unwrap()` on None
    if (k > n_clusters) return ctx->fail(SOUP_ERR_INVALID, "called `Option::unwrap()` on a `None` value (more log-probability columns than clusters)");
    if (n_cells >= (1ull << 31) || n_loci >= (1ull << 31)) return ctx->fail(SOUP_ERR_UNSUPPORTED, "soup_troublet: more than 2^31 cells / loci");
    if (row_ptr[0] != 0 || row_ptr[n_cells] != nnz) return ctx->fail(SOUP_ERR_INVALID, "soup_troublet: row_ptr does not span [0, nnz]");
    for (uint64_t c = 0; c < n_cells; ++c)
        if (row_ptr[c + 1] < row_ptr[c]) return ctx->fail(SOUP_ERR_INVALID, "soup_troublet: row_ptr not monotone at cell %llu", (unsigned long long)c);
    for (uint64_t j = 0; j < nnz; ++j)
        if (locus[j] >= n_loci) return ctx->fail(SOUP_ERR_INVALID, "soup_troublet: locus index out of range at entry %llu", (unsigned long long)j);
    CK(cudaSetDevice(ctx->device));
    FILE* log = (FILE*)log_file;
    cudaStream_t st = ctx->stream;
    const int K = k;
    const size_t N = n_cells, L = std::max<uint64_t>(n_loci, 1);
    std::vector<const char*> names(N, "");
    if (barcode_blob) { const char* q = barcode_blob; for (size_t c = 0; c < N; ++c) { names[c] = q; q += strlen(q) + 1; } }

    // load_clusters' best index = the singlet_assignment of T:154-160: first maximum of the loss columns
    std::vector<int32_t> cluster(N, 0);
    for (size_t c = 0; c < N; ++c) {
        double best = -10000000000.0;   // T:427
        int bi = 0;
        for (int j = 0; j < K; ++j) if (losses[c * K + j] > best) { best = losses[c * K + j]; bi = j; }
        cluster[c] = bi;
    }

    struct Bufs {
        uint64_t* row_ptr = nullptr; uint32_t *locus = nullptr, *ref = nullptr, *alt = nullptr, *lines = nullptr, *removed = nullptr;
        int32_t *cluster = nullptr, *pairs = nullptr;
        unsigned long long *S0 = nullptr, *A = nullptr;
        double *soup = nullptr, *sing = nullptr, *dbl = nullptr;
        ~Bufs() { for (void* q : {(void*)row_ptr, (void*)locus, (void*)ref, (void*)alt, (void*)lines, (void*)removed, (void*)cluster, (void*)pairs, (void*)S0, (void*)A, (void*)soup, (void*)sing, (void*)dbl}) cudaFree(q); }
    } b;
#define CKT(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) { cudaGetLastError(); return ctx->fail(e_ == cudaErrorMemoryAllocation ? SOUP_ERR_NOMEM : SOUP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } \
    } while (0)
    const size_t nz = std::max<uint64_t>(nnz, 1), tab = L * (size_t)K * 2;
    CKT(cudaMalloc(&b.row_ptr, (N + 1) * 8)); CKT(cudaMalloc(&b.locus, nz * 4)); CKT(cudaMalloc(&b.ref, nz * 4)); CKT(cudaMalloc(&b.alt, nz * 4));
    CKT(cudaMalloc(&b.lines, L * 4)); CKT(cudaMalloc(&b.removed, std::max<size_t>(N, 1) * 4)); CKT(cudaMalloc(&b.cluster, std::max<size_t>(N, 1) * 4));
    CKT(cudaMalloc(&b.pairs, std::max<size_t>(N, 1) * PAIRS * 2 * 4)); CKT(cudaMalloc(&b.S0, tab * 8)); CKT(cudaMalloc(&b.A, tab * 8));
    CKT(cudaMalloc(&b.soup, L * 8)); CKT(cudaMalloc(&b.sing, std::max<size_t>(N, 1) * K * 8)); CKT(cudaMalloc(&b.dbl, std::max<size_t>(N, 1) * PAIRS * 8));
    CKT(cudaMemcpyAsync(b.row_ptr, row_ptr, (N + 1) * 8, cudaMemcpyHostToDevice, st));
    if (nnz) {
        CKT(cudaMemcpyAsync(b.locus, locus, nnz * 4, cudaMemcpyHostToDevice, st));
        CKT(cudaMemcpyAsync(b.ref, ref, nnz * 4, cudaMemcpyHostToDevice, st));
        CKT(cudaMemcpyAsync(b.alt, alt, nnz * 4, cudaMemcpyHostToDevice, st));
    }
    if (N) CKT(cudaMemcpyAsync(b.cluster, cluster.data(), N * 4, cudaMemcpyHostToDevice, st));
    CKT(cudaMemsetAsync(b.lines, 0, L * 4, st));
    CKT(cudaMemsetAsync(b.S0, 0, tab * 8, st));
    {   // statrs tables with the HOST libm (FCACHE[n].ln())
        double lf[171], f = 1.0;
        lf[0] = std::log(1.0);
        for (int i = 1; i <= 170; ++i) { f *= (double)i; lf[i] = std::log(f); }
        static const double dk[11] = {2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469, 4.51227709466894823700,
                                      -2.98285225323576655721, 1.05639711577126713077, -1.95428773191645869583e-1, 1.70970543404441224307e-2,
                                      -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9};
        CKT(cudaMemcpyToSymbolAsync(c_ln_fact, lf, sizeof(lf), 0, cudaMemcpyHostToDevice, st));
        CKT(cudaMemcpyToSymbolAsync(c_gamma_dk, dk, sizeof(dk), 0, cudaMemcpyHostToDevice, st));
    }
    if (N) tally_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(b.row_ptr, b.locus, b.ref, b.alt, b.cluster, b.S0, b.lines, (uint32_t)N, K);
    soup_kernel<<<(unsigned)((L + 255) / 256), 256, 0, st>>>(b.S0, b.soup, (uint32_t)L, K);
    CKT(cudaMemcpyAsync(b.A, b.S0, tab * 8, cudaMemcpyDeviceToDevice, st));
    CKT(cudaGetLastError());

    Args a{b.row_ptr, b.locus, b.ref, b.alt, b.S0, b.A, b.soup, b.lines, (uint32_t)N, K, 0,
           Consts{p->p_soup, std::log(1.0 - p->doublet_prior), std::log(p->doublet_prior), 20u}};
    std::vector<double> sing(N * (size_t)K), dbl(N * (size_t)PAIRS);
    std::vector<int32_t> pairs(N * (size_t)PAIRS * 2);
    std::vector<uint8_t> all_removed(N, 0);
    std::vector<uint32_t> new_removed;
    auto lse = [](const double* x, int n) {   // log_sum_exp T:246-250 (host libm, as the reference)
        double m = -INFINITY;
        for (int i = 0; i < n; ++i) m = (x[i] > m || std::isnan(m)) ? x[i] : m;
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += std::exp(x[i] - m);
        return m + std::log(s);
    };
    r->rounds = 0; r->removed = 0; r->device_ms = 0;
    size_t any_removed = 1;
    while (any_removed > 0) {   // T:46
        a.round = r->rounds;
        r->rounds += 1;
        CKT(cudaEventRecord(ctx->ev0, st));
        if (N) singlet_kernel<<<(unsigned)((N * K + 3) / 4), 128, 0, st>>>(a, b.sing);   // one warp per (cell, cluster)
        CKT(cudaMemcpyAsync(sing.data(), b.sing, sing.size() * 8, cudaMemcpyDeviceToHost, st));
        CKT(cudaStreamSynchronize(st));
        CKT(cudaGetLastError());
        // T:88-92: stable descending sort of the singlets, the pairs among the top four
        std::vector<std::pair<double, int>> order(K);
        for (size_t c = 0; c < N; ++c) {
            for (int j = 0; j < K; ++j) {
                if (std::isnan(sing[c * K + j])) return ctx->fail(SOUP_ERR_INVALID, "called `Option::unwrap()` on a `None` value (NaN log-probability, T:90)");
                order[j] = {sing[c * K + j], j};
            }
            std::stable_sort(order.begin(), order.end(), [](const auto& x, const auto& y) { return -x.first < -y.first; });
            const int top = std::min(K, 4);
            int q = 0;
            for (int i = 0; i < top; ++i)
                for (int j = i + 1; j < top; ++j, ++q) { pairs[(c * PAIRS + q) * 2] = order[i].second; pairs[(c * PAIRS + q) * 2 + 1] = order[j].second; }
            for (; q < PAIRS; ++q) pairs[(c * PAIRS + q) * 2] = pairs[(c * PAIRS + q) * 2 + 1] = -1;
        }
        if (N) {
            CKT(cudaMemcpyAsync(b.pairs, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice, st));
            doublet_kernel<<<(unsigned)((N * PAIRS + 3) / 4), 128, 0, st>>>(a, b.pairs, b.dbl);   // one warp per (cell, pair)
            CKT(cudaMemcpyAsync(dbl.data(), b.dbl, dbl.size() * 8, cudaMemcpyDeviceToHost, st));
        }
        CKT(cudaEventRecord(ctx->ev1, st));
        CKT(cudaStreamSynchronize(st));
        CKT(cudaGetLastError());
        {
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
            r->device_ms += ms;
        }
        new_removed.clear();
        for (size_t c = 0; c < N; ++c) {
            int best_singlet = 0, d1 = 0, d2 = 0;
            double best_s = -INFINITY, best_d = -INFINITY;
            for (int j = 0; j < K; ++j) if (sing[c * K + j] > best_s) { best_s = sing[c * K + j]; best_singlet = j; }   // T:83-86
            for (int q = 0; q < PAIRS; ++q)
                if (pairs[(c * PAIRS + q) * 2] >= 0 && dbl[c * PAIRS + q] > best_d) { best_d = dbl[c * PAIRS + q]; d1 = pairs[(c * PAIRS + q) * 2]; d2 = pairs[(c * PAIRS + q) * 2 + 1]; }   // T:127-131
            const double dq[2] = {best_s, best_d};
            const double dq_denom = lse(dq, 2);                                   // T:145-151
            double top_singlet = -INFINITY;
            int singlet_assignment = 0;
            for (int j = 0; j < K; ++j) if (losses[c * K + j] > top_singlet) { top_singlet = losses[c * K + j]; singlet_assignment = j; }   // T:154-160
            const double sq_denom = lse(losses + c * K, K);
            const double sp = std::exp(top_singlet - sq_denom), dp = std::exp(best_d - dq_denom);   // T:166-167
            int status;
            if (best_s > best_d) status = sp > p->singlet_threshold ? SOUP_TROUBLET_SINGLET : SOUP_TROUBLET_UNASSIGNED;   // T:169-184
            else if (dp >= p->doublet_threshold) {                                                                          // T:186-195
                if (!all_removed[c]) {
                    all_removed[c] = 1;
                    new_removed.push_back((uint32_t)c);
                    if (log) fprintf(log, "removing %s as doublet\n", names[c]);
                }
                status = SOUP_TROUBLET_DOUBLET;
            } else status = SOUP_TROUBLET_UNASSIGNED_DOUBLET;
            r->status[c] = status; r->best_singlet[c] = best_singlet; r->doublet1[c] = d1; r->doublet2[c] = d2;
            r->singlet_posterior[c] = sp; r->doublet_posterior[c] = dp; r->log_prob_singleton[c] = best_s; r->log_prob_doublet[c] = best_d;
            if (singlet_assignment != best_singlet && !all_removed[c] && log) fprintf(log, "error\t%s\t%d\t%d\n", names[c], singlet_assignment, best_singlet);   // T:205-209
        }
        memcpy(r->cluster_logprobs, sing.data(), sing.size() * 8);
        if (log) fprintf(log, "%zu cells removed this round\n", new_removed.size());   // T:216
        any_removed = new_removed.size();
        r->removed += new_removed.size();
        if (any_removed) {   // T:218-236: the removed cells leave their cluster's counts; the fractions follow (computed on the fly from A)
            CKT(cudaMemcpyAsync(b.removed, new_removed.data(), new_removed.size() * 4, cudaMemcpyHostToDevice, st));
            remove_kernel<<<(unsigned)((new_removed.size() + 7) / 8), 256, 0, st>>>(b.row_ptr, b.locus, b.ref, b.alt, b.cluster, b.removed, (uint32_t)new_removed.size(), b.A, K);
            CKT(cudaStreamSynchronize(st));
            CKT(cudaGetLastError());
        }
    }
#undef CKT
    return SOUP_OK;
}
SOUP_CATCH_RETURN_CTX("soup_troublet")
