// soup_main: souporcell_main (/root/reference/souporcell/src/main.rs:60-131, "S:") on top of
// soup_solve.  Host-only logic: seed fan-out, the souporcell3 three-run loop, bad-cluster
// detection, best-of bookkeeping, restart sharding over contexts (GPUs) and processes.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "host_parallel.hpp"
#include "soup_common.hpp"

using namespace soup;

namespace {

struct Candidate {   // what ranks all-gather once per run
    float loss;
    int32_t solve;
    int32_t has_best;
    int32_t pad;
    uint64_t nnz_updates;
};

// reference best-of order inside a run: higher loss wins, ties keep the lowest solve index (S:110, S:120)
bool better(float loss, int solve, bool has, float cur_loss, int cur_solve, bool cur_has) {
    if (!has) return false;
    if (!cur_has) return true;
    return loss > cur_loss || (loss == cur_loss && solve < cur_solve);
}

void log_solves(FILE* log, const soup_main_params* p, int spt, const std::vector<float>& loss, const std::vector<uint8_t>& mine) {
    // "thread {} iteration {} done with {}, best so far {}" (S:115-116) for the solves this process ran;
    // "best so far" is the thread's running best over those solves.
    if (!log) return;
    for (uint32_t t = 0; t < p->threads; ++t) {
        float best = -INFINITY;
        for (int i = 0; i < spt; ++i) {
            int s = (int)t * spt + i;
            if (!mine[s]) continue;
            float l = loss[s];
            if (l > best) best = l;
            fprintf(log, "thread %u iteration %d done with %s, best so far %s\n", t, i, format_f32(l).c_str(), format_f32(best).c_str());
        }
    }
}

}  // namespace

// Cross-process winner selection of one run: all-gather (loss, solve, has_best), pick the reference's
// winner (higher loss; ties -> lowest solve index, S:110/S:120), fetch its buffers from the owning rank.
extern "C" int32_t soup_exchange_best(int32_t rank, int32_t world, soup_allgather_fn allgather, soup_bcast_fn bcast, void* user,
                                      float* loss, int32_t* solve, int32_t* has_best, float* log_probs, uint64_t n_log_probs,
                                      float* theta, uint64_t n_theta, int32_t* winner_rank) try {
    if (world <= 1) { if (winner_rank) *winner_rank = 0; return SOUP_OK; }
    if (!allgather || !bcast || !loss || !solve || !has_best || rank < 0 || rank >= world) { set_thread_error("soup_exchange_best: bad argument"); return SOUP_ERR_INVALID; }
    Candidate mine{*loss, *solve, *has_best, 0, 0};
    std::vector<Candidate> all((size_t)world);
    int32_t e = allgather(user, &mine, all.data(), sizeof(Candidate));
    if (e != SOUP_OK) { set_thread_error("soup_exchange_best: allgather callback failed (%d)", e); return e; }
    Candidate win{-INFINITY, -1, 0, 0, 0};
    int wr = 0;
    for (int q = 0; q < world; ++q)
        if (better(all[q].loss, all[q].solve, all[q].has_best != 0, win.loss, win.solve, win.has_best != 0)) { win = all[q]; wr = q; }
    if (winner_rank) *winner_rank = wr;
    *loss = win.loss; *solve = win.solve; *has_best = win.has_best;
    if (!win.has_best) return SOUP_OK;
    if (log_probs && n_log_probs) e = bcast(user, log_probs, n_log_probs * sizeof(float), wr);
    if (e == SOUP_OK && theta && n_theta) e = bcast(user, theta, n_theta * sizeof(float), wr);
    if (e != SOUP_OK) { set_thread_error("soup_exchange_best: bcast callback failed (%d)", e); return e; }
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_exchange_best")

extern "C" int32_t soup_main(soup_ctx* const* ctxs, int32_t n_ctx, const soup_main_params* p, soup_main_result* r, void* log_file) try {
    if (!ctxs || n_ctx <= 0 || !ctxs[0]) { set_thread_error("soup_main: need at least one context"); return SOUP_ERR_INVALID; }
    soup_ctx* ctx0 = ctxs[0];
    if (!p || !r || !r->best_log_probs) { set_thread_error("soup_main: null argument"); return SOUP_ERR_INVALID; }
    if (p->threads == 0 || p->k <= 0) { set_thread_error("soup_main: threads and k must be >= 1"); return SOUP_ERR_INVALID; }
    const int pworld = p->proc_world > 1 ? p->proc_world : 1, prank = pworld > 1 ? p->proc_rank : 0;
    const bool cell_mode = p->cell_allreduce != nullptr || p->cell_comms != nullptr;
    // Cell sharding over NCCL inside the library.  The contexts are G cell groups (soup_main_params.cell_groups, default 1) of SL
    // contexts each, group-major: context g*SL + s holds the s-th of this process's cell ranges for group g.  Every group spans the
    // same S ranges (the size of its communicator) and runs its own share of the restarts (solve mod groups) — two groups on the SAME
    // GPUs overlap one group's all-reduce with the other's E / M passes.  SL == S: every range is here (one process, the result covers
    // all cells); SL < S: one process per range.
    const int G = p->cell_comms && p->cell_groups > 1 ? p->cell_groups : 1;
    const int SL = n_ctx / G;
    int32_t cg_rank = 0, cg_world = 1;
    if (p->cell_comms) {
        if (n_ctx % G != 0) { set_thread_error("soup_main: %d contexts do not split into %d cell groups", n_ctx, G); return SOUP_ERR_INVALID; }
        for (int i = 0; i < n_ctx; ++i)
            if (!p->cell_comms[i] || !ctxs[i]) { set_thread_error("soup_main: cell_comms needs one communicator per context"); return SOUP_ERR_INVALID; }
        soup_comm_rank(p->cell_comms[0], &cg_rank, &cg_world);
        if (SL != 1 && SL != cg_world) { set_thread_error("soup_main: a cell group spans %d ranks but %d of its contexts were given (1 or all of them)", cg_world, SL); return SOUP_ERR_INVALID; }
    }
    const bool cell_local = p->cell_comms != nullptr && SL == cg_world;   // every cell range of a group is in this process
    if (cell_mode && !p->cell_comms && n_ctx != 1) { set_thread_error("soup_main: cell sharding through callbacks takes ONE context per process"); return SOUP_ERR_INVALID; }
    if (cell_mode && !p->cell_comms && (!p->cell_allgather || p->cell_group_size <= 0)) { set_thread_error("soup_main: cell sharding needs cell_allgather and cell_group_size"); return SOUP_ERR_INVALID; }
    if (pworld > 1 && !p->proc_comm && (!p->allgather || !p->bcast)) { set_thread_error("soup_main: proc_world > 1 needs proc_comm or the allgather and bcast callbacks"); return SOUP_ERR_INVALID; }
    FILE* log = (FILE*)log_file;
    const int K = p->k;
    uint64_t n_cells = 0, n_loci = 0, nnz = 0;
    int32_t rc = soup_get_dims(ctx0, &n_cells, &n_loci, &nnz);
    if (rc != SOUP_OK) { set_thread_error("%s", soup_last_error(ctx0)); return rc; }
    // cell group inside this process: context i holds the i-th contiguous range of the cells; results cover all of them in order
    std::vector<uint64_t> cell_off(SL + 1, 0);
    if (p->cell_comms) {
        for (int g = 0; g < G; ++g)
            for (int sI = 0; sI < SL; ++sI) {
                uint64_t nc = 0, nl = 0;
                rc = soup_get_dims(ctxs[g * SL + sI], &nc, &nl, nullptr);
                if (rc != SOUP_OK) { set_thread_error("%s", soup_last_error(ctxs[g * SL + sI])); return rc; }
                if (nl != n_loci) { set_thread_error("soup_main: the cell shards must hold the same loci"); return SOUP_ERR_INVALID; }
                if (g == 0) cell_off[sI + 1] = cell_off[sI] + nc;
                else if (cell_off[sI + 1] - cell_off[sI] != nc) { set_thread_error("soup_main: every cell group must hold the same cell ranges"); return SOUP_ERR_INVALID; }
            }
        n_cells = cell_off[SL];
    }
    const uint64_t n_cells_global = cell_local ? n_cells : p->n_cells_global;
    const int spt = (int)std::ceil((float)p->restarts / (float)p->threads);   // S:63 (f32 arithmetic)
    const int n_solves = (int)p->threads * spt;
    const int world = p->cell_comms ? pworld * G : pworld * n_ctx;   // restart shards (a cell group runs the same solves on every member)

    uint8_t seed[32];
    memset(seed, p->seed, 32);                   // S:61
    StdRng master = StdRng::from_seed(seed);     // S:62

    r->has_best = 0; r->runs = 0; r->best_loss = -INFINITY; r->nnz_updates = 0; r->kernel_launches = 0; r->solve_ms = 0;
    r->slot_iterations = 0; r->timed_iterations = 0; r->estep_ms = 0; r->mstep_ms = 0;
    std::vector<float> best_theta((size_t)K * n_loci);
    bool have_best = false;
    float best_loss = -INFINITY;
    std::vector<int32_t> bad(K);

    for (int run = 0; run < 3; ++run) {          // S:70
        int n_bad = 0;
        if (cell_mode && !cell_local && K >= 16 && run > 0) {
            // the cluster sizes S:156-162 over ALL cells: add the cell group's counts up (host all-gather, K u64 per process)
            std::vector<uint64_t> mine((size_t)K, 0);
            count_cluster_cells(r->best_log_probs, have_best ? n_cells : 0, K, mine.data());
            int32_t group = p->cell_comms ? cg_world : (p->cell_group_size > 0 ? p->cell_group_size : 1);
            std::vector<uint64_t> all((size_t)K * group, 0);
            int32_t e = p->cell_comms ? soup_comm_allgather(p->cell_comms[0], mine.data(), all.data(), (uint64_t)K * sizeof(uint64_t))
                                      : p->cell_allgather(p->cell_comm_user, mine.data(), all.data(), (uint64_t)K * sizeof(uint64_t));
            if (e != SOUP_OK) { if (!p->cell_comms) set_thread_error("soup_main: the cell_allgather callback failed (%d)", e); return e; }
            for (int32_t g = 1; g < group; ++g)
                for (int c = 0; c < K; ++c) all[c] += all[(size_t)g * K + c];
            n_bad = bad_clusters_from_counts(run, K, all.data(), bad.data(), log);
        } else {
            // (K < 16 or run 0: an empty list, S:151-153 — also when no solve was finite, where the reference just ends the runs)
            n_bad = soup_bad_cluster_detection(run, K, r->best_log_probs, have_best ? n_cells : 0, bad.data(), log);
        }
        if (n_bad < 0) return n_bad;
        if (run > 0 && n_bad > 0 && !have_best && n_solves > 0) {
            // S:93-95 copies the EMPTY best_cluster_centers and indexes it with the flagged clusters (K >= 16 flags some even with no
            // cell assigned): the reference panics as soon as a solve of this run starts
            set_thread_error("souporcell3 run %d: no finite solve in the previous run (the reference panics here)", run + 1);
            return SOUP_ERR_INVALID;
        }
        if (run != 0 && n_bad == 0) {            // S:73-76
            if (log) fprintf(log, "No outliers detected on run %d!, Exiting.\n", run + 1);
            break;
        }
        std::vector<uint8_t> seeds((size_t)p->threads * 32);
        for (uint32_t t = 0; t < p->threads; ++t)            // S:77-80, new_seed S:855-861
            for (int i = 0; i < 32; ++i) seeds[(size_t)t * 32 + i] = master.gen_u8();

        // ---- one soup_solve per context, concurrently
        struct PerCtx {
            soup_solve_result res{};
            std::vector<float> lp, th, loss;
            std::vector<int32_t> iters;
            int32_t rc = SOUP_OK;
            soup_stats stats{};
        };
        std::vector<PerCtx> pc(n_ctx);
        auto work = [&](int i) {
            PerCtx& c = pc[i];
            c.lp.resize((size_t)(p->cell_comms ? cell_off[i % SL + 1] - cell_off[i % SL] : n_cells) * K);
            c.th.resize((size_t)K * n_loci);
            c.loss.assign(n_solves, NAN);
            c.iters.assign(n_solves, 0);
            soup_solve_params sp{};
            sp.k = K; sp.method = p->method; sp.n_streams = (int)p->threads; sp.solves_per_stream = spt;
            sp.stream_seeds = seeds.data();
            sp.init_theta = p->known_theta;
            sp.init_strategy = p->init_strategy;
            sp.cell_allreduce = p->cell_allreduce; sp.comm_user = p->cell_comm_user; sp.n_cells_global = n_cells_global;
            sp.cell_comm = p->cell_comms ? p->cell_comms[i] : nullptr;
            sp.cell_ranges = G > 1 ? 1 : 0;      // several groups overlap each other: splitting the M pass into ranges only costs launches then
            sp.base_theta = run > 0 ? best_theta.data() : nullptr;
            sp.reinit_clusters = run > 0 ? bad.data() : nullptr;
            sp.n_reinit = run > 0 ? n_bad : 0;
            sp.iteration_cap = p->iteration_cap; sp.max_slots = p->max_slots; sp.lanes = p->lanes;
            sp.record_counts = p->record_counts; sp.keep_trace = p->verbose_trace; sp.time_kernels = p->time_kernels;
            sp.shard_rank = p->cell_comms ? prank * G + i / SL : prank * n_ctx + i; sp.shard_world = world;
            c.res.best_log_probs = c.lp.data(); c.res.best_theta = c.th.data();
            c.res.solve_loss = c.loss.data(); c.res.solve_iters = c.iters.data();
            c.rc = soup_solve(ctxs[i], &sp, &c.res);
            if (c.rc == SOUP_OK) soup_get_stats(ctxs[i], &c.stats);
        };
        const auto tw0 = std::chrono::steady_clock::now();
        if (n_solves == 0) {
            // --restarts 0: S:81 runs no solve, every thread reports -inf, the TSV is empty (S:142 zips over no rows) and the exit is clean
            for (auto& c : pc) { c.res.has_best = 0; c.res.best_solve = -1; c.res.best_loss = -INFINITY; }
        } else if (n_ctx == 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (int i = 0; i < n_ctx; ++i) pool.emplace_back(work, i);
            for (auto& t : pool) t.join();
        }
        if (getenv("SOUP_DEBUG_SLOW")) {
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
            if (ms > 50) fprintf(stderr, "soup_slow soup_main: solve call took %.1f ms\n", ms);
        }
        double run_ms = 0;
        std::vector<float> loss(n_solves, NAN);
        std::vector<int32_t> iters(n_solves, 0);
        std::vector<uint8_t> mine(n_solves, 0);
        int win_ctx = -1;
        for (int i = 0; i < n_ctx; ++i) {
            if (pc[i].rc != SOUP_OK) { set_thread_error("%s", soup_last_error(ctxs[i])); return pc[i].rc; }
            run_ms = std::max(run_ms, pc[i].stats.solve_ms);
            r->nnz_updates += pc[i].res.nnz_updates;
            r->kernel_launches += pc[i].stats.kernel_launches;
            if (!p->cell_comms || i % SL == 0) r->slot_iterations += pc[i].stats.slot_iterations;
            if (i == 0) { r->estep_ms += pc[i].stats.estep_ms; r->mstep_ms += pc[i].stats.mstep_ms; r->timed_iterations += pc[i].stats.timed_iterations; }
            if (p->cell_comms) {                   // every member of a cell group ran the same solves with the same reduced numbers:
                if (i % SL != 0) continue;         // the group's first context speaks for it
                for (int s = prank * G + i / SL; s < n_solves; s += world) { loss[s] = pc[i].loss[s]; iters[s] = pc[i].iters[s]; mine[s] = 1; }
                if (better(pc[i].res.best_loss, pc[i].res.best_solve, pc[i].res.has_best != 0,
                           win_ctx >= 0 ? pc[win_ctx].res.best_loss : -INFINITY, win_ctx >= 0 ? pc[win_ctx].res.best_solve : -1, win_ctx >= 0))
                    win_ctx = i;
                continue;
            }
            for (int s = prank * n_ctx + i; s < n_solves; s += world) { loss[s] = pc[i].loss[s]; iters[s] = pc[i].iters[s]; mine[s] = 1; }
            if (better(pc[i].res.best_loss, pc[i].res.best_solve, pc[i].res.has_best != 0,
                       win_ctx >= 0 ? pc[win_ctx].res.best_loss : -INFINITY, win_ctx >= 0 ? pc[win_ctx].res.best_solve : -1, win_ctx >= 0))
                win_ctx = i;
        }
        r->solve_ms += run_ms;
        if (log) {
            if (p->verbose_trace) {   // per-iteration lines S:264 / S:354, grouped per solve
                std::vector<soup_iter_record> rec(p->iteration_cap > 0 ? p->iteration_cap : 1000);
                for (int s = 0; s < n_solves; ++s) {
                    int owner = s % world - prank * (p->cell_comms ? G : n_ctx);
                    if (owner < 0 || owner >= (p->cell_comms ? G : n_ctx)) continue;
                    if (p->cell_comms) owner *= SL;
                    int32_t n = 0;
                    if (soup_get_trace(ctxs[owner], s, rec.data(), (int32_t)rec.size(), &n) != SOUP_OK) continue;
                    for (int j = 0; j < n; ++j) {
                        std::string a = p->method == SOUP_METHOD_EM ? format_f32_w8(rec[j].loss) : format_f32(rec[j].loss);
                        std::string b = p->method == SOUP_METHOD_EM ? format_f32_w8(rec[j].change) : format_f32(rec[j].change);
                        fprintf(log, "binomial\t%d\t%d\t%d\t%d\t%s\t%s\t%d\n", s / spt, s % spt, j + 1, rec[j].temp_step, a.c_str(), b.c_str(),
                                rec[j].clusters_above_threshold);
                    }
                }
            }
            log_solves(log, p, spt, loss, mine);
        }

        // ---- winner of this run: local first, then across processes
        float w_loss = win_ctx >= 0 ? pc[win_ctx].res.best_loss : -INFINITY;
        int32_t w_solve = win_ctx >= 0 ? pc[win_ctx].res.best_solve : -1, w_has = win_ctx >= 0 ? 1 : 0;
        std::vector<float> run_lp, run_th;
        if (win_ctx >= 0 && p->cell_comms) {       // the winning group's rows, in cell order
            run_lp.resize((size_t)n_cells * K);
            for (int sI = 0; sI < SL; ++sI) memcpy(run_lp.data() + (size_t)cell_off[sI] * K, pc[win_ctx + sI].lp.data(), pc[win_ctx + sI].lp.size() * sizeof(float));
            run_th.swap(pc[win_ctx].th);
        }
        else if (win_ctx >= 0) { run_lp.swap(pc[win_ctx].lp); run_th.swap(pc[win_ctx].th); }
        else { run_lp.assign((size_t)n_cells * K, 0.0f); run_th.assign((size_t)K * n_loci, 0.0f); }
        if (pworld > 1) {
            // over NCCL inside the library (proc_comm) or through the caller's host callbacks
            soup_allgather_fn ag = p->allgather;
            soup_bcast_fn bc = p->bcast;
            void* user = p->comm_user;
            if (p->proc_comm) {
                ag = [](void* u, const void* send, void* recv, uint64_t n) { return soup_comm_allgather((soup_comm*)u, send, recv, n); };
                bc = [](void* u, void* buf, uint64_t n, int32_t root) { return soup_comm_bcast((soup_comm*)u, buf, n, root); };
                user = p->proc_comm;
            }
            int32_t e = soup_exchange_best(prank, pworld, ag, bc, user, &w_loss, &w_solve, &w_has,
                                           run_lp.data(), run_lp.size(), run_th.data(), run_th.size(), nullptr);
            if (e != SOUP_OK) return e;
        }
        struct { float loss; int32_t has_best; } win{w_loss, w_has};
        // adopt only if strictly better than the best carried over from earlier runs (S:120)
        const bool adopt = win.has_best && win.loss > best_loss;
        if (adopt) {
            memcpy(r->best_log_probs, run_lp.data(), run_lp.size() * sizeof(float));
            best_theta.swap(run_th);
            best_loss = win.loss;
            have_best = true;
        }
        r->runs = run + 1;
        if (!p->souporcell3) break;              // S:127-129
    }
    if (log) fprintf(log, "best total log probability = %s\n", format_f32(best_loss).c_str());   // S:131
    r->has_best = have_best ? 1 : 0;
    r->best_loss = best_loss;
    if (r->best_theta && have_best) memcpy(r->best_theta, best_theta.data(), best_theta.size() * sizeof(float));
    return SOUP_OK;
}
SOUP_CATCH_RETURN("soup_main")
