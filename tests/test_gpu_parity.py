"""GPU parity: libsoupgpu (through the C ABI) vs the CPU oracle on identical seeded inputs.
Bit-exact (np.array_equal on the raw f32 bits) unless stated otherwise."""
import os
import subprocess

import numpy as np
import pytest

from helpers import random_theta, synth_csr
from oracle import oracle_py as orc
import souporcell_b200 as sp
from souporcell_b200 import api

pytestmark = pytest.mark.gpu


def bits(a):
    """raw f32 bit patterns; NaNs are canonicalised (x86 yields 0xffc00000 for 0*inf, the GPU 0x7fffffff — both print "NaN")."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


LONG_MODES = {
    # name: environment while the context is created and the matrix loaded
    "default": {},
    # the M kernel's one-warp long-locus role on this small matrix (threshold lowered), cooperative kernel off
    "long_role": {"SOUP_LONG_THRESHOLD": "12", "SOUP_LONG_COOP": "0"},
    # cooperative long-locus kernel (one CTA per locus, rows staged in shared memory) for every locus with >= 12 entries
    "long_coop": {"SOUP_LONG_THRESHOLD": "12", "SOUP_TAIL_THRESHOLD": "12", "SOUP_LONG_COOP": "2"},
    # the shipped policy (cooperative kernel only while <= 128 columns are live), thresholds lowered so that it is exercised
    "tail_coop": {"SOUP_LONG_THRESHOLD": "30", "SOUP_TAIL_THRESHOLD": "8"},
    # cell-sharded mode only: the M pass / all-reduce / theta pass pipelined over three locus ranges, with long loci in front
    "ranges3": {"SOUP_CELL_RANGES": "3", "SOUP_LONG_THRESHOLD": "30", "SOUP_TAIL_THRESHOLD": "8"},
}


def make_ctx(csr, mode):
    import os
    env = LONG_MODES[mode]
    os.environ.update(env)
    try:
        return sp.Context(0).load_csr(*csr)
    finally:
        for key in env:
            os.environ.pop(key, None)


@pytest.fixture(scope="module", params=list(LONG_MODES))
def small(request):
    v, csr = synth_csr()
    ctx = make_ctx(csr, request.param)
    od = orc.OracleData.from_csr(*csr)
    yield v, csr, ctx, od
    ctx.close()


def test_device_math_matches_host_libm(small):
    _, _, ctx, _ = small
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.random(200000, dtype=np.float32), np.float32(1) - rng.random(1000, dtype=np.float32) * 1e-3,
                        np.array([1e-4, 0.9999, 0.01, 0.99, 0.5, 1.0, 1e-30, 3e38], dtype=np.float32)])
    ref = np.array([orc.logf(float(t)) for t in x[:20000]], dtype=np.float32)
    assert np.array_equal(bits(ctx.device_math("ln", x[:20000])), bits(ref))
    y = np.concatenate([-rng.random(20000, dtype=np.float32) * 120, rng.standard_normal(5000).astype(np.float32),
                        np.array([0, -0.0, -87.5, -103.9, -104.5, -200, 88.0, 89.0], dtype=np.float32)])
    ref = np.array([orc.expf(float(t)) for t in y], dtype=np.float32)
    assert np.array_equal(bits(ctx.device_math("exp", y)), bits(ref))


def test_load_builds_csc_totals_lbc(small):
    _, csr, ctx, od = small
    n, L, row_ptr, locus, alt, ref = csr
    ex = od.export()
    assert np.array_equal(bits(ctx.get_cell_totals()), bits(ex["total_alleles"]))
    assert np.array_equal(bits(ctx.get_csr_lbc()), bits(ex["lbc"]))
    col_ptr, cell, calt, cref = ctx.get_csc()
    # reference transpose: stable sort of the CSR by locus
    rows = np.repeat(np.arange(n, dtype=np.uint32), np.diff(row_ptr).astype(np.int64))
    order = np.argsort(locus, kind="stable")
    assert np.array_equal(cell, rows[order]) and np.array_equal(calt, alt[order]) and np.array_equal(cref, ref[order])
    assert np.array_equal(col_ptr, np.concatenate([[0], np.cumsum(np.bincount(locus, minlength=L))]).astype(np.uint64))


@pytest.mark.parametrize("method", [sp.METHOD_EM, sp.METHOD_KHM])
@pytest.mark.parametrize("lanes", [0, 1, 2, 4, 8])
@pytest.mark.parametrize("k", [1, 3, 8])
def test_estep_mstep_bit_exact(small, method, lanes, k):
    _, csr, ctx, od = small
    rng = np.random.default_rng(100 + k)
    S = 5
    theta = random_theta(rng, (S, k, ctx.n_loci))
    temp_steps = np.array([0, 3, 8, 5, 1], dtype=np.int32)
    got = ctx.estep(theta, temp_steps, method=method, lanes=lanes)
    for s in range(S):
        want = od.step(theta[s], method=method, temp_step=int(temp_steps[s]))
        assert np.array_equal(bits(got["log_probs"][s]), bits(want["log_probs"])), f"log_probs slot {s}"
        assert np.array_equal(bits(got["lse"][s]), bits(want["lse"]))
        assert np.array_equal(bits(got["weights"][s]), bits(want["weights"]))
        assert bits(got["loss"][s:s + 1])[0] == bits(np.array([want["loss"]]))[0]
        th = ctx.mstep(got["weights"][s:s + 1], theta[s:s + 1], lanes=lanes)
        assert np.array_equal(bits(th[0]), bits(want["theta_out"]))


def test_mstep_locked_clusters(small):
    _, csr, ctx, od = small
    rng = np.random.default_rng(5)
    k = 6
    theta = random_theta(rng, (2, k, ctx.n_loci))
    got = ctx.estep(theta, [2, 2])
    th = ctx.mstep(got["weights"], theta, locked=[1, 4])
    for s in range(2):
        want = od.step(theta[s], temp_step=2, locked=[1, 4])
        assert np.array_equal(bits(th[s]), bits(want["theta_out"]))
        assert np.array_equal(bits(th[s][1]), bits(theta[s][1]))


@pytest.mark.parametrize("method", [sp.METHOD_EM, sp.METHOD_KHM])
def test_solve_matches_oracle_per_iteration(small, method):
    _, csr, ctx, od = small
    rng = np.random.default_rng(11)
    k, S = 3, 7
    theta0 = random_theta(rng, (S, k, ctx.n_loci))
    res = ctx.solve(k, method=method, n_streams=S, solves_per_stream=1, theta0=theta0, keep_trace=True, record_counts=True,
                    max_slots=3)   # 3 resident slots -> slots are refilled while others keep iterating
    best = None
    for s in range(S):
        want = od.solve(theta0[s], method=method)
        tr = ctx.get_trace(s)
        assert len(tr) == want["iterations"] == res.solve_iters[s]
        for (t0, l0, c0, a0), (t1, l1, c1, a1) in zip(tr, want["trace"]):
            assert t0 == t1 and a0 == a1
            assert np.float32(l0).view(np.uint32) == np.float32(l1).view(np.uint32)
            assert np.float32(c0).view(np.uint32) == np.float32(c1).view(np.uint32)
        assert bits(res.solve_loss[s:s + 1])[0] == bits(np.array([want["loss"]]))[0]
        if best is None or want["loss"] > best[0]:
            best = (want["loss"], s, want)
    assert res.has_best and res.best_solve == best[1]
    assert np.array_equal(bits(res.best_log_probs), bits(best[2]["log_probs"]))
    assert np.array_equal(bits(res.best_theta), bits(best[2]["theta"]))


# ---------------------------------------------------------------------------------------------- drivers, edge cases
def test_main_matches_oracle_with_chacha_seeds(small):
    """souporcell_main S:60-131 with the reference's own seeding: master StdRng([seed;32]) -> per-thread seeds ->
    theta0 drawn on the GPU from the ChaCha12 stream; restarts / threads combinations incl. ceil(restarts/threads)."""
    _, csr, ctx, od = small
    for method in (sp.METHOD_EM, sp.METHOD_KHM):
        for restarts, threads, seed, slots in [(5, 2, 4, 0), (4, 4, 9, 3), (3, 1, 200, 2)]:
            got = sp.main(ctx, 4, method=method, restarts=restarts, threads=threads, seed=seed, max_slots=slots)
            want = od.main(4, method=method, restarts=restarts, threads=threads, seed=seed)
            assert got.has_best and got.runs == want["runs"] == 1
            assert bits([got.best_loss])[0] == bits([want["best_loss"]])[0]
            assert np.array_equal(bits(got.best_log_probs), bits(want["best_log_probs"]))
            assert np.array_equal(bits(got.best_theta), bits(want["best_theta"]))
            assert got.nnz_updates == want["nnz_updates"]


def test_theta0_drawn_on_device_equals_stdrng_stream(small):
    _, csr, ctx, od = small
    k, seeds = 3, sp.thread_seeds(4, 2, 0)
    res = ctx.solve(k, n_streams=2, solves_per_stream=2, stream_seeds=seeds, iteration_cap=1, keep_trace=True)
    for s in range(4):
        stream, it = divmod(s, 2)
        th0 = orc.init_uniform(bytes(seeds[stream]), it * k * ctx.n_loci, k, ctx.n_loci)   # word offset it*K*L (S:554-563)
        want = od.solve(th0, iteration_cap=1)
        assert bits(res.solve_loss[s:s + 1])[0] == bits([want["loss"]])[0]


@pytest.mark.parametrize("method", [sp.METHOD_EM, sp.METHOD_KHM])
def test_souporcell3_three_runs_match_oracle(method):
    """-s true with k >= 16: bad-cluster detection, re-draw of the flagged clusters, locked rest (S:70-76, 92-97, 150-187)."""
    v, csr = synth_csr(n_cells=700, n_loci=2500, k=16, mean_loci=90, seed=5)
    od = orc.OracleData.from_csr(*csr)
    with sp.Context(0) as ctx:
        ctx.load_csr(*csr)
        got = sp.main(ctx, 16, method=method, restarts=4, threads=2, seed=4, souporcell3=True)
    want = od.main(16, method=method, restarts=4, threads=2, seed=4, souporcell3=True)
    assert got.runs == want["runs"] and got.runs >= 2
    assert bits([got.best_loss])[0] == bits([want["best_loss"]])[0]
    assert np.array_equal(bits(got.best_log_probs), bits(want["best_log_probs"]))
    assert np.array_equal(bits(got.best_theta), bits(want["best_theta"]))
    assert np.array_equal(got.assignments, np.argmax(want["best_log_probs"], axis=1))


def test_restart_shards_reproduce_the_single_shard_winner(small):
    _, csr, ctx, od = small
    seeds = sp.thread_seeds(4, 6, 0)
    full = ctx.solve(3, n_streams=6, solves_per_stream=1, stream_seeds=seeds)
    parts = [ctx.solve(3, n_streams=6, solves_per_stream=1, stream_seeds=seeds, shard_rank=r, shard_world=3) for r in range(3)]
    for r, p in enumerate(parts):
        mine = np.arange(r, 6, 3)
        assert np.array_equal(bits(p.solve_loss[mine]), bits(full.solve_loss[mine]))
        assert np.isnan(np.delete(p.solve_loss, mine)).all()
    best = max(parts, key=lambda p: (p.best_loss, -p.best_solve))
    assert best.best_solve == full.best_solve and np.array_equal(bits(best.best_log_probs), bits(full.best_log_probs))


def test_solve_is_deterministic_run_to_run(small):
    """Same inputs -> identical bits, every time (no atomics on floats, fixed accumulation order, ordered best-of)."""
    _, csr, ctx, od = small
    seeds = sp.thread_seeds(4, 5, 0)
    ref = None
    for rep in range(12):
        r = ctx.solve(3, method=rep % 2, n_streams=5, solves_per_stream=2, stream_seeds=seeds, max_slots=[0, 3, 4][rep % 3])
        key = (rep % 2,)
        sig = (r.best_solve, bits(r.solve_loss).tobytes(), r.solve_iters.tobytes(), bits(r.best_log_probs).tobytes(), bits(r.best_theta).tobytes())
        if ref is None:
            ref = {}
        if key in ref:
            assert ref[key] == sig, f"run {rep} differs"
        ref[key] = sig


def test_golden_micro_fixture_on_gpu():
    import json, os
    G = os.path.join(os.path.dirname(__file__), "golden")
    lm = json.load(open(os.path.join(G, "loader_micro.json")))["expect"]
    g = json.load(open(os.path.join(G, "em_micro.json")))
    with sp.Context(0) as ctx:
        ctx.load_csr(4, 3, np.array(lm["row_ptr"]), np.array(lm["locus"]), np.array(lm["alt"]), np.array(lm["ref"]))
        assert ctx.get_cell_totals().tolist() == lm["total_alleles"]
        for c in g["cases"]:
            m = sp.METHOD_KHM if c["method"] == "khm" else sp.METHOD_EM
            th0 = np.array(c["theta0"], dtype=np.float32)
            st = ctx.estep(th0[None], [c["temp_step"]], method=m)
            np.testing.assert_allclose(st["log_probs"][0], np.array(c["log_probs"]), rtol=g["tolerance"])
            np.testing.assert_allclose(st["loss"][0], c["loss"], rtol=g["tolerance"])
            res = ctx.solve(th0.shape[0], method=m, theta0=th0[None], keep_trace=True)
            assert res.solve_iters[0] == c["solve"]["iterations"]
            np.testing.assert_allclose(res.best_theta, np.array(c["solve"]["theta"]), rtol=400 * g["tolerance"])


def test_edge_cases_empty_cells_big_counts_nonfinite_theta():
    rng = np.random.default_rng(3)
    n, L = 40, 30
    rows = []
    for c in range(n):
        if c % 7 == 0:
            rows.append(np.array([], dtype=np.int64))                       # cells with no used locus at all
        else:
            rows.append(np.sort(rng.choice(L, size=rng.integers(1, 12), replace=False)))
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.uint64)
    locus = np.concatenate(rows).astype(np.uint32)
    alt = rng.integers(0, 4, size=locus.shape[0]).astype(np.uint32)
    ref = rng.integers(0, 4, size=locus.shape[0]).astype(np.uint32)
    alt[::5] = rng.integers(60, 400, size=alt[::5].shape[0])               # beyond the packed count field and the ln-factorial table
    ref[3::11] = 0
    alt[(alt == 0) & (ref == 0)] = 1
    zero_at = 4
    alt[zero_at] = ref[zero_at] = 0                                         # an explicit all-zero entry is dropped (S:664)
    od = orc.OracleData.from_csr(n, L, row_ptr, locus, alt, ref)
    with sp.Context(0) as ctx:
        ctx.load_csr(n, L, row_ptr, locus, alt, ref)
        assert ctx.nnz == od.nnz == locus.shape[0] - 1
        ex = od.export()
        assert np.array_equal(bits(ctx.get_csr_lbc()), bits(ex["lbc"])) and np.array_equal(bits(ctx.get_cell_totals()), bits(ex["total_alleles"]))
        k = 5
        theta = random_theta(rng, (2, k, L))
        theta[1, 0, :3] = [0.0, 1.0, 0.5]                                   # ln(0) = -inf: 0 * -inf = NaN in the reference's literal expression
        for method in (sp.METHOD_EM, sp.METHOD_KHM):
            got = ctx.estep(theta, [0, 8], method=method)
            for s in range(2):
                want = od.step(theta[s], method=method, temp_step=[0, 8][s])
                assert np.array_equal(bits(got["log_probs"][s]), bits(want["log_probs"]))
                assert np.array_equal(bits(got["weights"][s]), bits(want["weights"]))
                th = ctx.mstep(got["weights"][s:s + 1], theta[s:s + 1])
                assert np.array_equal(bits(th[0]), bits(want["theta_out"]))
        res = ctx.solve(k, theta0=theta[:1], keep_trace=True)
        want = od.solve(theta[0])
        assert res.solve_iters[0] == want["iterations"] and np.array_equal(bits(res.best_log_probs), bits(want["log_probs"]))
    # malformed CSR is rejected, not silently clustered
    with sp.Context(0) as ctx:
        bad = locus.copy()
        bad[1], bad[2] = bad[2], bad[1]
        with pytest.raises(sp.SoupError):
            ctx.load_csr(n, L, row_ptr, bad, alt, ref)
        with pytest.raises(sp.SoupError):
            ctx.load_csr(n, 5, row_ptr, locus, alt, ref)
    with sp.Context(0) as ctx:
        with pytest.raises(sp.SoupError):
            ctx.solve(3, stream_seeds=sp.thread_seeds(4, 1, 0))             # nothing loaded


def test_cli_end_to_end_matches_oracle_cli(tmp_path):
    """The drop-in executable vs the oracle's CLI on the same files and flags: clusters_tmp.tsv byte-identical,
    per-iteration stderr trace identical as a multiset of lines."""
    import subprocess
    from souporcell_b200 import build, synth
    v = synth.generate(n_cells=250, n_loci=1500, k=3, mean_loci=90, seed=8)
    alt, ref, bc = synth.write_files(v, str(tmp_path))
    for extra in (["-k", "3", "--restarts", "4", "-t", "2"], ["-k", "4", "--restarts", "3", "-m", "khm", "--seed", "11", "--min_alt", "6", "--min_ref=6"],
                  # init_cluster_centers_random_assignment S:565-594 (gen_range restated from rand 0.9.3, unpinned)
                  ["-k", "3", "--restarts", "5", "-t", "2", "--initialization_strategy", "random_cell_assignment"],
                  ["-k", "16", "--restarts", "2", "-t", "1", "-m", "khm", "-s", "true", "--seed", "7", "--initialization_strategy=random_cell_assignment"]):
        args = ["-a", alt, "-r", ref, "-b", bc] + extra
        ours = subprocess.run([build.CLI] + args, capture_output=True, text=True)
        theirs = subprocess.run([orc.CLI_PATH] + args, capture_output=True, text=True)
        assert ours.returncode == 0 and theirs.returncode == 0, ours.stderr[-500:]
        assert ours.stdout == theirs.stdout and ours.stdout.count("\n") == v.n_cells
        pick = lambda s: sorted(l for l in s.splitlines() if l.startswith(("binomial", "total loci", "best total", "thread ")))
        assert pick(ours.stderr) == pick(theirs.stderr)


def test_cli_no_restarts_and_no_outlier_edge_cases(tmp_path):
    """--restarts 0 (S:63 -> 0 solves per thread): no solve runs, `best total log probability = -inf`, an empty TSV and a clean exit,
    with and without souporcell3 (k < 16: "No outliers detected on run 2!"; k >= 16: clusters are flagged on every run, but S:93 is
    never reached because no solve starts) — byte for byte what the oracle CLI prints."""
    from souporcell_b200 import build, synth
    v = synth.generate(n_cells=60, n_loci=500, k=3, mean_loci=40, seed=2)
    alt, ref, bc = synth.write_files(v, str(tmp_path))
    for extra in (["-k", "3", "--restarts", "0"], ["-k", "3", "--restarts", "0", "-s", "true"], ["-k", "16", "--restarts", "0", "-s", "true", "-m", "khm"]):
        args = ["-a", alt, "-r", ref, "-b", bc] + extra
        ours = subprocess.run([build.CLI] + args, capture_output=True, text=True)
        theirs = subprocess.run([orc.CLI_PATH] + args, capture_output=True, text=True)
        assert ours.returncode == 0 and theirs.returncode == 0, ours.stderr[-500:]
        assert ours.stdout == theirs.stdout == ""
        keep = lambda s: [l for l in s.splitlines() if not l.startswith(("oracle_timing", "soup_", "soupgpu\t"))]
        assert keep(ours.stderr) == keep(theirs.stderr)
        assert "best total log probability = -inf" in ours.stderr


def test_full_size_properties_config2():
    """BASELINE configs[1] at full size (10k cells x 50k loci, k=8): too big for the oracle to cluster, so parity is
    checked on one capped solve (3 iterations, ~10 s of oracle time) bit for bit, plus size-independent properties."""
    import bench
    cfg, csr = bench.workload(2)
    n, L = csr[0], csr[1]
    rng = np.random.default_rng(0)
    th0 = random_theta(rng, (1, cfg["k"], L))
    with sp.Context(0) as ctx:
        ctx.load_csr(*csr)
        col_ptr, cell, calt, cref = ctx.get_csc()
        assert int(col_ptr[-1]) == csr[3].shape[0] and np.all(np.diff(col_ptr.astype(np.int64)) >= 0)
        seg = np.repeat(np.arange(L), np.diff(col_ptr.astype(np.int64)))
        same = seg[1:] == seg[:-1]
        assert np.all(cell[1:][same] > cell[:-1][same])                     # cells strictly ascending inside every locus
        assert int(calt.sum()) == int(csr[4].sum()) and int(cref.sum()) == int(csr[5].sum())
        e = ctx.estep(th0, [4])
        w = e["weights"][0]
        np.testing.assert_allclose(w.sum(axis=1), 1.0, rtol=1e-4)     # f32 softmax over k = 8, same arithmetic as the reference
        th1 = ctx.mstep(e["weights"], th0)
        assert th1.min() >= np.float32(0.01) and th1.max() <= np.float32(0.99)
        res = ctx.solve(cfg["k"], theta0=th0, iteration_cap=3, keep_trace=True)
        od = orc.OracleData.from_csr(*csr)
        want = od.solve(th0[0], iteration_cap=3)
        assert [bits([t[1]])[0] for t in ctx.get_trace(0)] == [bits([t[1]])[0] for t in want["trace"]]
        assert np.array_equal(bits(res.best_log_probs), bits(want["log_probs"])) and np.array_equal(bits(res.best_theta), bits(want["theta"]))
        # all 50 restarts: every solve runs its 9 temperature steps, finishes within the cap, and the winner is the max
        full = sp.main(ctx, cfg["k"], restarts=50, threads=50)
        assert full.has_best and full.nnz_updates % ctx.nnz == 0 and 9 * 50 <= full.nnz_updates // ctx.nnz <= 1000 * 50
        assert np.isfinite(full.best_log_probs).all() and (np.bincount(full.assignments, minlength=cfg["k"]) > 0).sum() >= 4


# ---------------------------------------------------------------------------------------------- cell-sharded mode
def _split_cells(csr, parts):
    """contiguous cell ranges balanced by nnz; each shard keeps all loci."""
    return api.split_cells(csr, parts)


@pytest.mark.parametrize("mode", ["default", "long_coop", "ranges3"])
@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("method", [sp.METHOD_EM, sp.METHOD_KHM])
def test_cell_sharded_mode_agrees_with_oracle_to_tolerance(world, method, mode):
    """`world` ranks emulated on one GPU: one context per cell shard, one thread per rank, the all-reduce callback is a
    rendezvous that sums the ranks' device buffers.  The pseudocount joins the reduced sum, so this mode is checked at
    north_star's tolerance (1e-5 relative per-iteration loss and theta, up to the first differing control decision)
    and on hard assignments, not bit for bit."""
    import threading
    import torch
    from souporcell_b200 import api

    v, csr = synth_csr(n_cells=400, n_loci=2500, k=3, mean_loci=110, seed=17)
    od = orc.OracleData.from_csr(*csr)
    cuts, shards = _split_cells(csr, world)
    k, S = 3, 4
    theta0 = random_theta(np.random.default_rng(2), (S, k, csr[1]))
    bar = threading.Barrier(world)
    bufs, tmp = [None] * world, [None]

    class Dev:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

    def make_cb(rank):
        def cb(user, buf, n, stream):
            try:
                torch.cuda.synchronize()                       # this rank's partial statistics are complete
                bufs[rank] = torch.as_tensor(Dev(buf, n), device="cuda")
                bar.wait()
                if rank == 0:
                    acc = bufs[0].clone()
                    for t in bufs[1:]:
                        acc += t                               # fixed rank order -> deterministic
                    tmp[0] = acc
                    torch.cuda.synchronize()
                bar.wait()
                bufs[rank].copy_(tmp[0])
                torch.cuda.synchronize()
                bar.wait()
                return 0
            except Exception:
                bar.abort()
                return -1
        return api.ALLREDUCE_FN(cb)

    results = [None] * world

    def run(rank):
        with sp.Context(0) as ctx:
            ctx.load_csr(*shards[rank])
            res = ctx.solve(k, method=method, n_streams=S, solves_per_stream=1, theta0=theta0, keep_trace=True, max_slots=3,
                            cell_allreduce=make_cb(rank), n_cells_global=csr[0])
            results[rank] = (res, [ctx.get_trace(s) for s in range(S)])

    import os
    os.environ.update(LONG_MODES[mode])       # read when the contexts are created / the shards loaded (inside the threads)
    try:
        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in threads]
        [t.join() for t in threads]
    finally:
        for key in LONG_MODES[mode]:
            os.environ.pop(key, None)
    assert all(r is not None for r in results)
    res0, traces0 = results[0]
    want = [od.solve(theta0[s], method=method) for s in range(S)]
    for s in range(S):
        got_tr, want_tr = traces0[s], want[s]["trace"]
        n_same = min(len(got_tr), len(want_tr))
        np.testing.assert_allclose([t[1] for t in got_tr[:n_same]], [t[1] for t in want_tr[:n_same]], rtol=1e-5)
        assert abs(len(got_tr) - len(want_tr)) <= 1                      # at most one borderline convergence decision
        for r in range(1, world):                                         # every rank takes identical decisions
            assert results[r][1][s] == got_tr
    best = int(np.argmax([w["loss"] for w in want]))
    assert res0.best_solve == best
    np.testing.assert_allclose(res0.best_theta, want[best]["theta"], rtol=1e-5, atol=1e-7)
    lp = np.concatenate([results[r][0].best_log_probs for r in range(world)])   # each rank returns its own cells
    np.testing.assert_allclose(lp, want[best]["log_probs"], rtol=1e-5)
    assert np.array_equal(np.argmax(lp, axis=1), np.argmax(want[best]["log_probs"], axis=1))


def test_cluster_allele_counts_exact(small):
    """soup_cluster_allele_counts vs a numpy restatement of consensus.py:296-331 (cells with cluster -1, the doublets there, skipped)."""
    _, csr, ctx, _ = small
    n, L, row_ptr, locus, alt, ref = csr
    rng = np.random.default_rng(12)
    k = 5
    cl = rng.integers(-1, k, size=n).astype(np.int32)
    got = ctx.cluster_allele_counts(cl, k)
    want = np.zeros((L, k, 2), dtype=np.uint64)
    cell_of = np.repeat(np.arange(n), np.diff(row_ptr).astype(np.int64))
    keep = cl[cell_of] >= 0
    np.add.at(want, (locus[keep].astype(np.int64), cl[cell_of][keep].astype(np.int64), 0), ref[keep].astype(np.uint64))
    np.add.at(want, (locus[keep].astype(np.int64), cl[cell_of][keep].astype(np.int64), 1), alt[keep].astype(np.uint64))
    assert np.array_equal(got, want)
    assert got.sum() == int(ref[keep].sum()) + int(alt[keep].sum())


def test_main_over_two_contexts_equals_one(small):
    """soup_main's in-process restart sharding (one context per GPU, one thread each), exercised with two contexts on
    the same device: solves split s mod 2, winner merged with the reference's tie rule -> identical to one context."""
    _, csr, ctx, od = small
    with sp.Context(0) as ctx2:
        ctx2.load_csr(*csr)
        for method, s3 in ((sp.METHOD_EM, False), (sp.METHOD_KHM, False)):
            one = sp.main(ctx, 4, method=method, restarts=7, threads=3, seed=4)
            two = sp.main([ctx, ctx2], 4, method=method, restarts=7, threads=3, seed=4)
            assert one.has_best and two.has_best and bits([one.best_loss])[0] == bits([two.best_loss])[0]
            assert np.array_equal(bits(one.best_log_probs), bits(two.best_log_probs))
            assert np.array_equal(bits(one.best_theta), bits(two.best_theta))
            assert one.nnz_updates == two.nnz_updates


def test_cli_known_genotypes_matches_oracle_cli(tmp_path):
    """--known_genotypes / --known_genotypes_sample_names (souporcell_pipeline.py:532-535): every restart starts from
    the VCF-derived centers (S:487); TSV byte-identical to the oracle CLI, also under souporcell3's re-draw rule."""
    import subprocess
    from souporcell_b200 import build, synth
    from test_host import write_vcf
    v = synth.generate(n_cells=300, n_loci=1200, k=3, mean_loci=90, seed=12)
    alt, ref, bc = synth.write_files(v, str(tmp_path))
    vcf = str(tmp_path / "known.vcf")
    write_vcf(vcf, 1200, ["s0", "s1", "s2"], np.random.default_rng(6))
    args = ["-a", alt, "-r", ref, "-b", bc, "-k", "3", "--restarts", "2", "-t", "2", "-g", vcf, "--known_genotypes_sample_names", "s2", "s0", "s1"]
    ours = subprocess.run([build.CLI] + args, capture_output=True, text=True)
    theirs = subprocess.run([orc.CLI_PATH] + args, capture_output=True, text=True)
    assert ours.returncode == 0 and theirs.returncode == 0, ours.stderr[-400:] + theirs.stderr[-400:]
    assert ours.stdout == theirs.stdout and ours.stdout.count("\n") == v.n_cells


@pytest.mark.parametrize("mode", ["long_role", "long_coop", "tail_coop"])
@pytest.mark.parametrize("k,n_solves,slots", [(5, 30, 0), (7, 20, 13), (64, 3, 0), (33, 9, 4), (2, 70, 0)])
def test_many_columns_straddling_groups(k, n_solves, slots, mode):
    """Column counts that are not multiples of the 256 / 64 / 32-column group widths (slots straddle groups, narrow last
    groups, several wide groups) with refills and tail compaction: every solve bit-identical to the oracle."""
    v, csr = synth_csr(n_cells=220, n_loci=1400, k=4, mean_loci=80, seed=29)
    od = orc.OracleData.from_csr(*csr)
    theta0 = random_theta(np.random.default_rng(k), (n_solves, k, csr[1]))
    ctx = make_ctx(csr, mode)
    with ctx:
        res = ctx.solve(k, method=sp.METHOD_KHM if k % 2 else sp.METHOD_EM, n_streams=n_solves, solves_per_stream=1, theta0=theta0, max_slots=slots)
        check = range(n_solves) if n_solves <= 10 else list(range(0, n_solves, 7)) + [n_solves - 1]
        for s in check:
            want = od.solve(theta0[s], method=1 if k % 2 else 0)
            assert res.solve_iters[s] == want["iterations"] and bits(res.solve_loss[s:s + 1])[0] == bits([want["loss"]])[0], s
        b = res.best_solve
        want = od.solve(theta0[b], method=1 if k % 2 else 0)
        assert np.array_equal(bits(res.best_log_probs), bits(want["log_probs"])) and np.array_equal(bits(res.best_theta), bits(want["theta"]))
        assert res.best_loss == np.nanmax(res.solve_loss)


@pytest.mark.parametrize("groups,shards", [(1, 2), (2, 2)])
def test_driver_cell_sharded_and_hybrid_souporcell3(groups, shards):
    """soup_main with cell sharding and the hybrid restarts x cells mode (row 3), emulated on one GPU: `groups` cell
    groups of `shards` processes (threads here, one context each).  k = 16 with souporcell3, so bad_cluster_detection runs on cluster
    sizes gathered over the cell group.  Checked against the single-context driver on the whole matrix at north_star's tolerance."""
    import ctypes as C
    import threading
    import torch
    from souporcell_b200 import api

    v, csr = synth_csr(n_cells=420, n_loci=2600, k=5, mean_loci=130, seed=23)
    K, restarts = 16, 4
    with sp.Context(0) as whole:
        whole.load_csr(*csr)
        want = sp.main(whole, K, method=sp.METHOD_KHM, restarts=restarts, threads=restarts, seed=9, souporcell3=True)
    assert want.runs >= 2                                   # the re-initialisation runs happened
    cuts, parts = _split_cells(csr, shards)

    class Dev:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

    class Rendezvous:
        def __init__(self, n):
            self.bar, self.slots, self.tmp = threading.Barrier(n), [None] * n, None

    cell_groups = [Rendezvous(shards) for _ in range(groups)]       # the processes of one cell group
    positions = [Rendezvous(groups) for _ in range(shards)]         # the processes holding the same cell range in every group

    def allreduce_cb(g, s):
        rv = cell_groups[g]
        def cb(user, buf, n, stream):
            try:
                torch.cuda.synchronize()
                rv.slots[s] = torch.as_tensor(Dev(buf, n), device="cuda")
                rv.bar.wait()
                if s == 0:
                    acc = rv.slots[0].clone()
                    for t in rv.slots[1:]:
                        acc += t
                    rv.tmp = acc
                    torch.cuda.synchronize()
                rv.bar.wait()
                rv.slots[s].copy_(rv.tmp)
                torch.cuda.synchronize()
                rv.bar.wait()
                return 0
            except Exception:
                rv.bar.abort()
                return -1
        return api.ALLREDUCE_FN(cb)

    def allgather_cb(rv, me):
        def cb(user, send, recv, nbytes):
            try:
                rv.slots[me] = C.string_at(send, nbytes)
                rv.bar.wait()
                data = b"".join(rv.slots)
                C.memmove(recv, data, len(data))
                rv.bar.wait()
                return 0
            except Exception:
                rv.bar.abort()
                return -1
        return api.ALLGATHER_FN(cb)

    def bcast_cb(rv, me):
        def cb(user, buf, nbytes, root):
            try:
                if me == root:
                    rv.tmp = C.string_at(buf, nbytes)
                rv.bar.wait()
                if me != root:
                    C.memmove(buf, rv.tmp, nbytes)
                rv.bar.wait()
                return 0
            except Exception:
                rv.bar.abort()
                return -1
        return api.BCAST_FN(cb)

    results, errors = {}, []

    def run(g, s):
        try:
            keep = (allreduce_cb(g, s), allgather_cb(cell_groups[g], s), allgather_cb(positions[s], g), bcast_cb(positions[s], g))
            with sp.Context(0) as ctx:
                ctx.load_csr(*parts[s])
                results[(g, s)] = sp.main(ctx, K, method=sp.METHOD_KHM, restarts=restarts, threads=restarts, seed=9, souporcell3=True,
                                          proc_rank=g, proc_world=groups, allgather=keep[2] if groups > 1 else None,
                                          bcast=keep[3] if groups > 1 else None, cell_allreduce=keep[0], cell_allgather=keep[1],
                                          cell_group_size=shards, n_cells_global=csr[0])
        except Exception as e:                              # noqa: BLE001
            errors.append(e)
            for rv in cell_groups + positions:
                rv.bar.abort()

    threads = [threading.Thread(target=run, args=(g, s)) for g in range(groups) for s in range(shards)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    for g in range(groups):
        lp = np.concatenate([results[(g, s)].best_log_probs for s in range(shards)])
        r0 = results[(g, 0)]
        assert r0.runs == want.runs
        np.testing.assert_allclose(r0.best_loss, want.best_loss, rtol=1e-5)
        np.testing.assert_allclose(r0.best_theta, want.best_theta, rtol=1e-5, atol=1e-7)      # north_star: 1e-5 relative
        np.testing.assert_allclose(lp, want.best_log_probs, rtol=1e-5)
        assert np.array_equal(np.argmax(lp, axis=1), np.argmax(want.best_log_probs, axis=1))


# ------------------------------------------------------------------ cell sharding over NCCL inside the library (soup_comm)
def test_cell_comm_single_rank_nccl_agrees_with_oracle():
    """The library's own NCCL path on ONE GPU (ncclCommInitAll over one context, every locus range all-reduced over a communicator of
    size 1): the same kernels, streams and statistics layout as a multi-GPU cell group.  Checked like the emulated ranks above."""
    v, csr = synth_csr(n_cells=400, n_loci=2500, k=3, mean_loci=110, seed=17)
    od = orc.OracleData.from_csr(*csr)
    k, S = 3, 5
    theta0 = random_theta(np.random.default_rng(3), (S, k, csr[1]))
    for method in (sp.METHOD_EM, sp.METHOD_KHM):
        env = LONG_MODES["ranges3"] if method == sp.METHOD_KHM else {}      # one range, and three pipelined ranges
        os.environ.update(env)
        try:
            with sp.Context(0) as ctx:
                ctx.load_csr(*csr)
                comm = sp.Comm.create_all([ctx])[0]
                try:
                    assert comm.rank() == (0, 1)
                    assert comm.allgather(b"abcd") == b"abcd"
                    res = ctx.solve(k, method=method, n_streams=S, solves_per_stream=1, theta0=theta0, keep_trace=True, max_slots=3,
                                    cell_comm=comm, n_cells_global=csr[0])
                    traces = [ctx.get_trace(s) for s in range(S)]
                finally:
                    comm.close()
        finally:
            for key in env:
                os.environ.pop(key, None)
        assert res.stats["allreduce_calls"] % (3 if env else 1) == 0
        assert res.stats["allreduce_calls"] > 0 and res.stats["allreduce_bytes"] > 0
        want = [od.solve(theta0[s], method=method) for s in range(S)]
        for s in range(S):
            n_same = min(len(traces[s]), len(want[s]["trace"]))
            np.testing.assert_allclose([t[1] for t in traces[s][:n_same]], [t[1] for t in want[s]["trace"][:n_same]], rtol=1e-5)
            assert abs(len(traces[s]) - len(want[s]["trace"])) <= 1
        best = int(np.argmax([w["loss"] for w in want]))
        assert res.best_solve == best
        np.testing.assert_allclose(res.best_theta, want[best]["theta"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(res.best_log_probs, want[best]["log_probs"], rtol=1e-5)
        assert np.array_equal(np.argmax(res.best_log_probs, axis=1), np.argmax(want[best]["log_probs"], axis=1))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("k,s3,method", [(4, False, sp.METHOD_EM), (16, True, sp.METHOD_KHM)])
def test_cell_sharded_two_gpus_real_nccl(k, s3, method, tmp_path):
    """Two B200s, one context each holding half of the cells, ncclAllReduce of the M-step sums every iteration (soup_main with
    cell_comms; also through the drop-in CLI with --shard cells).  Against the single-GPU run at north_star's 1e-5."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    from souporcell_b200 import api, build, synth
    v = synth.generate(n_cells=900, n_loci=5000, k=5, mean_loci=150, seed=31)
    row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v)
    csr = (v.n_cells, int(i2l.shape[0]), row_ptr, locus, alt, ref)
    with sp.Context(0) as whole:
        whole.load_csr(*csr)
        want = sp.main(whole, k, method=method, restarts=6, threads=3, seed=5, souporcell3=s3)
    cuts, parts = api.split_cells(csr, 2)
    ctxs = [sp.Context(0), sp.Context(1)]
    comms = []
    try:
        for c, part in zip(ctxs, parts):
            c.load_csr(*part)
        comms = sp.Comm.create_all(ctxs)
        got = sp.main(ctxs, k, method=method, restarts=6, threads=3, seed=5, souporcell3=s3, cell_comms=comms)
    finally:
        for c in comms:
            c.close()
        for c in ctxs:
            c.close()
    assert got.runs == want.runs and got.best_log_probs.shape == want.best_log_probs.shape
    np.testing.assert_allclose(got.best_loss, want.best_loss, rtol=1e-5)
    np.testing.assert_allclose(got.best_theta, want.best_theta, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got.best_log_probs, want.best_log_probs, rtol=1e-5)
    assert np.array_equal(got.assignments, want.assignments)
    assert got.nnz_updates == want.nnz_updates
    # the CLI: --shard cells over 2 GPUs prints the same assignments as one GPU
    a, r, b = synth.write_files(v, str(tmp_path))
    args = ["-a", a, "-r", r, "-b", b, "-k", str(k), "--restarts", "6", "-t", "3", "--seed", "5", "-m", "khm" if method == sp.METHOD_KHM else "em", "-s", "true" if s3 else "false"]
    one = subprocess.run([build.CLI] + args + ["--gpus", "1"], capture_output=True, text=True)
    two = subprocess.run([build.CLI] + args + ["--gpus", "2", "--shard", "cells"], capture_output=True, text=True)       # one thread per GPU
    assert one.returncode == 0 and two.returncode == 0, two.stderr[-600:]
    two_p = subprocess.run([build.CLI] + args + ["--gpus", "2", "--shard", "cells"], capture_output=True, text=True, env=dict(os.environ, SOUP_CLI_PROCESSES="1"))
    assert two_p.returncode == 0 and two_p.stdout == two.stdout, two_p.stderr[-600:]                                  # one process per GPU: the same bytes
    assert "shard\tcells" in two.stderr
    # restart sharding over 2 GPUs = one PROCESS per GPU meeting in shared memory: byte-identical TSV, every solve logged exactly once
    threads2 = subprocess.run([build.CLI] + args + ["--gpus", "2", "--shard", "restarts"], capture_output=True, text=True)
    assert threads2.returncode == 0 and threads2.stdout == one.stdout, threads2.stderr[-600:]
    procs = subprocess.run([build.CLI] + args + ["--gpus", "2", "--shard", "restarts"], capture_output=True, text=True, env=dict(os.environ, SOUP_CLI_PROCESSES="1"))
    assert procs.returncode == 0, procs.stderr[-600:]
    assert procs.stdout == one.stdout and "gpus\t2\tshard\trestarts" in procs.stderr
    solves = lambda err: sorted(l for l in err.splitlines() if l.startswith(("thread ", "binomial")))
    assert [l.split(" best so far")[0] for l in solves(procs.stderr) if l.startswith("thread ")] == [l.split(" best so far")[0] for l in solves(one.stderr) if l.startswith("thread ")]
    assert [l for l in solves(procs.stderr) if l.startswith("binomial")] == [l for l in solves(one.stderr) if l.startswith("binomial")]
    assert procs.stderr.count("best total log probability") == 1
    col = lambda out: [ln.split("\t")[:2] for ln in out.splitlines()]
    assert col(one.stdout) == col(two.stdout) and len(col(one.stdout)) == v.n_cells
    f1 = np.array([ln.split("\t")[2:] for ln in one.stdout.splitlines()], dtype=np.float64)
    f2 = np.array([ln.split("\t")[2:] for ln in two.stdout.splitlines()], dtype=np.float64)
    np.testing.assert_allclose(f2, f1, rtol=1e-5)


def _proc_comm_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)      # only carries the 128-byte NCCL id
    torch.cuda.set_device(rank)
    v, csr = synth_csr(n_cells=300, n_loci=2000, k=3, mean_loci=120, seed=7)
    with sp.Context(rank) as ctx:
        ctx.load_csr(*csr)
        comm = sp.Comm.from_torch(ctx, dist)
        try:
            r = sp.main(ctx, 4, restarts=7, threads=3, seed=4, proc_rank=rank, proc_world=world, proc_comm=comm)
        finally:
            comm.close()
    out[rank] = (r.best_loss, r.best_log_probs.tobytes(), r.best_theta.tobytes(), r.nnz_updates)
    dist.barrier()
    dist.destroy_process_group()


def test_restart_sharded_two_processes_winner_exchange_over_nccl():
    """One rank per GPU (what torchrun launches): restarts shard s mod 2 and the once-per-run winner exchange runs over the library's own
    NCCL communicator (proc_comm) — bit-identical to the single-process run on every rank."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    v, csr = synth_csr(n_cells=300, n_loci=2000, k=3, mean_loci=120, seed=7)
    with sp.Context(0) as ctx:
        ctx.load_csr(*csr)
        want = sp.main(ctx, 4, restarts=7, threads=3, seed=4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_proc_comm_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in range(2):
        loss, lp, th, upd = out[rank]
        assert bits([loss])[0] == bits([want.best_loss])[0]
        assert lp == want.best_log_probs.tobytes() and th == want.best_theta.tobytes()
    assert out[0][3] + out[1][3] == want.nnz_updates


def test_two_cell_groups_on_one_gpu_agree_with_one_context():
    """soup_main with two cell GROUPS on the same GPU (each its own one-rank NCCL communicator and context, restarts split solve mod 2,
    both solver loops running side by side — what overlaps one group's all-reduce with the other's passes on a multi-GPU box): same winner
    and, at north_star's tolerance, the same result as the plain single-context driver; k = 16 with souporcell3 exercises the re-runs."""
    v, csr = synth_csr(n_cells=420, n_loci=2600, k=5, mean_loci=130, seed=23)
    for K, method, s3 in ((4, sp.METHOD_EM, False), (16, sp.METHOD_KHM, True)):
        with sp.Context(0) as whole:
            whole.load_csr(*csr)
            want = sp.main(whole, K, method=method, restarts=5, threads=5, seed=9, souporcell3=s3)
        ctxs = [sp.Context(0), sp.Context(0)]
        comms = []
        try:
            for c in ctxs:
                c.load_csr(*csr)
                comms.append(sp.Comm.create_all([c])[0])
            got = sp.main(ctxs, K, method=method, restarts=5, threads=5, seed=9, souporcell3=s3, cell_comms=comms, cell_groups=2)
        finally:
            for c in comms:
                c.close()
            for c in ctxs:
                c.close()
        assert got.runs == want.runs and got.best_log_probs.shape == want.best_log_probs.shape
        np.testing.assert_allclose(got.best_loss, want.best_loss, rtol=1e-5)
        np.testing.assert_allclose(got.best_theta, want.best_theta, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(got.best_log_probs, want.best_log_probs, rtol=1e-5)
        assert np.array_equal(got.assignments, want.assignments) and got.nnz_updates == want.nnz_updates
