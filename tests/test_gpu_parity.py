"""GPU parity: libsoupgpu (through the C ABI) vs the CPU oracle on identical seeded inputs.
Bit-exact (np.array_equal on the raw f32 bits) unless stated otherwise."""
import numpy as np
import pytest

from helpers import random_theta, synth_csr
from oracle import oracle_py as orc
import souporcell_b200 as sp

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module", params=["default", "long_role"])
def small(request):
    """`long_role` lowers the long-locus threshold so the M kernel's deep narrow-group role runs on this small matrix."""
    import os
    v, csr = synth_csr()
    if request.param == "long_role":
        os.environ["SOUP_LONG_THRESHOLD"] = "12"
    try:
        ctx = sp.Context(0).load_csr(*csr)
    finally:
        os.environ.pop("SOUP_LONG_THRESHOLD", None)
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
