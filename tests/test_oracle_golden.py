"""CPU: pin the oracle (oracle/, C++ restatement of the reference) against tests/golden/.
The reference has no tests of its own; see tests/golden/make_golden.py for where each vector comes from."""
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle_py as orc

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name)) as fh:
        return json.load(fh)


def write_mtx(tmp_path, n_loci, n_cells, entries, crlf=False):
    nl = "\r\n" if crlf else "\n"
    for name, col in (("alt.mtx", 2), ("ref.mtx", 3)):
        with open(tmp_path / name, "w", newline="") as fh:
            fh.write("%%MatrixMarket matrix coordinate real general" + nl + "% golden" + nl)
            fh.write(f"{n_loci} {n_cells} {len(entries)}" + nl)
            for e in entries:
                fh.write(f"{e[0]} {e[1]} {e[col]}" + nl)
    return str(tmp_path / "alt.mtx"), str(tmp_path / "ref.mtx")


def test_chacha_public_vectors():
    kat = load("chacha_kat.json")
    key = bytes.fromhex(kat["key"])
    for rounds, hexs in kat["rounds"].items():
        assert orc.chacha_words(key, 16, 0, int(rounds)).astype("<u4").tobytes().hex() == hexs
    assert orc.chacha_words(key, 16, 16, 20).astype("<u4").tobytes().hex() == kat["rounds20_block1"]
    # word-granular seeking: any offset reads the same stream
    full = orc.chacha_words(key, 100, 0, 12)
    assert np.array_equal(orc.chacha_words(key, 37, 21, 12), full[21:58])


def test_stdrng_draws_are_one_word_each():
    # rand 0.9 StandardUniform: f32 = (u32 >> 8) * 2^-24, u8 = low byte; init_cluster_centers_uniform clamps to [1e-4, 0.9999]
    seed = bytes(range(32))
    words = orc.chacha_words(seed, 24, 5, 12)
    th = orc.init_uniform(seed, 5, 2, 12)
    want = np.clip((words >> 8).astype(np.float32) * np.float32(2.0 ** -24), np.float32(1e-4), np.float32(0.9999)).reshape(2, 12)
    assert np.array_equal(th, want)
    # thread seeds: 32 consecutive low bytes of the master stream per thread, runs continue the stream (S:61-62, 77-80, 855-861)
    master = orc.chacha_words(bytes([4] * 32), 32 * 6, 0, 12)
    s0, s1 = orc.thread_seeds(4, 3, 0), orc.thread_seeds(4, 3, 1)
    assert np.array_equal(s0.reshape(-1), (master[:96] & 0xFF).astype(np.uint8))
    assert np.array_equal(s1.reshape(-1), (master[96:192] & 0xFF).astype(np.uint8))


@pytest.mark.parametrize("crlf", [False, True])
def test_loader_micro_fixture(tmp_path, crlf):
    g = load("loader_micro.json")
    alt, ref = write_mtx(tmp_path, g["n_loci"], g["n_cells"], g["entries"], crlf)
    d = orc.OracleData.from_mtx(alt, ref, g["min_alt"], g["min_ref"])
    ex, want = d.export(), g["expect"]
    for k in ("row_ptr", "locus", "alt", "ref", "index_to_locus"):
        assert ex[k].tolist() == want[k], k
    assert ex["total_alleles"].tolist() == want["total_alleles"]
    lbc = {"0": 0.0, "ln2": math.log(2.0), "ln3": math.log(3.0)}
    assert np.array_equal(ex["lbc"], np.array([lbc[t] for t in want["lbc"]], dtype=np.float32))


def test_loader_duplicates_last_wins_but_all_lines_tally(tmp_path):
    g = load("loader_micro.json")
    alt, ref = write_mtx(tmp_path, 1, 4, g["dup_entries"])
    w = g["dup_expect"]
    ex = orc.OracleData.from_mtx(alt, ref, w["min_alt"], w["min_ref"]).export()
    for k in ("row_ptr", "locus", "alt", "ref", "index_to_locus"):
        assert ex[k].tolist() == w[k], k


def test_rust_float_formatting():
    for v, plain, w8 in load("format_f32.json"):
        assert orc.format_f32(v) == plain
        assert orc.format_f32(v, True) == w8


def test_bad_cluster_detection_cutoffs():
    for case in load("bad_clusters.json"):
        k = case["k"]
        rows = []
        for c, n in enumerate(case["sizes"]):
            r = np.full(k, -10.0, dtype=np.float32)
            r[c] = -1.0
            rows += [r] * n
        got = orc.bad_cluster_detection(case["run"], k, np.stack(rows))
        assert got == case["expect"], case


def test_argmax_tie_rules_differ():
    # per-iteration counts / bad-cluster detection keep the LAST maximum (S:159, S:254: max_by), the TSV the FIRST (S:134-141)
    lp = np.zeros((300, 16), dtype=np.float32)          # every cell tied across all clusters
    got = orc.bad_cluster_detection(1, 16, lp)
    # all 300 cells counted in cluster 15 -> ascending stable order 0..14 (size 0), 15 (size 300): positions 0, 1 and 15
    assert got == [0, 1, 15]


def test_em_khm_micro_against_independent_numpy():
    g = load("em_micro.json")
    tol = g["tolerance"]
    lm = load("loader_micro.json")["expect"]
    d = orc.OracleData.from_csr(4, 3, np.array(lm["row_ptr"]), np.array(lm["locus"]), np.array(lm["alt"]), np.array(lm["ref"]))
    for c in g["cases"]:
        m = 1 if c["method"] == "khm" else 0
        th0 = np.array(c["theta0"], dtype=np.float32)
        st = d.step(th0, method=m, temp_step=c["temp_step"])
        np.testing.assert_allclose(st["log_probs"], np.array(c["log_probs"]), rtol=tol)
        np.testing.assert_allclose(st["weights"], np.array(c["weights"]), rtol=20 * tol, atol=1e-30)
        np.testing.assert_allclose(st["loss"], c["loss"], rtol=tol)
        np.testing.assert_allclose(st["theta_out"], np.array(c["theta1"]), rtol=20 * tol)
        so = d.solve(th0, method=m)
        assert so["iterations"] == c["solve"]["iterations"]
        assert [t[0] for t in so["trace"]] == [t[0] for t in c["solve"]["trace"]]
        np.testing.assert_allclose([t[1] for t in so["trace"]], [t[1] for t in c["solve"]["trace"]], rtol=50 * tol)
        np.testing.assert_allclose(so["log_probs"], np.array(c["solve"]["log_probs"]), rtol=200 * tol)
        np.testing.assert_allclose(so["theta"], np.array(c["solve"]["theta"]), rtol=200 * tol)


def test_solve_structure_invariants():
    """Properties the reference's control flow implies (S:214-222): >= 9 iterations, 9 temp steps in order,
    loss of the last iteration returned, theta clamped to [0.01, 0.99]."""
    from helpers import random_theta, synth_csr
    _, csr = synth_csr(n_cells=120, n_loci=800, k=3, mean_loci=60, seed=11)
    d = orc.OracleData.from_csr(*csr)
    th0 = random_theta(np.random.default_rng(0), (3, csr[1]))
    for m in (0, 1):
        so = d.solve(th0, method=m)
        steps = [t[0] for t in so["trace"]]
        assert so["iterations"] >= 9 and steps == sorted(steps) and set(steps) == set(range(9))
        assert np.float32(so["trace"][-1][1]) == so["loss"]
        assert so["theta"].min() >= np.float32(0.01) and so["theta"].max() <= np.float32(0.99)
        capped = d.solve(th0, method=m, iteration_cap=5)
        assert capped["iterations"] == 5


def test_inner_threads_are_bit_identical_to_the_serial_loop():
    """The oracle's multi-threaded iteration (Params::inner_threads, used by the full-size parity tests) evaluates the reference's own
    per-cell and per-(cluster, locus) chains concurrently: every bit of the trace, the log-probabilities and the centers must equal
    the serial loop of S:225-243, for EM and KHM, with locked clusters, for any thread count."""
    from souporcell_b200 import synth
    v = synth.generate(n_cells=400, n_loci=2500, k=5, mean_loci=150, seed=31)
    row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v)
    od = orc.OracleData.from_csr(v.n_cells, int(i2l.shape[0]), row_ptr, locus, alt, ref)
    rng = np.random.default_rng(5)
    th0 = np.clip(rng.random((5, int(i2l.shape[0])), dtype=np.float32), 1e-4, 0.9999).astype(np.float32)
    for method in (0, 1):
        for locked in ((), (1, 3)):
            want = od.solve(th0, method=method, locked=locked, iteration_cap=25)
            for nt in (2, 3, 8):
                got = od.solve(th0, method=method, locked=locked, iteration_cap=25, inner_threads=nt)
                assert got["iterations"] == want["iterations"]
                assert [(t[0], np.float32(t[1]).tobytes(), np.float32(t[2]).tobytes(), t[3]) for t in got["trace"]] == \
                       [(t[0], np.float32(t[1]).tobytes(), np.float32(t[2]).tobytes(), t[3]) for t in want["trace"]]
                assert got["theta"].tobytes() == want["theta"].tobytes() and got["log_probs"].tobytes() == want["log_probs"].tobytes()


def test_complete_config0_solves_against_independent_restatement():
    """Whole em() / khm() solves at synth config 0 (300 cells, k = 3, the four restarts of --seed 4 --threads 4) from the independent
    pure-Python restatement of main.rs:189-483 (tests/golden/make_config0_golden.py): the oracle must take the same number of iterations
    through the same temperature steps, reach the same losses (float64-vs-libm log/exp: 1e-4 relative over a whole solve), pick the same best restart
    and give every cell the same hard assignment."""
    from souporcell_b200 import synth
    g = load("config0_solves.json")
    v = synth.generate_config(0)
    row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v)
    assert (v.n_cells, int(i2l.shape[0]), int(locus.shape[0])) == (g["n_cells"], g["loci_used"], g["nnz"])
    d = orc.OracleData.from_csr(v.n_cells, int(i2l.shape[0]), row_ptr, locus, alt, ref)
    seeds = orc.thread_seeds(g["seed"], g["threads"], 0)
    for name, m in (("em", 0), ("khm", 1)):
        want = g["methods"][name]
        losses = []
        for rec in want["solves"]:
            theta0 = orc.init_uniform(bytes(seeds[rec["thread"]]), 0, g["k"], g["loci_used"])
            so = d.solve(theta0, method=m)
            assert so["iterations"] == rec["iterations"], (name, rec["thread"])
            assert [t[0] for t in so["trace"]] == rec["temp_steps"]
            np.testing.assert_allclose([t[1] for t in so["trace"]], rec["losses"], rtol=1e-4)
            assert np.array_equal(np.argmax(so["log_probs"], axis=1), np.array(rec["assignments"]))
            losses.append(so["loss"])
        # the driver on the same seeds: same winner (S:110-125), same assignments
        r = d.main(g["k"], method=m, restarts=g["restarts"], threads=g["threads"], seed=g["seed"])
        best = want["best_thread"]
        assert r["has_best"] and np.float32(r["best_loss"]) == np.float32(max(losses))
        assert np.array_equal(np.argmax(r["best_log_probs"], axis=1), np.array(want["solves"][int(np.argmax(losses))]["assignments"]))
        assert abs(want["solves"][best]["loss"] - r["best_loss"]) <= 1e-4 * abs(r["best_loss"])
