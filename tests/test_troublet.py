"""troublet (/root/reference/troublet/src/main.rs "T:"): the CPU restatement (oracle/troublet_oracle.cpp), the host
pieces of libsoupgpu, and — on a GPU — soup_troublet / the `troublet` CLI against the oracle.

The reference holds no fixture for troublet and its third-party arithmetic (statrs ln_gamma / ln_binomial) is restated from the
published source: PARITY UNPINNED.  f64 tolerance: the device `log` is within 1 ulp of the host's, sums run over a few hundred
terms, so log-probabilities are compared at 1e-10 relative; calls (status / assignment) must be identical except where a
posterior sits within 1e-9 of a threshold."""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np
import pytest

import souporcell_b200 as sp
from souporcell_b200 import build, synth
from oracle import oracle_py as orc

TRB_LIB = os.path.join(os.path.dirname(orc.CLI_PATH), "libtroublet_oracle.so")
TRB_CLI = os.path.join(os.path.dirname(orc.CLI_PATH), "troublet_oracle")


def trb():
    orc.build()
    L = C.CDLL(TRB_LIB)
    L.trb_ln_gamma.restype = C.c_double
    L.trb_ln_gamma.argtypes = [C.c_double]
    L.trb_ln_binomial.restype = C.c_double
    L.trb_ln_binomial.argtypes = [C.c_uint64, C.c_uint64]
    L.trb_run.restype = C.c_int64
    return L


def oracle_run(alt, ref, clusters, n, k, **kw):
    L = trb()
    out = dict(status=np.zeros(n, np.int32), best_singlet=np.zeros(n, np.int32), doublet1=np.zeros(n, np.int32), doublet2=np.zeros(n, np.int32),
               singlet_posterior=np.zeros(n), doublet_posterior=np.zeros(n), log_prob_singleton=np.zeros(n), log_prob_doublet=np.zeros(n),
               cluster_logprobs=np.zeros((n, k)))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ko, rounds, removed = C.c_int32(), C.c_int32(), C.c_int64()
    err = C.create_string_buffer(256)
    L.trb_run.argtypes = [C.c_char_p] * 3 + [C.c_double] * 4 + [C.c_int64, C.c_int32] + [C.c_void_p] * 9 + [C.POINTER(C.c_int32)] * 2 + [C.POINTER(C.c_int64), C.c_char_p, C.c_int32]
    rc = L.trb_run(os.fsencode(alt), os.fsencode(ref), os.fsencode(clusters), kw.get("p_soup", 0.0), kw.get("doublet_prior", 0.5),
                   kw.get("doublet_threshold", 0.9), kw.get("singlet_threshold", 0.9), n, k, *[p(out[f]) for f in
                   ("status", "best_singlet", "doublet1", "doublet2", "singlet_posterior", "doublet_posterior", "log_prob_singleton", "log_prob_doublet", "cluster_logprobs")],
                   C.byref(ko), C.byref(rounds), C.byref(removed), err, 256)
    if rc < 0:
        raise RuntimeError(err.value.decode())
    out.update(rounds=rounds.value, removed=removed.value)
    return out


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    """A small matrix with 12 % inter-donor doublets, clustered by the CPU oracle of the clustering step (clusters_tmp.tsv)."""
    d = str(tmp_path_factory.mktemp("trb"))
    v = synth.generate(n_cells=500, n_loci=2500, k=4, mean_loci=220, seed=31, doublet_frac=0.12)
    alt, ref, bc = synth.write_files(v, d)
    orc.build()
    r = subprocess.run([orc.CLI_PATH, "-a", alt, "-r", ref, "-b", bc, "-k", "4", "--restarts", "6", "-t", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-300:]
    clusters = os.path.join(d, "clusters_tmp.tsv")
    with open(clusters, "w") as fh:
        fh.write(r.stdout)
    return dict(alt=alt, ref=ref, clusters=clusters, v=v)


# ------------------------------------------------------------------ CPU: oracle + host logic
def test_statrs_restatement_against_math():
    L = trb()
    for x in [1.0, 1.5, 2.0, 3.25, 10.0, 57.3, 171.0, 1234.5, 1e5]:
        assert L.trb_ln_gamma(x) == pytest.approx(math.lgamma(x), rel=1e-13, abs=1e-14)
    for n, k in [(0, 0), (1, 0), (1, 1), (5, 2), (40, 17), (170, 85), (171, 3), (400, 200)]:
        want = math.lgamma(n + 1) - math.lgamma(k + 1) - math.lgamma(n - k + 1)
        assert L.trb_ln_binomial(n, k) == pytest.approx(want, rel=1e-12, abs=1e-12)
    assert L.trb_ln_binomial(3, 4) == -math.inf


def test_clusters_loader_and_raw_matrix(dataset):
    names, losses, n_clusters = sp.load_clusters(dataset["clusters"])
    lines = open(dataset["clusters"]).read().splitlines()
    assert len(names) == len(lines) == dataset["v"].n_cells and n_clusters == 4
    for i in (0, 7, len(lines) - 1):
        tok = lines[i].split("\t")
        assert names[i] == tok[0] and np.array_equal(losses[i], np.array([float(t) for t in tok[2:]]))
    raw = sp.load_mtx(dataset["alt"], dataset["ref"], raw=True)
    v = dataset["v"]
    assert raw.n_loci == v.n_loci and raw.nnz == v.nnz          # no filter, every line
    order = np.lexsort((v.locus, v.cell))                        # file order is (locus, cell): per cell, ascending locus
    assert np.array_equal(raw.locus, v.locus[order].astype(np.uint32)) and np.array_equal(raw.alt, v.alt[order]) and np.array_equal(raw.ref, v.ref[order])


def test_raw_view_asserts_that_the_two_matrices_match(dataset, tmp_path):
    """T:315-320: troublet takes locus and cell from the REF line and asserts they equal the alt line's ("allele matrices do not
    match"); souporcell itself reads them from the alt line only (S:622-626), so the filtered view must keep accepting such files."""
    lines = open(dataset["ref"]).read().splitlines()
    tok = lines[5].split()
    for which, repl in (("locus", f"{int(tok[0]) + 1} {tok[1]} {tok[2]}"), ("cell", f"{tok[0]} {int(tok[1]) % dataset['v'].n_cells + 1} {tok[2]}")):
        bad = str(tmp_path / f"ref_bad_{which}.mtx")
        with open(bad, "w") as fh:
            fh.write("\n".join(lines[:5] + [repl] + lines[6:]) + "\n")
        with pytest.raises(sp.SoupError, match="allele matrices do not match"):
            sp.load_mtx(dataset["alt"], bad, raw=True)
        assert sp.load_mtx(dataset["alt"], bad).nnz == sp.load_mtx(dataset["alt"], dataset["ref"]).nnz


def test_oracle_calls_doublets_and_iterates(dataset):
    v = dataset["v"]
    got = oracle_run(dataset["alt"], dataset["ref"], dataset["clusters"], v.n_cells, 4)
    assert got["rounds"] >= 2 and got["removed"] > 0                  # at least one removal round, then a clean one
    assert (got["status"] == 2).sum() == got["removed"]
    # the doublet posterior is a two-way softmax of the best singlet and the best doublet (T:145-167)
    m = np.maximum(got["log_prob_singleton"], got["log_prob_doublet"])
    lse = m + np.log(np.exp(got["log_prob_singleton"] - m) + np.exp(got["log_prob_doublet"] - m))
    np.testing.assert_allclose(got["doublet_posterior"], np.exp(got["log_prob_doublet"] - lse), rtol=1e-12)
    assert np.array_equal(got["best_singlet"], np.argmax(got["cluster_logprobs"], axis=1))


def test_oracle_cli_output_shape(dataset):
    r = subprocess.run([TRB_CLI, "--alts", dataset["alt"], "--refs", dataset["ref"], "--clusters", dataset["clusters"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-300:]
    rows = r.stdout.splitlines()
    assert rows[0].split("\t")[:7] == ["barcode", "status", "assignment", "singlet_posterior", "doublet_posterior", "log_prob_singleton", "log_prob_doublet"]
    assert len(rows) == dataset["v"].n_cells + 1 and all(len(x.split("\t")) == 7 + 4 for x in rows[1:])
    assert "cells removed this round" in r.stderr


# ------------------------------------------------------------------ GPU: parity with the oracle
def _compare(got, want, thresholds=(0.9, 0.9)):
    assert got["rounds"] == want["rounds"] and got["removed"] == want["removed"]
    np.testing.assert_allclose(got["cluster_logprobs"], want["cluster_logprobs"], rtol=1e-10)
    np.testing.assert_allclose(got["log_prob_singleton"], want["log_prob_singleton"], rtol=1e-10)
    np.testing.assert_allclose(got["log_prob_doublet"], want["log_prob_doublet"], rtol=1e-10)
    np.testing.assert_allclose(got["singlet_posterior"], want["singlet_posterior"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(got["doublet_posterior"], want["doublet_posterior"], rtol=1e-7, atol=1e-300)
    edge = (np.abs(want["singlet_posterior"] - thresholds[1]) < 1e-9) | (np.abs(want["doublet_posterior"] - thresholds[0]) < 1e-9) | \
           (np.abs(want["log_prob_singleton"] - want["log_prob_doublet"]) < 1e-9 * np.abs(want["log_prob_singleton"]))
    for f in ("status", "best_singlet", "doublet1", "doublet2"):
        assert np.array_equal(got[f][~edge], want[f][~edge]), f


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [{}, {"doublet_prior": 0.2, "doublet_threshold": 0.8, "singlet_threshold": 0.95}])
def test_gpu_troublet_matches_oracle(dataset, kw):
    v = dataset["v"]
    want = oracle_run(dataset["alt"], dataset["ref"], dataset["clusters"], v.n_cells, 4, **kw)
    names, losses, n_clusters = sp.load_clusters(dataset["clusters"])
    raw = sp.load_mtx(dataset["alt"], dataset["ref"], raw=True)
    with sp.Context(0) as ctx:
        got = sp.troublet(ctx, raw, losses, n_clusters=n_clusters, barcodes=names, **kw)
    assert want["removed"] > 0
    _compare(got, want, (kw.get("doublet_threshold", 0.9), kw.get("singlet_threshold", 0.9)))
    assert got["device_ms"] > 0


@pytest.mark.gpu
def test_gpu_troublet_cli_matches_oracle_cli(dataset):
    args = ["--alts", dataset["alt"], "--refs", dataset["ref"], "--clusters", dataset["clusters"]]
    ours = subprocess.run([build.TROUBLET_CLI] + args, capture_output=True, text=True)
    theirs = subprocess.run([TRB_CLI] + args, capture_output=True, text=True)
    assert ours.returncode == 0 and theirs.returncode == 0, ours.stderr[-400:] + theirs.stderr[-400:]
    a, b = ours.stdout.splitlines(), theirs.stdout.splitlines()
    assert a[0] == b[0] and len(a) == len(b)
    same_calls = 0
    for x, y in zip(a[1:], b[1:]):
        tx, ty = x.split("\t"), y.split("\t")
        assert tx[0] == ty[0]
        same_calls += tx[1:3] == ty[1:3]
        np.testing.assert_allclose([float(t) for t in tx[5:]], [float(t) for t in ty[5:]], rtol=1e-10)
    assert same_calls >= len(a) - 2                     # a call may differ only for a posterior sitting on a threshold
    pick = lambda s: sorted(l for l in s.splitlines() if l.startswith("removing") or "cells removed" in l)
    assert pick(ours.stderr) == pick(theirs.stderr)


def _rust_f64(v):
    """Rust `{}` for f64: shortest round-trip digits (Python's repr gives the same digits), never in scientific notation."""
    v = float(v)
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "-inf" if v < 0 else "inf"
    if v == 0:
        return "-0" if math.copysign(1, v) < 0 else "0"
    from decimal import Decimal
    s = format(Decimal(repr(v)), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s


def test_tsv_writer_and_f64_formatting(tmp_path):
    """soup_troublet_write (T:36-41, T:233-243) and the oracle's printer on crafted values: Rust's f64 Display, the status names,
    `a/b` assignments for doublets."""
    from souporcell_b200 import api
    vals = [1.0, 0.1, -41.24912227493646, 1e-7, 1.1015978990922601e-07, 1e21, 123456789012345680.0, 0.0, -0.0, float("inf"), float("-inf"),
            float("nan"), 5e-324, 1.7976931348623157e308, 0.999999996610466, -0.5]
    L = trb()
    L.trb_format_f64.argtypes = [C.c_double, C.c_char_p, C.c_int32]
    buf = C.create_string_buffer(512)
    for v in vals:
        assert L.trb_format_f64(v, buf, 512) > 0 and buf.value.decode() == _rust_f64(v), v
    n, k = 4, 3
    out = dict(status=np.array([0, 1, 2, 3], np.int32), best_singlet=np.array([2, 0, 1, 1], np.int32), doublet1=np.array([0, 0, 1, 2], np.int32),
               doublet2=np.array([1, 2, 2, 0], np.int32), singlet_posterior=np.array(vals[:4]), doublet_posterior=np.array(vals[4:8]),
               log_prob_singleton=np.array(vals[8:12]), log_prob_doublet=np.array(vals[12:16]),
               cluster_logprobs=np.arange(n * k, dtype=np.float64).reshape(n, k) * -1.25)
    r = api._TroubletResult(*[out[f].ctypes.data_as(C.c_void_p) for f, _ in api._TroubletResult._fields_[:9]])
    names = ["AAAC-1", "AAAG-1", "AAAT-1", "AACA-1"]
    blob = b"".join(x.encode() + b"\0" for x in names) + b"\0"
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = str(tmp_path / "t.tsv")
    f = libc.fopen(path.encode(), b"w")
    assert sp.lib().soup_troublet_write(f, blob, n, k, C.byref(r)) == 0
    libc.fclose(f)
    rows = open(path).read().split("\n")
    assert rows[0] == "barcode\tstatus\tassignment\tsinglet_posterior\tdoublet_posterior\tlog_prob_singleton\tlog_prob_doublet\tcluster0\tcluster1\tcluster2"
    want_assign = ["2", "0", "1/2", "2/0"]
    want_status = ["singlet", "unassigned", "doublet", "unassigned"]
    for i in range(n):
        tok = rows[1 + i].split("\t")
        assert tok[:3] == [names[i], want_status[i], want_assign[i]]
        assert tok[3:7] == [_rust_f64(out[f][i]) for f in ("singlet_posterior", "doublet_posterior", "log_prob_singleton", "log_prob_doublet")]
        assert tok[7:] == [_rust_f64(x) for x in out["cluster_logprobs"][i]]
    assert rows[1 + n] == ""


def test_clusters_loader_edge_cases(tmp_path):
    p = tmp_path / "c.tsv"
    p.write_text("A-1\t0\t-1.5\t-2.5\r\nB-1\t1\t-3\t-0.25\n  C-1\t1\t-7e0\t-1e-3  \n")     # CRLF, surrounding blanks (line.trim())
    names, losses, n_clusters = sp.load_clusters(str(p))
    assert names == ["A-1", "B-1", "C-1"] and n_clusters == 2
    assert np.array_equal(losses, np.array([[-1.5, -2.5], [-3.0, -0.25], [-7.0, -0.001]]))
    p.write_text("A-1\t0\t-1.5\t-2.5\nB-1\t1\t-3\n")                                       # ragged rows
    with pytest.raises(sp.SoupError):
        sp.load_clusters(str(p))
    p.write_text("A-1\tx\t-1.5\n")                                                          # parse::<usize>() fails
    with pytest.raises(sp.SoupError):
        sp.load_clusters(str(p))
    with pytest.raises(sp.SoupError):
        sp.load_clusters(str(tmp_path / "missing.tsv"))


def test_oracle_against_independent_python_golden(tmp_path):
    """tests/golden/troublet_micro.json: call_doublets restated independently in pure Python (math.lgamma instead of the Lanczos series)
    on a 36 x 30 matrix with two planted doublets.  The C++ oracle must make the same calls and agree on every log-probability."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "troublet_micro.json")))
    n, L, k = g["n_cells"], g["n_loci"], g["k"]
    for name, col in (("alt.mtx", 3), ("ref.mtx", 2)):
        with open(tmp_path / name, "w") as fh:
            fh.write("%%MatrixMarket matrix coordinate real general\n% golden\n")
            fh.write(f"{L} {n} {len(g['entries'])}\n")
            for e in g["entries"]:
                fh.write(f"{e[0] + 1} {e[1] + 1} {e[col]}\n")
    with open(tmp_path / "clusters.tsv", "w") as fh:
        for c in range(n):
            best = int(np.argmax(g["losses"][c]))
            fh.write(f"CELL{c:03d}-1\t{best}\t" + "\t".join(repr(float(x)) for x in g["losses"][c]) + "\n")
    got = oracle_run(str(tmp_path / "alt.mtx"), str(tmp_path / "ref.mtx"), str(tmp_path / "clusters.tsv"), n, k, doublet_prior=g["doublet_prior"])
    assert got["rounds"] == g["rounds"] and got["removed"] == len(g["removed"])
    assert sorted(np.nonzero(got["status"] == 2)[0].tolist()) == g["removed"] == g["planted_doublets"]
    tol = g["tolerance"]
    for c, want in enumerate(g["cells"]):
        assert got["status"][c] == want["status"] and got["best_singlet"][c] == want["best_singlet"], c
        assert [got["doublet1"][c], got["doublet2"][c]] == want["doublet"], c
        np.testing.assert_allclose(got["cluster_logprobs"][c], want["cluster_logprobs"], rtol=tol)
        np.testing.assert_allclose([got["log_prob_singleton"][c], got["log_prob_doublet"][c]], [want["log_prob_singleton"], want["log_prob_doublet"]], rtol=tol)
        np.testing.assert_allclose([got["singlet_posterior"][c], got["doublet_posterior"][c]], [want["singlet_posterior"], want["doublet_posterior"]], rtol=1e-7, atol=1e-300)


def test_f64_display_against_the_reference_readme_sample(tmp_path):
    """tests/golden/readme_troublet_sample.json: ten rows of clusters.tsv as the reference's own troublet printed them (README.md:147-157).
    Every number there is Rust's f64 `{}`: parsing it and printing it again must give the same text — through the oracle's printer and
    through soup_troublet_write (T:233-243)."""
    from souporcell_b200 import api
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "readme_troublet_sample.json")))
    assert g["header"][:3] == ["barcode", "status", "assignment"] and g["header"][-2:] == ["cluster0", "cluster1"]
    rows = g["rows"]
    L = trb()
    L.trb_format_f64.argtypes = [C.c_double, C.c_char_p, C.c_int32]
    buf = C.create_string_buffer(512)
    n_checked = 0
    for r in rows:
        assert r[1] == "singlet" and r[2] in ("0", "1")
        for tok in r[3:]:
            assert L.trb_format_f64(float(tok), buf, 512) > 0 and buf.value.decode() == tok, tok
            assert _rust_f64(float(tok)) == tok
            n_checked += 1
    assert n_checked == 40
    # the same values through the library's writer: log_prob_singleton / log_prob_doublet / cluster0 / cluster1 columns
    n, k = len(rows), 2
    num = np.array([[float(t) for t in r[3:]] for r in rows])
    out = dict(status=np.zeros(n, np.int32), best_singlet=np.array([int(r[2]) for r in rows], np.int32), doublet1=np.zeros(n, np.int32),
               doublet2=np.ones(n, np.int32), singlet_posterior=np.ones(n), doublet_posterior=np.zeros(n),
               log_prob_singleton=np.ascontiguousarray(num[:, 0]), log_prob_doublet=np.ascontiguousarray(num[:, 1]),
               cluster_logprobs=np.ascontiguousarray(num[:, 2:]))
    res = api._TroubletResult(*[out[f].ctypes.data_as(C.c_void_p) for f, _ in api._TroubletResult._fields_[:9]])
    blob = b"".join(r[0].encode() + b"\0" for r in rows) + b"\0"
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = str(tmp_path / "readme.tsv")
    f = libc.fopen(path.encode(), b"w")
    assert sp.lib().soup_troublet_write(f, blob, n, k, C.byref(res)) == 0
    libc.fclose(f)
    got = [ln.split("\t") for ln in open(path).read().splitlines()[1:]]
    for r, t in zip(rows, got):
        assert t[:3] == r[:3] and t[5:] == r[3:], (r, t)          # columns 3-4 are the posteriors the sample's version did not print


def test_troublet_writer_many_blocks(tmp_path):
    """soup_troublet_write formats blocks of rows in parallel: several blocks, all four statuses, special values; every token against
    the independent Python rendering of Rust's f64 Display."""
    from souporcell_b200 import api
    n, k = 3000, 5
    rng = np.random.default_rng(3)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 5e-324, 1.7976931348623157e308, 1e21, 1e-7, 0.1, -41.24912227493646, 0.999999996610466])
    draw = lambda shape: np.where(rng.random(shape) < 0.05, special[rng.integers(0, special.shape[0], shape)], -rng.random(shape) * 10.0 ** rng.integers(-9, 6, shape))
    out = dict(status=rng.integers(0, 4, n).astype(np.int32), best_singlet=rng.integers(0, k, n).astype(np.int32), doublet1=rng.integers(0, k, n).astype(np.int32),
               doublet2=rng.integers(0, k, n).astype(np.int32), singlet_posterior=draw(n), doublet_posterior=draw(n), log_prob_singleton=draw(n),
               log_prob_doublet=draw(n), cluster_logprobs=np.ascontiguousarray(draw((n, k))))
    res = api._TroubletResult(*[out[f].ctypes.data_as(C.c_void_p) for f, _ in api._TroubletResult._fields_[:9]])
    names = [f"CELL{'AC' * int(rng.integers(0, 9))}{i}-1" for i in range(n)]
    blob = b"".join(x.encode() + b"\0" for x in names) + b"\0"
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = str(tmp_path / "big.tsv")
    f = libc.fopen(path.encode(), b"w")
    assert sp.lib().soup_troublet_write(f, blob, n, k, C.byref(res)) == 0
    libc.fclose(f)
    rows = open(path).read().split("\n")
    assert len(rows) == n + 2 and rows[-1] == "" and rows[0].endswith("\tcluster4")
    status = ["singlet", "unassigned", "doublet", "unassigned"]
    for i in range(n):
        tok = rows[1 + i].split("\t")
        st = int(out["status"][i])
        want_assign = str(out["best_singlet"][i]) if st in (0, 1) else f"{out['doublet1'][i]}/{out['doublet2'][i]}"
        assert tok[:3] == [names[i], status[st], want_assign], i
        want = [out[f][i] for f in ("singlet_posterior", "doublet_posterior", "log_prob_singleton", "log_prob_doublet")] + list(out["cluster_logprobs"][i])
        assert tok[3:] == [_rust_f64(x) for x in want], i


def test_clusters_loader_many_rows_and_earliest_error(tmp_path):
    """soup_clusters_load parses rows in parallel blocks: values must equal Python's float() of every token (both correctly rounded),
    whatever the spelling, and with several bad lines the EARLIEST one is reported, as a sequential reader would hit it."""
    n, k = 3000, 6
    rng = np.random.default_rng(8)
    vals = -rng.random((n, k)) * 10.0 ** rng.integers(-12, 9, (n, k))
    spell = [lambda v: repr(float(v)), lambda v: f"{v:.17e}", lambda v: f"{v:.6f}", lambda v: "+" + repr(float(abs(v))), lambda v: f"{v:.3E}"]
    rows, want = [], np.empty((n, k))
    for i in range(n):
        toks = []
        for c in range(k):
            t = spell[int(rng.integers(0, len(spell)))](vals[i, c])
            toks.append(t)
            want[i, c] = float(t)
        rows.append(f"BC{i}-1\t{i % 5}\t" + "\t".join(toks))
    rows[7] = "BC7-1\t+4\tinf\t-inf\tnan\t-0\t1e-400\t1e400"          # parse::<usize>() takes '+'; overflow / underflow go to inf / 0
    want[7] = [np.inf, -np.inf, np.nan, -0.0, 0.0, np.inf]
    p = tmp_path / "c.tsv"
    p.write_text("\n".join(rows) + "\n")
    names, losses, n_clusters = sp.load_clusters(str(p))
    assert names == [f"BC{i}-1" for i in range(n)] and n_clusters == 5 and losses.shape == (n, k)
    assert np.array_equal(losses, want, equal_nan=True) and np.signbit(losses[7, 3])
    bad = list(rows)
    bad[2500] = bad[2500].replace("\t", "\tx", 2)                    # cluster column "x..."
    bad[900] = bad[900] + "\t-1.0"                                   # one column too many
    bad[1700] = "lonely-barcode"
    p.write_text("\n".join(bad) + "\n")
    with pytest.raises(sp.SoupError) as e:
        sp.load_clusters(str(p))
    assert "line 901" in str(e.value), str(e.value)
    p.write_text("")                                                 # empty file: no cells, no columns
    names, losses, n_clusters = sp.load_clusters(str(p))
    assert names == [] and losses.shape[0] == 0
