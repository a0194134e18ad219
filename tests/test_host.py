"""CPU: the host side of libsoupgpu (no GPU needed) against the oracle and the golden fixtures; library surface."""
import ctypes
import gzip
import json
import os
import re
import subprocess

import numpy as np
import pytest

import souporcell_b200 as sp
from souporcell_b200 import api, build, synth
from oracle import oracle_py as orc
from test_oracle_golden import load, write_mtx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "soupgpu.h")).read()
    declared = set(re.findall(r"\b(soup_[a-z0-9_]+)\s*\(", hdr)) - {"soup_allgather_fn", "soup_bcast_fn"}
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    L = sp.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.soup_version() == 3000


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sp.SoupError) as e:
        sp.Context(0)
    assert e.value.code == api.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_chacha_and_seeds_match_public_vectors_and_oracle():
    kat = load("chacha_kat.json")
    key = bytes.fromhex(kat["key"])
    for rounds, hexs in kat["rounds"].items():
        assert sp.chacha_words(key, 16, 0, int(rounds)).astype("<u4").tobytes().hex() == hexs
    seed = bytes((7 * i + 3) % 256 for i in range(32))
    assert np.array_equal(sp.chacha_words(seed, 1000, 12345, 12), orc.chacha_words(seed, 1000, 12345, 12))
    for run in range(3):
        assert np.array_equal(sp.thread_seeds(4, 5, run), orc.thread_seeds(4, 5, run))
    assert np.array_equal(sp.thread_seeds(200, 1, 0), orc.thread_seeds(200, 1, 0))


def test_format_and_ln_binomial():
    for v, plain, w8 in load("format_f32.json"):
        assert sp.format_f32(v) == plain and sp.format_f32(v, True) == w8
    rng = np.random.default_rng(2)
    for v in np.concatenate([rng.standard_normal(300) * 1e3, rng.random(300), -rng.random(100) * 1e-5]).astype(np.float32):
        assert sp.format_f32(float(v)) == orc.format_f32(float(v))
        assert np.float32(float(sp.format_f32(float(v)))) == v            # round-trips
    for n, k in [(0, 0), (1, 0), (1, 1), (2, 1), (3, 2), (10, 3), (170, 85), (171, 3), (400, 200), (5, 6)]:
        assert np.float32(sp.ln_binomial_f32(n, k)) == np.float32(orc.ln_binomial(n, k)) or (k > n)
    # n > 170: statrs leaves its factorial table for the Lanczos ln_gamma (host_util.cpp / oracle.cpp restate it separately).
    # Library and oracle must agree bit for bit; math.lgamma is the independent check of the series (f64: 1e-12, hence a few f32 ulps).
    import math
    rng = np.random.default_rng(5)
    ns = np.concatenate([np.arange(171, 400), rng.integers(400, 70000, 300)])
    for n in ns.tolist():
        for k in {0, 1, n // 3, n // 2, n - 1, n}:
            got = np.float32(sp.ln_binomial_f32(n, k))
            assert got.view(np.uint32) == np.float32(orc.ln_binomial(n, k)).view(np.uint32), (n, k)
            want = math.lgamma(n + 1) - math.lgamma(k + 1) - math.lgamma(n - k + 1)
            assert orc.ln_binomial(n, k) == pytest.approx(want, rel=1e-12, abs=1e-8), (n, k)


def test_loader_golden_and_vs_oracle(tmp_path):
    g = load("loader_micro.json")
    alt, ref = write_mtx(tmp_path, g["n_loci"], g["n_cells"], g["entries"])
    m = sp.load_mtx(alt, ref, g["min_alt"], g["min_ref"])
    w = g["expect"]
    assert (m.n_cells, m.n_loci_total, m.n_loci) == (4, 5, 3)
    for k in ("row_ptr", "locus", "alt", "ref", "index_to_locus"):
        assert getattr(m, k).tolist() == w[k], k
    (tmp_path / "d").mkdir()
    alt, ref = write_mtx(tmp_path / "d", 1, 4, g["dup_entries"])
    w = g["dup_expect"]
    m = sp.load_mtx(alt, ref, w["min_alt"], w["min_ref"])
    for k in ("row_ptr", "locus", "alt", "ref", "index_to_locus"):
        assert getattr(m, k).tolist() == w[k], k


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("mins", [(4, 4, 0, 0), (1, 1, 0, 0), (2, 3, 5, 4), (10, 10, 0, 0)])
def test_loader_matches_oracle_on_synthetic_vartrix(tmp_path, shuffle, mins):
    v = synth.generate(n_cells=150, n_loci=900, k=3, mean_loci=70, seed=21)
    if shuffle:                                       # entry order must not matter (hash maps in the reference)
        p = np.random.default_rng(1).permutation(v.nnz)
        v.locus, v.cell, v.alt, v.ref = v.locus[p], v.cell[p], v.alt[p], v.ref[p]
    alt, ref, _ = synth.write_files(v, str(tmp_path))
    m = sp.load_mtx(alt, ref, *mins)
    ex = orc.OracleData.from_mtx(alt, ref, *mins).export()
    assert m.n_loci == ex["index_to_locus"].shape[0]
    for k in ("row_ptr", "locus", "alt", "ref", "index_to_locus"):
        assert np.array_equal(getattr(m, k), ex[k]), k


def test_loader_edge_cases(tmp_path):
    # empty body, header only
    alt, ref = write_mtx(tmp_path, 3, 2, [])
    m = sp.load_mtx(alt, ref)
    assert (m.n_cells, m.n_loci, m.nnz) == (2, 0, 0) and m.row_ptr.tolist() == [0, 0, 0]
    # out-of-range indices abort like the reference's assert!(locus < total_loci) (S:627-628)
    (tmp_path / "b").mkdir()
    alt, ref = write_mtx(tmp_path / "b", 3, 2, [[4, 1, 1, 1]])
    with pytest.raises(sp.SoupError):
        sp.load_mtx(alt, ref)
    with pytest.raises(sp.SoupError) as e:
        sp.load_mtx(str(tmp_path / "missing.mtx"), ref)
    assert e.value.code == api.ERR_IO
    # ragged: ref file shorter than alt -> the zip stops at the shorter one (S:616)
    (tmp_path / "r").mkdir()
    alt, ref = write_mtx(tmp_path / "r", 2, 2, [[1, 1, 1, 1], [1, 2, 1, 1], [2, 1, 1, 1]])
    lines = open(ref).read().splitlines()
    open(ref, "w").write("\n".join(lines[:-1]) + "\n")
    m = sp.load_mtx(alt, ref, 1, 1)
    ex = orc.OracleData.from_mtx(alt, ref, 1, 1).export()
    assert np.array_equal(m.locus, ex["locus"]) and np.array_equal(m.row_ptr, ex["row_ptr"])


def test_barcodes_plain_gz_crlf(tmp_path):
    names = ["AAAC-1", "CCGT-1", "", "TTTT-1"]
    p = tmp_path / "barcodes.tsv"
    p.write_text("\r\n".join(names) + "\r\n")
    assert sp.load_barcodes(str(p)) == names
    pz = tmp_path / "barcodes.tsv.gz"
    with gzip.open(pz, "wt") as fh:
        fh.write("\n".join(names))
    assert sp.load_barcodes(str(pz)) == names
    with pytest.raises(sp.SoupError):
        sp.load_barcodes(str(tmp_path / "nope.tsv"))


def test_bad_cluster_detection_matches_golden_and_oracle():
    for case in load("bad_clusters.json"):
        k = case["k"]
        rows = []
        for c, n in enumerate(case["sizes"]):
            r = np.full(k, -10.0, dtype=np.float32)
            r[c] = -1.0
            rows += [r] * n
        assert sp.bad_cluster_detection(case["run"], k, np.stack(rows)) == case["expect"]
    rng = np.random.default_rng(9)
    lp = -rng.random((4000, 32)).astype(np.float32) * 50
    lp[:500] = lp[0]                                   # ties and NaN handled by total_cmp like the oracle
    lp[7, 3] = np.nan
    for run in (1, 2):
        assert sp.bad_cluster_detection(run, 32, lp) == orc.bad_cluster_detection(run, 32, lp)


def test_write_clusters_tsv_format(tmp_path):
    lp = np.array([[-1.5, -0.25, -0.25], [np.float32(-1e-7), -3.0, -2.0], [-5.0, -5.0, -5.0]], dtype=np.float32)
    out = tmp_path / "clusters_tmp.tsv"
    sp.write_clusters(str(out), ["bc0", "bc1", "bc2", "extra-barcode"], lp)   # zip() truncates to the shorter (S:133)
    assert out.read_text() == "bc0\t1\t-1.5\t-0.25\t-0.25\nbc1\t0\t-0.0000001\t-3\t-2\nbc2\t0\t-5\t-5\t-5\n"


def test_cli_contract_without_gpu(tmp_path):
    """argv errors / -V / -h never need the GPU; a real run without one must fail loudly with nothing on stdout."""
    exe = build.CLI
    assert subprocess.run([exe, "-V"], capture_output=True, text=True).stdout == "souporcell 3.0\n"
    assert "--clustering_method" in subprocess.run([exe, "-h"], capture_output=True, text=True).stdout
    r = subprocess.run([exe, "-k", "3"], capture_output=True, text=True)
    assert r.returncode != 0 and r.stdout == "" and "required arguments" in r.stderr
    r = subprocess.run([exe, "--bogus", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "wasn't expected" in r.stderr
    g = load("loader_micro.json")
    alt, ref = write_mtx(tmp_path, g["n_loci"], g["n_cells"], g["entries"])
    bc = tmp_path / "bc.tsv"
    bc.write_text("a\nb\nc\nd\n")
    base = [exe, "-a", alt, "-r", ref, "-b", str(bc), "-k", "2", "--min_alt", "2", "--min_ref=2"]
    r = subprocess.run(base + ["-m", "kmeans"], capture_output=True, text=True)
    assert r.returncode == 101 and "Clustering method must be one of em, khm" in r.stderr and r.stdout == ""
    r = subprocess.run(base + ["--seed", "256"], capture_output=True, text=True)
    assert r.returncode == 101 and r.stdout == ""
    r = subprocess.run(base + ["--initialization_strategy", "kmeans++"], capture_output=True, text=True)
    assert r.returncode == 101 and "kmeans++ not yet implemented" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run(base, capture_output=True, text=True)
        assert r.returncode != 0 and r.stdout == "" and "total loci used 3" in r.stderr and "no CPU fallback" in r.stderr
        # the argv souporcell_pipeline.py:526-532 builds, in its order (values are strings there, `-s true`, `-m khm` lower-cased)
        pipeline = [exe, "-k", "2", "-a", alt, "-r", ref, "--restarts", "100", "-b", str(bc), "--min_ref", "2", "--min_alt", "2",
                    "--threads", "8", "-s", "true", "-m", "khm"]
        r = subprocess.run(pipeline, capture_output=True, text=True)
        assert r.returncode != 0 and r.stdout == "" and "total loci used 3" in r.stderr and "no CPU fallback" in r.stderr
        assert "wasn't expected" not in r.stderr and "required arguments" not in r.stderr


def write_vcf(path, n_records, samples, rng, gz=False):
    gts = ["0/0", "0/1", "1/1", "0|1", "1|0", "./.", "1/2", "2/2", ".|."]
    lines = ["##fileformat=VCFv4.2", "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">",
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples)]
    for i in range(n_records):
        fmt = "GT:DP" if i % 3 else "DP:GT"                       # GT is looked up by key, not by position
        cols = []
        for _ in samples:
            gt = gts[rng.integers(0, len(gts))]
            cols.append(f"{gt}:7" if i % 3 else f"7:{gt}")
        lines.append(f"chr1\t{100 + i}\t.\tA\tC,G\t50\tPASS\t.\t{fmt}\t" + "\t".join(cols))
    data = "\n".join(lines) + "\n"
    if gz:
        with gzip.open(path, "wt") as fh:
            fh.write(data)
    else:
        open(path, "w").write(data)


@pytest.mark.parametrize("gz", [False, True])
def test_known_genotypes_init_matches_oracle(tmp_path, gz):
    """init_cluster_centers_known_genotypes S:514-542: record i <-> original locus i, named sample s -> cluster s,
    (min(hap0,1)+min(hap1,1))/2 clamped to [.01,.99], '.' keeps 0.5, clusters without a sample stay 0.5."""
    v = synth.generate(n_cells=120, n_loci=400, k=3, mean_loci=60, seed=4)
    alt, ref, _ = synth.write_files(v, str(tmp_path))
    m = sp.load_mtx(alt, ref)
    od = orc.OracleData.from_mtx(alt, ref)
    vcf = str(tmp_path / ("known.vcf.gz" if gz else "known.vcf"))
    samples = ["donorB", "donorA", "donorC"]
    write_vcf(vcf, 400, samples, np.random.default_rng(5), gz)
    for names, k in ((["donorA", "donorC"], 4), (samples, 3)):
        got = sp.known_genotypes(vcf, names, k, m.index_to_locus)
        want = od.known_genotypes(vcf, names, k)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert set(np.unique(got)) <= {np.float32(0.01), np.float32(0.5), np.float32(0.99)}
        assert np.all(got[len(names):] == np.float32(0.5))
    for bad_names, k in ((["nobody"], 3), ([], 3), (samples, 2)):
        with pytest.raises(sp.SoupError):
            sp.known_genotypes(vcf, bad_names, k, m.index_to_locus)
    with pytest.raises(sp.SoupError):
        sp.known_genotypes(str(tmp_path / "missing.vcf"), samples, 3, m.index_to_locus)


def test_gz_input_and_binary_cache_roundtrip(tmp_path):
    """`.mtx.gz` inputs parse like the plain files, and a `.soupcsr` cache gives back exactly the parsed matrix — and
    refuses to be used for another filter, for OTHER alt / ref matrices than it was built from, or when damaged (`load_mtx(cache=)`
    and `souporcell --csr_cache` then rebuild it instead of printing results for the wrong data)."""
    import gzip
    import shutil
    v = synth.generate(n_cells=120, n_loci=900, k=3, mean_loci=60, seed=5)
    alt, ref, _ = synth.write_files(v, str(tmp_path))
    plain = sp.load_mtx(alt, ref)
    for src in (alt, ref):
        with open(src, "rb") as fi, gzip.open(src + ".gz", "wb") as fo:
            shutil.copyfileobj(fi, fo)
    gz = sp.load_mtx(alt + ".gz", ref + ".gz")
    fields = ("n_cells", "n_loci_total", "n_loci")
    same = lambda a, b: all(getattr(a, f) == getattr(b, f) for f in fields) and all(
        np.array_equal(getattr(a, f), getattr(b, f)) for f in ("row_ptr", "locus", "alt", "ref", "index_to_locus"))
    assert same(plain, gz)
    cache = str(tmp_path / "m.soupcsr")
    first = sp.load_mtx(alt, ref, cache=cache)                     # parses the text, writes the cache
    assert os.path.exists(cache) and same(plain, first)
    assert same(plain, sp.load_mtx_cache(cache, alt, ref))         # the cache itself
    assert same(plain, sp.load_mtx(alt, ref, cache=cache))
    with pytest.raises(sp.SoupError, match="another filter"):
        sp.load_mtx_cache(cache, alt, ref, min_alt=6)
    raw_cache = str(tmp_path / "raw.soupcsr")
    raw = sp.load_mtx(alt + ".gz", ref + ".gz", raw=True, cache=raw_cache)
    assert raw.nnz == v.nnz and same(raw, sp.load_mtx_cache(raw_cache, alt + ".gz", ref + ".gz", raw=True))
    with pytest.raises(sp.SoupError):
        sp.load_mtx_cache(raw_cache, alt + ".gz", ref + ".gz")     # RAW view asked for as the filtered one
    # a cache reused for other matrices (the advisor's reproduction): rejected, and load_mtx(cache=) returns the NEW data
    other = tmp_path / "other"
    v2 = synth.generate(n_cells=120, n_loci=900, k=3, mean_loci=60, seed=6)
    alt2, ref2, _ = synth.write_files(v2, str(other))
    plain2 = sp.load_mtx(alt2, ref2)
    assert not same(plain, plain2)
    with pytest.raises(sp.SoupError, match="built from another"):
        sp.load_mtx_cache(cache, alt2, ref2)
    with pytest.raises(sp.SoupError, match="built from another"):
        sp.load_mtx_cache(cache, alt, ref2)                        # only the ref matrix differs
    assert same(plain2, sp.load_mtx(alt2, ref2, cache=cache))      # rebuilt, not reused
    assert same(plain2, sp.load_mtx_cache(cache, alt2, ref2))
    with pytest.raises(sp.SoupError):
        sp.load_mtx_cache(cache, alt, ref)                         # ... and now it belongs to the second pair
    # a rewritten source file of the same size: the modification time / content hash differ
    data = open(alt2, "rb").read()
    os.utime(alt2, ns=(os.stat(alt2).st_atime_ns, os.stat(alt2).st_mtime_ns + 5_000_000_000))
    with pytest.raises(sp.SoupError, match="built from another"):
        sp.load_mtx_cache(cache, alt2, ref2)
    assert open(alt2, "rb").read() == data
    assert same(plain2, sp.load_mtx(alt2, ref2, cache=cache))
    blob = bytearray(open(cache, "rb").read())
    with open(cache, "wb") as fh:
        fh.write(blob[:-8])                                        # truncated
    with pytest.raises(sp.SoupError):
        sp.load_mtx_cache(cache, alt2, ref2)
    assert same(plain2, sp.load_mtx(alt2, ref2, cache=cache))


def test_header_is_plain_c_and_ctypes_mirrors_match_its_layout(tmp_path):
    """include/soupgpu.h is the drop-in boundary: it must compile as plain C (what bindgen / cgo / a Rust -sys crate consume), and every
    ctypes mirror in api.py must have the header's field names, order, offsets and sizes."""
    mirrors = {"soup_mtx": api._Mtx, "soup_solve_params": api._SolveParams, "soup_solve_result": api._SolveResult,
               "soup_iter_record": api._IterRecord, "soup_stats": api._Stats, "soup_main_params": api._MainParams,
               "soup_main_result": api._MainResult, "soup_troublet_params": api._TroubletParams, "soup_troublet_result": api._TroubletResult}
    hdr = open(os.path.join(ROOT, "include", "soupgpu.h")).read()
    declared = set(re.findall(r"^\}\s*(soup_[a-z0-9_]+)\s*;", hdr, flags=re.M))
    assert declared == set(mirrors), declared ^ set(mirrors)
    lines = ["#include <stddef.h>", "#include <stdio.h>", '#include "soupgpu.h"', "int main(void) {"]
    for name, cls in mirrors.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu %zu\\n", offsetof({name}, {field}), sizeof((({name}*)0)->{field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict((ln.split()[0], tuple(int(x) for x in ln.split()[1:])) for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in mirrors.items():
        assert got[name] == (ctypes.sizeof(cls),), (name, got[name], ctypes.sizeof(cls))
        body = re.search(r"typedef struct " + name + r"\s*\{(.*?)\}\s*" + name + r"\s*;", hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        n_members = sum(d.count(",") + 1 for d in body.split(";") if d.strip())
        assert n_members == len(cls._fields_), (name, n_members, len(cls._fields_))       # no header field without a mirror
        for field, _ in cls._fields_:
            d = getattr(cls, field)
            assert got[f"{name}.{field}"] == (d.offset, d.size), (name, field, got[f"{name}.{field}"], (d.offset, d.size))


def test_integration_md_rust_mirror_is_current():
    """INTEGRATION.md shows the -sys crate a maintainer would add: its #[repr(C)] structs must list the header's fields in order and
    its extern block may only name exported functions."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    mirrors = {"soup_mtx": api._Mtx, "soup_main_params": api._MainParams, "soup_main_result": api._MainResult}
    for name, cls in mirrors.items():
        body = re.search(r"#\[repr\(C\)\] pub struct " + name + r" \{(.*?)\n\}", doc, flags=re.S).group(1)
        body = re.sub(r"//[^\n]*", "", body)
        fields = [f.replace("r#", "") for f in re.findall(r"pub ((?:r#)?[a-z_0-9]+)\s*:", body)]
        assert fields == [f for f, _ in cls._fields_], (name, fields)
    fns = set(re.findall(r"\b(soup_[a-z0-9_]+)\(", doc))
    assert fns and fns <= set(api.EXPORTS), fns - set(api.EXPORTS)


@pytest.mark.parametrize("tool", ["souporcell", "troublet"])
def test_cli_accepts_every_flag_of_the_reference_params_yml(tool):
    """tests/golden/cli_flags.json is read from the reference's params.yml files (clap 2).  Every long / short flag must be known to
    the drop-in (in `--flag value`, `--flag=value` and `-f value` forms), -h must list them all, -V must print the reference's version,
    and leaving out any required flag must be the usage error.  argv errors come before any GPU work, so this runs without one."""
    spec = load("cli_flags.json")[tool]
    exe = build.CLI if tool == "souporcell" else build.TROUBLET_CLI
    run = lambda args: subprocess.run([exe] + args, capture_output=True, text=True)
    assert run(["-V"]).stdout == f"{spec['name']} {spec['version']}\n"
    help_text = run(["-h"]).stdout
    for a in spec["args"]:
        assert "--" + a["long"] in help_text, a["long"]
        if a["short"]:
            assert re.search(r"-" + a["short"] + r"\b", help_text), a["short"]
    required = [a for a in spec["args"] if a["required"]]
    assert required
    for a in spec["args"]:
        assert a["takes_value"]                                   # every flag of both tools takes a value
        forms = [["--" + a["long"], "x"], ["--" + a["long"] + "=x"]] + ([["-" + a["short"], "x"]] if a["short"] else [])
        for form in forms:
            r = run(form)                                         # the other required flags are missing: usage error, not "unknown flag"
            assert r.returncode != 0 and r.stdout == "", form
            if not (a["required"] and len(required) == 1):
                assert "required arguments" in r.stderr and "wasn't expected" not in r.stderr, (form, r.stderr)
    for skip in required:                                         # all required but one
        args = sum((["--" + a["long"], "/nonexistent"] for a in required if a is not skip), [])
        r = run(args)
        assert r.returncode != 0 and r.stdout == "" and "required arguments" in r.stderr and "--" + skip["long"] in r.stderr, (skip["long"], r.stderr)
    r = run(["--definitely_not_a_flag", "1"])
    assert r.returncode != 0 and "wasn't expected" in r.stderr


def test_troublet_help_is_the_text_the_reference_readme_shows():
    """tests/golden/readme_troublet_help.json = README.md:322-340 of the reference (clap 2 layout: flags sorted by long name, help column
    aligned).  Only the version line differs (the README's build printed 3.0, troublet/src/params.yml says 2.4)."""
    want = load("readme_troublet_help.json")["lines"]
    got = subprocess.run([build.TROUBLET_CLI, "-h"], capture_output=True, text=True).stdout.split("\n")
    assert got[-1] == "" and got[0] == "troublet 2.4" and want[0] == "troublet 3.0"
    assert [ln.rstrip() for ln in got[1:-1]] == [ln.rstrip() for ln in want[1:]]


def test_souporcell_help_is_the_text_the_reference_readme_shows():
    """tests/golden/readme_souporcell_help.json = README.md:253-297 of the reference: clap 2's layout on 120 columns (options sorted by
    long name; a help text that does not fit beside the longest option moves to the next line).  Two help strings lost their
    'NOT YET IMPLEMENTED ' prefix in params.yml since; the soupgpu extras follow in a section of their own."""
    want = [ln.replace("NOT YET IMPLEMENTED ", "").rstrip() for ln in load("readme_souporcell_help.json")["lines"]]
    got = [ln.rstrip() for ln in subprocess.run([build.CLI, "-h"], capture_output=True, text=True).stdout.split("\n")]
    cut = got.index("SOUPGPU OPTIONS:")
    assert got[:cut - 1] == want and got[cut - 1] == ""
    assert all("[soupgpu]" in ln for ln in got[cut + 1:] if ln)


def test_loader_multichunk_irregular_lines_any_thread_count(tmp_path):
    """A body long enough to split into several parse chunks and counting-sort parts (> 65536 lines each side of a cut), shuffled, with
    duplicates, and with lines the fast path must hand to the general tokenizer (tabs, padding, '+', CRLF, extra tokens, long zero-padded
    numbers, junk in the unparsed ref tokens): identical to the oracle's loader and identical for every thread count."""
    v = synth.generate(n_cells=1400, n_loci=5000, k=3, mean_loci=300, seed=33)
    rng = np.random.default_rng(4)
    p = rng.permutation(v.nnz)
    lo, ce, al, rf = (x[p] for x in (v.locus, v.cell, v.alt, v.ref))
    dup = rng.integers(0, v.nnz, 500)                              # duplicates of existing coordinates with other counts: last one wins
    lo, ce = np.concatenate([lo, lo[dup]]), np.concatenate([ce, ce[dup]])
    al, rf = np.concatenate([al, rng.integers(0, 4, 500).astype(al.dtype)]), np.concatenate([rf, rng.integers(0, 4, 500).astype(rf.dtype)])
    n = lo.shape[0]
    assert n > 4 * 65536, n
    forms = ["{l} {c} {x}\n", "{l}\t{c}\t{x}\n", "  {l}   {c}  {x}  \n", "+{l} +{c} +{x}\n", "{l} {c} {x}\r\n", "{l} {c} {x} trailing tokens 9\n",
             "0000000000{l} {c} 000000000000{x}\n", "{l} {c} {x}\t\r\n"]
    kind = np.where(rng.random(n) < 0.02, rng.integers(1, len(forms), n), 0)
    # ref_t.mtx: the same matrix without the junk tokens — troublet parses the ref line's locus and cell too (T:315-320)
    with open(tmp_path / "alt.mtx", "w", newline="") as fa, open(tmp_path / "ref.mtx", "w", newline="") as fr, open(tmp_path / "ref_t.mtx", "w", newline="") as ft:
        for fh in (fa, fr, ft):
            fh.write("%%MatrixMarket matrix coordinate real general\n% irregular\n")
            fh.write(f"{v.n_loci} {v.n_cells} {n}\n")
        for i in range(n):
            fa.write(forms[kind[i]].format(l=lo[i] + 1, c=ce[i] + 1, x=al[i]))
            ft.write(forms[kind[i]].format(l=lo[i] + 1, c=ce[i] + 1, x=rf[i]))
            if kind[i] == 5:
                fr.write(f"junk -7.5e3 {rf[i]}\n")                 # only the third ref token is ever parsed (S:625)
            else:
                fr.write(forms[kind[i]].format(l=lo[i] + 1, c=ce[i] + 1, x=rf[i]))
        fa.write(f"{lo[0] + 1} {ce[0] + 1} {al[0]}")               # no newline at the end of either file: still a line
        fr.write(f"{lo[0] + 1} {ce[0] + 1} {rf[0]}")
        ft.write(f"{lo[0] + 1} {ce[0] + 1} {rf[0]}")
    alt, ref, ref_t = str(tmp_path / "alt.mtx"), str(tmp_path / "ref.mtx"), str(tmp_path / "ref_t.mtx")
    ex = orc.OracleData.from_mtx(alt, ref, 3, 3).export()
    old = os.environ.get("SOUP_LOADER_THREADS")
    try:
        for threads in ("1", "3", "8", None):
            if threads is None:
                os.environ.pop("SOUP_LOADER_THREADS", None)
            else:
                os.environ["SOUP_LOADER_THREADS"] = threads
            m = sp.load_mtx(alt, ref, 3, 3)
            for k in ("row_ptr", "locus", "alt", "ref", "index_to_locus"):
                assert np.array_equal(getattr(m, k), ex[k]), (threads, k)
            raw = sp.load_mtx(alt, ref_t, raw=True)                # troublet's view: every line, file order inside a cell
            assert raw.nnz == n + 1 and raw.n_loci == v.n_loci
            order = np.argsort(np.concatenate([ce, ce[:1]]), kind="stable")
            assert np.array_equal(raw.locus, np.concatenate([lo, lo[:1]])[order].astype(raw.locus.dtype)), threads
            assert np.array_equal(raw.alt, np.concatenate([al, al[:1]])[order].astype(raw.alt.dtype)), threads
    finally:
        if old is None:
            os.environ.pop("SOUP_LOADER_THREADS", None)
        else:
            os.environ["SOUP_LOADER_THREADS"] = old
    with pytest.raises(sp.SoupError):                              # troublet's view unwraps the ref line's first two tokens as well
        sp.load_mtx(alt, ref, raw=True)
    # a malformed line in a late chunk is still the reference's panic
    bad = open(alt).read().split("\n")
    bad[len(bad) - 10] = "12 x 3"
    open(alt, "w").write("\n".join(bad))
    with pytest.raises(sp.SoupError):
        sp.load_mtx(alt, ref, 3, 3)


def test_write_clusters_many_blocks_matches_oracle_formatting(tmp_path):
    """The TSV writer formats blocks of cells in parallel and writes them in order: enough rows for several waves of blocks, ragged
    barcode lengths, special values; every token must be the oracle's rendering, the cluster column the FIRST maximum (S:134-141)."""
    n, k = 40000, 3
    rng = np.random.default_rng(12)
    lp = (-rng.random((n, k)) * 10.0 ** rng.integers(-6, 5, (n, k))).astype(np.float32)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-45, -1.17549435e-38, 3.4028235e38, -16777216.0, 0.1, -1e-7, 123456.79], dtype=np.float32)
    idx = rng.integers(0, n * k, 3000)
    lp.reshape(-1)[idx] = special[rng.integers(0, special.shape[0], 3000)]
    lp[17] = [-5.0, -5.0, -5.0]                                      # ties: strict '>' keeps the first
    lp[18] = [np.nan, np.nan, np.nan]                                # nothing beats -inf: cluster 0
    names = ["BC" + "ACGT" * int(rng.integers(0, 6)) + f"{i}-1" for i in range(n)]
    out = tmp_path / "big.tsv"
    sp.write_clusters(str(out), names, lp)
    rows = out.read_text().split("\n")
    assert rows[-1] == "" and len(rows) == n + 1
    fmt = {}
    for i in range(n):
        tok = rows[i].split("\t")
        assert tok[0] == names[i] and len(tok) == k + 2, i
        best, best_v = 0, -np.inf
        for c in range(k):
            if lp[i, c] > best_v:
                best, best_v = c, lp[i, c]
        assert tok[1] == str(best), i
        for c in range(k):
            v = float(lp[i, c])
            key = lp[i, c].tobytes()
            if key not in fmt:
                fmt[key] = orc.format_f32(v)
            assert tok[2 + c] == fmt[key], (i, c, v)


def test_sass_exactness_guards():
    """Static check of the built sm_100a code (cuobjdump runs without a GPU).  Bit-exactness against the reference's f32 arithmetic rests on
    the compiler never fusing a multiply into an add on its own (Rust never contracts; ptxas was seen to turn a packed multiply feeding a
    packed add into FFMA2 even under -fmad=false).  The E / M walks do issue fused operations themselves, but only through `vfma` with a
    multiplier that is 0 or a power of two (the product is then exact, so fused and unfused round alike); every other multiply must stay a
    scalar FMUL followed by a packed add.  So: FFMA2 may appear in the E / M kernels only, those kernels must still hold the unfused
    FMUL + FADD2 pairs of the general classes, and no other kernel may hold an FFMA2.  (Scalar FFMA inside the M / post kernels is the
    IEEE division and double -> float code, exact by construction; in the one-column E body it is `vfma` itself.)  The bit-exact GPU
    suite (counts 1..200 of either allele, test_gpu_parity.py) is the dynamic half of this guard.  The cubin must be sm_100a."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not installed")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB], capture_output=True, text=True, check=True).stdout
    assert "arch = sm_100a" in sass
    per = {}
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = {"FFMA": 0, "FFMA2": 0, "FADD2": 0, "FMUL": 0}
        elif cur:
            for op in ("FFMA2", "FADD2"):
                if re.search(r"\b" + op + r"\b", ln):
                    per[cur][op] += 1
            if re.search(r"\bFFMA\b", ln):
                per[cur]["FFMA"] += 1
            if re.search(r"\bFMUL\b", ln):
                per[cur]["FMUL"] += 1
    estep = {k: v for k, v in per.items() if "estep_kernel" in k}
    mstep = {k: v for k, v in per.items() if "mstep_kernel" in k}
    assert len(estep) == 5 and len(mstep) == 5, (sorted(estep), sorted(mstep))      # V = 1, 2, 4, 8, 16 columns per lane
    for k, v in estep.items():
        assert v["FMUL"] > 0, (k, v)
        assert v["FADD2"] > 0 or "ILi1E" in k, (k, v)                                # one column per lane has nothing to pair
        assert v["FFMA2"] > 0 or "ILi1E" in k, (k, v)
        if "ILi1E" not in k:
            assert v["FADD2"] > v["FFMA2"], (k, v)                                   # the general classes were not contracted away
    for k, v in per.items():
        if "estep_kernel" not in k and "mstep_kernel" not in k:
            assert v["FFMA2"] == 0, (k, v)
    assert all(v["FADD2"] > 0 for v in mstep.values())


def test_mtx_text_writer_matches_the_python_writer(tmp_path):
    """soup_mtx_write_text (parallel blocks, written in order) produces the bytes synth.write_files' plain-Python path writes — header,
    comment, dimensions, one `locus cell count` line per entry — over several waves of blocks, and the loader reads them back."""
    from souporcell_b200 import api
    rng = np.random.default_rng(8)
    n_loci, n_cells, nnz = 70000, 900, 700000                    # > 2 blocks of 2^18 entries
    locus = np.sort(rng.integers(0, n_loci, nnz)).astype(np.int64)
    cell = rng.integers(0, n_cells, nnz).astype(np.int64)
    cnt = rng.integers(0, 5000, nnz).astype(np.uint32)
    cnt[:10] = [0, 1, 9, 10, 99, 100, 4294967295, 12345678, 7, 70]
    path = str(tmp_path / "w.mtx")
    api.write_mtx_text(path, n_loci, n_cells, locus, cell, cnt, comment="written by souporcell_b200.synth")
    want = "%%MatrixMarket matrix coordinate real general\n% written by souporcell_b200.synth\n" + f"{n_loci} {n_cells} {nnz}\n" + \
           "".join(f"{l + 1} {c + 1} {x}\n" for l, c, x in zip(locus.tolist(), cell.tolist(), cnt.tolist()))
    assert open(path).read() == want
    empty = str(tmp_path / "e.mtx")
    api.write_mtx_text(empty, 3, 2, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint32))
    assert open(empty).read().splitlines()[-1] == "3 2 0"


def test_bench_workload_cache_round_trips(tmp_path, monkeypatch):
    """bench.py keeps the generated matrices in a scratch cache (config 4 takes numpy half a minute): the cached arrays are the
    generator's, a second call reads the file, and a half-written file can never be picked up (written under another name, then renamed)."""
    import importlib
    import bench
    monkeypatch.setattr(bench, "CACHE_DIR", str(tmp_path / "cache"))
    cfg, csr = bench.workload(0)
    v = synth.generate_config(0)
    row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v)
    assert cfg == synth.CONFIGS[0] and csr[0] == v.n_cells and csr[1] == int(i2l.shape[0])
    for got, want in zip(csr[2:], (row_ptr, locus, alt, ref)):
        assert got.dtype == want.dtype and np.array_equal(got, want)
    files = os.listdir(tmp_path / "cache")
    assert len(files) == 1 and files[0].endswith(".npz") and ".tmp" not in files[0]
    monkeypatch.setattr(synth, "generate_config", lambda *a, **k: (_ for _ in ()).throw(AssertionError("regenerated")))
    cfg2, csr2 = bench.workload(0)                                 # served from the file
    assert np.array_equal(csr2[3], csr[3])


def test_cli_without_a_usable_gpu_fails_loudly_in_every_process_model(tmp_path):
    """No CPU fallback: with no usable CUDA device the drop-in exits 101 with a message and an empty stdout — as one process, with
    `--gpus 2` (two devices named but not there), and as one process per GPU, where a rank that fails must take the others down with it
    instead of leaving them waiting in the shared-memory barrier."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    v = synth.generate(n_cells=60, n_loci=400, k=2, mean_loci=40, seed=3)
    a, r, b = synth.write_files(v, str(tmp_path))
    args = [build.CLI, "-a", a, "-r", r, "-b", b, "-k", "2", "--restarts", "4", "-t", "2"]
    for extra, env in (([], {}), (["--gpus", "2"], {"CUDA_VISIBLE_DEVICES": "0,1"}), (["--gpus", "2"], {"CUDA_VISIBLE_DEVICES": "0,1", "SOUP_CLI_PROCESSES": "1"}),
                       (["--gpus", "2", "--shard", "cells"], {"CUDA_VISIBLE_DEVICES": "0,1", "SOUP_CLI_PROCESSES": "1"})):
        p = subprocess.run(args + extra, capture_output=True, text=True, env=dict(os.environ, **env), timeout=60)
        assert p.returncode == 101 and p.stdout == "", (extra, env, p.returncode)
        assert "no CUDA device" in p.stderr and "no CPU fallback" in p.stderr
