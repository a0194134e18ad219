"""Generates tests/golden/*.json.  Three kinds of fixtures:

  chacha_kat.json      PUBLIC known-answer vectors (draft-strombergson-chacha-test-vectors TC1: zero key, zero IV,
                       8 / 12 / 20 rounds; second block of the 20-round stream) — typed in, not computed.
  loader_micro.json    a 5-locus x 4-cell vartrix pair whose filtered CSR was worked out BY HAND from
                       load_cell_data (/root/reference/souporcell/src/main.rs:601-681) — typed in, not computed.
  format_f32.json      Rust `{}` / `{:8}` renderings of f32 values (known Rust behaviour) — typed in.
  bad_clusters.json    bad_cluster_detection cut-offs (main.rs:150-187) for K = 16 / 64 worked out by hand.
  troublet_micro.json  troublet's call_doublets (/root/reference/troublet/src/main.rs:32-244) on a 36-cell x 30-locus matrix with two planted
                       doublets, from an INDEPENDENT pure-Python restatement (this script; ln_gamma = math.lgamma, not the Lanczos
                       series the oracle restates from statrs): log-probabilities compared at 1e-9 relative, calls exactly.
  em_micro.json        one E+M pass and a full em()/khm() on the micro matrix from an INDEPENDENT numpy restatement
                       of main.rs:189-483 (this script; float32 arithmetic, float64 log/exp rounded to f32, so it can
                       differ from glibc logf/expf by an ulp: compared at 2e-6 relative, iteration counts exactly).

  cli_flags.json       the clap flag tables of both binaries, READ from souporcell/src/params.yml and troublet/src/params.yml.
  readme_souporcell_help.json  the `souporcell -h` text the reference's README.md:253-297 shows.
  readme_troublet_help.json    the `troublet -h` text the reference's README.md:322-340 shows.
  readme_troublet_sample.json  the ten sample rows of clusters.tsv in the reference's README.md:147-157 (its own troublet's output).

The reference itself has no tests or fixtures and cannot be built here (no Rust toolchain), so none of
these come from running it: PARITY UNPINNED.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as fh:
        json.dump(obj, fh, indent=1)
        fh.write("\n")


# ---------------------------------------------------------------- typed-in fixtures
CHACHA = {
    "source": "draft-strombergson-chacha-test-vectors-01, TC1 (all-zero 256-bit key, all-zero IV, block counter 0)",
    "key": "00" * 32,
    "rounds": {
        "8": "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42",
        "12": "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be",
        "20": "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586",
    },
    "rounds20_block1": "9f07e7be5551387a98ba977c732d080dcb0f29a048e3656912c6533e32ee7aed29b721769ce64e43d57133b074d839d531ed1f28510afb45ace10a1f4b794d6f",
}

# entries as (locus, cell, alt, ref), 1-based, deliberately NOT sorted (the loader must not depend on order)
LOADER = {
    "n_loci": 5, "n_cells": 4, "min_alt": 2, "min_ref": 2,
    "entries": [
        [4, 3, 1, 0], [1, 1, 1, 0], [1, 2, 0, 3], [1, 3, 2, 1],
        [2, 1, 1, 0], [2, 4, 1, 0],
        [4, 1, 0, 1], [4, 2, 0, 1], [4, 4, 1, 0],
        [5, 1, 2, 1], [5, 2, 0, 0], [5, 3, 1, 1],
    ],
    # by hand: locus 1 (ref cells 2,3 / alt cells 1,3) kept; locus 2 (no ref cell) dropped; locus 3 absent;
    # locus 4 (ref cells 1,2 / alt cells 3,4) kept; locus 5 (cells 1,3 both) kept, its explicit (0,0) entry dropped (main.rs:664)
    "expect": {
        "index_to_locus": [0, 3, 4],
        "row_ptr": [0, 3, 5, 8, 9],
        "locus": [0, 1, 2, 0, 1, 0, 1, 2, 1],
        "alt": [1, 0, 2, 0, 0, 2, 1, 1, 1],
        "ref": [0, 1, 1, 3, 1, 1, 0, 1, 0],
        "total_alleles": [5.0, 4.0, 6.0, 1.0],
        # ln C(n, k): C(1,*) = 1, C(3,2) = C(3,0)... = 3 or 1, C(2,1) = 2
        "lbc": ["0", "0", "ln3", "0", "0", "ln3", "0", "ln2", "0"],
    },
    # duplicates: the LAST line wins for the stored counts, every line counts in the per-locus tallies (main.rs:629-634)
    "dup_entries": [[1, 1, 1, 0], [1, 2, 0, 1], [1, 1, 0, 2], [1, 3, 3, 0]],
    "dup_expect": {"min_alt": 2, "min_ref": 2, "index_to_locus": [0], "row_ptr": [0, 1, 2, 3, 3], "locus": [0, 0, 0],
                   "alt": [0, 0, 3], "ref": [2, 1, 0]},
}

FORMAT = [
    [1.0, "1", "       1"], [0.1, "0.1", "     0.1"], [1e-7, "0.0000001", "0.0000001"], [1e10, "10000000000", "10000000000"],
    [-0.0, "-0", "      -0"], [float("inf"), "inf", "     inf"], [float("-inf"), "-inf", "    -inf"], [float("nan"), "NaN", "     NaN"],
    [16777216.0, "16777216", "16777216"], [0.3, "0.3", "     0.3"], [-2.5, "-2.5", "    -2.5"], [1234.5678, "1234.5677", "1234.5677"],
    [3.4028234663852886e38, "340282350000000000000000000000000000000", "340282350000000000000000000000000000000"],
    [-12345.678, "-12345.678", "-12345.678"], [10000.0, "10000", "   10000"], [0.5, "0.5", "     0.5"],
]

# cluster sizes chosen distinct so the stable sort has no ties except where stated
BAD = [
    {"k": 16, "run": 1, "sizes": [50, 3, 70, 9, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 5],
     # ascending by size: 1(3) 15(5) 3(9) ... 2(70): run_value 2 -> positions 0,1 and position 15
     "expect": [1, 15, 2]},
    {"k": 16, "run": 2, "sizes": [50, 3, 70, 9, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 5], "expect": [1]},
    {"k": 8, "run": 1, "sizes": [1, 2, 3, 4, 5, 6, 7, 8], "expect": []},
    {"k": 16, "run": 0, "sizes": [1] * 16, "expect": []},
    # ties: equal sizes keep cluster-index order (stable sort); run 1 -> positions 0,1 and 15
    {"k": 16, "run": 1, "sizes": [4] * 16, "expect": [0, 1, 15]},
    {"k": 64, "run": 1, "sizes": list(range(100, 164)), "expect": list(range(0, 8)) + list(range(57, 64))},
    {"k": 64, "run": 2, "sizes": list(range(100, 164)), "expect": list(range(0, 4)) + list(range(61, 64))},
]


# ---------------------------------------------------------------- independent numpy restatement of em()/khm()
def logf(x):
    return f32(np.log(np.float64(x)))


def expf(x):
    return f32(np.exp(np.float64(x)))


def ln_binomial(n, k):
    from math import lgamma
    return f32(lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1))


def lse(v):
    m = f32(-np.inf)
    for x in v:
        m = max(m, x)
    s = f32(0)
    for x in v:
        s = f32(s + expf(f32(x - m)))
    return f32(m + logf(s))


def cells_from(expect):
    rp, lo, al, rf = expect["row_ptr"], expect["locus"], expect["alt"], expect["ref"]
    cells = []
    for c in range(len(rp) - 1):
        ent = [(lo[j], al[j], rf[j], ln_binomial(al[j] + rf[j], al[j])) for j in range(rp[c], rp[c + 1])]
        tot = f32(0)
        for _, a, r, _ in ent:
            tot = f32(tot + f32(a + r))
        cells.append((ent, tot))
    return cells


def e_pass(cells, theta, method, temp_step):
    K = theta.shape[0]
    log_prior = logf(f32(1.0) / f32(K))
    lps, ws, loss = [], [], f32(0)
    for ent, tot in cells:
        lp = []
        for k in range(K):
            acc = log_prior
            for l, a, r, lbc in ent:
                t = f32(f32(lbc + f32(f32(a) * logf(theta[k, l]))) + f32(f32(r) * logf(f32(f32(1.0) - theta[k, l]))))
                acc = f32(acc + t)
            lp.append(acc)
        loss = f32(loss + lse(lp))
        if method == "em":
            temp = max(f32(tot / f32(20.0 * 2.0 ** temp_step)), f32(1))
            x = lp
        else:
            win = int(np.argmax(np.array(lp, dtype=np.float32)))     # first max
            wl = f32(-lp[win])
            d = [f32(f32(25.0) * f32(wl + v)) if i != win else f32(0) for i, v in enumerate(lp)]
            qd = lse(d)
            q = [f32(f32(f32(50.0) * wl - f32(f32(27.0) * f32(-v))) - f32(f32(2.0) * qd)) for v in lp]
            qs = lse(q)
            x = [f32(v - qs) for v in q]
            temp = max(f32(tot / f32(0.5 * 2.0 ** temp_step)), f32(1))
        if temp_step == 8:
            temp = f32(1)
        y = [f32(v / temp) for v in x]
        s = lse(y)
        ws.append([expf(f32(v - s)) for v in y])
        lps.append(lp)
    return lps, ws, loss


def m_pass(cells, ws, K, L, theta, locked=()):
    sums = np.ones((K, L), dtype=np.float32)
    den = np.full((K, L), 2.0, dtype=np.float32)
    for (ent, _), w in zip(cells, ws):
        for l, a, r, _ in ent:
            for k in range(K):
                sums[k, l] = f32(sums[k, l] + f32(w[k] * f32(a)))
                den[k, l] = f32(den[k, l] + f32(w[k] * f32(a + r)))
    out = theta.copy()
    for k in range(K):
        if k in locked:
            continue
        out[k] = np.maximum(np.minimum(sums[k] / den[k], f32(0.99)), f32(0.01))
    return out


def solve(cells, theta, method, L):
    K = theta.shape[0]
    limit = f32(0.01) * f32(len(cells))
    last, iters, trace, lps = f32(-np.inf), 0, [], None
    for t in range(9):
        change = f32(10000)
        while change > limit and iters < 1000:
            lps, ws, loss = e_pass(cells, theta, method, t)
            change = f32(loss - last)
            last = loss
            theta = m_pass(cells, ws, K, L, theta)
            iters += 1
            trace.append([t, float(loss)])
    return theta, lps, float(last), iters, trace


# ---------------------------------------------------------------- troublet: independent pure-Python restatement (T:32-244, T:261-412)
def troublet_micro():
    import math
    rng = np.random.default_rng(2024)
    n_cells, n_loci, K = 36, 30, 3
    geno = rng.integers(0, 3, size=(K, n_loci)) / 2.0
    donor = np.arange(n_cells) % K
    donor2 = donor.copy()
    donor2[[4, 17]] = (donor[[4, 17]] + 1) % K                      # two planted doublets
    entries = []                                                     # (locus, cell, ref, alt), file order = (locus, cell)
    for l in range(n_loci):
        for c in range(n_cells):
            if rng.random() < 0.8:
                depth = int(rng.integers(1, 5)) + (3 if c in (4, 17) else 0)
                g = 0.5 * (geno[donor[c], l] + geno[donor2[c], l])
                alt = int(rng.binomial(depth, 0.98 * g + 0.01))
                entries.append([l, c, depth - alt, alt])
    # clusters file: a confident assignment to the (first) donor
    losses = [[-5.0 if k == donor[c] else -60.0 - 3.0 * k for k in range(K)] for c in range(n_cells)]
    prior, thr_d, thr_s = 0.5, 0.9, 0.9
    lnb = lambda a, b: math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
    lnbin = lambda n, k: math.lgamma(n + 1) - math.lgamma(k + 1) - math.lgamma(n - k + 1)
    term = lambda r, a, f, cnt: lnbin(r + a, r) + lnb(r + f * cnt + 1.0, a + (1.0 - f) * cnt + 1.0) - lnb(f * cnt + 1.0, (1.0 - f) * cnt + 1.0)
    clamp = lambda f: min(max(f, 0.01), 0.99)
    lse = lambda xs: max(xs) + math.log(sum(math.exp(x - max(xs)) for x in xs))
    best_col = [max(range(K), key=lambda k: (losses[c][k], -k)) for c in range(n_cells)]
    S0 = np.zeros((K, n_loci, 2))
    lines = np.zeros(n_loci, dtype=int)
    cells = [[] for _ in range(n_cells)]
    for l, c, r, a in entries:
        S0[best_col[c], l] += (r, a)
        lines[l] += 1
        cells[c].append((l, r, a))
    A = S0.copy()
    soup = [S0[:, l, 0].sum() / S0[:, l].sum() if S0[:, l].sum() > 0 else 0.0 for l in range(n_loci)]   # p_soup = 0: unused
    removed, rounds, out = set(), 0, None
    while True:
        first = rounds == 0
        rounds += 1
        out, new = [], []
        for c in range(n_cells):
            sing = []
            for k in range(K):
                lp = math.log(1.0 - prior)
                for l, r, a in cells[c]:
                    if lines[l] < 20:
                        continue
                    t = A[k, l].sum()
                    f = 0.5 if t <= 0 else clamp(A[k, l, 0] / t)
                    lp += term(r, a, f, t)
                sing.append(lp)
            order = sorted(range(K), key=lambda k: (-sing[k], k))    # stable descending
            best_s = max(range(K), key=lambda k: (sing[k], -k))
            best_d, pair = -math.inf, (0, 0)
            top = order[:min(K, 4)]
            for i in range(len(top)):
                for j in range(i + 1, len(top)):
                    c1, c2 = top[i], top[j]
                    lp = math.log(prior)
                    for l, r, a in cells[c]:
                        if lines[l] < 20:
                            continue
                        t1, t2 = A[c1, l].sum(), A[c2, l].sum()
                        if first:
                            f = 0.5 if t1 <= 0 else (clamp(0.5 * A[c1, l, 0] / t1 + 0.5 * A[c2, l, 0] / t2) if t2 > 0 else clamp(A[c1, l, 0] / t1))
                        else:
                            rr, aa = A[c1, l, 0] + S0[c2, l, 0], A[c1, l, 1] + S0[c2, l, 1]
                            f = 0.5 if rr + aa <= 0 else clamp(rr / (rr + aa))
                        lp += term(r, a, f, (t1 + t2) / 2.0)
                    if lp > best_d:
                        best_d, pair = lp, (c1, c2)
            sp_ = math.exp(max(losses[c]) - lse(losses[c]))
            dp_ = math.exp(best_d - lse([sing[best_s], best_d]))
            if sing[best_s] > best_d:
                status = 0 if sp_ > thr_s else 1
            elif dp_ >= thr_d:
                status = 2
                if c not in removed:
                    removed.add(c)
                    new.append(c)
            else:
                status = 3
            out.append(dict(status=status, best_singlet=best_s, doublet=list(pair), singlet_posterior=sp_, doublet_posterior=dp_,
                            log_prob_singleton=sing[best_s], log_prob_doublet=best_d, cluster_logprobs=sing))
        for c in new:
            for l, r, a in cells[c]:
                A[best_col[c], l] -= (r, a)
        if not new:
            break
    return {"n_cells": n_cells, "n_loci": n_loci, "k": K, "entries": entries, "losses": losses, "doublet_prior": prior,
            "planted_doublets": [4, 17], "rounds": rounds, "removed": sorted(removed), "tolerance": 1e-9, "cells": out}


REFERENCE = "/root/reference"


def reference_derived():
    """Fixtures READ from the reference tree (not computed): the clap flag tables of both binaries and the sample clusters.tsv the
    README prints.  Only regenerated where /root/reference exists; the committed JSON files travel."""
    import yaml
    flags = {}
    for tool in ("souporcell", "troublet"):
        doc = yaml.safe_load(open(os.path.join(REFERENCE, tool, "src", "params.yml")))
        args = []
        for a in doc["args"]:
            (name, spec), = a.items()
            args.append({"name": name, "long": spec["long"], "short": spec.get("short"), "takes_value": bool(spec.get("takes_value", False)),
                         "required": bool(spec.get("required", False)), "multiple": bool(spec.get("multiple", False))})
        flags[tool] = {"source": f"{tool}/src/params.yml", "name": doc["name"], "version": str(doc["version"]), "args": args}
    dump("cli_flags.json", flags)
    lines = open(os.path.join(REFERENCE, "README.md")).read().split("\n")[146:157]
    rows = [ln.split() for ln in lines]
    assert rows[0][0] == "barcode" and len(rows) == 11 and all(len(r) == 7 for r in rows)
    dump("readme_troublet_sample.json", {
        "source": "README.md:147-157 of the reference: the sample clusters.tsv printed by its own troublet binary (an earlier column set: no "
                  "posterior columns). The only reference-produced numbers in the tree; whitespace-split because the README shows the tabs expanded.",
        "header": rows[0], "rows": rows[1:]})
    readme = open(os.path.join(REFERENCE, "README.md")).read().split("\n")
    assert readme[320] == "troublet -h" and readme[339].lstrip().startswith("--singlet_threshold")
    assert readme[251] == "souporcell -h" and readme[296].lstrip().startswith("-t, --threads")
    dump("readme_souporcell_help.json", {"source": "README.md:253-297 of the reference: what `souporcell -h` printed (clap 2.x on a 120-column "
                                                   "terminal; two help strings still said NOT YET IMPLEMENTED, params.yml:50,56 no longer do)",
                                         "lines": readme[252:297]})
    dump("readme_troublet_help.json", {"source": "README.md:322-340 of the reference: what `troublet -h` printed (clap 2.x; the README's build "
                                                 "called itself 3.0, troublet/src/params.yml says 2.4)", "lines": readme[321:340]})


def main():
    if os.path.isdir(REFERENCE):
        reference_derived()
    dump("troublet_micro.json", troublet_micro())
    dump("chacha_kat.json", CHACHA)
    dump("loader_micro.json", LOADER)
    dump("format_f32.json", FORMAT)
    dump("bad_clusters.json", BAD)
    cells = cells_from(LOADER["expect"])
    rng = np.random.default_rng(123)
    out = {"tolerance": 2e-6, "cases": []}
    for method in ("em", "khm"):
        for K in (2, 3):
            theta = np.clip(rng.random((K, 3)), 0.05, 0.95).astype(np.float32)
            lps, ws, loss = e_pass(cells, theta, method, 2)
            th1 = m_pass(cells, ws, K, 3, theta)
            thf, lpf, lossf, iters, trace = solve(cells, theta, method, 3)
            out["cases"].append({
                "method": method, "k": K, "theta0": theta.tolist(), "temp_step": 2,
                "log_probs": np.array(lps, dtype=np.float64).tolist(), "weights": np.array(ws, dtype=np.float64).tolist(),
                "loss": float(loss), "theta1": th1.astype(np.float64).tolist(),
                "solve": {"theta": thf.astype(np.float64).tolist(), "log_probs": np.array(lpf, dtype=np.float64).tolist(),
                          "loss": lossf, "iterations": iters, "trace": trace},
            })
    dump("em_micro.json", out)


if __name__ == "__main__":
    main()
