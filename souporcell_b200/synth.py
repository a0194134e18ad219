"""Synthetic vartrix-shaped input generator.

Produces what `vartrix` would hand to the clustering binary: a pair of Matrix
Market coordinate files (`alt.mtx`, `ref.mtx`; rows = loci, columns = cells,
1-based, identical coordinate order in both files, explicit zeros kept) and a
`barcodes.tsv`, plus the ground-truth donor of every cell.

The reference ships no data and no generator; the shapes follow the
file contract read by `load_cell_data` (/root/reference/souporcell/src/main.rs:601-681).
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

BASE_SEED = 20240611

# (method, k, restarts, cells, loci, mean covered loci per cell, max_loci)
CONFIGS = {
    0: dict(method="em", k=3, restarts=4, n_cells=300, n_loci=2000, mean_loci=120, max_loci=2048),   # tiny test shape
    1: dict(method="em", k=4, restarts=10, n_cells=2000, n_loci=20000, mean_loci=500, max_loci=2048),
    2: dict(method="em", k=8, restarts=50, n_cells=10000, n_loci=50000, mean_loci=900, max_loci=2048),
    3: dict(method="khm", k=16, restarts=100, n_cells=40000, n_loci=100000, mean_loci=1000, max_loci=2048),
    4: dict(method="khm", k=64, restarts=100, n_cells=100000, n_loci=200000, mean_loci=1000, max_loci=2048),
    5: dict(method="em", k=32, restarts=100, n_cells=10000, n_loci=50000, mean_loci=900, max_loci=2048),
}


@dataclasses.dataclass
class Vartrix:
    """COO allele counts sorted by (locus, cell) — the order the mtx files are written in."""
    n_loci: int
    n_cells: int
    locus: np.ndarray   # int64 [nnz], 0-based
    cell: np.ndarray    # int64 [nnz], 0-based
    alt: np.ndarray     # uint32 [nnz]
    ref: np.ndarray     # uint32 [nnz]
    donor: np.ndarray   # int32 [n_cells] ground truth
    barcodes: list

    @property
    def nnz(self) -> int:
        return int(self.locus.shape[0])


def generate(n_cells: int, n_loci: int, k: int, mean_loci: float, max_loci: int = 2048, seed: int = BASE_SEED,
             doublet_frac: float = 0.0, **_unused) -> Vartrix:
    """`doublet_frac` > 0 turns that fraction of the cells into inter-donor doublets (genotype = mean of two donors; for the
    troublet tests).  They are drawn from a separate generator, so doublet_frac = 0 reproduces the historical data bit for bit."""
    rng = np.random.default_rng(seed)
    donor = rng.integers(0, k, size=n_cells).astype(np.int32)
    donor2 = donor.copy()
    if doublet_frac > 0 and k > 1:
        rng2 = np.random.default_rng(seed + 777)
        dbl = rng2.random(n_cells) < doublet_frac
        donor2[dbl] = (donor[dbl] + rng2.integers(1, k, size=int(dbl.sum()))) % k
    f = np.clip(rng.beta(0.5, 0.5, size=n_loci), 0.05, 0.95)
    geno = rng.binomial(2, f[None, :].repeat(k, axis=0)) / 2.0            # [k][L] in {0, .5, 1}
    w = np.arange(1, n_loci + 1, dtype=np.float64) ** -0.8
    rng.shuffle(w)
    cdf = np.cumsum(w / w.sum())
    target = np.clip(np.rint(rng.lognormal(np.log(mean_loci) - 0.18, 0.6, size=n_cells)), 20, max_loci)
    draws = rng.poisson(target)                                           # with-replacement draws per cell
    cell_of_draw = np.repeat(np.arange(n_cells, dtype=np.int64), draws)
    locus_of_draw = np.searchsorted(cdf, rng.random(cell_of_draw.shape[0]), side="right").astype(np.int64)
    np.minimum(locus_of_draw, n_loci - 1, out=locus_of_draw)
    key = cell_of_draw * n_loci + locus_of_draw                           # dedupe, sorted by (cell, locus): the values np.unique returns,
    key.sort()                                                            # without its hash pass (5x slower than the sort at 4e7 keys)
    if key.shape[0]:
        key = key[np.concatenate([[True], key[1:] != key[:-1]])]
    cell = key // n_loci
    locus = key % n_loci
    # cap at max_loci covered loci per cell (keep a random subset)
    counts = np.bincount(cell, minlength=n_cells)
    if counts.max() > max_loci:
        keep = np.ones(key.shape[0], dtype=bool)
        starts = np.concatenate([[0], np.cumsum(counts)])
        for c in np.nonzero(counts > max_loci)[0]:
            drop = rng.choice(counts[c], size=counts[c] - max_loci, replace=False)
            keep[starts[c] + drop] = False
        cell, locus = cell[keep], locus[keep]
    depth = rng.geometric(0.7, size=cell.shape[0])
    eps = 0.01
    g = geno[donor[cell], locus] if doublet_frac <= 0 else 0.5 * (geno[donor[cell], locus] + geno[donor2[cell], locus])
    p_alt = 0.95 * (g * (1 - 2 * eps) + eps) + 0.05 * f[locus]
    alt = rng.binomial(depth, p_alt).astype(np.uint32)
    ref = (depth - alt).astype(np.uint32)
    order = np.lexsort((cell, locus))                                     # mtx order: (locus, cell)
    letters = np.array(list("ACGT"))
    bc = letters[rng.integers(0, 4, size=(n_cells, 16))]
    barcodes = ["".join(row) + "-1" for row in bc]
    return Vartrix(n_loci, n_cells, locus[order], cell[order], alt[order], ref[order], donor, barcodes)


def generate_config(config: int, scale: float = 1.0) -> Vartrix:
    c = dict(CONFIGS[config])
    if scale != 1.0:
        c["n_cells"] = max(8, int(c["n_cells"] * scale))
        c["n_loci"] = max(64, int(c["n_loci"] * scale))
    return generate(seed=BASE_SEED + config, **c)


def write_files(v: Vartrix, outdir: str):
    """alt.mtx / ref.mtx / barcodes.tsv in the layout S:616-639 expects."""
    os.makedirs(outdir, exist_ok=True)
    hdr = "%%MatrixMarket matrix coordinate real general\n% written by souporcell_b200.synth\n"
    dims = f"{v.n_loci} {v.n_cells} {v.nnz}\n"
    if v.nnz > (1 << 20):          # large matrices: the library's parallel writer (same bytes; minutes of str() otherwise)
        from . import api
        for name, vals in (("alt.mtx", v.alt), ("ref.mtx", v.ref)):
            api.write_mtx_text(os.path.join(outdir, name), v.n_loci, v.n_cells, v.locus, v.cell, vals)
    for name, vals in (("alt.mtx", v.alt), ("ref.mtx", v.ref)) if v.nnz <= (1 << 20) else ():
        path = os.path.join(outdir, name)
        with open(path, "w") as fh:
            fh.write(hdr)
            fh.write(dims)
            block = 1 << 20
            for s in range(0, v.nnz, block):
                e = min(v.nnz, s + block)
                arr = np.stack([v.locus[s:e] + 1, v.cell[s:e] + 1, vals[s:e].astype(np.int64)], axis=1)
                fh.write("\n".join(" ".join(map(str, r)) for r in arr.tolist()))
                fh.write("\n")
    with open(os.path.join(outdir, "barcodes.tsv"), "w") as fh:
        fh.write("\n".join(v.barcodes) + "\n")
    return os.path.join(outdir, "alt.mtx"), os.path.join(outdir, "ref.mtx"), os.path.join(outdir, "barcodes.tsv")


def filter_to_csr(v: Vartrix, min_alt: int = 4, min_ref: int = 4, min_alt_umis: int = 0, min_ref_umis: int = 0):
    """numpy restatement of the locus filter + re-index of S:629-675 for duplicate-free input.

    Returns (row_ptr u64[N+1], locus u32[nnz'], alt u32, ref u32, index_to_locus i64[L'])."""
    L = v.n_loci
    cells_ref = np.bincount(v.locus, weights=(v.ref > 0), minlength=L)
    cells_alt = np.bincount(v.locus, weights=(v.alt > 0), minlength=L)
    umis_ref = np.bincount(v.locus, weights=v.ref, minlength=L)
    umis_alt = np.bincount(v.locus, weights=v.alt, minlength=L)
    seen = np.bincount(v.locus, minlength=L) > 0
    keep_locus = seen & (cells_ref >= min_ref) & (cells_alt >= min_alt) & (umis_ref >= min_ref_umis) & (umis_alt >= min_alt_umis)
    index_to_locus = np.nonzero(keep_locus)[0]
    remap = np.full(L, -1, dtype=np.int64)
    remap[index_to_locus] = np.arange(index_to_locus.shape[0])
    keep = keep_locus[v.locus] & ((v.alt + v.ref) > 0)
    cell, loc = v.cell[keep], remap[v.locus[keep]]
    alt, ref = v.alt[keep], v.ref[keep]
    order = np.lexsort((loc, cell))
    cell, loc, alt, ref = cell[order], loc[order], alt[order], ref[order]
    row_ptr = np.zeros(v.n_cells + 1, dtype=np.uint64)
    row_ptr[1:] = np.cumsum(np.bincount(cell, minlength=v.n_cells)).astype(np.uint64)
    return row_ptr, loc.astype(np.uint32), alt.astype(np.uint32), ref.astype(np.uint32), index_to_locus
