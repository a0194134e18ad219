"""Shared test inputs: seeded synthetic vartrix data, filtered to the CSR both sides consume."""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from souporcell_b200 import synth  # noqa: E402


@functools.lru_cache(maxsize=8)
def synth_csr(n_cells=300, n_loci=2000, k=3, mean_loci=120, seed=7, max_loci=2048):
    v = synth.generate(n_cells=n_cells, n_loci=n_loci, k=k, mean_loci=mean_loci, max_loci=max_loci, seed=seed)
    row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v)
    return v, (v.n_cells, int(i2l.shape[0]), row_ptr, locus, alt, ref)


def random_theta(rng, shape):
    return np.clip(rng.random(shape, dtype=np.float32), 1e-4, 0.9999).astype(np.float32)
