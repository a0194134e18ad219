"""souporcell_b200 — B200-native (sm_100a) replacement for the genotype-clustering hot path of
wheaton5/souporcell (the EM / k-harmonic-means loop of the `souporcell` Rust binary).

The product is `lib/libsoupgpu.so` (C ABI, include/soupgpu.h) and `bin/souporcell` (drop-in CLI);
this package is the thin ctypes host mirror used by the tests and the benchmark."""
from .api import (  # noqa: F401
    INIT_RANDOM_ASSIGNMENT, INIT_RANDOM_UNIFORM, METHOD_EM, METHOD_KHM, CellCSR, Comm, Context, MainResult, SolveResult, SoupError, bad_cluster_detection, chacha_words,
    format_f32, known_genotypes, lib, ln_binomial_f32, load_barcodes, load_clusters, load_mtx, load_mtx_cache, main, thread_seeds, troublet, write_clusters,
)
