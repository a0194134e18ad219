"""CPU probe (oracle only, no GPU): how sparse are the M-step weights along the annealing schedule, and how many
iterations does a solve spend in each temperature step?  Decides whether an exact "skip the +0 weights" M pass is worth
building.  p = exp(y - lse(y)) underflows to exactly +0 once y_max - y > ~104, and s + 0*alt == s.

    python tools/prof_weight_sparsity.py            # a few minutes on 8 cores
"""
import collections
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py as orc            # noqa: E402  (test infrastructure)
from souporcell_b200 import synth              # noqa: E402

SHAPES = [(8, 2000, 50000, 900, 0), (16, 3000, 100000, 1000, 1), (64, 6000, 200000, 1000, 1)]   # k, cells, loci, mean loci, method


def main():
    for k, n_cells, n_loci, mean_loci, method in SHAPES:
        v = synth.generate(n_cells=n_cells, n_loci=n_loci, k=k, mean_loci=mean_loci, seed=5)
        row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v, 4, 4)
        L = int(i2l.shape[0])
        od = orc.OracleData.from_csr(v.n_cells, L, row_ptr, locus, alt, ref)
        print(f"k={k} cells={n_cells} loci_used={L} nnz={locus.shape[0]} method={'em' if method == 0 else 'khm'}")
        th = np.clip(np.random.default_rng(1).random((k, L), dtype=np.float32), 1e-4, 0.9999)
        for t in range(9):                      # two iterations per temperature step
            for _ in range(2):
                r = od.step(th, method=method, temp_step=t)
                th, w = r["theta_out"], r["weights"]
            nz = (w != 0).sum(1)
            print(f"  temp_step {t}: weights == +0 {np.mean(w == 0):.3f}, non-zero per cell {nz.mean():.2f}, one-hot cells {np.mean(nz == 1):.3f}")
        if k <= 16:                             # real solves: iterations per temperature step (S:215-250)
            cnt = collections.Counter()
            for s in range(6):
                th0 = np.clip(np.random.default_rng(s).random((k, L), dtype=np.float32), 1e-4, 0.9999)
                for t, *_ in od.solve(th0, method=method)["trace"]:
                    cnt[t] += 1
            print("  iterations per temp step over 6 solves:", dict(sorted(cnt.items())), "total", sum(cnt.values()))


if __name__ == "__main__":
    main()
