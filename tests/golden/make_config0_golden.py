"""Generates tests/golden/config0_solves.json: COMPLETE em() / khm() solves of synth config 0 (300 cells x 2000 loci, k = 3, the four
restarts of `--seed 4 --threads 4`) from the independent pure-Python restatement of main.rs:189-483 in make_golden.py (float32
arithmetic, float64 log / exp rounded to f32) — iteration counts, temperature-step traces, final losses and hard assignments.
The initial centers are the reference's seeded draws (StdRng = ChaCha12, pinned by the public KAT vectors in chacha_kat.json; drawn here
through the oracle's init_uniform).  A few minutes of pure Python; run once, the JSON is committed.

    python tests/golden/make_config0_golden.py
"""
import json, os, sys, time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as mg
from oracle import oracle_py as orc
from souporcell_b200 import synth

cfg = synth.CONFIGS[0]
v = synth.generate_config(0)
row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v)
L, K, restarts = int(i2l.shape[0]), cfg["k"], cfg["restarts"]
cells = mg.cells_from({"row_ptr": row_ptr.astype(int).tolist(), "locus": locus.astype(int).tolist(), "alt": alt.astype(int).tolist(), "ref": ref.astype(int).tolist()})
seeds = orc.thread_seeds(4, restarts, 0)
out = {"config": 0, "k": K, "restarts": restarts, "seed": 4, "threads": restarts, "n_cells": v.n_cells, "loci_used": L, "nnz": int(locus.shape[0]),
       "tolerance": 2e-6, "methods": {}}
for method in ("em", "khm"):
    recs = []
    for t in range(restarts):
        theta0 = orc.init_uniform(bytes(seeds[t]), 0, K, L)          # S:554-563: thread t's first solve
        t0 = time.time()
        thf, lpf, lossf, iters, trace = mg.solve(cells, np.array(theta0, dtype=np.float32), method, L)
        lp = np.array(lpf, dtype=np.float32)
        recs.append({"thread": t, "iterations": iters, "loss": lossf, "temp_steps": [tr[0] for tr in trace], "losses": [tr[1] for tr in trace],
                     "assignments": np.argmax(lp, axis=1).astype(int).tolist(),
                     "margin_min": float(np.min(np.sort(lp, axis=1)[:, -1] - np.sort(lp, axis=1)[:, -2]))})
        print(method, t, iters, lossf, f"{time.time() - t0:.0f}s", flush=True)
    best = max(range(restarts), key=lambda i: (recs[i]["loss"], -i))
    out["methods"][method] = {"solves": recs, "best_thread": best}
with open(os.path.join(HERE, "config0_solves.json"), "w") as fh:
    json.dump(out, fh)
    fh.write("\n")
