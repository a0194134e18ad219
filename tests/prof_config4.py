"""Developer probe: BASELINE config 4 (KHM, k=64, 100k cells x 200k loci, 100 restarts, souporcell3) on ONE B200, restart engine only."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import souporcell_b200 as sp
t0 = time.time()
cfg, csr = bench.workload(4)
print(f"generated config 4 in {time.time()-t0:.0f}s: cells {csr[0]} loci_used {csr[1]} nnz {csr[3].shape[0]}", file=sys.stderr, flush=True)
ctx = sp.Context(0)
t0 = time.time(); ctx.load_csr(*csr); print(f"load_csr {time.time()-t0:.2f}s", file=sys.stderr, flush=True)
restarts = int(sys.argv[1]) if len(sys.argv) > 1 else 100
t0 = time.time()
r = sp.main(ctx, cfg["k"], method=sp.METHOD_KHM, restarts=restarts, threads=restarts, souporcell3=True)
dt = time.time() - t0
from collections import Counter
print(f"main: {dt:.2f}s wall, solve_ms {r.solve_ms:.0f}, runs {r.runs}, nnz_updates {r.nnz_updates:.3e}, {r.nnz_updates/r.solve_ms/1e6:.1f} G/s, best {r.best_loss}", file=sys.stderr)
v = bench.synth.generate_config(4) if False else None
print("clusters used:", len(Counter(r.assignments.tolist())), file=sys.stderr)
