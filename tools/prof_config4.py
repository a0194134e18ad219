"""Developer probe: BASELINE config 4 (KHM, k=64, 100k cells x 200k loci, souporcell3) on ONE B200, restart engine only.
usage: prof_config4.py [restarts] [s3: 0|1]; caches the generated CSR in /tmp/soup_cfg4.npz"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import souporcell_b200 as sp
t0 = time.time()
cache = "/tmp/soup_cfg4.npz"
if os.path.exists(cache):
    z = np.load(cache)
    csr = (int(z["n"]), int(z["L"]), z["row_ptr"], z["locus"], z["alt"], z["ref"])
    cfg = dict(bench.synth.CONFIGS[4])
else:
    cfg, csr = bench.workload(4)
    np.savez(cache, n=csr[0], L=csr[1], row_ptr=csr[2], locus=csr[3], alt=csr[4], ref=csr[5])
print(f"config 4 ready in {time.time()-t0:.0f}s: cells {csr[0]} loci_used {csr[1]} nnz {csr[3].shape[0]}", file=sys.stderr, flush=True)
ctx = sp.Context(0)
t0 = time.time(); ctx.load_csr(*csr); print(f"load_csr {time.time()-t0:.2f}s", file=sys.stderr, flush=True)
restarts = int(sys.argv[1]) if len(sys.argv) > 1 else 100
s3 = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
t0 = time.time()
r = sp.main(ctx, cfg["k"], method=sp.METHOD_KHM, restarts=restarts, threads=restarts, souporcell3=s3, time_kernels=True)
dt = time.time() - t0
print(f"main: {dt:.2f}s wall, solve_ms {r.solve_ms:.0f} (E {r.estep_ms:.0f} M {r.mstep_ms:.0f} over {r.timed_iterations} iterations), runs {r.runs}, "
      f"nnz_updates {r.nnz_updates:.3e}, {r.nnz_updates/r.solve_ms/1e6:.1f} G/s, best {r.best_loss}", file=sys.stderr)
