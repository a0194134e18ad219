"""Developer probe: BASELINE config 4 (KHM, k=64, 100k cells x 200k loci, 100 restarts, souporcell3) restart-sharded over all
visible GPUs inside ONE process (soup_main with one context per GPU).  usage: prof_config4_multi.py [restarts] [s3: 0|1]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import souporcell_b200 as sp
n_gpu = torch.cuda.device_count()
t0 = time.time()
cfg, csr = bench.workload(4)
print(f"config 4 ready in {time.time()-t0:.0f}s: cells {csr[0]} loci_used {csr[1]} nnz {csr[3].shape[0]}; {n_gpu} GPUs", file=sys.stderr, flush=True)
t0 = time.time()
ctxs = [sp.Context(i) for i in range(n_gpu)]
for c in ctxs:
    c.load_csr(*csr)
print(f"{n_gpu} contexts + load_csr {time.time()-t0:.2f}s", file=sys.stderr, flush=True)
restarts = int(sys.argv[1]) if len(sys.argv) > 1 else 100
s3 = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
for rep in range(2):
    t0 = time.time()
    r = sp.main(ctxs, cfg["k"], method=sp.METHOD_KHM, restarts=restarts, threads=restarts, souporcell3=s3)
    dt = time.time() - t0
    print(f"rep {rep}: main {dt:.2f}s wall on {n_gpu} GPUs, solve_ms (slowest context, summed over runs) {r.solve_ms:.0f}, runs {r.runs}, "
          f"nnz_updates {r.nnz_updates:.3e}, {r.nnz_updates/dt/1e9:.1f} G/s wall, best {r.best_loss}", file=sys.stderr, flush=True)
