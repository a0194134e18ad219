"""Developer probe (not a test): per-iteration E / M times of one config-2 job.  usage: SOUP_DEBUG_ITERS=1 python tools/prof_iter.py [config] [max_slots] [lanes]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import souporcell_b200 as sp

cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 2
max_slots = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg, csr = bench.workload(cfgno)
ctx = sp.Context(0).load_csr(*csr)
method = sp.METHOD_KHM if cfg["method"] == "khm" else sp.METHOD_EM
for rep in range(2):
    r = sp.main(ctx, cfg["k"], method=method, restarts=cfg["restarts"], threads=cfg["restarts"], time_kernels=True, max_slots=max_slots, lanes=lanes)
    print("solve_ms", r.solve_ms, "E", r.estep_ms, "M", r.mstep_ms, "iters", r.timed_iterations, "slot_iters", r.slot_iterations,
          "Gupd/s", r.nnz_updates / r.solve_ms / 1e6, file=sys.stderr)
