"""Developer probe: phase timings of load_csr + main (SOUP_DEBUG_ITERS=1)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import souporcell_b200 as sp
cfg, csr = bench.workload(2)
ctx = sp.Context(0)
for rep in range(int(os.environ.get("REPS", "3"))):
    t0 = time.perf_counter(); ctx.load_csr(*csr); t1 = time.perf_counter()
    r = sp.main(ctx, cfg["k"], restarts=cfg["restarts"], threads=cfg["restarts"]); t2 = time.perf_counter()
    print(f"rep {rep}: load {1e3*(t1-t0):.1f} ms main {1e3*(t2-t1):.1f} ms solve_ms {r.solve_ms:.1f}", file=sys.stderr)
