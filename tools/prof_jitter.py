"""Developer probe: host-side jitter of repeated config-2 jobs (SOUP_DEBUG_SLOW=1 prints where a slow call spent its time)."""
import sys, os, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
import souporcell_b200 as sp
cfg, csr = bench.workload(2)
ctx = sp.Context(0).load_csr(*csr)
for _ in range(3): sp.main(ctx, 8, restarts=50, threads=50)
gc.disable()
rows = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    t = time.perf_counter(); r = sp.main(ctx, 8, restarts=50, threads=50); w = (time.perf_counter() - t) * 1e3
    rows.append((w, r.solve_ms))
rows = np.array(rows)
print("wall  : min %.1f med %.1f p90 %.1f max %.1f" % (rows[:,0].min(), np.median(rows[:,0]), np.percentile(rows[:,0],90), rows[:,0].max()))
print("solve : min %.1f med %.1f p90 %.1f max %.1f" % (rows[:,1].min(), np.median(rows[:,1]), np.percentile(rows[:,1],90), rows[:,1].max()))
print("outliers (wall, solve_ms):", np.round(rows[rows[:,0] > np.median(rows[:,0]) + 3][:12], 1).tolist())
