import sys, os, time, threading
sys.path.insert(0, "/root/repo")
import numpy as np, bench, pynvml
import souporcell_b200 as sp
cfg, csr = bench.workload(2)
ctx = sp.Context(0).load_csr(*csr)
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def job(): return sp.main(ctx, 8, restarts=50, threads=50)
for _ in range(3): job()
def run(n, label, fn=None, period=0.25):
    stop = [False]; lat = []
    def loop():
        while not stop[0]:
            t = time.perf_counter(); fn(); lat.append((time.perf_counter() - t) * 1e3); time.sleep(period)
    th = threading.Thread(target=loop, daemon=True) if fn else None
    if th: th.start()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); job(); ts.append((time.perf_counter() - t) * 1e3)
    stop[0] = True
    print(f"{label:40s} steps ms: min {min(ts):.1f} med {np.median(ts):.1f} max {max(ts):.1f}; call ms: {np.round(lat[:6],2).tolist() if lat else '-'}")
import gc
gc.disable()
os.environ["SOUP_POLL_CHUNK"] = "4"; os.environ["SOUP_DEBUG_SLOW"] = "1"
rows = []
for _ in range(120):
    t = time.perf_counter(); r = sp.main(ctx, 8, restarts=50, threads=50, time_kernels=False); w = (time.perf_counter() - t) * 1e3
    rows.append((w, r.solve_ms))
rows = np.array(rows)
print("wall  : min %.1f med %.1f p90 %.1f max %.1f" % (rows[:,0].min(), np.median(rows[:,0]), np.percentile(rows[:,0],90), rows[:,0].max()))
print("solve : min %.1f med %.1f p90 %.1f max %.1f" % (rows[:,1].min(), np.median(rows[:,1]), np.percentile(rows[:,1],90), rows[:,1].max()))
print("outlier rows (wall, solve_ms):", np.round(rows[rows[:,0] > 46][:12], 1).tolist())
print("first 20 walls:", np.round(rows[:20,0],1).tolist())
