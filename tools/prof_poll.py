"""Developer probe: what the speculative iterations cost.  The host enqueues chunks of `poll_chunk` iterations and always runs one chunk
ahead of the chunk it polls, so up to 2 x poll_chunk iterations of kernels that find no active slot trail every soup_solve.
Whole-job device time of BASELINE config 1 (4.9 ms job) and config 2 for poll_chunk = 1, 2, 4 (default), 8.
    python tools/prof_poll.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import souporcell_b200 as sp

for cfgno in (1, 2):
    cfg, csr = bench.workload(cfgno)
    for chunk in (1, 2, 4, 8):
        os.environ["SOUP_POLL_CHUNK"] = str(chunk)
        ctx = sp.Context(0).load_csr(*csr)
        ms, launches = [], 0
        for rep in range(6):
            r = sp.main(ctx, cfg["k"], restarts=cfg["restarts"], threads=cfg["restarts"], seed=4)
            ms.append(r.solve_ms); launches = r.kernel_launches
        ctx.close()
        print(f"config {cfgno} poll_chunk {chunk}: solve_ms min {min(ms[1:]):.3f} median {sorted(ms[1:])[2]:.3f}  kernel launches {launches}  slot-iterations {r.slot_iterations}", flush=True)
