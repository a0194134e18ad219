"""Developer probe: repeat one solve many times and report any run-to-run difference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from helpers import synth_csr
import souporcell_b200 as sp
if len(sys.argv) > 1:
    os.environ["SOUP_LONG_THRESHOLD"] = sys.argv[1]
v, csr = synth_csr()
ctx = sp.Context(0).load_csr(*csr)
seeds = sp.thread_seeds(4, 5, 0)
ref = {}
bad = 0
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 200):
    m, slots = rep % 2, [0, 3, 4][rep % 3]
    r = ctx.solve(3, method=m, n_streams=5, solves_per_stream=2, stream_seeds=seeds, max_slots=slots)
    sig = (r.best_solve, r.solve_loss.copy(), r.solve_iters.copy(), r.best_log_probs.copy(), r.best_theta.copy())
    if m in ref:
        a = ref[m]
        same = a[0] == sig[0] and all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(a[1:], sig[1:]))
        if not same:
            bad += 1
            print(f"rep {rep} method {m} slots {slots}: best {a[0]} vs {sig[0]}; loss diff at {np.nonzero(a[1].view(np.uint32) != sig[1].view(np.uint32))[0]}"
                  f" iters {a[2]} vs {sig[2]}; lp diff {int((a[3].view(np.uint32) != sig[3].view(np.uint32)).sum())} theta diff {int((a[4].view(np.uint32) != sig[4].view(np.uint32)).sum())}")
    else:
        ref[m] = sig
print("mismatching runs:", bad)
