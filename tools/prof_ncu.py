"""Developer probe for ncu captures: a short config-2 job.  usage: prof_ncu.py [max_slots] [restarts] [iteration_cap]
(max_slots 5 at k = 8 reproduces the tail of a job: 40 live columns in every iteration)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import souporcell_b200 as sp
max_slots = int(sys.argv[1]) if len(sys.argv) > 1 else 0
restarts = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 12
cfg, csr = bench.workload(2)
ctx = sp.Context(0).load_csr(*csr)
r = sp.main(ctx, cfg["k"], restarts=restarts, threads=restarts, iteration_cap=cap, max_slots=max_slots, time_kernels=True)
print(f"solve_ms {r.solve_ms:.2f} E {r.estep_ms:.2f} M {r.mstep_ms:.2f} over {r.timed_iterations} iterations, slot_iters {r.slot_iterations}", file=sys.stderr)
