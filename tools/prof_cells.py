"""Developer probe (torchrun, one rank per GPU): the cell-sharded job of bench.py's `cells` sub-record (config-4 matrix, KHM k = 64, 8 restarts)
for several settings of SOUP_CELL_RANGES, and the raw NCCL all-reduce rate at the message sizes that job uses.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/prof_cells.py [config] [restarts] [k]
"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
import souporcell_b200 as sp
from souporcell_b200 import api

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 4
restarts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg, csr = bench.workload(cfgno)
K = int(sys.argv[3]) if len(sys.argv) > 3 else cfg["k"]
n, L = csr[0], csr[1]
method = sp.METHOD_KHM if cfg["method"] == "khm" else sp.METHOD_EM
part = api.split_cells(csr, world)[1][rank]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
out = {"world": world, "config": cfgno, "restarts": restarts, "k": K, "cells": n, "loci": L, "jobs": [], "nccl_allreduce": []}
for combo in os.environ.get("PROF_COMBOS", "4:1,4:2,1:2,1:1,2:2").split(","):       # locus ranges : cell groups
    ranges, groups = combo.split(":")
    os.environ["SOUP_CELL_RANGES"] = ranges
    ctxs, comms = [], []
    for g in range(int(groups)):
        c = sp.Context(local, stream=stream.cuda_stream) if g == 0 else sp.Context(local)
        c.load_csr(*part)
        ctxs.append(c); comms.append(sp.Comm.from_torch(c, dist))
    ms = []
    for rep in range(3):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = sp.main(ctxs, K, method=method, restarts=restarts, threads=restarts, seed=4, cell_comms=comms, cell_groups=int(groups), n_cells_global=n)
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t[0]))
    st = ctxs[0].stats()
    out["jobs"].append({"ranges": int(ranges), "cell_groups": int(groups), "ms": ms, "iterations_enqueued_group0": int(st["estep_launches"]),
                        "allreduce_bytes_group0": int(st["allreduce_bytes"]), "best_loss": r.best_loss, "slot_iterations": r.slot_iterations})
    if rank == 0:
        print("#", json.dumps(out["jobs"][-1]), file=sys.stderr, flush=True)
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()
for mb in (45, 90, 180, 360, 720):
    x = torch.ones(mb * (1 << 20) // 4, dtype=torch.float32, device="cuda")
    for _ in range(2):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dist.all_reduce(x)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5
    out["nccl_allreduce"].append({"MB": mb, "ms": t, "algbw_GBs": mb * (1 << 20) / t / 1e6, "busbw_GBs": mb * (1 << 20) / t / 1e6 * 2 * (world - 1) / world})
    del x
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
