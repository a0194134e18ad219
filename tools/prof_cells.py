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
for ranges in os.environ.get("PROF_RANGES", "4,1,2,8").split(","):
    os.environ["SOUP_CELL_RANGES"] = ranges
    ctx = sp.Context(local, stream=stream.cuda_stream)
    ctx.load_csr(*part)
    comm = sp.Comm.from_torch(ctx, dist)
    ms = []
    for rep in range(3):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = sp.main(ctx, K, method=method, restarts=restarts, threads=restarts, seed=4, cell_comms=[comm], n_cells_global=n, time_kernels=(rep == 2))
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(float(t[0]))
    st = ctx.stats()
    out["jobs"].append({"ranges": int(ranges), "ms": ms, "iterations_enqueued": int(st["estep_launches"]), "allreduce_bytes": int(st["allreduce_bytes"]),
                        "E_ms_timed": r.estep_ms, "M_plus_comm_enqueue_ms_timed": r.mstep_ms, "best_loss": r.best_loss, "slot_iterations": r.slot_iterations})
    comm.close(); ctx.close()
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
