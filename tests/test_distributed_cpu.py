"""CPU, world_size 2 over gloo: the only cross-rank step of the restart-sharded path — agreeing on a run's winner and
fetching its log-probabilities / centers from the owning rank (soup_exchange_best, used once per run by soup_main).
Restarts themselves shard with no data-path collective (solve s runs on rank s mod world)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cases, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from souporcell_b200 import api
    ag, bc = api.torch_exchange_callbacks(dist)
    res = []
    for case in cases:
        loss, solve, has = case["cand"][rank]
        lp = np.full((5, 3), float(rank + 1), dtype=np.float32)
        th = np.full((3, 7), float(10 * (rank + 1)), dtype=np.float32)
        got = api.exchange_best(rank, world, ag, bc, loss, solve, has, lp, th)
        res.append((got, float(lp[0, 0]), float(th[0, 0])))
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_best_two_ranks_gloo():
    ninf = float("-inf")
    cases = [
        {"cand": [(-100.0, 4, True), (-50.0, 7, True)], "want": (-50.0, 7, True, 1)},      # higher loss wins
        {"cand": [(-50.0, 9, True), (-50.0, 2, True)], "want": (-50.0, 2, True, 1)},       # tie -> lowest solve index (S:110, S:120)
        {"cand": [(-50.0, 1, True), (ninf, -1, False)], "want": (-50.0, 1, True, 0)},      # a rank with nothing finite
        {"cand": [(ninf, -1, False), (ninf, -1, False)], "want": (ninf, -1, False, 0)},    # nobody beat -inf: buffers untouched
    ]
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, cases, out), nprocs=world, join=True)
    for rank in range(world):
        for case, (got, lp0, th0) in zip(cases, out[rank]):
            assert got == case["want"], (rank, got, case)
            if case["want"][2]:
                w = case["want"][3]
                assert lp0 == float(w + 1) and th0 == float(10 * (w + 1))          # every rank now holds the winner's buffers
            else:
                assert lp0 == float(rank + 1) and th0 == float(10 * (rank + 1))


def test_restart_shards_partition_the_solves():
    # solve s -> shard s mod world: disjoint, complete, and every stream's solves keep their own seed stream / word offset
    for n_solves, world in [(50, 1), (50, 2), (100, 8), (7, 8), (1000, 4)]:
        owned = [list(range(r, n_solves, world)) for r in range(world)]
        flat = sorted(s for o in owned for s in o)
        assert flat == list(range(n_solves))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
