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


def _cell_worker(rank, world, port, out):
    """One rank of the cell-sharded iteration, on the CPU: this rank's cells' share of the M-step sums S:476-483 (float32, cells ascending,
    starting from 0 like mstep_kernel's partial mode) and of the loss S:225-230, ONE all-reduce of [S | D | loss] over gloo, then
    theta = clamp((1 + S) / (2 + D)) as theta_kernel does (S:386-396)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as orc
    from souporcell_b200 import api
    from helpers import synth_csr, random_theta
    v, csr = synth_csr(n_cells=260, n_loci=1800, k=3, mean_loci=100, seed=41)
    n, L, row_ptr, locus, alt, ref = csr
    K = 4
    theta = random_theta(np.random.default_rng(5), (K, L))
    res = {}
    for method in (0, 1):
        whole = orc.OracleData.from_csr(*csr).step(theta, method=method, temp_step=2)
        cuts, shards = api.split_cells(csr, world)
        a, b = cuts[rank], cuts[rank + 1]
        _, _, rp, lo, al, rf = shards[rank]
        S = np.zeros((K, L), dtype=np.float32)
        D = np.zeros((K, L), dtype=np.float32)
        for c in range(b - a):                                  # cells ascending, entries in locus order: the reference's order inside the shard
            p = whole["weights"][a + c]
            for j in range(int(rp[c]), int(rp[c + 1])):
                for k in range(K):
                    S[k, lo[j]] = np.float32(S[k, lo[j]] + np.float32(p[k] * np.float32(al[j])))
                    D[k, lo[j]] = np.float32(D[k, lo[j]] + np.float32(p[k] * np.float32(al[j] + rf[j])))
        loss = np.float32(0)
        for c in range(a, b):
            loss = np.float32(loss + whole["lse"][c])
        buf = torch.from_numpy(np.concatenate([S.ravel(), D.ravel(), [loss]]).astype(np.float32))
        dist.all_reduce(buf)                                    # the one collective of the iteration
        red = buf.numpy()
        Sr, Dr = red[:K * L].reshape(K, L), red[K * L:2 * K * L].reshape(K, L)
        th = np.clip((np.float32(1) + Sr) / (np.float32(2) + Dr), np.float32(0.01), np.float32(0.99))
        res[method] = (float(np.max(np.abs(th - whole["theta_out"]) / whole["theta_out"])), float(red[-1]), float(whole["loss"]), th.tobytes())
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_cell_sharded_iteration_allreduce_two_ranks_gloo():
    """The data-path collective of the cell-sharded mode over a real process group: per-rank partial M sums and
    losses, one all-reduce, every rank ends with the same theta, equal to the unsharded oracle iteration to north_star's 1e-5."""
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_cell_worker, args=(world, port, out), nprocs=world, join=True)
    for method in (0, 1):
        err0, loss0, want_loss, th0 = out[0][method]
        err1, loss1, _, th1 = out[1][method]
        assert th0 == th1 and loss0 == loss1                    # identical reduced numbers -> identical control decisions on every rank
        assert err0 < 1e-5, err0
        assert abs(loss0 - want_loss) <= 1e-5 * abs(want_loss)


def test_restart_shards_partition_the_solves():
    # solve s -> shard s mod world: disjoint, complete, and every stream's solves keep their own seed stream / word offset
    for n_solves, world in [(50, 1), (50, 2), (100, 8), (7, 8), (1000, 4)]:
        owned = [list(range(r, n_solves, world)) for r in range(world)]
        flat = sorted(s for o in owned for s in o)
        assert flat == list(range(n_solves))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_split_cells_partitions_the_matrix_by_nnz():
    # cell sharding: contiguous ranges, every cell and entry exactly once, nnz balanced to within one row
    sys.path.insert(0, ROOT)
    from souporcell_b200 import api, synth
    v = synth.generate(n_cells=400, n_loci=3000, k=3, mean_loci=150, seed=9)
    row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v)
    csr = (v.n_cells, int(i2l.shape[0]), row_ptr, locus, alt, ref)
    nnz, longest = int(row_ptr[-1]), int(np.diff(row_ptr.astype(np.int64)).max())
    for parts in (1, 2, 3, 8):
        cuts, shards = api.split_cells(csr, parts)
        assert cuts[0] == 0 and cuts[-1] == v.n_cells and all(a <= b for a, b in zip(cuts[:-1], cuts[1:])) and len(shards) == parts
        assert sum(s[0] for s in shards) == v.n_cells
        for name, col in (("locus", 3), ("alt", 4), ("ref", 5)):
            assert np.array_equal(np.concatenate([s[col] for s in shards]), csr[col]), name
        for (a, b), s in zip(zip(cuts[:-1], cuts[1:]), shards):
            assert s[1] == csr[1] and s[2][0] == 0 and s[2].dtype == np.uint64
            assert np.array_equal(s[2].astype(np.int64), row_ptr[a:b + 1].astype(np.int64) - int(row_ptr[a]))
            assert abs(int(s[2][-1]) - nnz / parts) <= 2 * longest
    # more parts than cells with data: empty shards are legal, nothing is lost
    tiny = (3, 5, np.array([0, 2, 2, 3], dtype=np.uint64), np.array([0, 4, 1], dtype=np.uint32), np.ones(3, np.uint32), np.zeros(3, np.uint32))
    cuts, shards = api.split_cells(tiny, 5)
    assert sum(s[0] for s in shards) == 3 and sum(int(s[2][-1]) for s in shards) == 3
    import pytest
    with pytest.raises(ValueError):
        api.split_cells(tiny, 0)
