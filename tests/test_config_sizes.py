"""Bit-exact parity at the sizes BASELINE.json names (reference loop main.rs:189-358, driver :60-148).

config 1 (the survey's parity config) is clustered COMPLETELY by the CLI and by the oracle CLI and compared byte for byte; configs 3, 4
and 5 are too large for a complete oracle job, so one solve from a shared theta0 is capped at 1-2 iterations and every bit of the
per-iteration loss, the log-probabilities and the centers is compared (the oracle's independent chains run on all host cores,
Params::inner_threads — bit-identical to its serial loop, tests/test_oracle_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import random_theta
from oracle import oracle_py as orc
import souporcell_b200 as sp

pytestmark = pytest.mark.gpu


def bits(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def test_config1_complete_job_cli_vs_oracle_cli(tmp_path):
    """BASELINE configs[0]: EM, k=4, 10 restarts, 2k cells x 20k loci, from the text files through both executables: the
    clusters TSV must be byte-identical and every `binomial` / `thread` / `best total` stderr line equal, for two thread counts
    (--threads is the seed fan-out S:63,78: 10 streams x 1 solve, and 4 streams x ceil(10/4)=3 solves = 12 >= restarts)."""
    from souporcell_b200 import build, synth
    v = synth.generate_config(1)
    alt, ref, bc = synth.write_files(v, str(tmp_path))
    pick = lambda s: sorted(l for l in s.splitlines() if l.startswith(("binomial", "total loci", "best total", "thread ")))
    for threads in ("10", "4"):
        args = ["-a", alt, "-r", ref, "-b", bc, "-k", "4", "--restarts", "10", "-t", threads]
        ours = subprocess.run([build.CLI] + args, capture_output=True, text=True)
        theirs = subprocess.run([orc.CLI_PATH] + args, capture_output=True, text=True)
        assert ours.returncode == 0 and theirs.returncode == 0, (ours.stderr[-500:], theirs.stderr[-500:])
        assert ours.stdout.count("\n") == v.n_cells
        assert ours.stdout == theirs.stdout
        po, pt = pick(ours.stderr), pick(theirs.stderr)
        assert len([l for l in po if l.startswith("binomial")]) >= 9 * (10 if threads == "10" else 12)
        assert po == pt


@pytest.mark.parametrize("config,cap", [(5, 2), (3, 2), (4, 1)])
def test_full_size_capped_solve_is_bit_exact(config, cap):
    """BASELINE configs[4] (k=32, config-2 shape), configs[2] (KHM k=16, 40k x 100k) and configs[3] (KHM k=64, 100k x 200k) at full
    matrix size: one em()/khm() call from a shared theta0, capped at `cap` iterations."""
    import bench
    cfg, csr = bench.workload(config)
    n, L, nnz = csr[0], csr[1], int(csr[3].shape[0])
    assert n == cfg["n_cells"] and 0.5 * cfg["n_loci"] < L <= cfg["n_loci"] and nnz > 500 * n
    method = sp.METHOD_KHM if cfg["method"] == "khm" else sp.METHOD_EM
    rng = np.random.default_rng(100 + config)
    th0 = random_theta(rng, (1, cfg["k"], L))
    with sp.Context(0) as ctx:
        ctx.load_csr(*csr)
        res = ctx.solve(cfg["k"], method=method, theta0=th0, iteration_cap=cap, keep_trace=True, record_counts=True)
        got_trace = ctx.get_trace(0)
    od = orc.OracleData.from_csr(*csr)
    want = od.solve(th0[0], method=method, iteration_cap=cap, inner_threads=os.cpu_count() or 1)
    assert want["iterations"] == cap and len(got_trace) == cap
    assert [(t[0], bits([t[1]])[0], t[3]) for t in got_trace] == [(t[0], bits([t[1]])[0], t[3]) for t in want["trace"]]
    assert np.float32(res.best_loss).tobytes() == np.float32(want["loss"]).tobytes()
    assert np.array_equal(bits(res.best_log_probs), bits(want["log_probs"]))
    assert np.array_equal(bits(res.best_theta), bits(want["theta"]))
