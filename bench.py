#!/usr/bin/env python
"""Benchmark of the souporcell clustering hot path (EM / KHM restarts) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2]

One "step" is one complete clustering job of BASELINE.json's config (default configs[1]:
EM, k=8, 10k cells x 50k loci, 50 restarts per GPU) on synthetic vartrix-shaped data:
every restart runs its 9 annealing steps to convergence.  The metric is E+M nnz-updates/s
(1 update = one stored (cell, locus) entry carried through one E+M iteration of one solve).

  value   : whole-job nnz-updates/s with the CSR already resident in HBM (device time, CUDA events)
  e2e     : same metric through the public API from HOST buffers: CSR upload (+ CSC transpose on the
            GPU) + solve + best log-probs/centers read back, every step
  roofline: the dominant kernel's algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the CPU oracle (C++ restatement of the Rust reference, which cannot
            be built here: no cargo/rustc) on the host cores, bounded sample of the same workload.

N > 1 (torchrun, one rank per GPU): restarts shard across ranks with no data-path collective
(weak scaling: 50 restarts per GPU); ranks only exchange the winner once per run.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from souporcell_b200 import synth  # noqa: E402

METRIC = "E+M nnz-updates/s (clustering hot path)"
UNIT = "nnz-updates/s"


CACHE_DIR = os.environ.get("SOUP_BENCH_CACHE", "/tmp/soupgpu_bench_cache")


def workload(config: int):
    """The filtered CSR of BASELINE config `config` (seeded synthetic vartrix data).  Generating the larger configs takes numpy 20-80 s, so
    the arrays are kept in a scratch cache: under torchrun local rank 0 writes it and the other ranks wait for the file."""
    c = dict(synth.CONFIGS[config])
    path = os.path.join(CACHE_DIR, f"config{config}_seed{synth.BASE_SEED + config}_v2.npz")
    if not os.path.exists(path):
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            v = synth.generate_config(config)
            row_ptr, locus, alt, ref, i2l = synth.filter_to_csr(v)
            os.makedirs(CACHE_DIR, exist_ok=True)
            tmp = path + f".tmp{os.getpid()}.npz"
            np.savez(tmp, n_cells=np.int64(v.n_cells), n_loci=np.int64(i2l.shape[0]), row_ptr=row_ptr, locus=locus, alt=alt, ref=ref)
            os.replace(tmp, path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > 1800:
                    raise SystemExit("bench.py: timed out waiting for local rank 0 to generate the synthetic matrix")
                time.sleep(0.5)
    z = np.load(path)
    return c, (int(z["n_cells"]), int(z["n_loci"]), z["row_ptr"], z["locus"], z["alt"], z["ref"])


def algorithmic_bytes(nnz, K, L, N):
    """compulsory HBM bytes per solve-iteration, split by pass."""
    e = 8 * nnz + 8 * K * L + 4 * K * N + 12 * N
    m = 8 * nnz + 4 * K * N + 8 * K * L + 4 * L
    return e, m


def profiled_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as fh:
            return json.load(fh)[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def profiled_limits(kernel, avg_launch_ms):
    """What the committed ncu --set full capture says binds `kernel` (profiles/ncu_traffic.json): real DRAM GB/s and the share of the L2 -> SM
    crossbar cap, from a full-occupancy launch; dram_gbs_real_live = the same DRAM bytes over this run's own average launch time."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as fh:
            d = json.load(fh)[kernel]
    except Exception:
        return {}
    out = {k: d[k] for k in ("dram_gbs_real", "l2_xbar_tbs", "l2_xbar_frac", "issue_active_pct", "capture_launch_us") if k in d}
    if avg_launch_ms > 0 and "dram_bytes_per_launch" in d:
        out["dram_gbs_real_live_upper"] = d["dram_bytes_per_launch"] / (avg_launch_ms * 1e-3) / 1e9
    return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled in-process
    (a sampling thread, 4 Hz): spawning `nvidia-smi -lms 200` measurably stalls kernel launches on multi-GPU boxes,
    which would perturb the very number it is there to vouch for.  Falls back to nvidia-smi when pynvml is missing."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, index: int):
        self.index, self.sm, self.max_mhz, self.reasons, self._stop, self._thr, self.proc = index, [], None, set(), False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        get = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
                        mask = int(get(h))
                        for bit, name in self.REASONS.items():
                            if mask & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.25)
            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
        except Exception:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown"
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "500"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                threading.Thread(target=self._read, daemon=True).start()
            except Exception:
                self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [t.strip() for t in line.split(",")]
            try:
                self.sm.append(float(r[0])); self.max_mhz = float(r[1])
            except Exception:
                continue
            for name, val in zip(("sw_power_cap", "hw_slowdown", "sw_thermal_slowdown", "hw_thermal_slowdown"), r[2:6]):
                if val.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "samples": 0}
        busy = [x for x in self.sm if self.max_mhz and x > 0.3 * self.max_mhz] or self.sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---------------------------------------------------------------------------------- reference arm (CPU oracle)
def run_oracle_sample(csr, cfg, cores, budget_s):
    """Time the CPU oracle's souporcell_main on a bounded sample: `cores` restarts (one per thread),
    iteration cap chosen so the sample costs about `budget_s` seconds.  Returns (updates/s, desc, secs)."""
    from oracle import oracle_py as orc

    od = orc.OracleData.from_csr(*csr)
    method = 1 if cfg["method"] == "khm" else 0
    t0 = time.perf_counter()
    r = od.main(cfg["k"], method=method, restarts=cores, threads=cores, seed=4, iteration_cap=1)
    t1 = time.perf_counter() - t0                       # one iteration per thread (all threads in parallel)
    cap = int(max(1, min(1000, budget_s / max(t1, 1e-3))))
    if cap > 1:
        t0 = time.perf_counter()
        r = od.main(cfg["k"], method=method, restarts=cores, threads=cores, seed=4, iteration_cap=cap)
        t1 = time.perf_counter() - t0
    rate = r["nnz_updates"] / t1
    return rate, f"{cores} restarts (1 per thread) x iteration_cap {cap} of the same matrix, {r['nnz_updates']} nnz-updates in {t1:.1f}s", t1


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, csr = workload(args.config)
    cores = os.cpu_count() or 1
    budget = max(2.0, min(25.0, 150.0 / (args.steps + args.warmup)))
    vals, desc = [], ""
    for i in range(args.warmup + args.steps):
        rate, desc, secs = run_oracle_sample(csr, cfg, cores, budget)
        if i >= args.warmup:
            vals.append((rate, secs))
    value = float(np.mean([v for v, _ in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean([s for _, s in vals]) * 1e3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, cfg, csr, restarts_per_gpu=cfg["restarts"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = C++ restatement of souporcell/src/main.rs (oracle/); the Rust binary cannot be built in this image",
    }
    print(json.dumps(line), flush=True)


def config_dict(args, cfg, csr, restarts_per_gpu):
    n, L, row_ptr, locus, _, _ = csr
    return {"workload": f"BASELINE.json configs[{args.config - 1}]: {cfg['method'].upper()} k={cfg['k']}, {n} cells x {cfg['n_loci']} loci "
                        f"({L} after the min_alt/min_ref=4 filter), nnz={int(locus.shape[0])}, {restarts_per_gpu} restarts per GPU, synthetic vartrix (seed {synth.BASE_SEED + args.config})",
            "k": cfg["k"], "method": cfg["method"], "cells": n, "loci_used": L, "nnz": int(locus.shape[0]),
            "restarts_per_gpu": restarts_per_gpu, "sharding": "restarts (solve mod world), no data-path collective",
            "l2": "inputs larger than L2 (tables+weights of the resident restarts) and a 512 MB L2 flush between steps"}


# ---------------------------------------------------------------------------------- our arm
def ours(args):
    import torch
    import souporcell_b200 as sp
    from souporcell_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, csr = workload(args.config)
    n, L, row_ptr, locus, alt, ref = csr
    nnz, K = int(locus.shape[0]), cfg["k"]
    method = sp.METHOD_KHM if cfg["method"] == "khm" else sp.METHOD_EM
    restarts = cfg["restarts"] * world                       # weak scaling: fixed restarts per GPU
    stream = torch.cuda.Stream()                            # a real (non-NULL) stream: the library launches on it,
    torch.cuda.set_stream(stream)                           # so torch.cuda.Event brackets the library's kernels
    ctx = sp.Context(local, stream=stream.cuda_stream)
    ctx.load_csr(*csr)
    # one NCCL communicator inside libsoupgpu per rank: the once-per-run winner exchange of restart sharding and the per-iteration
    # all-reduce of cell sharding run over it (torch.distributed only carries its 128-byte id and the bench's own barriers)
    comm = sp.Comm.from_torch(ctx, dist) if world > 1 else None
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def job(time_kernels=False):
        return sp.main(ctx, K, method=method, restarts=restarts, threads=restarts, seed=4, time_kernels=time_kernels,
                       proc_rank=rank, proc_world=world, proc_comm=comm)

    def timed(fn):
        flush.zero_()
        gc.collect()          # collect between steps, never inside one (a gen-2 pass over this process's heap costs ~50 ms)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        out = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_dev, ms_wall = e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3
        t = torch.tensor([ms_dev, ms_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t[0]), float(t[1])

    def total(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t)
        return float(t[0])

    gc.disable()
    # ---- device-resident leg ("value")
    for _ in range(args.warmup):
        timed(job)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms = wall_ms = updates = launches = 0.0
    for _ in range(args.steps):
        r, ms_d, ms_w = timed(job)
        if os.environ.get("SOUP_BENCH_VERBOSE"):
            print(f"bench value step: wall {ms_w:.2f} ms device {ms_d:.2f} ms solve {r.solve_ms:.2f} ms", file=sys.stderr)
        dev_ms += ms_d; wall_ms += ms_w
        updates += total(r.nnz_updates); launches += total(r.kernel_launches)
    value = updates / (dev_ms * 1e-3)
    # same K steps again with every E / M launch bracketed by CUDA events on the launching stream (roofline leg).
    # Kept apart from the `value` steps: a timing event between two kernels stops the next launch from overlapping
    # the previous kernel's tail, which costs ~10 % of the step.
    em = [0.0, 0.0, 0, 0]
    ev_ms = 0.0
    for _ in range(0 if os.environ.get("SOUP_BENCH_NO_EVENTS") else args.steps):
        r, ms_d, _ = timed(lambda: job(time_kernels=True))
        ev_ms += ms_d
        em[0] += r.estep_ms; em[1] += r.mstep_ms; em[2] += r.timed_iterations; em[3] += r.slot_iterations
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end leg: host CSR (pinned) -> upload + transpose on the GPU -> solve -> results on the host, every step
    pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (row_ptr.view(np.int64), locus.view(np.int32), alt.view(np.int32), ref.view(np.int32))]
    csr_pinned = (n, L, pinned[0].numpy().view(np.uint64), pinned[1].numpy().view(np.uint32), pinned[2].numpy().view(np.uint32), pinned[3].numpy().view(np.uint32))

    def e2e_job():
        ctx.load_csr(*csr_pinned)
        return job()
    for _ in range(min(args.warmup, 1)):
        timed(e2e_job)
    e_ms = e_upd = 0.0
    for _ in range(args.steps):
        r, ms_d, ms_w = timed(e2e_job)
        if os.environ.get("SOUP_BENCH_VERBOSE"):
            print(f"bench e2e step: wall {ms_w:.2f} ms device {ms_d:.2f} ms", file=sys.stderr)
        e_ms += ms_w; e_upd += total(r.nnz_updates)       # wall clock: includes host-side packing and copies
    h2d = (n + 1) * 8 + nnz * 12 + restarts // world * 32
    d2h = n * K * 4 + K * L * 4

    # ---- sub-records: strong scaling of restart sharding (config 3, 100 restarts fixed) and the cell-sharded mode (config-4 shape)
    subs = {}
    if not args.no_sub:
        ctx.close()
        ctx = None
        del flush
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def sub_leg(config, restarts_total, mode, k=None, steps=2, warm=1):
            """One fixed job on the whole box.  mode 'restarts': solve s on rank s mod world, no data-path collective.
            mode 'cells': rank r holds the r-th contiguous cell range, every rank runs every restart, ncclAllReduce of the M-step sums."""
            scfg, scsr = workload(config)
            sn, sL, srp, slo, sal, srf = scsr
            sk = k or scfg["k"]
            smethod = sp.METHOD_KHM if scfg["method"] == "khm" else sp.METHOD_EM
            c2 = sp.Context(local, stream=stream.cuda_stream)
            cm = None
            extra = []                      # the second cell group's context and communicator (same GPU, its own streams)
            try:
                if mode == "cells" and world > 1:
                    part = api.split_cells(scsr, world)[1][rank]
                    c2.load_csr(*part)
                    cm = sp.Comm.from_torch(c2, dist)
                    groups = int(os.environ.get("SOUP_CELL_GROUPS", "2"))
                    ctxs_, comms_ = [c2], [cm]
                    for _ in range(groups - 1):
                        cx = sp.Context(local)
                        cx.load_csr(*part)
                        extra.append((cx, sp.Comm.from_torch(cx, dist)))
                        ctxs_.append(cx); comms_.append(extra[-1][1])
                    run = lambda: sp.main(ctxs_, sk, method=smethod, restarts=restarts_total, threads=restarts_total, seed=4, cell_comms=comms_,
                                          cell_groups=groups, n_cells_global=sn)
                else:
                    c2.load_csr(*scsr)
                    cm = sp.Comm.from_torch(c2, dist) if world > 1 else None
                    run = lambda: sp.main(c2, sk, method=smethod, restarts=restarts_total, threads=restarts_total, seed=4,
                                          proc_rank=rank, proc_world=world, proc_comm=cm)
                for _ in range(warm):
                    timed(run)
                ms = upd = 0.0
                r = None
                for _ in range(steps):
                    r, ms_d, _ = timed(run)
                    ms += ms_d
                    # cells: every rank counts its own entries (the sum over ranks is the whole matrix); restarts: its own solves
                    upd += total(r.nnz_updates)
                st = c2.stats()
                be_, bm_ = algorithmic_bytes(int(slo.shape[0]), sk, sL, sn)
                rec = {"value": upd / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warm, "scaling": "strong",
                       "n_gpus": world, "sharding": mode if world > 1 else "none (one GPU: the base of the strong-scaling curve)",
                       "workload": f"BASELINE.json configs[{config - 1}] matrix: {scfg['method'].upper()} k={sk}, {sn} cells x {sL} loci used, nnz={int(slo.shape[0])}, "
                                   f"{restarts_total} restarts in all (fixed as GPUs are added), every restart to convergence",
                       "best_loss": float(r.best_loss), "solve_iterations": int(total(r.slot_iterations) if mode != "cells" else r.slot_iterations),
                       "hbm_equiv_frac_per_gpu": upd / (ms * 1e-3) / world * (be_ + bm_) / int(slo.shape[0]) / 1e9 / measured_peak()[0]}
                if mode == "cells" and world > 1:
                    rec["cell_groups"] = 1 + len(extra)
                    rec["sharding"] = f"cells: every rank holds 1/{world} of the cells; the restarts run as {1 + len(extra)} cell groups on the same GPUs (one group's all-reduce overlaps the other's E / M passes)"
                    iters = max(int(st["estep_launches"]), 1)
                    rec["allreduce"] = {"bytes_per_rank_per_job": int(st["allreduce_bytes"]), "calls_per_job": int(st["allreduce_calls"]),
                                        "mean_bytes_per_iteration": int(st["allreduce_bytes"]) / iters,
                                        "full_width_bytes_per_iteration": (2 * sL * int(st["padded_cols"]) + int(st["slots"])) * 4,
                                        "note": "per cell group: one ncclAllReduce per locus range per iteration, overlapped with the M pass; only the live restarts' columns travel"}
                return rec
            finally:
                for cx, cc in extra:
                    cc.close(); cx.close()
                if cm is not None:
                    cm.close()
                c2.close()

        subs["strong"] = sub_leg(3, 100, "restarts")
        subs["cells"] = sub_leg(4, 8, "cells", k=64)

    if rank == 0:
        be, bm = algorithmic_bytes(nnz, K, L, n)
        peak, peak_src = measured_peak()
        kern, k_ms, k_bytes = ("estep_kernel", em[0], be) if em[0] >= em[1] else ("mstep_kernel", em[1], bm)
        per_launch_bytes = k_bytes * em[3] / max(em[2], 1)            # active solves per launch x bytes per solve-iteration
        achieved = per_launch_bytes / (k_ms / max(em[2], 1) * 1e-3) / 1e9 if k_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args, cfg, csr, cfg["restarts"]),
            "wall_ms_per_step": wall_ms / args.steps, "clocks": clocks,
            "e2e": {"value": e_upd / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e_ms / args.steps},
            "gpu_launches": int(launches),
            # bound: the roofline the contract's figures are taken against (algorithmic bytes over the measured HBM bandwidth).
            # limited_by: what MEASURABLY limits the kernel (ncu: L2 -> SM row gathers + the walks' dependent chains; DRAM is at a few percent).
            "roofline": {"bound": "hbm", "limited_by": "l2_gather+issue", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profiled_traffic(kern), "peak_source": peak_src, "bytes_per_launch": per_launch_bytes,
                         "limiter_note": "achieved/peak is an HBM-EQUIVALENT figure (algorithmic un-batched bytes / time).  Batched restarts share one pass over the "
                                         "matrix, so the real DRAM traffic (dram_gbs_real, ncu) is a few percent of HBM; what binds is the L2->SM row-gather path "
                                         "(l2_xbar_tbs; l2_xbar_frac = its share of the 21 TB/s a pure gather reaches on this GPU) and the walks' own "
                                         "dependent chains (issue_active_pct); figures from the committed ncu --set full capture, profiles/ncu_traffic.json",
                         **profiled_limits(kern, k_ms / max(em[2], 1)),
                         "avg_launch_ms": k_ms / max(em[2], 1), "estep_ms": em[0], "mstep_ms": em[1], "launches_timed": em[2],
                         "timing": f"per-launch CUDA events on the launching stream over {args.steps} extra steps of the same job ({ev_ms / args.steps:.2f} ms/step with events)",
                         "whole_step": {"bytes_per_nnz_update": (be + bm) / nnz, "achieved_per_gpu": value / world * (be + bm) / nnz / 1e9,
                                        "frac": value / world * (be + bm) / nnz / 1e9 / peak,
                                        "note": "whole job (all kernels, partial-occupancy tail included) per GPU vs the same peak"}},
        }
        line.update(subs)
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            rate, desc, _ = run_oracle_sample(csr, cfg, cores, 15.0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line), flush=True)
    if ctx is not None:
        ctx.close()
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ours_sweep(args):
    """BASELINE.json configs[4]: restart sweep 10-1000 restarts at k = 32 (restart-parallel, any number of ranks).  One JSON line whose
    `sweep` list holds, per restart count, the whole-job rate and time; `value` is the rate at the largest count."""
    import torch
    import souporcell_b200 as sp

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, csr = workload(args.config)
    n, L, row_ptr, locus, alt, ref = csr
    K = cfg["k"]
    method = sp.METHOD_KHM if cfg["method"] == "khm" else sp.METHOD_EM
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = sp.Context(local, stream=stream.cuda_stream)
    ctx.load_csr(*csr)
    comm = sp.Comm.from_torch(ctx, dist) if world > 1 else None
    out = []
    for restarts in (10, 30, 100, 300, 1000):
        def job():
            return sp.main(ctx, K, method=method, restarts=restarts, threads=restarts, seed=4, proc_rank=rank, proc_world=world, proc_comm=comm)
        ms = upd = 0.0
        reps = 1 if restarts >= 300 else 2
        for i in range(1 + reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r = job()
            e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1), float(r.nnz_updates)], dtype=torch.float64, device="cuda")
            if world > 1:
                tm = t.clone()
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                dist.all_reduce(t)
                t[0] = tm[0]
            if i > 0:
                ms += float(t[0]); upd += float(t[1])
        out.append({"restarts": restarts, "value": upd / (ms * 1e-3), "ms_per_job": ms / reps, "best_loss": float(r.best_loss)})
    if rank == 0:
        line = {"metric": METRIC, "value": out[-1]["value"], "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": out[-1]["ms_per_job"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config_dict(args, cfg, csr, "10..1000 in all"), sharding="restarts (solve mod world), no data-path collective"),
                "sweep": out}
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ours_cells(args):
    """Cell-sharded leg (optional: `--shard cells`, N > 1): every rank holds a contiguous range of the cells,
    runs all restarts, and all-reduces the M-pass statistics [2][L][columns] + losses over NCCL every iteration.
    Strong scaling: the job (config restarts on the whole matrix) is fixed, `value` = nnz-updates of the whole job / time."""
    import torch
    import torch.distributed as dist
    import souporcell_b200 as sp
    from souporcell_b200 import api

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, csr = workload(args.config)
    n, L, row_ptr, locus, alt, ref = csr
    shard = api.split_cells(csr, world)[1][rank]
    K, restarts = cfg["k"], cfg["restarts"]
    method = sp.METHOD_KHM if cfg["method"] == "khm" else sp.METHOD_EM
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = sp.Context(local, stream=stream.cuda_stream)
    ctx.load_csr(*shard)
    comm = sp.Comm.from_torch(ctx, dist)       # NCCL inside libsoupgpu; torch.distributed only carries its id
    seeds = sp.thread_seeds(4, restarts, 0)

    def job():
        return ctx.solve(K, method=method, n_streams=restarts, solves_per_stream=1, stream_seeds=seeds, cell_comm=comm, n_cells_global=n)

    ms = upd = launches = 0.0
    for i in range(args.warmup + args.steps):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = job()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if i >= args.warmup:
            ms += float(t[0]); upd += float(int(locus.shape[0]) * int(r.solve_iters.sum())); launches += r.stats["kernel_launches"]
    if rank == 0:
        be, bm = algorithmic_bytes(int(locus.shape[0]), K, L, n)
        peak, src = measured_peak()
        value = upd / (ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": dict(config_dict(args, cfg, csr, restarts), sharding="cells (contiguous ranges balanced by nnz) + NCCL all-reduce of the M statistics per iteration"),
                "gpu_launches": int(launches * world),
                "roofline": {"bound": "hbm", "achieved": value / world * (be + bm) / int(locus.shape[0]) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": value / world * (be + bm) / int(locus.shape[0]) / 1e9 / peak, "traffic": None, "peak_source": src,
                             "note": "whole job per GPU"},
                "allreduce_bytes_per_job_per_rank": int(r.stats["allreduce_bytes"]), "allreduce_calls_per_job": int(r.stats["allreduce_calls"]),
                "allreduce_full_width_floats_per_iteration": int(r.stats["padded_cols"]) * L * 2 + int(r.stats["slots"])}
        print(json.dumps(line), flush=True)
    comm.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config number (1-based)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="restart sweep 10..1000 (BASELINE configs[4]; use with --config 5)")
    ap.add_argument("--no-sub", action="store_true", help="skip the strong-scaling (config 3) and cell-sharded (config-4 shape) sub-records")
    ap.add_argument("--shard", default="restarts", choices=["restarts", "cells"], help="multi-GPU partition (default: restarts, weak scaling)")
    args = ap.parse_args()
    from souporcell_b200 import build
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        build.build_all()
    if args.impl == "reference":
        reference_arm(args)
    elif args.sweep:
        ours_sweep(args)
    elif args.shard == "cells" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        ours_cells(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
