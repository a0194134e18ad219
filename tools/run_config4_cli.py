"""North-star run, whole process: BASELINE config 4 (KHM, k=64, 100k cells x 200k loci, 100 restarts, souporcell3) through the drop-in
CLI from the vartrix TEXT files to the TSV, on all visible GPUs — wall time and stage times (SOUP_DEBUG_CLI), restart-sharded and
cell-sharded, against the single-GPU run of the same files.

    python tools/run_config4_cli.py [--gpus N] [--modes restarts,cells,single] [--restarts 100] [--out gpurun_out/config4_cli]

Also fills bench.py's matrix cache for config 4 (the filtered CSR), so a bench run after it does not generate the data again."""
import argparse, hashlib, json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from souporcell_b200 import synth, build
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--modes", default="restarts,cells,single")
ap.add_argument("--restarts", type=int, default=100)
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--dir", default="/tmp/soup_config4")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config4_cli"))
args = ap.parse_args()
build.build_all()
os.makedirs(args.out, exist_ok=True)
import torch
n_gpu = args.gpus or torch.cuda.device_count()
cfg = synth.CONFIGS[args.config]
rec = {"config": args.config, "gpus_visible": torch.cuda.device_count(), "host_cores": os.cpu_count(), "runs": {}}
alt, ref, bc = (os.path.join(args.dir, n) for n in ("alt.mtx", "ref.mtx", "barcodes.tsv"))
if not os.path.exists(bc):
    t0 = time.time()
    v = synth.generate_config(args.config)
    rec["generate_s"] = time.time() - t0
    t0 = time.time()
    synth.write_files(v, args.dir)
    rec["write_text_s"] = time.time() - t0
    np.save(os.path.join(args.dir, "donor.npy"), v.donor)          # ground truth of the synthetic cells
    cache = os.path.join(bench.CACHE_DIR, f"config{args.config}_seed{synth.BASE_SEED + args.config}_v2.npz")
    if not os.path.exists(cache):       # what bench.workload() would compute
        row_ptr, locus, a, r, i2l = synth.filter_to_csr(v)
        os.makedirs(bench.CACHE_DIR, exist_ok=True)
        np.savez(cache + ".tmp.npz", n_cells=np.int64(v.n_cells), n_loci=np.int64(i2l.shape[0]), row_ptr=row_ptr, locus=locus, alt=a, ref=r)
        os.replace(cache + ".tmp.npz", cache)
    del v
rec["text_bytes"] = [os.path.getsize(alt), os.path.getsize(ref)]
print(json.dumps({k: rec[k] for k in rec if k != "runs"}), flush=True)
base = ["-a", alt, "-r", ref, "-b", bc, "-k", str(cfg["k"]), "--restarts", str(args.restarts), "-t", str(args.restarts), "-m", cfg["method"],
        "-s", "true" if args.config == 4 else "false", "--quiet_iterations"]
plans = {"restarts": ["--gpus", str(n_gpu), "--shard", "restarts"], "cells": ["--gpus", str(n_gpu), "--shard", "cells"], "single": ["--gpus", "1"],
         "auto": [],                                      # the CLI picks the number of GPUs (sqrt rule) and the sharding itself
         "restarts4": ["--gpus", "4", "--shard", "restarts"], "restarts2": ["--gpus", "2", "--shard", "restarts"], "cells4": ["--gpus", "4", "--shard", "cells"]}
tsv = {}
for mode in args.modes.split(","):
    if mode not in plans or (mode != "single" and n_gpu < 2):
        continue
    for rep in range(args.reps if mode != "single" else 1):       # the second run reads the text from the page cache, like a pipeline stage that follows vartrix
        t0 = time.time()
        p = subprocess.run([build.CLI] + base + plans[mode], capture_output=True, text=True, env=dict(os.environ, SOUP_DEBUG_CLI="1"))
        wall = time.time() - t0
        stages = {l.split("cli", 1)[1].rsplit(None, 2)[0].strip(): float(l.split()[-2]) for l in p.stderr.splitlines() if l.startswith("soup_debug cli")}
        info = [l for l in p.stderr.splitlines() if l.startswith(("soupgpu", "total loci", "best total", "No outliers"))]
        rows = [l for l in p.stdout.splitlines() if "\t" in l]
        tsv[mode] = rows
        rec["runs"][f"{mode}_{rep}"] = {"argv": plans[mode], "wall_s": wall, "rc": p.returncode, "stages_ms": stages, "log": info, "rows": len(rows),
                                        "tsv_md5": hashlib.md5(p.stdout.encode()).hexdigest()}
        print(mode, rep, json.dumps(rec["runs"][f"{mode}_{rep}"]), flush=True)
        if p.returncode != 0:
            print(p.stderr[-2000:], flush=True)
def compare(a, b):
    """Two labelings of the same cells: agreement after the best one-to-one matching of the (arbitrary) labels, adjusted Rand index."""
    from scipy.optimize import linear_sum_assignment
    a, b = np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64)
    kk = int(max(a.max(), b.max())) + 1
    cont = np.zeros((kk, kk), dtype=np.int64)
    np.add.at(cont, (a, b), 1)
    ri, ci = linear_sum_assignment(-cont)
    comb = lambda x: x * (x - 1) / 2.0
    sum_ij, sa, sb, tot = comb(cont).sum(), comb(cont.sum(1)).sum(), comb(cont.sum(0)).sum(), comb(len(a))
    return {"agree_matched_labels": int(cont[ri, ci].sum()) / max(len(a), 1), "adjusted_rand_index": float((sum_ij - sa * sb / tot) / (0.5 * (sa + sb) - sa * sb / tot))}


truth = np.load(os.path.join(args.dir, "donor.npy")) if os.path.exists(os.path.join(args.dir, "donor.npy")) else None
labels = {mode: [int(r.split("\t")[1]) for r in rows] for mode, rows in tsv.items() if rows}
for mode, lab in labels.items():
    if truth is not None and len(lab) == len(truth):
        rec["runs"][f"{mode}_vs_ground_truth"] = compare(lab, truth)
        print(mode, "vs the synthetic donors:", rec["runs"][f"{mode}_vs_ground_truth"], flush=True)
ref_mode = "single" if "single" in labels else ("restarts" if "restarts" in labels else None)   # (restart sharding prints the single-GPU run's bytes)
for mode, lab in labels.items():
    if ref_mode and mode != ref_mode:
        rec["runs"][f"{mode}_vs_{ref_mode}"] = dict(compare(lab, labels[ref_mode]), tsv_identical=tsv[mode] == tsv[ref_mode])
        print(mode, "vs", ref_mode, rec["runs"][f"{mode}_vs_{ref_mode}"], flush=True)
with open(os.path.join(args.out, f"config{args.config}_gpus{n_gpu}.json"), "w") as fh:
    json.dump(rec, fh, indent=1)
