"""Developer probe: troublet at config-2 size (10k cells x 50k loci, k = 8, 10 % synthetic doublets): GPU path vs the CPU oracle.
usage: prof_troublet.py [cells] [loci] [k]"""
import sys, os, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import souporcell_b200 as sp
from souporcell_b200 import build, synth
from oracle import oracle_py as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d = tempfile.mkdtemp()
t0 = time.time()
v = synth.generate(n_cells=n, n_loci=L, k=k, mean_loci=900, seed=77, doublet_frac=0.10)
alt, ref, bc = synth.write_files(v, d)
print(f"generated + wrote {v.nnz} entries in {time.time()-t0:.0f}s", file=sys.stderr, flush=True)
# cluster with the GPU CLI
t0 = time.time()
r = subprocess.run([build.CLI, "-a", alt, "-r", ref, "-b", bc, "-k", str(k), "--restarts", "20", "-t", "20", "--quiet_iterations"], capture_output=True, text=True)
assert r.returncode == 0, r.stderr[-300:]
clusters = os.path.join(d, "clusters_tmp.tsv")
open(clusters, "w").write(r.stdout)
print(f"souporcell CLI {time.time()-t0:.2f}s", file=sys.stderr, flush=True)
t0 = time.time()
names, losses, nc = sp.load_clusters(clusters)
raw = sp.load_mtx(alt, ref, raw=True)
t_load = time.time() - t0
with sp.Context(0) as ctx:
    for rep in range(3):
        t0 = time.time()
        got = sp.troublet(ctx, raw, losses, n_clusters=nc, barcodes=names)
        dt = time.time() - t0
        print(f"GPU troublet rep {rep}: {dt*1e3:.1f} ms wall ({got['device_ms']:.1f} ms in the singlet/doublet kernels), rounds {got['rounds']}, removed {got['removed']}; "
              f"loaders {t_load*1e3:.0f} ms", file=sys.stderr, flush=True)
t0 = time.time()
ours = subprocess.run([build.TROUBLET_CLI, "--alts", alt, "--refs", ref, "--clusters", clusters], capture_output=True, text=True)
t_cli = time.time() - t0
t0 = time.time()
theirs = subprocess.run([os.path.join(os.path.dirname(orc.CLI_PATH), "troublet_oracle"), "--alts", alt, "--refs", ref, "--clusters", clusters], capture_output=True, text=True)
t_orc = time.time() - t0
a, b = ours.stdout.splitlines(), theirs.stdout.splitlines()
same = sum(x.split("\t")[1:3] == y.split("\t")[1:3] for x, y in zip(a[1:], b[1:]))
from collections import Counter
print(f"troublet CLI (GPU) {t_cli:.2f}s vs oracle CLI (1 thread, like the reference) {t_orc:.2f}s; identical calls {same}/{len(a)-1}; "
      f"calls {dict(Counter(x.split(chr(9))[1] for x in a[1:]))}", file=sys.stderr)
