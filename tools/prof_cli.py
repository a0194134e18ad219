"""Developer probe: the drop-in CLI vs the oracle CLI on BASELINE config 2 written to vartrix files (wall time, TSV identity)."""
import sys, os, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from souporcell_b200 import synth, build
from oracle import oracle_py as orc
cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = synth.CONFIGS[cfgno]
d = tempfile.mkdtemp()
t0 = time.time(); v = synth.generate_config(cfgno); alt, ref, bc = synth.write_files(v, d)
print(f"files written in {time.time()-t0:.0f}s ({os.path.getsize(alt)/1e6:.0f} MB each)", flush=True)
threads = os.cpu_count()
args = ["-a", alt, "-r", ref, "-b", bc, "-k", str(c["k"]), "--restarts", str(c["restarts"]), "-t", str(threads), "-m", c["method"]]
t0 = time.time(); ours = subprocess.run([build.CLI] + args + ["--quiet_iterations"], capture_output=True, text=True); t_ours = time.time() - t0
print(f"soupgpu CLI: {t_ours:.2f}s wall, rc {ours.returncode}; {[l for l in ours.stderr.splitlines() if l.startswith(('soupgpu','total loci','best total'))]}", flush=True)
t0 = time.time(); ours2 = subprocess.run([build.CLI] + args + ["--quiet_iterations"], capture_output=True, text=True, env=dict(os.environ, SOUP_DEBUG_ITERS="1")); print(f"soupgpu CLI again: {time.time()-t0:.2f}s", [l for l in ours2.stderr.splitlines() if "cli" in l], flush=True)
if "--no-oracle" not in sys.argv:
    t0 = time.time(); theirs = subprocess.run([orc.CLI_PATH] + args, capture_output=True, text=True); t_ref = time.time() - t0
    print(f"oracle CLI ({threads} threads): {t_ref:.2f}s wall, rc {theirs.returncode}; identical TSV: {ours.stdout == theirs.stdout} ({ours.stdout.count(chr(10))} rows)")
    print(f"speed-up (process wall, incl. parsing the mtx text): {t_ref / t_ours:.1f}x")
