"""Compile-time experiments on the E / M kernels: build libsoupgpu variants (here, no GPU needed) and time them (on the GPU box).

    python tools/variants.py build NAME "-DSOUP_E_U8=8 -DSOUP_E_MINB=2" [NAME2 "..."]     # -> souporcell_b200/lib_variants/NAME/libsoupgpu.so
    python tools/variants.py run [NAME ...]                                                # per-iteration E / M times + a bit-exact check

`run` prints, per variant: the mean E and M time of the full-width iterations (2-12) of a config-2 job, the whole-job solve time, and
whether the bit-exact E/M parity tests pass with that library (SOUP_LIB selects it)."""
import os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "souporcell_b200", "lib_variants")


def build(name, defines):
    from souporcell_b200 import build as b
    out = os.path.join(VDIR, name)
    os.makedirs(out, exist_ok=True)
    flags = [f for f in b.NVCC_FLAGS] + defines.split()
    cmd = [b.NVCC] + flags + ["-Xptxas", "-v", "-shared", "-o", os.path.join(out, "libsoupgpu.so")] + [os.path.join(b.CSRC, s) for s in b.LIB_SOURCES] + b.LINK_LIBS
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        print(p.stdout[-3000:], p.stderr[-3000:])
        raise SystemExit(f"variant {name}: build failed")
    with open(os.path.join(out, "defines.txt"), "w") as fh:
        fh.write(defines + "\n")
    # registers / spills of the E and M kernels
    lines = p.stderr.splitlines()
    for i, l in enumerate(lines):
        m = re.search(r"Compiling entry function '_ZN4soup(12estep_kernel|12mstep_kernel)ILi8E", l)
        if m:
            info = " ".join(x.strip() for x in lines[i + 1:i + 4])
            regs = re.search(r"Used (\d+) registers", info)
            spill = re.search(r"(\d+) bytes spill stores", info)
            print(f"  {name}: {m.group(1)[2:]}<8> registers {regs.group(1) if regs else '?'} spill stores {spill.group(1) if spill else '?'}")


def run(names):
    if not names:
        names = sorted(os.listdir(VDIR))
    for name in names:
        lib = os.path.join(VDIR, name, "libsoupgpu.so")
        env = dict(os.environ, SOUP_LIB=lib, SOUP_DEBUG_ITERS="1")
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_iter.py")], capture_output=True, text=True, env=env)
        e, m, solve = [], [], []
        rep = 0
        for l in p.stderr.splitlines():
            g = re.match(r"soup_debug iter (\d+) E ([\d.]+) ms M ([\d.]+) ms", l)
            if g and rep == 1 and 2 <= int(g.group(1)) <= 12:
                e.append(float(g.group(2))); m.append(float(g.group(3)))
            g = re.match(r"solve_ms ([\d.]+)", l)
            if g:
                solve.append(float(g.group(1))); rep += 1
        t = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k",
                            "estep_mstep_bit_exact or solve_matches_oracle_per_iteration"], capture_output=True, text=True, env=dict(os.environ, SOUP_LIB=lib))
        ok = t.stdout.strip().splitlines()[-1] if t.stdout.strip() else t.stderr[-200:]
        d = open(os.path.join(VDIR, name, "defines.txt")).read().strip()
        if e:
            print(f"{name:14s} E {sum(e)/len(e):.3f} ms  M {sum(m)/len(m):.3f} ms  job {min(solve):.2f} ms   parity: {ok}   [{d}]", flush=True)
        else:
            print(f"{name:14s} FAILED rc {p.returncode}: {p.stderr[-400:]}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        a = sys.argv[2:]
        for i in range(0, len(a), 2):
            build(a[i], a[i + 1] if i + 1 < len(a) else "")
    else:
        run(sys.argv[2:])
