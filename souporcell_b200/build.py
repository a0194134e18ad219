"""In-tree build of libsoupgpu (sm_100a CUDA + C++ host), the `souporcell` CLI and the CPU oracle.

Everything is compiled with explicit nvcc / g++ command lines (no JIT cache) so the built
artefacts live in the tree and travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "souporcell_b200", "csrc")
LIB_DIR = os.path.join(ROOT, "souporcell_b200", "lib")
BIN_DIR = os.path.join(ROOT, "souporcell_b200", "bin")
LIB = os.path.join(LIB_DIR, "libsoupgpu.so")
CLI = os.path.join(BIN_DIR, "souporcell")
TROUBLET_CLI = os.path.join(BIN_DIR, "troublet")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
ORACLE_CLI = os.path.join(ORACLE_DIR, "_build", "souporcell_oracle")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
# Exact IEEE arithmetic everywhere: no FMA contraction (Rust never contracts), IEEE div, no FTZ.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-pthread",
]
# Compile-time experiments (e.g. SOUP_NVCC_DEFINES="-DSOUP_E_U8=4 -DSOUP_E_MINB=3" python -m souporcell_b200.build --force); empty = the measured build.
NVCC_FLAGS += os.environ.get("SOUP_NVCC_DEFINES", "").split()
LINK_LIBS = ["-lz", "-lnccl"]
LIB_SOURCES = ["soup_device.cu", "host_loader.cpp", "host_util.cpp", "soup_driver.cpp"]
LIB_HEADERS = ["soup_kernels.cuh", "troublet_kernels.cuh", "soup_common.hpp", "host_parallel.hpp", "glibc_flt32.h", os.path.join("..", "..", "include", "soupgpu.h")]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, cwd=None):
    proc = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("build command failed: " + " ".join(cmd))
    return proc.stdout


FLAGS_STAMP = LIB + ".flags"


def _flags_changed() -> bool:
    """The effective nvcc command line is part of the staleness check: a library built with SOUP_NVCC_DEFINES (an experiment) must not
    survive into a plain build, test or bench run, and changing the defines must take effect without --force."""
    try:
        with open(FLAGS_STAMP) as fh:
            return fh.read() != " ".join(NVCC_FLAGS)
    except OSError:
        return True


def build_lib(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in LIB_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in LIB_HEADERS]
    if force or _newer(LIB, deps) or _flags_changed():
        os.makedirs(LIB_DIR, exist_ok=True)
        _run([NVCC] + NVCC_FLAGS + ["-shared", "-o", LIB] + srcs + LINK_LIBS)
        with open(FLAGS_STAMP, "w") as fh:
            fh.write(" ".join(NVCC_FLAGS))
    return LIB


def build_cli(force: bool = False) -> str:
    src = os.path.join(CSRC, "souporcell_cli.cpp")
    if force or _newer(CLI, [src, LIB]):
        os.makedirs(BIN_DIR, exist_ok=True)
        _run([NVCC, "-O2", "-std=c++17", "-Xcompiler", "-Wall", "-o", CLI, src, "-L" + LIB_DIR, "-lsoupgpu",
              "-Xlinker", "-rpath,$ORIGIN/../lib"])
    tsrc = os.path.join(CSRC, "troublet_cli.cpp")
    if force or _newer(TROUBLET_CLI, [tsrc, LIB]):
        _run([NVCC, "-O2", "-std=c++17", "-Xcompiler", "-Wall", "-o", TROUBLET_CLI, tsrc, "-L" + LIB_DIR, "-lsoupgpu",
              "-Xlinker", "-rpath,$ORIGIN/../lib"])
    return CLI


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle.cpp", "oracle_capi.cpp", "oracle_cli.cpp", "oracle.hpp", "troublet_oracle.cpp", "troublet_oracle.hpp", "Makefile")]
    if force or _newer(ORACLE_LIB, srcs) or _newer(ORACLE_CLI, srcs) or _newer(os.path.join(ORACLE_DIR, "_build", "troublet_oracle"), srcs):
        _run(["make", "-C", ORACLE_DIR] + (["-B"] if force else []))
    return ORACLE_LIB


def build_all(force: bool = False):
    build_lib(force)
    build_cli(force)
    build_oracle(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB, CLI, ORACLE_LIB, sep="\n")
