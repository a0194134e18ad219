"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/_build/liboracle.so, the CPU
restatement of /root/reference/souporcell/src/main.rs.  PARITY UNPINNED (see oracle.hpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
CLI_PATH = os.path.join(_HERE, "_build", "souporcell_oracle")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_data_from_csr.restype = C.c_void_p
        L.orc_data_from_csr.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_data_load_mtx.restype = C.c_void_p
        L.orc_data_load_mtx.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_char_p, C.c_int]
        L.orc_data_free.argtypes = [C.c_void_p]
        L.orc_data_free.restype = None
        for f in ("orc_data_n_cells", "orc_data_n_loci", "orc_data_nnz"):
            getattr(L, f).restype = C.c_uint64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_data_export.argtypes = [C.c_void_p] * 8
        L.orc_data_export.restype = None
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_solve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_main.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint8, C.c_int, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_char_p, C.c_int]
        L.orc_known_genotypes.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_int]
        L.orc_chacha_words.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_chacha_words.restype = None
        L.orc_thread_seeds.argtypes = [C.c_uint8, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_thread_seeds.restype = None
        L.orc_init_uniform.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p]
        L.orc_init_uniform.restype = None
        L.orc_ln_binomial.restype = C.c_double
        L.orc_ln_binomial.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_format_f32.argtypes = [C.c_float, C.c_int, C.c_char_p, C.c_int]
        L.orc_bad_cluster_detection.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_logf.restype = C.c_float
        L.orc_logf.argtypes = [C.c_float]
        L.orc_expf.restype = C.c_float
        L.orc_expf.argtypes = [C.c_float]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleData:
    """Vec<CellData> of the reference (S:683-703), built from mtx files (S:601-681) or a filtered CSR."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: data construction failed")
        self._h = C.c_void_p(handle)
        L = lib()
        self.n_cells = int(L.orc_data_n_cells(self._h))
        self.n_loci = int(L.orc_data_n_loci(self._h))
        self.nnz = int(L.orc_data_nnz(self._h))

    @classmethod
    def from_csr(cls, n_cells, n_loci, row_ptr, locus, alt, ref):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        locus = np.ascontiguousarray(locus, dtype=np.uint32)
        alt = np.ascontiguousarray(alt, dtype=np.uint32)
        ref = np.ascontiguousarray(ref, dtype=np.uint32)
        return cls(lib().orc_data_from_csr(n_cells, n_loci, _p(row_ptr), _p(locus), _p(alt), _p(ref)))

    @classmethod
    def from_mtx(cls, alt_path, ref_path, min_alt=4, min_ref=4, min_alt_umis=0, min_ref_umis=0):
        err = C.create_string_buffer(512)
        h = lib().orc_data_load_mtx(os.fsencode(alt_path), os.fsencode(ref_path), min_alt, min_ref, min_alt_umis, min_ref_umis, err, 512)
        if not h:
            raise RuntimeError("oracle: " + err.value.decode())
        return cls(h)

    def __del__(self):
        try:
            if self._h:
                lib().orc_data_free(self._h)
                self._h = None
        except Exception:
            pass

    def export(self):
        row_ptr = np.empty(self.n_cells + 1, dtype=np.uint64)
        locus = np.empty(self.nnz, dtype=np.uint32)
        alt = np.empty(self.nnz, dtype=np.uint32)
        ref = np.empty(self.nnz, dtype=np.uint32)
        lbc = np.empty(self.nnz, dtype=np.float32)
        tot = np.empty(self.n_cells, dtype=np.float32)
        i2l = np.empty(self.n_loci, dtype=np.uint64)
        lib().orc_data_export(self._h, _p(row_ptr), _p(locus), _p(alt), _p(ref), _p(lbc), _p(tot), _p(i2l))
        return dict(row_ptr=row_ptr, locus=locus, alt=alt, ref=ref, lbc=lbc, total_alleles=tot, index_to_locus=i2l)

    def known_genotypes(self, vcf_path, sample_names, k):
        """init_cluster_centers_known_genotypes S:514-542 -> centers [k][n_loci]."""
        th = np.empty((k, self.n_loci), dtype=np.float32)
        blob = b"".join(s.encode() + b"\0" for s in sample_names) + b"\0"
        err = C.create_string_buffer(512)
        if lib().orc_known_genotypes(self._h, os.fsencode(vcf_path), blob, len(sample_names), k, _p(th), err, 512) != 0:
            raise RuntimeError("oracle: " + err.value.decode())
        return th

    def step(self, theta, method=0, temp_step=0, locked=()):
        """One E pass + M pass (S:226-249).  theta [K][L] -> dict(theta_out, log_probs, weights, lse, loss)."""
        th = np.array(theta, dtype=np.float32, order="C", copy=True)
        K = th.shape[0]
        lp = np.empty((self.n_cells, K), dtype=np.float32)
        w = np.empty((self.n_cells, K), dtype=np.float32)
        lse = np.empty(self.n_cells, dtype=np.float32)
        loss = C.c_float()
        lk = np.ascontiguousarray(list(locked), dtype=np.int32)
        lib().orc_step(self._h, K, method, temp_step, _p(th), _p(lk), lk.size, _p(lp), _p(w), _p(lse), C.byref(loss))
        return dict(theta_out=th, log_probs=lp, weights=w, lse=lse, loss=np.float32(loss.value))

    def solve(self, theta, method=0, locked=(), iteration_cap=1000, inner_threads=1):
        """One em()/khm() call (S:189 / S:280) from theta [K][L].  inner_threads > 1: the iteration's independent chains run
        concurrently (bit-identical to the serial loop; for the full-size parity tests)."""
        th = np.array(theta, dtype=np.float32, order="C", copy=True)
        K = th.shape[0]
        lp = np.empty((self.n_cells, K), dtype=np.float32)
        loss = C.c_float()
        iters = C.c_int()
        cap = max(iteration_cap, 1)
        tl = np.empty(cap, dtype=np.float32)
        tc = np.empty(cap, dtype=np.float32)
        tt = np.empty(cap, dtype=np.int32)
        ta = np.empty(cap, dtype=np.int32)
        lk = np.ascontiguousarray(list(locked), dtype=np.int32)
        lib().orc_solve_mt(self._h, K, method, _p(th), _p(lk), lk.size, iteration_cap, max(1, int(inner_threads)), _p(lp), C.byref(loss),
                           C.byref(iters), cap, _p(tl), _p(tc), _p(tt), _p(ta))
        n = iters.value
        return dict(theta=th, log_probs=lp, loss=np.float32(loss.value), iterations=n,
                    trace=[(int(tt[i]), np.float32(tl[i]), np.float32(tc[i]), int(ta[i])) for i in range(n)])

    def main(self, k, method=0, restarts=100, threads=1, seed=4, souporcell3=False, iteration_cap=1000, theta0=None):
        """souporcell_main S:60-131."""
        spt = int(np.ceil(np.float32(restarts) / np.float32(threads)))
        cap = threads * spt * 3
        lp = np.zeros((self.n_cells, k), dtype=np.float32)
        th = np.zeros((k, self.n_loci), dtype=np.float32)
        sl = np.empty(cap, dtype=np.float32)
        si = np.empty(cap, dtype=np.int32)
        best = C.c_float()
        ns, runs = C.c_int(), C.c_int()
        upd = C.c_uint64()
        err = C.create_string_buffer(512)
        t0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float32)
        rc = lib().orc_main(self._h, k, method, restarts, threads, seed, int(souporcell3), iteration_cap, _p(t0), C.byref(best),
                            _p(lp), _p(th), _p(sl), _p(si), cap, C.byref(ns), C.byref(runs), C.byref(upd), err, 512)
        if rc < 0:
            raise RuntimeError("oracle: " + err.value.decode())
        n = ns.value
        return dict(has_best=(rc == 0), best_loss=np.float32(best.value), best_log_probs=lp, best_theta=th,
                    solve_loss=sl[:n].copy(), solve_iters=si[:n].copy(), runs=runs.value, nnz_updates=int(upd.value))


def chacha_words(seed: bytes, n: int, word_offset: int = 0, rounds: int = 12):
    out = np.empty(n, dtype=np.uint32)
    sb = (C.c_uint8 * 32)(*seed)
    lib().orc_chacha_words(sb, rounds, word_offset, _p(out), n)
    return out


def thread_seeds(seed: int, threads: int, run: int = 0):
    out = np.empty((threads, 32), dtype=np.uint8)
    lib().orc_thread_seeds(seed, threads, run, _p(out))
    return out


def init_uniform(seed32: bytes, word_offset: int, k: int, n_loci: int):
    out = np.empty((k, n_loci), dtype=np.float32)
    sb = (C.c_uint8 * 32)(*seed32)
    lib().orc_init_uniform(sb, word_offset, k, n_loci, _p(out))
    return out


def ln_binomial(n, k):
    return float(lib().orc_ln_binomial(n, k))


def format_f32(v, width8=False):
    buf = C.create_string_buffer(96)
    if lib().orc_format_f32(C.c_float(v), int(width8), buf, 96) < 0:
        raise RuntimeError("format buffer")
    return buf.value.decode()


def bad_cluster_detection(run, k, best_lp):
    lp = np.ascontiguousarray(best_lp, dtype=np.float32)
    out = np.empty(max(k, 1), dtype=np.int32)
    n = lib().orc_bad_cluster_detection(run, k, _p(lp), lp.shape[0], _p(out))
    return out[:n].tolist()


def logf(x):
    return np.float32(lib().orc_logf(C.c_float(x)))


def expf(x):
    return np.float32(lib().orc_expf(C.c_float(x)))
