"""ctypes binding of libsoupgpu's C ABI (include/soupgpu.h) — the host-side mirror of the
reference's solver entry points (/root/reference/souporcell/src/main.rs, "S:"):

    load_mtx()            load_cell_data   S:601-681      (host, hash-free parallel parser)
    Context.load_csr()    Vec<CellData> -> device CSR + CSC transpose built on the GPU
    Context.solve()       one run of souporcell_main's restart loop  S:77-125  (em S:189 / khm S:280)
    Context.estep/mstep() one pass of the while-body  S:226-249  (parity hooks)
    main()                souporcell_main  S:60-131
    write_clusters()      the TSV writer   S:133-147

There is no CPU fallback: every device entry point raises SoupError when the CUDA library or
a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOUP_LIB") or os.path.join(_HERE, "lib", "libsoupgpu.so")

OK, ERR_INVALID, ERR_CUDA, ERR_IO, ERR_UNSUPPORTED, ERR_NOMEM = 0, -1, -2, -3, -4, -5
METHOD_EM, METHOD_KHM = 0, 1
INIT_RANDOM_UNIFORM, INIT_RANDOM_ASSIGNMENT = 0, 1
MTX_PINNED = 1

EXPORTS = [
    "soup_version", "soup_last_error", "soup_mtx_load", "soup_mtx_free", "soup_barcodes_load", "soup_free",
    "soup_known_genotypes_load", "soup_chacha_words", "soup_thread_seeds", "soup_format_f32", "soup_ln_binomial_f32", "soup_bad_cluster_detection",
    "soup_ctx_create", "soup_ctx_destroy", "soup_load_csr", "soup_get_dims", "soup_get_csc", "soup_get_cell_totals",
    "soup_get_csr_lbc", "soup_cluster_allele_counts", "soup_solve", "soup_get_trace", "soup_get_stats", "soup_estep", "soup_mstep", "soup_device_math",
    "soup_main", "soup_exchange_best", "soup_write_clusters", "soup_clusters_load", "soup_troublet", "soup_troublet_write", "soup_mtx_save", "soup_mtx_load_cache",
    "soup_mtx_write_text", "soup_comm_unique_id", "soup_comm_create", "soup_comm_create_all", "soup_comm_dup_all", "soup_comm_destroy", "soup_comm_rank", "soup_comm_allgather", "soup_comm_bcast",
]


class SoupError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsoupgpu error {code}: {msg}")
        self.code = code


class _Mtx(C.Structure):
    _fields_ = [("n_cells", C.c_uint64), ("n_loci_total", C.c_uint64), ("n_loci", C.c_uint64), ("nnz", C.c_uint64),
                ("row_ptr", C.POINTER(C.c_uint64)), ("locus", C.POINTER(C.c_uint32)), ("alt", C.POINTER(C.c_uint32)),
                ("ref", C.POINTER(C.c_uint32)), ("index_to_locus", C.POINTER(C.c_uint64)), ("pinned", C.c_int32)]


class _SolveParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("method", C.c_int32), ("n_streams", C.c_int32), ("solves_per_stream", C.c_int32),
                ("stream_seeds", C.c_void_p), ("theta0", C.c_void_p), ("init_theta", C.c_void_p), ("base_theta", C.c_void_p),
                ("reinit_clusters", C.c_void_p), ("n_reinit", C.c_int32), ("iteration_cap", C.c_int32),
                ("max_slots", C.c_int32), ("lanes", C.c_int32), ("record_counts", C.c_int32), ("keep_trace", C.c_int32),
                ("poll_chunk", C.c_int32), ("time_kernels", C.c_int32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32),
                ("cell_allreduce", C.c_void_p), ("comm_user", C.c_void_p), ("n_cells_global", C.c_uint64), ("init_strategy", C.c_int32),
                ("cell_comm", C.c_void_p), ("cell_ranges", C.c_int32)]


class _SolveResult(C.Structure):
    _fields_ = [("has_best", C.c_int32), ("best_solve", C.c_int32), ("best_loss", C.c_float),
                ("best_log_probs", C.c_void_p), ("best_theta", C.c_void_p), ("solve_loss", C.c_void_p),
                ("solve_iters", C.c_void_p), ("nnz_updates", C.c_uint64)]


class _IterRecord(C.Structure):
    _fields_ = [("temp_step", C.c_int32), ("loss", C.c_float), ("change", C.c_float), ("clusters_above_threshold", C.c_int32)]


class _Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("estep_launches", C.c_uint64), ("mstep_launches", C.c_uint64),
                ("solve_ms", C.c_double), ("estep_ms", C.c_double), ("mstep_ms", C.c_double), ("timed_iterations", C.c_uint64),
                ("slot_iterations", C.c_uint64), ("slots", C.c_int32), ("lanes", C.c_int32),
                ("groups", C.c_int32), ("padded_cols", C.c_int32), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("allreduce_bytes", C.c_uint64), ("allreduce_calls", C.c_uint64)]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
BCAST_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32)


class _MainParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("method", C.c_int32), ("restarts", C.c_uint32), ("threads", C.c_uint32),
                ("seed", C.c_uint8), ("souporcell3", C.c_int32), ("iteration_cap", C.c_int32), ("max_slots", C.c_int32),
                ("lanes", C.c_int32), ("record_counts", C.c_int32), ("verbose_trace", C.c_int32), ("time_kernels", C.c_int32),
                ("known_theta", C.c_void_p), ("proc_rank", C.c_int32), ("proc_world", C.c_int32), ("allgather", ALLGATHER_FN), ("bcast", BCAST_FN),
                ("comm_user", C.c_void_p), ("init_strategy", C.c_int32), ("cell_allreduce", C.c_void_p), ("cell_comm_user", C.c_void_p),
                ("n_cells_global", C.c_uint64), ("cell_allgather", ALLGATHER_FN), ("cell_group_size", C.c_int32),
                ("cell_comms", C.POINTER(C.c_void_p)), ("proc_comm", C.c_void_p), ("cell_groups", C.c_int32)]


class _MainResult(C.Structure):
    _fields_ = [("has_best", C.c_int32), ("runs", C.c_int32), ("best_loss", C.c_float), ("best_log_probs", C.c_void_p),
                ("best_theta", C.c_void_p), ("nnz_updates", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("slot_iterations", C.c_uint64), ("timed_iterations", C.c_uint64), ("solve_ms", C.c_double),
                ("estep_ms", C.c_double), ("mstep_ms", C.c_double)]


_lib = None


def _preload_nccl():
    """libsoupgpu links libnccl.so.2.  torch ships its own copy under the same soname; whichever is loaded first serves both, and torch
    needs ITS version.  So in a Python process the bundled one is mapped first (a no-op when torch is already imported); the
    `souporcell` executable, which has no torch, takes the system library."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def lib():
    """Load libsoupgpu.so (built in-tree by souporcell_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SoupError(ERR_CUDA, f"{LIB_PATH} is missing — run `python -m souporcell_b200.build`; there is no fallback path")
        _preload_nccl()
        L = C.CDLL(LIB_PATH)
        L.soup_last_error.restype = C.c_char_p
        L.soup_last_error.argtypes = [C.c_void_p]
        L.soup_ln_binomial_f32.restype = C.c_float
        L.soup_ln_binomial_f32.argtypes = [C.c_uint32, C.c_uint32]
        L.soup_format_f32.argtypes = [C.c_float, C.c_int32, C.c_char_p, C.c_int32]
        L.soup_mtx_load.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.POINTER(C.POINTER(_Mtx))]
        L.soup_mtx_free.argtypes = [C.POINTER(_Mtx)]
        L.soup_mtx_free.restype = None
        L.soup_free.argtypes = [C.c_void_p]
        L.soup_free.restype = None
        L.soup_barcodes_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.soup_known_genotypes_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p]
        L.soup_chacha_words.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p, C.c_uint64]
        L.soup_thread_seeds.argtypes = [C.c_uint8, C.c_uint32, C.c_uint32, C.c_void_p]
        L.soup_bad_cluster_detection.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.soup_ctx_create.argtypes = [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
        L.soup_ctx_destroy.argtypes = [C.c_void_p]
        L.soup_ctx_destroy.restype = None
        L.soup_load_csr.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.soup_get_dims.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.soup_get_csc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.soup_get_cell_totals.argtypes = [C.c_void_p, C.c_void_p]
        L.soup_get_csr_lbc.argtypes = [C.c_void_p, C.c_void_p]
        L.soup_cluster_allele_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.soup_solve.argtypes = [C.c_void_p, C.POINTER(_SolveParams), C.POINTER(_SolveResult)]
        L.soup_get_trace.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.soup_get_stats.argtypes = [C.c_void_p, C.POINTER(_Stats)]
        L.soup_estep.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 7
        L.soup_mstep.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.soup_device_math.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.soup_main.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(_MainParams), C.POINTER(_MainResult), C.c_void_p]
        L.soup_exchange_best.argtypes = [C.c_int32, C.c_int32, ALLGATHER_FN, BCAST_FN, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_int32)]
        L.soup_write_clusters.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int32]
        L.soup_clusters_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.soup_troublet.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p]
        L.soup_troublet_write.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int32, C.c_void_p]
        L.soup_mtx_write_text.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
        L.soup_comm_unique_id.argtypes = [C.c_void_p]
        L.soup_comm_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.soup_comm_create_all.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]
        L.soup_comm_dup_all.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(C.c_void_p)]
        L.soup_comm_destroy.argtypes = [C.c_void_p]
        L.soup_comm_destroy.restype = None
        L.soup_comm_rank.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.soup_comm_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.soup_comm_bcast.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32]
        _lib = L
    return _lib


def _check(rc, ctx=None):
    if rc < 0:
        msg = lib().soup_last_error(ctx)
        raise SoupError(rc, msg.decode("utf-8", "replace") if msg else "")
    return rc


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ host side
@dataclass
class CellCSR:
    """What load_cell_data returns (S:680): the filtered per-cell sparse vectors as a CSR."""
    n_cells: int
    n_loci_total: int
    n_loci: int
    row_ptr: np.ndarray         # uint64 [n_cells + 1]
    locus: np.ndarray           # uint32 [nnz]   used-locus index, ascending inside a cell
    alt: np.ndarray             # uint32 [nnz]
    ref: np.ndarray             # uint32 [nnz]
    index_to_locus: np.ndarray  # uint64 [n_loci]

    @property
    def nnz(self) -> int:
        return int(self.locus.shape[0])


MTX_RAW = 2


def _mtx_to_csr(m) -> "CellCSR":
    L = lib()
    try:
        s = m.contents
        n, nnz, nl = int(s.n_cells), int(s.nnz), int(s.n_loci)
        cp = lambda p, cnt, dt: np.ctypeslib.as_array(p, shape=(max(cnt, 1),))[:cnt].astype(dt, copy=True)
        return CellCSR(n, int(s.n_loci_total), nl, cp(s.row_ptr, n + 1, np.uint64), cp(s.locus, nnz, np.uint32),
                       cp(s.alt, nnz, np.uint32), cp(s.ref, nnz, np.uint32), cp(s.index_to_locus, nl, np.uint64))
    finally:
        L.soup_mtx_free(m)


def _cache_argtypes(L):
    L.soup_mtx_load_cache.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.POINTER(_Mtx))]
    L.soup_mtx_save.argtypes = [C.POINTER(_Mtx), C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]


def load_mtx_cache(cache, alt_path, ref_path, min_alt=4, min_ref=4, min_alt_umis=0, min_ref_umis=0, raw=False) -> "CellCSR":
    """soup_mtx_load_cache: the matrix stored in the .soupcsr file `cache`, which must have been built with this filter from these
    very alt / ref files (size, mtime and content fingerprint) — SoupError otherwise."""
    L = lib()
    _cache_argtypes(L)
    m = C.POINTER(_Mtx)()
    _check(L.soup_mtx_load_cache(os.fsencode(cache), os.fsencode(alt_path), os.fsencode(ref_path), min_alt, min_ref, min_alt_umis, min_ref_umis,
                                 MTX_RAW if raw else 0, C.byref(m)))
    return _mtx_to_csr(m)


def load_mtx(alt_path, ref_path, min_alt=4, min_ref=4, min_alt_umis=0, min_ref_umis=0, raw=False, cache=None) -> CellCSR:
    """load_cell_data S:601-681 (defaults S:764-767, 811-816); raw=True: troublet's view of the files (SOUP_MTX_RAW).
    `cache`: path of a .soupcsr file — read if present and built with the same filter from these very files (size / mtime / content
    fingerprint), else (re)written after the text parse, exactly as `souporcell --csr_cache` does."""
    L = lib()
    _cache_argtypes(L)
    m = C.POINTER(_Mtx)()
    flags = MTX_RAW if raw else 0
    a, r = os.fsencode(alt_path), os.fsencode(ref_path)
    hit = False
    if cache is not None and os.path.exists(cache):
        rc = L.soup_mtx_load_cache(os.fsencode(cache), a, r, min_alt, min_ref, min_alt_umis, min_ref_umis, flags, C.byref(m))
        if rc == ERR_INVALID:      # another filter, other source matrices or an older layout: rebuild it
            m = C.POINTER(_Mtx)()
        else:
            _check(rc)
            hit = True
    if not hit:
        _check(L.soup_mtx_load(a, r, min_alt, min_ref, min_alt_umis, min_ref_umis, flags, C.byref(m)))
        if cache is not None:
            _check(L.soup_mtx_save(m, os.fsencode(cache), a, r, min_alt, min_ref, min_alt_umis, min_ref_umis, flags))
    return _mtx_to_csr(m)


def load_barcodes(path):
    """load_barcodes S:707-718 (plain or .gz)."""
    L = lib()
    blob = C.c_void_p()
    n = C.c_uint64()
    _check(L.soup_barcodes_load(os.fsencode(path), C.byref(blob), C.byref(n)))
    try:
        out, off = [], 0
        for _ in range(n.value):
            s = C.string_at(blob.value + off)
            out.append(s.decode("utf-8", "replace"))
            off += len(s) + 1
        return out
    finally:
        L.soup_free(blob)


def known_genotypes(vcf_path, sample_names, k: int, index_to_locus) -> np.ndarray:
    """init_cluster_centers_known_genotypes S:514-542 -> centers [k][n_loci]."""
    i2l = np.ascontiguousarray(index_to_locus, dtype=np.uint64)
    out = np.empty((k, i2l.shape[0]), dtype=np.float32)
    blob = b"".join(s.encode() + b"\0" for s in sample_names) + b"\0"
    _check(lib().soup_known_genotypes_load(os.fsencode(vcf_path), blob, len(sample_names), k, _ptr(i2l), i2l.shape[0], _ptr(out)))
    return out


def chacha_words(seed: bytes, n: int, word_offset: int = 0, rounds: int = 12) -> np.ndarray:
    out = np.empty(n, dtype=np.uint32)
    sb = (C.c_uint8 * 32)(*seed)
    _check(lib().soup_chacha_words(sb, rounds, word_offset, _ptr(out), n))
    return out


def thread_seeds(seed: int, threads: int, run: int = 0) -> np.ndarray:
    out = np.empty((threads, 32), dtype=np.uint8)
    _check(lib().soup_thread_seeds(seed, threads, run, _ptr(out)))
    return out


def format_f32(v: float, width8: bool = False) -> str:
    buf = C.create_string_buffer(96)
    _check(lib().soup_format_f32(C.c_float(v), 1 if width8 else 0, buf, 96))
    return buf.value.decode()


def ln_binomial_f32(n: int, k: int) -> float:
    return float(lib().soup_ln_binomial_f32(n, k))


def bad_cluster_detection(run: int, k: int, best_log_probs: np.ndarray):
    lp = np.ascontiguousarray(best_log_probs, dtype=np.float32)
    out = np.empty(max(k, 1), dtype=np.int32)
    n = _check(lib().soup_bad_cluster_detection(run, k, _ptr(lp), lp.shape[0] if lp.ndim == 2 else 0, _ptr(out), None))
    return out[:n].tolist()


class _TroubletParams(C.Structure):
    _fields_ = [("p_soup", C.c_double), ("doublet_prior", C.c_double), ("doublet_threshold", C.c_double), ("singlet_threshold", C.c_double)]


class _TroubletResult(C.Structure):
    _fields_ = [("status", C.c_void_p), ("best_singlet", C.c_void_p), ("doublet1", C.c_void_p), ("doublet2", C.c_void_p),
                ("singlet_posterior", C.c_void_p), ("doublet_posterior", C.c_void_p), ("log_prob_singleton", C.c_void_p),
                ("log_prob_doublet", C.c_void_p), ("cluster_logprobs", C.c_void_p), ("rounds", C.c_int32), ("removed", C.c_uint64),
                ("device_ms", C.c_double)]


def load_clusters(path):
    """load_clusters T:414-449 -> (barcodes, losses f64 [n_cells][k], n_clusters)."""
    L = lib()
    blob, losses = C.c_void_p(), C.c_void_p()
    n, k, nc = C.c_uint64(), C.c_int32(), C.c_int32()
    _check(L.soup_clusters_load(os.fsencode(path), C.byref(blob), C.byref(losses), C.byref(n), C.byref(k), C.byref(nc)))
    try:
        names, off = [], 0
        for _ in range(n.value):
            sname = C.string_at(blob.value + off)
            names.append(sname.decode())
            off += len(sname) + 1
        arr = np.ctypeslib.as_array(C.cast(losses, C.POINTER(C.c_double)), shape=(max(n.value * k.value, 1),))[:n.value * k.value]
        return names, arr.reshape(n.value, k.value).copy(), nc.value
    finally:
        L.soup_free(blob)
        L.soup_free(losses)


def troublet(ctx, raw: "CellCSR", losses, n_clusters=None, barcodes=None, p_soup=0.0, doublet_prior=0.5, doublet_threshold=0.9,
             singlet_threshold=0.9, out_path=None):
    """call_doublets T:32-244 on the GPU over the RAW matrix (load_mtx(..., raw=True)).  Returns a dict of per-cell arrays."""
    losses = np.ascontiguousarray(losses, dtype=np.float64)
    n, k = losses.shape
    assert n == raw.n_cells
    p = _TroubletParams(p_soup, doublet_prior, doublet_threshold, singlet_threshold)
    out = dict(status=np.zeros(n, np.int32), best_singlet=np.zeros(n, np.int32), doublet1=np.zeros(n, np.int32), doublet2=np.zeros(n, np.int32),
               singlet_posterior=np.zeros(n), doublet_posterior=np.zeros(n), log_prob_singleton=np.zeros(n), log_prob_doublet=np.zeros(n),
               cluster_logprobs=np.zeros((n, k)))
    r = _TroubletResult(*[_ptr(out[f]) for f, _ in _TroubletResult._fields_[:9]])
    blob = None if barcodes is None else b"".join(b.encode() + b"\0" for b in barcodes) + b"\0"
    ctx._ck(lib().soup_troublet(ctx._h, raw.n_cells, raw.n_loci, raw.nnz, _ptr(raw.row_ptr), _ptr(raw.locus), _ptr(raw.alt), _ptr(raw.ref),
                                        _ptr(losses), k, k if n_clusters is None else n_clusters, C.byref(p), C.byref(r), blob, None))
    out.update(rounds=r.rounds, removed=int(r.removed), device_ms=float(r.device_ms))
    if out_path is not None:
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        f = libc.fopen(os.fsencode(out_path), b"w")
        if not f:
            raise OSError(f"cannot open {out_path}")
        try:
            _check(lib().soup_troublet_write(f, blob if blob is not None else b"\0" * (n + 1), n, k, C.byref(r)))
        finally:
            libc.fclose(f)
    return out


def write_clusters(path, barcodes, best_log_probs: np.ndarray):
    """TSV writer S:133-147 into `path`."""
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    lp = np.ascontiguousarray(best_log_probs, dtype=np.float32)
    blob = b"".join(b.encode() + b"\0" for b in barcodes) + b"\0"
    f = libc.fopen(os.fsencode(path), b"w")
    if not f:
        raise OSError(f"cannot open {path}")
    try:
        _check(lib().soup_write_clusters(f, blob, len(barcodes), _ptr(lp), lp.shape[0], lp.shape[1]))
    finally:
        libc.fclose(f)


# ------------------------------------------------------------------ device side
@dataclass
class SolveResult:
    has_best: bool
    best_solve: int
    best_loss: float
    best_log_probs: np.ndarray   # [N][K]
    best_theta: np.ndarray       # [K][L]
    solve_loss: np.ndarray       # [n_solves] (NaN for solves of other shards)
    solve_iters: np.ndarray
    nnz_updates: int
    stats: dict


class Context:
    """One GPU: owns the device copy of the matrix and the resident-slot workspace."""

    def __init__(self, device: int = 0, stream=None):
        self._h = C.c_void_p()
        _check(lib().soup_ctx_create(device, stream, C.byref(self._h)))
        self.n_cells = self.n_loci = self.nnz = 0

    def close(self):
        if self._h:
            lib().soup_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        return _check(rc, self._h)

    def load_csr(self, n_cells, n_loci, row_ptr, locus, alt, ref):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        locus = np.ascontiguousarray(locus, dtype=np.uint32)
        alt = np.ascontiguousarray(alt, dtype=np.uint32)
        ref = np.ascontiguousarray(ref, dtype=np.uint32)
        self._ck(lib().soup_load_csr(self._h, n_cells, n_loci, locus.shape[0], _ptr(row_ptr), _ptr(locus), _ptr(alt), _ptr(ref)))
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ck(lib().soup_get_dims(self._h, C.byref(a), C.byref(b), C.byref(c)))
        self.n_cells, self.n_loci, self.nnz = a.value, b.value, c.value
        return self

    def load(self, m: CellCSR):
        return self.load_csr(m.n_cells, m.n_loci, m.row_ptr, m.locus, m.alt, m.ref)

    def get_csc(self):
        col_ptr = np.empty(self.n_loci + 1, dtype=np.uint64)
        cell = np.empty(self.nnz, dtype=np.uint32)
        alt = np.empty(self.nnz, dtype=np.uint32)
        ref = np.empty(self.nnz, dtype=np.uint32)
        self._ck(lib().soup_get_csc(self._h, _ptr(col_ptr), _ptr(cell), _ptr(alt), _ptr(ref)))
        return col_ptr, cell, alt, ref

    def get_cell_totals(self):
        out = np.empty(self.n_cells, dtype=np.float32)
        self._ck(lib().soup_get_cell_totals(self._h, _ptr(out)))
        return out

    def get_csr_lbc(self):
        out = np.empty(self.nnz, dtype=np.float32)
        self._ck(lib().soup_get_csr_lbc(self._h, _ptr(out)))
        return out

    def cluster_allele_counts(self, cell_cluster, k):
        """[n_loci][k][2] = (ref, alt) sums over the cells assigned to each cluster (consensus.py:296-331); cluster < 0 skips a cell."""
        cl = np.ascontiguousarray(cell_cluster, dtype=np.int32).reshape(self.n_cells)
        out = np.zeros((self.n_loci, k, 2), dtype=np.uint64)
        self._ck(lib().soup_cluster_allele_counts(self._h, _ptr(cl), k, _ptr(out)))
        return out

    def stats(self) -> dict:
        s = _Stats()
        self._ck(lib().soup_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in _Stats._fields_}

    def solve(self, k, method=METHOD_EM, n_streams=1, solves_per_stream=1, stream_seeds=None, theta0=None,
              base_theta=None, reinit_clusters=None, iteration_cap=0, max_slots=0, lanes=0, record_counts=False,
              keep_trace=False, poll_chunk=0, time_kernels=False, shard_rank=0, shard_world=1, cell_allreduce=None,
              n_cells_global=0, init_strategy=INIT_RANDOM_UNIFORM, cell_comm=None) -> SolveResult:
        n_solves = n_streams * solves_per_stream
        p = _SolveParams()
        p.k, p.method, p.n_streams, p.solves_per_stream = k, method, n_streams, solves_per_stream
        keep = []
        if stream_seeds is not None:
            ss = np.ascontiguousarray(stream_seeds, dtype=np.uint8).reshape(n_streams, 32)
            keep.append(ss)
            p.stream_seeds = _ptr(ss)
        if theta0 is not None:
            t0 = np.ascontiguousarray(theta0, dtype=np.float32).reshape(n_solves, k, self.n_loci)
            keep.append(t0)
            p.theta0 = _ptr(t0)
        if base_theta is not None:
            bt = np.ascontiguousarray(base_theta, dtype=np.float32).reshape(k, self.n_loci)
            keep.append(bt)
            p.base_theta = _ptr(bt)
        if reinit_clusters is not None and len(reinit_clusters):
            rc_ = np.ascontiguousarray(reinit_clusters, dtype=np.int32)
            keep.append(rc_)
            p.reinit_clusters = _ptr(rc_)
            p.n_reinit = rc_.shape[0]
        p.iteration_cap, p.max_slots, p.lanes = iteration_cap, max_slots, lanes
        p.record_counts, p.keep_trace, p.poll_chunk = int(record_counts), int(keep_trace), poll_chunk
        p.shard_rank, p.shard_world, p.time_kernels = shard_rank, shard_world, int(time_kernels)
        p.init_strategy = init_strategy
        if cell_allreduce is not None:
            p.cell_allreduce = C.cast(cell_allreduce, C.c_void_p)
            p.n_cells_global = n_cells_global
            keep.append(cell_allreduce)
        if cell_comm is not None:
            p.cell_comm = cell_comm._h
            p.n_cells_global = n_cells_global
        r = _SolveResult()
        lp = np.zeros((self.n_cells, k), dtype=np.float32)
        th = np.zeros((k, self.n_loci), dtype=np.float32)
        sl = np.empty(n_solves, dtype=np.float32)
        si = np.empty(n_solves, dtype=np.int32)
        r.best_log_probs, r.best_theta, r.solve_loss, r.solve_iters = _ptr(lp), _ptr(th), _ptr(sl), _ptr(si)
        self._ck(lib().soup_solve(self._h, C.byref(p), C.byref(r)))
        return SolveResult(bool(r.has_best), r.best_solve, float(r.best_loss), lp, th, sl, si, int(r.nnz_updates), self.stats())

    def get_trace(self, solve: int, cap: int = 1000):
        rec = (_IterRecord * cap)()
        n = C.c_int32()
        self._ck(lib().soup_get_trace(self._h, solve, rec, cap, C.byref(n)))
        return [(rec[i].temp_step, rec[i].loss, rec[i].change, rec[i].clusters_above_threshold) for i in range(min(n.value, cap))]

    def estep(self, theta, temp_step, method=METHOD_EM, lanes=0):
        """One E pass for n_slots center sets. theta [S][K][L]; returns dict of log_probs/weights [S][N][K], lse [S][N], loss [S], counts [S][K]."""
        theta = np.ascontiguousarray(theta, dtype=np.float32)
        S, K, Lc = theta.shape
        assert Lc == self.n_loci
        ts = np.ascontiguousarray(np.broadcast_to(np.asarray(temp_step, dtype=np.int32), (S,)))
        lp = np.empty((S, self.n_cells, K), dtype=np.float32)
        w = np.empty((S, self.n_cells, K), dtype=np.float32)
        lse = np.empty((S, self.n_cells), dtype=np.float32)
        loss = np.empty(S, dtype=np.float32)
        cnt = np.empty((S, K), dtype=np.int32)
        self._ck(lib().soup_estep(self._h, K, method, S, lanes, _ptr(theta), _ptr(ts), _ptr(lp), _ptr(w), _ptr(lse), _ptr(loss), _ptr(cnt)))
        return dict(log_probs=lp, weights=w, lse=lse, loss=loss, counts=cnt)

    def mstep(self, weights, theta_in, locked=(), lanes=0):
        weights = np.ascontiguousarray(weights, dtype=np.float32)
        theta_in = np.ascontiguousarray(theta_in, dtype=np.float32)
        S, N, K = weights.shape
        assert N == self.n_cells and theta_in.shape == (S, K, self.n_loci)
        lk = np.ascontiguousarray(list(locked), dtype=np.int32)
        out = np.empty_like(theta_in)
        self._ck(lib().soup_mstep(self._h, K, S, lanes, _ptr(weights), _ptr(theta_in), _ptr(lk) if lk.size else None, lk.size, _ptr(out)))
        return out

    def device_math(self, op: str, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._ck(lib().soup_device_math(self._h, 0 if op == "ln" else 1, _ptr(x), _ptr(out), x.size))
        return out


@dataclass
class MainResult:
    has_best: bool
    runs: int
    best_loss: float
    best_log_probs: np.ndarray
    best_theta: np.ndarray
    nnz_updates: int
    kernel_launches: int
    solve_ms: float
    slot_iterations: int = 0
    timed_iterations: int = 0
    estep_ms: float = 0.0
    mstep_ms: float = 0.0

    @property
    def assignments(self) -> np.ndarray:
        """FIRST-max cluster per cell, the rule of the TSV writer (S:134-141)."""
        return np.argmax(self.best_log_probs, axis=1)


def main(ctxs, k, method=METHOD_EM, restarts=100, threads=1, seed=4, souporcell3=False, iteration_cap=0, max_slots=0,
         lanes=0, record_counts=False, verbose_trace=False, time_kernels=False, proc_rank=0, proc_world=1, allgather=None,
         bcast=None, known_theta=None, init_strategy=INIT_RANDOM_UNIFORM, cell_allreduce=None, cell_allgather=None, cell_group_size=0,
         n_cells_global=0, cell_comms=None, proc_comm=None, cell_groups=1) -> MainResult:
    """souporcell_main S:60-131 over one context per GPU: restart-sharded, or — with `cell_comms`, one Comm per context — cell-sharded
    (context i holds the i-th contiguous range of the cells; the result covers all of them in order)."""
    if isinstance(ctxs, Context):
        ctxs = [ctxs]
    if isinstance(cell_comms, Comm):
        cell_comms = [cell_comms]
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    p = _MainParams()
    p.k, p.method, p.restarts, p.threads, p.seed = k, method, restarts, threads, seed
    p.souporcell3, p.iteration_cap, p.max_slots, p.lanes = int(souporcell3), iteration_cap, max_slots, lanes
    p.record_counts, p.verbose_trace, p.time_kernels = int(record_counts), int(verbose_trace), int(time_kernels)
    p.proc_rank, p.proc_world = proc_rank, proc_world
    p.init_strategy = init_strategy
    if cell_allreduce is not None:   # cell sharding under the driver (hybrid when proc_world > 1)
        p.cell_allreduce = C.cast(cell_allreduce, C.c_void_p)
        p.cell_allgather = cell_allgather
        p.cell_group_size, p.n_cells_global = cell_group_size, n_cells_global
    if known_theta is not None:
        kt = np.ascontiguousarray(known_theta, dtype=np.float32).reshape(k, ctxs[0].n_loci)
        p.known_theta = _ptr(kt)
    if allgather is not None:
        p.allgather = allgather
    if bcast is not None:
        p.bcast = bcast
    if cell_comms is not None:
        comm_arr = (C.c_void_p * len(ctxs))(*[c._h for c in cell_comms])
        p.cell_comms = comm_arr
        p.n_cells_global = n_cells_global
        p.cell_groups = cell_groups
    if proc_comm is not None:
        p.proc_comm = proc_comm._h
    n, L = (sum(c.n_cells for c in ctxs[:len(ctxs) // max(cell_groups, 1)]) if cell_comms is not None else ctxs[0].n_cells), ctxs[0].n_loci
    lp = np.zeros((n, k), dtype=np.float32)
    th = np.zeros((k, L), dtype=np.float32)
    r = _MainResult()
    r.best_log_probs, r.best_theta = _ptr(lp), _ptr(th)
    _check(lib().soup_main(arr, len(ctxs), C.byref(p), C.byref(r), None))
    return MainResult(bool(r.has_best), r.runs, float(r.best_loss), lp, th, int(r.nnz_updates), int(r.kernel_launches), float(r.solve_ms),
                      int(r.slot_iterations), int(r.timed_iterations), float(r.estep_ms), float(r.mstep_ms))


def write_mtx_text(path, n_loci, n_cells, locus, cell, count, comment="written by souporcell_b200.synth"):
    """soup_mtx_write_text: one vartrix-style .mtx from 0-based int64 coordinates and uint32 counts."""
    lo, ce = np.ascontiguousarray(locus, dtype=np.int64), np.ascontiguousarray(cell, dtype=np.int64)
    cn = np.ascontiguousarray(count, dtype=np.uint32)
    _check(lib().soup_mtx_write_text(os.fsencode(path), n_loci, n_cells, lo.shape[0], _ptr(lo), _ptr(ce), _ptr(cn), comment.encode()))


class Comm:
    """soup_comm: one rank of an NCCL communicator inside libsoupgpu, bound to a context's device."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _check(lib().soup_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def create(cls, ctx, unique_id: bytes, rank: int, world: int) -> "Comm":
        """One rank per process (ncclCommInitRank): `unique_id` comes from Comm.unique_id() on one rank, sent to the others by the launcher."""
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().soup_comm_create(ctx._h, buf, rank, world, C.byref(h)))
        return cls(h)

    @classmethod
    def create_all(cls, ctxs):
        """The contexts of this process (one GPU each) as the ranks of one communicator (ncclCommInitAll)."""
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        out = (C.c_void_p * len(ctxs))()
        _check(lib().soup_comm_create_all(arr, len(ctxs), out))
        return [cls(C.c_void_p(out[i])) for i in range(len(ctxs))]

    @classmethod
    def dup_all(cls, comms):
        """A second communicator over the same ranks of this process (ncclCommSplit), e.g. for another cell group on the same GPUs."""
        arr = (C.c_void_p * len(comms))(*[c._h for c in comms])
        out = (C.c_void_p * len(comms))()
        _check(lib().soup_comm_dup_all(arr, len(comms), out))
        return [cls(C.c_void_p(out[i])) for i in range(len(comms))]

    @classmethod
    def from_torch(cls, ctx, dist) -> "Comm":
        """One rank per process under torchrun: rank 0's unique id travels over the torch.distributed process group (any backend)."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls.create(ctx, box[0], rank, world)

    def rank(self):
        r, w = C.c_int32(), C.c_int32()
        _check(lib().soup_comm_rank(self._h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def allgather(self, data: bytes) -> bytes:
        _, w = self.rank()
        out = (C.c_uint8 * (len(data) * w))()
        _check(lib().soup_comm_allgather(self._h, data, out, len(data)))
        return bytes(out)

    def close(self):
        if self._h:
            lib().soup_comm_destroy(self._h)
            self._h = None


def split_cells(csr, parts: int):
    """Cell sharding: cut the cells into `parts` contiguous ranges balanced by nnz; every shard keeps all loci.
    csr = (n_cells, n_loci, row_ptr, locus, alt, ref).  Returns (cuts [parts + 1], [csr tuple per shard]) — shard r holds cells
    cuts[r]:cuts[r + 1], its row_ptr rebased to 0; a shard may be empty when parts exceeds what the matrix can feed."""
    n, L, row_ptr, locus, alt, ref = csr
    if parts < 1:
        raise ValueError("split_cells: parts must be >= 1")
    rp = np.asarray(row_ptr).astype(np.int64)
    cuts = [0] + [int(np.searchsorted(rp, rp[-1] * i // parts)) for i in range(1, parts)] + [int(n)]
    cuts = [min(max(c, 0), int(n)) for c in cuts]
    out = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        s, e = int(rp[a]), int(rp[b])
        out.append((b - a, L, (rp[a:b + 1] - s).astype(np.uint64), locus[s:e], alt[s:e], ref[s:e]))
    return cuts, out


def exchange_best(rank, world, allgather, bcast, loss, solve, has_best, log_probs=None, theta=None):
    """soup_exchange_best: returns (loss, solve, has_best, winner_rank); log_probs / theta (float32 arrays) are overwritten in place."""
    l, s_, h, w = C.c_float(loss), C.c_int32(solve), C.c_int32(int(has_best)), C.c_int32(-1)
    _check(lib().soup_exchange_best(rank, world, allgather, bcast, None, C.byref(l), C.byref(s_), C.byref(h),
                                    _ptr(log_probs), 0 if log_probs is None else log_probs.size,
                                    _ptr(theta), 0 if theta is None else theta.size, C.byref(w)))
    return float(l.value), int(s_.value), bool(h.value), int(w.value)


def torch_allreduce_callback(dist):
    """soup_allreduce_fn over torch.distributed (NCCL): sums the device buffer in place on torch's current stream —
    create the Context on that stream (Context(device, stream=torch.cuda.current_stream().cuda_stream))."""
    import torch

    class _Dev:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

    def _ar(user, buf, n, stream):
        try:
            t = torch.as_tensor(_Dev(buf, n), device="cuda")
            dist.all_reduce(t)
            return OK
        except Exception:
            return ERR_INVALID

    return ALLREDUCE_FN(_ar)


def torch_exchange_callbacks(dist):
    """allgather / bcast callbacks for `main(..., proc_world>1)` on top of torch.distributed (host tensors
    travel through the process group's backend; used once per souporcell3 run, never in the data path)."""
    import torch

    backend_dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    world = dist.get_world_size()

    def _ag(user, send, recv, nbytes):
        try:
            src = torch.frombuffer((C.c_uint8 * nbytes).from_address(send), dtype=torch.uint8).clone().to(backend_dev)
            outs = [torch.empty_like(src) for _ in range(world)]
            dist.all_gather(outs, src)
            flat = torch.cat(outs).cpu().numpy().tobytes()
            C.memmove(recv, flat, nbytes * world)
            return OK
        except Exception:
            return ERR_INVALID

    def _bc(user, buf, nbytes, root):
        try:
            t = torch.frombuffer((C.c_uint8 * nbytes).from_address(buf), dtype=torch.uint8)
            d = t.clone().to(backend_dev)
            dist.broadcast(d, src=root)
            if dist.get_rank() != root:
                C.memmove(buf, d.cpu().numpy().tobytes(), nbytes)
            return OK
        except Exception:
            return ERR_INVALID

    return ALLGATHER_FN(_ag), BCAST_FN(_bc)
