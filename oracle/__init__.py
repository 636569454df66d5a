"""oracle — TEST INFRASTRUCTURE ONLY (not part of the product).

ctypes binding of ``oracle/liborc.so`` (``graph_oracle.cpp``): the CPU restatement of the
reference's vector path. May be imported only by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

Parity status (see the header of graph_oracle.cpp): HNSW / Vamana search and the two file formats
are **parity unpinned** (third-party crates absent, no reference golden vectors); the exact scan,
BM25, hybrid fusion and the filter are pinned by in-repo reference source and tests
(``oracle/text_oracle.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

METRIC_IP, METRIC_L2SQ, METRIC_IP_CLAMP = 0, 1, 2


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "graph_oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp, sz, u64p, f32p, u32p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        L.orc_hnsw_build.restype = vp
        L.orc_hnsw_build.argtypes = [f32p, sz, sz, sz, sz, C.c_uint64, C.c_int]
        L.orc_hnsw_save.argtypes = [vp, C.c_char_p]
        L.orc_hnsw_load.restype = vp
        L.orc_hnsw_load.argtypes = [C.c_char_p, sz, C.c_char_p, sz]
        L.orc_hnsw_free.argtypes = [vp]
        L.orc_hnsw_len.restype = sz
        L.orc_hnsw_len.argtypes = [vp]
        L.orc_hnsw_info.argtypes = [vp, u64p]
        L.orc_hnsw_search.argtypes = [vp, f32p, sz, sz, sz, C.c_int, u64p, sz, u64p, f32p, u32p, u64p, C.c_int]
        L.orc_vamana_build.restype = vp
        L.orc_vamana_build.argtypes = [f32p, sz, sz, sz, sz, C.c_float, C.c_uint64, C.c_int]
        L.orc_vamana_save.argtypes = [vp, C.c_char_p]
        L.orc_vamana_load.restype = vp
        L.orc_vamana_load.argtypes = [C.c_char_p, C.c_char_p, sz]
        L.orc_vamana_free.argtypes = [vp]
        L.orc_vamana_info.argtypes = [vp, u64p]
        L.orc_vamana_set_metric.argtypes = [vp, C.c_int]
        L.orc_vamana_search.argtypes = [vp, f32p, sz, sz, sz, C.c_int, u64p, sz, u64p, f32p, u32p, u64p, C.c_int]
        L.orc_exact_scan.argtypes = [f32p, sz, f32p, sz, sz, sz, C.c_int, u64p, u64p, f32p, u32p, C.c_int]
        L.orc_exact_f64.argtypes = [f32p, sz, f32p, sz, sz, sz, C.c_int, u64p, C.c_int]
        L.orc_distance.restype = C.c_float
        L.orc_distance.argtypes = [f32p, f32p, sz, C.c_int, C.c_int]
        L.orc_hardware_threads.restype = C.c_int
        L.orc_compat_flags.restype = C.c_uint32
        L.orc_compat_default.restype = C.c_uint32
        L.orc_set_compat.argtypes = [C.c_uint32]
        L.orc_set_compat.restype = None
        _LIB = L
    return _LIB


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _mask_ptr(mask):
    if mask is None:
        return None, None
    m = np.ascontiguousarray(mask, dtype=np.uint64)
    return m, m.ctypes.data_as(C.POINTER(C.c_uint64))


def pack_mask(bits) -> np.ndarray:
    """bool[N] -> uint64 words, bit i of word i//64 (little-endian bit order)."""
    bits = np.asarray(bits, dtype=bool)
    n = bits.shape[0]
    pad = (-n) % 64
    b = np.concatenate([bits, np.zeros(pad, dtype=bool)]).reshape(-1, 64)
    w = (b.astype(np.uint64) << np.arange(64, dtype=np.uint64)[None, :]).sum(axis=1, dtype=np.uint64)
    return w


class _Graph:
    _search_fn = None
    _free_fn = None

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle: null handle")
        self.h = C.c_void_p(handle)

    def __del__(self):
        try:
            if self.h:
                getattr(lib(), self._free_fn)(self.h)
                self.h = None
        except Exception:
            pass

    def search(self, queries, k, ef, lanes=0, mask=None, next_cap=0, nthreads=1):
        """Returns keys[nq,k] u64, dists[nq,k] f32, counts[nq] u32, stats[nq,4] u64
        (n_dist, n_hops0, n_hops_upper, queue_dropped)."""
        q, qp = _f32(queries)
        nq = q.shape[0]
        keys = np.empty((nq, k), dtype=np.uint64)
        dists = np.empty((nq, k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        stats = np.zeros((nq, 4), dtype=np.uint64)
        m, mp = _mask_ptr(mask)
        getattr(lib(), self._search_fn)(
            self.h, qp, nq, k, ef, lanes, mp, next_cap,
            keys.ctypes.data_as(C.POINTER(C.c_uint64)), dists.ctypes.data_as(C.POINTER(C.c_float)),
            counts.ctypes.data_as(C.POINTER(C.c_uint32)), stats.ctypes.data_as(C.POINTER(C.c_uint64)), nthreads)
        return keys, dists, counts, stats


class Hnsw(_Graph):
    """usearch-semantics HNSW (hnsw.rs:18-139 call sites)."""
    _search_fn = "orc_hnsw_search"
    _free_fn = "orc_hnsw_free"

    @classmethod
    def build(cls, vecs, M=32, ef_add=64, seed=1, metric=METRIC_IP):
        v, vp = _f32(vecs)
        return cls(lib().orc_hnsw_build(vp, v.shape[0], v.shape[1], M, ef_add, seed, metric))

    @classmethod
    def load(cls, path, dims=0):
        err = C.create_string_buffer(256)
        h = lib().orc_hnsw_load(os.fsencode(path), dims, err, 256)
        if not h:
            raise RuntimeError("oracle hnsw_load: " + err.value.decode())
        return cls(h)

    def save(self, path):
        if lib().orc_hnsw_save(self.h, os.fsencode(path)) != 0:
            raise RuntimeError("oracle hnsw_save failed")

    def info(self):
        out = (C.c_uint64 * 7)()
        lib().orc_hnsw_info(self.h, out)
        return dict(zip(["n", "d", "M", "M0", "max_level", "entry", "metric"], [int(x) for x in out]))


class Vamana(_Graph):
    """diskann-rs-semantics Vamana (diskann.rs:21-105 call sites)."""
    _search_fn = "orc_vamana_search"
    _free_fn = "orc_vamana_free"

    @classmethod
    def build(cls, vecs, R=64, L=100, alpha=1.2, seed=1, metric=METRIC_IP_CLAMP):
        v, vp = _f32(vecs)
        return cls(lib().orc_vamana_build(vp, v.shape[0], v.shape[1], R, L, alpha, seed, metric))

    @classmethod
    def load(cls, path):
        err = C.create_string_buffer(256)
        h = lib().orc_vamana_load(os.fsencode(path), err, 256)
        if not h:
            raise RuntimeError("oracle vamana_load: " + err.value.decode())
        return cls(h)

    def save(self, path):
        if lib().orc_vamana_save(self.h, os.fsencode(path)) != 0:
            raise RuntimeError("oracle vamana_save failed")

    def set_metric(self, metric):
        lib().orc_vamana_set_metric(self.h, metric)

    def info(self):
        out = (C.c_uint64 * 4)()
        lib().orc_vamana_info(self.h, out)
        return dict(zip(["n", "d", "R", "medoid"], [int(x) for x in out]))


def exact_scan(queries, db, k, metric=0, mask=None, nthreads=1):
    """recompute.rs:96-110 semantics. metric 0: dot desc; 1: L2sq asc; 2: 1-dot asc."""
    q, qp = _f32(queries)
    x, xp = _f32(db)
    nq = q.shape[0]
    idx = np.empty((nq, k), dtype=np.uint64)
    sc = np.empty((nq, k), dtype=np.float32)
    counts = np.zeros(nq, dtype=np.uint32)
    m, mp = _mask_ptr(mask)
    lib().orc_exact_scan(qp, nq, xp, x.shape[0], x.shape[1], k, metric, mp,
                         idx.ctypes.data_as(C.POINTER(C.c_uint64)), sc.ctypes.data_as(C.POINTER(C.c_float)),
                         counts.ctypes.data_as(C.POINTER(C.c_uint32)), nthreads)
    return idx, sc, counts


def exact_f64(queries, db, k, metric=0, nthreads=0):
    q, qp = _f32(queries)
    x, xp = _f32(db)
    nq = q.shape[0]
    idx = np.empty((nq, k), dtype=np.uint64)
    lib().orc_exact_f64(qp, nq, xp, x.shape[0], x.shape[1], k, metric,
                        idx.ctypes.data_as(C.POINTER(C.c_uint64)), nthreads or hardware_threads())
    return idx


def distance(a, b, metric=0, lanes=0) -> float:
    a, ap = _f32(a)
    b, bp = _f32(b)
    return float(lib().orc_distance(ap, bp, a.shape[0], metric, lanes))


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads())


# Recalled third-party behaviours behind named switches (graph_oracle.cpp CompatBits; the product mirrors them in
# leann_rs_b200/csrc/compat.h and reports them through leann_cuda_compat_flags()).
COMPAT_BITS = {"usearch_stop_strict": 1, "diskann_stop_strict": 2, "top_newcomer_before_equals": 4,
               "next_fifo_among_equals": 8, "distdot_clamp_at_zero": 16}


def compat_flags() -> int:
    return int(lib().orc_compat_flags())


def compat_default() -> int:
    return int(lib().orc_compat_default())


def set_compat(flags: int) -> None:
    """Tests only; not thread-safe (set between searches)."""
    lib().orc_set_compat(int(flags))
