"""leann_rs_b200 — host-side mirror of leann-rs's search interfaces over ``libleann_cuda.so``.

The product is the C-ABI library (``include/leann_cuda.h``, sources in ``csrc/``). This module is
the thin host layer the tests and ``bench.py`` drive it through; class and method names follow the
reference (``src/backend/traits.rs``, ``src/backend/{hnsw,diskann}.rs``, ``src/index/{searcher,
bm25,filter,recompute}.rs``). PyTorch appears only as the owner of device buffers and streams.

There is no CPU fallback: importing works anywhere, but every compute call raises ``LeannCudaError``
when the CUDA library or a B200 is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LEANN_CUDA_LIB") or os.path.join(_HERE, "libleann_cuda.so")   # LEANN_CUDA_LIB: A/B builds of the same ABI

OK = 0
BACKEND_HNSW, BACKEND_VAMANA, BACKEND_FLAT = 0, 1, 2
METRIC_DEFAULT, METRIC_IP, METRIC_L2SQ, METRIC_IP_CLAMP, METRIC_DOT_DESC = -1, 0, 1, 2, 3
MASK_NONE, MASK_INLINE = 0, 1
ERR_NOT_FOUND, ERR_BAD_FORMAT, ERR_DIM_MISMATCH, ERR_CUDA, ERR_NCCL, ERR_INVALID_ARG, ERR_FAISS_FORMAT, ERR_PARSE = (
    -1, -2, -3, -4, -5, -6, -7, -8)


class LeannCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[leann_cuda {code}] {msg}")
        self.code = code
        self.message = msg


_lib = None


def _prefer_bundled_nccl() -> None:
    """The library loads libnccl.so.2 with dlopen on first sharded use. In a Python process that also imports torch, the copy
    torch was built against (the `nvidia-nccl` wheel) must be the one that gets mapped: whichever libnccl.so.2 is loaded first
    satisfies every later request for that SONAME, and an older system copy lacks symbols torch needs. Point the loader at the
    wheel's file unless the caller chose one (LEANN_CUDA_NCCL_LIB)."""
    if os.environ.get("LEANN_CUDA_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["LEANN_CUDA_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def lib():
    """Loads libleann_cuda.so; fails loudly when it has not been built (no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LeannCudaError(ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    _prefer_bundled_nccl()
    L = C.CDLL(LIB_PATH)
    vp, sz, cp = C.c_void_p, C.c_size_t, C.c_char_p
    u64p, f32p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    pp = C.POINTER(C.c_void_p)
    szp = C.POINTER(C.c_size_t)
    cpp = C.POINTER(C.c_char_p)
    sig = {
        "leann_cuda_open": (C.c_int, [cp, C.c_int, sz, C.c_int, C.c_int, pp, cp, sz]),
        "leann_cuda_flat_from_host": (C.c_int, [vp, sz, sz, C.c_int, C.c_int, pp, cp, sz]),
        "leann_cuda_flat_from_device": (C.c_int, [vp, sz, sz, C.c_int, C.c_int, pp, cp, sz]),
        "leann_cuda_hnsw_build": (C.c_int, [vp, C.c_int, sz, sz, sz, sz, C.c_int, C.c_uint64, C.c_int, pp, cp, sz]),
        "leann_cuda_hnsw_add": (C.c_int, [vp, vp, C.c_int, sz, C.c_uint64, sz, C.c_uint64, cp, sz]),
        "leann_cuda_vamana_build": (C.c_int, [vp, C.c_int, sz, sz, sz, sz, C.c_float, C.c_int, C.c_uint64, C.c_int, pp, cp, sz]),
        "leann_cuda_save": (C.c_int, [vp, cp, cp, sz]),
        "leann_cuda_check_index_file": (C.c_int, [cp, C.c_int, sz, u64p, cp, sz]),
        "leann_cuda_write_layout_cache": (C.c_int, [vp, cp, cp, sz]),
        "leann_cuda_layout_cache_used": (C.c_int, [vp]),
        "leann_cuda_len": (sz, [vp]),
        "leann_cuda_dims": (sz, [vp]),
        "leann_cuda_info": (C.c_int, [vp, u64p]),
        "leann_cuda_reduction_lanes": (C.c_int, [sz]),
        "leann_cuda_queue_capacity": (sz, [sz, C.c_int]),
        "leann_cuda_search": (C.c_int, [vp, vp, sz, sz, sz, vp, C.c_int, vp, vp, vp, cp, sz]),
        "leann_cuda_search_device": (C.c_int, [vp, vp, sz, sz, sz, vp, C.c_int, vp, vp, vp, vp, vp, cp, sz]),
        "leann_cuda_topk_merge_device": (C.c_int, [vp, vp, sz, sz, sz, C.c_int, vp, vp, vp, vp, cp, sz]),
        "leann_cuda_close": (None, [vp]),
        "leann_cuda_shards_open": (C.c_int, [cpp, sz, C.c_int, sz, C.c_int, C.POINTER(C.c_int), u64p, C.c_int, pp, cp, sz]),
        "leann_cuda_shards_from_indexes": (C.c_int, [pp, sz, u64p, C.c_int, C.c_int, pp, cp, sz]),
        "leann_cuda_comm_unique_id": (C.c_int, [vp, sz, cp, sz]),
        "leann_cuda_shards_join": (C.c_int, [vp, C.c_int, vp, sz, C.c_int, C.c_int, C.c_uint64, pp, cp, sz]),
        "leann_cuda_shards_len": (sz, [vp]),
        "leann_cuda_shards_count": (sz, [vp]),
        "leann_cuda_shards_info": (C.c_int, [vp, u64p]),
        "leann_cuda_shards_search": (C.c_int, [vp, vp, sz, sz, sz, pp, vp, vp, vp, cp, sz]),
        "leann_cuda_shards_search_device": (C.c_int, [vp, vp, sz, sz, sz, vp, vp, vp, vp, vp, cp, sz]),
        "leann_cuda_shards_dims": (sz, [vp]),
        "leann_cuda_shards_hybrid_search": (C.c_int, [vp, vp, vp, cpp, szp, sz, sz, sz, C.c_int, C.c_float, vp, sz, vp, vp, vp, cp, sz]),
        "leann_cuda_shards_close": (None, [vp]),
        "leann_cuda_set_visited_hash": (C.c_int, [vp, sz]),
        "leann_cuda_set_coalescing": (C.c_int, [vp, sz, C.c_uint]),
        "leann_cuda_coalescing_stats": (C.c_int, [vp, u64p, u64p]),
        "leann_cuda_workspace_stats": (C.c_int, [vp, u64p]),
        "leann_cuda_device_count": (C.c_int, []),
        "leann_cuda_version": (cp, []),
        "leann_cuda_compat_flags": (C.c_uint, []),
        "leann_cuda_bm25_build": (C.c_int, [cpp, szp, sz, C.c_int, pp, cp, sz]),
        "leann_cuda_bm25_len": (sz, [vp]),
        "leann_cuda_bm25_dense_rows": (sz, [vp]),
        "leann_cuda_bm25_stats": (C.c_int, [vp, u64p, f32p]),
        "leann_cuda_tokenize": (sz, [cp, sz, cp, sz]),
        "leann_cuda_bm25_score": (C.c_int, [vp, cp, sz, vp, cp, sz]),
        "leann_cuda_bm25_search": (C.c_int, [vp, cpp, szp, sz, sz, vp, vp, vp, cp, sz]),
        "leann_cuda_bm25_shard_stats": (C.c_int, [cpp, szp, sz, C.c_int, vp, sz, szp, cp, sz]),
        "leann_cuda_bm25_stats_merge": (C.c_int, [vp, szp, sz, vp, sz, szp, cp, sz]),
        "leann_cuda_bm25_build_sharded": (C.c_int, [cpp, szp, sz, vp, sz, C.c_int, pp, cp, sz]),
        "leann_cuda_bm25_search_shard": (C.c_int, [vp, cpp, szp, sz, sz, C.c_uint64, vp, vp, sz, vp, vp, vp, vp, vp, vp, cp, sz]),
        "leann_cuda_hybrid_fuse": (C.c_int, [vp, vp, vp, sz, sz, vp, vp, vp, vp, sz, vp, vp, C.c_int, C.c_float, vp, sz, sz, C.c_int,
                                             vp, vp, vp, cp, sz]),
        "leann_cuda_bm25_last_batch": (C.c_int, [vp, u64p, f32p]),
        "leann_cuda_bm25_last_batch_bytes": (C.c_uint64, [vp]),
        "leann_cuda_hybrid_rerank": (C.c_int, [vp, vp, sz, vp, sz, C.c_float, C.c_int, vp, vp, cp, sz]),
        "leann_cuda_bm25_free": (None, [vp]),
        "leann_cuda_hybrid_search": (C.c_int, [vp, vp, vp, cpp, szp, sz, sz, sz, C.c_int, C.c_float, vp, vp, vp, vp, cp, sz]),
        "leann_cuda_filter_parse": (C.c_int, [cp, pp, cp, sz]),
        "leann_cuda_filter_describe": (sz, [vp, cp, sz]),
        "leann_cuda_filter_matches": (C.c_int, [vp, cp, sz, C.POINTER(C.c_int), cp, sz]),
        "leann_cuda_filter_mask": (C.c_int, [vp, cpp, szp, sz, vp, cp, sz]),
        "leann_cuda_filter_free": (None, [vp]),
        "leann_cuda_metacols_build": (C.c_int, [cpp, szp, sz, pp, cp, sz]),
        "leann_cuda_metacols_mask": (C.c_int, [vp, vp, vp, cp, sz]),
        "leann_cuda_metacols_len": (sz, [vp]),
        "leann_cuda_metacols_fields": (sz, [vp]),
        "leann_cuda_metacols_free": (None, [vp]),
        "leann_cuda_searcher_load": (C.c_int, [cp, cp, sz, C.c_int, pp, cp, sz]),
        "leann_cuda_searcher_len": (sz, [vp]),
        "leann_cuda_searcher_id": (sz, [vp, C.c_uint64, cp, sz]),
        "leann_cuda_searcher_search": (C.c_int, [vp, vp, cpp, szp, sz, sz, sz, cp, C.c_int, C.c_float, vp, vp, vp, cp, sz]),
        "leann_cuda_searcher_close": (None, [vp]),
        "leann_cuda_searcher_set_honor_complexity": (C.c_int, [vp, C.c_int]),
        "leann_cuda_searcher_bm25_search": (C.c_int, [vp, cp, sz, sz, vp, vp, vp, cp, sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name, None)
        if fn is None:
            continue  # reported by tests/test_abi.py
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


ABI_SYMBOLS = None  # filled by tests from include/leann_cuda.h


def _check(code: int, err) -> None:
    if code != OK:
        raise LeannCudaError(code, err.value.decode(errors="replace"))


def _err():
    return C.create_string_buffer(1024)


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


def pack_mask(bits) -> np.ndarray:
    """bool[N] -> u64 words (bit s%64 of word s//64)."""
    bits = np.asarray(bits, dtype=bool)
    pad = (-bits.shape[0]) % 64
    b = np.concatenate([bits, np.zeros(pad, dtype=bool)]).reshape(-1, 64)
    return (b.astype(np.uint64) << np.arange(64, dtype=np.uint64)[None, :]).sum(axis=1, dtype=np.uint64)


def reduction_lanes(dims: int) -> int:
    return int(lib().leann_cuda_reduction_lanes(dims))


def queue_capacity(ef: int, masked: bool) -> int:
    return int(lib().leann_cuda_queue_capacity(ef, 1 if masked else 0))


def device_count() -> int:
    return int(lib().leann_cuda_device_count())


def _strs(items: Sequence):
    from .text import _strs as impl   # one implementation (joined buffer + vectorised pointer table)
    return impl(items)


# ------------------------------------------------------------------------------------------------
# backend/traits.rs:11-30  trait BackendSearcher
# ------------------------------------------------------------------------------------------------
class BackendSearcher:
    """`trait BackendSearcher` (src/backend/traits.rs:11-30) over a device-resident index."""

    backend = BACKEND_HNSW

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    # -- construction ----------------------------------------------------------------------------
    @classmethod
    def _open(cls, base_path: str, backend: int, dimensions: int, metric: int = METRIC_DEFAULT, device: int = 0):
        h = C.c_void_p()
        e = _err()
        _check(lib().leann_cuda_open(os.fsencode(base_path), backend, dimensions, metric, device, C.byref(h), e, 1024), e)
        return cls(h.value)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().leann_cuda_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- trait methods ---------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(lib().leann_cuda_len(self._h))

    def len(self) -> int:
        return len(self)

    def is_empty(self) -> bool:
        return len(self) == 0

    @property
    def dims(self) -> int:
        return int(lib().leann_cuda_dims(self._h))

    def info(self) -> dict:
        out = (C.c_uint64 * 8)()
        lib().leann_cuda_info(self._h, out)
        names = ["n", "dims", "backend", "metric", "M", "M0", "max_level", "entry"]
        return dict(zip(names, [int(x) for x in out]))

    def _effective_ef(self, complexity: int) -> int:
        return complexity

    def search(self, query, top_k: int, complexity: int = 64):
        """One query, exactly the trait call: returns (keys: list[int], distances: list[float])."""
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(1, -1)
        keys, dists, counts = self.search_batch(q, top_k, self._effective_ef(complexity))
        c = int(counts[0])
        return [int(x) for x in keys[0, :c]], [float(x) for x in dists[0, :c]]

    def search_batch(self, queries, top_k: int, ef: int, mask: Optional[np.ndarray] = None):
        """Host buffers in, host buffers out (H2D/D2H inside the call)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.dims:
            raise LeannCudaError(ERR_DIM_MISMATCH, f"queries must be [nq, {self.dims}]")
        nq = q.shape[0]
        keys = np.empty((nq, top_k), dtype=np.uint64)
        dists = np.empty((nq, top_k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint64)
        e = _err()
        _check(lib().leann_cuda_search(self._h, _np_ptr(q), nq, top_k, ef, None if m is None else _np_ptr(m),
                                       MASK_NONE if m is None else MASK_INLINE, _np_ptr(keys), _np_ptr(dists),
                                       _np_ptr(counts), e, 1024), e)
        return keys, dists, counts

    def search_device(self, queries, top_k: int, ef: int, mask=None, out=None, stats=None, stream=None):
        """torch CUDA tensors in/out, enqueued on torch's current stream, no host sync."""
        import torch

        assert queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous()
        nq = queries.shape[0]
        if out is None:
            keys = torch.empty((nq, top_k), dtype=torch.int64, device=queries.device)
            dists = torch.empty((nq, top_k), dtype=torch.float32, device=queries.device)
            counts = torch.empty((nq,), dtype=torch.int32, device=queries.device)
        else:
            keys, dists, counts = out
        st = stream if stream is not None else torch.cuda.current_stream(queries.device).cuda_stream
        e = _err()
        _check(lib().leann_cuda_search_device(
            self._h, C.c_void_p(queries.data_ptr()), nq, top_k, ef,
            None if mask is None else C.c_void_p(mask.data_ptr()), MASK_NONE if mask is None else MASK_INLINE,
            C.c_void_p(keys.data_ptr()), C.c_void_p(dists.data_ptr()), C.c_void_p(counts.data_ptr()),
            None if stats is None else C.c_void_p(stats.data_ptr()), C.c_void_p(st), e, 1024), e)
        return keys, dists, counts

    def set_coalescing(self, max_batch: int, max_wait_us: int = 200):
        """Merge concurrent single-query `search` calls into batched launches (serve.rs traffic)."""
        lib().leann_cuda_set_coalescing(self._h, max_batch, max_wait_us)

    def set_visited_hash(self, capacity: int):
        """Visited set of the traversal: 0 auto, 1 byte maps only, 2 / 3 stand-alone shared-memory tables where supported
        (3: 256-entry limit), 4 / 5 shared-memory first level + hash overflow level (5: tiny levels; 3 and 5 are for tests),
        >= 1024 force per-warp hash tables of this capacity."""
        rc = lib().leann_cuda_set_visited_hash(self._h, capacity)
        if rc != 0:
            raise LeannCudaError(rc, "invalid visited-hash capacity")

    def workspace_stats(self) -> dict:
        out = (C.c_uint64 * 4)()
        lib().leann_cuda_workspace_stats(self._h, out)
        return dict(zip(["reallocs", "bytes", "large_mode", "n_warps"], [int(x) for x in out]))

    def coalescing_stats(self):
        b, r = C.c_uint64(), C.c_uint64()
        lib().leann_cuda_coalescing_stats(self._h, C.byref(b), C.byref(r))
        return int(b.value), int(r.value)

    def save(self, base_path: str):
        e = _err()
        _check(lib().leann_cuda_save(self._h, os.fsencode(base_path), e, 1024), e)

    def write_layout_cache(self, base_path: str):
        """Persist the parsed adjacency as `<base>.cuda-layout` (bound to the `.index` at the same base) so later opens skip the parse."""
        e = _err()
        _check(lib().leann_cuda_write_layout_cache(self._h, os.fsencode(base_path), e, 1024), e)

    @property
    def layout_cache_used(self) -> bool:
        return bool(lib().leann_cuda_layout_cache_used(self._h))


class HnswSearcher(BackendSearcher):
    """src/backend/hnsw.rs:12-93. `search` ignores `complexity` and runs ef = max(64, top_k)
    exactly as hnsw.rs:49,83 do; `search_batch`/`search_device` take a real ef."""

    backend = BACKEND_HNSW

    @classmethod
    def load(cls, index_path: str, dimensions: int, device: int = 0, metric: int = METRIC_DEFAULT):
        return cls._open(index_path, BACKEND_HNSW, dimensions, metric, device)

    def _effective_ef(self, complexity: int) -> int:
        return 64  # expansion_search: 64 (hnsw.rs:49); `_complexity` unused (hnsw.rs:83)

    @classmethod
    def build(cls, embeddings, graph_degree: int = 32, complexity: int = 64, metric: int = METRIC_DEFAULT,
              seed: int = 1, device: int = 0):
        """hnsw::build_index (hnsw.rs:96-139) on the GPU; `embeddings` numpy [n,d] or CUDA tensor."""
        h = C.c_void_p()
        e = _err()
        if isinstance(embeddings, np.ndarray):
            x = np.ascontiguousarray(embeddings, dtype=np.float32)
            ptr, on_dev, n, d = _np_ptr(x), 0, x.shape[0], x.shape[1]
        else:
            assert embeddings.is_cuda and embeddings.is_contiguous()
            ptr, on_dev, n, d = C.c_void_p(embeddings.data_ptr()), 1, embeddings.shape[0], embeddings.shape[1]
            device = embeddings.device.index or 0
        _check(lib().leann_cuda_hnsw_build(ptr, on_dev, n, d, graph_degree, complexity, metric, seed, device,
                                           C.byref(h), e, 1024), e)
        return cls(h.value)


    def add(self, embeddings, start_id: int = None, complexity: int = 64, seed: int = 1):
        """hnsw::add_to_index (hnsw.rs:142-191) on the resident index: keys start_id .. start_id+m-1
        (default: continue from len(), as cli/update.rs:221-232 does with passage_count)."""
        e = _err()
        if start_id is None:
            start_id = len(self)
        if isinstance(embeddings, np.ndarray):
            x = np.ascontiguousarray(embeddings, dtype=np.float32)
            ptr, on_dev, m = _np_ptr(x), 0, x.shape[0]
        else:
            assert embeddings.is_cuda and embeddings.is_contiguous()
            ptr, on_dev, m = C.c_void_p(embeddings.data_ptr()), 1, embeddings.shape[0]
        _check(lib().leann_cuda_hnsw_add(self._h, ptr, on_dev, m, start_id, complexity, seed, e, 1024), e)
        return self


def add_to_index(embeddings, index_path: str, dimensions: int, start_id: int, device: int = 0) -> None:
    """hnsw::add_to_index (hnsw.rs:142-191): load `<index_path>.index`, append, save it back."""
    s = HnswSearcher.load(index_path, dimensions, device)
    try:
        s.add(embeddings, start_id, complexity=64)
        s.save(index_path)
    finally:
        s.close()


class DiskAnnSearcher(BackendSearcher):
    """src/backend/diskann.rs:12-66: beam = max(complexity, top_k) (:54), DistDot metric (:16)."""

    backend = BACKEND_VAMANA

    @classmethod
    def load(cls, index_path: str, dimensions: int = 0, device: int = 0, metric: int = METRIC_DEFAULT):
        return cls._open(index_path, BACKEND_VAMANA, dimensions, metric, device)

    @classmethod
    def build(cls, embeddings, graph_degree: int = 64, complexity: int = 100, alpha: float = 1.2,
              metric: int = METRIC_DEFAULT, seed: int = 1, device: int = 0):
        h = C.c_void_p()
        e = _err()
        if isinstance(embeddings, np.ndarray):
            x = np.ascontiguousarray(embeddings, dtype=np.float32)
            ptr, on_dev, n, d = _np_ptr(x), 0, x.shape[0], x.shape[1]
        else:
            ptr, on_dev, n, d = C.c_void_p(embeddings.data_ptr()), 1, embeddings.shape[0], embeddings.shape[1]
            device = embeddings.device.index or 0
        _check(lib().leann_cuda_vamana_build(ptr, on_dev, n, d, graph_degree, complexity, alpha, metric, seed, device,
                                             C.byref(h), e, 1024), e)
        return cls(h.value)


class FlatSearcher(BackendSearcher):
    """Exact scan with RecomputeSearcher's scoring (src/index/recompute.rs:96-110): raw dot,
    descending, ties by ascending index."""

    backend = BACKEND_FLAT

    @classmethod
    def load(cls, index_path: str, dimensions: int, device: int = 0, metric: int = METRIC_DEFAULT):
        return cls._open(index_path, BACKEND_FLAT, dimensions, metric, device)

    @classmethod
    def from_vectors(cls, vectors, metric: int = METRIC_DOT_DESC, device: int = 0):
        h = C.c_void_p()
        e = _err()
        if isinstance(vectors, np.ndarray):
            x = np.ascontiguousarray(vectors, dtype=np.float32)
            _check(lib().leann_cuda_flat_from_host(_np_ptr(x), x.shape[0], x.shape[1], metric, device, C.byref(h), e, 1024), e)
        else:
            assert vectors.is_cuda and vectors.is_contiguous()
            device = vectors.device.index or 0
            _check(lib().leann_cuda_flat_from_device(C.c_void_p(vectors.data_ptr()), vectors.shape[0], vectors.shape[1],
                                                     metric, device, C.byref(h), e, 1024), e)
        return cls(h.value)


EXCHANGE_AUTO, EXCHANGE_NCCL, EXCHANGE_PEER = 0, 1, 2


class ShardedBackend:
    """A BackendSearcher (src/backend/traits.rs:11-30) over sub-indexes on several GPUs (`leann_cuda_shards_*`):
    every shard searches all queries, the per-shard lists are exchanged (NCCL all_gather, or peer-memory loads inside the
    merge kernel when one process owns all devices) and merged per query on the device. Keys are global."""

    def __init__(self, handle: int, parts=()):
        self._h = C.c_void_p(handle)
        self._parts = list(parts)   # keeps the shard searchers alive when the handle does not own them

    @classmethod
    def open(cls, base_paths: Sequence[str], backend: int, dimensions: int, devices: Sequence[int],
             metric: int = METRIC_DEFAULT, key_offsets=None, exchange: int = EXCHANGE_AUTO):
        paths = [os.fsencode(p) for p in base_paths]
        arr = (C.c_char_p * len(paths))(*paths)   # NUL-terminated C strings (no length array in this entry point)
        devs = (C.c_int * len(devices))(*devices)
        offs = None if key_offsets is None else (C.c_uint64 * len(key_offsets))(*key_offsets)
        h, e = C.c_void_p(), _err()
        _check(lib().leann_cuda_shards_open(arr, len(base_paths), backend, dimensions, metric, devs, offs, exchange,
                                            C.byref(h), e, 1024), e)
        return cls(h.value)

    @classmethod
    def from_searchers(cls, parts: Sequence[BackendSearcher], key_offsets=None, exchange: int = EXCHANGE_AUTO):
        """One process, one already built/opened searcher per device."""
        hs = (C.c_void_p * len(parts))(*[p._h for p in parts])
        offs = None if key_offsets is None else (C.c_uint64 * len(key_offsets))(*key_offsets)
        h, e = C.c_void_p(), _err()
        _check(lib().leann_cuda_shards_from_indexes(hs, len(parts), offs, 0, exchange, C.byref(h), e, 1024), e)
        return cls(h.value, parts)

    @staticmethod
    def unique_id() -> bytes:
        buf, e = C.create_string_buffer(128), _err()
        _check(lib().leann_cuda_comm_unique_id(buf, 128, e, 1024), e)
        return buf.raw

    @classmethod
    def join(cls, local: BackendSearcher, unique_id: bytes, rank: int, n_ranks: int, key_offset: int):
        """One process per GPU: every rank joins with its own shard (ncclCommInitRank inside the library)."""
        h, e = C.c_void_p(), _err()
        idb = C.create_string_buffer(bytes(unique_id), 128)
        _check(lib().leann_cuda_shards_join(local._h, 0, idb, 128, rank, n_ranks, key_offset, C.byref(h), e, 1024), e)
        return cls(h.value, [local])

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().leann_cuda_shards_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(lib().leann_cuda_shards_len(self._h))

    def len(self) -> int:
        return len(self)

    def is_empty(self) -> bool:
        return len(self) == 0

    def info(self) -> dict:
        out = (C.c_uint64 * 4)()
        lib().leann_cuda_shards_info(self._h, out)
        return {"shards": int(out[0]), "exchange": {0: "none", 1: "nccl_all_gather", 2: "peer_memory_merge"}[int(out[1])],
                "exchanges": int(out[2]), "exchange_bytes": int(out[3])}

    def search(self, query, top_k: int, complexity: int = 64):
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(1, -1)
        keys, dists, counts = self.search_batch(q, top_k, complexity)
        c = int(counts[0])
        return [int(x) for x in keys[0, :c]], [float(x) for x in dists[0, :c]]

    def search_batch(self, queries, top_k: int, ef: int, shard_masks=None):
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        keys = np.empty((nq, top_k), dtype=np.uint64)
        dists = np.empty((nq, top_k), dtype=np.float32)
        counts = np.zeros(nq, dtype=np.uint32)
        mp = None
        if shard_masks is not None:
            ms = [None if m is None else np.ascontiguousarray(m, dtype=np.uint64) for m in shard_masks]
            mp = (C.c_void_p * len(ms))(*[None if m is None else m.ctypes.data for m in ms])
        e = _err()
        _check(lib().leann_cuda_shards_search(self._h, _np_ptr(q), nq, top_k, ef, mp, _np_ptr(keys), _np_ptr(dists),
                                              _np_ptr(counts), e, 1024), e)
        return keys, dists, counts

    def hybrid_search(self, bm25_shard, queries, query_texts, top_k: int, ef: int, hybrid: bool, alpha: float, filter_mask=None):
        """search_with_options (index/searcher.rs:123-210) over document-range shards, one process per GPU: vector candidates
        from this sharded backend, BM25 from `bm25_shard` (this rank's documents, corpus-wide statistics), one all_gather,
        fusion on the device. Host arrays in/out; every rank gets the same answer."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        nq = q.shape[0]
        idx = np.empty((nq, top_k), dtype=np.uint64)
        sc = np.empty((nq, top_k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        arr = lens = None
        if query_texts is not None:
            keep, arr, lens = _strs(query_texts)
        m = None if filter_mask is None else np.ascontiguousarray(filter_mask, dtype=np.uint64)
        e = _err()
        _check(lib().leann_cuda_shards_hybrid_search(self._h, None if bm25_shard is None else bm25_shard._h, _np_ptr(q), arr, lens, nq, top_k, ef,
                                                     1 if hybrid else 0, C.c_float(alpha), None if m is None else _np_ptr(m),
                                                     0 if m is None else m.shape[0] * 64, _np_ptr(idx), _np_ptr(sc), _np_ptr(cnt), e, 1024), e)
        return idx, sc, cnt

    def search_device(self, queries, top_k: int, ef: int, mask=None, out=None, stream=None):
        """torch CUDA tensors on the local shard's device; search + exchange + merge on torch's current stream."""
        import torch

        assert queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous()
        nq = queries.shape[0]
        if out is None:
            keys = torch.empty((nq, top_k), dtype=torch.int64, device=queries.device)
            dists = torch.empty((nq, top_k), dtype=torch.float32, device=queries.device)
            counts = torch.empty((nq,), dtype=torch.int32, device=queries.device)
        else:
            keys, dists, counts = out
        st = stream if stream is not None else torch.cuda.current_stream(queries.device).cuda_stream
        e = _err()
        _check(lib().leann_cuda_shards_search_device(
            self._h, C.c_void_p(queries.data_ptr()), nq, top_k, ef, None if mask is None else C.c_void_p(mask.data_ptr()),
            C.c_void_p(keys.data_ptr()), C.c_void_p(dists.data_ptr()), C.c_void_p(counts.data_ptr()), C.c_void_p(st), e, 1024), e)
        return keys, dists, counts


class BackendType:
    """src/backend/mod.rs:16-45."""

    Hnsw = "hnsw"
    DiskAnn = "diskann"

    @staticmethod
    def load_searcher(backend_name: str, index_path: str, dimensions: int, device: int = 0) -> BackendSearcher:
        if backend_name == "hnsw":
            return HnswSearcher.load(index_path, dimensions, device)
        if backend_name == "diskann":
            return DiskAnnSearcher.load(index_path, dimensions, device)
        if backend_name == "flat":
            return FlatSearcher.load(index_path, dimensions, device)
        raise LeannCudaError(ERR_INVALID_ARG, f"Unknown backend: {backend_name}")  # searcher.rs:98


def check_index_file(base_path: str, backend: int, dimensions: int = 0) -> dict:
    """Host-only validation of `<base>.index` / `.diskann` / `.embeddings` (no GPU needed): the reader of leann_cuda_open."""
    out = (C.c_uint64 * 8)()
    e = _err()
    _check(lib().leann_cuda_check_index_file(os.fsencode(base_path), backend, dimensions, out, e, 1024), e)
    names = ["n", "dims", "M", "M0", "max_level", "entry", "n_upper_lists", "adjacency_hash"]
    return dict(zip(names, [int(x) for x in out]))


def topk_merge_device(keys_in, dists_in, descending: bool = False):
    """keys_in/dists_in: CUDA tensors [n_shards, nq, k] (an all_gather result) -> merged [nq, k]."""
    import torch

    g, nq, k = keys_in.shape
    keys = torch.empty((nq, k), dtype=torch.int64, device=keys_in.device)
    dists = torch.empty((nq, k), dtype=torch.float32, device=keys_in.device)
    counts = torch.empty((nq,), dtype=torch.int32, device=keys_in.device)
    e = _err()
    st = torch.cuda.current_stream(keys_in.device).cuda_stream
    _check(lib().leann_cuda_topk_merge_device(C.c_void_p(keys_in.data_ptr()), C.c_void_p(dists_in.data_ptr()), g, nq, k,
                                              1 if descending else 0, C.c_void_p(keys.data_ptr()),
                                              C.c_void_p(dists.data_ptr()), C.c_void_p(counts.data_ptr()),
                                              C.c_void_p(st), e, 1024), e)
    return keys, dists, counts


from . import text  # noqa: E402,F401
from .text import Bm25Scorer, MetadataFilter, MetadataColumns, IndexSearcher, SearchOptions, SearchResult, hybrid_rerank, hybrid_fuse, tokenize  # noqa: E402,F401
