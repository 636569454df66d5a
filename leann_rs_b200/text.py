"""Host mirror of src/index/{bm25,filter,searcher}.rs over the C ABI (text path).

Names and argument meaning follow the reference: `Bm25Scorer::{build,score_query,search}`,
`tokenize`, `hybrid_rerank`, `MetadataFilter::{parse,matches}`, `SearchOptions`,
`IndexSearcher::{load,search,search_with_options,bm25_search}`. All arithmetic runs in
libleann_cuda.so (BM25 / fusion on the GPU, filter parsing and evaluation on the host side of the
library); nothing here computes.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np


def _core():
    import leann_rs_b200 as P
    return P


def _err():
    return C.create_string_buffer(1024)


def _check(code, e):
    P = _core()
    if code != 0:
        raise P.LeannCudaError(code, e.value.decode(errors="replace"))


def _strs(items: Sequence):
    """Python strings / bytes -> (keep-alive, `const char* const*`, `const size_t*`) for the C ABI. One joined buffer plus a
    vectorised pointer table: building a ctypes array element by element cost ~10 ms per 10k query texts, which sat inside
    the end-to-end time of every hybrid batch."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in items]
    n = len(bs)
    lens = np.fromiter(map(len, bs), dtype=np.uint64, count=n) if n else np.zeros(1, dtype=np.uint64)
    blob = C.create_string_buffer(b"".join(bs) + b"\0")
    base = C.addressof(blob)
    ptrs = np.zeros(max(n, 1), dtype=np.uint64)
    if n:
        ptrs[0] = base
        if n > 1:
            np.cumsum(lens[:-1], out=ptrs[1:])
            ptrs[1:] += np.uint64(base)
    arr = C.cast(C.c_void_p(ptrs.ctypes.data), C.POINTER(C.c_char_p))
    lens_p = C.cast(C.c_void_p(lens.ctypes.data), C.POINTER(C.c_size_t))
    return (bs, blob, ptrs, lens), arr, lens_p


def tokenize(text: str) -> List[str]:
    """index/bm25.rs:127-132."""
    L = _core().lib()
    b = text.encode()
    cap = len(b) + 2
    out = C.create_string_buffer(cap)
    n = L.leann_cuda_tokenize(b, len(b), out, cap)
    return out.value.decode().split("\n")[:n] if n else []


class Bm25Scorer:
    """index/bm25.rs:17-122 with the inverted index resident in HBM."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    @classmethod
    def build(cls, documents: Sequence[str], device: int = 0) -> "Bm25Scorer":
        L = _core().lib()
        keep, arr, lens = _strs(documents)
        h = C.c_void_p()
        e = _err()
        _check(L.leann_cuda_bm25_build(arr, lens, len(documents), device, C.byref(h), e, 1024), e)
        return cls(h.value)

    # ---- document-range shards: corpus-wide N / token count / df travel as opaque blobs ----
    @staticmethod
    def shard_stats(documents: Sequence[str], device: int = 0) -> bytes:
        L = _core().lib()
        keep, arr, lens = _strs(documents)
        need = C.c_size_t()
        e = _err()
        _check(L.leann_cuda_bm25_shard_stats(arr, lens, len(documents), device, None, 0, C.byref(need), e, 1024), e)
        buf = (C.c_ubyte * max(need.value, 1))()
        _check(L.leann_cuda_bm25_shard_stats(arr, lens, len(documents), device, buf, need.value, C.byref(need), e, 1024), e)
        return bytes(buf[: need.value])

    @staticmethod
    def merge_stats(blobs: Sequence[bytes]) -> bytes:
        L = _core().lib()
        n = len(blobs)
        bufs = [(C.c_ubyte * max(len(b), 1)).from_buffer_copy(b if b else b"\0") for b in blobs]
        arr = (C.POINTER(C.c_ubyte) * max(n, 1))(*[C.cast(b, C.POINTER(C.c_ubyte)) for b in bufs])
        lens = (C.c_size_t * max(n, 1))(*[len(b) for b in blobs])
        need = C.c_size_t()
        e = _err()
        _check(L.leann_cuda_bm25_stats_merge(arr, lens, n, None, 0, C.byref(need), e, 1024), e)
        out = (C.c_ubyte * max(need.value, 1))()
        _check(L.leann_cuda_bm25_stats_merge(arr, lens, n, out, need.value, C.byref(need), e, 1024), e)
        return bytes(out[: need.value])

    @classmethod
    def build_sharded(cls, documents: Sequence[str], global_stats: bytes, device: int = 0) -> "Bm25Scorer":
        """This rank's documents indexed with the corpus-wide statistics: scores equal the unsharded index bit for bit."""
        L = _core().lib()
        keep, arr, lens = _strs(documents)
        st = (C.c_ubyte * max(len(global_stats), 1)).from_buffer_copy(global_stats or b"\0")
        h = C.c_void_p()
        e = _err()
        _check(L.leann_cuda_bm25_build_sharded(arr, lens, len(documents), st, len(global_stats), device, C.byref(h), e, 1024), e)
        return cls(h.value)

    def search_shard(self, queries: Sequence[str], top_k: int, doc_offset: int, cand_idx=None, cand_cnt=None):
        """One shard's BM25 part of the hybrid step. Returns (top_idx[nq,k] global ids, top_score, top_cnt, cand_bm or None,
        bmax[nq], bmin[nq]); reduce over shards with top-k merge / sum / max / min."""
        L = _core().lib()
        keep, arr, lens = _strs(queries)
        nq = len(queries)
        ti = np.empty((nq, top_k), dtype=np.uint64)
        ts = np.empty((nq, top_k), dtype=np.float32)
        tc = np.zeros(nq, dtype=np.uint32)
        bx = np.zeros(nq, dtype=np.float32)
        bn = np.zeros(nq, dtype=np.float32)
        ci = cc = cb = None
        fk = 0
        if cand_idx is not None:
            ci = np.ascontiguousarray(cand_idx, dtype=np.uint64)
            cc = np.ascontiguousarray(cand_cnt, dtype=np.uint32)
            fk = ci.shape[1]
            cb = np.zeros((nq, fk), dtype=np.float32)
        ptr = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        e = _err()
        _check(L.leann_cuda_bm25_search_shard(self._h, arr, lens, nq, top_k, doc_offset, ptr(ci), ptr(cc), fk, ptr(ti), ptr(ts), ptr(tc),
                                              ptr(cb), ptr(bx), ptr(bn), e, 1024), e)
        return ti, ts, tc, cb, bx, bn

    def __len__(self):
        return int(_core().lib().leann_cuda_bm25_len(self._h))

    def dense_rows(self) -> int:
        """Terms also kept as dense score rows (K3d); LEANN_CUDA_BM25_DENSE_FRAC / _MAX at build time."""
        return int(_core().lib().leann_cuda_bm25_dense_rows(self._h))

    def stats(self) -> dict:
        st = (C.c_uint64 * 4)()
        avg = C.c_float()
        _core().lib().leann_cuda_bm25_stats(self._h, st, C.byref(avg))
        return {"num_docs": int(st[0]), "n_terms": int(st[1]), "n_postings": int(st[2]), "total_tokens": int(st[3]),
                "avg_doc_len": float(avg.value)}

    def score_query(self, query: str) -> np.ndarray:
        """bm25.rs:77-106 -> dense f32[num_docs]."""
        L = _core().lib()
        out = np.zeros(len(self), dtype=np.float32)
        b = query.encode()
        e = _err()
        _check(L.leann_cuda_bm25_score(self._h, b, len(b), C.c_void_p(out.ctypes.data), e, 1024), e)
        return out

    def search(self, query: str, top_k: int) -> List[Tuple[int, float]]:
        """bm25.rs:109-122 for one query."""
        idx, sc, cnt = self.search_batch([query], top_k)
        return [(int(idx[0, j]), float(sc[0, j])) for j in range(int(cnt[0]))]

    def search_batch(self, queries: Sequence[str], top_k: int):
        L = _core().lib()
        keep, arr, lens = _strs(queries)
        nq = len(queries)
        idx = np.empty((nq, top_k), dtype=np.uint64)
        sc = np.empty((nq, top_k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        e = _err()
        _check(L.leann_cuda_bm25_search(self._h, arr, lens, nq, top_k, C.c_void_p(idx.ctypes.data), C.c_void_p(sc.ctypes.data),
                                        C.c_void_p(cnt.ctypes.data), e, 1024), e)
        return idx, sc, cnt

    def last_batch_bytes(self) -> int:
        """Bytes the last search_batch's tokens make the kernel stream (8 per posting, 4 per document of a dense row)."""
        return int(_core().lib().leann_cuda_bm25_last_batch_bytes(self._h))

    def last_batch(self):
        """(postings covered by the last search_batch's tokens, device ms of its query kernel)."""
        n, ms = C.c_uint64(), C.c_float()
        _core().lib().leann_cuda_bm25_last_batch(self._h, C.byref(n), C.byref(ms))
        return int(n.value), float(ms.value)

    def close(self):
        if getattr(self, "_h", None):
            _core().lib().leann_cuda_bm25_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hybrid_rerank(vector_results: Sequence[Tuple[int, float]], bm25_scores, alpha: float, device: int = 0):
    """index/bm25.rs:135-170 (device kernel): returns [(idx, combined)] stable-sorted descending."""
    L = _core().lib()
    n = len(vector_results)
    if n == 0:
        return []
    idx = np.asarray([i for i, _ in vector_results], dtype=np.uint64)
    vs = np.asarray([s for _, s in vector_results], dtype=np.float32)
    bm = np.ascontiguousarray(bm25_scores, dtype=np.float32)
    oi = np.empty(n, dtype=np.uint64)
    os_ = np.empty(n, dtype=np.float32)
    e = _err()
    _check(L.leann_cuda_hybrid_rerank(C.c_void_p(idx.ctypes.data), C.c_void_p(vs.ctypes.data), n, C.c_void_p(bm.ctypes.data),
                                      bm.shape[0], C.c_float(alpha), device, C.c_void_p(oi.ctypes.data),
                                      C.c_void_p(os_.ctypes.data), e, 1024), e)
    return [(int(a), float(b)) for a, b in zip(oi, os_)]


class MetadataFilter:
    """index/filter.rs:35-39, 52-134, 319-325."""

    def __init__(self, handle, expr):
        self._h = C.c_void_p(handle)
        self.expr = expr

    @classmethod
    def parse(cls, filter_str: str) -> Optional["MetadataFilter"]:
        """Returns None where the reference returns None."""
        P = _core()
        h = C.c_void_p()
        e = _err()
        rc = P.lib().leann_cuda_filter_parse(filter_str.encode(), C.byref(h), e, 1024)
        if rc == P.ERR_PARSE:
            return None
        _check(rc, e)
        return cls(h.value, filter_str)

    def describe(self) -> Any:
        L = _core().lib()
        n = L.leann_cuda_filter_describe(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        L.leann_cuda_filter_describe(self._h, buf, n + 1)
        return json.loads(buf.value.decode())

    def matches(self, metadata: Any) -> bool:
        L = _core().lib()
        b = json.dumps(metadata).encode()
        res = C.c_int()
        e = _err()
        _check(L.leann_cuda_filter_matches(self._h, b, len(b), C.byref(res), e, 1024), e)
        return bool(res.value)

    def mask(self, metadata_docs: Sequence[Any]) -> np.ndarray:
        """Evaluate once over all passages -> bitmask words for the kernels."""
        L = _core().lib()
        docs = [m if isinstance(m, (str, bytes)) else json.dumps(m) for m in metadata_docs]
        keep, arr, lens = _strs(docs)
        out = np.zeros((len(docs) + 63) // 64, dtype=np.uint64)
        e = _err()
        _check(L.leann_cuda_filter_mask(self._h, arr, lens, len(docs), C.c_void_p(out.ctypes.data), e, 1024), e)
        return out

    def close(self):
        if getattr(self, "_h", None):
            _core().lib().leann_cuda_filter_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MetadataColumns:
    """Columnar side-car of the passages' metadata: built once from one JSON document per
    passage (None = no metadata); `mask(filter)` evaluates a parsed MetadataFilter column-wise into the
    N-bit mask, bit-identical to MetadataFilter.mask (filter.rs:319-439 row by row)."""

    def __init__(self, metadata_docs: Sequence[Any]):
        L = _core().lib()
        n = len(metadata_docs)
        enc = [None if m is None else (m if isinstance(m, bytes) else (m if isinstance(m, str) else json.dumps(m)).encode())
               for m in metadata_docs]
        arr = (C.c_char_p * max(n, 1))(*enc) if n else (C.c_char_p * 1)()
        lens = (C.c_size_t * max(n, 1))(*[0 if b is None else len(b) for b in enc]) if n else (C.c_size_t * 1)()
        h = C.c_void_p()
        e = _err()
        _check(L.leann_cuda_metacols_build(arr, lens, n, C.byref(h), e, 1024), e)
        self._h, self.n = h, n

    @property
    def fields(self) -> int:
        return int(_core().lib().leann_cuda_metacols_fields(self._h))

    def mask(self, flt: "MetadataFilter") -> np.ndarray:
        out = np.zeros((self.n + 63) // 64, dtype=np.uint64)
        e = _err()
        _check(_core().lib().leann_cuda_metacols_mask(self._h, flt._h, C.c_void_p(out.ctypes.data), e, 1024), e)
        return out

    def close(self):
        if getattr(self, "_h", None):
            _core().lib().leann_cuda_metacols_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class SearchOptions:
    """index/searcher.rs:25-63."""
    top_k: int = 0
    complexity: int = 0
    filter: Optional[str] = None
    hybrid: bool = False
    hybrid_alpha: float = 0.0
    query_text: Optional[str] = None

    @classmethod
    def new(cls, top_k: int, complexity: int) -> "SearchOptions":
        return cls(top_k=top_k, complexity=complexity, hybrid_alpha=0.7)  # searcher.rs:47

    def with_filter(self, filter_str: str) -> "SearchOptions":
        self.filter = filter_str
        return self

    def with_hybrid(self, query_text: str, alpha: float) -> "SearchOptions":
        self.hybrid, self.hybrid_alpha, self.query_text = True, alpha, query_text
        return self


@dataclass
class SearchResult:
    """index/searcher.rs:15-21."""
    id: str
    score: float
    text: str
    metadata: Any = field(default=None)


def hybrid_search(index, bm25: Optional[Bm25Scorer], queries, query_texts: Optional[Sequence[str]], top_k: int, ef: int,
                  hybrid: bool, alpha: float, filter_mask: Optional[np.ndarray] = None):
    """Batched search_with_options core (C ABI leann_cuda_hybrid_search): host arrays in/out."""
    L = _core().lib()
    q = np.ascontiguousarray(queries, dtype=np.float32)
    nq = q.shape[0]
    idx = np.empty((nq, top_k), dtype=np.uint64)
    sc = np.empty((nq, top_k), dtype=np.float32)
    cnt = np.zeros(nq, dtype=np.uint32)
    if query_texts is not None:
        keep, arr, lens = _strs(query_texts)
    else:
        arr, lens = None, None
    m = None if filter_mask is None else np.ascontiguousarray(filter_mask, dtype=np.uint64)
    e = _err()
    _check(L.leann_cuda_hybrid_search(index._h, None if bm25 is None else bm25._h, C.c_void_p(q.ctypes.data), arr, lens, nq,
                                      top_k, ef, 1 if hybrid else 0, C.c_float(alpha),
                                      None if m is None else C.c_void_p(m.ctypes.data), C.c_void_p(idx.ctypes.data),
                                      C.c_void_p(sc.ctypes.data), C.c_void_p(cnt.ctypes.data), e, 1024), e)
    return idx, sc, cnt


def hybrid_fuse(vkeys, vdists, vcnt, top_k: int, hybrid: bool, alpha: float, cand_bm=None, bm_idx=None, bm_score=None, bm_cnt=None,
                bmax=None, bmin=None, filter_mask: Optional[np.ndarray] = None, mask_bits: int = 0, device: int = 0):
    """Batched hybrid_rerank + BM25-only additions + post-filter walk (bm25.rs:135-170, searcher.rs:156-207) over gathered
    inputs (C ABI leann_cuda_hybrid_fuse). vkeys/vdists [nq, fetch_k]; bm_* [nq, bm_k]."""
    L = _core().lib()
    vk = np.ascontiguousarray(vkeys, dtype=np.uint64)
    vd = np.ascontiguousarray(vdists, dtype=np.float32)
    vc = np.ascontiguousarray(vcnt, dtype=np.uint32)
    nq, fk = vk.shape
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=t) for a, t in
            ((cand_bm, np.float32), (bm_idx, np.uint64), (bm_score, np.float32), (bm_cnt, np.uint32), (bmax, np.float32), (bmin, np.float32))]
    m = None if filter_mask is None else np.ascontiguousarray(filter_mask, dtype=np.uint64)
    idx = np.empty((nq, top_k), dtype=np.uint64)
    sc = np.empty((nq, top_k), dtype=np.float32)
    cnt = np.zeros(nq, dtype=np.uint32)
    ptr = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
    e = _err()
    _check(L.leann_cuda_hybrid_fuse(ptr(vk), ptr(vd), ptr(vc), nq, fk, ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]), ptr(arrs[3]),
                                    0 if arrs[1] is None else arrs[1].shape[1], ptr(arrs[4]), ptr(arrs[5]), 1 if hybrid else 0,
                                    C.c_float(alpha), ptr(m), mask_bits if m is not None else 0, top_k, device, ptr(idx), ptr(sc), ptr(cnt),
                                    e, 1024), e)
    return idx, sc, cnt


class IndexSearcher:
    """index/searcher.rs:66-257. `load(index_path, backend_name, dimensions)` takes the two meta.json
    fields the reference reads (meta.backend_name, meta.dimensions)."""

    def __init__(self, handle, base):
        self._h = C.c_void_p(handle)
        self.base = base
        self._jsonl = None

    @classmethod
    def load(cls, index_path: str, backend_name: str, dimensions: int, device: int = 0) -> "IndexSearcher":
        L = _core().lib()
        h = C.c_void_p()
        e = _err()
        _check(L.leann_cuda_searcher_load(os.fsencode(index_path), backend_name.encode(), dimensions, device, C.byref(h), e, 1024), e)
        return cls(h.value, index_path)

    def __len__(self):
        return int(_core().lib().leann_cuda_searcher_len(self._h))

    def honor_complexity(self, on: bool = True):
        _core().lib().leann_cuda_searcher_set_honor_complexity(self._h, 1 if on else 0)

    def _id(self, idx: int) -> str:
        buf = C.create_string_buffer(512)
        _core().lib().leann_cuda_searcher_id(self._h, int(idx), buf, 512)
        return buf.value.decode()

    def search_batch(self, query_embeddings, opts: SearchOptions, query_texts: Optional[Sequence[str]] = None):
        L = _core().lib()
        q = np.ascontiguousarray(query_embeddings, dtype=np.float32)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        idx = np.empty((nq, opts.top_k), dtype=np.uint64)
        sc = np.empty((nq, opts.top_k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        texts = query_texts
        if texts is None and opts.hybrid and opts.query_text is not None:
            texts = [opts.query_text] * nq
        if texts is not None:
            keep, arr, lens = _strs(texts)
        else:
            arr, lens = None, None
        e = _err()
        _check(L.leann_cuda_searcher_search(self._h, C.c_void_p(q.ctypes.data), arr, lens, nq, opts.top_k, opts.complexity,
                                            None if not opts.filter else opts.filter.encode(), 1 if opts.hybrid else 0,
                                            C.c_float(opts.hybrid_alpha), C.c_void_p(idx.ctypes.data), C.c_void_p(sc.ctypes.data),
                                            C.c_void_p(cnt.ctypes.data), e, 1024), e)
        return idx, sc, cnt

    def _passage(self, pid: str):
        # PassageStore::get (passages.rs:90-105): text/metadata storage stays with the reference's files
        if self._jsonl is None:
            base = self.base.rsplit(".", 1)[0] if "." in os.path.basename(self.base) else self.base
            with open(base + ".passages.idx.json") as f:
                self._offsets = json.load(f)
            self._jsonl = open(base + ".passages.jsonl", "rb")
        self._jsonl.seek(self._offsets[pid])
        return json.loads(self._jsonl.readline())

    def search_with_options(self, query_embedding, opts: SearchOptions) -> List[SearchResult]:
        idx, sc, cnt = self.search_batch(query_embedding, opts)
        out = []
        for j in range(int(cnt[0])):
            pid = self._id(idx[0, j])
            p = self._passage(pid)
            out.append(SearchResult(pid, float(sc[0, j]), p.get("text", ""), p.get("metadata")))
        return out

    def search(self, query_embedding, top_k: int, complexity: int) -> List[SearchResult]:
        return self.search_with_options(query_embedding, SearchOptions.new(top_k, complexity))

    def bm25_search(self, query: str, top_k: int) -> List[str]:
        """searcher.rs:228-246: texts of the BM25 top-k passages."""
        L = _core().lib()
        idx = np.empty(top_k, dtype=np.uint64)
        sc = np.empty(top_k, dtype=np.float32)
        cnt = C.c_uint32()
        b = query.encode()
        e = _err()
        _check(L.leann_cuda_searcher_bm25_search(self._h, b, len(b), top_k, C.c_void_p(idx.ctypes.data),
                                                 C.c_void_p(sc.ctypes.data), C.byref(cnt), e, 1024), e)
        return [self._passage(self._id(idx[j])).get("text", "") for j in range(cnt.value)]

    def close(self):
        if getattr(self, "_h", None):
            _core().lib().leann_cuda_searcher_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
