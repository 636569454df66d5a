"""Multi-GPU layouts of the search path: one process per GPU, `torch.distributed`
for the plumbing (NCCL on GPUs; gloo in the CPU tests of this host logic).

* ``ShardedSearcher`` — the database and graph are split into per-rank sub-indexes (independent
  graphs, own entry points); every rank searches ALL queries in its shard, the per-shard top-k lists
  are exchanged with one all_gather and merged per query by the K4 kernel
  (`leann_cuda_topk_merge_device`). Keys are made global by adding the shard's row offset.
* ``ReplicaSearcher`` — every rank holds the whole index; the query batch is split across ranks and
  the results are all_gathered (no data-path collective, the layout that scales QPS when the index
  fits one GPU: 1M x 768 is 3.4 GB of a B200's 180 GB).

* ``ShardedHybridSearcher`` — search_with_options (index/searcher.rs:123-210) over document-range shards:
  vector candidates from a ``ShardedSearcher``, BM25 from per-shard inverted indexes built with the
  corpus-wide statistics (`Bm25Scorer.build_sharded`), one exchange step — all_gather of the per-shard
  BM25 top lists, all_reduce of the candidates' BM25 scores (sum: only the owner is non-zero) and of
  max / min of the dense score vector (bm25.rs:152-153 needs them over the WHOLE corpus) — then the fusion
  kernel on every rank.

All take the local backend and the merge / fuse functions by injection so the host logic is testable on CPU.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; the first n % world shards hold one extra row."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def numpy_topk_merge(keys, dists, descending=False):
    """Reference merge for tests: keys/dists [G, nq, k] -> [nq, k]; ties by (shard, rank) order."""
    g, nq, k = keys.shape
    out_k = np.empty((nq, k), dtype=keys.dtype)
    out_d = np.empty((nq, k), dtype=dists.dtype)
    for i in range(nq):
        fd = dists[:, i, :].reshape(-1)
        fk = keys[:, i, :].reshape(-1)
        order = np.argsort(-fd if descending else fd, kind="stable")[:k]
        out_k[i], out_d[i] = fk[order], fd[order]
    return out_k, out_d


class ShardedSearcher:
    def __init__(self, local_search: Callable, row_offset: int, world: int, rank: int, descending: bool = False,
                 merge_fn: Callable = None, dist_module=None, group=None):
        self.local_search = local_search  # (queries, k, ef) -> (keys[nq,k] int64, dists[nq,k] f32) torch tensors
        self.row_offset, self.world, self.rank = int(row_offset), world, rank
        self.descending = descending
        self.merge_fn = merge_fn
        self.dist = dist_module
        self.group = group

    def search(self, queries, k: int, ef: int):
        import torch

        keys, dists = self.local_search(queries, k, ef)
        invalid = keys < 0  # UINT64_MAX read as int64
        keys = torch.where(invalid, keys, keys + self.row_offset)
        if self.world == 1:
            return keys, dists
        nq, k = keys.shape
        gk = torch.empty((self.world * nq, k), dtype=keys.dtype, device=keys.device)
        gd = torch.empty((self.world * nq, k), dtype=dists.dtype, device=dists.device)
        self.dist.all_gather_into_tensor(gk, keys.contiguous(), group=self.group)
        self.dist.all_gather_into_tensor(gd, dists.contiguous(), group=self.group)
        return self.merge_fn(gk.view(self.world, nq, k), gd.view(self.world, nq, k), self.descending)


class ReplicaSearcher:
    def __init__(self, local_search: Callable, world: int, rank: int, dist_module=None, group=None):
        self.local_search, self.world, self.rank = local_search, world, rank
        self.dist, self.group = dist_module, group

    def search(self, queries, k: int, ef: int, gather: bool = True):
        """`queries` is the GLOBAL batch (same on every rank); this rank answers its slice."""
        import torch

        nq = queries.shape[0]
        lo, hi = shard_bounds(nq, self.world, self.rank)
        keys, dists = self.local_search(queries[lo:hi].contiguous(), k, ef)
        if self.world == 1 or not gather:
            return keys, dists
        per = -(-nq // self.world)
        pk = torch.full((per, k), -1, dtype=keys.dtype, device=keys.device)
        pd = torch.full((per, k), float("inf"), dtype=dists.dtype, device=dists.device)
        pk[: hi - lo], pd[: hi - lo] = keys, dists
        gk = torch.empty((self.world * per, k), dtype=keys.dtype, device=keys.device)
        gd = torch.empty((self.world * per, k), dtype=dists.dtype, device=dists.device)
        self.dist.all_gather_into_tensor(gk, pk, group=self.group)
        self.dist.all_gather_into_tensor(gd, pd, group=self.group)
        gk, gd = gk.view(self.world, per, k), gd.view(self.world, per, k)
        outk, outd = [], []
        for r in range(self.world):
            a, b = shard_bounds(nq, self.world, r)
            outk.append(gk[r, : b - a])
            outd.append(gd[r, : b - a])
        return torch.cat(outk), torch.cat(outd)


class ShardedHybridSearcher:
    """`bm25_search_shard(texts, k, doc_offset, cand_idx, cand_cnt) -> (top_idx[nq,k] uint64 global ids (~0 = none),
    top_score, top_cnt, cand_bm[nq,fk], bmax[nq], bmin[nq])` (Bm25Scorer.search_shard);
    `fuse(vkeys, vdists, vcnt, top_k, hybrid, alpha, cand_bm, bm_idx, bm_score, bm_cnt, bmax, bmin, mask, mask_bits)`
    (text.hybrid_fuse); `merge_desc(keys[G,nq,k] int64, scores[G,nq,k]) -> (keys[nq,k], scores[nq,k])` descending,
    ties by (shard, rank) = ascending global id for contiguous shards."""

    def __init__(self, vector: ShardedSearcher, bm25_search_shard: Callable, fuse: Callable, merge_desc: Callable,
                 row_offset: int, world: int, rank: int, dist_module=None, group=None, exchange_device=None):
        self.vector, self.bm25_search_shard, self.fuse, self.merge_desc = vector, bm25_search_shard, fuse, merge_desc
        self.row_offset, self.world, self.rank = int(row_offset), world, rank
        self.dist, self.group, self.xdev = dist_module, group, exchange_device

    def search(self, queries, texts, top_k: int, ef: int, hybrid: bool, alpha: float, filter_mask=None, mask_bits: int = 0):
        import torch

        fk = top_k * 5 if (filter_mask is not None or hybrid) else top_k      # searcher.rs:129-133
        if hybrid and texts is None:
            hybrid = False                                                     # searcher.rs:147
        keys, dists = self.vector.search(queries, fk, ef)                      # merged over shards, identical on every rank
        vk = keys.cpu().numpy().astype(np.int64)
        vd = dists.cpu().numpy().astype(np.float32)
        vc = (vk >= 0).sum(axis=1).astype(np.uint32)
        cb = ti = ts = tc = bx = bn = None
        if hybrid:
            ti, ts, tc, cb, bx, bn = self.bm25_search_shard(texts, fk, self.row_offset, vk.view(np.uint64), vc)
            if self.world > 1:
                dev = self.xdev if self.xdev is not None else "cpu"
                t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt) if a.dtype == np.uint64 else np.ascontiguousarray(a)).to(dev)
                lk, ls = t(ti, np.int64), t(ts, np.float32)
                nq = lk.shape[0]
                gk = torch.empty((self.world * nq, fk), dtype=lk.dtype, device=lk.device)
                gs = torch.empty((self.world * nq, fk), dtype=ls.dtype, device=ls.device)
                self.dist.all_gather_into_tensor(gk, lk, group=self.group)
                self.dist.all_gather_into_tensor(gs, ls, group=self.group)
                mk, ms = self.merge_desc(gk.view(self.world, nq, fk), gs.view(self.world, nq, fk))
                tcb, tbx, tbn = t(cb, np.float32), t(bx, np.float32), t(bn, np.float32)
                self.dist.all_reduce(tcb, op=self.dist.ReduceOp.SUM, group=self.group)    # x + 0 + ... + 0: exact
                self.dist.all_reduce(tbx, op=self.dist.ReduceOp.MAX, group=self.group)
                self.dist.all_reduce(tbn, op=self.dist.ReduceOp.MIN, group=self.group)
                ti = mk.cpu().numpy().astype(np.int64)
                ts = ms.cpu().numpy().astype(np.float32)
                tc = (ti >= 0).sum(axis=1).astype(np.uint32)
                ti = ti.view(np.uint64)
                cb, bx, bn = tcb.cpu().numpy(), tbx.cpu().numpy(), tbn.cpu().numpy()
        return self.fuse(vk.view(np.uint64), vd, vc, top_k, hybrid, alpha, cb, ti, ts, tc, bx, bn, filter_mask, mask_bits)
