"""Multi-GPU layouts of the search path: one process per GPU, `torch.distributed`
for the plumbing (NCCL on GPUs; gloo in the CPU tests of this host logic).

* ``ShardedSearcher`` — the database and graph are split into per-rank sub-indexes (independent
  graphs, own entry points); every rank searches ALL queries in its shard, the per-shard top-k lists
  are exchanged with one all_gather and merged per query by the K4 kernel
  (`leann_cuda_topk_merge_device`). Keys are made global by adding the shard's row offset.
* ``ReplicaSearcher`` — every rank holds the whole index; the query batch is split across ranks and
  the results are all_gathered (no data-path collective, the layout that scales QPS when the index
  fits one GPU: 1M x 768 is 3.4 GB of a B200's 180 GB).

Both take the local backend and a merge function by injection so the host logic is testable on CPU.
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
