"""world_size-2 gloo test (CPU) of the multi-GPU host logic in leann-rs_b200/shards.py: the sharded
layout (every rank searches all queries in its rows, all_gather, per-query merge, global keys) and the
replica layout (queries split, results all_gathered) must both equal a single-process search."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _brute(x, q, k):
    d = 1.0 - q @ x.T
    idx = np.argsort(d, axis=1, kind="stable")[:, :k]
    return idx.astype(np.int64), np.take_along_axis(d, idx, axis=1).astype(np.float32)


def _worker(rank, world, port, n, nq, k, out_dir):
    sys.path.insert(0, ROOT)
    from leann_rs_b200 import shards as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, 16)).astype(np.float32)
    q = rng.standard_normal((nq, 16)).astype(np.float32)
    lo, hi = S.shard_bounds(n, world, rank)

    def local_shard(qt, kk, ef):
        i, d = _brute(x[lo:hi], qt.numpy(), min(kk, hi - lo))
        pad = kk - i.shape[1]
        if pad:
            i = np.concatenate([i, -np.ones((i.shape[0], pad), dtype=np.int64)], axis=1)
            d = np.concatenate([d, np.full((d.shape[0], pad), np.inf, dtype=np.float32)], axis=1)
        return torch.from_numpy(i), torch.from_numpy(d)

    def merge(gk, gd, desc):
        a, b = S.numpy_topk_merge(gk.numpy(), gd.numpy(), desc)
        return torch.from_numpy(a), torch.from_numpy(b)

    sh = S.ShardedSearcher(local_shard, lo, world, rank, False, merge, dist)
    sk, sd = sh.search(torch.from_numpy(q), k, 0)

    def local_full(qt, kk, ef):
        i, d = _brute(x, qt.numpy(), kk)
        return torch.from_numpy(i), torch.from_numpy(d)

    rp = S.ReplicaSearcher(local_full, world, rank, dist)
    rk, rd = rp.search(torch.from_numpy(q), k, 0)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), sk=sk.numpy(), sd=sd.numpy(), rk=rk.numpy(), rd=rd.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("n,nq", [(1001, 37), (64, 5)])
def test_sharded_and_replica_layouts_equal_single_process(tmp_path, n, nq):
    world, k = 2, 10
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, n, nq, k, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((n, 16)).astype(np.float32)
    q = rng.standard_normal((nq, 16)).astype(np.float32)
    gi, gd = _brute(x, q, k)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert np.array_equal(z["sk"], gi) and np.allclose(z["sd"], gd)
        assert np.array_equal(z["rk"], gi) and np.allclose(z["rd"], gd)


def test_shard_bounds_cover_and_balance():
    sys.path.insert(0, ROOT)
    from leann_rs_b200 import shards as S
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [S.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
