"""world_size-2 gloo test (CPU) of the multi-GPU host logic in leann_rs_b200/shards.py: the sharded
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


# ---- hybrid search over document-range shards ------------------------------------------
def _hy_corpus(n, seed=5, vocab=120):
    rng = np.random.default_rng(seed)
    words = np.array([f"w{i}" for i in range(vocab)])
    p = 1.0 / np.arange(1, vocab + 1) ** 1.1
    p /= p.sum()
    docs = [" ".join(words[rng.choice(vocab, size=int(rng.integers(3, 20)), p=p)]) for _ in range(n)]
    queries = [" ".join(words[rng.choice(vocab, size=int(rng.integers(1, 4)), p=p)]) for _ in range(24)]
    x = rng.standard_normal((n, 16)).astype(np.float32)
    q = rng.standard_normal((len(queries), 16)).astype(np.float32)
    passes = rng.random(n) < 0.5
    return docs, queries, x, q, passes


def _hy_worker(rank, world, port, n, k, out_dir):
    sys.path.insert(0, ROOT)
    from leann_rs_b200 import shards as S
    from oracle import text_oracle as T
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    docs, queries, x, q, passes = _hy_corpus(n)
    lo, hi = S.shard_bounds(n, world, rank)
    ref = T.Bm25Scorer(docs)          # corpus-wide statistics: what build_sharded gives every shard

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

    def bm25_search_shard(texts, kk, doc_offset, cand_idx, cand_cnt):      # stand-in for Bm25Scorer.search_shard
        nq = len(texts)
        ti = np.full((nq, kk), np.uint64(2**64 - 1), dtype=np.uint64)
        ts = np.zeros((nq, kk), dtype=np.float32)
        tc = np.zeros(nq, dtype=np.uint32)
        cb = np.zeros(cand_idx.shape, dtype=np.float32)
        bx, bn = np.zeros(nq, dtype=np.float32), np.zeros(nq, dtype=np.float32)
        for i, t in enumerate(texts):
            dense = ref.score_query_fast(t)[lo:hi]
            pos = np.nonzero(dense > 0)[0]
            order = pos[np.argsort(-dense[pos], kind="stable")][:kk]
            ti[i, :len(order)] = order + doc_offset
            ts[i, :len(order)] = dense[order]
            tc[i] = len(order)
            for j in range(int(cand_cnt[i])):
                g = int(cand_idx[i, j])
                if lo <= g < hi:
                    cb[i, j] = dense[g - lo]
            bx[i], bn[i] = (dense.max(), dense.min()) if dense.size else (-np.inf, np.inf)
        return ti, ts, tc, cb, bx, bn

    def fuse(vk, vd, vc, top_k, hybrid, alpha, cb, ti, ts, tc, bx, bn, mask, mask_bits):   # stand-in for text.hybrid_fuse
        out = []
        for i in range(vk.shape[0]):
            res = [(int(vk[i, j]), vd[i, j]) for j in range(int(vc[i]))]
            if hybrid:
                dense = np.zeros(n, dtype=np.float32)
                for j in range(int(vc[i])):
                    dense[int(vk[i, j])] = cb[i, j]
                for j in range(int(tc[i])):
                    dense[int(ti[i, j])] = ts[i, j]
                assert dense.max() == bx[i] and dense.min() == bn[i]
                seen = {a for a, _ in res}
                res += [(int(ti[i, j]), np.float32(0.0)) for j in range(int(tc[i])) if int(ti[i, j]) not in seen]
                res = T.hybrid_rerank(res, dense, alpha)
            out.append([(a, float(b)) for a, b in res if mask is None or mask[a]][:top_k])
        return out

    vec = S.ShardedSearcher(local_shard, lo, world, rank, False, merge, dist)
    hy = S.ShardedHybridSearcher(vec, bm25_search_shard, fuse, lambda gk, gs: merge(gk, gs, True), lo, world, rank, dist)
    got = {}
    for name, hybrid, use_mask in (("hybrid", True, False), ("hybrid_filter", True, True), ("filter", False, True), ("plain", False, False)):
        got[name] = hy.search(torch.from_numpy(q), queries, k, 0, hybrid, 0.5, passes if use_mask else None, n)
    import pickle
    pickle.dump(got, open(os.path.join(out_dir, f"hy{rank}.pkl"), "wb"))
    dist.destroy_process_group()


def test_sharded_hybrid_equals_single_process_oracle(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import text_oracle as T
    import pickle
    world, k, n = 2, 5, 333
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_hy_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    docs, queries, x, q, passes = _hy_corpus(n)
    ref = T.Bm25Scorer(docs)
    for name, hybrid, use_mask in (("hybrid", True, False), ("hybrid_filter", True, True), ("filter", False, True), ("plain", False, False)):
        fk = T.fetch_k(k, use_mask, hybrid)
        gi, gd = _brute(x, q, fk)
        want = [T.search_with_options(gi[i].tolist(), gd[i].tolist(), k, ref, queries[i], hybrid, 0.5,
                                      (lambda j: bool(passes[j])) if use_mask else (lambda j: True), fk, fast=True) for i in range(len(queries))]
        for r in range(world):
            got = pickle.load(open(os.path.join(str(tmp_path), f"hy{r}.pkl"), "rb"))[name]
            for i in range(len(queries)):
                assert [a for a, _ in got[i]] == [a for a, _ in want[i]], (name, r, i)
                assert np.allclose([b for _, b in got[i]], [float(b) for _, b in want[i]], rtol=0, atol=0), (name, r, i)
