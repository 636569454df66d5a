"""GPU tests of the sharded backend behind the C ABI (`leann_cuda_shards_*`): sub-index per GPU, exchange
(NCCL all_gather or peer-memory loads inside the merge kernel), per-query top-k merge. The merged answer must equal a
numpy merge of the per-shard answers bit for bit, and the unsharded search of the whole database up to exact-score ties."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import make_data

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def _n_gpus(pkg):
    return pkg.device_count()


def test_one_shard_handle_equals_backend(pkg):
    x, q = make_data(20000, 96, 3, nq=50)
    flat = pkg.FlatSearcher.from_vectors(x, metric=pkg.METRIC_DOT_DESC)
    k0, d0, c0 = flat.search_batch(q, 20, 0)
    sh = pkg.ShardedBackend.from_searchers([flat], key_offsets=[1000])
    assert len(sh) == 20000 and sh.info()["shards"] == 1 and sh.info()["exchange"] == "none"
    k1, d1, c1 = sh.search_batch(q, 20, 0)
    assert np.array_equal(k1, k0 + np.uint64(1000)) and np.array_equal(d1.view(np.uint32), d0.view(np.uint32)) and np.array_equal(c0, c1)
    ids, dd = sh.search(q[0], 5)
    assert ids == [int(v) + 1000 for v in k0[0, :5]]
    sh.close()
    # the process-per-GPU entry point with a world of one
    sj = pkg.ShardedBackend.join(flat, b"\0" * 128, 0, 1, 7)
    k2, d2, _ = sj.search_batch(q, 20, 0)
    assert np.array_equal(k2, k0 + np.uint64(7)) and np.array_equal(d2.view(np.uint32), d0.view(np.uint32))
    sj.close()
    flat.close()


def test_sharded_hybrid_world_of_one_equals_hybrid_search(pkg):
    """leann_cuda_shards_hybrid_search with a single shard runs the whole device-resident pipeline (candidate localisation,
    packed block, merge, reductions, fusion) and must reproduce leann_cuda_hybrid_search bit for bit."""
    rng = np.random.default_rng(3)
    nd, d, nq = 5000, 96, 64
    x, q = make_data(nd, d, 8, nq=nq)
    vocab = [f"tok{i}" for i in range(400)]
    docs = [" ".join(rng.choice(vocab, size=int(rng.integers(5, 50))).tolist()) for _ in range(nd)]
    texts = [" ".join(rng.choice(vocab[:80], size=int(rng.integers(1, 6))).tolist()) for _ in range(nq)]
    mask = pkg.pack_mask(rng.random(nd) < 0.3)
    idx = pkg.HnswSearcher.build(x, graph_degree=16, complexity=64, seed=2)
    bm = pkg.Bm25Scorer.build(docs)
    sh = pkg.ShardedBackend.join(idx, b"\0" * 128, 0, 1, 0)
    for hybrid, alpha, m in ((True, 0.5, mask), (True, 0.7, None), (False, 0.5, mask), (False, 0.5, None)):
        ri, rs, rc = pkg.text.hybrid_search(idx, bm, q, texts, 10, 64, hybrid, alpha, m)
        si, ss, sc = sh.hybrid_search(bm, q, texts, 10, 64, hybrid, alpha, m)
        assert np.array_equal(rc, sc) and np.array_equal(ri, si) and np.array_equal(rs.view(np.uint32), ss.view(np.uint32)), (hybrid, alpha)
    sh.close(); bm.close(); idx.close()


def test_sharded_errors(pkg):
    x, _ = make_data(3000, 64, 1)
    a = pkg.FlatSearcher.from_vectors(x, metric=pkg.METRIC_DOT_DESC)
    b = pkg.FlatSearcher.from_vectors(x, metric=pkg.METRIC_DOT_DESC)
    with pytest.raises(pkg.LeannCudaError) as ei:   # two shards on one device
        pkg.ShardedBackend.from_searchers([a, b])
    assert ei.value.code == pkg.ERR_INVALID_ARG
    with pytest.raises(pkg.LeannCudaError) as ei:
        pkg.ShardedBackend.open(["/nonexistent/a.leann"], pkg.BACKEND_HNSW, 64, [0])
    assert ei.value.code == pkg.ERR_NOT_FOUND
    a.close(); b.close()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_single_process_shards(pkg, exchange, tmp_path):
    from leann_rs_b200.shards import numpy_topk_merge, shard_bounds
    g = _n_gpus(pkg)
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(g, 4)
    mode = pkg.EXCHANGE_PEER if exchange == "peer" else pkg.EXCHANGE_NCCL
    n, d, k, nq = 40000, 128, 10, 200
    x, q = make_data(n, d, 5, nq=nq)
    bounds = [shard_bounds(n, world, r) for r in range(world)]
    # exact scan shards
    parts = [pkg.FlatSearcher.from_vectors(x[lo:hi], metric=pkg.METRIC_DOT_DESC, device=r) for r, (lo, hi) in enumerate(bounds)]
    try:
        sh = pkg.ShardedBackend.from_searchers(parts, exchange=mode)
    except pkg.LeannCudaError as e:
        if exchange == "peer" and e.code == pkg.ERR_CUDA:
            pytest.skip("no peer access between the devices of this box")
        raise
    assert sh.info()["exchange"] == ("peer_memory_merge" if exchange == "peer" else "nccl_all_gather")
    mk, md, mc = sh.search_batch(q, k, 0)
    per = [p.search_batch(q, k, 0) for p in parts]
    lk = np.stack([np.where(pk == NONE, pk, pk + np.uint64(lo)) for (pk, _, _), (lo, _) in zip(per, bounds)])
    ld = np.stack([pd for _, pd, _ in per])
    rk, rd = numpy_topk_merge(lk, ld, descending=True)
    assert np.array_equal(mk, rk) and np.array_equal(md.view(np.uint32), rd.view(np.uint32)) and (mc == k).all()
    whole = pkg.FlatSearcher.from_vectors(x, metric=pkg.METRIC_DOT_DESC, device=0)
    wk, wd, _ = whole.search_batch(q, k, 0)
    same = wk == mk
    assert same.mean() > 0.995 and np.allclose(wd, md, rtol=0, atol=2e-6)
    assert np.all(np.abs(wd[~same] - md[~same]) <= 1e-5 * np.abs(wd[~same]) + 1e-7)
    whole.close(); sh.close()
    for p in parts:
        p.close()
    # HNSW shards opened from files through leann_cuda_shards_open, with an inline mask on shard 1
    bases = []
    for r, (lo, hi) in enumerate(bounds):
        s = pkg.HnswSearcher.build(x[lo:hi], graph_degree=16, complexity=64, seed=9, device=r)
        base = str(tmp_path / f"shard{r}" / "documents.leann")
        os.makedirs(os.path.dirname(base))
        s.save(base)
        s.close()
        bases.append(base)
    sh = pkg.ShardedBackend.open(bases, pkg.BACKEND_HNSW, d, list(range(world)), exchange=mode)
    assert len(sh) == n
    rng = np.random.default_rng(2)
    m1 = pkg.pack_mask(rng.random(bounds[1][1] - bounds[1][0]) < 0.3)
    masks = [None, m1] + [None] * (world - 2)
    mk, md, mc = sh.search_batch(q, k, 64, shard_masks=masks)
    parts = [pkg.HnswSearcher.load(b, d, device=r) for r, b in enumerate(bases)]
    per = [p.search_batch(q, k, 64, mask=masks[r]) for r, p in enumerate(parts)]
    lk = np.stack([np.where(pk == NONE, pk, pk + np.uint64(lo)) for (pk, _, _), (lo, _) in zip(per, bounds)])
    ld = np.stack([pd for _, pd, _ in per])
    rk, rd = numpy_topk_merge(lk, ld, descending=False)
    assert np.array_equal(mk, rk) and np.array_equal(md.view(np.uint32), rd.view(np.uint32))
    assert sh.info()["exchanges"] == 1 and sh.info()["exchange_bytes"] > 0
    sh.close()
    for p in parts:
        p.close()


def test_process_per_gpu_join(pkg):
    g = _n_gpus(pkg)
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(g, 4)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29653", os.path.join(ROOT, "tests", "mp", "shards_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDS_WORKER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
