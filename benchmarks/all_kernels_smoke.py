"""Small end-to-end pass over every kernel family (graph pool / CTA-per-query / append / masks, exact scan resident / streamed / f32,
BM25 build / search / shards / fusion). compute-sanitizer is closed on this GPU pool, so this is only a quick functional sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import leann_rs_b200 as P
rng = np.random.default_rng(0)
def data(n, d, nq):
    W = rng.standard_normal((16, d), dtype=np.float32)
    f = lambda m: (rng.standard_normal((m, 16), dtype=np.float32) @ W + 0.3 * rng.standard_normal((m, d), dtype=np.float32))
    x, q = f(n), f(nq)
    return x / np.linalg.norm(x, axis=1, keepdims=True), q / np.linalg.norm(q, axis=1, keepdims=True)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "graph"):
    for d, M in ((96, 8), (768, 8)):
        x, q = data(3000, d, 700)
        s = P.HnswSearcher.build(x[:2000], graph_degree=M, complexity=32)
        s.add(x[2000:], start_id=2000)
        for nq in (1, 200, 700):                                  # 8-warp CTA, 4-warp CTA, warp pool
            k, dd, c = s.search_batch(q[:nq], 10, 48)
            assert (c == 10).all()
        mask = P.pack_mask(rng.random(3000) < 0.3)
        s.search_batch(q[:64], 5, 32, mask=mask)
        v = P.DiskAnnSearcher.build(x, graph_degree=16, complexity=32, metric=P.METRIC_L2SQ)
        v.search_batch(q[:300], 10, 40)
    print("graph ok")
if which in ("all", "exact"):
    x, q = data(20000, 384, 130)
    f = P.FlatSearcher.from_vectors(x, metric=P.METRIC_DOT_DESC)
    for nq in (1, 130):
        k, sc, c = f.search_batch(q[:nq], 20, 0)
        ref = np.argsort(-(q[:nq] @ x.T), axis=1, kind="stable")[:, :20]
        assert (k.astype(np.int64) == ref).mean() > 0.99
    x2, q2 = data(18000, 520, 70)                                  # streamed query tile
    P.FlatSearcher.from_vectors(x2, metric=P.METRIC_IP).search_batch(q2, 5, 0)
    P.FlatSearcher.from_vectors(x[:3000], metric=P.METRIC_L2SQ).search_batch(q, 7, 0)
    print("exact ok")
if which in ("all", "text"):
    words = [f"w{i}" for i in range(500)]
    p = 1.0 / np.arange(1, 501) ** 1.07; p /= p.sum()
    docs = [" ".join(rng.choice(words, size=int(rng.integers(1, 50)), p=p)) for _ in range(20000)]
    qs = [" ".join(rng.choice(words, size=int(rng.integers(1, 5)), p=p)) for _ in range(64)] + ["", "zzz"]
    bm = P.Bm25Scorer.build(docs)
    bm.search_batch(qs, 50); bm.score_query(qs[0])
    st = P.Bm25Scorer.merge_stats([P.Bm25Scorer.shard_stats(docs[:9000]), P.Bm25Scorer.shard_stats(docs[9000:])])
    sh = P.Bm25Scorer.build_sharded(docs[9000:], st)
    x, q = data(20000, 128, len(qs))
    idx = P.HnswSearcher.build(x, graph_degree=8, complexity=32)
    mask = P.pack_mask(rng.random(20000) < 0.4)
    P.text.hybrid_search(idx, bm, q, qs, 10, 64, True, 0.5, mask)
    vk, vd, vc = idx.search_batch(q, 50, 64)
    ti, ts, tc, cb, bx, bn = sh.search_shard(qs, 50, 9000, vk, vc)
    P.hybrid_fuse(vk, vd, vc, 10, True, 0.5, cb, ti, ts, tc, bx, bn, mask, 20000)
    print("text ok")
