"""BASELINE config C5: hybrid vector + BM25 (alpha = 0.5) with a metadata-filter bitmask over N passages,
batched queries. Text: 64-256 tokens per passage drawn Zipf(1.07) from a 200k-word vocabulary (seed 777),
queries 2-6 tokens; metadata keys as the reference's chunkers write them (chunker/simple.rs:42-46).
  python benchmarks/bench_hybrid.py [--n 1000000] [--nq 10000] [--d 768]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import leann_rs_b200 as P

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000); ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--d", type=int, default=768); ap.add_argument("--k", type=int, default=10)
ap.add_argument("--vocab", type=int, default=200_000); ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(777)
t0 = time.time()
p = 1.0 / np.arange(1, a.vocab + 1) ** 1.07; cdf = np.cumsum(p / p.sum())
vocab = np.array([f"w{i}" for i in range(a.vocab)])
lens = rng.integers(64, 257, size=a.n)
ids = np.searchsorted(cdf, rng.random(int(lens.sum())))
offs = np.concatenate([[0], np.cumsum(lens)])
docs = [" ".join(vocab[ids[offs[i]:offs[i + 1]]].tolist()) for i in range(a.n)]
qlens = rng.integers(2, 7, size=a.nq)
qids = np.searchsorted(cdf, rng.random(int(qlens.sum())))
qoffs = np.concatenate([[0], np.cumsum(qlens)])
texts = [" ".join(vocab[qids[qoffs[i]:qoffs[i + 1]]].tolist()) for i in range(a.nq)]
exts = ["rs", "py", "md", "txt"]
metas = [json.dumps({"source": f"dir{i % 100}/f{i}.{exts[i % 4]}", "chunk_index": i % 7, "chunk_type": ["simple", "ast", "context"][i % 3],
                     "lines": int(l)}) for i, l in enumerate(rng.integers(1, 501, size=a.n))]
t_corpus = time.time() - t0
t0 = time.time(); bm = P.Bm25Scorer.build(docs); t_bm = time.time() - t0
st = bm.stats()
masks = {}
EXPRS = ("source:*.rs", "chunk_type=ast,lines>100", "lines>=490")
t0 = time.time()
rowwise = {expr: P.MetadataFilter.parse(expr).mask(metas) for expr in EXPRS}   # JSON parse + tree walk per passage and filter
t_mask = time.time() - t0
t0 = time.time(); cols = P.MetadataColumns(metas); t_cols = time.time() - t0       # side-car: parsed once
t_eval = {}
for expr in EXPRS:
    f = P.MetadataFilter.parse(expr)
    t0 = time.time(); masks[expr] = cols.mask(f); t_eval[expr] = round((time.time() - t0) * 1e3, 2)
    assert np.array_equal(masks[expr], rowwise[expr]), expr
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((32, a.d), generator=g, device=dev)
def gen(m, seed):
    gg = torch.Generator(device=dev); gg.manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn((m, 32), generator=gg, device=dev) @ W + 0.3 * torch.randn((m, a.d), generator=gg, device=dev), dim=1)
x = torch.cat([gen(min(1 << 18, a.n - s), 10 + s) for s in range(0, a.n, 1 << 18)])
t0 = time.time(); index = P.HnswSearcher.build(x, 32, 64); torch.cuda.synchronize(); t_idx = time.time() - t0
q = gen(a.nq, 4321).cpu().numpy()
# algorithmic bytes of the BM25 part: sum over query tokens of df * 8
tok_df = bm.search_batch(texts[:1], 1)  # warm
import ctypes as C
import gc
gc.collect(); gc.freeze(); gc.disable()   # millions of live corpus strings: a generational GC pass inside a timed call costs 100s of ms
rows = []
for name, mask in [("hybrid", None)] + [("hybrid+filter " + e, m) for e, m in masks.items()] + [("filter-only lines>=490", masks["lines>=490"])]:
    hybrid = not name.startswith("filter-only")
    P.text.hybrid_search(index, bm, q, texts if hybrid else None, a.k, 64, hybrid, 0.5, mask)   # warm
    ts = []
    for _ in range(a.steps):
        t0 = time.time()
        idx, sc, cnt = P.text.hybrid_search(index, bm, q, texts if hybrid else None, a.k, 64, hybrid, 0.5, mask)
        ts.append(time.time() - t0)
    dt = sorted(ts)[len(ts) // 2]
    rows.append({"mode": name, "pass_frac": None if mask is None else round(float(np.unpackbits(mask.view(np.uint8)).sum()) / a.n, 4),
                 "ms_per_batch_e2e": round(dt * 1e3, 2), "step_ms": [round(t * 1e3, 1) for t in ts], "qps_e2e": round(a.nq / dt), "mean_results": round(float(cnt.mean()), 2)})
# small batches: the reference API is one query per call (IndexSearcher::search_with_options)
small = []
for m in (1, 16, 128):
    ts = []
    for i in range(24):
        lo = (i * m) % (a.nq - m)
        t0 = time.time()
        P.text.hybrid_search(index, bm, q[lo:lo + m], texts[lo:lo + m], a.k, 64, True, 0.5, masks["source:*.rs"])
        ts.append(time.time() - t0)
    ts = sorted(ts[4:])
    small.append({"nq": m, "p50_ms": round(ts[len(ts) // 2] * 1e3, 3), "p90_ms": round(ts[int(len(ts) * 0.9)] * 1e3, 3)})
# BM25-only batched top-50 (Bm25Scorer::search) and consistency with the dense score_query path
k3 = []
for _ in range(4):
    t0 = time.time(); bi, bs, bc = bm.search_batch(texts, 50); t_bm25 = time.time() - t0
    npost, kms = bm.last_batch()
    k3.append({"host_call_ms": round(t_bm25 * 1e3, 1), "kernel_ms": round(kms, 2), "postings": npost,
               "algorithmic_GBps": round(npost * 8 / kms / 1e6, 1)})
ok = True
for i in range(16):
    dense = bm.score_query(texts[i])
    pos = np.nonzero(dense > 0)[0]
    order = pos[np.argsort(-dense[pos], kind="stable")][:50]
    ok &= bi[i, :bc[i]].tolist() == order.tolist() and np.array_equal(bs[i, :bc[i]], dense[order])
print(json.dumps({"bench": "hybrid", "n": a.n, "nq": a.nq, "k": a.k, "alpha": 0.5, "bm25_stats": st, "corpus_s": round(t_corpus, 1),
                  "bm25_build_s": round(t_bm, 1), "filter_masks_rowwise_s": round(t_mask, 2), "metadata_columns_build_s": round(t_cols, 2), "filter_mask_columnar_ms": t_eval, "hnsw_build_s": round(t_idx, 2),
                  "bm25_top50_batch_ms": round(t_bm25 * 1e3, 1), "bm25_top50_qps": round(a.nq / t_bm25), "k3_batches": k3, "hybrid_filter_small_batches": small,
                  "dense_vs_topk_consistent": bool(ok), "results": rows}))
