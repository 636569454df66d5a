"""BASELINE config C5 over document-range shards (SURVEY §8e, BM25 / hybrid row): every rank indexes its slice of the
passages (HNSW sub-index + BM25 postings built with the corpus-wide statistics), answers all queries, and one exchange
step (all_gather of vector and BM25 top lists, all_reduce of the candidates' BM25 scores and of min/max) precedes the
fusion kernel. Same synthetic corpus as bench_hybrid.py.
  torchrun --nproc-per-node G benchmarks/bench_hybrid_sharded.py [--rows 1000000] [--nq 10000]"""
import argparse, gc, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
import leann_rs_b200 as P
from leann_rs_b200 import shards as S

ap = argparse.ArgumentParser()
ap.add_argument("--rows", dest="n", type=int, default=1_000_000); ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--d", type=int, default=768); ap.add_argument("--k", type=int, default=10)
ap.add_argument("--vocab", type=int, default=200_000); ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=dev)
lo, hi = S.shard_bounds(a.n, world, rank)
rng = np.random.default_rng(777)
p = 1.0 / np.arange(1, a.vocab + 1) ** 1.07; cdf = np.cumsum(p / p.sum())
vocab = np.array([f"w{i}" for i in range(a.vocab)])
lens = rng.integers(64, 257, size=a.n)
ids = np.searchsorted(cdf, rng.random(int(lens.sum())))
offs = np.concatenate([[0], np.cumsum(lens)])
docs = [" ".join(vocab[ids[offs[i]:offs[i + 1]]].tolist()) for i in range(lo, hi)]      # this rank's passages only
qlens = rng.integers(2, 7, size=a.nq)
qids = np.searchsorted(cdf, rng.random(int(qlens.sum())))
qoffs = np.concatenate([[0], np.cumsum(qlens)])
texts = [" ".join(vocab[qids[qoffs[i]:qoffs[i + 1]]].tolist()) for i in range(a.nq)]
exts = ["rs", "py", "md", "txt"]
metas = [json.dumps({"source": f"dir{i % 100}/f{i}.{exts[i % 4]}", "chunk_index": i % 7, "chunk_type": ["simple", "ast", "context"][i % 3],
                     "lines": int(l)}) for i, l in enumerate(rng.integers(1, 501, size=a.n))]
mask = P.MetadataColumns(metas).mask(P.MetadataFilter.parse("source:*.rs"))      # global bitmask, same on every rank
del metas, ids
t0 = time.time()
blob = P.Bm25Scorer.shard_stats(docs)
blobs = [None] * world
if world > 1: dist.all_gather_object(blobs, blob)
else: blobs = [blob]
bm = P.Bm25Scorer.build_sharded(docs, P.Bm25Scorer.merge_stats(blobs), device=lr)
t_bm = time.time() - t0
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((32, a.d), generator=g, device=dev)
def gen(m, seed):
    gg = torch.Generator(device=dev); gg.manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn((m, 32), generator=gg, device=dev) @ W + 0.3 * torch.randn((m, a.d), generator=gg, device=dev), dim=1)
parts = []
for s0 in range(0, a.n, 1 << 18):
    m = min(1 << 18, a.n - s0)
    if s0 + m <= lo or s0 >= hi: continue
    blk = gen(m, 10 + s0)
    parts.append(blk[max(lo - s0, 0): min(hi - s0, m)])
x = torch.cat(parts); del parts
t0 = time.time(); index = P.HnswSearcher.build(x, 32, 64); torch.cuda.synchronize(); t_idx = time.time() - t0
q = gen(a.nq, 4321)
merge = lambda gk, gd, desc: P.topk_merge_device(gk, gd, desc)[:2]
vec = S.ShardedSearcher(lambda qq, k, ef: index.search_device(qq, k, ef)[:2], lo, world, rank, False, merge, dist)
hy = S.ShardedHybridSearcher(vec, bm.search_shard, lambda *args: P.hybrid_fuse(*args, device=lr), lambda gk, gs: merge(gk, gs, True),
                             lo, world, rank, dist, exchange_device=dev)
# the same through the library's own sharded handle (leann_cuda_shards_join + leann_cuda_shards_hybrid_search): one all_gather, no host hop
uid = [P.ShardedBackend.unique_id() if rank == 0 else None]
if world > 1: dist.broadcast_object_list(uid, src=0)
sh = P.ShardedBackend.join(index, uid[0], rank, world, lo)
qh = q.cpu().numpy()
gc.collect(); gc.freeze(); gc.disable()
rows = []
for name, hybrid, m, abi in (("hybrid", True, None, False), ("hybrid+filter source:*.rs", True, mask, False), ("filter-only source:*.rs", False, mask, False),
                             ("C ABI: hybrid", True, None, True), ("C ABI: hybrid+filter source:*.rs", True, mask, True)):
    run = (lambda: sh.hybrid_search(bm, qh, texts, a.k, 64, hybrid, 0.5, m)) if abi else (lambda: hy.search(q, texts, a.k, 64, hybrid, 0.5, m, a.n))
    run()   # warm
    ts = []
    for _ in range(a.steps):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.time()
        idx, sc, cnt = run()
        dt = time.time() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = t.item()
        ts.append(dt)
    dt = sorted(ts)[len(ts) // 2]
    rows.append({"mode": name, "ms_per_batch_e2e": round(dt * 1e3, 2), "qps_e2e": round(a.nq / dt), "mean_results": round(float(cnt.mean()), 2),
                 "checksum": int(idx[cnt > 0, 0].astype(np.int64).sum() & 0xFFFFFFFF)})
if rank == 0:
    print(json.dumps({"bench": "hybrid_sharded", "gpus": world, "n": a.n, "rows_per_gpu": hi - lo, "nq": a.nq, "k": a.k, "alpha": 0.5,
                      "bm25_shard_stats": bm.stats(), "bm25_build_s": round(t_bm, 1), "hnsw_build_s": round(t_idx, 2), "results": rows}))
if world > 1: dist.destroy_process_group()
