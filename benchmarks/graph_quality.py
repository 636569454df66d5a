"""Graph quality A/B: sequential CPU build (oracle restatement of usearch add) vs the GPU batched builder on
the same data; recall@10 per ef, searched by the same GPU kernel."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import leann_rs_b200 as P, oracle
n, d = int(sys.argv[1]), 768
rng = np.random.default_rng(0)
W = rng.standard_normal((32, d), dtype=np.float32)
f = lambda m: rng.standard_normal((m, 32), dtype=np.float32) @ W + 0.3 * rng.standard_normal((m, d), dtype=np.float32)
x, q = f(n), f(2000)
x /= np.linalg.norm(x, axis=1, keepdims=True); q /= np.linalg.norm(q, axis=1, keepdims=True)
flat = P.FlatSearcher.from_vectors(x, metric=P.METRIC_IP); gt = flat.search_batch(q, 10, 0)[0]
def rec(keys): return float(np.mean([len(set(keys[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(q))]))
t = time.time(); g = oracle.Hnsw.build(x, M=32, ef_add=64, seed=1); t_cpu = time.time() - t
with tempfile.TemporaryDirectory() as td:
    g.save(os.path.join(td, "documents.index")); s_cpu = P.HnswSearcher.load(os.path.join(td, "documents.leann"), d)
t = time.time(); s_gpu = P.HnswSearcher.build(torch.from_numpy(x).cuda(), 32, 64); torch.cuda.synchronize(); t_gpu = time.time() - t
print("n", n, "cpu build %.1fs" % t_cpu, "gpu build %.2fs" % t_gpu)
for ef in (32, 64, 96, 128, 160, 256):
    print("ef", ef, "recall cpu-built %.4f" % rec(s_cpu.search_batch(q, 10, ef)[0]), "gpu-built %.4f" % rec(s_gpu.search_batch(q, 10, ef)[0]))
