"""Short K1 run for ncu: Vamana n x 96 L2 (C4 shard shape), R = 64, beam L, a few 10k-query launches.
Variant via LEANN_K1_TUNE / LEANN_CUDA_DISABLE_REG_LISTS (see graph_search.cu)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leann_rs_b200 as P
from benchmarks import secondary as S2
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=12_500_000); ap.add_argument("--d", type=int, default=96)
ap.add_argument("--L", type=int, default=100); ap.add_argument("--launches", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda", 0)
W = S2._make_W(torch, dev, 16, a.d)
x = S2._gen(torch, dev, a.n, a.d, 1234, W, normalize=False)
q = S2._gen(torch, dev, 10_000, a.d, 4321, W, normalize=False)
idx = P.DiskAnnSearcher.build(x, graph_degree=64, complexity=100, alpha=1.2, metric=P.METRIC_L2SQ)
del x
for _ in range(a.launches):
    idx.search_device(q, 10, a.L)
torch.cuda.synchronize()
print("ok")
