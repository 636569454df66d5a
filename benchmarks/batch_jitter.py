"""Per-batch device time of K1 over many consecutive batches (looks for periodic slow batches, e.g. the epoch wipe of
the visited byte map every 255 queries per warp).  python benchmarks/batch_jitter.py [n] [d] [ef] [batches]"""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import leann_rs_b200 as P
n, d, ef, nb = (int(v) for v in (sys.argv[1:5] + ["2000000", "96", "50", "200"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((16, d), generator=g, device=dev)
def gen(m, seed):
    gg = torch.Generator(device=dev); gg.manual_seed(seed)
    return torch.randn((m, 16), generator=gg, device=dev) @ W + 0.3 * torch.randn((m, d), generator=gg, device=dev)
x = gen(n, 1234); q = gen(10000, 4321)
idx = P.DiskAnnSearcher.build(x, graph_degree=64, complexity=100, alpha=1.2, metric=P.METRIC_L2SQ)
idx.search_device(q, 10, ef); torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(nb + 1)]
ev[0].record()
for i in range(nb):
    idx.search_device(q, 10, ef); ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(nb)]
srt = sorted(ms)
print(json.dumps({"n": n, "d": d, "ef": ef, "batches": nb, "median_ms": round(srt[nb // 2], 3), "mean_ms": round(sum(ms) / nb, 3),
                  "max_ms": round(srt[-1], 3), "slow_batches": [(i, round(v, 2)) for i, v in enumerate(ms) if v > 1.5 * srt[nb // 2]]}))
