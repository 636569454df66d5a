"""K1 short rows: cost of the visited set. Same index (C4 shard shape), visited set forced to byte maps (mode 1), automatic,
or per-warp hash tables of a given capacity (L2-resident when small enough); results must be identical."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leann_rs_b200 as P
from benchmarks import secondary as S2
n, d, nq = int(os.environ.get("N", 12_500_000)), 96, 10_000
dev = torch.device("cuda", 0)
W = S2._make_W(torch, dev, 16, d)
x = S2._gen(torch, dev, n, d, 1234, W, normalize=False)
q = S2._gen(torch, dev, nq, d, 4321, W, normalize=False)
idx = P.DiskAnnSearcher.build(x, graph_degree=64, complexity=100, alpha=1.2, metric=P.METRIC_L2SQ)
del x
ref = {}
for mode in (1, 0, 8192, 16384, 32768):
    idx.set_visited_hash(mode)
    for ef in (50, 100):
        keys, dists, _ = idx.search_device(q, 10, ef)
        torch.cuda.synchronize()
        if mode == 1: ref[ef] = keys.clone()
        ms, _ = S2._timed(torch, lambda: idx.search_device(q, 10, ef), 5, 3)
        print(json.dumps({"visited": {1: "byte maps", 0: "auto"}.get(mode, f"hash {mode}"), "ef": ef, "ms": round(ms, 3), "qps": round(nq / ms * 1e3),
                          "identical": bool(torch.equal(keys, ref[ef])), "ws": idx.workspace_stats()}), flush=True)
