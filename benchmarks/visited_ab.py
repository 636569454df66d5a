"""A/B of the traversal's visited-set representations on one index: per-warp byte maps (auto on small indexes) against forced
per-warp hash tables of several capacities.  python benchmarks/visited_ab.py [n] [d]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import leann_rs_b200 as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 96
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((16, d), generator=g, device=dev)
def gen(m, seed):
    gg = torch.Generator(device=dev); gg.manual_seed(seed)
    return torch.randn((m, 16), generator=gg, device=dev) @ W + 0.3 * torch.randn((m, d), generator=gg, device=dev)
x = gen(n, 1234); q = gen(10000, 4321)
idx = P.DiskAnnSearcher.build(x, graph_degree=64, complexity=100, alpha=1.2, metric=P.METRIC_L2SQ)
rows = []
for mode in (0, 8192, 16384, 32768, 0):
    idx.set_visited_hash(mode)
    for ef in (50, 100):
        for _ in range(3): idx.search_device(q, 10, ef)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): idx.search_device(q, 10, ef)
        e1.record(); torch.cuda.synchronize()
        rows.append({"mode": mode, "ef": ef, "ms": round(e0.elapsed_time(e1) / 8, 3)})
print(json.dumps({"bench": "visited_ab", "n": n, "d": d, "rows": rows}))
