"""Exact scan latency for small query batches (RecomputeSearcher::search is one query per call, recompute.rs:52-123).
  python benchmarks/exact_small_batches.py [n] [d] [k]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, leann_rs_b200 as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 384
k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((32, d), generator=g, device=dev)
def gen(m, seed):
    gg = torch.Generator(device=dev); gg.manual_seed(seed)
    out = torch.empty((m, d), device=dev)
    for s0 in range(0, m, 1 << 20):
        mm = min(1 << 20, m - s0)
        out[s0:s0 + mm] = torch.nn.functional.normalize(torch.randn((mm, 32), generator=gg, device=dev) @ W + 0.3 * torch.randn((mm, d), generator=gg, device=dev), dim=1)
    return out
x = gen(n, 1); q = gen(4096, 2)
flat = P.FlatSearcher.from_vectors(x, metric=P.METRIC_DOT_DESC); del x
rows = []
for nq in (1, 8, 32, 63, 64, 256, 1024):
    qq = q[:nq].contiguous()
    for _ in range(2): flat.search_device(qq, k, 0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); flat.search_device(qq, k, 0); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    rows.append({"nq": nq, "ms": round(ms, 3), "qps": round(nq / ms * 1e3), "db_GBps_f32": round(n * d * 4 / ms / 1e6, 1)})
print(json.dumps({"bench": "exact_small", "n": n, "d": d, "k": k, "rows": rows}))
