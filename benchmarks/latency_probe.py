"""Latency of small host-buffer calls (the reference's API is one query per call, traits.rs:16-21):
leann_cuda_search with nq = 1, 4, 16, 64, 256 on the 1M x 768 HNSW index at ef = 64 (what HnswSearcher::search runs).
  python benchmarks/latency_probe.py [n] [ef]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, leann_rs_b200 as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 64
d, k = 768, 10
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((32, d), generator=g, device=dev)
gen = lambda m: torch.nn.functional.normalize(torch.randn((m, 32), generator=g, device=dev) @ W + 0.3 * torch.randn((m, d), generator=g, device=dev), dim=1)
x = torch.cat([gen(1 << 18) for _ in range((n + (1 << 18) - 1) >> 18)])[:n].contiguous()
idx = P.HnswSearcher.build(x, 32, 64)
q = gen(4096).cpu().numpy()
L = P.lib(); err = C.create_string_buffer(1024)
rows = []
for nq in (1, 4, 16, 64, 256, 1024):
    hk = np.empty((nq, k), dtype=np.uint64); hd = np.empty((nq, k), dtype=np.float32); hc = np.empty(nq, dtype=np.uint32)
    ts = []
    for i in range(60):
        qq = np.ascontiguousarray(q[(i * nq) % 2048:(i * nq) % 2048 + nq])
        t0 = time.perf_counter()
        rc = L.leann_cuda_search(idx._h, C.c_void_p(qq.ctypes.data), nq, k, ef, None, 0, C.c_void_p(hk.ctypes.data), C.c_void_p(hd.ctypes.data), C.c_void_p(hc.ctypes.data), err, 1024)
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0
    ts = sorted(ts[10:])
    rows.append({"nq": nq, "p50_ms": round(ts[len(ts) // 2], 3), "p90_ms": round(ts[int(len(ts) * 0.9)], 3), "qps_at_p50": round(nq / ts[len(ts) // 2] * 1e3)})
print(json.dumps({"bench": "latency", "n": n, "d": d, "ef": ef, "k": k, "rows": rows}))
