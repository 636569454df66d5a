"""A/B harness for K1 on short rows (C4 shard: Vamana n x 96 L2, R = 64): one index, the round-2 changes switched off one by
one through the library's A/B environment switches, results compared bit for bit with the round-1 kernel.
  base     LEANN_CUDA_DISABLE_REG_LISTS      shared-memory top / next lists (round 1)
  single   LEANN_CUDA_DISABLE_Q16            one list with an expanded bit, u32 hash / byte maps
  default                                    + bucketed 16-bit quotiented visited tables
(The (unroll, CTAs per SM) sweeps of profiles/r2_k1_tune*.log were made with a temporary LEANN_K1_TUNE switch.)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leann_rs_b200 as P
from benchmarks import secondary as S2

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=12_500_000); ap.add_argument("--d", type=int, default=96)
ap.add_argument("--nq", type=int, default=10_000); ap.add_argument("--variants", default="base;single;default")
ap.add_argument("--efs", default="50,100"); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
W = S2._make_W(torch, dev, 16, a.d)
x = S2._gen(torch, dev, a.n, a.d, 1234, W, normalize=False)
q = S2._gen(torch, dev, a.nq, a.d, 4321, W, normalize=False)
t0 = time.time()
cache = os.environ.get("K1_INDEX_CACHE")   # e.g. /dev/shm/k1c4: several A/B processes (LEANN_CUDA_LIB=...) of one session share the index
if cache and os.path.exists(cache + ".diskann"):
    idx = P.DiskAnnSearcher.load(cache + ".leann", a.d, metric=P.METRIC_L2SQ)
    print("load_s", round(time.time() - t0, 1), os.path.basename(P.LIB_PATH), flush=True)
else:
    idx = P.DiskAnnSearcher.build(x, graph_degree=64, complexity=100, alpha=1.2, metric=P.METRIC_L2SQ)
    torch.cuda.synchronize()
    print("build_s", round(time.time() - t0, 1), os.path.basename(P.LIB_PATH), flush=True)
    if cache:
        idx.save(cache + ".leann")
del x
info = idx.info()
rows = []
ref = {}
variants = [v for v in a.variants.replace("base,", "base;").split(";") if v]
SW = {"base": ["LEANN_CUDA_DISABLE_REG_LISTS"], "single": ["LEANN_CUDA_DISABLE_Q16"], "default": []}
for v in variants:
    for k in ("LEANN_CUDA_DISABLE_REG_LISTS", "LEANN_CUDA_DISABLE_SINGLE_LIST", "LEANN_CUDA_DISABLE_Q16"):
        os.environ.pop(k, None)
    for k in SW[v]:
        os.environ[k] = "1"
    for ef in [int(e) for e in a.efs.split(",")]:
        st = torch.zeros((a.nq, 4), dtype=torch.int64, device=dev)
        keys, dists, _ = idx.search_device(q, 10, ef, stats=st)
        torch.cuda.synchronize()
        if v == "base":
            ref[ef] = (keys.clone(), dists.clone(), st.clone())
        same = None
        if ef in ref:   # ids, distance bits and the work counters (evaluations, hops); the queue-drop flag is 0 by construction in single-list mode
            same = bool(torch.equal(keys, ref[ef][0]) and torch.equal(dists.view(torch.int32), ref[ef][1].view(torch.int32)) and torch.equal(st[:, :3], ref[ef][2][:, :3]))
        tot = st.sum(0).tolist()
        byts = tot[0] * ((a.d + 3) // 4) * 16 + tot[1] * info["M0"] * 4
        ms, step_ms = S2._timed(torch, lambda: idx.search_device(q, 10, ef), a.steps, 3)
        import zlib
        crc = zlib.crc32(keys.cpu().numpy().tobytes()) ^ zlib.crc32(dists.cpu().numpy().tobytes()) ^ zlib.crc32(st[:, :3].cpu().numpy().tobytes())
        rows.append({"variant": v, "lib": os.path.basename(P.LIB_PATH), "crc": crc, "ef": ef, "ms": round(ms, 3), "qps": round(a.nq / ms * 1e3), "frac": round(byts / ms / 1e6 / 6538.0, 4), "identical_to_base": same})
        print(json.dumps(rows[-1]), flush=True)
