"""Graph-search benchmark for the non-headline configs (BASELINE C1 / C4): builds the index on the GPU,
computes exact ground truth, reports recall@k, QPS and the HBM roofline fraction of K1.
  python benchmarks/bench_graph.py --backend vamana --n 12500000 --d 96 --metric l2 --deg 64 --L 100 --ef 100
  python benchmarks/bench_graph.py --backend hnsw --n 10000 --d 768 --deg 32 --L 64 --ef 64 --nq 1000     (C1)
  torchrun --nproc-per-node 8 benchmarks/bench_graph.py --rows 100000000 ...   (C4: --n rows split into one sub-index per rank,
      every rank searches all queries, NCCL all_gather + top-k merge kernel; timing = max over ranks)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import leann_rs_b200 as P
from leann_rs_b200 import shards as S

ap = argparse.ArgumentParser()
ap.add_argument("--backend", default="vamana"); ap.add_argument("--n", "--rows", dest="n", type=int, default=12_500_000)
ap.add_argument("--d", type=int, default=96); ap.add_argument("--metric", default="l2")
ap.add_argument("--deg", type=int, default=64); ap.add_argument("--L", type=int, default=100)
ap.add_argument("--ef", type=int, nargs="+", default=[100]); ap.add_argument("--k", type=int, default=10)
ap.add_argument("--nq", type=int, default=10_000); ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--rank", type=int, default=16, help="latent rank of the synthetic data")
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=dev)
lo, hi = S.shard_bounds(a.n, world, rank)
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((a.rank, a.d), generator=g, device=dev)
def gen(m, seed, normalize):
    gg = torch.Generator(device=dev); gg.manual_seed(seed)
    out = torch.empty((m, a.d), device=dev)
    for s0 in range(0, m, 1 << 20):
        mm = min(1 << 20, m - s0)
        v = torch.randn((mm, a.rank), generator=gg, device=dev) @ W + 0.3 * torch.randn((mm, a.d), generator=gg, device=dev)
        out[s0:s0 + mm] = torch.nn.functional.normalize(v, dim=1) if normalize else v
    return out
norm = a.metric != "l2"
metric = {"l2": P.METRIC_L2SQ, "ip": P.METRIC_IP, "dot": P.METRIC_IP_CLAMP}[a.metric]
x = gen(hi - lo, 1234 + rank, norm); q = gen(a.nq, 4321, norm)
torch.cuda.synchronize(); t0 = time.time()
if a.backend == "vamana":
    idx = P.DiskAnnSearcher.build(x, graph_degree=a.deg, complexity=a.L, alpha=1.2, metric=metric)
else:
    idx = P.HnswSearcher.build(x, graph_degree=a.deg, complexity=a.L, metric=metric)
torch.cuda.synchronize(); t_build = time.time() - t0
flat = P.FlatSearcher.from_vectors(x, metric=metric if a.metric == "l2" else P.METRIC_IP)
# sharded layout through the library's own handle (leann_cuda_shards_join: ncclCommInitRank inside, one all_gather + merge per batch)
def joined(local):
    uid = [P.ShardedBackend.unique_id() if rank == 0 else None]
    if world > 1: dist.broadcast_object_list(uid, src=0)
    return P.ShardedBackend.join(local, uid[0] if world > 1 else b"\0" * 128, rank, world, lo)
class Eng:
    def __init__(self, sh): self.sh = sh
    def search(self, qq, k, ef): return self.sh.search_device(qq, k, ef)[:2]
gsh = joined(flat)
gt = gsh.search_device(q, a.k, 0)[0]
torch.cuda.synchronize(); gsh.close(); flat.close(); del flat, x; torch.cuda.empty_cache()
eng = Eng(joined(idx))
info = idx.info()
peak = 6538.0
try: peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception: pass
rows = []
for ef in a.ef:
    st = torch.zeros((a.nq, 4), dtype=torch.int64, device=dev)
    idx.search_device(q, a.k, ef, stats=st); keys = eng.search(q, a.k, ef)[0]; torch.cuda.synchronize()
    rec = (keys.unsqueeze(2) == gt.unsqueeze(1)).any(2).float().mean().item()
    if world > 1: dist.all_reduce(st)
    tot = st.sum(0).tolist()
    byts = tot[0] * ((a.d + 3) // 4) * 16 + tot[1] * info["M0"] * 4 + tot[2] * info["M"] * 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): eng.search(q, a.k, ef)          # warm-up
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    evs[0].record()
    for i in range(a.steps): eng.search(q, a.k, ef); evs[i + 1].record()
    torch.cuda.synchronize()
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]
    ms = sorted(step_ms)[a.steps // 2]
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
    rows.append({"ef": ef, "recall": round(rec, 4), "qps": round(a.nq / ms * 1e3), "ms": round(ms, 3), "step_ms": [round(v, 2) for v in step_ms], "n_dist": round(tot[0] / a.nq, 1),
                 "hops": round(tot[1] / a.nq, 1), "algorithmic_GBps": round(byts / ms / 1e6, 1), "frac_of_hbm_peak": round(byts / ms / 1e6 / peak / world, 4)})
if world > 1: dist.destroy_process_group()
if rank == 0: print(json.dumps({"bench": "graph", "gpus": world, "exchange": eng.sh.info()["exchange"], "backend": a.backend, "n": a.n, "d": a.d, "metric": a.metric, "degree": a.deg, "L_build": a.L, "k": a.k,
                  "nq": a.nq, "build_s": round(t_build, 2), "info": info, "results": rows}))
