#!/usr/bin/env python
"""bench.py — headline benchmark of the leann-rs search hot path on B200.

Metric (BASELINE.json): QPS at recall@10 >= 0.95 on 1M x 768 cosine HNSW (M=32), 10k-query batches,
plus the achieved HBM GB/s of the traversal kernel against the measured B200 peak.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--layout replica|shard]

A step = one batch of 10 000 queries per GPU through the K1 graph-search kernel.
`value`   : whole-job QPS with the queries already resident in HBM (CUDA events, max over ranks).
`e2e`     : the same through the host-buffer C-ABI call (`leann_cuda_search`), pinned host
            queries -> H2D -> kernel -> D2H of keys/distances inside the timed region.
`roofline`: algorithmic bytes the kernel itself counted (distance evaluations x row bytes +
            adjacency rows read) / average launch duration, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` / `--impl reference`: the CPU oracle (port of the usearch search loop the
            reference calls, hnsw.rs:85) on the box's host cores, same index file, same ef.
Data is synthetic (no embedding service offline): unit-normalised low-intrinsic-dimension
embeddings x = normalise(z W + 0.3 g), z in R^32 (seeded); queries are fresh draws.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DB, DIM, M_DEG, EF_ADD, TOP_K, NQ = 1_000_000, 768, 32, 64, 10, 10_000
RECALL_TARGET = 0.95
EF_SWEEP = (64, 96, 128, 160, 192, 256)
N_QUERY_BATCHES = 4
DB_SEED, Q_SEED = 1234, 4321


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layout", default="replica", choices=["replica", "shard"])
    ap.add_argument("--n", type=int, default=N_DB, help="database rows per index (default = the named config)")
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--cpu-sample", type=int, default=4000, help="queries in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v == "Active":
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def gen_lowrank(n, d, seed, device, W):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    step = 1 << 18
    for lo in range(0, n, step):
        m = min(step, n - lo)
        z = torch.randn((m, W.shape[0]), generator=g, device=device)
        x = z @ W + 0.3 * torch.randn((m, d), generator=g, device=device)
        out[lo:lo + m] = torch.nn.functional.normalize(x, dim=1)
    return out


def make_W(d, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(DB_SEED)
    return torch.randn((32, d), generator=g, device=device)


def recall_at_k(keys, gt):
    return (keys.unsqueeze(2) == gt.unsqueeze(1)).any(2).float().mean().item()


def load_measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def load_ncu_traffic(ef):
    p = os.path.join(ROOT, "profiles", "k1_ncu_summary.json")
    try:
        j = json.load(open(p))
        if j.get("n") == N_DB and j.get("dim") == DIM and j.get("ef") == ef:
            return j.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


def cpu_oracle_qps(index, queries_np, ef, steps, warmup, tmpdir):
    """Times the CPU oracle (port of the reference's usearch search) on the same index file."""
    import oracle
    base = os.path.join(tmpdir, "documents.leann")
    index.save(base)
    g = oracle.Hnsw.load(base.replace(".leann", ".index"), DIM)
    os.remove(base.replace(".leann", ".index"))
    cores = oracle.hardware_threads()
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        keys, dists, counts, stats = g.search(queries_np, TOP_K, ef, lanes=-1, next_cap=0, nthreads=cores)
        t1 = time.perf_counter()
        if s >= warmup:
            times.append(t1 - t0)
    return queries_np.shape[0] * len(times) / sum(times), cores, keys


def main():
    a = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference" and rank != 0:
        return 0
    import numpy as np
    import torch
    import torch.distributed as dist

    import leann_rs_b200 as P
    from leann_rs_b200 import shards as S  # noqa

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    use_dist = world > 1 and a.impl == "ours"
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n, nq = a.n, a.nq
    layout = a.layout if world > 1 else "single"

    # ---- data + index (setup, untimed) ------------------------------------------------------------
    W = make_W(DIM, dev)
    db_seed = DB_SEED + (rank if layout == "shard" else 0)   # shard layout: every rank owns different rows
    t0 = time.time()
    x = gen_lowrank(n, DIM, db_seed, dev, W)
    qb = [gen_lowrank(nq, DIM, Q_SEED + i, dev, W) for i in range(N_QUERY_BATCHES)]
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    index = P.HnswSearcher.build(x, graph_degree=M_DEG, complexity=EF_ADD, seed=DB_SEED)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    flat = P.FlatSearcher.from_vectors(x, metric=P.METRIC_IP)
    gts = [flat.search_device(q, TOP_K, 0)[0] for q in qb]
    torch.cuda.synchronize()
    flat.close()
    del flat
    torch.cuda.empty_cache()

    # ---- recall sweep: smallest ef with recall@10 >= 0.95 (untimed) ----------------------------------
    sweep, ef_star, rec_star = [], None, None

    def probe(ef):
        index.search_device(qb[0], TOP_K, ef)  # warm
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        keys, dists, counts = index.search_device(qb[0], TOP_K, ef)
        e1.record()
        torch.cuda.synchronize()
        r = recall_at_k(keys, gts[0])
        sweep.append({"ef": ef, "recall_at_10": round(r, 4), "qps_1batch": round(nq / e0.elapsed_time(e1) * 1e3)})
        return r

    prev = None
    for ef in EF_SWEEP:
        r = probe(ef)
        if r >= RECALL_TARGET:
            ef_star, rec_star = ef, r
            break
        prev = ef
    if ef_star is None:
        ef_star, rec_star = EF_SWEEP[-1], sweep[-1]["recall_at_10"]
    elif prev is not None:
        # refine between the last failing and the first passing grid point (multiples of 4)
        lo, hi = prev, ef_star
        while hi - lo > 4:
            mid = (lo + hi) // 8 * 4
            if mid <= lo:
                mid = lo + 4
            r = probe(mid)
            if r >= RECALL_TARGET:
                hi, ef_star, rec_star = mid, mid, r
            else:
                lo = mid
    for ef in EF_SWEEP:   # the rest of the published grid, for the recall/QPS curve
        if all(e["ef"] != ef for e in sweep):
            probe(ef)
    sweep.sort(key=lambda e: e["ef"])
    # the timed region must meet the target on every cycled batch, not only on batch 0
    while ef_star < EF_SWEEP[-1] and min(recall_at_k(index.search_device(q, TOP_K, ef_star)[0], g) for q, g in zip(qb, gts)) < RECALL_TARGET:
        ef_star += 4
        rec_star = probe(ef_star)

    tmpdir = tempfile.mkdtemp(prefix="leann_bench_")
    try:
        if a.impl == "reference":
            sample = qb[0][: a.cpu_sample].cpu().numpy()
            qps, cores, ckeys = cpu_oracle_qps(index, sample, ef_star, a.steps, a.warmup, tmpdir)
            crec = recall_at_k(torch.from_numpy(ckeys.astype(np.int64)).to(dev), gts[0][: a.cpu_sample])
            line = {
                "impl": "reference", "metric": "QPS at recall@10>=0.95, 1Mx768 cosine HNSW", "value": round(qps, 1),
                "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": round(a.cpu_sample / qps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: 1Mx768 cosine HNSW, M=32, ef_add=64, top-10", "n": n, "dim": DIM, "M": M_DEG,
                           "ef": ef_star, "recall_at_10": round(crec, 4), "queries_per_step": a.cpu_sample,
                           "note": "CPU oracle = port of usearch's search loop (reference: hnsw.rs:85); the index file "
                                   "was written by the GPU builder as untimed setup because a sequential CPU build of 1M x 768 "
                                   "takes hours; search itself runs only oracle code on host threads"},
                "cpu_baseline": {"value": round(qps, 1), "unit": "queries/s", "cores": cores, "kind": "port",
                                 "sample": f"{a.cpu_sample} queries of batch 0 per step, all host threads, SIMD-shaped f32 dot"},
                "e2e": {"value": round(qps, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
            }
            print(json.dumps(line))
            return 0

        # ---- algorithmic bytes per batch, counted by the kernel ------------------------------------------
        info = index.info()
        row_bytes = ((DIM + 3) // 4) * 16
        bytes_per_batch, ndist_mean = [], []
        for q in qb:
            st = torch.zeros((nq, 4), dtype=torch.int64, device=dev)
            index.search_device(q, TOP_K, ef_star, stats=st)
            torch.cuda.synchronize()
            tot = st.sum(0).tolist()
            bytes_per_batch.append(tot[0] * row_bytes + tot[1] * info["M0"] * 4 + tot[2] * info["M"] * 4)
            ndist_mean.append(tot[0] / nq)

        # ---- the search callable of this layout -------------------------------------------------------------
        def local_search(q, k, ef):
            kk, dd, _ = index.search_device(q, k, ef)
            return kk, dd

        if layout == "shard":
            eng = S.ShardedSearcher(local_search, rank * n, world, rank, False,
                                    lambda gk, gd, desc: P.topk_merge_device(gk, gd, desc)[:2], dist)
            step_fn = lambda i: eng.search(qb[i % N_QUERY_BATCHES], TOP_K, ef_star)
            queries_per_step_job = nq
        else:
            step_fn = lambda i: local_search(qb[i % N_QUERY_BATCHES], TOP_K, ef_star)
            queries_per_step_job = nq * world

        def barrier():
            if use_dist:
                dist.barrier()
            torch.cuda.synchronize()

        for i in range(a.warmup):
            step_fn(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
        barrier()
        if a.profile_range:
            torch.cuda.profiler.start()
        evs[0].record()
        for i in range(a.steps):
            step_fn(i)
            evs[i + 1].record()
        barrier()
        if a.profile_range:
            torch.cuda.profiler.stop()
        total_ms = evs[0].elapsed_time(evs[-1])
        step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps)]

        # ---- e2e: host buffers through the C ABI -------------------------------------------------------------
        hq = [q.cpu().pin_memory() for q in qb]
        hk = torch.empty((nq, TOP_K), dtype=torch.int64).pin_memory()
        hd = torch.empty((nq, TOP_K), dtype=torch.float32).pin_memory()
        hc = torch.empty((nq,), dtype=torch.int32).pin_memory()
        err = C.create_string_buffer(1024)
        L = P.lib()

        def e2e_step(i):
            q = hq[i % N_QUERY_BATCHES]
            rc = L.leann_cuda_search(index._h, C.c_void_p(q.data_ptr()), nq, TOP_K, ef_star, None, 0,
                                     C.c_void_p(hk.data_ptr()), C.c_void_p(hd.data_ptr()), C.c_void_p(hc.data_ptr()), err, 1024)
            if rc != 0:
                raise RuntimeError(err.value.decode())

        for i in range(max(1, a.warmup)):
            e2e_step(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            e2e_step(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop()
        e2e_rec = recall_at_k(hk.to(dev), gts[(a.steps - 1) % N_QUERY_BATCHES])

        if use_dist:
            t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms, e2e_s = t.tolist()
        value = queries_per_step_job * a.steps / (total_ms / 1e3)
        e2e_value = (nq * world if layout != "shard" else nq) * a.steps / e2e_s
        if layout == "shard":
            e2e_value = None  # the host-buffer call is per rank; the sharded merge runs on device tensors

        # ---- roofline of the dominant kernel (K1) --------------------------------------------------------------
        peak, peak_src = load_measured_peak()
        mean_bytes = sum(bytes_per_batch[i % N_QUERY_BATCHES] for i in range(a.steps)) / a.steps
        mean_ms = sum(step_ms) / len(step_ms)
        achieved = mean_bytes / (mean_ms / 1e3) / 1e9

        cpu = None
        if rank == 0 and world == 1 and not a.no_cpu_baseline:
            sample = qb[0][: a.cpu_sample].cpu().numpy()
            qps, cores, ckeys = cpu_oracle_qps(index, sample, ef_star, 1, 0, tmpdir)
            same = float((torch.from_numpy(ckeys.astype(np.int64)).to(dev) == index.search_device(qb[0][: a.cpu_sample].contiguous(), TOP_K, ef_star)[0]).float().mean().item())
            cpu = {"value": round(qps, 1), "unit": "queries/s", "cores": cores, "kind": "port",
                   "sample": f"{a.cpu_sample} queries of batch 0, one pass, all host threads, same index file and ef; "
                             f"id agreement with the GPU result {same:.4f} (different f32 summation order)"}

        if rank == 0:
            line = {
                "metric": "QPS at recall@10>=0.95, 1Mx768 cosine HNSW", "value": round(value, 1), "unit": "queries/s",
                "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(total_ms / a.steps, 4),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {
                    "workload": "configs[1]: 1Mx768 cosine HNSW (usearch format), M=32, ef_add=64, top-10, 10k-query batches",
                    "n": n, "dim": DIM, "M": M_DEG, "M0": info["M0"], "ef": ef_star, "recall_at_10": round(rec_star, 4),
                    "recall_target": RECALL_TARGET, "queries_per_step_per_gpu": nq, "query_batches_cycled": N_QUERY_BATCHES,
                    "parallelism": {"single": "1 GPU", "replica": f"replica x{world}: index replicated, queries split, no data-path collective",
                                    "shard": f"db-shard x{world}: {n} rows per GPU, all queries on every shard, NCCL all_gather + top-k merge kernel"}[layout],
                    "l2": "inputs larger than L2: 3.07 GB of vectors, ~%.0f GB gathered per step" % (mean_bytes / 1e9),
                    "generator": "x = normalise(z W + 0.3 g), z~N(0,I_32); DB seed 1234, query seeds 4321..",
                    "index_build_s": round(t_build, 2), "datagen_s": round(t_gen, 2),
                    "mean_distance_evals_per_query": round(sum(ndist_mean) / len(ndist_mean), 1),
                },
                "recall_sweep": sweep,
                "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": load_ncu_traffic(ef_star),
                             "peak_source": peak_src, "kernel": "graph_search_kernel<32,6,4>",
                             "algorithmic_bytes_per_launch": int(mean_bytes), "launch_ms": round(mean_ms, 4)},
                "cpu_baseline": cpu,
                "e2e": {"value": None if e2e_value is None else round(e2e_value, 1), "unit": "queries/s",
                        "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * TOP_K * 12 + nq * 4,
                        "recall_at_10": round(e2e_rec, 4)},
                "gpu_launches": a.steps * (2 if layout == "shard" else 1),
                "clocks": clocks,
            }
            print(json.dumps(line))
    finally:
        shutil.rmtree(tmpdir, ignore_errors=True)
        if use_dist:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
