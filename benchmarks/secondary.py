"""Secondary records of bench.py's N = 1 line: BASELINE.json configs[2], [3] (one shard) and [4], each measured on the
GPU with its roofline and checked against the CPU oracle on a sample of the same inputs.

  c3  exact brute-force kNN 10M x 384 dot-product, top-100, 10k queries            -> K2 (tcgen05), tensor roofline
  c4  one shard of the DiskANN/Vamana config: 12.5M x 96 L2, R = 64, L = 100       -> K1 short rows, HBM roofline
  c5  hybrid vector + BM25 (alpha = 0.5) + metadata-filter bitmask, 1M passages     -> K1 + K3 + K3f, e2e through host buffers

Every function returns a plain dict; bench.py puts them under "secondary". Nothing here is a product path: the oracle
(`oracle/`) is imported as the checker and, for the cpu_* figures, as the timed CPU baseline on the sample.
"""
from __future__ import annotations

import json
import os
import shutil
import tempfile
import time

import numpy as np

NONE = np.uint64(0xFFFFFFFFFFFFFFFF)


def _gen(torch, dev, m, d, seed, W, normalize=True, noise=0.3):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((m, d), dtype=torch.float32, device=dev)
    step = 1 << 20
    for lo in range(0, m, step):
        mm = min(step, m - lo)
        v = torch.randn((mm, W.shape[0]), generator=g, device=dev) @ W + noise * torch.randn((mm, d), generator=g, device=dev)
        out[lo:lo + mm] = torch.nn.functional.normalize(v, dim=1) if normalize else v
    return out


def _make_W(torch, dev, rank, d, seed=1234):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    return torch.randn((rank, d), generator=g, device=dev)


def _timed(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return sum(ms) / len(ms), ms


def _scratch_dir():
    for base in ("/dev/shm", None):
        try:
            if base is None or (os.path.isdir(base) and shutil.disk_usage(base).free > (24 << 30)):
                return tempfile.mkdtemp(prefix="leann_bench_", dir=base)
        except Exception:
            pass
    return tempfile.mkdtemp(prefix="leann_bench_")


# -----------------------------------------------------------------------------------------------------------------
def run_c3(P, torch, dev, orc, peaks, steps=5, warmup=3, n=10_000_000, d=384, k=100, nq=10_000, sample=8):
    """configs[2] on one GPU. flops = 2 N d per query (counted once); roofline = sustained bf16 tensor peak."""
    t0 = time.time()
    W = _make_W(torch, dev, 32, d)
    x = _gen(torch, dev, n, d, 1000, W)
    q = _gen(torch, dev, nq, d, 4321, W)
    flat = P.FlatSearcher.from_vectors(x, metric=P.METRIC_DOT_DESC)
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    out = None

    def step():
        nonlocal out
        out = flat.search_device(q, k, 0)

    ms, step_ms = _timed(torch, step, steps, warmup)
    keys, scores, _ = out
    tflops = 2.0 * n * d * nq / (ms / 1e3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", 1395.4))
    # e2e through the host-buffer ABI call (one batch)
    hq = q.cpu().numpy()
    flat.search_batch(hq, k, 0)   # warm-up at the timed size (workspace, pinned staging)
    e2e_steps = []
    for _ in range(3):
        t0 = time.perf_counter()
        hk, hs, hc = flat.search_batch(hq, k, 0)
        e2e_steps.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = sum(e2e_steps) / len(e2e_steps)
    # oracle on a sample of the SAME database (sequential f32 fold of recompute.rs:137-139, stable sort :106)
    xs = x.cpu().numpy()
    cores = orc.hardware_threads()
    t0 = time.perf_counter()
    oi, osc, oc = orc.exact_scan(hq[:sample], xs, k, metric=0, nthreads=cores)
    cpu_s = time.perf_counter() - t0
    gk = keys[:sample].cpu().numpy().astype(np.uint64)
    gs = scores[:sample].cpu().numpy()
    neq = gk != oi
    tol = 1e-5 * np.abs(osc) + np.spacing(np.abs(osc))
    excused = int(np.sum(neq & (np.abs(gs - osc) <= tol)))
    unexcused = int(neq.sum()) - excused
    res = {
        "workload": f"configs[2]: exact kNN {n}x{d} dot-product, top-{k}, {nq}-query batch, 1 GPU",
        "ms": round(ms, 3), "step_ms": [round(v, 2) for v in step_ms], "qps": round(nq / ms * 1e3, 1),
        "tflops": round(tflops, 1), "peak_tflops": peak, "frac": round(tflops / peak, 4),
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained", "bound": "tensor",
        "kernel": "scan_tc_kernel<true> (tcgen05 cta_group::2) + rerank_kernel + select_kernel",
        "flops_per_launch": 2 * n * d * nq, "e2e_ms_host_buffers": round(e2e_ms, 2), "e2e_step_ms": [round(v, 1) for v in e2e_steps],
        "h2d_bytes": nq * d * 4, "d2h_bytes": nq * k * 12 + nq * 4,
        "oracle": {"sample_queries": sample, "id_agreement": round(float(1.0 - neq.mean()), 6),
                   "tie_excused_positions": excused, "outside_tie_rule": unexcused, "positions": int(neq.size),
                   "rule": "ids equal the oracle's f32 result except ties within 1e-5 relative score"},
        "cpu_baseline": {"value": round(sample / cpu_s, 3), "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} queries over the full {n}x{d} database, all host threads"},
        "setup_s": round(t_setup, 1), "gpu_launches_per_step": "3 per round (scan, rerank, select) x rounds + 4",
    }
    flat.close()
    del x, flat
    torch.cuda.empty_cache()
    return res


# -----------------------------------------------------------------------------------------------------------------
def run_c4(P, torch, dev, orc, peaks, steps=5, warmup=3, n=12_500_000, d=96, R=64, L=100, k=10, nq=10_000, sample=200,
           efs=(50, 100)):
    """One shard of configs[3] (100M x 96 L2 over 8 GPUs = 12.5M rows per GPU): Vamana R = 64 built with L = 100 on the
    GPU, searched with beam L. HBM roofline from the kernel's own counters."""
    t0 = time.time()
    W = _make_W(torch, dev, 16, d)
    x = _gen(torch, dev, n, d, 1234, W, normalize=False)
    q = _gen(torch, dev, nq, d, 4321, W, normalize=False)
    torch.cuda.synchronize()
    tb = time.time()
    idx = P.DiskAnnSearcher.build(x, graph_degree=R, complexity=L, alpha=1.2, metric=P.METRIC_L2SQ)
    torch.cuda.synchronize()
    t_build = time.time() - tb
    flat = P.FlatSearcher.from_vectors(x, metric=P.METRIC_L2SQ)
    gt = flat.search_device(q, k, 0)[0]
    torch.cuda.synchronize()
    flat.close()
    del flat, x
    torch.cuda.empty_cache()
    t_setup = time.time() - t0
    info = idx.info()
    peak = float(peaks.get("hbm_gbs", 6538.0))
    row_bytes = ((d + 3) // 4) * 16
    rows = []
    for ef in efs:
        st = torch.zeros((nq, 4), dtype=torch.int64, device=dev)
        keys = idx.search_device(q, k, ef, stats=st)[0]
        torch.cuda.synchronize()
        rec = (keys.unsqueeze(2) == gt.unsqueeze(1)).any(2).float().mean().item()
        tot = st.sum(0).tolist()
        byts = tot[0] * row_bytes + tot[1] * info["M0"] * 4 + tot[2] * info["M"] * 4
        ms, step_ms = _timed(torch, lambda: idx.search_device(q, k, ef), steps, warmup)
        rows.append({"L": ef, "recall_at_10": round(rec, 4), "qps": round(nq / ms * 1e3), "ms": round(ms, 3),
                     "step_ms": [round(v, 2) for v in step_ms], "distance_evals_per_query": round(tot[0] / nq, 1),
                     "hops_per_query": round(tot[1] / nq, 1), "algorithmic_bytes_per_launch": int(byts),
                     "achieved_gbs": round(byts / ms / 1e6, 1), "frac": round(byts / ms / 1e6 / peak, 4)})
    main = rows[-1]
    # oracle (diskann-rs search_with_dists restated) on the file the product wrote, sample of the same queries
    tmp = _scratch_dir()
    orec = {}
    try:
        base = os.path.join(tmp, "documents.leann")
        t0 = time.time()
        idx.save(base)
        g = orc.Vamana.load(base.replace(".leann", ".diskann"))
        g.set_metric(1)   # L2 squared (BASELINE C4; the reference itself hard-wires DistDot, diskann.rs:16)
        t_io = time.time() - t0
        hq = q[:sample].cpu().numpy()
        cap = P.queue_capacity(max(L, k), False)
        cores = orc.hardware_threads()
        t0 = time.perf_counter()
        ok, od, oc, ost = g.search(hq, k, L, lanes=P.reduction_lanes(d), next_cap=cap, nthreads=cores)
        cpu_s = time.perf_counter() - t0
        gk, gd, gc = idx.search_batch(hq, k, L)
        orec = {"sample_queries": sample, "ids_identical": bool(np.array_equal(gk, ok)),
                "id_agreement": round(float(np.mean(gk == ok)), 6),
                "distance_bits_identical": bool(np.array_equal(gd.view(np.uint32), od.view(np.uint32))),
                "save_load_s": round(t_io, 1)}
        cpu = {"value": round(sample / cpu_s, 1), "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"{sample} queries, beam {L}, same .diskann file, all host threads"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    res = {
        "workload": f"configs[3], one of 8 shards: Vamana {n}x{d} L2, R={R}, L={L}, top-{k}, {nq}-query batch, 1 GPU",
        "qps": main["qps"], "ms": main["ms"], "recall_at_10": main["recall_at_10"], "bound": "hbm",
        "achieved_gbs": main["achieved_gbs"], "peak_gbs": peak, "frac": main["frac"],
        "peak_source": "MEASURED_PEAKS.json hbm_gbs", "kernel": "graph_search_kernel (8 lanes per vector, short rows)",
        "beams": rows, "build_s": round(t_build, 1), "setup_s": round(t_setup, 1), "oracle": orec, "cpu_baseline": cpu,
        "gpu_launches_per_step": 1,
    }
    idx.close()
    torch.cuda.empty_cache()
    return res


# -----------------------------------------------------------------------------------------------------------------
def _corpus(n, nq, vocab_n, rng):
    """Passages of 64-256 tokens drawn Zipf(1.07) from a `vocab_n`-word vocabulary ("w000123"), queries of 2-6 tokens.
    Returns (docs: list[bytes], texts: list[bytes], token ids + offsets for the oracle)."""
    p = 1.0 / np.arange(1, vocab_n + 1) ** 1.07
    cdf = np.cumsum(p / p.sum())
    vb = np.frombuffer("".join(f"w{i:06d} " for i in range(vocab_n)).encode(), dtype=np.uint8).reshape(vocab_n, 8)
    lens = rng.integers(64, 257, size=n)
    ids = np.searchsorted(cdf, rng.random(int(lens.sum()))).astype(np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = vb[ids].reshape(-1).tobytes()
    docs = [flat[offs[i] * 8: offs[i + 1] * 8 - 1] for i in range(n)]
    qlens = rng.integers(2, 7, size=nq)
    qids = np.searchsorted(cdf, rng.random(int(qlens.sum()))).astype(np.int32)
    qoffs = np.concatenate([[0], np.cumsum(qlens)]).astype(np.int64)
    qflat = vb[qids].reshape(-1).tobytes()
    texts = [qflat[qoffs[i] * 8: qoffs[i + 1] * 8 - 1] for i in range(nq)]
    return docs, texts, ids, offs, lens, qids, qoffs


def _oracle_scorer_for(T, tokens, ids, offs, lens, n):
    """An oracle Bm25Scorer (bm25.rs:33-122 restated in oracle/text_oracle.py) holding the corpus statistics and the
    postings of `tokens` only — computed with numpy from the token-id corpus instead of re-tokenising 1M strings."""
    s = object.__new__(T.Bm25Scorer)
    s.num_docs = n
    s.doc_lengths = lens.astype(np.int64).tolist()
    s.avg_doc_len = np.float32(int(lens.sum())) / np.float32(n)
    s.doc_freq, s._post, s.term_freqs = {}, {}, None
    for t in sorted(set(int(v) for v in tokens)):
        pos = np.nonzero(ids == t)[0]
        docs = np.searchsorted(offs, pos, side="right") - 1
        ud, cnt = np.unique(docs, return_counts=True)
        name = f"w{t:06d}"
        s.doc_freq[name] = int(ud.size)
        s._post[name] = (ud.tolist(), cnt.tolist())
    return s


def run_c5(P, torch, dev, orc, peaks, index, q_dev, steps=3, warmup=1, k=10, ef=64, alpha=0.5, vocab_n=200_000, sample=8):
    """configs[4]: hybrid vector + BM25 (alpha = 0.5) with a metadata-filter bitmask over the passages of `index`
    (the 1M x 768 HNSW index of the headline config), batched queries, host buffers end to end."""
    from oracle import text_oracle as T
    n, nq = len(index), q_dev.shape[0]
    rng = np.random.default_rng(777)
    t0 = time.time()
    docs, texts, ids, offs, lens, qids, qoffs = _corpus(n, nq, vocab_n, rng)
    exts = ["rs", "py", "md", "txt"]
    metas = [json.dumps({"source": f"dir{i % 100}/f{i}.{exts[i % 4]}", "chunk_index": i % 7,
                         "chunk_type": ["simple", "ast", "context"][i % 3], "lines": int(l)})
             for i, l in enumerate(rng.integers(1, 501, size=n))]
    t_corpus = time.time() - t0
    t0 = time.time()
    bm = P.Bm25Scorer.build(docs)
    t_bm = time.time() - t0
    t0 = time.time()
    cols = P.MetadataColumns(metas)
    t_cols = time.time() - t0
    exprs = ("source:*.rs", "chunk_type=ast,lines>100", "lines>=490")
    masks, t_mask = {}, {}
    for e in exprs:
        f = P.MetadataFilter.parse(e)
        t0 = time.time()
        masks[e] = cols.mask(f)
        t_mask[e] = round((time.time() - t0) * 1e3, 2)
    q = q_dev.cpu().pin_memory().numpy()   # pinned host queries, as in the headline's e2e leg
    import gc
    gc.collect(); gc.freeze(); gc.disable()   # millions of live corpus objects: a GC pass inside a timed call costs 100s of ms
    rows = []
    try:
        for name, mask in [("hybrid", None)] + [("hybrid+filter " + e, masks[e]) for e in exprs]:
            for _ in range(warmup):
                P.text.hybrid_search(index, bm, q, texts, k, ef, True, alpha, mask)
            ts = []
            for _ in range(steps):
                t0 = time.perf_counter()
                idx_, sc_, cnt_ = P.text.hybrid_search(index, bm, q, texts, k, ef, True, alpha, mask)
                ts.append((time.perf_counter() - t0) * 1e3)
            rows.append({"mode": name, "pass_frac": None if mask is None else round(float(np.unpackbits(mask.view(np.uint8)).sum()) / n, 4),
                         "e2e_ms": round(sum(ts) / len(ts), 2), "step_ms": [round(t, 1) for t in ts],
                         "qps_e2e": round(nq / (sum(ts) / len(ts)) * 1e3), "mean_results": round(float(cnt_.mean()), 2)})
        # K3 alone (Bm25Scorer::search batched, top fetch_k = 5k): algorithmic bytes = postings covered x 8
        k3 = []
        for _ in range(3):
            bm.search_batch(texts, 5 * k)
            npost, kms = bm.last_batch()
            k3.append((npost, kms))
        npost, kms = k3[-1]
        k3_bytes, k3_rows = bm.last_batch_bytes(), bm.dense_rows()
    finally:
        gc.enable()
    peak = float(peaks.get("hbm_gbs", 6538.0))
    # oracle on a sample: BM25 (bm25.rs) + hybrid_rerank + the post-filter walk of searcher.rs:123-210, vector
    # candidates as the backend returned them (their parity is the headline's own check)
    main_expr = exprs[0]
    mbits = np.unpackbits(masks[main_expr].view(np.uint8), bitorder="little")[:n].astype(bool)
    toks = qids[qoffs[0]:qoffs[sample]]
    t0 = time.perf_counter()
    scorer = _oracle_scorer_for(T, toks, ids, offs, lens, n)
    fk = T.fetch_k(k, True, True)
    vk, vd, vc = index.search_batch(q[:sample], fk, ef)
    gi, gs, gc_ = P.text.hybrid_search(index, bm, q[:sample], texts[:sample], k, ef, True, alpha, masks[main_expr])
    ids_ok = bits_ok = 0
    for i in range(sample):
        c = int(vc[i])
        ref = T.search_with_options(vk[i, :c].astype(np.int64).tolist(), vd[i, :c].tolist(), k, scorer, texts[i].decode(), True, alpha,
                                    lambda j: bool(mbits[j]), fk, fast=True)
        ri = np.array([r[0] for r in ref], dtype=np.uint64)
        rs = np.array([r[1] for r in ref], dtype=np.float32)
        m = int(gc_[i])
        ids_ok += int(m == len(ref) and np.array_equal(gi[i, :m], ri))
        bits_ok += int(m == len(ref) and np.array_equal(gs[i, :m].view(np.uint32), rs.view(np.uint32)))
    cpu_s = time.perf_counter() - t0
    main = rows[1]
    res = {
        "workload": f"configs[4]: hybrid vector+BM25 (alpha={alpha}) + metadata-filter bitmask '{main_expr}', {n} passages, "
                    f"{nq}-query batch, top-{k} (fetch_k {5 * k}), host buffers end to end",
        "e2e_ms": main["e2e_ms"], "qps_e2e": main["qps_e2e"], "modes": rows,
        "h2d_bytes": nq * q.shape[1] * 4 + sum(len(t) for t in texts) + len(masks[main_expr]) * 8, "d2h_bytes": nq * k * 12 + nq * 4,
        "k3": {"kernel": "bm25_query_kernel", "kernel_ms": round(kms, 3), "postings": int(npost), "dense_rows": int(k3_rows),
               "algorithmic_bytes": int(k3_bytes), "algorithmic_bytes_note": "8 B per posting of a sparse token, 4 B per document of a token "
               "whose term has a dense row (K3d); the all-postings figure is postings x 8",
               "postings_x8_gbs": round(npost * 8 / kms / 1e6, 1),
               "achieved_gbs": round(k3_bytes / kms / 1e6, 1), "peak_gbs": peak, "frac": round(k3_bytes / kms / 1e6 / peak, 4),
               "bound": "hbm (postings and rows cross L2->SM once; see profiles/)"},
        "bm25_stats": bm.stats(), "corpus_s": round(t_corpus, 1), "bm25_build_s": round(t_bm, 1),
        "metadata_columns_build_s": round(t_cols, 2), "filter_mask_ms": t_mask,
        "oracle": {"sample_queries": sample, "ids_identical": ids_ok, "score_bits_identical": bits_ok,
                   "what": "oracle BM25 + hybrid_rerank + post-filter walk (text_oracle.py) on the product's vector candidates",
                   "oracle_s": round(cpu_s, 1)},
    }
    bm.close()
    return res
