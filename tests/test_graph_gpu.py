"""GPU parity of K1/K1f (graph beam search) against the CPU oracle on the same index file and ef:
keys, f32 distance bits and the algorithmic-work counters must be identical."""
import numpy as np
import pytest

from conftest import make_data

pytestmark = pytest.mark.gpu


def _hnsw_case(orc, pkg, tmp_path, n, d, M, nq, efs, k=10, seed=7, mask_frac=None):
    x, q = make_data(n, d, seed, nq=nq)
    g = orc.Hnsw.build(x, M=M, ef_add=64, seed=seed)
    base = str(tmp_path / "documents.leann")
    g.save(base.replace(".leann", ".index"))
    s = pkg.HnswSearcher.load(base, d)
    assert len(s) == n and s.info()["M0"] == 2 * M
    lanes = pkg.reduction_lanes(d)
    mask = None
    if mask_frac is not None:
        rng = np.random.default_rng(seed + 1)
        mask = pkg.pack_mask(rng.random(n) < mask_frac)
    for ef in efs:
        cap = pkg.queue_capacity(max(ef, k), mask is not None)
        ok, od, oc, ost = g.search(q, k, ef, lanes=lanes, mask=mask, next_cap=cap)
        keys, dists, counts = s.search_batch(q, k, ef, mask=mask)
        assert np.array_equal(counts, oc)
        assert np.array_equal(keys, ok), f"ef={ef}: {np.mean(keys == ok):.4f} id agreement"
        assert np.array_equal(dists.view(np.uint32), od.view(np.uint32))
        # the bounded queue the kernel uses must equal the reference's unbounded heap here
        uk, ud, uc, ust = g.search(q, k, ef, lanes=lanes, mask=mask, next_cap=0)
        if mask is None:
            assert np.array_equal(uk, ok)
    return s, g, x, q


def test_hnsw_parity_d128(orc, pkg, tmp_path):
    _hnsw_case(orc, pkg, tmp_path, n=4000, d=128, M=16, nq=300, efs=(16, 64, 128))


def test_hnsw_parity_d768_m32(orc, pkg, tmp_path):
    s, g, x, q = _hnsw_case(orc, pkg, tmp_path, n=2500, d=768, M=32, nq=200, efs=(64, 256))
    # trait call: one query, complexity ignored -> ef = 64 (hnsw.rs:49,83)
    keys, dists = s.search(q[0], 5, 999)
    ok, od, _, _ = g.search(q[:1], 5, 64, lanes=pkg.reduction_lanes(768), next_cap=64)
    assert keys == [int(v) for v in ok[0]] and np.allclose(dists, od[0])


def test_hnsw_parity_odd_dims(orc, pkg, tmp_path):
    _hnsw_case(orc, pkg, tmp_path, n=1500, d=100, M=8, nq=100, efs=(32,))
    _hnsw_case(orc, pkg, tmp_path, n=1500, d=70, M=8, nq=100, efs=(32,))
    _hnsw_case(orc, pkg, tmp_path, n=1200, d=1536, M=8, nq=50, efs=(32,))


def test_hnsw_inline_mask(orc, pkg, tmp_path):
    _hnsw_case(orc, pkg, tmp_path, n=4000, d=128, M=16, nq=200, efs=(64,), mask_frac=0.25)
    _hnsw_case(orc, pkg, tmp_path, n=4000, d=128, M=16, nq=200, efs=(64,), mask_frac=0.02)


def test_hnsw_stats_counters(orc, pkg, tmp_path):
    import torch
    n, d, k, ef = 3000, 128, 10, 64
    x, q = make_data(n, d, 3, nq=128)
    g = orc.Hnsw.build(x, M=16, ef_add=64, seed=3)
    base = str(tmp_path / "documents.leann")
    g.save(base.replace(".leann", ".index"))
    s = pkg.HnswSearcher.load(base, d)
    qt = torch.from_numpy(q).cuda()
    stats = torch.zeros((q.shape[0], 4), dtype=torch.int64, device="cuda")
    keys, dists, counts = s.search_device(qt, k, ef, stats=stats)
    torch.cuda.synchronize()
    ok, od, oc, ost = g.search(q, k, ef, lanes=pkg.reduction_lanes(d), next_cap=ef)
    assert np.array_equal(keys.cpu().numpy().astype(np.uint64), ok)
    assert np.array_equal(stats.cpu().numpy().astype(np.uint64), ost)


def test_hnsw_small_and_k_gt_n(orc, pkg, tmp_path):
    x, q = make_data(5, 64, 1, nq=4)
    g = orc.Hnsw.build(x, M=4, ef_add=16, seed=1)
    base = str(tmp_path / "documents.leann")
    g.save(base.replace(".leann", ".index"))
    s = pkg.HnswSearcher.load(base, 64)
    keys, dists, counts = s.search_batch(q, 10, 16)
    ok, od, oc, _ = g.search(q, 10, 16, lanes=pkg.reduction_lanes(64), next_cap=16)
    assert np.array_equal(counts, oc) and np.array_equal(keys, ok) and (counts == 5).all()
    assert (keys[:, 5:] == np.uint64(2**64 - 1)).all() and np.isinf(dists[:, 5:]).all()


def test_vamana_parity(orc, pkg, tmp_path):
    n, d, R, L, k = 3000, 96, 32, 50, 10
    x, q = make_data(n, d, 11, nq=200)
    g = orc.Vamana.build(x, R=R, L=L, alpha=1.2, seed=11)
    base = str(tmp_path / "documents.leann")
    g.save(base.replace(".leann", ".diskann"))
    s = pkg.DiskAnnSearcher.load(base, d)
    assert s.info()["M0"] == R and s.info()["entry"] == g.info()["medoid"]
    lanes = pkg.reduction_lanes(d)
    for beam in (10, 50, 100):
        ok, od, oc, _ = g.search(q, k, beam, lanes=lanes, next_cap=max(beam, k))
        keys, dists, counts = s.search_batch(q, k, beam)
        assert np.array_equal(keys, ok) and np.array_equal(dists.view(np.uint32), od.view(np.uint32))
    # trait call honours complexity: beam = max(complexity, top_k)  (diskann.rs:54)
    keys, dists = s.search(q[0], 10, 3)
    ok, od, _, _ = g.search(q[:1], 10, 10, lanes=lanes, next_cap=10)
    assert keys == [int(v) for v in ok[0]]
    # L2 variant (BASELINE config C4) on un-normalised vectors
    x2, q2 = make_data(n, d, 12, nq=100, normalize=False)
    g2 = orc.Vamana.build(x2, R=R, L=L, alpha=1.2, seed=12, metric=orc.METRIC_L2SQ)
    base2 = str(tmp_path / "l2.leann")
    g2.save(base2.replace(".leann", ".diskann"))
    s2 = pkg.DiskAnnSearcher.load(base2, d, metric=pkg.METRIC_L2SQ)
    ok, od, oc, _ = g2.search(q2, k, 64, lanes=lanes, next_cap=64)
    keys, dists, counts = s2.search_batch(q2, k, 64)
    assert np.array_equal(keys, ok) and np.array_equal(dists.view(np.uint32), od.view(np.uint32))


def test_request_coalescing_native_threads(orc, pkg, tmp_path):
    """serve.rs shares one searcher between request threads (src/cli/serve.rs:84,260-311). A native driver
    (tests/native/coalesce_driver.cpp, plain C ABI) issues concurrent nq=1 calls on one handle: every answer
    must equal the batched call, and with coalescing on they must run in far fewer launches."""
    import json, os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "coalesce_driver")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(root, "tests/native/coalesce_driver.cpp"), "-o", exe,
                           "-L" + os.path.dirname(pkg.LIB_PATH), "-lleann_cuda", "-Wl,-rpath," + os.path.dirname(pkg.LIB_PATH)])
    n, d = 20000, 128
    x, _ = make_data(n, d, 19)
    s = pkg.HnswSearcher.build(x, graph_degree=16, complexity=64)
    base = str(tmp_path / "documents.leann")
    s.save(base)
    out = subprocess.check_output([exe, base, str(d), "64", "32"], text=True, timeout=240)
    r = json.loads(out.strip().splitlines()[-1])
    assert r["mismatch_or_fail"] == 0 and r["coalesced_requests"] == r["requests"]
    assert r["batches"] < r["requests"] / 4, r
    assert r["qps_coalesced"] > r["qps_uncoalesced"], r
    print(r)


def test_cpp_host_mirror(orc, pkg, tmp_path):
    """The header-only C++ host layer (leann_rs_b200/host/leann_cuda.hpp) mirrors the reference's Rust types; a
    native program runs the reference's bm25/filter unit-test assertions and a trait-level search through it."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_mirror_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", os.path.join(root, "tests/native/host_mirror_test.cpp"), "-o", exe,
                           "-L" + os.path.dirname(pkg.LIB_PATH), "-lleann_cuda", "-Wl,-rpath," + os.path.dirname(pkg.LIB_PATH)])
    n, d = 3000, 128
    x, _ = make_data(n, d, 5)
    g = orc.Hnsw.build(x, M=16, ef_add=64, seed=5)
    base = str(tmp_path / "documents.leann")
    g.save(base.replace(".leann", ".index"))
    q = np.sin(0.37 * np.arange(1, d + 1, dtype=np.float32)).astype(np.float32)
    q.tofile(str(tmp_path / "query.f32"))
    out = subprocess.check_output([exe, base, str(d), str(tmp_path / "query.f32"), str(tmp_path / "scratch.leann")], text=True, timeout=120)
    assert out.strip().endswith("OK")       # includes hnsw::build_index / add_to_index and MetadataColumns through the C++ mirror
    keys = [int(t) for t in [l for l in out.splitlines() if l.startswith("KEYS")][0].split()[1:]]
    ok, _, _, _ = g.search(q[None, :], 5, 64, lanes=pkg.reduction_lanes(d), next_cap=64)
    assert keys == [int(v) for v in ok[0]]


def test_small_batch_cta_per_query_matches_pool_and_oracle(orc, pkg, tmp_path):
    """Batches of at most two queries per SM run one 4-warp CTA per query (distance evaluations split across the warps);
    larger batches run one warp per query. Both must give the oracle's ids, distance bits and work counters."""
    import torch
    for d, M in ((768, 16), (96, 16)):
        n, k, ef = 5000, 10, 64
        x, q = make_data(n, d, 23, nq=900)
        g = orc.Hnsw.build(x, M=M, ef_add=64, seed=9)
        base = str(tmp_path / f"documents{d}.leann")
        g.save(base.replace(".leann", ".index"))
        s = pkg.HnswSearcher.load(base, d)
        ok, od, oc, ost = g.search(q, k, ef, lanes=pkg.reduction_lanes(d), next_cap=pkg.queue_capacity(ef, False))
        qt = torch.from_numpy(q).cuda()
        for lo, hi in ((0, 900), (0, 1), (1, 8), (8, 150), (150, 446)):     # pool kernel, then CTA-per-query batches
            st = torch.zeros((hi - lo, 4), dtype=torch.int64, device="cuda")
            keys, dists, counts = s.search_device(qt[lo:hi].contiguous(), k, ef, stats=st)
            assert np.array_equal(keys.cpu().numpy().view(np.uint64), ok[lo:hi]), (d, lo, hi)
            assert np.array_equal(dists.cpu().numpy().view(np.uint32), od[lo:hi].view(np.uint32)), (d, lo, hi)
            assert np.array_equal(st.cpu().numpy()[:, :3], ost[lo:hi, :3].astype(np.int64)), (d, lo, hi)
        # inline mask on a small batch
        bits = np.random.default_rng(1).random(n) < 0.3
        mk, md, mc = s.search_batch(q[:50], k, ef, mask=pkg.pack_mask(bits))
        wk, wd, wc, _ = g.search(q[:50], k, ef, lanes=pkg.reduction_lanes(d), mask=pkg.pack_mask(bits), next_cap=pkg.queue_capacity(ef, True))
        assert np.array_equal(mk, wk) and np.array_equal(md.view(np.uint32), wd.view(np.uint32))


def test_randomised_configs_match_oracle(orc, pkg, tmp_path):
    """Seeded sweep over shapes the fixed cases do not hit: odd dimensions on both sides of the 8-lane / 32-lane split, tiny
    and large degrees, ef below and above the queue sizes, k = 1 .. ef, iid data, exact duplicate vectors (distance ties),
    batch sizes on both sides of the CTA-per-query / warp-pool switch, inline masks, both backends and all three metrics."""
    rng = np.random.default_rng(2024)
    dims = [24, 64, 96, 200, 256, 260, 384, 1000]
    for trial in range(10):
        d = int(dims[trial % len(dims)])
        n = int(rng.integers(600, 4000))
        nq = int(rng.choice([1, 7, 120, 310, 500]))
        ef = int(rng.integers(8, 260))
        k = int(rng.integers(1, min(ef, 40) + 1))
        x, q = make_data(n, d, 100 + trial, kind="iid" if trial % 3 == 0 else "lowrank", nq=nq, normalize=trial % 4 != 1)
        dup = rng.integers(0, n, size=n // 10)
        x[rng.integers(0, n, size=n // 10)] = x[dup]                      # exact duplicates: equal distances
        mask_bits = (rng.random(n) < 0.35) if trial % 2 else None
        mask = None if mask_bits is None else pkg.pack_mask(mask_bits)
        base = str(tmp_path / f"t{trial}.leann")
        if trial % 2 == 0:
            M = int(rng.choice([4, 8, 16, 32]))
            g = orc.Hnsw.build(x, M=M, ef_add=int(rng.integers(16, 80)), seed=trial)
            g.save(base.replace(".leann", ".index"))
            s = pkg.HnswSearcher.load(base, d)
        else:
            R = int(rng.choice([8, 24, 64]))
            metric_o, metric_p = [(orc.METRIC_IP_CLAMP, pkg.METRIC_DEFAULT), (orc.METRIC_L2SQ, pkg.METRIC_L2SQ)][trial % 4 == 1]
            g = orc.Vamana.build(x, R=R, L=int(rng.integers(R, 90)), alpha=1.2, seed=trial, metric=metric_o)
            g.save(base.replace(".leann", ".diskann"))
            s = pkg.DiskAnnSearcher.load(base, d, metric=metric_p)
            g.set_metric(metric_o)
        cap = pkg.queue_capacity(max(ef, k), mask is not None)
        ok, od, oc, ost = g.search(q, k, ef, lanes=pkg.reduction_lanes(d), mask=mask, next_cap=cap)
        keys, dists, counts = s.search_batch(q, k, ef, mask=mask)
        ctx = (trial, d, n, nq, ef, k, mask is not None)
        assert np.array_equal(counts, oc), ctx
        assert np.array_equal(keys, ok), ctx
        assert np.array_equal(dists.view(np.uint32), od.view(np.uint32)), ctx
        if mask_bits is not None:
            valid = keys != np.uint64(2**64 - 1)
            assert mask_bits[keys[valid].astype(np.int64)].all(), ctx


def test_visited_set_modes_are_equivalent(orc, pkg, tmp_path):
    """The traversal's exact visited set has two representations: per-warp byte maps (default) and, for indexes whose byte
    maps would not fit, per-warp hash tables with a pooled byte-map spill. Forced through the tuning hook on a small index:
    a tiny table (every traversal spills through the pool), a roomy table, byte maps only — all must reproduce the oracle
    (ids, distance bits, work counters), for the warp pool and the CTA-per-query kernel, short and long rows."""
    import torch
    for d in (96, 768):
        n, k = 30000, 10
        x, q = make_data(n, d, 19, nq=500, normalize=False)
        base = str(tmp_path / f"v{d}.leann")
        pkg.DiskAnnSearcher.build(x, graph_degree=64, complexity=100, metric=pkg.METRIC_L2SQ).save(base)   # CPU build of this size is slow
        g = orc.Vamana.load(base.replace(".leann", ".diskann"))
        g.set_metric(orc.METRIC_L2SQ)
        s = pkg.DiskAnnSearcher.load(base, d, metric=pkg.METRIC_L2SQ)
        qt = torch.from_numpy(q).cuda()
        for ef in (32, 200):
            ok, od, oc, ost = g.search(q, k, ef, lanes=pkg.reduction_lanes(d), next_cap=pkg.queue_capacity(ef, False))
            for mode in (0, 1024, 1, 65536, 2, 3, 4, 5):   # 2-5: shared-memory tables (short rows, ef <= 128); 3, 5 = every traversal overflows / spills
                s.set_visited_hash(mode)
                for lo, hi in ((0, 500), (0, 40)):            # warp pool, CTA per query
                    stats = torch.zeros((hi - lo, 4), dtype=torch.int64, device="cuda")
                    keys, dists, counts = s.search_device(qt[lo:hi].contiguous(), k, ef, stats=stats)
                    assert np.array_equal(keys.cpu().numpy().view(np.uint64), ok[lo:hi]), (d, ef, mode, hi)
                    assert np.array_equal(dists.cpu().numpy().view(np.uint32), od[lo:hi].view(np.uint32)), (d, ef, mode, hi)
                    assert np.array_equal(stats.cpu().numpy()[:, :3], ost[lo:hi, :3].astype(np.int64)), (d, ef, mode, hi)
        assert ost[:, 0].max() > 768          # the 1024-slot tables did have to spill
        with pytest.raises(pkg.LeannCudaError):
            s.set_visited_hash(100)
        s.set_visited_hash(0)


def test_row_ring_form_is_bit_identical(pkg, tmp_path):
    """A/B form of the short-row traversal: rows of a hop as bulk async copies through a shared-memory ring
    (LEANN_CUDA_RING=1, read once per process, hence the child process). Same index file, same queries: keys, distance bits
    and work counters must equal the default form's, for both visited-set forms the register-list kernel supports."""
    import subprocess, sys, json, zlib, os
    import torch
    n, d, k = 40000, 96, 10
    x, q = make_data(n, d, 29, nq=600, normalize=False)
    base = str(tmp_path / "ring.leann")
    pkg.DiskAnnSearcher.build(x, graph_degree=64, complexity=100, metric=pkg.METRIC_L2SQ).save(base)
    np.save(str(tmp_path / "q.npy"), q)
    prog = (
        "import sys, json, zlib, numpy as np, torch\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "import leann_rs_b200 as P\n"
        f"s = P.DiskAnnSearcher.load({base!r}, {d}, metric=P.METRIC_L2SQ)\n"
        f"q = torch.from_numpy(np.load({str(tmp_path / 'q.npy')!r})).cuda()\n"
        "out = {}\n"
        "for mode in (0, 65536):\n"
        "    s.set_visited_hash(mode)\n"
        "    for ef in (32, 100):\n"
        "        st = torch.zeros((q.shape[0], 4), dtype=torch.int64, device='cuda')\n"
        f"        kk, dd, cc = s.search_device(q, {k}, ef, stats=st)\n"
        "        out[f'{mode}/{ef}'] = [zlib.crc32(kk.cpu().numpy().tobytes()), zlib.crc32(dd.cpu().numpy().tobytes()), zlib.crc32(st[:, :3].cpu().numpy().tobytes())]\n"
        "print('RESULT ' + json.dumps(out))\n")
    res = {}
    for ring in ("0", "1"):
        env = dict(os.environ, LEANN_CUDA_RING=ring)
        r = subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res[ring] = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert res["0"] == res["1"], res
    assert len({tuple(v) for kq, v in res["0"].items() if kq.endswith("/100")}) == 1      # visited-set forms agree as well


def test_handmade_index_hand_traced_on_gpu(pkg, tmp_path):
    """The hand-traced known answers of tests/handmade.py (file written byte by byte from the published usearch layout,
    search traced by hand from the published loop) through the product: reader, descent, beam, counters."""
    import torch
    import handmade as H
    vecs, keys, levels, adj, q = H.ring_case()
    base = str(tmp_path / "ring.leann")
    H.write_usearch_index(base.replace(".leann", ".index"), vecs, keys, levels, adj, M=2, M0=4, entry=0, max_level=1)
    s = pkg.HnswSearcher.load(base, 2)
    assert len(s) == 8 and s.info()["max_level"] == 1
    qt = torch.from_numpy(np.repeat(q, 400, axis=0)).cuda()       # 400 copies: warp-pool kernel; 1 copy: CTA-per-query kernel
    for ef, k, exp in ((2, 2, H.EXPECT_EF2), (1, 1, H.EXPECT_EF1)):
        for nq in (400, 1):
            st = torch.zeros((nq, 4), dtype=torch.int64, device="cuda")
            kk, dd, cc = s.search_device(qt[:nq].contiguous(), k, ef, stats=st)
            assert kk.cpu().numpy()[0].tolist() == exp["keys"] and (kk == kk[0]).all()
            assert st.cpu().numpy()[0, :3].tolist() == [exp["n_dist"], exp["hops0"], exp["hops_upper"]]
            want = 1.0 - np.cos(np.deg2rad(70.0 - 10.0 * (np.array(exp["keys"]) - 100)))
            assert np.allclose(dd.cpu().numpy()[0], want, atol=1e-6)
    # the trait call (hnsw.rs:79-88): complexity ignored, ef = max(64, k) -> every node is reached, nearest first
    keys5, dists5 = s.search(q[0], 5, 1)
    assert keys5 == [107, 106, 105, 104, 103]


def test_handmade_diskann_hand_traced_on_gpu(pkg, tmp_path):
    """Hand-written `.diskann` file + hand-traced diskann-rs beam search (tests/handmade.py) through the product."""
    import torch
    import handmade as H
    vecs, keys, levels, adj, q = H.ring_case()
    base = str(tmp_path / "ring.leann")
    H.write_diskann(base.replace(".leann", ".diskann"), vecs, [a[0] for a in adj], R=2, medoid=0)
    s = pkg.DiskAnnSearcher.load(base, 2)
    assert len(s) == 8
    qt = torch.from_numpy(np.repeat(q, 400, axis=0)).cuda()
    for L, k, exp in ((2, 2, H.VAMANA_L2), (1, 1, H.VAMANA_L1)):
        for nq in (400, 1):
            st = torch.zeros((nq, 4), dtype=torch.int64, device="cuda")
            kk, dd, cc = s.search_device(qt[:nq].contiguous(), k, L, stats=st)
            assert kk.cpu().numpy()[0].tolist() == exp["keys"] and (kk == kk[0]).all()
            assert st.cpu().numpy()[0, :2].tolist() == [exp["n_dist"], exp["hops0"]]
    # trait call (diskann.rs:47-62): beam = max(complexity, top_k)
    k3, d3 = s.search(q[0], 3, 1)
    assert k3 == [7, 6, 5]


def test_workspace_sized_once_and_streams_chain(orc, pkg):
    """(i) the host-buffer path must not size the traversal workspace with other parameters than the launch
    (a repeated identical search performs no reallocation); (ii) device-pointer calls on different streams of one handle
    share that workspace and must execute in call order (the library chains them with an event)."""
    import torch
    n, d, k, ef = 6000, 96, 10, 64
    x, q = make_data(n, d, 21, nq=3000)
    s = pkg.HnswSearcher.build(x, graph_degree=16, complexity=64, seed=3)
    k0, d0, c0 = s.search_batch(q, k, ef)
    r0 = s.workspace_stats()["reallocs"]
    assert r0 >= 1
    for _ in range(3):
        k1, d1, _ = s.search_batch(q, k, ef)
        assert np.array_equal(k1, k0) and np.array_equal(d1.view(np.uint32), d0.view(np.uint32))
    assert s.workspace_stats()["reallocs"] == r0
    # forced large-index mode (hash tables): same rule
    s.set_visited_hash(4096)
    s.search_batch(q, k, ef)
    r1 = s.workspace_stats()["reallocs"]
    assert s.workspace_stats()["large_mode"] == 1
    k2, _, _ = s.search_batch(q, k, ef)
    assert s.workspace_stats()["reallocs"] == r1 and np.array_equal(k2, k0)
    s.set_visited_hash(0)
    # two streams, interleaved launches, no host synchronisation in between
    qt = torch.from_numpy(q).cuda()
    st = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    torch.cuda.synchronize()
    for i in range(6):
        with torch.cuda.stream(st[i % 2]):
            outs.append(s.search_device(qt, k, ef))
    torch.cuda.synchronize()
    for kk, dd, cc in outs:
        assert np.array_equal(kk.cpu().numpy().astype(np.uint64), k0)
        assert np.array_equal(dd.cpu().numpy().view(np.uint32), d0.view(np.uint32))
    s.close()


def test_reduction_orders_agree_at_100k(orc, pkg, tmp_path):
    """The graph parity tests run the oracle in `lanes` mode (a bit-exact model of the kernel's reduction tree). The
    reference's usearch uses SimSIMD's order instead; the two differ in the last ulp of a distance, which can only swap
    near-ties. On a 100k-vector index the ids under both orders (and under the plain sequential fold) must agree >= 0.999."""
    n, d, k, ef = 100_000, 128, 10, 64
    x, q = make_data(n, d, 41, nq=2000)
    s = pkg.HnswSearcher.build(x, graph_degree=16, complexity=64, seed=4)
    base = str(tmp_path / "documents.leann")
    s.save(base)
    g = orc.Hnsw.load(base.replace(".leann", ".index"), d)
    lanes = pkg.reduction_lanes(d)
    kl, dl, _, _ = g.search(q, k, ef, lanes=lanes, next_cap=pkg.queue_capacity(ef, False), nthreads=8)
    gk, gd, _ = s.search_batch(q, k, ef)
    assert np.array_equal(gk, kl) and np.array_equal(gd.view(np.uint32), dl.view(np.uint32))
    for other in (-1, 0):     # SimSIMD-shaped, sequential
        ko, do_, _, _ = g.search(q, k, ef, lanes=other, next_cap=0, nthreads=8)
        agree = float(np.mean(ko == kl))
        assert agree >= 0.999, (other, agree)
        assert float(np.max(np.abs(do_ - dl))) < 1e-6
    s.close()
