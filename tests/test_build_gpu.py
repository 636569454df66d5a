"""GPU index construction: the file it writes is a valid usearch `.index`, searching
it on the GPU equals the oracle searching the same file, and graph quality matches the sequential
CPU build of the same data."""
import numpy as np
import pytest

from conftest import make_data

pytestmark = pytest.mark.gpu


def _recall(keys, gt, k):
    return float(np.mean([len(set(keys[i].tolist()) & set(gt[i].tolist())) / k for i in range(keys.shape[0])]))


def test_gpu_hnsw_build_quality_and_file(orc, pkg, tmp_path):
    n, d, M, k = 20000, 128, 16, 10
    x, q = make_data(n, d, 31, nq=300)
    s = pkg.HnswSearcher.build(x, graph_degree=M, complexity=64, seed=5)
    info = s.info()
    assert info["n"] == n and info["M"] == M and info["M0"] == 2 * M and info["max_level"] >= 1
    gt = orc.exact_f64(q, x, k)
    keys, dists, counts = s.search_batch(q, k, 64)
    r_gpu = _recall(keys, gt, k)
    g_cpu = orc.Hnsw.build(x, M=M, ef_add=64, seed=5)
    r_cpu = _recall(g_cpu.search(q, k, 64)[0], gt, k)
    assert r_gpu > 0.9 and r_gpu > r_cpu - 0.03, (r_gpu, r_cpu)
    # the written file obeys the format (oracle reader checks every field and the size equation)
    base = str(tmp_path / "documents.leann")
    s.save(base)
    g = orc.Hnsw.load(base.replace(".leann", ".index"), d)
    assert g.info()["n"] == n and g.info()["max_level"] == info["max_level"] and g.info()["entry"] == info["entry"]
    ok, od, oc, _ = g.search(q, k, 64, lanes=pkg.reduction_lanes(d), next_cap=64)
    assert np.array_equal(keys, ok) and np.array_equal(dists.view(np.uint32), od.view(np.uint32))
    # and re-opens through the product reader with identical results
    s2 = pkg.HnswSearcher.load(base, d)
    k2, d2, _ = s2.search_batch(q, k, 64)
    assert np.array_equal(k2, keys)


def test_gpu_hnsw_build_tiny_and_device_input(orc, pkg):
    import torch
    for n in (1, 2, 17):
        x, q = make_data(n, 64, n, nq=3)
        s = pkg.HnswSearcher.build(x, graph_degree=4, complexity=16)
        keys, dists, counts = s.search_batch(q, 5, 16)
        assert (counts == min(n, 5)).all()
        gt = orc.exact_f64(q, x, min(n, 5))
        assert all(set(keys[i, :counts[i]].tolist()) == set(gt[i].tolist()) for i in range(3))
    x, q = make_data(5000, 768, 3, nq=100)
    s = pkg.HnswSearcher.build(torch.from_numpy(x).cuda(), graph_degree=32, complexity=64)
    gt = orc.exact_f64(q, x, 10)
    assert _recall(s.search_batch(q, 10, 64)[0], gt, 10) > 0.9


def test_gpu_vamana_build_quality_and_file(orc, pkg, tmp_path):
    n, d, R, L, k = 20000, 96, 32, 64, 10
    x, q = make_data(n, d, 41, nq=300)
    s = pkg.DiskAnnSearcher.build(x, graph_degree=R, complexity=L, alpha=1.2)
    info = s.info()
    assert info["n"] == n and info["M0"] == R and info["max_level"] == 0
    gt = orc.exact_f64(q, x, k)
    keys, dists, counts = s.search_batch(q, k, L)
    r_gpu = _recall(keys, gt, k)
    g_cpu = orc.Vamana.build(x[:6000], R=R, L=L, alpha=1.2, seed=3)     # CPU reference build is slow: smaller set
    r_cpu = _recall(g_cpu.search(q, k, L)[0], orc.exact_f64(q, x[:6000], k), k)
    assert r_gpu > 0.9 and r_gpu > r_cpu - 0.05, (r_gpu, r_cpu)
    base = str(tmp_path / "documents.leann")
    s.save(base)
    g = orc.Vamana.load(base.replace(".leann", ".diskann"))      # checks header + size equation
    assert g.info() == {"n": n, "d": d, "R": R, "medoid": info["entry"]}
    ok, od, oc, _ = g.search(q, k, L, lanes=pkg.reduction_lanes(d), next_cap=L)
    assert np.array_equal(keys, ok) and np.array_equal(dists.view(np.uint32), od.view(np.uint32))
    s2 = pkg.DiskAnnSearcher.load(base, d)
    assert np.array_equal(s2.search_batch(q, k, L)[0], keys)
    # L2 metric on un-normalised data (BASELINE C4)
    x2, q2 = make_data(n, d, 42, nq=200, normalize=False)
    s3 = pkg.DiskAnnSearcher.build(x2, graph_degree=R, complexity=L, metric=pkg.METRIC_L2SQ)
    gt2 = orc.exact_f64(q2, x2, k, metric=1)
    assert _recall(s3.search_batch(q2, k, L)[0], gt2, k) > 0.9


def test_gpu_hnsw_add_to_index(orc, pkg, tmp_path):
    """hnsw::add_to_index (hnsw.rs:142-191): append to a saved index, keys continue from start_id, the updated
    file is a valid usearch `.index`, GPU search on it equals the oracle on the same file, and recall over the
    whole set matches a one-shot build."""
    n0, m, d, M, k = 12000, 8000, 128, 16, 10
    x, q = make_data(n0 + m, d, 77, nq=300)
    base = str(tmp_path / "documents.leann")
    s0 = pkg.HnswSearcher.build(x[:n0], graph_degree=M, complexity=64, seed=3)
    s0.save(base)
    s0.close()
    pkg.add_to_index(x[n0:], base, d, start_id=n0)          # load -> add -> save, as the reference does
    s = pkg.HnswSearcher.load(base, d)
    assert len(s) == n0 + m and s.info()["M"] == M
    gt = orc.exact_f64(q, x, k)
    keys, dists, counts = s.search_batch(q, k, 64)
    r_add = _recall(keys, gt, k)
    full = pkg.HnswSearcher.build(x, graph_degree=M, complexity=64, seed=3)
    r_full = _recall(full.search_batch(q, k, 64)[0], gt, k)
    assert r_add > 0.9 and r_add > r_full - 0.03, (r_add, r_full)
    assert (keys[:, 0] < n0 + m).all() and (keys >= n0).any()           # appended rows are reachable
    g = orc.Hnsw.load(base.replace(".leann", ".index"), d)                 # format + size equation
    assert g.info()["n"] == n0 + m
    ok, od, oc, _ = g.search(q, k, 64, lanes=pkg.reduction_lanes(d), next_cap=64)
    assert np.array_equal(keys, ok) and np.array_equal(dists.view(np.uint32), od.view(np.uint32))
    # non-contiguous keys: start_id beyond the current size
    s.add(x[:100] * np.float32(2.0), start_id=10**6)      # dot with the doubled copy beats the original
    k2, _, _ = s.search_batch(x[:20], 3, 64)
    assert len(s) == n0 + m + 100 and (k2[:, 0] == 10**6 + np.arange(20)).mean() > 0.9
    # adding to an empty-then-built tiny index and to a Vamana handle
    t = pkg.HnswSearcher.build(x[:1], graph_degree=4, complexity=16)
    t.add(x[1:40])
    kk, _, cc = t.search_batch(q[:4], 5, 16)
    assert len(t) == 40 and (cc == 5).all()
    v = pkg.DiskAnnSearcher.build(x[:500], graph_degree=8, complexity=16)
    with pytest.raises(pkg.LeannCudaError):
        pkg.lib().leann_cuda_hnsw_add  # symbol exists
        pkg.HnswSearcher.add(v, x[:3])


def test_gpu_builders_in_large_index_mode(orc, pkg, monkeypatch):
    """The builders' insert-time searches use the same visited set as K1; forcing the large-index representation (hash tables +
    pooled byte-map spill) must give a graph of the same quality as the byte maps (batched insertion resolves concurrent
    reverse links by atomics, so two builds are not bit-identical; recall is the invariant)."""
    n, d, k = 12000, 96, 10
    xn, qn = make_data(n, d, 51, nq=200)                       # unit vectors, inner product (HNSW)
    xl, ql = make_data(n, d, 52, nq=200, normalize=False)      # raw vectors, squared L2 (Vamana)
    gt_n, gt_l = orc.exact_f64(qn, xn, k), orc.exact_f64(ql, xl, k, metric=1)

    def build_pair():
        h = pkg.HnswSearcher.build(xn, graph_degree=16, complexity=64, seed=7)
        v = pkg.DiskAnnSearcher.build(xl, graph_degree=32, complexity=64, metric=pkg.METRIC_L2SQ)
        return h, v, _recall(h.search_batch(qn, k, 64)[0], gt_n, k), _recall(v.search_batch(ql, k, 64)[0], gt_l, k)

    h0, v0, rh0, rv0 = build_pair()
    monkeypatch.setenv("LEANN_CUDA_FORCE_LARGE_INDEX_MODE", "1")
    h1, v1, rh1, rv1 = build_pair()
    monkeypatch.delenv("LEANN_CUDA_FORCE_LARGE_INDEX_MODE")
    assert h0.info()["n"] == h1.info()["n"] == n and v0.info() == v1.info()
    assert min(rh0, rh1, rv0, rv1) > 0.9, (rh0, rh1, rv0, rv1)
    assert abs(rh0 - rh1) < 0.02 and abs(rv0 - rv1) < 0.02, (rh0, rh1, rv0, rv1)
