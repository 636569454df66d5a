"""GPU parity of K2 (exact scan + fused top-k) and K4 (shard merge) against the oracle restatement of
src/index/recompute.rs:96-110. IDs must match except for ties within 1e-5 relative score
(BASELINE.json north_star); scores within 1e-5 relative."""
import numpy as np
import pytest

from conftest import make_data

pytestmark = pytest.mark.gpu

REL = 1e-5


def _check_topk(keys, scores, okeys, oscores, descending):
    assert keys.shape == okeys.shape
    scale = np.maximum(np.abs(oscores), 1e-6)
    assert np.all(np.abs(scores - oscores) <= REL * np.maximum(scale, 1.0) + 1e-6)
    bad = 0
    for i in range(keys.shape[0]):
        if np.array_equal(keys[i], okeys[i]):
            continue
        # every mismatching id must sit in a tie group (score within REL of the oracle's at that rank)
        mine = dict(zip(keys[i].tolist(), scores[i].tolist()))
        for j, kk in enumerate(okeys[i].tolist()):
            if keys[i, j] == kk:
                continue
            ref = oscores[i, j]
            assert abs(scores[i, j] - ref) <= REL * max(abs(ref), 1.0), (i, j)
            bad += 1
    return bad


@pytest.mark.parametrize("metric_name,k,d,n", [("dot", 100, 384, 20000), ("dot", 10, 768, 6000), ("l2", 10, 96, 9000), ("ip", 17, 70, 5000)])
def test_exact_scan_parity(orc, pkg, metric_name, k, d, n):
    x, q = make_data(n, d, 5, nq=150, normalize=(metric_name != "l2"))
    pm = {"dot": pkg.METRIC_DOT_DESC, "l2": pkg.METRIC_L2SQ, "ip": pkg.METRIC_IP}[metric_name]
    om = {"dot": 0, "l2": 1, "ip": 2}[metric_name]
    s = pkg.FlatSearcher.from_vectors(x, metric=pm)
    keys, scores, counts = s.search_batch(q, k, 0)
    oi, osc, oc = orc.exact_scan(q, x, k, metric=om, nthreads=8)
    assert np.array_equal(counts, oc)
    _check_topk(keys, scores, oi, osc, metric_name == "dot")
    # recall against f64 brute force
    gt = orc.exact_f64(q, x, k, metric=1 if metric_name == "l2" else 0)
    rec = np.mean([len(set(keys[i].tolist()) & set(gt[i].tolist())) / k for i in range(len(q))])
    assert rec > 0.999


def test_exact_scan_mask_prefilter_and_small(orc, pkg):
    x, q = make_data(3000, 64, 9, nq=40)
    rng = np.random.default_rng(0)
    bits = rng.random(3000) < 0.1
    mask = pkg.pack_mask(bits)
    s = pkg.FlatSearcher.from_vectors(x)
    keys, scores, counts = s.search_batch(q, 20, 0, mask=mask)
    oi, osc, oc = orc.exact_scan(q, x, 20, metric=0, mask=orc.pack_mask(bits))
    assert np.array_equal(counts, oc)
    _check_topk(keys, scores, oi, osc, True)
    assert bits[keys[keys != np.uint64(2**64 - 1)].astype(np.int64)].all()
    # k > n
    s2 = pkg.FlatSearcher.from_vectors(x[:7])
    keys, scores, counts = s2.search_batch(q, 10, 0)
    assert (counts == 7).all() and (keys[:, 7:] == np.uint64(2**64 - 1)).all()


def test_exact_scan_duplicates_tie_order(orc, pkg):
    """Equal scores rank by ascending index (stable sort over enumerate(), recompute.rs:106)."""
    x, q = make_data(500, 64, 2, nq=10)
    x = np.concatenate([x, x, x])  # every vector three times
    s = pkg.FlatSearcher.from_vectors(x)
    keys, scores, counts = s.search_batch(q, 30, 0)
    oi, osc, oc = orc.exact_scan(q, x, 30, metric=0)
    # duplicates have bit-identical scores on both sides, so order inside a group is by index
    for i in range(len(q)):
        for j in range(0, 30, 3):
            grp = keys[i, j:j + 3].astype(np.int64)
            assert grp[0] < grp[1] < grp[2] and grp[1] == grp[0] + 500 and grp[2] == grp[0] + 1000


def test_exact_scan_sorted_database_overflow_path(orc, pkg):
    """Rows sorted by ascending similarity to the query make every later chunk beat the threshold:
    exercises the candidate-overflow re-run."""
    d, n = 32, 60000
    rng = np.random.default_rng(4)
    qv = rng.standard_normal(d).astype(np.float32)
    qv /= np.linalg.norm(qv)
    t = np.linspace(-1, 1, n, dtype=np.float32)[:, None]
    x = (t * qv[None, :] + 0.001 * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)
    q = np.repeat(qv[None, :], 4, axis=0)
    s = pkg.FlatSearcher.from_vectors(x)
    keys, scores, counts = s.search_batch(q, 10, 0)
    oi, osc, oc = orc.exact_scan(q, x, 10, metric=0)
    _check_topk(keys, scores, oi, osc, True)


def test_topk_merge(pkg):
    import torch
    g, nq, k = 8, 64, 10
    rng = np.random.default_rng(1)
    d = np.sort(rng.random((g, nq, k)).astype(np.float32), axis=2)
    keys = rng.integers(0, 1 << 40, size=(g, nq, k)).astype(np.int64)
    # ragged: some shards return fewer than k
    d[3, :, 6:] = np.inf
    keys[3, :, 6:] = -1
    mk, md, mc = pkg.topk_merge_device(torch.from_numpy(keys).cuda(), torch.from_numpy(d).cuda(), descending=False)
    mk, md, mc = mk.cpu().numpy(), md.cpu().numpy(), mc.cpu().numpy()
    for i in range(nq):
        flat_d = d[:, i, :].reshape(-1)
        flat_k = keys[:, i, :].reshape(-1)
        order = np.argsort(flat_d, kind="stable")[:k]
        assert np.array_equal(md[i], flat_d[order]) and np.array_equal(mk[i], flat_k[order])
    assert (mc == k).all()
    # descending
    dd = -d
    dd[np.isinf(dd)] = -np.inf
    mk2, md2, _ = pkg.topk_merge_device(torch.from_numpy(keys).cuda(), torch.from_numpy(dd).cuda(), descending=True)
    assert np.array_equal(mk2.cpu().numpy(), mk)


# ---- tensor-core path (tcgen05 first pass + fp32 re-rank): n >= 16384, d >= 64, any batch size -----------------
@pytest.mark.parametrize("metric_name,k,d,n,nq", [("dot", 100, 384, 50000, 200), ("dot", 10, 768, 30000, 130), ("ip", 10, 100, 40000, 96),
                                                   ("dot", 1, 64, 20000, 64), ("dot", 1000, 128, 70000, 257),
                                                   # query tile resident up to 7 k-blocks (d <= 448), streamed beyond; ragged query pairs
                                                   ("dot", 10, 448, 30000, 300), ("ip", 10, 450, 30000, 513), ("dot", 5, 520, 20000, 64),
                                                   # single query and a handful: the query tile is mostly TMA zero fill
                                                   ("dot", 100, 384, 40000, 1), ("ip", 10, 768, 30000, 3), ("dot", 7, 128, 20000, 17)])
def test_exact_scan_tensor_path_parity(orc, pkg, metric_name, k, d, n, nq):
    x, q = make_data(n, d, 17, nq=nq)
    pm = {"dot": pkg.METRIC_DOT_DESC, "ip": pkg.METRIC_IP}[metric_name]
    om = {"dot": 0, "ip": 2}[metric_name]
    s = pkg.FlatSearcher.from_vectors(x, metric=pm)
    keys, scores, counts = s.search_batch(q, k, 0)
    oi, osc, oc = orc.exact_scan(q, x, k, metric=om, nthreads=8)
    assert np.array_equal(counts, oc)
    _check_topk(keys, scores, oi, osc, metric_name == "dot")
    # second call reuses the bf16 copy
    keys2, scores2, _ = s.search_batch(q, k, 0)
    assert np.array_equal(keys, keys2) and np.array_equal(scores, scores2)


def test_exact_scan_tensor_path_unnormalised_mask_and_sorted(orc, pkg):
    # un-normalised vectors with widely varying norms: the error margin scales with |q| * max|x|
    x, q = make_data(30000, 128, 23, nq=80, normalize=False)
    x *= np.linspace(0.1, 30.0, 30000, dtype=np.float32)[:, None]
    rng = np.random.default_rng(5)
    bits = rng.random(30000) < 0.03
    s = pkg.FlatSearcher.from_vectors(x)
    keys, scores, counts = s.search_batch(q, 20, 0, mask=pkg.pack_mask(bits))
    oi, osc, oc = orc.exact_scan(q, x, 20, metric=0, mask=orc.pack_mask(bits), nthreads=8)
    assert np.array_equal(counts, oc)
    _check_topk(keys, scores, oi, osc, True)
    # rows sorted by ascending similarity: every later chunk beats the threshold (overflow re-run on the tensor path)
    d, n = 64, 80000
    qv = rng.standard_normal(d).astype(np.float32)
    qv /= np.linalg.norm(qv)
    t = np.linspace(-1, 1, n, dtype=np.float32)[:, None]
    xs = (t * qv[None, :] + 0.001 * rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)
    qs = np.repeat(qv[None, :], 64, axis=0) + 0.01 * rng.standard_normal((64, d)).astype(np.float32)
    s2 = pkg.FlatSearcher.from_vectors(xs)
    keys, scores, counts = s2.search_batch(qs, 10, 0)
    oi, osc, oc = orc.exact_scan(qs, xs, 10, metric=0, nthreads=8)
    _check_topk(keys, scores, oi, osc, True)


def test_randomised_exact_scan_matches_oracle(orc, pkg):
    """Seeded sweep: sizes on both sides of every path switch (f32 tiles below 16384 rows or d < 64, tensor path with the query
    tile resident up to d = 448 and streamed beyond), ragged batches, k up to several hundred, all metrics, pre-filter masks,
    exact duplicate rows (ties resolve by ascending index)."""
    rng = np.random.default_rng(77)
    dims = [8, 64, 70, 128, 384, 448, 456, 1000]
    for trial in range(12):
        d = int(dims[trial % len(dims)])
        n = int(rng.choice([150, 5000, 16383, 16384, 40000, 70000]))
        nq = int(rng.choice([1, 5, 64, 129, 300]))
        k = int(rng.integers(1, 200))
        metric_name = ["dot", "ip", "l2"][trial % 3]
        x, q = make_data(n, d, 300 + trial, nq=nq, normalize=(metric_name != "l2"))
        x[rng.integers(0, n, size=n // 20)] = x[rng.integers(0, n, size=n // 20)]
        bits = (rng.random(n) < 0.2) if trial % 4 == 3 else None
        pm = {"dot": pkg.METRIC_DOT_DESC, "l2": pkg.METRIC_L2SQ, "ip": pkg.METRIC_IP}[metric_name]
        om = {"dot": 0, "l2": 1, "ip": 2}[metric_name]
        s = pkg.FlatSearcher.from_vectors(x, metric=pm)
        keys, scores, counts = s.search_batch(q, k, 0, mask=None if bits is None else pkg.pack_mask(bits))
        oi, osc, oc = orc.exact_scan(q, x, k, metric=om, mask=None if bits is None else orc.pack_mask(bits), nthreads=8)
        ctx = (trial, d, n, nq, k, metric_name, bits is not None)
        assert np.array_equal(counts, oc), ctx
        valid = np.arange(k)[None, :] < oc[:, None]
        assert np.array_equal(keys[~valid], oi[~valid]), ctx
        kk, ss, ok_, os_ = keys.copy(), scores.copy(), oi.copy(), osc.copy()
        for i in range(nq):                      # compare only the filled part of each row
            c = int(oc[i])
            _check_topk(kk[i:i + 1, :c], ss[i:i + 1, :c], ok_[i:i + 1, :c], os_[i:i + 1, :c], metric_name == "dot")
        s.close()
