"""CPU checks that pin the graph oracle where the reference offers no golden vectors:
recall against f64 brute force, bounded == unbounded candidate queue, reduction-order models,
file round trips with the exact size equation, and the host-side error paths of the C ABI."""
import os
import struct

import numpy as np
import pytest

from conftest import make_data


@pytest.fixture(scope="module")
def small(orc):
    x, q = make_data(3000, 96, 21, nq=200)
    g = orc.Hnsw.build(x, M=16, ef_add=64, seed=21)
    return x, q, g


def _recall(keys, gt, k):
    return float(np.mean([len(set(keys[i].tolist()) & set(gt[i].tolist())) / k for i in range(keys.shape[0])]))


def test_hnsw_recall_and_monotone_ef(orc, small):
    x, q, g = small
    gt = orc.exact_f64(q, x, 10)
    r = [_recall(g.search(q, 10, ef)[0], gt, 10) for ef in (16, 64, 256)]
    assert r[0] <= r[1] + 1e-9 <= r[2] + 2e-9 and r[1] > 0.9 and r[2] > 0.97


def test_bounded_queue_equals_reference_heap(orc, small):
    x, q, g = small
    for ef in (10, 64, 200):
        a = g.search(q, 10, ef, lanes=8, next_cap=0)
        b = g.search(q, 10, ef, lanes=8, next_cap=max(ef, 10))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert np.array_equal(a[3][:, :3], b[3][:, :3])


def test_results_sorted_unique_and_counts(orc, small):
    x, q, g = small
    keys, dists, counts, stats = g.search(q, 10, 64)
    assert (counts == 10).all() and (np.diff(dists, axis=1) >= 0).all()
    assert all(len(set(r.tolist())) == 10 for r in keys)
    # distances are what the metric says: 1 - <q, x>
    d0 = 1.0 - np.einsum("ij,ij->i", q, x[keys[:, 0].astype(np.int64)])
    assert np.allclose(d0, dists[:, 0], atol=2e-6)
    assert (stats[:, 0] > 0).all() and (stats[:, 1] > 0).all()


def test_reduction_orders_agree_to_rounding(orc):
    rng = np.random.default_rng(0)
    for d in (70, 96, 100, 384, 768, 1536):
        a = rng.standard_normal(d).astype(np.float32)
        b = rng.standard_normal(d).astype(np.float32)
        ref = float(1.0 - np.dot(a.astype(np.float64), b.astype(np.float64)))
        for lanes in (0, 8, 32, -1):
            assert abs(orc.distance(a, b, 0, lanes) - ref) < 1e-4 * max(1.0, abs(ref))
        assert orc.distance(a, b, 2, 0) == max(0.0, orc.distance(a, b, 0, 0))
        l2 = float(np.sum((a.astype(np.float64) - b.astype(np.float64)) ** 2))
        for lanes in (0, 8, 32):
            assert abs(orc.distance(a, b, 1, lanes) - l2) < 1e-4 * l2


def test_sequential_fold_is_the_rust_sum(orc):
    """recompute.rs:137-139: a.iter().zip(b).map(|(x,y)| x*y).sum() — f32 mul then f32 add, in order."""
    rng = np.random.default_rng(1)
    a = rng.standard_normal(257).astype(np.float32)
    b = rng.standard_normal(257).astype(np.float32)
    s = np.float32(0)
    for x, y in zip(a, b):
        s = np.float32(s + np.float32(x * y))
    assert orc.distance(a, b, 0, 0) == float(np.float32(1.0) - s)


def test_usearch_file_roundtrip_and_size_equation(orc, small, tmp_path):
    x, q, g = small
    p = str(tmp_path / "a.index")
    g.save(p)
    info = g.info()
    n, d, M, M0 = info["n"], info["d"], info["M"], info["M0"]
    data = open(p, "rb").read()
    rows, cols = struct.unpack_from("<II", data, 0)
    assert (rows, cols) == (n, d * 4)
    head = 8 + rows * cols
    assert data[head:head + 7] == b"usearch" and data[head + 13:head + 17] == bytes([ord("i"), 11, 14, 15])
    size, conn, conn_base, max_level, entry = struct.unpack_from("<5Q", data, head + 64)
    assert (size, conn, conn_base) == (n, M, M0)
    levels = np.frombuffer(data, dtype="<i2", count=n, offset=head + 104)
    node_bytes = 10 + 4 * (1 + M0) + levels.astype(np.int64) * 4 * (1 + M)
    assert len(data) == head + 64 + 40 + 2 * n + int(node_bytes.sum())   # Appendix A.1 item 6
    g2 = orc.Hnsw.load(p, d)
    assert g2.info() == info
    a, b = g.search(q, 10, 64), g2.search(q, 10, 64)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # truncated and padded files are rejected
    open(str(tmp_path / "t.index"), "wb").write(data[:-5])
    with pytest.raises(RuntimeError):
        orc.Hnsw.load(str(tmp_path / "t.index"), d)
    open(str(tmp_path / "u.index"), "wb").write(data + b"\0\0")
    with pytest.raises(RuntimeError):
        orc.Hnsw.load(str(tmp_path / "u.index"), d)
    with pytest.raises(RuntimeError):
        orc.Hnsw.load(p, d + 1)


def test_vamana_recall_roundtrip(orc, tmp_path):
    x, q = make_data(2000, 64, 5, nq=100)
    g = orc.Vamana.build(x, R=24, L=40, alpha=1.2, seed=5)
    gt = orc.exact_f64(q, x, 10)
    keys, dists, counts, stats = g.search(q, 10, 64)
    assert _recall(keys, gt, 10) > 0.9 and (np.diff(dists, axis=1) >= 0).all() and (dists >= 0).all()
    p = str(tmp_path / "v.diskann")
    g.save(p)
    sz = os.path.getsize(p)
    assert sz == (1 << 20) + 2000 * 64 * 4 + 2000 * 24 * 4   # Appendix A.3 self-check
    g2 = orc.Vamana.load(p)
    assert g2.info() == g.info()
    assert np.array_equal(g2.search(q, 10, 64)[0], keys)
    a = g.search(q, 10, 64, lanes=8, next_cap=0)
    b = g.search(q, 10, 64, lanes=8, next_cap=64)
    assert np.array_equal(a[0], b[0])


def test_exact_scan_oracle_matches_f64_and_ties(orc):
    x, q = make_data(4000, 48, 8, nq=50)
    idx, sc, cnt = orc.exact_scan(q, x, 25, metric=0)
    gt = orc.exact_f64(q, x, 25)
    assert _recall(idx, gt, 25) > 0.999 and (np.diff(sc, axis=1) <= 0).all() and (cnt == 25).all()
    xx = np.concatenate([x[:100], x[:100]])
    idx, sc, cnt = orc.exact_scan(q[:5], xx, 10, metric=0)
    assert all(idx[i, j] + 100 == idx[i, j + 1] for i in range(5) for j in range(0, 10, 2))


# ---- C-ABI host-side error paths (no GPU needed: files are validated before the device is touched)
def test_abi_open_error_paths(pkg, orc, small, tmp_path):
    x, q, g = small
    base = str(tmp_path / "documents.leann")
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.HnswSearcher.load(base, 96)
    assert e.value.code == pkg.ERR_NOT_FOUND and "Index file not found" in e.value.message
    idx = base.replace(".leann", ".index")
    open(idx, "wb").write(b"IxHN" + b"\0" * 100)        # FAISS magic (compat.rs:15-38)
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.HnswSearcher.load(base, 96)
    assert e.value.code == pkg.ERR_FAISS_FORMAT and "FAISS" in e.value.message
    g.save(idx)
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.HnswSearcher.load(base, 97)
    assert e.value.code == pkg.ERR_DIM_MISMATCH
    data = bytearray(open(idx, "rb").read())
    head = 8 + 3000 * 96 * 4
    bad = bytearray(data); bad[head:head + 7] = b"notsear"
    open(idx, "wb").write(bad)
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.HnswSearcher.load(base, 96)
    assert e.value.code == pkg.ERR_BAD_FORMAT and "magic" in e.value.message
    open(idx, "wb").write(data[:-3])
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.HnswSearcher.load(base, 96)
    assert e.value.code == pkg.ERR_BAD_FORMAT
    open(idx, "wb").write(data + b"\0")
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.HnswSearcher.load(base, 96)
    assert e.value.code == pkg.ERR_BAD_FORMAT and "size equation" in e.value.message
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.DiskAnnSearcher.load(base, 96)
    assert e.value.code == pkg.ERR_NOT_FOUND
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.BackendType.load_searcher("faiss", base, 96)
    assert "Unknown backend" in e.value.message
    # a well-formed file with no GPU present must fail loudly, never fall back to the CPU
    open(idx, "wb").write(data)
    if pkg.device_count() == 0:
        with pytest.raises(pkg.LeannCudaError) as e:
            pkg.HnswSearcher.load(base, 96)
        assert e.value.code == pkg.ERR_CUDA and "no CPU fallback" in e.value.message


def test_handmade_index_and_hand_traced_search(orc, tmp_path):
    """Known-answer test derived by hand from the published usearch loop on a file written byte by byte from the published
    layout (tests/handmade.py): pins the reader, the greedy descent, the level-0 beam, the bounded `top` and the counters."""
    import handmade as H
    vecs, keys, levels, adj, q = H.ring_case()
    path = str(tmp_path / "ring.index")
    H.write_usearch_index(path, vecs, keys, levels, adj, M=2, M0=4, entry=0, max_level=1)
    g = orc.Hnsw.load(path, 2)
    assert g.info()["n"] == 8 and g.info()["max_level"] == 1 and g.info()["entry"] == 0
    for ef, k, exp in ((2, 2, H.EXPECT_EF2), (1, 1, H.EXPECT_EF1)):
        for cap in (0, ef):                          # unbounded heap and the kernel's bounded queue
            kk, dd, cc, st = g.search(q, k, ef, next_cap=cap)
            assert kk[0].tolist() == exp["keys"] and int(cc[0]) == k
            assert (int(st[0, 0]), int(st[0, 1]), int(st[0, 2])) == (exp["n_dist"], exp["hops0"], exp["hops_upper"])
            want = 1.0 - np.cos(np.deg2rad(70.0 - 10.0 * (np.array(exp["keys"]) - 100)))
            assert np.allclose(dd[0], want, atol=1e-6)


def test_handmade_diskann_file_and_hand_traced_search(orc, tmp_path):
    import handmade as H
    vecs, keys, levels, adj, q = H.ring_case()
    path = str(tmp_path / "ring.diskann")
    H.write_diskann(path, vecs, [a[0] for a in adj], R=2, medoid=0)
    g = orc.Vamana.load(path)
    assert g.info() == {"n": 8, "d": 2, "R": 2, "medoid": 0}
    for L, k, exp in ((2, 2, H.VAMANA_L2), (1, 1, H.VAMANA_L1)):
        for cap in (0, L):
            kk, dd, cc, st = g.search(q, k, L, next_cap=cap)
            assert kk[0].tolist() == exp["keys"], (L, kk)
            assert (int(st[0, 0]), int(st[0, 1])) == (exp["n_dist"], exp["hops0"]), (L, st[0])
            want = np.maximum(0.0, 1.0 - np.cos(np.deg2rad(70.0 - 10.0 * np.array(exp["keys"]))))
            assert np.allclose(dd[0], want, atol=1e-6)


def test_compat_switches_change_only_tie_handling(orc):
    """The recalled behaviours sit behind named switches (graph_oracle.cpp CompatBits). On data with exact duplicate rows the
    tie switches reorder equal-distance results and nothing else; on duplicate-free data they change nothing."""
    x, q = make_data(1500, 32, 3, nq=120)
    xd = x.copy()
    xd[500:1000] = xd[:500]                       # every one of 500 rows twice: equal distances everywhere
    base = orc.compat_default()
    try:
        for data, expect_change in ((xd, True), (x, False)):
            g = orc.Hnsw.build(data, M=8, ef_add=32, seed=3)
            orc.set_compat(base)
            k0, d0, _, _ = g.search(q, 10, 32)
            changed = False
            for bit in (orc.COMPAT_BITS["top_newcomer_before_equals"], orc.COMPAT_BITS["next_fifo_among_equals"]):
                orc.set_compat(base ^ bit)
                k1, d1, _, _ = g.search(q, 10, 32)
                changed |= not np.array_equal(k0, k1)
                assert np.mean(np.sort(d0, axis=1) == np.sort(d1, axis=1)) > 0.97   # only the order inside tie groups moves
            assert changed == expect_change
            # the stop rule (> vs >=) decides whether the current worst entry of `top` is still expanded: it may move a
            # boundary result even without ties, but never more than a few
            orc.set_compat(base ^ orc.COMPAT_BITS["usearch_stop_strict"])
            k1, d1, _, _ = g.search(q, 10, 32)
            assert np.mean(k0 == k1) > 0.95
    finally:
        orc.set_compat(base)
    assert orc.compat_flags() == base
