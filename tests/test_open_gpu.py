"""leann_cuda_open streams index files from disk to HBM (csrc/open_stream.cu) and can reuse a cached device layout
(`<base>.cuda-layout`). The streamed paths must load exactly what the whole-file readers describe, keep
every format check, and never trust a stale cache."""
import os
import struct

import numpy as np
import pytest

from conftest import make_data

pytestmark = pytest.mark.gpu


def _search_equal(pkg, s, g, q, d, k=10, ef=48):
    ok, od, oc, _ = g.search(q, k, ef, lanes=pkg.reduction_lanes(d), next_cap=pkg.queue_capacity(ef, False))
    keys, dists, counts = s.search_batch(q, k, ef)
    assert np.array_equal(keys, ok) and np.array_equal(dists.view(np.uint32), od.view(np.uint32)) and np.array_equal(counts, oc)
    return keys


@pytest.mark.parametrize("d", [70, 128])     # 70: rows are padded on the device while streaming
def test_layout_cache_round_trip(orc, pkg, tmp_path, d):
    n = 6000
    x, q = make_data(n, d, 31, nq=100)
    g = orc.Hnsw.build(x, M=16, ef_add=64, seed=2)
    base = str(tmp_path / "documents.leann")
    idx_file, cache = base.replace(".leann", ".index"), base.replace(".leann", ".cuda-layout")
    g.save(idx_file)
    s = pkg.HnswSearcher.load(base, d)
    assert not s.layout_cache_used and not os.path.exists(cache)
    k0 = _search_equal(pkg, s, g, q, d)
    base1 = str(tmp_path / "parsed" / "documents.leann")
    os.makedirs(os.path.dirname(base1))
    s.save(base1)                                   # what a handle loaded by parsing the node block writes back
    s.write_layout_cache(base)
    assert os.path.exists(cache)
    s.close()
    s = pkg.HnswSearcher.load(base, d)               # adjacency from the cache, vectors streamed from the .index
    assert s.layout_cache_used
    assert np.array_equal(_search_equal(pkg, s, g, q, d), k0)
    # save() from a cache-loaded handle writes the same .index bytes as the parse-loaded one (levels, keys, lists survive)
    base2 = str(tmp_path / "copy" / "documents.leann")
    os.makedirs(os.path.dirname(base2))
    s.save(base2)
    assert open(base2.replace(".leann", ".index"), "rb").read() == open(base1.replace(".leann", ".index"), "rb").read()
    s.close()
    # a touched .index (same bytes, new mtime) invalidates the cache: parsed again, same answers
    st = os.stat(idx_file)
    os.utime(idx_file, ns=(st.st_atime_ns, st.st_mtime_ns + 5_000_000_000))
    s = pkg.HnswSearcher.load(base, d)
    assert not s.layout_cache_used
    _search_equal(pkg, s, g, q, d)
    s.write_layout_cache(base)
    s.close()
    # a truncated cache file is ignored
    data = open(cache, "rb").read()
    open(cache, "wb").write(data[:-16])
    s = pkg.HnswSearcher.load(base, d)
    assert not s.layout_cache_used
    _search_equal(pkg, s, g, q, d)
    s.close()
    # a cache bound to another index (different graph of the same shape) is ignored
    open(cache, "wb").write(data)
    g2 = orc.Hnsw.build(x[::-1].copy(), M=16, ef_add=64, seed=3)
    g2.save(idx_file)
    s = pkg.HnswSearcher.load(base, d)
    assert not s.layout_cache_used
    _search_equal(pkg, s, g2, q, d)
    s.close()


def test_layout_cache_env_autowrite(orc, pkg, tmp_path, monkeypatch):
    x, q = make_data(3000, 64, 5, nq=50)
    g = orc.Hnsw.build(x, M=8, ef_add=32, seed=2)
    base = str(tmp_path / "documents.leann")
    g.save(base.replace(".leann", ".index"))
    monkeypatch.setenv("LEANN_CUDA_LAYOUT_CACHE", "1")
    s = pkg.HnswSearcher.load(base, 64)
    assert not s.layout_cache_used and os.path.exists(base.replace(".leann", ".cuda-layout"))
    s.close()
    s = pkg.HnswSearcher.load(base, 64)
    assert s.layout_cache_used
    _search_equal(pkg, s, g, q, 64)
    s.close()
    monkeypatch.setenv("LEANN_CUDA_NO_LAYOUT_CACHE", "1")
    s = pkg.HnswSearcher.load(base, 64)
    assert not s.layout_cache_used
    s.close()


def test_streamed_open_keeps_format_checks(orc, pkg, tmp_path):
    """Checks that used to run on the host copy of the whole file now run in the parser threads / on the device."""
    x, q = make_data(2000, 32, 9, nq=10)
    g = orc.Hnsw.build(x, M=8, ef_add=32, seed=2)
    base = str(tmp_path / "documents.leann")
    idx_file = base.replace(".leann", ".index")
    g.save(idx_file)
    data = bytearray(open(idx_file, "rb").read())
    nodes_off = 8 + 2000 * 32 * 4 + 64 + 40 + 2000 * 2
    bad = bytearray(data)
    bad[nodes_off + 10 + 4: nodes_off + 10 + 8] = struct.pack("<I", 2000)        # first neighbour slot of node 0 out of range
    open(idx_file, "wb").write(bad)
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.HnswSearcher.load(base, 32)
    assert e.value.code == pkg.ERR_BAD_FORMAT and "out of range" in e.value.message
    bad = bytearray(data)
    bad[nodes_off + 10: nodes_off + 14] = struct.pack("<I", 17)                  # neighbour count above connectivity_base (16)
    open(idx_file, "wb").write(bad)
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.HnswSearcher.load(base, 32)
    assert e.value.code == pkg.ERR_BAD_FORMAT and "count" in e.value.message
    # .diskann: neighbour ids are range-checked on the device
    v = orc.Vamana.build(x, R=16, L=32, alpha=1.2, seed=1)
    dk = base.replace(".leann", ".diskann")
    v.save(dk)
    s = pkg.DiskAnnSearcher.load(base, 32)
    ok, od, oc, _ = v.search(q, 5, 32, lanes=pkg.reduction_lanes(32), next_cap=pkg.queue_capacity(32, False))
    keys, dists, _ = s.search_batch(q, 5, 32)
    assert np.array_equal(keys, ok) and np.array_equal(dists.view(np.uint32), od.view(np.uint32))
    s.close()
    raw = bytearray(open(dk, "rb").read())
    raw[-4:] = struct.pack("<I", 2000)                                           # last adjacency entry = n
    open(dk, "wb").write(raw)
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.DiskAnnSearcher.load(base, 32)
    assert e.value.code == pkg.ERR_BAD_FORMAT and "out of range" in e.value.message
    # .embeddings (exact scan): count = file length / (4 * dims), trailing bytes ignored (embeddings.rs:26-27)
    emb = base.replace(".leann", ".embeddings")
    open(emb, "wb").write(x.tobytes() + b"\x01\x02\x03")
    f = pkg.FlatSearcher.load(base, 32)
    assert len(f) == 2000
    fk, fs, _ = f.search_batch(q, 5, 0)
    oi, osc, _ = orc.exact_scan(q, x, 5, metric=0)
    assert np.array_equal(fk, oi)
    f.close()


def test_save_is_atomic_and_reports_failures(pkg, tmp_path):
    """writers go through `<file>.tmp`, check every write and rename only a complete file over the target."""
    x, _ = make_data(800, 32, 4)
    s = pkg.HnswSearcher.build(x, graph_degree=8, complexity=32, seed=1)
    base = str(tmp_path / "documents.leann")
    s.save(base)
    first = open(base.replace(".leann", ".index"), "rb").read()
    assert not os.path.exists(base.replace(".leann", ".index") + ".tmp")
    with pytest.raises(pkg.LeannCudaError) as e:
        s.save(str(tmp_path / "no_such_dir" / "documents.leann"))
    assert e.value.code == pkg.ERR_NOT_FOUND
    s.save(base)                                   # overwriting in place (what add_to_index does) leaves one complete file
    assert open(base.replace(".leann", ".index"), "rb").read() == first
    assert sorted(os.listdir(tmp_path)) == ["documents.index"]
    s.close()
