"""GPU-free coverage of the product's index-file readers (csrc/formats.cpp) through leann_cuda_check_index_file: files written
byte by byte from the published layouts (tests/handmade.py) and files written by the oracle must parse to exactly the graph
that was written; every corruption the reader guards against must be reported with its error class (the format
module had no GPU-free test)."""
import os
import struct

import numpy as np
import pytest

import handmade
from conftest import make_data

MASK64 = (1 << 64) - 1
SEP = 0xFFFFFFFFFF


def _fnv(lists_per_node, keys):
    """Mirror of the hash in leann_cuda_check_index_file: per node, each list's slots in order then a separator; then the key."""
    h = 0xCBF29CE484222325
    for lists, key in zip(lists_per_node, keys):
        for nb in lists:
            for s in nb:
                h = ((h ^ s) * 0x100000001B3) & MASK64
            h = ((h ^ SEP) * 0x100000001B3) & MASK64
        h = ((h ^ key) * 0x100000001B3) & MASK64
    return h


def test_handmade_usearch_file_parses_to_the_written_graph(pkg, tmp_path):
    vecs, keys, levels, adj, q = handmade.ring_case()
    base = str(tmp_path / "documents.leann")
    handmade.write_usearch_index(base.replace(".leann", ".index"), vecs, keys, levels, adj, M=2, M0=4, entry=0, max_level=1)
    info = pkg.check_index_file(base, pkg.BACKEND_HNSW, 2)
    assert (info["n"], info["dims"], info["M"], info["M0"], info["max_level"], info["entry"], info["n_upper_lists"]) == (8, 2, 2, 4, 1, 0, 2)
    assert info["adjacency_hash"] == _fnv(adj, keys)          # list order, list lengths and keys exactly as written


def test_handmade_diskann_file_parses_to_the_written_graph(pkg, tmp_path):
    vecs, keys, levels, adj, q = handmade.ring_case()
    chain = [a[0] for a in adj]
    base = str(tmp_path / "documents.leann")
    handmade.write_diskann(base.replace(".leann", ".diskann"), vecs, chain, R=2, medoid=0)
    info = pkg.check_index_file(base, pkg.BACKEND_VAMANA, 2)
    assert (info["n"], info["dims"], info["M"], info["entry"]) == (8, 2, 2, 0)
    assert info["adjacency_hash"] == _fnv([[c] for c in chain], list(range(8)))
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.check_index_file(base, pkg.BACKEND_VAMANA, 3)
    assert e.value.code == pkg.ERR_DIM_MISMATCH


def test_oracle_written_files_and_corruptions(pkg, orc, tmp_path):
    x, _ = make_data(1500, 24, 6)
    g = orc.Hnsw.build(x, M=8, ef_add=32, seed=6)
    base = str(tmp_path / "documents.leann")
    idx = base.replace(".leann", ".index")
    g.save(idx)
    info, oinfo = pkg.check_index_file(base, pkg.BACKEND_HNSW, 24), g.info()
    for k in ("n", "M", "M0", "max_level", "entry"):
        assert info[k] == oinfo[k], k
    data = bytearray(open(idx, "rb").read())
    nodes_off = 8 + 1500 * 24 * 4 + 64 + 40 + 1500 * 2

    def expect(mutated, code, word):
        open(idx, "wb").write(mutated)
        with pytest.raises(pkg.LeannCudaError) as e:
            pkg.check_index_file(base, pkg.BACKEND_HNSW, 24)
        assert e.value.code == code and word in e.value.message, e.value.message

    bad = bytearray(data); bad[nodes_off + 14: nodes_off + 18] = struct.pack("<I", 1500)
    expect(bad, pkg.ERR_BAD_FORMAT, "out of range")                      # neighbour slot == n
    bad = bytearray(data); bad[nodes_off + 10: nodes_off + 14] = struct.pack("<I", 17)
    expect(bad, pkg.ERR_BAD_FORMAT, "count")                             # count above connectivity_base
    bad = bytearray(data); bad[nodes_off + 8: nodes_off + 10] = struct.pack("<h", 7)
    expect(bad, pkg.ERR_BAD_FORMAT, "level")                             # node level disagrees with the level table
    expect(data[:-1], pkg.ERR_BAD_FORMAT, "truncated")
    expect(data + b"\0", pkg.ERR_BAD_FORMAT, "trailing")
    bad = bytearray(data); bad[8 + 1500 * 24 * 4 + 14] = 12
    expect(bad, pkg.ERR_BAD_FORMAT, "kinds")                             # f16 scalars
    expect(b"IxHN" + bytes(200), pkg.ERR_FAISS_FORMAT, "FAISS")
    os.remove(idx)
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.check_index_file(base, pkg.BACKEND_HNSW, 24)
    assert e.value.code == pkg.ERR_NOT_FOUND
    # .embeddings: count = floor(len / (4 * dims))
    open(base.replace(".leann", ".embeddings"), "wb").write(x.tobytes() + b"abc")
    assert pkg.check_index_file(base, pkg.BACKEND_FLAT, 24)["n"] == 1500
