"""A usearch `.index` written byte by byte from the published layout, independent of both the
oracle's and the product's writers, over a graph small enough to trace the published search loop by hand."""
import struct

import numpy as np


def write_usearch_index(path, vecs, keys, levels, adj, M, M0, entry, max_level):
    """adj[i][l] = list of neighbour slots of node i on level l (list order is traversal order)."""
    n, d = vecs.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<II", n, d * 4))
        f.write(np.ascontiguousarray(vecs, dtype="<f4").tobytes())
        head = bytearray(64)
        head[0:7] = b"usearch"
        struct.pack_into("<HHH", head, 7, 2, 23, 0)
        head[13] = ord("i"); head[14] = 11; head[15] = 14; head[16] = 15      # metric ip, f32, u64 keys, u32 slots
        struct.pack_into("<QQQ", head, 17, n, 0, d)
        head[41] = 0
        f.write(bytes(head))
        f.write(struct.pack("<QQQQQ", n, M, M0, max_level, entry))
        f.write(np.asarray(levels, dtype="<i2").tobytes())
        for i in range(n):
            f.write(struct.pack("<Qh", keys[i], levels[i]))
            for l in range(levels[i] + 1):
                cap = M0 if l == 0 else M
                nb = list(adj[i][l])
                f.write(struct.pack("<I", len(nb)))
                f.write(np.asarray(nb + [0xDEADBEEF] * (cap - len(nb)), dtype="<u4").tobytes())   # unused slots are undefined in usearch


def ring_case():
    """Eight unit vectors on a circle at 0, 10, .., 70 degrees, linked as a chain 0-1-2-...-7 (each list: lower neighbour first).
    Keys are 100 + slot. Node 3 also lives on level 1 together with node 0; the entry point is node 0 on level 1.
    Query at 70 degrees: distance(i) = 1 - cos(70 - 10 i degrees), strictly decreasing along the chain."""
    ang = np.deg2rad(np.arange(8) * 10.0)
    vecs = np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32)
    levels = [1, 0, 0, 1, 0, 0, 0, 0]
    adj = []
    for i in range(8):
        l0 = [j for j in (i - 1, i + 1) if 0 <= j < 8]
        lists = [l0]
        if levels[i] == 1:
            lists.append([3] if i == 0 else [0])
        adj.append(lists)
    q = np.array([[np.cos(np.deg2rad(70.0)), np.sin(np.deg2rad(70.0))]], dtype=np.float32)
    return vecs, [100 + i for i in range(8)], levels, adj, q


# Hand trace of usearch search (Appendix A.2) for ring_case, query at 70 degrees, wanted k, expansion ef:
#   level 1: closest = 0; pass 1 measures its level-1 neighbour 3 (1 evaluation): closer -> closest = 3, changed;
#            pass 2 measures 3's level-1 neighbour 0: not closer -> stop.           n_dist = 1 (entry) + 2, upper hops = 2
#   level 0 from node 3, ef = 2: top = [3], next = [3]
#     pop 3: neighbours 2, 4 both new; 2: top not full -> next/top insert; 4: top full (3,2)? after inserting 2 top = [3,2] is full,
#            d(4) < radius = d(2) -> insert, top = [4,3], radius = d(3)
#     pop 4 (nearest in next): neighbours 3 (seen), 5: d(5) < d(3) -> top = [5,4], radius = d(4)
#     pop 5: neighbours 4 (seen), 6 -> top = [6,5];  pop 6: 5 (seen), 7 -> top = [7,6];  pop 7: 6 seen, nothing new
#     next still holds 3? no: 3 was popped; it holds 2 with d(2) > radius = d(6) -> stop.
#   level-0 hops: 3, 4, 5, 6, 7 = 5; level-0 evaluations: 2, 4, 5, 6, 7 = 5  -> n_dist = 3 + 5 = 8
EXPECT_EF2 = {"keys": [107, 106], "n_dist": 8, "hops0": 5, "hops_upper": 2}
#   ef = 1 (k = 1): top = [3]; pop 3: 2: top full, d(2) > d(3) -> rejected; 4: d(4) < d(3) -> top = [4]; pop 4 -> 5; pop 5 -> 6; pop 6 -> 7;
#   pop 7: nothing; next empty -> stop. hops 3,4,5,6,7 = 5; evaluations 2,4,5,6,7 = 5 -> n_dist = 8
EXPECT_EF1 = {"keys": [107], "n_dist": 8, "hops0": 5, "hops_upper": 2}


def write_diskann(path, vecs, adj, R, medoid, name=b"DistDot"):
    """diskann-rs single-file layout: u64 metadata length, bincode metadata
    {dim, n, max_degree u64; medoid u32; vectors_offset, adjacency_offset u64; distance name (u64 length + bytes)},
    vectors, fixed-degree adjacency padded with u32::MAX."""
    n, d = vecs.shape
    meta_len = 52 + len(name)
    voff = 8 + meta_len
    aoff = voff + n * d * 4
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", meta_len))
        f.write(struct.pack("<QQQIQQQ", d, n, R, medoid, voff, aoff, len(name)) + name)
        f.write(np.ascontiguousarray(vecs, dtype="<f4").tobytes())
        for i in range(n):
            nb = list(adj[i])
            f.write(np.asarray(nb + [0xFFFFFFFF] * (R - len(nb)), dtype="<u4").tobytes())


# Hand trace of diskann-rs search_with_dists (Appendix A.3) on the chain 0-1-...-7 of ring_case (level-0 lists only, R = 2,
# medoid 0), query at 70 degrees, beam L = max(complexity, k):
#   L = 2: results = {0}; pop 0 (results not full): 1 is new, results = {0,1}; pop 1: full, d(1) < worst = d(0): go on;
#          2 new, d(2) < d(0): results {0,1,2} -> drop worst -> {1,2}; ... each pop i reveals i+1 which replaces i-1 ...
#          pop 7: d(7) < worst = d(6): go on; its only neighbour 6 is visited; queue empty -> stop.
#          pops 0..7 = 8 hops, evaluations medoid + 1..7 = 8, answer [7, 6].
VAMANA_L2 = {"keys": [7, 6], "n_dist": 8, "hops0": 8}
#   L = 1: results = {0} is full at once; first pop: d(0) >= worst = d(0) -> the non-strict stop rule ends the search at the medoid.
VAMANA_L1 = {"keys": [0], "n_dist": 1, "hops0": 0}
