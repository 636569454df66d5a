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
