#!/usr/bin/env python
"""oracle/pin_graph_golden.py — TEST INFRASTRUCTURE. Produces the golden vectors that PIN the graph half of the oracle.

The reference delegates HNSW search to usearch 2.23.0 (Cargo.lock:4381-4388; call sites src/backend/hnsw.rs:43-55,85,
122-134) and this container has neither the crate's sources nor a `usearch` wheel, so `oracle/graph_oracle.cpp` is a
restatement from the published algorithm ("parity unpinned"). Run THIS script on any machine with

    pip install usearch==2.23.0 numpy

to build a small index with the reference's exact options (hnsw.rs:43-51: MetricKind::IP, f32, connectivity 32,
expansion_add 64, expansion_search 64; keys = ordinals added one by one, hnsw.rs:128-130), save it in usearch's own
`.index` format and record what `Index.search(query, k)` returns:

    tests/golden/graph_golden.index   the file usearch wrote (pins the on-disk format: SURVEY.md Appendix A.1)
    tests/golden/graph_golden.json    queries (f32 bits), keys and distance bits per query, library version

Commit both. `tests/test_graph_golden.py` (skipped while they are absent) then loads the file through the oracle's reader
and the product's reader, runs the same queries through both and compares keys and distance bits with the recording; on
disagreement it searches the oracle's named behaviour switches (graph_oracle.cpp CompatBits) for the combination that
reproduces the recording and prints it, so the fix is flipping one constant in oracle/graph_oracle.cpp and
leann_rs_b200/csrc/compat.h. A `.diskann` golden file for the Vamana path needs the Rust crate diskann-rs 0.3.4 (no Python
binding): `--diskann-from FILE.json` imports a recording made by the 20-line Rust program printed by `--print-rust`.
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
N, D, NQ, K = 2000, 64, 200, 10

RUST = r'''
// cargo add diskann-rs@0.3.4 anndists@0.1.3 serde_json ; cargo run --release -- out.json
use anndists::dist::DistDot;
use diskann_rs::{DiskANN, DiskAnnParams};
fn main() {
    let (n, d, nq, k) = (2000usize, 64usize, 200usize, 10usize);
    let mut s = 12345u64;                                   // xorshift64*, same stream as pin_graph_golden.py --seed-stream
    let mut next = || { s ^= s >> 12; s ^= s << 25; s ^= s >> 27; ((s.wrapping_mul(0x2545F4914F6CDD1D) >> 40) as f32) / 16777216.0 - 0.5 };
    let mut rows: Vec<Vec<f32>> = (0..n + nq).map(|_| { let v: Vec<f32> = (0..d).map(|_| next()).collect();
        let nrm = v.iter().map(|x| x * x).sum::<f32>().sqrt(); v.iter().map(|x| x / nrm).collect() }).collect();
    let queries = rows.split_off(n);
    let params = DiskAnnParams { max_degree: 32, build_beam_width: 64, alpha: 1.2 };       // diskann.rs:88-92
    let idx = DiskANN::<DistDot>::build_index_with_params(&rows, DistDot {}, "graph_golden.diskann", params).unwrap();
    let out: Vec<_> = queries.iter().map(|q| idx.search_with_dists(q, k, 64)).collect();  // diskann.rs:54-56
    let j: Vec<_> = out.iter().map(|r| (r.iter().map(|x| x.0).collect::<Vec<u32>>(), r.iter().map(|x| x.1.to_bits()).collect::<Vec<u32>>())).collect();
    std::fs::write(std::env::args().nth(1).unwrap(), serde_json::to_string(&j).unwrap()).unwrap();
}
'''


def data(seed=12345):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((8, D)).astype(np.float32)
    f = lambda m: (rng.standard_normal((m, 8)).astype(np.float32) @ W + 0.3 * rng.standard_normal((m, D)).astype(np.float32))
    x, q = f(N), f(NQ)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    x[100] = x[7]      # exact duplicates: equal distances exercise the tie rules the oracle had to recall
    x[101] = x[7]
    return np.ascontiguousarray(x, np.float32), np.ascontiguousarray(q, np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--print-rust", action="store_true")
    ap.add_argument("--out", default=GOLDEN)
    a = ap.parse_args()
    if a.print_rust:
        print(RUST)
        return 0
    try:
        import usearch
        from usearch.index import Index, MetricKind, ScalarKind
    except ImportError:
        print("usearch is not installed here: run this script where `pip install usearch==2.23.0` is possible", file=sys.stderr)
        return 2
    x, q = data()
    index = Index(ndim=D, metric=MetricKind.IP, dtype=ScalarKind.F32, connectivity=32, expansion_add=64, expansion_search=64,
                  multi=False)                                       # hnsw.rs:43-51
    for i in range(N):                                               # hnsw.rs:128-130: one add per passage, key = ordinal
        index.add(i, x[i], threads=1)
    os.makedirs(a.out, exist_ok=True)
    path = os.path.join(a.out, "graph_golden.index")
    index.save(path)                                                 # hnsw.rs:134
    loaded = Index(ndim=D, metric=MetricKind.IP, dtype=ScalarKind.F32, connectivity=32, expansion_add=64, expansion_search=64)
    loaded.load(path)                                                # hnsw.rs:53-55
    keys, dists = [], []
    for i in range(NQ):                                              # hnsw.rs:85: search(query, top_k), one query per call
        m = loaded.search(q[i], K, threads=1)
        keys.append([int(v) for v in m.keys])
        dists.append([int(v) for v in np.asarray(m.distances, dtype=np.float32).view(np.uint32)])
    rec = {"library": "usearch", "version": getattr(usearch, "__version__", "?"), "n": N, "dim": D, "k": K, "connectivity": 32,
           "expansion_add": 64, "expansion_search": 64, "metric": "ip", "queries_f32_bits": q.view(np.uint32).tolist(),
           "keys": keys, "distance_f32_bits": dists, "made_by": "oracle/pin_graph_golden.py"}
    json.dump(rec, open(os.path.join(a.out, "graph_golden.json"), "w"))
    print(f"wrote {path} and graph_golden.json (usearch {rec['version']})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
