"""K3 alone: Bm25Scorer::search batched (top-50) over the C5 corpus; prints the device time of bm25_query_kernel, the
algorithmic postings rate and a checksum of the results (identical across A/B builds: LEANN_CUDA_LIB=...)."""
import json, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import leann_rs_b200 as P
from benchmarks import secondary as S2
n, nq = int(os.environ.get("N", 1_000_000)), 10_000
rng = np.random.default_rng(777)
docs, texts, *_ = S2._corpus(n, nq, 200_000, rng)
t0 = time.time(); bm = P.Bm25Scorer.build(docs); t_build = time.time() - t0
rows = []
for _ in range(4):
    t0 = time.time(); bi, bs, bc = bm.search_batch(texts, 50); host_ms = (time.time() - t0) * 1e3
    npost, kms = bm.last_batch()
    rows.append((round(kms, 3), round(host_ms, 1)))
crc = zlib.crc32(bi.tobytes()) ^ zlib.crc32(bs.tobytes()) ^ zlib.crc32(bc.tobytes())
print(json.dumps({"lib": os.path.basename(P.LIB_PATH), "kernel_ms/host_ms": rows, "postings": npost, "algorithmic_GBps": round(npost * 8 / rows[-1][0] / 1e6, 1),
                  "build_s": round(t_build, 1), "crc": crc}))
