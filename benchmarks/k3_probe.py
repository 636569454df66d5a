"""K3 alone: Bm25Scorer::search batched (top-50) over the C5 corpus; prints the device time of bm25_query_kernel, the
algorithmic postings rate and a checksum of the results (identical across A/B builds: LEANN_CUDA_LIB=...)."""
import json, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import leann_rs_b200 as P
from benchmarks import secondary as S2
n, nq = int(os.environ.get("N", 1_000_000)), 10_000
cache = os.environ.get("K3_CORPUS_CACHE")   # e.g. /dev/shm/k3corpus.npz: several A/B processes of one session share the corpus
if cache and os.path.exists(cache):
    z = np.load(cache)
    fb, fo, qb, qo = z["fb"].tobytes(), z["fo"], z["qb"].tobytes(), z["qo"]
    docs = [fb[fo[i]:fo[i + 1]] for i in range(len(fo) - 1)]
    texts = [qb[qo[i]:qo[i + 1]] for i in range(len(qo) - 1)]
else:
    rng = np.random.default_rng(777)
    docs, texts, *_ = S2._corpus(n, nq, 200_000, rng)
    if cache:
        cat = lambda xs: (np.frombuffer(b"".join(xs), dtype=np.uint8), np.concatenate([[0], np.cumsum([len(x) for x in xs])]).astype(np.int64))
        (fb, fo), (qb, qo) = cat(docs), cat(texts)
        np.savez(cache, fb=fb, fo=fo, qb=qb, qo=qo)
t0 = time.time(); bm = P.Bm25Scorer.build(docs); t_build = time.time() - t0
rows = []
for _ in range(4):
    t0 = time.time(); bi, bs, bc = bm.search_batch(texts, 50); host_ms = (time.time() - t0) * 1e3
    npost, kms = bm.last_batch()
    rows.append((round(kms, 3), round(host_ms, 1)))
crc = zlib.crc32(bi.tobytes()) ^ zlib.crc32(bs.tobytes()) ^ zlib.crc32(bc.tobytes())
print(json.dumps({"lib": os.path.basename(P.LIB_PATH), "dense_frac": os.environ.get("LEANN_CUDA_BM25_DENSE_FRAC"), "dense_rows": bm.dense_rows(), "kernel_ms/host_ms": rows, "postings": npost, "stream_bytes": bm.last_batch_bytes(), "stream_GBps": round(bm.last_batch_bytes() / rows[-1][0] / 1e6, 1), "algorithmic_GBps": round(npost * 8 / rows[-1][0] / 1e6, 1),
                  "build_s": round(t_build, 1), "crc": crc}))
