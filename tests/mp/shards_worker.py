"""Worker of tests/test_shards_gpu.py::test_process_per_gpu_join (launched by torch.distributed.run, one process per GPU).
Every rank builds its own exact-scan and HNSW shard, joins the library's NCCL communicator (`leann_cuda_shards_join`) and
checks the merged answer against a numpy merge of the per-shard answers gathered over gloo."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import leann_rs_b200 as P
    from conftest import make_data
    from leann_rs_b200.shards import numpy_topk_merge, shard_bounds

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    n, d, k, nq = 6000, 128, 10, 64
    x, q = make_data(n, d, 11, nq=nq)
    lo, hi = shard_bounds(n, world, rank)
    uid = [P.ShardedBackend.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    for kind in ("flat", "hnsw"):
        if kind == "flat":
            part = P.FlatSearcher.from_vectors(x[lo:hi], metric=P.METRIC_DOT_DESC, device=local)
            ef = 0
        else:
            part = P.HnswSearcher.build(x[lo:hi], graph_degree=16, complexity=64, seed=5, device=local)
            ef = 48
        if kind == "hnsw":   # a communicator per handle: fresh id
            uid = [P.ShardedBackend.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
        sh = P.ShardedBackend.join(part, uid[0], rank, world, lo)
        assert len(sh) == n and sh.info()["shards"] == world
        lk, ld, lc = part.search_batch(q, k, ef)
        # host-buffer call and device-buffer call
        mk, md, mc = sh.search_batch(q, k, ef)
        qt = torch.from_numpy(q).cuda()
        dk, dd, dc = sh.search_device(qt, k, ef)
        torch.cuda.synchronize()
        assert np.array_equal(dk.cpu().numpy().astype(np.uint64), mk) and np.array_equal(dd.cpu().numpy(), md)
        # reference merge from the per-shard answers (gloo all_gather of numpy arrays)
        allk, alld = [None] * world, [None] * world
        dist.all_gather_object(allk, np.where(lk == np.uint64(0xFFFFFFFFFFFFFFFF), lk, lk + np.uint64(lo)))
        dist.all_gather_object(alld, ld)
        rk, rd = numpy_topk_merge(np.stack(allk), np.stack(alld), descending=(kind == "flat"))
        assert np.array_equal(rk, mk), f"{kind}: merged ids differ on rank {rank}"
        assert np.array_equal(rd.view(np.uint32), md.view(np.uint32))
        assert sh.info()["exchange"] == ("nccl_all_gather" if world > 1 else "none")
        sh.close()
        part.close()
    # ---- hybrid path over document-range shards: bit-identical to the unsharded hybrid search ----
    nd, dd, nqh = 4000, 64, 48
    xv, qv = make_data(nd, dd, 21, nq=nqh)
    rng = np.random.default_rng(5)
    vocab = [f"tok{i}" for i in range(300)]
    docs = [" ".join(rng.choice(vocab, size=int(rng.integers(5, 40))).tolist()) for _ in range(nd)]
    texts = [" ".join(rng.choice(vocab[:60], size=int(rng.integers(1, 5))).tolist()) for _ in range(nqh)]
    mask = P.pack_mask(rng.random(nd) < 0.4)
    lo, hi = shard_bounds(nd, world, rank)
    whole = P.FlatSearcher.from_vectors(xv, metric=P.METRIC_IP, device=local)       # exact backend: the same candidates sharded or not
    bm_whole = P.Bm25Scorer.build(docs, device=local)
    part = P.FlatSearcher.from_vectors(xv[lo:hi], metric=P.METRIC_IP, device=local)
    blobs = [None] * world
    dist.all_gather_object(blobs, P.Bm25Scorer.shard_stats(docs[lo:hi], device=local))
    bm_part = P.Bm25Scorer.build_sharded(docs[lo:hi], P.Bm25Scorer.merge_stats(blobs), device=local)
    uid = [P.ShardedBackend.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sh = P.ShardedBackend.join(part, uid[0], rank, world, lo)
    for hybrid, alpha, m in ((True, 0.5, mask), (True, 0.7, None), (False, 0.5, mask), (True, 0.0, None)):
        ri, rs, rc = P.text.hybrid_search(whole, bm_whole, qv, texts, 10, 0, hybrid, alpha, m)
        si, ss, sc = sh.hybrid_search(bm_part, qv, texts, 10, 0, hybrid, alpha, m)
        assert np.array_equal(rc, sc), (hybrid, alpha, rank)
        assert np.array_equal(ri, si), (hybrid, alpha, rank, float(np.mean(ri == si)))
        assert np.array_equal(rs.view(np.uint32), ss.view(np.uint32)), (hybrid, alpha, rank)
    sh.close(); part.close(); whole.close(); bm_part.close(); bm_whole.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDS_WORKER_OK")


if __name__ == "__main__":
    main()
