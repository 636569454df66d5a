"""GPU parity of K3/K3f (BM25 scoring, BM25 top-k, hybrid fusion, post-filter walk) and of the
IndexSearcher glue against the text oracle: scores must be bit-identical f32, ids identical."""
import json
import os

import numpy as np
import pytest

from conftest import make_data
from oracle import text_oracle as T

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "text_golden.json")))


def test_bm25_golden_bits(pkg):
    for case in GOLD["bm25"]:
        sc = pkg.Bm25Scorer.build(case["docs"])
        st = sc.stats()
        assert st["num_docs"] == len(case["docs"])
        assert np.float32(st["avg_doc_len"]).view(np.uint32) == case["avg_doc_len_bits"]
        for q, want, top in zip(case["queries"], case["scores_bits"], case["top5"]):
            assert sc.score_query(q).view(np.uint32).tolist() == want, q
            got = sc.search(q, 5)
            assert [[i, int(np.float32(s).view(np.uint32))] for i, s in got] == top, q


def test_bm25_reference_unit_tests_on_gpu(pkg):
    """bm25.rs:199-280 assertions, through the GPU path."""
    fox = ["the quick brown fox jumps over the lazy dog", "a quick brown dog outpaces a swift fox", "the dog chases the fox around the yard"]
    r = pkg.Bm25Scorer.build(fox).search("quick fox", 3)
    assert r and len(r) <= 3
    s = pkg.Bm25Scorer.build(["rust rust rust programming", "rust programming"]).score_query("rust")
    assert s[0] > s[1]
    s = pkg.Bm25Scorer.build(["common rare", "common", "common"]).score_query("rare")
    assert s[0] > 0 and s[1] == 0 and s[2] == 0
    assert pkg.Bm25Scorer.build(["hello world"]).score_query("")[0] == 0
    assert pkg.Bm25Scorer.build(["hello world"]).search("xyz", 5) == []
    r = pkg.Bm25Scorer.build(["apple banana", "apple cherry", "banana cherry", "apple apple apple"]).search("apple", 2)
    assert len(r) == 2 and r[0][0] == 3 and r[1][0] == 0


def test_hybrid_rerank_golden_bits(pkg):
    for h in GOLD["hybrid"]:
        r = pkg.hybrid_rerank([(i, s) for i, s in h["vec"]], np.asarray(h["bm25"], dtype=np.float32), h["alpha"])
        assert [[i, int(np.float32(s).view(np.uint32))] for i, s in r] == h["result"]


def _corpus(n, seed, vocab=3000):
    rng = np.random.default_rng(seed)
    words = np.array([f"t{i}" for i in range(vocab)])
    p = 1.0 / np.arange(1, vocab + 1) ** 1.07
    p /= p.sum()
    docs = [" ".join(words[rng.choice(vocab, size=int(rng.integers(8, 60)), p=p)]) for _ in range(n)]
    queries = [" ".join(words[rng.choice(vocab, size=int(rng.integers(1, 6)), p=p)]) for _ in range(64)]
    return docs, queries


def test_bm25_medium_corpus_bits_and_topk(pkg):
    docs, queries = _corpus(20000, 3)
    queries += ["", "zzz unknown", "t0 t0 t0", "t1"]
    ref = T.Bm25Scorer(docs)
    sc = pkg.Bm25Scorer.build(docs)
    for q in queries[:12] + queries[-4:]:
        assert np.array_equal(sc.score_query(q).view(np.uint32), ref.score_query_fast(q).view(np.uint32)), q
    idx, scores, cnt = sc.search_batch(queries, 50)
    for i, q in enumerate(queries):
        want = ref.search(q, 50, fast=True)
        assert int(cnt[i]) == len(want)
        assert idx[i, :len(want)].tolist() == [d for d, _ in want], q
        assert scores[i, :len(want)].view(np.uint32).tolist() == [int(np.float32(s).view(np.uint32)) for _, s in want]
    # a very common term exercises the in-kernel prune (more positives than the shared buffer holds)
    big = sc.search("t0", 1000)
    want = ref.search("t0", 1000, fast=True)
    assert [d for d, _ in big] == [d for d, _ in want]
    # long queries (duplicates included): the tile-boundary table is rebuilt every few tiles; every document scores
    rng = np.random.default_rng(11)
    longq = [" ".join(f"t{int(i)}" for i in rng.integers(0, 3000, size=m)) for m in (300, 700, 1024)]
    idx, scores, cnt = sc.search_batch(longq, 100)
    for i, q in enumerate(longq):
        want = ref.search(q, 100, fast=True)
        assert int(cnt[i]) == len(want) and idx[i, :len(want)].tolist() == [d for d, _ in want], len(q)
        assert scores[i, :len(want)].view(np.uint32).tolist() == [int(np.float32(s).view(np.uint32)) for _, s in want]
    with pytest.raises(pkg.LeannCudaError) as e:
        sc.search(" ".join(["t1"] * 1025), 5)
    assert "indexed tokens" in e.value.message


def test_bm25_dense_rows_do_not_change_a_bit(pkg, monkeypatch):
    """K3d: frequent terms are added to the accumulator tile from dense rows instead of posting slices. Whatever the
    threshold (off, default, nearly every term, a row limit of 2), ids, score bits and counts are the same."""
    docs, queries = _corpus(20000, 5)
    queries += ["t0", "t0 t1", "t1 t0 t1", "t900 t0 t2 t901 t1", "t0 t1 t2 t3 t4 t5 t6 t7", "t2500", ""]
    rng = np.random.default_rng(12)
    queries += [" ".join(f"t{int(i)}" for i in rng.integers(0, 60, size=m)) for m in (40, 400)]
    results, rows = [], []
    for frac, mx in (("0", None), (None, None), ("0.002", "1000"), ("0.3", "2")):
        if frac is None: monkeypatch.delenv("LEANN_CUDA_BM25_DENSE_FRAC", raising=False)
        else: monkeypatch.setenv("LEANN_CUDA_BM25_DENSE_FRAC", frac)
        if mx is None: monkeypatch.delenv("LEANN_CUDA_BM25_DENSE_MAX", raising=False)
        else: monkeypatch.setenv("LEANN_CUDA_BM25_DENSE_MAX", mx)
        sc = pkg.Bm25Scorer.build(docs)
        rows.append(sc.dense_rows())
        idx, scores, cnt = sc.search_batch(queries, 50)
        big = sc.search_batch(queries[-12:], 1000)   # many candidates: the scan's full-buffer / resume path
        results.append((idx.copy(), scores.view(np.uint32).copy(), cnt.copy(), big[0].copy(), big[1].view(np.uint32).copy(), big[2].copy()))
    assert rows[0] == 0 and rows[1] >= 3 and rows[2] > rows[1] and rows[3] == 2, rows
    for r in results[1:]:
        for a, b in zip(results[0], r):
            assert np.array_equal(a, b)
    ref = T.Bm25Scorer(docs)
    for i in (len(queries) - 9, len(queries) - 6, len(queries) - 2):
        want = ref.search(queries[i], 50, fast=True)
        assert results[1][0][i, :len(want)].tolist() == [d for d, _ in want]
        assert results[1][1][i, :len(want)].tolist() == [int(np.float32(s).view(np.uint32)) for _, s in want]


def test_bm25_min_max_of_the_dense_vector(pkg, monkeypatch):
    """bm25.rs:152-153 takes max / min over the whole dense score vector. The kernel never materialises it: the minimum is
    0.0 as soon as one document scores 0.0 and the smallest positive score only when every document matched — with and
    without dense rows, over several tiles with a partial last one."""
    rng = np.random.default_rng(21)
    n = 20000
    docs = [" ".join(["every"] * int(rng.integers(1, 4)) + [f"x{int(v)}" for v in rng.integers(0, 50, size=int(rng.integers(3, 30)))]) for _ in range(n)]
    docs[n - 1] = "every tail"
    queries = ["every", "every x1", "x1 every x2 every", "x1", "x1 x2 x3 x4 x5 x6 x7 x8 x9 x10 x11 x12", "tail", "nothing here", "tail every"]
    ref = T.Bm25Scorer(docs)
    dense = [ref.score_query_fast(q) for q in queries]
    for frac in ("0", "0.5", "0.05"):
        monkeypatch.setenv("LEANN_CUDA_BM25_DENSE_FRAC", frac)
        sc = pkg.Bm25Scorer.build(docs)
        assert (sc.dense_rows() > 0) == (frac != "0")
        ti, ts, tc, _, bx, bn = sc.search_shard(queries, 20, 0)
        for i, q in enumerate(queries):
            assert bx[i].view(np.uint32) == dense[i].max().view(np.uint32), (frac, q)
            assert bn[i].view(np.uint32) == dense[i].min().view(np.uint32), (frac, q)
            want = ref.search(q, 20, fast=True)
            assert int(tc[i]) == len(want) and ti[i, :len(want)].tolist() == [d for d, _ in want], (frac, q)
    assert dense[0].min() > 0 and dense[3].min() == 0


def _fixture_dir(tmp_path, orc, n=3000, d=64, missing=(), with_ids=True, seed=13):
    x, q = make_data(n, d, seed, nq=40)
    g = orc.Hnsw.build(x, M=8, ef_add=32, seed=seed)
    base = str(tmp_path / "documents.leann")
    stem = base[: -len(".leann")]
    g.save(stem + ".index")
    docs, queries = _corpus(n, seed, vocab=800)
    rng = np.random.default_rng(seed)
    metas, offsets, pos = [], {}, 0
    exts = ["rs", "py", "md", "txt"]
    with open(stem + ".passages.jsonl", "wb") as f:
        for i in range(n):
            md = {"source": f"dir{i % 100}/f{i}.{exts[i % 4]}", "chunk_index": i % 7, "chunk_type": ["simple", "ast", "context"][i % 3],
                  "lines": int(rng.integers(1, 500))}
            metas.append(md)
            if i in missing:
                continue
            line = json.dumps({"id": f"p{i}", "text": docs[i], "metadata": md}).encode() + b"\n"
            offsets[f"p{i}"] = pos
            f.write(line)
            pos += len(line)
    json.dump(offsets, open(stem + ".passages.idx.json", "w"))
    if with_ids:
        open(stem + ".ids.txt", "w").write("\n".join(f"p{i}" for i in range(n)) + "\n")
    return base, x, q, g, docs, queries, metas


def test_hybrid_search_matches_reference_glue(orc, pkg, tmp_path):
    base, x, q, g, docs, queries, metas = _fixture_dir(tmp_path, orc)
    n, d, k = len(docs), x.shape[1], 10
    s = pkg.HnswSearcher.load(base, d)
    bm = pkg.Bm25Scorer.build(docs)
    ref_bm = T.Bm25Scorer(docs)
    texts = queries[: len(q)]
    filt = T.parse_filter("chunk_type=ast,lines>100")
    passes_bits = np.array([T.filter_matches(filt, m) for m in metas])
    fmask = pkg.MetadataFilter.parse("chunk_type=ast,lines>100").mask(metas)
    assert np.array_equal(fmask, pkg.pack_mask(passes_bits))
    for hybrid, use_filter, alpha in [(True, False, 0.5), (True, True, 0.7), (False, True, 0.7), (False, False, 0.7), (True, False, 1.0), (True, False, 0.0)]:
        fk = T.fetch_k(k, use_filter, hybrid)
        bk, bd, bc = s.search_batch(q, fk, 64)            # what backend.search returns (parity-checked elsewhere)
        idx, sc, cnt = pkg.text.hybrid_search(s, bm, q, texts if hybrid else None, k, 64, hybrid, alpha, fmask if use_filter else None)
        for i in range(len(q)):
            want = T.search_with_options(bk[i, :bc[i]].tolist(), bd[i, :bc[i]].tolist(), k, ref_bm, texts[i], hybrid, alpha,
                                         (lambda j: bool(passes_bits[j])) if use_filter else (lambda j: True), fk, fast=True)
            assert int(cnt[i]) == len(want), (hybrid, use_filter, i)
            assert idx[i, :len(want)].tolist() == [a for a, _ in want], (hybrid, use_filter, alpha, i)
            assert sc[i, :len(want)].view(np.uint32).tolist() == [int(np.float32(b).view(np.uint32)) for _, b in want]


def test_index_searcher_end_to_end(orc, pkg, tmp_path):
    missing = {5, 17, 1234}
    base, x, q, g, docs, queries, metas = _fixture_dir(tmp_path, orc, missing=missing)
    d, k = x.shape[1], 5
    s = pkg.IndexSearcher.load(base, "hnsw", d)
    assert len(s) == len(docs)
    ref_docs = ["" if i in missing else docs[i] for i in range(len(docs))]   # get_all_texts: missing -> empty (searcher.rs:219)
    ref_bm = T.Bm25Scorer(ref_docs)
    lanes = pkg.reduction_lanes(d)
    for filt_expr, hybrid in [(None, False), ("source:*.rs", False), ("lines>=400", True), (None, True)]:
        filt = T.parse_filter(filt_expr) if filt_expr else None
        opts = pkg.SearchOptions.new(k, 999)
        if filt_expr:
            opts = opts.with_filter(filt_expr)
        fk = T.fetch_k(k, filt is not None, hybrid)
        ok, od, oc, _ = g.search(q, fk, 64, lanes=lanes, next_cap=max(64, fk))   # HNSW ignores complexity: ef = 64
        idx, sc, cnt = s.search_batch(q, opts if not hybrid else opts.with_hybrid("unused", 0.6), queries[: len(q)] if hybrid else None)
        passes = lambda j: (j not in missing) and j < len(metas) and (filt is None or T.filter_matches(filt, metas[j]))
        for i in range(len(q)):
            want = T.search_with_options(ok[i, :oc[i]].astype(np.int64).tolist(), od[i, :oc[i]].tolist(), k, ref_bm, queries[i], hybrid, 0.6,
                                         passes, fk, fast=True)
            assert idx[i, :cnt[i]].tolist() == [a for a, _ in want], (filt_expr, hybrid, i)
            assert sc[i, :cnt[i]].view(np.uint32).tolist() == [int(np.float32(b).view(np.uint32)) for _, b in want]
    # single-query reference-shaped calls
    res = s.search(q[0], 3, 64)
    assert len(res) == 3 and res[0].id.startswith("p") and res[0].text and "source" in res[0].metadata
    hits = s.bm25_search(queries[0], 5)
    want = ref_bm.search(queries[0], 5, fast=True)
    assert hits == [ref_docs[i] for i, _ in want]
    with pytest.raises(pkg.LeannCudaError) as e:
        s.search_batch(q, pkg.SearchOptions.new(k, 64).with_filter("nonsense"))
    assert e.value.code == pkg.ERR_PARSE
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.IndexSearcher.load(base, "faiss", d)
    assert "Unknown backend" in e.value.message


def test_sharded_bm25_and_fuse_equal_unsharded(orc, pkg, tmp_path):
    """two document-range shards built with the merged corpus statistics, reduced the way the
    collectives of ShardedHybridSearcher reduce them, reproduce the unsharded BM25 top-k and the unsharded hybrid
    result bit for bit."""
    import torch
    base, x, q, g, docs, queries, metas = _fixture_dir(tmp_path, orc)
    n, d, k = len(docs), x.shape[1], 10
    fk = 5 * k
    s = pkg.HnswSearcher.load(base, d)
    full = pkg.Bm25Scorer.build(docs)
    cut = 1234
    st = pkg.Bm25Scorer.merge_stats([pkg.Bm25Scorer.shard_stats(docs[:cut]), pkg.Bm25Scorer.shard_stats(docs[cut:])])
    assert st == pkg.Bm25Scorer.shard_stats(docs)
    sh = [pkg.Bm25Scorer.build_sharded(docs[:cut], st), pkg.Bm25Scorer.build_sharded(docs[cut:], st)]
    assert len(sh[0]) == cut and len(sh[1]) == n - cut
    assert np.float32(sh[0].stats()["avg_doc_len"]).view(np.uint32) == np.float32(full.stats()["avg_doc_len"]).view(np.uint32)
    texts = queries[: len(q)]
    vk, vd, vc = s.search_batch(q, fk, 64)
    parts = [sh[0].search_shard(texts, fk, 0, vk, vc), sh[1].search_shard(texts, fk, cut, vk, vc)]
    gk = torch.from_numpy(np.stack([p[0].view(np.int64) for p in parts])).cuda()
    gs = torch.from_numpy(np.stack([p[1] for p in parts])).cuda()
    mk, ms, mc = pkg.topk_merge_device(gk, gs, descending=True)
    mk, ms, mc = mk.cpu().numpy(), ms.cpu().numpy(), mc.cpu().numpy()
    fi, fs, fc = full.search_batch(texts, fk)
    for i in range(len(texts)):
        c = int(fc[i])
        assert int(mc[i]) == c and mk[i, :c].tolist() == fi[i, :c].astype(np.int64).tolist(), texts[i]
        assert ms[i, :c].view(np.uint32).tolist() == fs[i, :c].view(np.uint32).tolist()
        dense = full.score_query(texts[i])
        cb = parts[0][3][i] + parts[1][3][i]
        assert np.array_equal(cb[: vc[i]].view(np.uint32), dense[vk[i, : vc[i]].astype(np.int64)].view(np.uint32))
        assert max(parts[0][4][i], parts[1][4][i]) == dense.max() and min(parts[0][5][i], parts[1][5][i]) == dense.min()
    cb = parts[0][3] + parts[1][3]
    bx, bn = np.maximum(parts[0][4], parts[1][4]), np.minimum(parts[0][5], parts[1][5])
    bits = np.array([m["chunk_type"] == "ast" for m in metas])
    for hybrid, mask in ((True, None), (True, pkg.pack_mask(bits)), (False, pkg.pack_mask(bits))):
        idx, sc, cnt = pkg.hybrid_fuse(vk, vd, vc, k, hybrid, 0.5, cb, mk.view(np.uint64), ms, mc.astype(np.uint32), bx, bn, mask, n)
        ridx, rsc, rcnt = pkg.text.hybrid_search(s, full, q, texts if hybrid else None, k, 64, hybrid, 0.5, mask)
        assert np.array_equal(cnt, rcnt) and np.array_equal(idx, ridx) and np.array_equal(sc.view(np.uint32), rsc.view(np.uint32))


def test_gpu_index_build_edge_cases(pkg):
    """The device tokenizer / index builder against the oracle on awkward inputs: empty and one-letter passages, non-ASCII
    bytes, mixed case, tokens that straddle the 4096-byte scan chunks, one very long token, an empty corpus."""
    rng = np.random.default_rng(3)
    words = ["rust", "Rust", "RUST", "x", "ab", "a1", "42", "é東京", "naïve", "foo_bar", "C++", "zz" * 300]
    docs = ["", "a", "a b c", "é 東 京", "Rust RUST rust x ab", "ab" + "_" * 4093 + "cd ef", "q" * 4095 + " tail end", "w" * 10000,
            "x" * 4094 + " ab cd", "ends with token zz", " " * 5000 + "late token"]
    docs += [" ".join(rng.choice(words, size=int(rng.integers(0, 40)))) for _ in range(300)]
    docs += ["pad " * 1000 + "needle"]          # > one chunk of ordinary text
    ref = T.Bm25Scorer(docs)
    sc = pkg.Bm25Scorer.build(docs)
    st = sc.stats()
    assert st["num_docs"] == len(docs)
    assert st["total_tokens"] == int(sum(ref.doc_lengths)) and st["n_terms"] == len(ref.doc_freq)
    assert st["n_postings"] == int(sum(ref.doc_freq.values()))
    assert np.float32(st["avg_doc_len"]).view(np.uint32) == np.float32(ref.avg_doc_len).view(np.uint32)
    for q in ["rust", "RUST ab", "zz" * 300, "q" * 4095, "w" * 10000, "tail end cd ef", "needle pad", "42 a1", "x", "", "東京", "naïve", "ve na"]:
        assert np.array_equal(sc.score_query(q).view(np.uint32), ref.score_query_fast(q).view(np.uint32)), q[:20]
    empty = pkg.Bm25Scorer.build([])
    assert len(empty) == 0 and empty.search("anything", 3) == []
    blank = pkg.Bm25Scorer.build(["", " ", "a b"])
    assert blank.stats()["n_terms"] == 0 and blank.search("a b", 3) == [] and blank.score_query("ab").tolist() == [0.0, 0.0, 0.0]
