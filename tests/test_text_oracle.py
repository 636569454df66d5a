"""Pins the text oracle (oracle/text_oracle.py) against the reference's own unit tests — every test of
src/index/bm25.rs:176-329 and src/index/filter.rs:446-551 is ported below with the same inputs and
assertions — and against the f32 known-answer values of SURVEY.md §4 / tests/golden/text_golden.json."""
import json
import os

import numpy as np
import pytest

from oracle import text_oracle as T

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "text_golden.json")))


# ---- bm25.rs tests ---------------------------------------------------------------------------------
def test_tokenize_basic():  # bm25.rs:176-184
    t = T.tokenize("Hello, World! This is a test.")
    assert "hello" in t and "world" in t and "test" in t and "a" not in t


def test_tokenize_empty():  # bm25.rs:186-190
    assert T.tokenize("") == []


def test_tokenize_numbers():  # bm25.rs:192-197
    t = T.tokenize("test123 456abc")
    assert "test123" in t and "456abc" in t


FOX = ["the quick brown fox jumps over the lazy dog", "a quick brown dog outpaces a swift fox",
       "the dog chases the fox around the yard"]


def test_bm25_basic_scoring():  # bm25.rs:199-213
    r = T.Bm25Scorer(FOX).search("quick fox", 3)
    assert r and len(r) <= 3


def test_bm25_term_frequency_matters():  # bm25.rs:215-227
    s = T.Bm25Scorer(["rust rust rust programming", "rust programming"]).score_query("rust")
    assert s[0] > s[1]


def test_bm25_idf_matters():  # bm25.rs:229-244
    s = T.Bm25Scorer(["common rare", "common", "common"]).score_query("rare")
    assert s[0] > 0 and s[1] == 0 and s[2] == 0


def test_bm25_empty_query():  # bm25.rs:246-253
    assert T.Bm25Scorer(["hello world"]).score_query("")[0] == 0


def test_bm25_no_match():  # bm25.rs:255-262
    assert T.Bm25Scorer(["hello world"]).search("xyz", 5) == []


APPLE = ["apple banana", "apple cherry", "banana cherry", "apple apple apple"]


def test_bm25_search_top_k():  # bm25.rs:264-280
    r = T.Bm25Scorer(APPLE).search("apple", 2)
    assert len(r) == 2 and r[0][0] == 3


def test_hybrid_rerank_basic():  # bm25.rs:282-299
    r = T.hybrid_rerank([(0, 0.9), (1, 0.8), (2, 0.7)], [0.5, 0.9, 0.3], 0.5)
    assert len(r) == 3 and all(0.0 <= s <= 1.0000001 for _, s in r)


def test_hybrid_rerank_vector_only():  # bm25.rs:301-314
    assert T.hybrid_rerank([(0, 0.9), (1, 0.5)], [0.1, 0.9], 1.0)[0][0] == 0


def test_hybrid_rerank_bm25_only():  # bm25.rs:316-329
    assert T.hybrid_rerank([(0, 0.9), (1, 0.5)], [0.1, 0.9], 0.0)[0][0] == 1


# ---- known-answer f32 values -----------------------------------------------------
def test_golden_bm25_values():
    s = T.Bm25Scorer(APPLE).score_query("apple")
    assert np.array_equal(s, np.array([0.37365952, 0.37365952, 0.0, 0.5231233], dtype=np.float32))
    assert [i for i, _ in T.Bm25Scorer(APPLE).search("apple", 2)] == [3, 0]   # tie 0/1 -> lower index (stable sort)
    sc = T.Bm25Scorer(FOX)
    assert sc.doc_lengths == [9, 6, 8]                                          # "a" dropped
    assert np.array_equal(sc.score_query("quick fox"), np.array([0.56344783, 0.6624485, 0.13119788], dtype=np.float32))
    # SURVEY.md §4 lists [0.26740503, 0.21110922] for this case: those came from numpy's SIMD float32 log,
    # which is 1 ulp off for ln(1.2f). Rust's f32::ln lowers to libm logf (correctly rounded here), so the
    # reference produces the values below (bit patterns 0x3E88E94F, 0x3E582D03).
    assert T.Bm25Scorer(["rust rust rust programming", "rust programming"]).score_query("rust").view(np.uint32).tolist() == [
        1049160015, 1045966083]
    assert np.array_equal(T.Bm25Scorer(["common rare", "common", "common"]).score_query("rare"),
                          np.array([0.8142733, 0.0, 0.0], dtype=np.float32))


def test_golden_hybrid_values():
    r = T.hybrid_rerank([(0, 0.9), (1, 0.8), (2, 0.7)], [0.5, 0.9, 0.3], 0.5)
    assert [i for i, _ in r] == [1, 0, 2]
    assert np.array_equal(np.array([s for _, s in r], dtype=np.float32), np.array([0.7500001, 0.6666667, 0.0], dtype=np.float32))
    assert [(i, float(s)) for i, s in T.hybrid_rerank([(0, 0.9), (1, 0.5)], [0.1, 0.9], 1.0)] == [(0, 1.0), (1, 0.0)]
    assert [(i, float(s)) for i, s in T.hybrid_rerank([(0, 0.9), (1, 0.5)], [0.1, 0.9], 0.0)] == [(1, 1.0), (0, 0.0)]


def test_golden_file_matches_oracle():
    for case in GOLD["bm25"]:
        sc = T.Bm25Scorer(case["docs"])
        for q, want in zip(case["queries"], case["scores_bits"]):
            assert sc.score_query(q).view(np.uint32).tolist() == want
            assert sc.score_query_fast(q).view(np.uint32).tolist() == want
    for case in GOLD["filters"]:
        f = T.parse_filter(case["expr"])
        assert f == case["tree"]
        if f is not None:
            assert [T.filter_matches(f, m) for m in GOLD["metadata"]] == case["matches"]


def test_duplicate_query_tokens_count_twice():  # bm25.rs:81 (Q4)
    sc = T.Bm25Scorer(APPLE)
    a, b = sc.score_query("apple"), sc.score_query("apple apple")
    assert np.array_equal(b, (a + a).astype(np.float32))


# ---- filter.rs tests --------------------------------------------------------------------------------------
def test_filter_parse():  # filter.rs:446-450
    assert "field" in T.parse_filter("source:*.rs")


def test_filter_matches():  # filter.rs:452-469
    md = {"source": "main.rs", "type": "code", "lines": 100}
    for e in ("source:*.rs", "type=code", "lines>50"):
        assert T.filter_matches(T.parse_filter(e), md)


def test_filter_in():  # filter.rs:471-483
    md = {"type": "code", "lang": "rust"}
    assert T.filter_matches(T.parse_filter("type in [code,text,doc]"), md)
    assert not T.filter_matches(T.parse_filter("type in [text,doc]"), md)


def test_filter_not_in():  # filter.rs:485-496
    md = {"type": "code"}
    assert T.filter_matches(T.parse_filter("type not_in [text,doc]"), md)
    assert not T.filter_matches(T.parse_filter("type not_in [code,text]"), md)


def test_filter_and():  # filter.rs:498-514
    md = {"type": "code", "lines": 100}
    assert T.filter_matches(T.parse_filter("type=code,lines>50"), md)
    assert T.filter_matches(T.parse_filter("type=code AND lines>50"), md)
    assert not T.filter_matches(T.parse_filter("type=code,lines>200"), md)


def test_filter_or():  # filter.rs:516-527
    md = {"type": "code"}
    assert T.filter_matches(T.parse_filter("type=code OR type=text"), md)
    assert not T.filter_matches(T.parse_filter("type=text OR type=doc"), md)


def test_filter_contains():  # filter.rs:529-540
    md = {"source": "/path/to/main.rs"}
    assert T.filter_matches(T.parse_filter("source~main"), md)
    assert T.filter_matches(T.parse_filter("source:*main*"), md)


def test_filter_exists():  # filter.rs:542-551
    md = {"source": "main.rs"}
    assert T.filter_matches(T.parse_filter("source?"), md)
    assert not T.filter_matches(T.parse_filter("missing?"), md)


def test_filter_quirks():
    # type mismatch compares as Equal: Gte/Lte true, Gt/Lt false (filter.rs:402-418)
    assert T.filter_matches(T.parse_filter("lines>=abc"), {"lines": 5})
    assert not T.filter_matches(T.parse_filter("lines>abc"), {"lines": 5})
    # Ne / NotIn are true when the field is missing (filter.rs:335,349)
    assert T.filter_matches(T.parse_filter("x!=1"), {}) and T.filter_matches(T.parse_filter("x not_in [1]"), {})
    # no trimming around '=' (filter.rs:273-286)
    assert T.parse_filter("type = code") == {"field": "type ", "op": "eq", "value": " code"}
    # '^' is ignored when '>=' is present (filter.rs:197)
    assert T.parse_filter("a^b>=3")["op"] == "gte"
    assert T.parse_value("1e3") == 1000.0 and T.parse_value("inf") == "inf" and T.parse_value("+7") == 7
    assert T.filter_matches(T.parse_filter("a.b.c=1"), {"a": {"b": {"c": 1.0}}})
