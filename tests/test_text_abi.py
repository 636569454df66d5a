"""CPU tests of the C ABI's host-side text logic (no GPU compute): tokenizer and filter DSL against the
golden fixtures and the oracle, export of every declared symbol, and loud failure without a GPU."""
import json
import os
import re
import subprocess

import pytest
from hypothesis import given, settings, strategies as st

from oracle import text_oracle as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "text_golden.json")))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "leann_cuda.h")).read()
    declared = sorted(set(re.findall(r"\b(leann_cuda_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 38
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (leann_cuda_[a-z0-9_]+)", out))
    missing = [d for d in declared if d not in exported]
    assert not missing, f"declared in include/leann_cuda.h but not exported: {missing}"
    L = pkg.lib()
    for d in declared:
        assert getattr(L, d) is not None
    assert b"sm_100a" in L.leann_cuda_version()


def test_tokenize_golden(pkg):
    for c in GOLD["tokenize"]:
        assert pkg.tokenize(c["text"]) == c["tokens"]


@settings(max_examples=300, deadline=None)
@given(st.text(alphabet=st.sampled_from(list("abcXYZ019 _-.,!?é東\n\t")), max_size=60))
def test_tokenize_matches_oracle(text):
    import leann_rs_b200 as P
    assert P.tokenize(text) == T.tokenize(text)


def test_filter_golden(pkg):
    for c in GOLD["filters"]:
        f = pkg.MetadataFilter.parse(c["expr"])
        if c["tree"] is None:
            assert f is None, c["expr"]
            continue
        assert f is not None, c["expr"]
        assert f.describe() == c["tree"], c["expr"]
        assert [f.matches(m) for m in GOLD["metadata"]] == c["matches"], c["expr"]
        words = f.mask(GOLD["metadata"])
        assert [bool((int(words[i // 64]) >> (i % 64)) & 1) for i in range(len(GOLD["metadata"]))] == c["matches"]


def test_filter_trim_is_rust_unicode_white_space(pkg):
    """filter.rs trims with str::trim = the Unicode White_Space property: NBSP, NEL, U+1680, U+2000..200A, U+2028/9, U+202F,
    U+205F and U+3000 are stripped like ASCII blanks; U+001C..U+001F and U+200B (zero width space) are NOT white space in
    Rust (Python's str.strip() would strip the former), so they stay part of the field / value."""
    ws = ["\u00a0", "\u0085", "\u1680", "\u2000", "\u2005", "\u200a", "\u2028", "\u2029", "\u202f", "\u205f", "\u3000", " \t\u00a0\r\n"]
    for w in ws:
        for expr in (f"{w}type=code{w}", f"type=code,{w}lines>3{w}", f"type{w} in [{w}code{w},{w}5{w}]", f"{w}lines>3 OR{w} {w}type=5{w}"):
            ref = T.parse_filter(expr)
            f = pkg.MetadataFilter.parse(expr)
            assert ref is not None and f is not None, repr(expr)
            assert f.describe() == ref, repr(expr)
            assert "\u00a0" not in str(ref) and "\u3000" not in str(ref)
    plain = pkg.MetadataFilter.parse("type=code").describe()
    assert pkg.MetadataFilter.parse("\u3000type=code\u2003").describe() == plain
    for keep in ("\x1c", "\x1f", "\u200b", "\u180e"):
        expr = f"{keep}type=code"
        ref = T.parse_filter(expr)
        f = pkg.MetadataFilter.parse(expr)
        assert f is not None and f.describe() == ref
        assert f.describe() != plain and not f.matches({"type": "code"})      # the field name keeps the character


_field = st.sampled_from(["type", "lines", "source", "a.b", "flag", "x"])
_val = st.sampled_from(["code", "5", "5.0", "-3", "1e2", "true", "false", "*.rs", "*ain*", "ma*", "", " 7", "abc", "inf", "+4", "\u00a07", "\u30005\u2009",
                        "\x1f5"])
_op = st.sampled_from(["=", ":", ">", "<", ">=", "<=", "!=", "~", "^", "$", " in [", " not_in ["])


@st.composite
def _expr(draw):
    parts = []
    for _ in range(draw(st.integers(1, 3))):
        f, o, v = draw(_field), draw(_op), draw(_val)
        if o.endswith("["):
            parts.append(f + o + ",".join(draw(st.lists(_val, min_size=0, max_size=3))) + draw(st.sampled_from(["]", ""])))
        elif draw(st.booleans()) and o == "=":
            parts.append(f + "?")
        else:
            parts.append(f + o + v)
    return draw(st.sampled_from([",", " AND ", " OR "])).join(parts)


_MD = [{"type": "code", "lines": 5, "source": "main.rs", "a": {"b": 5.0}, "flag": True}, {"type": "5", "lines": "5", "x": None},
       {"lines": 100.0, "source": "src/domain.py"}, {}]


@settings(max_examples=400, deadline=None)
@given(_expr())
def test_filter_matches_oracle(expr):
    import leann_rs_b200 as P
    ref = T.parse_filter(expr)
    f = P.MetadataFilter.parse(expr)
    if ref is None:
        assert f is None
        return
    assert f is not None and f.describe() == ref
    for md in _MD:
        assert f.matches(md) == T.filter_matches(ref, md), (expr, md)


def _bits(words, n):
    return [bool((int(words[i // 64]) >> (i % 64)) & 1) for i in range(n)]


_MD_COLS = _MD + [None, {"a": {"b": "x", "c": {"d": 1}}, "a.b": 7, "": {"": 3}}, {"a": 5, "lines": [1, 2], "source": {"k": 1}},
                  {"lines": -3, "flag": False, "type": None}, {"type": "code", "lines": 5.0000000000000001, "x": "abc"},
                  "not an object", 17, {"source": "main.rs", "lines": 9007199254740993}]


def test_metadata_columns_golden(pkg):
    """Column-wise evaluation over the side-car == MetadataFilter.matches row by row on the golden metadata."""
    cols = pkg.MetadataColumns(GOLD["metadata"])
    assert cols.fields >= 3
    for c in GOLD["filters"]:
        if c["tree"] is None:
            continue
        f = pkg.MetadataFilter.parse(c["expr"])
        assert _bits(cols.mask(f), len(GOLD["metadata"])) == c["matches"], c["expr"]


@settings(max_examples=400, deadline=None)
@given(_expr())
def test_metadata_columns_match_rowwise_and_oracle(expr):
    import leann_rs_b200 as P
    ref = T.parse_filter(expr)
    f = P.MetadataFilter.parse(expr)
    if ref is None:
        assert f is None
        return
    cols = _cols_cache.setdefault("c", P.MetadataColumns([None if m is None else json.dumps(m) for m in _MD_COLS]))
    got = _bits(cols.mask(f), len(_MD_COLS))
    want = [T.filter_matches(ref, md) for md in _MD_COLS]     # passages without metadata are matched against null
    assert got == want, (expr, got, want)
    rows = _bits(f.mask(["null" if m is None else json.dumps(m) for m in _MD_COLS]), len(_MD_COLS))
    assert got == rows, expr


_cols_cache = {}


def test_metadata_columns_paths_and_sizes(pkg):
    f = pkg.MetadataFilter.parse
    cols = pkg.MetadataColumns([None if m is None else json.dumps(m) for m in _MD_COLS])
    n = len(_MD_COLS)
    pick = lambda e: [i for i, b in enumerate(_bits(cols.mask(f(e)), n)) if b]
    assert pick("a.c.d=1") == [5] and pick("a.b=x") == [5] and pick("a.b=5") == [0]
    assert pick("a?") == [0, 5, 6] and pick("nosuch?") == [] and pick("nosuch!=1") == list(range(n))
    assert pick("lines>=5") == [i for i, m in enumerate(_MD_COLS) if isinstance(m, dict) and "lines" in m and not (isinstance(m["lines"], (int, float)) and not isinstance(m["lines"], bool) and m["lines"] < 5)]
    # ragged sizes around the word boundary, empty set
    for k in (0, 1, 63, 64, 65, 130):
        docs = [{"i": i, "s": "v%d" % (i % 7)} for i in range(k)]
        c = pkg.MetadataColumns(docs)
        w = c.mask(f("i>=10,s!=v3"))
        assert len(w) == (k + 63) // 64
        assert _bits(w, k) == [i >= 10 and i % 7 != 3 for i in range(k)]
        if k % 64:
            assert int(w[-1]) >> (k % 64) == 0       # no stray bits beyond n


def test_text_path_fails_loudly_without_gpu(pkg):
    if pkg.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(pkg.LeannCudaError) as e:
        pkg.Bm25Scorer.build(["hello world"])
    assert e.value.code == pkg.ERR_CUDA and "no CPU fallback" in e.value.message
    with pytest.raises(pkg.LeannCudaError):
        pkg.hybrid_rerank([(0, 0.9)], [0.1], 0.5)
    with pytest.raises(pkg.LeannCudaError):
        pkg.FlatSearcher.from_vectors(__import__("numpy").zeros((4, 8), dtype="float32"))


def _stats_blob(num_docs, total_tokens, df):
    import struct
    out = struct.pack("<QQQQ", 0x314D42534E41454C, num_docs, total_tokens, len(df))
    for term in sorted(df):
        t = term.encode()
        out += struct.pack("<I", len(t)) + t + struct.pack("<Q", df[term])
    return out


def test_bm25_shard_statistics_merge(pkg):
    """Corpus-wide BM25 statistics of document-range shards travel as blobs: merging is a host-side sum over
    N, token count and per-term document frequency, deterministic in its byte output, and rejects malformed input."""
    a = _stats_blob(3, 17, {"fox": 2, "quick": 1, "the": 3})
    b = _stats_blob(2, 9, {"dog": 1, "the": 2, "zebra": 1})
    want = _stats_blob(5, 26, {"fox": 2, "quick": 1, "the": 5, "dog": 1, "zebra": 1})
    assert pkg.Bm25Scorer.merge_stats([a, b]) == want == pkg.Bm25Scorer.merge_stats([b, a])
    assert pkg.Bm25Scorer.merge_stats([a]) == a
    assert pkg.Bm25Scorer.merge_stats([]) == _stats_blob(0, 0, {})
    for bad in (a[:-3], b"\x00" * 32, a + b"x"):
        with pytest.raises(pkg.LeannCudaError) as e:
            pkg.Bm25Scorer.merge_stats([a, bad])
        assert e.value.code == pkg.ERR_BAD_FORMAT


def test_public_header_is_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: include/leann_cuda.h must compile as C99 (no C++ or torch types in the signatures)."""
    src = tmp_path / "abi.c"
    src.write_text('#include "include/leann_cuda.h"\nint main(void) { return leann_cuda_device_count() < 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-c", str(src), "-I", ROOT, "-o", str(tmp_path / "abi.o")])
