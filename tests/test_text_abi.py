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


_field = st.sampled_from(["type", "lines", "source", "a.b", "flag", "x"])
_val = st.sampled_from(["code", "5", "5.0", "-3", "1e2", "true", "false", "*.rs", "*ain*", "ma*", "", " 7", "abc", "inf", "+4"])
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
