"""oracle/text_oracle.py — TEST INFRASTRUCTURE ONLY (not part of the product).

Python/numpy restatement of the reference's in-repo text arithmetic, every function citing the
leann-rs source it follows. All floating point is IEEE f32 via numpy scalars/arrays (no FMA
contraction, same operation order as the Rust source); `ln` is glibc `logf`, which is what Rust's
`f32::ln` lowers to on Linux.

PARITY STATUS: pinned. These functions are checked against the reference's own unit tests
(src/index/bm25.rs:176-329, src/index/filter.rs:446-551 — ported in tests/test_text_oracle.py) and
against the f32 values recorded in SURVEY.md §4.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import json
import re
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32
K1 = f32(1.2)   # bm25.rs:9
B = f32(0.75)   # bm25.rs:10

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.logf.restype = ctypes.c_float
_libm.logf.argtypes = [ctypes.c_float]


def logf(x) -> np.float32:
    return f32(_libm.logf(float(f32(x))))


_TOKEN = re.compile(r"[a-zA-Z0-9]+")   # bm25.rs:13-15


def tokenize(text: str) -> List[str]:
    """bm25.rs:127-132: regex matches, lower-cased, byte length > 1."""
    return [m.lower() for m in _TOKEN.findall(text) if len(m) > 1]


class Bm25Scorer:
    """bm25.rs:17-122, data structures as in the reference (per-document term-frequency maps)."""

    def __init__(self, documents: Sequence[str]):  # build, bm25.rs:33-74
        self.num_docs = len(documents)
        self.doc_freq = {}
        self.doc_lengths = []
        self.term_freqs = []
        total = 0
        for doc in documents:
            toks = tokenize(doc)
            self.doc_lengths.append(len(toks))
            total += len(toks)
            tf = {}
            for t in toks:
                tf[t] = tf.get(t, 0) + 1
            for t in tf:
                self.doc_freq[t] = self.doc_freq.get(t, 0) + 1
            self.term_freqs.append(tf)
        self.avg_doc_len = f32(total) / f32(self.num_docs) if self.num_docs > 0 else f32(1.0)
        # inverted view for the vectorised scorer (same numbers, different traversal order)
        self._post = {}
        for d, tf in enumerate(self.term_freqs):
            for t, c in tf.items():
                self._post.setdefault(t, ([], []))
                self._post[t][0].append(d)
                self._post[t][1].append(c)

    def idf(self, token: str) -> np.float32:
        df = f32(self.doc_freq.get(token, 0))
        return logf((f32(self.num_docs) - df + f32(0.5)) / (df + f32(0.5)) + f32(1.0))  # bm25.rs:88

    def score_query(self, query: str) -> np.ndarray:
        """bm25.rs:77-106, literal loop (small corpora)."""
        scores = np.zeros(self.num_docs, dtype=np.float32)
        for token in tokenize(query):
            df = f32(self.doc_freq.get(token, 0))
            if df == 0:
                continue
            idf = self.idf(token)
            for doc_id, tf_map in enumerate(self.term_freqs):
                tf = f32(tf_map.get(token, 0))
                if tf == 0:
                    continue
                doc_len = f32(self.doc_lengths[doc_id])
                norm = f32(1.0) - B + B * (doc_len / self.avg_doc_len)
                score = idf * (tf * (K1 + f32(1.0))) / (tf + K1 * norm)
                scores[doc_id] = scores[doc_id] + score
        return scores

    def score_query_fast(self, query: str) -> np.ndarray:
        """Same arithmetic, vectorised over each token's postings (large corpora). Per document the
        contributions are still added in query-token order, so the f32 results are identical."""
        scores = np.zeros(self.num_docs, dtype=np.float32)
        lens = np.asarray(self.doc_lengths, dtype=np.float32)
        for token in tokenize(query):
            if token not in self._post:
                continue
            docs = np.asarray(self._post[token][0], dtype=np.int64)
            tf = np.asarray(self._post[token][1], dtype=np.float32)
            idf = self.idf(token)
            norm = (f32(1.0) - B) + B * (lens[docs] / self.avg_doc_len)
            s = (idf * (tf * (K1 + f32(1.0)))) / (tf + K1 * norm)
            scores[docs] = scores[docs] + s.astype(np.float32)
        return scores

    def search(self, query: str, top_k: int, fast: bool = False) -> List[Tuple[int, np.float32]]:
        """bm25.rs:109-122: positives only, stable sort descending, truncate."""
        scores = self.score_query_fast(query) if fast else self.score_query(query)
        pos = np.nonzero(scores > 0)[0]
        order = pos[np.argsort(-scores[pos], kind="stable")][:top_k]
        return [(int(i), scores[i]) for i in order]


def hybrid_rerank(vector_results: Sequence[Tuple[int, float]], bm25_scores: np.ndarray, alpha) -> List[Tuple[int, np.float32]]:
    """bm25.rs:135-170."""
    alpha = f32(alpha)
    vs = [f32(s) for _, s in vector_results]
    max_v = f32(-np.inf)
    min_v = f32(np.inf)
    for s in vs:
        max_v = max(max_v, s)
        min_v = min(min_v, s)
    with np.errstate(invalid="ignore"):
        v_range = max(f32(max_v - min_v), f32(1e-6))
    bm = np.asarray(bm25_scores, dtype=np.float32)
    max_b = f32(bm.max()) if bm.size else f32(-np.inf)
    min_b = f32(bm.min()) if bm.size else f32(np.inf)
    with np.errstate(invalid="ignore"):
        b_range = max(f32(max_b - min_b), f32(1e-6))
    combined = []
    for (idx, _), v in zip(vector_results, vs):
        norm_vec = (v - min_v) / v_range
        b = bm[idx] if 0 <= idx < bm.size else f32(0.0)
        norm_b = (b - min_b) / b_range
        combined.append((int(idx), f32(alpha * norm_vec) + f32((f32(1.0) - alpha) * norm_b)))
    order = sorted(range(len(combined)), key=lambda i: -float(combined[i][1]))  # stable
    return [combined[i] for i in order]


# ------------------------------------------------------------------------------------------------
# filter.rs
# ------------------------------------------------------------------------------------------------
def _parse_i64(s: str):
    if re.fullmatch(r"[+-]?[0-9]+", s):
        v = int(s)
        if -(2 ** 63) <= v < 2 ** 63:
            return v
    return None


_F64 = re.compile(r"[+-]?(?:[0-9]+\.?[0-9]*|\.[0-9]+)(?:[eE][+-]?[0-9]+)?")


def parse_value(s: str) -> Any:  # filter.rs:420-439
    v = _parse_i64(s)
    if v is not None:
        return v
    if _F64.fullmatch(s):
        x = float(s)
        if np.isfinite(x):
            return x
    if s == "true":
        return True
    if s == "false":
        return False
    return s


# Rust str::trim strips exactly the chars with the Unicode White_Space property (Python's str.strip() would also strip
# U+001C..U+001F, which Rust keeps).
_RUST_WS = "\t\n\x0b\x0c\r \x85\xa0\u1680\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008\u2009\u200a\u2028\u2029\u202f\u205f\u3000"


def _cond(field, op, value):
    return {"field": field, "op": op, "value": value}


def parse_single(s: str) -> Optional[dict]:  # filter.rs:137-316
    s = s.strip(_RUST_WS)
    if s.endswith("?"):
        return _cond(s[:-1], "exists", None)
    for key, op in ((" in [", "in"), (" not_in [", "notin")):
        idx = s.find(key)
        if idx >= 0:
            field = s[:idx].strip(_RUST_WS)
            rest = s[idx + len(key):]
            end = rest.find("]")
            if end >= 0:
                return _cond(field, op, [parse_value(v.strip(_RUST_WS)) for v in rest[:end].split(",")])
    if "~" in s:
        a, b = s.split("~", 1)
        return _cond(a, "contains", b)
    if "^" in s and ">=" not in s:
        a, b = s.split("^", 1)
        return _cond(a, "startswith", b)
    if "$" in s:
        a, b = s.split("$", 1)
        return _cond(a, "endswith", b)
    for sep, op in (("!=", "ne"), (">=", "gte"), ("<=", "lte"), (">", "gt"), ("<", "lt")):
        if sep in s:
            a, b = s.split(sep, 1)
            return _cond(a, op, parse_value(b))
    if "=" in s:
        field, value = s.split("=", 1)
    elif ":" in s:
        field, value = s.split(":", 1)
    else:
        return None
    if "*" in value:
        if value.startswith("*") and value.endswith("*") and len(value) > 2:
            return _cond(field, "contains", value[1:-1])
        if value.startswith("*"):
            return _cond(field, "endswith", value[1:])
        if value.endswith("*"):
            return _cond(field, "startswith", value[:-1])
    return _cond(field, "eq", parse_value(value))


def parse_filter(s: str) -> Optional[dict]:  # filter.rs:52-134
    s = s.strip(_RUST_WS)
    if " OR " in s:
        fs = [f for f in (parse_filter(p.strip(_RUST_WS)) for p in s.split(" OR ")) if f is not None]
        if len(fs) > 1:
            return {"or": fs}
        return fs[0] if fs else None
    has_and = " AND " in s
    depth, has_comma = 0, False
    for c in s:
        if c == "[":
            depth += 1
        elif c == "]":
            depth -= 1
        elif c == "," and depth == 0:
            has_comma = True
            break
    if has_and or has_comma:
        if has_and:
            parts = s.split(" AND ")
        else:
            parts, cur, depth = [], "", 0
            for c in s:
                if c == "[":
                    depth += 1
                    cur += c
                elif c == "]":
                    depth -= 1
                    cur += c
                elif c == "," and depth == 0:
                    parts.append(cur)
                    cur = ""
                else:
                    cur += c
            if cur:
                parts.append(cur)
        fs = [f for f in (parse_single(p.strip(_RUST_WS)) for p in parts) if f is not None]
        if len(fs) > 1:
            return {"and": fs}
        return fs[0] if fs else None
    return parse_single(s)


_MISSING = object()


def _nested(md, path):  # filter.rs:376-388
    cur = md
    for part in path.split("."):
        if isinstance(cur, dict) and part in cur:
            cur = cur[part]
        else:
            return _MISSING
    return cur


def _is_num(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool)


def values_equal(a, b) -> bool:  # filter.rs:390-400
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    if _is_num(a) and _is_num(b):
        return abs(float(a) - float(b)) < 2.220446049250313e-16
    if isinstance(a, bool) and isinstance(b, bool):
        return a == b
    if a is None and b is None:
        return True
    return False


def compare_values(a, b) -> int:  # filter.rs:402-418
    if _is_num(a) and _is_num(b):
        x, y = float(a), float(b)
        return -1 if x < y else (1 if x > y else 0)
    if isinstance(a, str) and isinstance(b, str):
        x, y = a.encode(), b.encode()
        return -1 if x < y else (1 if x > y else 0)
    return 0


def filter_matches(f: dict, md) -> bool:  # filter.rs:319-373
    if "and" in f:
        return all(filter_matches(c, md) for c in f["and"])
    if "or" in f:
        return any(filter_matches(c, md) for c in f["or"])
    fv = _nested(md, f["field"])
    op, val = f["op"], f["value"]
    have = fv is not _MISSING
    pat = val if isinstance(val, str) else ""
    if op == "exists":
        return have
    if op == "eq":
        return have and values_equal(fv, val)
    if op == "ne":
        return (not have) or not values_equal(fv, val)
    if op == "gt":
        return have and compare_values(fv, val) > 0
    if op == "gte":
        return have and compare_values(fv, val) >= 0
    if op == "lt":
        return have and compare_values(fv, val) < 0
    if op == "lte":
        return have and compare_values(fv, val) <= 0
    if op == "in":
        return isinstance(val, list) and have and any(values_equal(fv, it) for it in val)
    if op == "notin":
        if not isinstance(val, list):
            return True
        return (not have) or not any(values_equal(fv, it) for it in val)
    if op == "contains":
        return have and isinstance(fv, str) and pat in fv
    if op == "startswith":
        return have and isinstance(fv, str) and fv.startswith(pat)
    if op == "endswith":
        return have and isinstance(fv, str) and fv.endswith(pat)
    raise ValueError(op)


# ------------------------------------------------------------------------------------------------
# searcher.rs:123-210 — the glue, given the backend's answer for fetch_k
# ------------------------------------------------------------------------------------------------
def fetch_k(top_k: int, has_filter: bool, hybrid: bool) -> int:  # searcher.rs:129-133
    return top_k * 5 if (has_filter or hybrid) else top_k


def search_with_options(backend_idx: Sequence[int], backend_dist: Sequence[float], top_k: int,
                        scorer: Optional[Bm25Scorer], query_text: Optional[str], hybrid: bool, alpha,
                        passes, fk: int, fast: bool = False) -> List[Tuple[int, np.float32]]:
    """`passes(idx) -> bool` stands for "passage loads and the filter (if any) matches"
    (searcher.rs:186-205). Returns [(idx, score)] of length <= top_k."""
    results = [(int(i), f32(d)) for i, d in zip(backend_idx, backend_dist)]
    if hybrid and query_text is not None:
        bm = scorer.score_query_fast(query_text) if fast else scorer.score_query(query_text)
        top = scorer.search(query_text, fk, fast=fast)
        seen = {i for i, _ in results}
        for i, _ in top:
            if i not in seen:
                results.append((i, f32(0.0)))
        results = hybrid_rerank(results, bm, alpha)
    out = []
    for idx, score in results:
        if len(out) >= top_k:
            break
        if not passes(idx):
            continue
        out.append((idx, score))
    return out


def describe(f: dict) -> str:
    return json.dumps(f, sort_keys=True)
