"""Generates tests/golden/text_golden.json with the text oracle (oracle/text_oracle.py), which is pinned
by the ported reference unit tests in tests/test_text_oracle.py. Run: python tests/golden/make_text_golden.py
The product (C ABI / GPU) is compared against these fixtures in tests/test_text_abi.py and test_text_gpu.py."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import text_oracle as T  # noqa: E402

rng = np.random.default_rng(7)
vocab = [f"w{i}" for i in range(300)] + ["Rust", "HNSW", "vector", "x", "a", "B2", "naive-bayes", "snake_case", "über"]


def doc(n):
    z = rng.zipf(1.3, size=n)
    return " ".join(vocab[int(v) % len(vocab)] for v in z)


bm25 = [
    {"docs": ["apple banana", "apple cherry", "banana cherry", "apple apple apple"], "queries": ["apple", "banana cherry", "", "zzz", "apple apple"]},
    {"docs": ["the quick brown fox jumps over the lazy dog", "a quick brown dog outpaces a swift fox", "the dog chases the fox around the yard"],
     "queries": ["quick fox", "the", "Dog, DOG; dog!"]},
    {"docs": [doc(int(rng.integers(1, 60))) for _ in range(200)] + ["", "a b c"], "queries": [doc(int(rng.integers(1, 6))) for _ in range(12)]},
]
for c in bm25:
    sc = T.Bm25Scorer(c["docs"])
    c["scores_bits"] = [sc.score_query(q).view(np.uint32).tolist() for q in c["queries"]]
    c["top5"] = [[[i, int(np.float32(s).view(np.uint32))] for i, s in sc.search(q, 5)] for q in c["queries"]]
    c["avg_doc_len_bits"] = int(np.float32(sc.avg_doc_len).view(np.uint32))

metadata = [
    {"source": "main.rs", "type": "code", "lines": 100},
    {"source": "/path/to/main.rs", "chunk_type": "ast", "lines": 7, "name": "parse", "language": "rust"},
    {"source": "docs/readme.md", "type": "text", "lines": 50.5, "nested": {"a": {"b": 3}}, "flag": True},
    {"type": "doc", "lines": "many", "tags": ["x", "y"], "none": None},
    {},
    None,
    {"source": "dir3/f12.py", "chunk_index": 4, "chunk_type": "simple", "lines": 490},
]
exprs = [
    "source:*.rs", "type=code", "lines>50", "type in [code,text,doc]", "type in [text,doc]", "type not_in [text,doc]",
    "type not_in [code,text]", "type=code,lines>50", "type=code AND lines>50", "type=code,lines>200", "type=code OR type=text",
    "type=text OR type=doc", "source~main", "source:*main*", "source?", "missing?", "lines>=490", "lines<=7", "lines<50.5",
    "lines!=100", "source^docs", "source$.py", "source:dir*", "nested.a.b=3", "nested.a.b>2,flag=true", "flag=false",
    "lines>=abc", "lines>abc", "type = code", "a^b>=3", "lines in [7, 100,490]", "chunk_type=ast,lines>5 OR type=doc",
    "nonsense", "", "   ", ",", "x=1,,y=2", "lines=1e2", "lines=100.0", "none?", "tags=x", "source:*", "source:**", "a=b=c",
    "type in [code", "lines not_in [100]", "chunk_index>=4 AND chunk_type=simple AND lines>=490", "name~ars OR language$ust",
]
filters = []
for e in exprs:
    f = T.parse_filter(e)
    filters.append({"expr": e, "tree": f, "matches": None if f is None else [T.filter_matches(f, m) for m in metadata]})

tok = ["Hello, World! This is a test.", "", "test123 456abc", "a I x9 A1b __init__ fooBar", "naïve café 東京 abc", "x" * 3 + " y z zz"]
out = {"bm25": bm25, "metadata": metadata, "filters": filters, "tokenize": [{"text": t, "tokens": T.tokenize(t)} for t in tok],
       "hybrid": [
           {"vec": [[0, 0.9], [1, 0.8], [2, 0.7]], "bm25": [0.5, 0.9, 0.3], "alpha": 0.5},
           {"vec": [[0, 0.9], [1, 0.5]], "bm25": [0.1, 0.9], "alpha": 1.0},
           {"vec": [[0, 0.9], [1, 0.5]], "bm25": [0.1, 0.9], "alpha": 0.0},
           {"vec": [[5, 0.25], [2, 0.25], [9, 0.0], [7, 0.31]], "bm25": [0.0, 0.0, 1.5, 0.0, 0.0, 0.2, 0.0, 0.0], "alpha": 0.7},
           {"vec": [[1, 0.4]], "bm25": [0.0, 0.0], "alpha": 0.5},
       ]}
for h in out["hybrid"]:
    r = T.hybrid_rerank([(i, s) for i, s in h["vec"]], np.asarray(h["bm25"], dtype=np.float32), h["alpha"])
    h["result"] = [[i, int(np.float32(s).view(np.uint32))] for i, s in r]
json.dump(out, open(os.path.join(os.path.dirname(__file__), "text_golden.json"), "w"), indent=1)
print("written", len(filters), "filters")
