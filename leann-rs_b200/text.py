"""Host mirror of src/index/{bm25,filter,searcher}.rs over the C ABI (filled in with the text path)."""
from __future__ import annotations


class _Pending:
    def __init__(self, *a, **k):
        raise NotImplementedError("text path lands with bm25.cu/text.cpp")


Bm25Scorer = MetadataFilter = IndexSearcher = SearchOptions = SearchResult = _Pending


def hybrid_rerank(*a, **k):
    raise NotImplementedError


def tokenize(*a, **k):
    raise NotImplementedError
