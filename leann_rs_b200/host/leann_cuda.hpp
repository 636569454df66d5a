// leann_cuda.hpp — header-only C++ host layer above the C ABI (include/leann_cuda.h), mirroring the
// reference's Rust interfaces one to one (the reference is compiled code and its toolchain is absent
// from this image, so the host side is written in C++; rust/leann-cuda/ holds the same thing as a crate):
//   trait BackendSearcher                 src/backend/traits.rs:11-30
//   HnswSearcher / DiskAnnSearcher        src/backend/hnsw.rs:12-93, src/backend/diskann.rs:12-66
//   BackendType::load_searcher            src/backend/mod.rs:16-45
//   hnsw::{build_index, add_to_index}     src/backend/hnsw.rs:96-191
//   diskann::build_index                  src/backend/diskann.rs:70-105
//   Bm25Scorer, hybrid_rerank             src/index/bm25.rs:17-170
//   MetadataFilter                        src/index/filter.rs:35-39,52-134,319-325
//   SearchOptions, IndexSearcher          src/index/searcher.rs:25-63,66-257
// Errors surface as leann::Error (anyhow::Result in the reference). Nothing here computes.
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/leann_cuda.h"

namespace leann {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc, const char* err) {
    if (rc != LEANN_OK) throw Error(rc, err);
}

// ---- src/backend/traits.rs ------------------------------------------------------------------------
class BackendSearcher {
public:
    virtual ~BackendSearcher() = default;
    /// (indices, distances) for one query; ascending distance, length <= top_k.
    virtual std::pair<std::vector<uint64_t>, std::vector<float>> search(const std::vector<float>& query, size_t top_k,
                                                                        size_t complexity) const = 0;
    virtual size_t len() const = 0;
    bool is_empty() const { return len() == 0; }
};

class CudaSearcher : public BackendSearcher {
public:
    CudaSearcher(const std::string& index_path, int backend, size_t dimensions, int device, std::optional<size_t> fixed_ef)
        : dims_(dimensions), fixed_ef_(fixed_ef) {
        char err[1024];
        check(leann_cuda_open(index_path.c_str(), backend, dimensions, LEANN_METRIC_DEFAULT, device, &h_, err, sizeof err), err);
    }
    ~CudaSearcher() override { leann_cuda_close(h_); }
    CudaSearcher(const CudaSearcher&) = delete;
    CudaSearcher& operator=(const CudaSearcher&) = delete;

    std::pair<std::vector<uint64_t>, std::vector<float>> search(const std::vector<float>& query, size_t top_k,
                                                                size_t complexity) const override {
        std::vector<uint64_t> keys(top_k);
        std::vector<float> dists(top_k);
        uint32_t count = 0;
        char err[1024];
        check(leann_cuda_search(h_, query.data(), 1, top_k, fixed_ef_.value_or(complexity), nullptr, LEANN_MASK_NONE, keys.data(),
                                dists.data(), &count, err, sizeof err), err);
        keys.resize(count);
        dists.resize(count);
        return {std::move(keys), std::move(dists)};
    }
    /// Batched form (new): queries is nq x dims row-major; outputs nq x top_k, counts per query.
    void search_batch(const float* queries, size_t nq, size_t top_k, size_t ef, uint64_t* keys, float* dists, uint32_t* counts,
                      const uint64_t* mask_bits = nullptr) const {
        char err[1024];
        check(leann_cuda_search(h_, queries, nq, top_k, ef, mask_bits, mask_bits ? LEANN_MASK_INLINE : LEANN_MASK_NONE, keys, dists,
                                counts, err, sizeof err), err);
    }
    size_t len() const override { return leann_cuda_len(h_); }
    size_t dims() const { return dims_; }
    /// Merge concurrent single-query calls (the axum handlers of serve.rs) into batched launches.
    void set_coalescing(size_t max_batch, unsigned max_wait_us) { leann_cuda_set_coalescing(h_, max_batch, max_wait_us); }
    /// Visited-set representation of the traversal (tuning hook; results never depend on it): 0 automatic, 1 byte maps only,
    /// 2..5 the shared-memory table forms, >= 1024 per-warp hash tables of that capacity. False for a rejected capacity.
    bool set_visited_hash(size_t capacity) { return leann_cuda_set_visited_hash(h_, capacity) == LEANN_OK; }
    /// Persist the parsed adjacency as `<index_path>.cuda-layout` so later loads stream it instead of parsing the node block.
    void write_layout_cache(const std::string& index_path) const {
        char err[1024];
        check(leann_cuda_write_layout_cache(h_, index_path.c_str(), err, sizeof err), err);
    }
    bool layout_cache_used() const { return leann_cuda_layout_cache_used(h_) != 0; }
    const leann_cuda_index* handle() const { return h_; }

protected:
    leann_cuda_index* h_ = nullptr;
    size_t dims_;
    std::optional<size_t> fixed_ef_;
};

/// src/backend/hnsw.rs: `complexity` is ignored, expansion_search is 64 (hnsw.rs:49,83).
class HnswSearcher : public CudaSearcher {
public:
    static std::unique_ptr<HnswSearcher> load(const std::string& index_path, size_t dimensions, int device = 0) {
        return std::unique_ptr<HnswSearcher>(new HnswSearcher(index_path, dimensions, device));
    }
private:
    HnswSearcher(const std::string& p, size_t d, int dev) : CudaSearcher(p, LEANN_BACKEND_HNSW, d, dev, 64) {}
};
/// src/backend/diskann.rs: beam = max(complexity, top_k) (diskann.rs:54).
class DiskAnnSearcher : public CudaSearcher {
public:
    static std::unique_ptr<DiskAnnSearcher> load(const std::string& index_path, size_t dimensions, int device = 0) {
        return std::unique_ptr<DiskAnnSearcher>(new DiskAnnSearcher(index_path, dimensions, device));
    }
private:
    DiskAnnSearcher(const std::string& p, size_t d, int dev) : CudaSearcher(p, LEANN_BACKEND_VAMANA, d, dev, std::nullopt) {}
};

/// A BackendSearcher over sub-indexes on several GPUs of one box (`leann_cuda_shards_*`): what load_searcher returns
/// when the index directory holds `<base>.shardNN.index` files. One host process owns all devices; the per-shard
/// top-k lists are merged on the first device (peer-memory loads inside the merge kernel, or NCCL all_gather).
class ShardedSearcher : public BackendSearcher {
public:
    ShardedSearcher(const std::vector<std::string>& base_paths, int backend, size_t dimensions, const std::vector<int>& devices,
                    std::optional<size_t> fixed_ef, int exchange = 0)
        : dims_(dimensions), fixed_ef_(fixed_ef) {
        if (base_paths.size() != devices.size()) throw Error(LEANN_ERR_INVALID_ARG, "one device per shard");
        std::vector<const char*> paths;
        for (auto& p : base_paths) paths.push_back(p.c_str());
        char err[1024];
        check(leann_cuda_shards_open(paths.data(), paths.size(), backend, dimensions, LEANN_METRIC_DEFAULT, devices.data(), nullptr,
                                     exchange, &h_, err, sizeof err), err);
    }
    ~ShardedSearcher() override { leann_cuda_shards_close(h_); }
    ShardedSearcher(const ShardedSearcher&) = delete;
    ShardedSearcher& operator=(const ShardedSearcher&) = delete;
    std::pair<std::vector<uint64_t>, std::vector<float>> search(const std::vector<float>& query, size_t top_k,
                                                                size_t complexity) const override {
        std::vector<uint64_t> keys(top_k);
        std::vector<float> dists(top_k);
        uint32_t count = 0;
        char err[1024];
        check(leann_cuda_shards_search(h_, query.data(), 1, top_k, fixed_ef_.value_or(complexity), nullptr, keys.data(), dists.data(),
                                       &count, err, sizeof err), err);
        keys.resize(count);
        dists.resize(count);
        return {std::move(keys), std::move(dists)};
    }
    void search_batch(const float* queries, size_t nq, size_t top_k, size_t ef, uint64_t* keys, float* dists, uint32_t* counts) const {
        char err[1024];
        check(leann_cuda_shards_search(h_, queries, nq, top_k, ef, nullptr, keys, dists, counts, err, sizeof err), err);
    }
    size_t len() const override { return leann_cuda_shards_len(h_); }
    size_t shards() const { return leann_cuda_shards_count(h_); }

private:
    leann_cuda_shards* h_ = nullptr;
    size_t dims_;
    std::optional<size_t> fixed_ef_;
};

// ---- src/backend/mod.rs:16-45 -------------------------------------------------------------------------
enum class BackendType { Hnsw, DiskAnn };
inline std::unique_ptr<BackendSearcher> load_searcher(BackendType t, const std::string& index_path, size_t dimensions, int device = 0) {
    if (t == BackendType::Hnsw) return HnswSearcher::load(index_path, dimensions, device);
    return DiskAnnSearcher::load(index_path, dimensions, device);
}

// ---- src/backend/hnsw.rs:96-191, src/backend/diskann.rs:70-105 -----------------------------------------
namespace detail {
inline std::vector<float> flatten(const std::vector<std::vector<float>>& embeddings, size_t dimensions) {
    std::vector<float> flat;
    flat.reserve(embeddings.size() * dimensions);
    for (auto& e : embeddings) {
        if (e.size() != dimensions) throw Error(LEANN_ERR_DIM_MISMATCH, "Dimension mismatch: expected " + std::to_string(dimensions) + ", got " + std::to_string(e.size()));
        flat.insert(flat.end(), e.begin(), e.end());
    }
    return flat;
}
struct IndexHandle {   // closes on scope exit
    leann_cuda_index* h = nullptr;
    ~IndexHandle() { if (h) leann_cuda_close(h); }
};
}  // namespace detail

namespace hnsw {
/// hnsw.rs:96-139: same arguments, writes `<index_path>.index` (with_extension("index")).
inline void build_index(const std::vector<std::vector<float>>& embeddings, const std::vector<std::string>& /*ids*/, const std::string& index_path,
                        size_t dimensions, size_t graph_degree, size_t complexity, int device = 0) {
    auto flat = detail::flatten(embeddings, dimensions);
    detail::IndexHandle ix;
    char err[1024];
    check(leann_cuda_hnsw_build(flat.data(), 0, embeddings.size(), dimensions, graph_degree, complexity, LEANN_METRIC_DEFAULT, 1, device, &ix.h, err,
                                sizeof err), err);
    check(leann_cuda_save(ix.h, index_path.c_str(), err, sizeof err), err);
}
/// hnsw.rs:142-191: load, append with keys start_id.., save back (connectivity from the file, expansion_add 64).
inline void add_to_index(const std::vector<std::vector<float>>& embeddings, const std::string& index_path, size_t dimensions, size_t start_id,
                         int device = 0) {
    auto flat = detail::flatten(embeddings, dimensions);
    detail::IndexHandle ix;
    char err[1024];
    check(leann_cuda_open(index_path.c_str(), LEANN_BACKEND_HNSW, dimensions, LEANN_METRIC_DEFAULT, device, &ix.h, err, sizeof err), err);
    check(leann_cuda_hnsw_add(ix.h, flat.data(), 0, embeddings.size(), start_id, 64, 1, err, sizeof err), err);
    check(leann_cuda_save(ix.h, index_path.c_str(), err, sizeof err), err);
}
}  // namespace hnsw

namespace diskann {
/// diskann.rs:70-105: writes `<index_path>.diskann`; alpha = 1.2 (diskann.rs:91).
inline void build_index(const std::vector<std::vector<float>>& embeddings, const std::vector<std::string>& /*ids*/, const std::string& index_path,
                        size_t dimensions, size_t graph_degree, size_t complexity, int device = 0) {
    auto flat = detail::flatten(embeddings, dimensions);
    detail::IndexHandle ix;
    char err[1024];
    check(leann_cuda_vamana_build(flat.data(), 0, embeddings.size(), dimensions, graph_degree, complexity, 1.2f, LEANN_METRIC_DEFAULT, 1, device,
                                  &ix.h, err, sizeof err), err);
    check(leann_cuda_save(ix.h, index_path.c_str(), err, sizeof err), err);
}
}  // namespace diskann

// ---- src/index/filter.rs ---------------------------------------------------------------------------------
class MetadataFilter {
public:
    /// nullopt where the reference returns None.
    static std::optional<MetadataFilter> parse(const std::string& filter_str) {
        leann_cuda_filter* f = nullptr;
        char err[1024];
        int rc = leann_cuda_filter_parse(filter_str.c_str(), &f, err, sizeof err);
        if (rc == LEANN_ERR_PARSE) return std::nullopt;
        check(rc, err);
        return MetadataFilter(f);
    }
    bool matches(const std::string& metadata_json) const {
        int r = 0;
        char err[1024];
        check(leann_cuda_filter_matches(f_.get(), metadata_json.data(), metadata_json.size(), &r, err, sizeof err), err);
        return r != 0;
    }
    std::vector<uint64_t> mask(const std::vector<std::string>& metadata_json) const {
        std::vector<const char*> p(metadata_json.size());
        std::vector<size_t> n(metadata_json.size());
        for (size_t i = 0; i < p.size(); ++i) { p[i] = metadata_json[i].data(); n[i] = metadata_json[i].size(); }
        std::vector<uint64_t> out((p.size() + 63) / 64);
        char err[1024];
        check(leann_cuda_filter_mask(f_.get(), p.data(), n.data(), p.size(), out.data(), err, sizeof err), err);
        return out;
    }
    const leann_cuda_filter* handle() const { return f_.get(); }
private:
    explicit MetadataFilter(leann_cuda_filter* f) : f_(f, leann_cuda_filter_free) {}
    std::shared_ptr<leann_cuda_filter> f_;
};

/// Typed columns of the passages' metadata: any parsed filter -> N-bit mask without touching JSON again
/// (replaces the per-candidate passages.get + filter.matches of searcher.rs:186-194).
class MetadataColumns {
public:
    explicit MetadataColumns(const std::vector<std::string>& metadata_json) : n_(metadata_json.size()) {
        std::vector<const char*> p(n_);
        std::vector<size_t> n(n_);
        for (size_t i = 0; i < n_; ++i) { p[i] = metadata_json[i].data(); n[i] = metadata_json[i].size(); }
        leann_cuda_metacols* c = nullptr;
        char err[1024];
        check(leann_cuda_metacols_build(p.data(), n.data(), n_, &c, err, sizeof err), err);
        c_.reset(c, leann_cuda_metacols_free);
    }
    std::vector<uint64_t> mask(const MetadataFilter& f) const {
        std::vector<uint64_t> out((n_ + 63) / 64);
        char err[1024];
        check(leann_cuda_metacols_mask(c_.get(), f.handle(), out.data(), err, sizeof err), err);
        return out;
    }
private:
    size_t n_;
    std::shared_ptr<leann_cuda_metacols> c_;
};

// ---- src/index/bm25.rs -------------------------------------------------------------------------------------
class Bm25Scorer {
public:
    static Bm25Scorer build(const std::vector<std::string>& documents, int device = 0) {
        std::vector<const char*> p(documents.size());
        std::vector<size_t> n(documents.size());
        for (size_t i = 0; i < p.size(); ++i) { p[i] = documents[i].data(); n[i] = documents[i].size(); }
        leann_cuda_bm25* b = nullptr;
        char err[1024];
        check(leann_cuda_bm25_build(p.data(), n.data(), p.size(), device, &b, err, sizeof err), err);
        return Bm25Scorer(b);
    }
    size_t dense_rows() const { return leann_cuda_bm25_dense_rows(b_.get()); }   // frequent terms also kept row-wise (K3d)
    std::vector<float> score_query(const std::string& query) const {
        std::vector<float> s(leann_cuda_bm25_len(b_.get()));
        char err[1024];
        check(leann_cuda_bm25_score(b_.get(), query.data(), query.size(), s.data(), err, sizeof err), err);
        return s;
    }
    std::vector<std::pair<size_t, float>> search(const std::string& query, size_t top_k) const {
        std::vector<uint64_t> idx(top_k);
        std::vector<float> sc(top_k);
        uint32_t cnt = 0;
        const char* q = query.data();
        size_t n = query.size();
        char err[1024];
        check(leann_cuda_bm25_search(b_.get(), &q, &n, 1, top_k, idx.data(), sc.data(), &cnt, err, sizeof err), err);
        std::vector<std::pair<size_t, float>> out;
        for (uint32_t i = 0; i < cnt; ++i) out.emplace_back((size_t)idx[i], sc[i]);
        return out;
    }
    const leann_cuda_bm25* handle() const { return b_.get(); }
private:
    explicit Bm25Scorer(leann_cuda_bm25* b) : b_(b, leann_cuda_bm25_free) {}
    std::shared_ptr<leann_cuda_bm25> b_;
};

inline std::vector<std::pair<size_t, float>> hybrid_rerank(const std::vector<std::pair<size_t, float>>& vector_results,
                                                           const std::vector<float>& bm25_scores, float alpha, int device = 0) {
    std::vector<uint64_t> idx(vector_results.size()), oi(vector_results.size());
    std::vector<float> vs(vector_results.size()), os(vector_results.size());
    for (size_t i = 0; i < idx.size(); ++i) { idx[i] = vector_results[i].first; vs[i] = vector_results[i].second; }
    char err[1024];
    check(leann_cuda_hybrid_rerank(idx.data(), vs.data(), idx.size(), bm25_scores.data(), bm25_scores.size(), alpha, device, oi.data(),
                                   os.data(), err, sizeof err), err);
    std::vector<std::pair<size_t, float>> out;
    for (size_t i = 0; i < oi.size(); ++i) out.emplace_back((size_t)oi[i], os[i]);
    return out;
}

// ---- src/index/searcher.rs -----------------------------------------------------------------------------------
struct SearchOptions {
    size_t top_k = 0, complexity = 0;
    std::optional<std::string> filter;
    bool hybrid = false;
    float hybrid_alpha = 0.7f;  // searcher.rs:47
    std::optional<std::string> query_text;
    static SearchOptions make(size_t top_k, size_t complexity) { SearchOptions o; o.top_k = top_k; o.complexity = complexity; return o; }
    SearchOptions& with_filter(std::string f) { filter = std::move(f); return *this; }
    SearchOptions& with_hybrid(std::string text, float alpha) { hybrid = true; hybrid_alpha = alpha; query_text = std::move(text); return *this; }
};
struct SearchHit { std::string id; uint64_t ordinal; float score; };

class IndexSearcher {
public:
    static IndexSearcher load(const std::string& index_path, const std::string& backend_name, size_t dimensions, int device = 0) {
        leann_cuda_searcher* s = nullptr;
        char err[1024];
        check(leann_cuda_searcher_load(index_path.c_str(), backend_name.c_str(), dimensions, device, &s, err, sizeof err), err);
        return IndexSearcher(s);
    }
    std::vector<SearchHit> search_with_options(const std::vector<float>& query_embedding, const SearchOptions& o) const {
        std::vector<uint64_t> idx(o.top_k);
        std::vector<float> sc(o.top_k);
        uint32_t cnt = 0;
        const char* t = o.query_text ? o.query_text->data() : nullptr;
        size_t tn = o.query_text ? o.query_text->size() : 0;
        char err[1024];
        check(leann_cuda_searcher_search(s_.get(), query_embedding.data(), t ? &t : nullptr, t ? &tn : nullptr, 1, o.top_k, o.complexity,
                                         o.filter ? o.filter->c_str() : nullptr, o.hybrid ? 1 : 0, o.hybrid_alpha, idx.data(), sc.data(),
                                         &cnt, err, sizeof err), err);
        std::vector<SearchHit> out;
        for (uint32_t i = 0; i < cnt; ++i) {
            char id[512];
            leann_cuda_searcher_id(s_.get(), idx[i], id, sizeof id);
            out.push_back({id, idx[i], sc[i]});
        }
        return out;
    }
    std::vector<SearchHit> search(const std::vector<float>& q, size_t top_k, size_t complexity) const {
        return search_with_options(q, SearchOptions::make(top_k, complexity));
    }
    size_t len() const { return leann_cuda_searcher_len(s_.get()); }
    /// searcher.rs:228-246: BM25-only top-k ordinals and scores (the reference returns the passages' texts).
    std::vector<std::pair<size_t, float>> bm25_search(const std::string& query, size_t top_k) const {
        std::vector<uint64_t> idx(top_k);
        std::vector<float> sc(top_k);
        uint32_t cnt = 0;
        char err[1024];
        check(leann_cuda_searcher_bm25_search(s_.get(), query.data(), query.size(), top_k, idx.data(), sc.data(), &cnt, err, sizeof err), err);
        std::vector<std::pair<size_t, float>> out;
        for (uint32_t i = 0; i < cnt; ++i) out.emplace_back((size_t)idx[i], sc[i]);
        return out;
    }
private:
    explicit IndexSearcher(leann_cuda_searcher* s) : s_(s, leann_cuda_searcher_close) {}
    std::shared_ptr<leann_cuda_searcher> s_;
};

}  // namespace leann
