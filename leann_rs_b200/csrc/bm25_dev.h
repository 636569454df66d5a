// bm25_dev.h — device view of the BM25 inverted index and kernel launchers (bm25.cu).
#pragma once
#include <vector>

#include "internal.h"
#include "text.h"

namespace leann {

struct Bm25Dev {
    uint32_t n_docs;
    const uint64_t* term_off;  // [n_terms + 1]
    const uint32_t* post_doc;  // [n_postings] ascending inside a term
    const float* post_score;   // [n_postings] per-posting BM25 contribution (query independent)
    // dense rows of the most frequent terms (K3d, bm25.cu): dense_of[term] = row or 0xFFFFFFFF; row r holds the term's
    // per-posting score at [r * n_pad + doc] and 0.0f for documents outside its posting list
    const uint32_t* dense_of = nullptr;   // [n_terms] or null
    const float* dense_rows = nullptr;    // [n_dense][n_pad]
    uint32_t n_pad = 0;                   // n_docs rounded up to whole accumulator tiles
};

void launch_bm25_dense(const Bm25Dev& b, const uint32_t* terms, const uint64_t* dfs, size_t n_tokens, float* d_scores,
                       cudaStream_t s);
size_t bm25_max_query_tokens();   // longest query (known tokens, duplicates included) the query kernel takes
// Builds the dense rows of `b` from its postings (terms present in at least a quarter of the documents, most frequent first,
// bounded by LEANN_CUDA_BM25_DENSE_MAX rows, default 64, and an eighth of the free memory). LEANN_CUDA_BM25_DENSE_FRAC (default
// 0.25) is the df / n_docs threshold; 0 disables the rows.
void bm25_build_dense_rows(leann_cuda_bm25* b);
int bm25_query_ctas_per_sm();   // resident CTAs per SM the query kernel is built for (persistent pool size)
void launch_bm25_query(const Bm25Dev& b, const uint64_t* qtok_off, const uint32_t* qtok_term, uint32_t nq, uint32_t K,
                       int n_ctas, const uint64_t* cand_idx, const uint32_t* cand_cnt, uint32_t fk,
                       float* cand_bm, uint64_t* top_idx, float* top_score, uint32_t* top_cnt, float* bmax, float* bmin,
                       uint32_t* qcounter, cudaStream_t s);
void launch_bm25_candidates(const Bm25Dev& b, const uint64_t* qtok_off, const uint32_t* qtok_term, uint32_t nq, const uint64_t* cand_idx,
                            const uint32_t* cand_cnt, uint32_t fk, float* cand_bm, cudaStream_t s);
void launch_hybrid_fuse(const uint64_t* vkeys, const float* vdists, const uint32_t* vcnt, uint32_t fk, const float* cand_bm,
                        const uint64_t* bm_idx, const float* bm_score, const uint32_t* bm_cnt, uint32_t bm_k,
                        const float* bmax, const float* bmin, int hybrid, float alpha, const uint64_t* mask, uint64_t mask_bits,
                        uint32_t top_k, uint64_t* out_idx, float* out_score, uint32_t* out_cnt, uint32_t nq, cudaStream_t s);
// document-range shards: candidate ids -> this shard's local ordinals (~0 elsewhere); sum / max / min over the shards' blocks
void launch_localize_candidates(const uint64_t* global_idx, uint64_t doc_offset, uint64_t n_local, size_t count, uint64_t* local_idx, cudaStream_t s);
void launch_shard_reduce(const unsigned char* gathered, size_t block_bytes, uint32_t g, uint32_t nq, uint32_t fk, float* cand_bm, float* bmax,
                         float* bmin, cudaStream_t s);
void launch_dense_minmax_gather(const float* dense, uint32_t n, const uint64_t* idx, uint32_t m, float* cand_bm, float* bmax,
                                float* bmin, uint32_t* scratch2, cudaStream_t s);

// GPU index construction (bm25_build.cu). stats_only: fill `stats` (N, token count, per-term df) and return.
void bm25_build_device(const char* const* docs, const size_t* doc_bytes, size_t n_docs, const Bm25GlobalStats* glob,
                       leann_cuda_bm25* out, Bm25GlobalStats* stats, bool stats_only);

// search entry shared between api.cu and text_api.cu: enqueue a backend search with device buffers
void backend_search_device(const leann_cuda_index* ix, const float* d_queries, size_t nq, size_t k, size_t ef,
                           const uint64_t* d_mask, uint64_t* d_keys, float* d_dists, uint32_t* d_counts, cudaStream_t stream);
cudaStream_t backend_stream(const leann_cuda_index* ix);

}  // namespace leann

struct leann_cuda_bm25 {
    int device = 0;
    leann::Bm25Host host;   // dictionary, term offsets, idf, counters (the postings live on the device only)
    uint64_t* d_term_off = nullptr;
    uint32_t* d_post_doc = nullptr;
    float* d_post_score = nullptr;
    uint32_t* d_dense_of = nullptr;   // see Bm25Dev
    float* d_dense_rows = nullptr;
    uint32_t n_dense = 0, n_pad = 0;
    std::vector<uint8_t> is_dense;    // host copy, per term (empty without dense rows)
    // per-handle workspace
    mutable std::mutex mu;
    mutable float* d_acc = nullptr;   // [n_docs] dense score vector of score_query, zero between calls
    mutable int n_ctas = 0;
    mutable uint32_t* d_qcounter = nullptr;
    mutable cudaStream_t stream = nullptr;
    // measurement of the last leann_cuda_bm25_search batch: postings its tokens cover, device time of the query kernel
    mutable cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    mutable cudaEvent_t ev_join = nullptr;   // BM25 top-k kernel done (joins the vector search's stream in the hybrid path)
    mutable uint64_t last_postings = 0;
    mutable uint64_t last_stream_bytes = 0;   // what those tokens make the kernel read: 8 B per posting, or 4 B per document of a dense row
    mutable float last_kernel_ms = 0.0f;
    leann::Bm25Dev view() const {
        return leann::Bm25Dev{(uint32_t)host.num_docs, d_term_off, d_post_doc, d_post_score, n_dense ? d_dense_of : nullptr, d_dense_rows, n_pad};
    }
};
