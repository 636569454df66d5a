/*
 * leann_cuda.h — C ABI of the B200-native (sm_100a) search path for decisiongraph/leann-rs.
 *
 * Every entry point names the reference interface it replaces (paths relative to the leann-rs
 * repository). Plain pointers and sizes only; no C++/torch types cross this boundary. All functions
 * return LEANN_OK (0) or a negative error class and copy a message into `err` (nullable).
 * Nothing here aborts or throws across the ABI. There is NO CPU fallback: every search entry point
 * returns LEANN_ERR_CUDA when no sm_100 device is usable.
 *
 * Threading: `*_search*` calls may be issued concurrently on one handle from any number of threads
 * (src/cli/serve.rs:84,289 shares one searcher behind read locks); the library serialises the launches of a
 * handle and merges concurrent single-query calls (see leann_cuda_set_coalescing). open/build/close must not
 * race with searches on the same handle.
 */
#ifndef LEANN_CUDA_H
#define LEANN_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEANN_OK 0
#define LEANN_ERR_NOT_FOUND (-1)    /* hnsw.rs:34-40, diskann.rs:26-32 "index not found" */
#define LEANN_ERR_BAD_FORMAT (-2)   /* hnsw.rs:55-70 "incompatible format" */
#define LEANN_ERR_DIM_MISMATCH (-3) /* usearch load / diskann assert on dimension */
#define LEANN_ERR_CUDA (-4)
#define LEANN_ERR_NCCL (-5)
#define LEANN_ERR_INVALID_ARG (-6)
#define LEANN_ERR_FAISS_FORMAT (-7) /* backend/compat.rs:15-38 + hnsw.rs:24-32 */
#define LEANN_ERR_PARSE (-8)        /* MetadataFilter::parse returned None (filter.rs:52) */

/* backend/mod.rs:16-19 `enum BackendType { Hnsw, DiskAnn }` + the exact scan of index/recompute.rs */
#define LEANN_BACKEND_HNSW 0
#define LEANN_BACKEND_VAMANA 1
#define LEANN_BACKEND_FLAT 2

/* Distance / score conventions.
 *  IP        1 - <q,x>, ascending     usearch MetricKind::IP            (hnsw.rs:45)
 *  L2SQ      |q-x|^2,   ascending     BASELINE config C4
 *  IP_CLAMP  max(0, 1 - <q,x>)        anndists DistDot                  (diskann.rs:16,36); deviation: anndists
 *                                     asserts 1 - <q,x> >= -2e-6 and panics on non-unit vectors, this library clamps
 *  DOT_DESC  <q,x>,     descending    RecomputeSearcher score           (recompute.rs:96-106)
 * LEANN_METRIC_DEFAULT takes the metric recorded in the file / the reference default of the backend. */
#define LEANN_METRIC_DEFAULT (-1)
#define LEANN_METRIC_IP 0
#define LEANN_METRIC_L2SQ 1
#define LEANN_METRIC_IP_CLAMP 2
#define LEANN_METRIC_DOT_DESC 3

#define LEANN_MASK_NONE 0
#define LEANN_MASK_INLINE 1 /* bit i set = passage i may be returned; tested inside the traversal */

typedef struct leann_cuda_index leann_cuda_index;
typedef struct leann_cuda_shards leann_cuda_shards;
typedef struct leann_cuda_bm25 leann_cuda_bm25;
typedef struct leann_cuda_filter leann_cuda_filter;
typedef struct leann_cuda_metacols leann_cuda_metacols;
typedef struct leann_cuda_searcher leann_cuda_searcher;

/* ---------------------------------------------------------------------------------------------
 * Backend: replaces BackendType::load_searcher (backend/mod.rs:23-45), HnswSearcher::load
 * (backend/hnsw.rs:18-75), DiskAnnSearcher::load (backend/diskann.rs:21-43).
 * `base_path` is the extension-less base (".../documents.leann"); ".index" / ".diskann" /
 * ".embeddings" is appended exactly as `with_extension` does in the reference.
 * ------------------------------------------------------------------------------------------- */
int leann_cuda_open(const char* base_path, int backend, size_t dims, int metric, int device,
                    leann_cuda_index** out, char* err, size_t errlen);

/* In-memory constructors (same device layout as open): host row-major f32 vectors. */
int leann_cuda_flat_from_host(const float* vectors, size_t n, size_t dims, int metric, int device,
                              leann_cuda_index** out, char* err, size_t errlen);
int leann_cuda_flat_from_device(const float* d_vectors, size_t n, size_t dims, int metric, int device,
                                leann_cuda_index** out, char* err, size_t errlen);

/* hnsw::build_index (backend/hnsw.rs:96-139): builds on the GPU; keys are ordinals 0..n-1.
 * `vectors_on_device` != 0: `vectors` is a device pointer on `device`. */
int leann_cuda_hnsw_build(const float* vectors, int vectors_on_device, size_t n, size_t dims,
                          size_t graph_degree, size_t complexity, int metric, uint64_t seed, int device,
                          leann_cuda_index** out, char* err, size_t errlen);
/* hnsw::add_to_index (backend/hnsw.rs:142-191, called from cli/update.rs:221-232): appends `m` vectors to a
 * GPU-resident HNSW index with keys start_id .. start_id + m - 1, as `index.add(id, embedding)` after
 * `index.load`. Connectivity is the index's; `complexity` is expansion_add (the reference passes 64).
 * Follow with leann_cuda_save to write the updated `.index`. Must not run concurrently with a search on
 * the same handle. */
int leann_cuda_hnsw_add(leann_cuda_index* index, const float* vectors, int vectors_on_device, size_t m,
                        uint64_t start_id, size_t complexity, uint64_t seed, char* err, size_t errlen);
/* diskann::build_index (backend/diskann.rs:70-105), alpha as DiskAnnParams.alpha (:91). */
int leann_cuda_vamana_build(const float* vectors, int vectors_on_device, size_t n, size_t dims,
                            size_t graph_degree, size_t complexity, float alpha, int metric,
                            uint64_t seed, int device, leann_cuda_index** out, char* err, size_t errlen);
/* index.save(<base>.index) (hnsw.rs:134) / build_index_with_params file output (diskann.rs:94-99) /
 * EmbeddingsWriter (index/embeddings.rs:126-147). Writes the reference's on-disk formats. */
int leann_cuda_save(const leann_cuda_index* index, const char* base_path, char* err, size_t errlen);

/* Host-only validation of an index file: the header pass and node parser of leann_cuda_open, without touching a device
 * (same error classes and messages). info[0..7] = n, dims, M (or R), M0, max_level, entry (or medoid), number of
 * upper-level lists, FNV-1a over every neighbour list in list order (terminated per list) and every key. */
int leann_cuda_check_index_file(const char* base_path, int backend, size_t dims, uint64_t* info8, char* err, size_t errlen);

/* Cached device layout. leann_cuda_open streams the vectors block of `.index` / `.diskann` /
 * `.embeddings` files from disk to HBM through pinned staging buffers and parses the usearch node block on host threads
 * meanwhile. write_layout_cache stores the parsed, fixed-stride adjacency of an HNSW handle as `<base>.cuda-layout`, bound to
 * the `.index` file at the same base by (size, mtime, hash of its headers); a later leann_cuda_open finds it, verifies the
 * binding and streams it instead of reading and parsing the node block (a stale or foreign cache file is ignored).
 * The environment variable LEANN_CUDA_LAYOUT_CACHE=1 makes leann_cuda_open write the file after a parse;
 * LEANN_CUDA_NO_LAYOUT_CACHE=1 makes it ignore an existing one. */
int leann_cuda_write_layout_cache(const leann_cuda_index* index, const char* base_path, char* err, size_t errlen);
/* 1 when this handle's adjacency was loaded from `<base>.cuda-layout`. */
int leann_cuda_layout_cache_used(const leann_cuda_index* index);

/* BackendSearcher::len (backend/traits.rs:24). */
size_t leann_cuda_len(const leann_cuda_index* index);
size_t leann_cuda_dims(const leann_cuda_index* index);
/* info[0..7] = n, dims, backend, metric, M (or R), M0, max_level, entry (or medoid). */
int leann_cuda_info(const leann_cuda_index* index, uint64_t* info8);
/* Number of lanes that cooperate on one distance for dimension `dims` (the oracle mirrors this
 * reduction order bit for bit in tests). */
int leann_cuda_reduction_lanes(size_t dims);
/* Capacity of the bounded candidate queue the traversal uses for expansion `ef`
 * (ef without a mask; min(4*ef, 2048) with an inline mask). Exposed so tests can drive the oracle's
 * bounded-queue mode with the same number. */
size_t leann_cuda_queue_capacity(size_t ef, int masked);

/* BackendSearcher::search (backend/traits.rs:16-21; hnsw.rs:79-88; diskann.rs:47-62), batched.
 * queries: nq x dims row-major f32 (host). keys/dists: nq x k, ascending distance (descending score
 * for DOT_DESC); unused tail = UINT64_MAX / +inf (-inf for DOT_DESC); counts[i] = valid entries.
 * ef: HNSW expansion_search / Vamana beam; effective value is max(ef, k) as in both reference
 * engines. (HnswSearcher ignores `complexity` and always runs ef = 64, hnsw.rs:49,83 — callers that
 * want that behaviour pass ef = 64.)
 * mask_bits: nullable, ceil(n/64) words, bit s%64 of word s/64 belongs to slot s. */
int leann_cuda_search(const leann_cuda_index* index, const float* queries, size_t nq, size_t k,
                      size_t ef, const uint64_t* mask_bits, int mask_mode, uint64_t* keys,
                      float* dists, uint32_t* counts, char* err, size_t errlen);

/* Request coalescing. The reference API is one query per call and `leann serve` issues
 * those calls concurrently on one shared searcher (src/cli/serve.rs:84,260-311). Calls with nq == 1 (no mask)
 * that arrive while another launch of the same handle is running are merged into one batched launch; each
 * caller still gets exactly its own result. Default: max_batch = 256, max_wait_us = 0 — a lone caller never
 * waits, concurrent callers batch naturally. max_wait_us > 0 makes a leader wait that long for company;
 * max_batch <= 1 disables coalescing (calls then serialise on the handle). */
/* Visited set of the graph traversal. By default every resident warp owns a byte map over the n nodes (exact, never cleared);
 * when n * resident warps would not fit in a third of device memory (tens of millions of short vectors per GPU) the library
 * switches to per-warp hash tables of about 2 * ef * degree node ids with a small shared pool of byte maps for traversals
 * that outgrow theirs. Short rows (d <= 128) under the diskann-rs stop rule keep (the first level of) their table in shared
 * memory (up to 2^24 rows per GPU). capacity 0 = automatic, 1 = byte maps only (fewer warps on large indexes), 2 / 3 =
 * stand-alone shared-memory tables wherever the kernel supports them (3: 256-entry limit), 4 / 5 = shared-memory first level
 * + hash-table overflow level (5: tiny levels) — 3 and 5 exist for tests, every traversal then overflows and spills —,
 * >= 1024 = force hash tables of this size (tests use a tiny table to exercise the spill). 6..1023 are invalid.
 * Results never depend on the choice. */
int leann_cuda_set_visited_hash(leann_cuda_index* index, size_t capacity);
int leann_cuda_set_coalescing(leann_cuda_index* index, size_t max_batch, unsigned max_wait_us);
int leann_cuda_coalescing_stats(const leann_cuda_index* index, uint64_t* batches, uint64_t* requests);
/* stats[0] = (re)allocations of the traversal workspace so far, [1] = its bytes, [2] = 1 in large-index mode,
 * [3] = warps it is sized for. A repeated identical search must not change stats[0]. */
int leann_cuda_workspace_stats(const leann_cuda_index* index, uint64_t* stats4);

/* Same, all pointers in device memory of the index's device, enqueued on `cuda_stream`
 * (a cudaStream_t; NULL = default stream). Graph backends return without host synchronisation; the exact scan
 * (FLAT) synchronises the stream once per call to read its overflow flag. Launches on one handle share one
 * workspace: the library orders them itself (each call's stream first waits for the previous call's work on
 * this handle), so calls from different streams or threads are safe but do not overlap each other.
 * d_stats: nullable, nq x 4 u64 = (distance evaluations, level-0 hops, upper-level hops, queue drops). */
int leann_cuda_search_device(const leann_cuda_index* index, const float* d_queries, size_t nq,
                             size_t k, size_t ef, const uint64_t* d_mask_bits, int mask_mode,
                             uint64_t* d_keys, float* d_dists, uint32_t* d_counts, uint64_t* d_stats,
                             void* cuda_stream, char* err, size_t errlen);

/* Per-query top-k merge of `n_shards` result sets laid out [shard][nq][k] (the buffer an
 * ncclAllGather of each shard's keys/dists produces). descending != 0 for DOT_DESC. Device pointers. */
int leann_cuda_topk_merge_device(const uint64_t* d_keys_in, const float* d_dists_in, size_t n_shards,
                                 size_t nq, size_t k, int descending, uint64_t* d_keys_out,
                                 float* d_dists_out, uint32_t* d_counts_out, void* cuda_stream,
                                 char* err, size_t errlen);

void leann_cuda_close(leann_cuda_index* index);

/* ---------------------------------------------------------------------------------------------
 * Sharded backend (BASELINE north_star): the database and graph are split into sub-indexes, one per
 * GPU; every shard searches all queries, the per-shard top-k lists are exchanged and merged per query (K4).
 * What BackendType::load_searcher (backend/mod.rs:23-45) returns for an index that spans several GPUs;
 * `search` keeps the contract of BackendSearcher::search (backend/traits.rs:16-21). Returned keys are global:
 * shard key + key_offsets[shard] (key_offsets == NULL: running sum of the shard lengths, i.e. the shards are
 * consecutive row ranges of one corpus). At most 16 shards.
 *
 * One host process, several devices: leann_cuda_shards_open / leann_cuda_shards_from_indexes. `exchange`:
 *   0 automatic (peer memory when every pair of devices can map each other: the merge kernel on the first device
 *     reads the shards' result blocks directly over NVLink, no collective launch; otherwise NCCL),
 *   1 ncclCommInitAll + ncclAllGather + merge,   2 peer memory or fail.
 * One process per GPU (torchrun-style): rank 0 calls leann_cuda_comm_unique_id, the caller broadcasts the 128 bytes
 * with its own plumbing, every rank calls leann_cuda_shards_join with its local index (ncclCommInitRank inside);
 * every rank then issues the same searches and every rank receives the merged answer.
 * libnccl.so.2 is loaded with dlopen on first use (LEANN_CUDA_NCCL_LIB overrides the path); failures are
 * LEANN_ERR_NCCL.
 * ------------------------------------------------------------------------------------------- */
int leann_cuda_shards_open(const char* const* base_paths, size_t n_shards, int backend, size_t dims, int metric,
                           const int* devices, const uint64_t* key_offsets, int exchange, leann_cuda_shards** out,
                           char* err, size_t errlen);
/* Same over already opened / built handles (one per device). take_ownership != 0: shards_close closes them. */
int leann_cuda_shards_from_indexes(leann_cuda_index* const* shards, size_t n_shards, const uint64_t* key_offsets,
                                   int take_ownership, int exchange, leann_cuda_shards** out, char* err, size_t errlen);
int leann_cuda_comm_unique_id(unsigned char* id, size_t id_bytes /* >= 128 */, char* err, size_t errlen);
int leann_cuda_shards_join(leann_cuda_index* local, int take_ownership, const unsigned char* id, size_t id_bytes,
                           int rank, int n_ranks, uint64_t key_offset, leann_cuda_shards** out, char* err, size_t errlen);
/* BackendSearcher::len over all shards / number of shards. */
size_t leann_cuda_shards_len(const leann_cuda_shards* shards);
size_t leann_cuda_shards_count(const leann_cuda_shards* shards);
size_t leann_cuda_shards_dims(const leann_cuda_shards* shards);
/* info[0] = shards, [1] = exchange in use (0 none, 1 NCCL all_gather, 2 peer-memory merge),
 * [2] = exchange steps so far, [3] = bytes moved between GPUs by them. */
int leann_cuda_shards_info(const leann_cuda_shards* shards, uint64_t* info4);
/* Host buffers, as leann_cuda_search. shard_masks: nullable; one entry per LOCAL shard (nullable each), bits indexed
 * by the shard's own slots, tested inside the traversal (LEANN_MASK_INLINE). */
int leann_cuda_shards_search(leann_cuda_shards* shards, const float* queries, size_t nq, size_t k, size_t ef,
                             const uint64_t* const* shard_masks, uint64_t* keys, float* dists, uint32_t* counts,
                             char* err, size_t errlen);
/* Device buffers on the local shard's device, everything (search, all_gather, merge) enqueued on `cuda_stream`;
 * handles with exactly one local shard (the process-per-GPU layout). */
int leann_cuda_shards_search_device(leann_cuda_shards* shards, const float* d_queries, size_t nq, size_t k, size_t ef,
                                    const uint64_t* d_mask_bits, uint64_t* d_keys, float* d_dists, uint32_t* d_counts,
                                    void* cuda_stream, char* err, size_t errlen);
/* IndexSearcher::search_with_options (index/searcher.rs:123-210) over document-range shards, one process per GPU
 * (a handle from leann_cuda_shards_join): vector candidates from the sharded backend, BM25 from this rank's postings
 * (leann_cuda_bm25_build_sharded, documents = the rows of this rank's vector shard), ONE ncclAllGather of the per-rank
 * block (BM25 top list, scores of the candidates the shard owns, max / min of its dense score vector), then merge,
 * fusion and post-filter walk on the device. Every rank issues the same call and receives the same answer, which is
 * bit-identical to leann_cuda_hybrid_search over the unsharded index. filter_mask: global passage bits (nullable). */
int leann_cuda_shards_hybrid_search(leann_cuda_shards* shards, const leann_cuda_bm25* bm25_shard, const float* queries,
                                    const char* const* query_texts, const size_t* query_text_bytes, size_t nq, size_t top_k,
                                    size_t ef, int hybrid, float alpha, const uint64_t* filter_mask, size_t mask_bits,
                                    uint64_t* idx, float* scores, uint32_t* counts, char* err, size_t errlen);
void leann_cuda_shards_close(leann_cuda_shards* shards);

/* ---------------------------------------------------------------------------------------------
 * BM25: replaces Bm25Scorer::{build,score_query,search} (index/bm25.rs:33-122), tokenize (:127-132)
 * and hybrid_rerank (:135-170). The inverted index is built once and kept in HBM.
 * ------------------------------------------------------------------------------------------- */
int leann_cuda_bm25_build(const char* const* docs, const size_t* doc_bytes, size_t n_docs, int device,
                          leann_cuda_bm25** out, char* err, size_t errlen);
size_t leann_cuda_bm25_len(const leann_cuda_bm25* bm25);
/* stats[0..3] = n_docs, n_terms, n_postings, total_tokens; avg_doc_len out (bm25.rs:61-65). */
int leann_cuda_bm25_stats(const leann_cuda_bm25* bm25, uint64_t* stats4, float* avg_doc_len);
/* Number of terms the handle also keeps as dense score rows (K3d: terms present in at least a quarter of the documents are
 * added to the accumulator tile row-wise, 4 B per document instead of 8 B per posting; results are bit-identical).
 * Tuning, read at build time: LEANN_CUDA_BM25_DENSE_FRAC (df / n_docs threshold, default 0.25, 0 = off),
 * LEANN_CUDA_BM25_DENSE_MAX (row limit, default 64). */
size_t leann_cuda_bm25_dense_rows(const leann_cuda_bm25* bm25);
/* tokenize (bm25.rs:127-132): writes tokens separated by '\n' into out (truncated to cap),
 * returns the number of tokens. */
size_t leann_cuda_tokenize(const char* text, size_t text_bytes, char* out, size_t cap);
/* Bm25Scorer::score_query (bm25.rs:77-106): dense scores[n_docs] (host). */
int leann_cuda_bm25_score(const leann_cuda_bm25* bm25, const char* query, size_t query_bytes,
                          float* scores, char* err, size_t errlen);
/* Bm25Scorer::search (bm25.rs:109-122), batched: idx/scores nq x top_k (host), score > 0 only,
 * stable descending (ties by ascending document index). */
int leann_cuda_bm25_search(const leann_cuda_bm25* bm25, const char* const* queries,
                           const size_t* query_bytes, size_t nq, size_t top_k, uint64_t* idx,
                           float* scores, uint32_t* counts, char* err, size_t errlen);
/* Measurement of the last leann_cuda_bm25_search batch on this handle: postings covered by its query tokens
 * (algorithmic bytes of K3 = postings * 8) and the device time of the query kernel (CUDA events). */
int leann_cuda_bm25_last_batch(const leann_cuda_bm25* bm25, uint64_t* postings, float* kernel_ms);
/* Bytes the tokens of that batch make the query kernel stream: 8 per posting, or 4 per document for a token whose term has a
 * dense row (the roofline numerator of K3 once dense rows exist). */
uint64_t leann_cuda_bm25_last_batch_bytes(const leann_cuda_bm25* bm25);
/* hybrid_rerank (bm25.rs:135-170) for one candidate list, evaluated on the device. bm25_scores is
 * the dense host vector of n_docs scores. Output has n entries, stable descending. */
int leann_cuda_hybrid_rerank(const uint64_t* idx, const float* vec_scores, size_t n,
                             const float* bm25_scores, size_t n_docs, float alpha, int device,
                             uint64_t* out_idx, float* out_scores, char* err, size_t errlen);
void leann_cuda_bm25_free(leann_cuda_bm25* bm25);

/* IndexSearcher::search_with_options (index/searcher.rs:123-210) without the passage fetch:
 * fetch_k = 5k when a filter or hybrid is present (:129-133); backend search (:136); BM25 union with
 * vector score 0.0 (:156-165); hybrid_rerank (:167); post-filter walk until k results (:174-207).
 * query_texts nullable (no hybrid). filter_mask nullable: bit i = passage i passes
 * MetadataFilter::matches (post-filter, reference semantics; may return < k).
 * ef: the backend `complexity` (pass 64 for HNSW to reproduce hnsw.rs:49,83).
 * Outputs nq x top_k (host): passage ordinals and the score searcher.rs returns (distance, or fused
 * score when hybrid). */
/* Document-range shards of the BM25 / hybrid path. Every shard indexes its own documents with the
 * corpus-wide N, total token count and per-term document frequency, so idf (bm25.rs:88), avg_doc_len (:61-65) and
 * each posting's score are bit-identical to the unsharded index. The statistics travel between ranks as opaque
 * blobs (all_gather of bytes): shard_stats on every rank -> stats_merge of all blobs -> build_sharded.
 * Sizing call: out == NULL returns the byte count in *needed (shard_stats tokenises on `device` during the sizing call
 * and hands the same blob to the fill call that follows on the calling thread). */
int leann_cuda_bm25_shard_stats(const char* const* docs, const size_t* doc_bytes, size_t n_docs, int device,
                                unsigned char* out, size_t cap, size_t* needed, char* err, size_t errlen);
int leann_cuda_bm25_stats_merge(const unsigned char* const* blobs, const size_t* blob_bytes, size_t n_blobs,
                                unsigned char* out, size_t cap, size_t* needed, char* err, size_t errlen);
int leann_cuda_bm25_build_sharded(const char* const* docs, const size_t* doc_bytes, size_t n_docs,
                                  const unsigned char* global_stats, size_t stats_bytes, int device,
                                  leann_cuda_bm25** out, char* err, size_t errlen);
/* One shard's part of search_with_options' hybrid step (searcher.rs:146-167) for a batch: BM25 top-k of the shard
 * (ids + doc_offset), BM25 score of the global vector candidates the shard owns (0 elsewhere), min / max of the
 * shard's dense score vector (bm25.rs:152-153). Reduce over shards: top lists -> leann_cuda_topk_merge_device
 * (descending), cand_bm -> sum, bmax -> max, bmin -> min; then leann_cuda_hybrid_fuse. cand_idx may be NULL. */
int leann_cuda_bm25_search_shard(const leann_cuda_bm25* bm25, const char* const* queries, const size_t* query_bytes,
                                 size_t nq, size_t top_k, uint64_t doc_offset, const uint64_t* cand_idx,
                                 const uint32_t* cand_cnt, size_t fk, uint64_t* top_idx, float* top_score,
                                 uint32_t* top_cnt, float* cand_bm, float* bmax, float* bmin, char* err, size_t errlen);
/* Batched hybrid_rerank (bm25.rs:135-170) + BM25-only additions (searcher.rs:156-165) + post-filter walk
 * (searcher.rs:174-207) over gathered inputs; host pointers. vkeys/vdists: nq x fk vector results; bm_*: nq x bm_k. */
int leann_cuda_hybrid_fuse(const uint64_t* vkeys, const float* vdists, const uint32_t* vcnt, size_t nq, size_t fk,
                           const float* cand_bm, const uint64_t* bm_idx, const float* bm_score, const uint32_t* bm_cnt,
                           size_t bm_k, const float* bmax, const float* bmin, int hybrid, float alpha,
                           const uint64_t* mask, size_t mask_bits, size_t top_k, int device, uint64_t* out_idx,
                           float* out_score, uint32_t* out_cnt, char* err, size_t errlen);
int leann_cuda_hybrid_search(const leann_cuda_index* index, const leann_cuda_bm25* bm25,
                             const float* queries, const char* const* query_texts,
                             const size_t* query_text_bytes, size_t nq, size_t top_k, size_t ef,
                             int hybrid, float alpha, const uint64_t* filter_mask, uint64_t* idx,
                             float* scores, uint32_t* counts, char* err, size_t errlen);

/* ---------------------------------------------------------------------------------------------
 * Metadata filter: replaces MetadataFilter::parse / matches (index/filter.rs:52-134,137-316,319-439).
 * Evaluated once on the host into a bitmask; kernels test bits.
 * ------------------------------------------------------------------------------------------- */
int leann_cuda_filter_parse(const char* expr, leann_cuda_filter** out, char* err, size_t errlen);
/* JSON rendering of the parsed tree in serde's shape ({"field","op","value"} / {"and":[..]} / {"or":[..]}). */
size_t leann_cuda_filter_describe(const leann_cuda_filter* f, char* out, size_t cap);
/* matches(&serde_json::Value): metadata given as JSON text. *result = 0/1. */
int leann_cuda_filter_matches(const leann_cuda_filter* f, const char* metadata_json, size_t bytes,
                              int* result, char* err, size_t errlen);
/* Bulk: one JSON document per passage -> bitmask words (ceil(n/64)). */
int leann_cuda_filter_mask(const leann_cuda_filter* f, const char* const* metadata_json,
                           const size_t* bytes, size_t n, uint64_t* mask_bits, char* err, size_t errlen);
void leann_cuda_filter_free(leann_cuda_filter* f);

/* Columnar side-car of the passages' metadata. The reference opens the passage file and parses
 * its JSON for every candidate a filter is tested on (index/passages.rs:90-105 + filter.rs:319-439); here each
 * dotted field path becomes one typed column (kind / f64 / dictionary-coded string) built once, and a parsed
 * filter is evaluated column-wise into the N-bit mask the kernels test. metadata_json[i] == NULL: passage i has
 * no metadata. Results are identical to leann_cuda_filter_mask row by row. Host only. */
int leann_cuda_metacols_build(const char* const* metadata_json, const size_t* bytes, size_t n,
                              leann_cuda_metacols** out, char* err, size_t errlen);
int leann_cuda_metacols_mask(const leann_cuda_metacols* cols, const leann_cuda_filter* f, uint64_t* mask_bits,
                             char* err, size_t errlen);
size_t leann_cuda_metacols_len(const leann_cuda_metacols* cols);
size_t leann_cuda_metacols_fields(const leann_cuda_metacols* cols);
void leann_cuda_metacols_free(leann_cuda_metacols* cols);

/* ---------------------------------------------------------------------------------------------
 * IndexSearcher (index/searcher.rs:76-109 load; :123-210 search; :228-246 bm25_search):
 * opens <base>.passages.jsonl / .passages.idx.json / .ids.txt + the backend; BM25 is built once
 * on first hybrid use (the reference rebuilds it per query, searcher.rs:149-150).
 * backend_name: "hnsw" | "diskann" (index/meta.rs backend_name) | "flat" (pruned index, recompute.rs).
 * ------------------------------------------------------------------------------------------- */
int leann_cuda_searcher_load(const char* base_path, const char* backend_name, size_t dims, int device,
                             leann_cuda_searcher** out, char* err, size_t errlen);
size_t leann_cuda_searcher_len(const leann_cuda_searcher* s);
/* id_map[idx] or idx.to_string() when out of range (searcher.rs:180-184). Returns bytes written. */
size_t leann_cuda_searcher_id(const leann_cuda_searcher* s, uint64_t idx, char* out, size_t cap);
int leann_cuda_searcher_search(const leann_cuda_searcher* s, const float* queries,
                               const char* const* query_texts, const size_t* query_text_bytes,
                               size_t nq, size_t top_k, size_t complexity, const char* filter_expr,
                               int hybrid, float alpha, uint64_t* idx, float* scores,
                               uint32_t* counts, char* err, size_t errlen);
/* IndexSearcher::bm25_search (searcher.rs:228-246): BM25-only top-k passage ordinals + scores. */
int leann_cuda_searcher_bm25_search(const leann_cuda_searcher* s, const char* query, size_t query_bytes,
                                    size_t top_k, uint64_t* idx, float* scores, uint32_t* count,
                                    char* err, size_t errlen);
/* HnswSearcher discards `complexity` and always searches with ef = 64 (hnsw.rs:49,83); that is the
 * default here too. on != 0 makes the HNSW backend honour `complexity` as ef. */
int leann_cuda_searcher_set_honor_complexity(leann_cuda_searcher* s, int on);
void leann_cuda_searcher_close(leann_cuda_searcher* s);

/* Library / device probe: returns the CUDA device count usable by the library (0 = none). */
int leann_cuda_device_count(void);
const char* leann_cuda_version(void);
/* The recalled third-party behaviours the traversal kernels were compiled with (leann_rs_b200/csrc/compat.h):
 * bit 0 usearch stop rule is strict (cand.d > radius), bit 1 diskann-rs stop rule is strict, bit 2 a newcomer enters
 * `top` before equal distances, bit 3 equal distances leave the candidate queue in arrival order, bit 4 DistDot clamps
 * at zero (the anndists assert on 1 - dot < -2e-6 is NOT reproduced: non-unit vectors are clamped, not rejected). */
unsigned leann_cuda_compat_flags(void);

#ifdef __cplusplus
}
#endif
#endif /* LEANN_CUDA_H */
