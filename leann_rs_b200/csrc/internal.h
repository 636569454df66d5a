// internal.h — shared host-side declarations of libleann_cuda (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/leann_cuda.h"

namespace leann {

constexpr uint32_t SENT = 0xFFFFFFFFu;  // padding slot in every adjacency row
constexpr int MAX_DEG = 128;            // largest adjacency row the kernels stage in shared memory
constexpr int MAX_EF = 1024;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define LEANN_CUDA_CHECK(expr)                                                                      \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::leann::Error(LEANN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---- host-side images of the reference's files (formats.cpp) --------------------------------
struct HostHnsw {  // usearch `.index` (hnsw.rs:55,134), SURVEY.md Appendix A.1
    size_t n = 0, d = 0, M = 0, M0 = 0;
    int64_t max_level = 0;
    uint64_t entry = 0;
    int metric = LEANN_METRIC_IP;
    std::vector<float> vecs;       // n*d
    std::vector<int16_t> levels;   // n
    std::vector<uint64_t> keys;    // n
    std::vector<uint32_t> adj0;    // n*M0, SENT padded, list order preserved
    std::vector<uint32_t> upper_base;  // n: first upper list of the node (valid when level > 0)
    std::vector<uint32_t> adjU;    // n_upper_lists*M, SENT padded
};
struct HostVamana {  // diskann-rs `.diskann` (diskann.rs:34-37,94-99), Appendix A.3
    size_t n = 0, d = 0, R = 0;
    uint32_t medoid = 0;
    std::string distance_name;
    std::vector<float> vecs;
    std::vector<uint32_t> adj;  // n*R, SENT padded
};

// Header-only passes (formats.cpp): every header field verified, byte ranges of the blocks returned; the blocks themselves
// are streamed to the device by open_stream.cu.
struct UsearchPlan {
    size_t n = 0, d = 0, M = 0, M0 = 0;
    int64_t max_level = 0;
    uint64_t entry = 0;
    int metric = LEANN_METRIC_IP;
    size_t vec_off = 0, vec_bytes = 0, levels_off = 0, nodes_off = 0, file_size = 0;
    int64_t mtime_ns = 0;
    uint64_t head_hash = 0;   // FNV-1a of (matrix shape, dense head, graph header)
};
struct DiskannPlan {
    size_t n = 0, d = 0, R = 0;
    uint32_t medoid = 0;
    std::string distance_name;
    size_t vec_off = 0, adj_off = 0, file_size = 0;
};
UsearchPlan usearch_probe(const std::string& path, size_t dims);
void usearch_layout(const UsearchPlan& pl, const std::string& path, const int16_t* levels, std::vector<uint32_t>& upper_base,
                    std::vector<uint64_t>& node_off, size_t& n_upper);
void usearch_parse_nodes(const UsearchPlan& pl, const std::string& path, const unsigned char* nodes, const uint64_t* node_off,
                         const int16_t* levels, const uint32_t* upper_base, size_t i0, size_t i1, uint64_t* keys, uint32_t* adj0,
                         uint32_t* adjU);
DiskannPlan diskann_probe(const std::string& path, size_t dims);
uint64_t fnv1a64(const void* data, size_t bytes, uint64_t h);

bool is_faiss_index(const std::string& index_file);                       // backend/compat.rs:15-38
void read_usearch_index(const std::string& path, size_t dims, HostHnsw& out);   // throws Error
void write_usearch_index(const std::string& path, const HostHnsw& g);
void read_diskann(const std::string& path, size_t dims, HostVamana& out);
void write_diskann(const std::string& path, const HostVamana& g);
void read_embeddings(const std::string& path, size_t dims, std::vector<float>& out, size_t& n);  // index/embeddings.rs:21-36
void write_embeddings(const std::string& path, const float* v, size_t n, size_t dims);
std::string with_extension(const std::string& base, const std::string& ext);  // Path::with_extension

// ---- device-resident index --------------------------------------------------------------------
struct GraphView {  // passed by value to kernels
    const float4* vecs;          // [n][d4] rows padded with zeros to a multiple of 4 floats
    const uint32_t* adj0;        // [n][deg0]
    const uint32_t* upper_base;  // [n]
    const uint32_t* adjU;        // [n_upper_lists][degU]
    const uint64_t* keys;        // [n] or nullptr (key == slot)
    uint32_t n, d, d4, deg0, degU;
    int max_level;
    uint32_t entry;
    int metric;
};

struct SearchWorkspace {
    uint8_t* visited = nullptr;  // [n_warps][n_pad] epoch tags
    uint32_t* epochs = nullptr;  // [n_warps]
    uint32_t* counter = nullptr; // dynamic query scheduler
    // large-index mode: visited hash tables [n_warps][vhash_cap]; `visited` / `epochs` then hold a small pool of byte maps
    uint32_t* vhash = nullptr; size_t vhash_words = 0; uint32_t vhash_cap = 0;
    uint32_t* pool_locks = nullptr; uint32_t pool_slots = 0;
    bool large_mode = false;
    size_t mem_total = 0;        // device memory size (queried once)
    uint64_t reallocs = 0;       // (re)allocations of the traversal workspace so far (leann_cuda_workspace_stats)
    int n_warps = 0;
    int warp_cap = 0;            // memory-bounded maximum pool size (0 = not computed yet)
    size_t n_pad = 0;
    cudaStream_t stream = nullptr;  // private stream for host-pointer calls
    float* d_queries = nullptr; uint64_t* d_keys = nullptr; float* d_dists = nullptr; uint32_t* d_counts = nullptr;
    uint64_t* d_mask = nullptr;
    size_t cap_q = 0, cap_out = 0, cap_mask = 0, cap_counts = 0;
};

struct SearchParams {
    const float* queries;  // [nq][d]
    uint32_t nq, k, ef, next_cap, next_capp;
    const uint64_t* mask;  // nullable
    int nonstrict_term;    // 0: usearch (cand.d > radius), 1: diskann-rs (full && cand.d >= worst)
    uint64_t* out_keys; float* out_dists; uint32_t* out_counts; uint64_t* out_stats;
    uint8_t* visited; uint32_t* epochs; uint32_t* counter;
    size_t n_pad;
    int n_warps;
    uint32_t* vhash; uint32_t vhash_cap;              // nullable: per-warp visited hash tables (large-index mode)
    uint32_t vhash16, q_rem_bits, q_key_bits, q_inv;  // vhash16 != 0: the tables hold vhash_cap 16-bit quotiented entries in buckets of 8
    uint32_t* pool_locks; uint32_t pool_slots;        // byte-map spill pool of the large-index mode (maps in `visited`)
    int row_ring;                                     // != 0: rows of a hop through a shared-memory ring of bulk async copies (A/B form)
    int smem_vis; uint32_t smv_limit;                 // smem_vis != 0: register-list kernel with the visited table in shared memory (spill above smv_limit entries)
    int coop_warps;  // 4 or 8 warps per CTA of the small-batch kernel
    int coop_ctas;   // > 0: small batch, launch this many multi-warp CTAs (one query each at a time) instead of the warp pool
};

int reduction_lanes(size_t dims);
// Enqueues the K1/K1f beam search on `stream`. Throws Error.
void launch_graph_search(const GraphView& g, const SearchParams& p, cudaStream_t stream);
size_t graph_search_smem_per_warp(uint32_t ef, uint32_t next_capp);
int graph_search_max_warps(int device);
int graph_search_warps_per_sm(const GraphView& g, uint32_t ef, uint32_t next_capp, int nonstrict_term = 0, int smem_vis = 0, int row_ring = 0);
bool graph_search_uses_reg_lists(const GraphView& g, const SearchParams& p);   // register-list kernel (short rows): supports the q16 table

// Exact scan (K2 + K2r)
struct FlatView { const float4* vecs; uint32_t n, d, d4; int metric; };
struct TcIndexView;
// helper stream / events / pinned flag words a handle lends to the exact scan (two query halves on two streams)
struct ScanAux {
    cudaStream_t helper = nullptr;   // nullptr: never split the batch
    cudaEvent_t fork = nullptr, join = nullptr;
    uint32_t* h_flags = nullptr;     // pinned, 4 words
};
void launch_exact_scan(const FlatView& f, const float* d_queries, uint32_t nq, uint32_t k, const uint64_t* d_mask,
                       uint64_t* d_keys, float* d_dists, uint32_t* d_counts, void* scratch, size_t scratch_bytes,
                       cudaStream_t stream, const TcIndexView* tv, int sms, const ScanAux& aux);
size_t exact_scan_scratch_bytes(uint32_t d4, uint32_t nq, uint32_t k);

void launch_topk_merge(const uint64_t* keys_in, const float* dists_in, uint32_t n_shards, uint32_t nq, uint32_t k,
                       int descending, uint64_t* keys_out, float* dists_out, uint32_t* counts_out, cudaStream_t stream);

void launch_pad_rows(const float* src, float4* dst, size_t n, uint32_t d, uint32_t d4, cudaStream_t stream);

}  // namespace leann

namespace leann {
// Request coalescing: concurrent nq=1 callers of one handle (the axum handlers of
// src/cli/serve.rs:260-311 share one searcher) are merged into one batched launch by a leader/follower
// scheme; no background thread.
struct CoalesceReq {
    const float* query; size_t k, ef;
    uint64_t* keys; float* dists; uint32_t* count;
    int rc = 0; std::string err; bool done = false;
    bool taken = false;   // already part of a batch some leader is running
};
struct Coalescer {
    std::mutex m;
    std::condition_variable cv_leader, cv_done;
    bool leader_active = false;
    size_t max_batch = 256;      // <= 1 = disabled
    unsigned max_wait_us = 0;    // 0: never wait for company, only merge what queued up while the previous batch ran
    std::vector<CoalesceReq*> queue;
    uint64_t batches = 0, requests = 0;
};
}  // namespace leann

struct leann_cuda_index {
    int backend = LEANN_BACKEND_HNSW;
    int device = 0;
    int metric = LEANN_METRIC_IP;
    size_t n = 0, d = 0;
    uint32_t d4 = 0;
    // device memory
    float4* vecs = nullptr;
    uint32_t* adj0 = nullptr;
    uint32_t* upper_base = nullptr;
    uint32_t* adjU = nullptr;
    uint64_t* keys = nullptr;
    std::vector<int16_t> h_levels;  // host copy kept for save()
    std::string distance_name;      // diskann metadata string
    size_t n_upper_lists = 0;
    uint32_t M = 0, M0 = 0;
    int max_level = 0;
    uint32_t entry = 0;
    bool identity_keys = true;
    bool layout_cache_used = false;   // adjacency came from <base>.cuda-layout (open_stream.cu)
    // workspaces (guarded by mu: concurrent callers serialise on the GPU queue)
    mutable std::mutex mu;
    mutable leann::SearchWorkspace ws;
    mutable void* scan_scratch = nullptr;
    mutable size_t scan_scratch_bytes = 0;
    mutable uint32_t* scan_pinned = nullptr;   // pinned words for the overflow read-back (2 per query half)
    mutable leann::ScanAux scan_aux;           // helper stream + events of the two-half exact scan
    // tensor-path copy of a flat database (bf16 rows, row norms, max norm), built on first use
    mutable void* tc_bf16 = nullptr;
    mutable float* tc_norms = nullptr;
    mutable uint32_t* tc_xmax = nullptr;
    mutable bool tc_disabled = false;
    mutable leann::Coalescer coalescer;
    // stream chaining: event recorded after the last enqueue on this handle, and the stream it was recorded on
    mutable cudaEvent_t chain_ev = nullptr;
    mutable cudaStream_t chain_stream = nullptr;
    mutable bool chain_valid = false;
    bool coop_small_batches = true;   // CTA-per-query kernel for nq <= 2 per SM
    size_t vhash_mode = 0;   // visited set: 0 auto (byte maps unless they would not fit), 1 byte maps only, >= 1024 force hash tables of this capacity
    leann::GraphView view() const {
        leann::GraphView g;
        g.vecs = vecs; g.adj0 = adj0; g.upper_base = upper_base; g.adjU = adjU;
        g.keys = identity_keys ? nullptr : keys;
        g.n = (uint32_t)n; g.d = (uint32_t)d; g.d4 = d4; g.deg0 = M0; g.degU = M;
        g.max_level = max_level; g.entry = entry; g.metric = metric;
        return g;
    }
};
