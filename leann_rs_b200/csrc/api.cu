// api.cu — the C ABI of the vector path (include/leann_cuda.h): open/build/save/search/merge/close.
// Host logic only; kernels live in graph_search.cu, exact_scan.cu, hnsw_build.cu.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <functional>
#include <memory>

#include "compat.h"
#include "scan_common.h"

using namespace leann;

namespace leann {
int guard_impl(char* err, size_t errlen, const std::function<void()>& f);
// builders (hnsw_build.cu / vamana_build.cu)
void gpu_hnsw_build(leann_cuda_index* ix, size_t M, size_t ef_add, uint64_t seed);
void gpu_vamana_build(leann_cuda_index* ix, size_t R, size_t L, float alpha, uint64_t seed);
void gpu_hnsw_add(leann_cuda_index* ix, const float4* new_rows, size_t m, uint64_t start_id, size_t ef_add, uint64_t seed);
// file -> HBM loaders (open_stream.cu)
leann_cuda_index* open_hnsw_streamed(const std::string& base, size_t dims, int device, int metric, bool* used_cache);
leann_cuda_index* open_vamana_streamed(const std::string& base, size_t dims, int device, int metric);
leann_cuda_index* open_flat_streamed(const std::string& base, size_t dims, int device, int metric);
void write_layout_cache(const leann_cuda_index* ix, const std::string& base);
}  // namespace leann

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        cudaError_t e = cudaSetDevice(dev);
        if (e != cudaSuccess) throw Error(LEANN_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

void require_gpu(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw Error(LEANN_ERR_CUDA, "no CUDA device available: libleann_cuda has no CPU fallback");
    }
    if (device < 0 || device >= n) throw Error(LEANN_ERR_INVALID_ARG, "device ordinal out of range");
    cudaDeviceProp p;
    LEANN_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) throw Error(LEANN_ERR_CUDA, std::string("device is sm_") + std::to_string(p.major * 10 + p.minor) + ", this library is built for sm_100a only");
}

template <typename T>
T* dalloc(size_t count) {
    T* p = nullptr;
    if (count == 0) count = 1;
    LEANN_CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
    return p;
}

void upload_vectors(leann_cuda_index* ix, const float* src, bool on_device) {
    ix->d4 = (uint32_t)((ix->d + 3) / 4);
    ix->vecs = dalloc<float4>(ix->n * ix->d4);
    if (ix->n == 0) return;
    if (on_device) {
        launch_pad_rows(src, ix->vecs, ix->n, (uint32_t)ix->d, ix->d4, nullptr);
    } else if (ix->d % 4 == 0) {
        LEANN_CUDA_CHECK(cudaMemcpy(ix->vecs, src, ix->n * ix->d * 4, cudaMemcpyHostToDevice));
    } else {
        float* tmp = dalloc<float>(ix->n * ix->d);
        LEANN_CUDA_CHECK(cudaMemcpy(tmp, src, ix->n * ix->d * 4, cudaMemcpyHostToDevice));
        launch_pad_rows(tmp, ix->vecs, ix->n, (uint32_t)ix->d, ix->d4, nullptr);
        LEANN_CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(tmp);
    }
    LEANN_CUDA_CHECK(cudaDeviceSynchronize());
}

template <typename T>
T* upload(const std::vector<T>& v) {
    T* p = dalloc<T>(v.size());
    if (!v.empty()) LEANN_CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

leann_cuda_index* from_hnsw(HostHnsw& h, int device, int metric) {
    std::unique_ptr<leann_cuda_index> ix(new leann_cuda_index());
    ix->backend = LEANN_BACKEND_HNSW; ix->device = device;
    ix->metric = metric == LEANN_METRIC_DEFAULT ? h.metric : metric;
    ix->n = h.n; ix->d = h.d; ix->M = (uint32_t)h.M; ix->M0 = (uint32_t)h.M0;
    ix->max_level = (int)h.max_level; ix->entry = (uint32_t)h.entry;
    ix->n_upper_lists = h.M ? h.adjU.size() / h.M : 0;
    ix->h_levels = h.levels;
    upload_vectors(ix.get(), h.vecs.data(), false);
    ix->adj0 = upload(h.adj0);
    ix->upper_base = upload(h.upper_base);
    ix->adjU = upload(h.adjU);
    ix->identity_keys = true;
    for (size_t i = 0; i < h.n; ++i) if (h.keys[i] != i) { ix->identity_keys = false; break; }
    ix->keys = upload(h.keys);
    return ix.release();
}

leann_cuda_index* from_vamana(HostVamana& h, int device, int metric) {
    std::unique_ptr<leann_cuda_index> ix(new leann_cuda_index());
    ix->backend = LEANN_BACKEND_VAMANA; ix->device = device;
    // diskann.rs:16,36 hard-wires DistDot
    ix->metric = metric == LEANN_METRIC_DEFAULT ? LEANN_METRIC_IP_CLAMP : metric;
    ix->n = h.n; ix->d = h.d; ix->M = (uint32_t)h.R; ix->M0 = (uint32_t)h.R;
    ix->max_level = 0; ix->entry = h.medoid; ix->distance_name = h.distance_name;
    upload_vectors(ix.get(), h.vecs.data(), false);
    ix->adj0 = upload(h.adj);
    ix->identity_keys = true;
    return ix.release();
}

uint32_t next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }

// Large-index mode of the traversal workspace: one byte map per resident warp would not fit (n * warps bytes), so every warp
// gets a visited hash table sized from ef * degree instead, and a small pool of byte maps serves the traversals that outgrow
// their table (graph_device.cuh VisitedSet).
void ensure_workspace_large(const leann_cuda_index* ix, int want, size_t n_pad, size_t budget, size_t ef) {
    SearchWorkspace& ws = ix->ws;
    uint32_t cap = ix->vhash_mode >= 1024 ? next_pow2((uint32_t)std::min<size_t>(ix->vhash_mode, (size_t)1 << 22))
                                         : std::max<uint32_t>(1024u, next_pow2((uint32_t)std::min<size_t>(2 * ef * ix->M0, (size_t)1 << 22)));
    if (!ws.large_mode || ws.n_pad != n_pad) {   // (re)create the byte-map pool
        ws.reallocs++;
        if (ws.visited) cudaFree(ws.visited);
        if (ws.epochs) cudaFree(ws.epochs);
        if (ws.pool_locks) cudaFree(ws.pool_locks);
        ws.visited = nullptr; ws.epochs = nullptr; ws.pool_locks = nullptr; ws.n_warps = 0;
        uint32_t slots = (uint32_t)std::min<size_t>(64, std::max<size_t>(1, budget / 2 / std::max<size_t>(n_pad, 1)));
        ws.visited = dalloc<uint8_t>((size_t)slots * n_pad);
        ws.epochs = dalloc<uint32_t>(slots);
        ws.pool_locks = dalloc<uint32_t>(slots);
        LEANN_CUDA_CHECK(cudaMemset(ws.visited, 0, (size_t)slots * n_pad));
        LEANN_CUDA_CHECK(cudaMemset(ws.epochs, 0, (size_t)slots * 4));
        LEANN_CUDA_CHECK(cudaMemset(ws.pool_locks, 0, (size_t)slots * 4));
        ws.pool_slots = slots;
        ws.n_pad = n_pad;
        ws.large_mode = true;
    }
    if (!ws.counter) ws.counter = dalloc<uint32_t>(1);
    ws.n_warps = std::max(ws.n_warps, want);
    const size_t words = (size_t)cap * (size_t)ws.n_warps;
    if (ws.vhash_words < words) {
        ws.reallocs++;
        if (ws.vhash) cudaFree(ws.vhash);
        ws.vhash = nullptr; ws.vhash_words = 0;
        ws.vhash = dalloc<uint32_t>(words);
        ws.vhash_words = words;
    }
    ws.vhash_cap = cap;
}

void ensure_stream(const leann_cuda_index* ix) {
    if (!ix->ws.stream) LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&ix->ws.stream, cudaStreamNonBlocking));
}

// Sizes the traversal workspace for a launch of `nq` queries with the kernel's real resident-warp count and ef. Called
// from exactly one place (search_device_impl): the byte-map / large-index decision depends on warps_per_sm, so a second
// caller with other defaults would flip the mode back and forth (free + malloc + memset of GBs per call).
void ensure_workspace(const leann_cuda_index* ix, size_t nq, int warps_per_sm, size_t ef) {
    SearchWorkspace& ws = ix->ws;
    if (ix->backend == LEANN_BACKEND_FLAT) return;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
    int max_warps = sms * std::max(warps_per_sm, 4);
    size_t n_pad = (ix->n + 127) & ~(size_t)127;
    int want = (int)std::min<size_t>((size_t)max_warps, std::max<size_t>(nq, 1));
    want = (want + 3) & ~3;
    if (ix->vhash_mode != 1) {
        // byte maps for the full pool of resident warps must fit in a third of device memory, else: large-index mode
        if (!ws.mem_total) { size_t f = 0; LEANN_CUDA_CHECK(cudaMemGetInfo(&f, &ws.mem_total)); }
        const size_t budget = ws.mem_total / 3;
        if (ix->vhash_mode >= 1024 || (size_t)max_warps * n_pad > budget) {
            ensure_workspace_large(ix, want, n_pad, budget, ef);
            return;
        }
    }
    if (ws.large_mode) {   // leaving large-index mode (tuning hook): start over with per-warp byte maps
        ws.reallocs++;
        if (ws.visited) cudaFree(ws.visited);
        if (ws.epochs) cudaFree(ws.epochs);
        ws.visited = nullptr; ws.epochs = nullptr; ws.n_warps = 0; ws.warp_cap = 0; ws.n_pad = 0; ws.large_mode = false;
    }
    if (ws.n_warps >= want && ws.n_pad == n_pad) return;   // fast path: no driver queries
    if (ws.warp_cap && ws.n_warps >= ws.warp_cap && ws.n_pad == n_pad) return;  // already at the memory-bounded maximum
    // bound the visited workspace to ~1/3 of device memory
    size_t free_b = 0, total_b = 0;
    LEANN_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    size_t budget = total_b / 3;
    while (max_warps > 64 && (size_t)max_warps * n_pad > budget) max_warps /= 2;
    ws.warp_cap = max_warps;
    want = std::min(want, max_warps);
    ws.reallocs++;
    if (ws.visited) cudaFree(ws.visited);
    if (ws.epochs) cudaFree(ws.epochs);
    if (ws.pool_locks) cudaFree(ws.pool_locks);
    if (!ws.counter) ws.counter = dalloc<uint32_t>(1);
    // allocate for the full pool once a large batch has been seen, else just what is needed
    ws.n_warps = want;
    ws.n_pad = n_pad;
    ws.visited = dalloc<uint8_t>((size_t)want * n_pad);
    ws.epochs = dalloc<uint32_t>(want);
    ws.pool_locks = dalloc<uint32_t>(want);   // the per-warp maps double as the spill pool of the L2-resident hash mode
    ws.pool_slots = (uint32_t)want;
    LEANN_CUDA_CHECK(cudaMemset(ws.visited, 0, (size_t)want * n_pad));
    LEANN_CUDA_CHECK(cudaMemset(ws.epochs, 0, (size_t)want * 4));
    LEANN_CUDA_CHECK(cudaMemset(ws.pool_locks, 0, (size_t)want * 4));
}

// Throughput batches: per-warp visited hash tables that (mostly) stay in L2 instead of byte maps in HBM, which cost one DRAM
// burst per neighbour test plus the write-back. Long rows (d > 256): 8192-entry tables (64 MB for 12 warps per SM) measured
// 4-7 % faster at ef = 64, 16384-entry tables 1.7 % faster at ef = 148 on 1M x 768; allowed while all tables fit in 128 MB.
// Short rows (d <= 256, register-list kernels): the traversal is bound by DRAM transactions, a third of which are visited
// tags; on the 12.5M x 96 Vamana shard 16384-entry tables (232 MB for 24 warps per SM, about half of it L2-resident)
// measured 5.54 -> 4.98 ms at L = 100 and 8192-entry tables 2.57 ms at L = 50 (benchmarks/k1_visited_ab.py), so up to
// 256 MB of tables are allowed there. A capacity is chosen only when the expected number of visited nodes (about
// 0.7 * ef * degree) leaves the table under 41 % full, so that spills into the byte maps (limit 75 %) stay exceptional.
// Returns the capacity, 0 = byte maps.
uint32_t l2_hash_capacity(const leann_cuda_index* ix, size_t ef) {
    const SearchWorkspace& ws = ix->ws;
    if (ix->vhash_mode != 0 || ws.large_mode || ws.n_warps <= 0) return 0;
    if ((size_t)ws.n_warps * ws.n_pad <= ((size_t)64 << 20)) return 0;   // small index: the byte maps themselves live in L2
    const size_t budget = reduction_lanes(ix->d) == 8 ? ((size_t)256 << 20) : ((size_t)128 << 20);
    for (uint32_t cap : {8192u, 16384u})
        if (70 * ef * ix->M0 <= 41 * (size_t)cap && (size_t)ws.n_warps * cap * 4 <= budget) return cap;
    return 0;
}
// Modular inverse of the table's hash multiplier 0x9E3779B1 (mod 2^32): reconstructs a slot from a quotiented entry.
constexpr uint32_t inv_mod32(uint32_t a) { uint32_t x = a; for (int i = 0; i < 5; ++i) x *= 2u - a * x; return x; }
constexpr uint32_t Q_HASH_INV = inv_mod32(0x9E3779B1u);
static_assert(Q_HASH_INV * 0x9E3779B1u == 1u, "hash multiplier inverse");

// q16 table for the register-list kernels: `slots` 16-bit entries per warp in buckets of 8. Sized for <= ~33 % load at
// the expected 0.7 * ef * degree visited nodes; the slot space must split into bucket bits + at most 14 remainder bits.
// `first_level` > 0: the table is the overflow level of the hybrid form, whose shared-memory level absorbs that many members;
// `tiny`: 1024 entries (tests: overflow and spill on every traversal).
bool q16_plan(const leann_cuda_index* ix, size_t ef, uint32_t* slots, uint32_t* rem_bits, uint32_t* key_bits, size_t first_level = 0,
              bool tiny = false) {
    const bool off = getenv("LEANN_CUDA_DISABLE_Q16") != nullptr;   // A/B switch for benchmarks (read per launch)
    const SearchWorkspace& ws = ix->ws;
    if (off || ix->vhash_mode == 1 || ws.n_warps <= 0 || ix->n < 2) return false;
    if (!ws.large_mode && ix->vhash_mode == 0 && (size_t)ws.n_warps * ws.n_pad <= ((size_t)64 << 20)) return false;   // byte maps already live in L2
    uint32_t B = 1;
    while (B < 32 && ((uint64_t)1 << B) < ix->n) ++B;
    uint32_t S = 8192;
    size_t expect10 = 7 * ef * ix->M0;                                        // 10 x the expected 0.7 * ef * degree members
    if (first_level) expect10 = expect10 > 10 * first_level + 5120 ? expect10 - 10 * first_level : 5120;
    while ((size_t)S * 10 < 3 * expect10 && S < 65536) S <<= 1;               // S >= 3 x expected
    if (tiny) S = 1024;
    if (ix->vhash_mode >= 1024) S = std::max<uint32_t>(1024u, std::min<uint32_t>(65536u, next_pow2((uint32_t)std::min<size_t>(ix->vhash_mode, 65536))));   // tests: tiny tables exercise the spill
    auto bucket_bits = [](uint32_t s) { uint32_t b = 0; while ((8u << b) < s) ++b; return b; };
    while (B > bucket_bits(S) + 14 && S < 65536) S <<= 1;
    const uint32_t bb = bucket_bits(S);
    if (B <= bb || B > bb + 14) return false;
    if ((size_t)ws.n_warps * S * 2 > ((size_t)256 << 20) && !ws.large_mode && ix->vhash_mode == 0) return false;
    *slots = S; *rem_bits = B - bb; *key_bits = B;
    return true;
}

// Shared-memory visited tables (graph_device.cuh "smv") for the register-list kernels (short rows, d <= 128, diskann-rs stop
// rule, throughput batches), up to 2^24 rows per GPU. Small indexes keep the byte maps (they live in L2).
//   hybrid (smem_vis = 2): 4096 two-choice entries per warp in shared memory take the first ~3000 members of a traversal, the
//       q16 table in global memory the rest; the usual 6 CTAs per SM;
//   stand-alone (smem_vis = 1): 8192 entries per warp, 3 CTAs per SM, byte-map spill; taken when the expected 0.7 * ef * degree
//       members stay under 75 % load.
// Both are bit-identical to the other forms and both are OFF by default: on the 12.5M x 96 shard the hybrid measured 1.2 %
// faster at L = 50 (where it never touches its q16 level) and 3.7 % slower at L = 100, the stand-alone form equal at L = 100
// and 3 % slower at L = 50 with unroll 3 (profiles/r2_k1_smv_ab.log, r2_k1_hybrid_ab.log) — taking the visited set off the L2 / DRAM path
// buys almost nothing, i.e. the short-row traversal is not bound by memory transactions (see also benchmarks/gather_probe.cu:
// random 384-byte rows alone stream at 6.9 TB/s). LEANN_CUDA_SMV = 0 / 1 / 2 selects none / stand-alone / hybrid (read once). Visited-set modes (leann_cuda_set_visited_hash):
// 2 / 3 force the stand-alone form (3: 256-entry limit, every traversal spills), 4 / 5 the hybrid (5: 64-entry first level and a
// 1024-entry q16 level, every traversal overflows and spills).
void smv_plan(const leann_cuda_index* ix, SearchParams& p) {
    static const int env_mode = [] { const char* e = getenv("LEANN_CUDA_SMV"); return e ? atoi(e) : 0; }();
    const size_t vm = ix->vhash_mode;
    int mode = vm == 0 ? env_mode : (vm == 2 || vm == 3) ? 1 : (vm == 4 || vm == 5) ? 2 : 0;
    if (mode <= 0 || mode > 2 || ix->n < 2 || ix->d4 > 32) return;
    if (vm == 0 && ix->n < 65536) return;
    if (!graph_search_uses_reg_lists(ix->view(), p)) return;
    uint32_t B = 1;
    while (B < 32 && ((uint64_t)1 << B) < ix->n) ++B;
    if (mode == 1) {
        B = std::max<uint32_t>(B, 11u);
        if (B > 24 || (size_t)7 * p.ef * ix->M0 > (size_t)10 * 6144) return;
        p.smem_vis = 1;
        p.smv_limit = vm == 3 ? 256u : 6656u;
    } else {
        if (B < 10 || B > 24) return;
        p.smem_vis = 2;                      // confirmed once the q16 level is planned (search_device_launch)
        p.smv_limit = vm == 5 ? 64u : 3072u;
    }
    p.q_key_bits = B; p.q_inv = Q_HASH_INV;
}

void ensure_l2_hash(const leann_cuda_index* ix, uint32_t cap) {
    SearchWorkspace& ws = ix->ws;
    const size_t words = (size_t)ws.n_warps * cap;
    if (ws.vhash_words >= words) return;
    ws.reallocs++;
    if (ws.vhash) cudaFree(ws.vhash);
    ws.vhash = nullptr; ws.vhash_words = 0;
    ws.vhash = dalloc<uint32_t>(words);
    ws.vhash_words = words;
}

// All launches on one handle share the traversal workspace (query counter, visited maps, hash tables), the exact-scan
// scratch and the lazily built bf16 copy. Calls arrive under ix->mu, but they may name different streams: each call
// first makes its stream wait for the previous call's work (an event recorded after the last enqueue), so launches
// on one handle execute in call order whatever streams carry them.
void chain_begin(const leann_cuda_index* ix, cudaStream_t stream) {
    if (ix->chain_valid && ix->chain_stream != stream) LEANN_CUDA_CHECK(cudaStreamWaitEvent(stream, ix->chain_ev, 0));
}
void chain_end(const leann_cuda_index* ix, cudaStream_t stream) {
    if (!ix->chain_ev) LEANN_CUDA_CHECK(cudaEventCreateWithFlags(&ix->chain_ev, cudaEventDisableTiming));
    LEANN_CUDA_CHECK(cudaEventRecord(ix->chain_ev, stream));
    ix->chain_stream = stream;
    ix->chain_valid = true;
}

void search_device_launch(const leann_cuda_index* ix, const float* d_queries, size_t nq, size_t k, size_t ef,
                          const uint64_t* d_mask, int mask_mode, uint64_t* d_keys, float* d_dists, uint32_t* d_counts,
                          uint64_t* d_stats, cudaStream_t stream);

void search_device_impl(const leann_cuda_index* ix, const float* d_queries, size_t nq, size_t k, size_t ef,
                        const uint64_t* d_mask, int mask_mode, uint64_t* d_keys, float* d_dists, uint32_t* d_counts,
                        uint64_t* d_stats, cudaStream_t stream) {
    if (!ix) throw Error(LEANN_ERR_INVALID_ARG, "null index");
    if (nq == 0) return;
    if (k == 0) throw Error(LEANN_ERR_INVALID_ARG, "k must be > 0");
    chain_begin(ix, stream);
    search_device_launch(ix, d_queries, nq, k, ef, d_mask, mask_mode, d_keys, d_dists, d_counts, d_stats, stream);
    chain_end(ix, stream);
}

void search_device_launch(const leann_cuda_index* ix, const float* d_queries, size_t nq, size_t k, size_t ef,
                          const uint64_t* d_mask, int mask_mode, uint64_t* d_keys, float* d_dists, uint32_t* d_counts,
                          uint64_t* d_stats, cudaStream_t stream) {
    if (mask_mode == LEANN_MASK_NONE) d_mask = nullptr;
    if (ix->backend == LEANN_BACKEND_FLAT) {
        size_t need = exact_scan_scratch_bytes(ix->d4, (uint32_t)nq, (uint32_t)k);
        if (ix->scan_scratch_bytes < need) {
            if (ix->scan_scratch) cudaFree(ix->scan_scratch);
            ix->scan_scratch = nullptr; ix->scan_scratch_bytes = 0;
            LEANN_CUDA_CHECK(cudaMalloc(&ix->scan_scratch, need));
            ix->scan_scratch_bytes = need;
        }
        FlatView f{ix->vecs, (uint32_t)ix->n, (uint32_t)ix->d, ix->d4, ix->metric};
        TcIndexView tv{};
        const TcIndexView* tvp = nullptr;
        static const bool tc_env_off = getenv("LEANN_CUDA_DISABLE_TC") != nullptr;  // A/B switch for benchmarks
        if (!tc_env_off && !ix->tc_disabled && exact_scan_tc_supported(f, (uint32_t)nq)) {
            const uint32_t dp8 = exact_scan_tc_dp8((uint32_t)ix->d, ix->d4, ix->metric);
            if (!ix->tc_bf16) {
                LEANN_CUDA_CHECK(cudaMalloc(&ix->tc_bf16, ix->n * (size_t)dp8 * 2));
                LEANN_CUDA_CHECK(cudaMalloc(&ix->tc_norms, ix->n * 4));
                LEANN_CUDA_CHECK(cudaMalloc(&ix->tc_xmax, 256));
                exact_scan_tc_prepare(ix->vecs, ix->n, ix->d4, dp8, ix->tc_bf16, ix->tc_norms, ix->tc_xmax, ix->metric, stream);
            }
            tv.x_bf16 = ix->tc_bf16; tv.xmax_bits = ix->tc_xmax; tv.dp8 = dp8;
            tvp = &tv;
        }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
        if (!ix->scan_pinned) LEANN_CUDA_CHECK(cudaMallocHost(&ix->scan_pinned, 64));
        if (!ix->scan_aux.helper) {
            static const bool no_split = getenv("LEANN_CUDA_SCAN_NO_SPLIT") != nullptr;   // A/B switch for benchmarks
            ix->scan_aux.h_flags = ix->scan_pinned;
            if (!no_split) {
                LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&ix->scan_aux.helper, cudaStreamNonBlocking));
                LEANN_CUDA_CHECK(cudaEventCreateWithFlags(&ix->scan_aux.fork, cudaEventDisableTiming));
                LEANN_CUDA_CHECK(cudaEventCreateWithFlags(&ix->scan_aux.join, cudaEventDisableTiming));
            }
        }
        launch_exact_scan(f, d_queries, (uint32_t)nq, (uint32_t)k, d_mask, d_keys, d_dists, d_counts, ix->scan_scratch,
                          ix->scan_scratch_bytes, stream, tvp, sms, ix->scan_aux);
        if (d_stats) LEANN_CUDA_CHECK(cudaMemsetAsync(d_stats, 0, nq * 4 * sizeof(uint64_t), stream));
        return;
    }
    if (ix->n == 0) throw Error(LEANN_ERR_INVALID_ARG, "index is empty");
    size_t eff = std::max(ef, k);
    if (eff > (size_t)MAX_EF) throw Error(LEANN_ERR_INVALID_ARG, "max(ef, k) exceeds 1024");
    SearchParams p;
    p.queries = d_queries;
    p.nq = (uint32_t)nq; p.k = (uint32_t)k; p.ef = (uint32_t)eff;
    p.next_cap = (uint32_t)leann_cuda_queue_capacity(eff, d_mask != nullptr);
    p.next_capp = next_pow2(p.next_cap);
    p.mask = d_mask;
    p.nonstrict_term = (ix->backend == LEANN_BACKEND_VAMANA ? compat::DISKANN_STOP_STRICT : compat::USEARCH_STOP_STRICT) ? 0 : 1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
    // batches of at most two queries per SM leave most of the machine idle with one warp per query: give each query a CTA
    const bool coop = nq <= (size_t)sms * 2 && ix->coop_small_batches;
    p.coop_ctas = coop ? 1 : 0;
    p.coop_warps = nq <= (size_t)sms ? 8 : 4;
    p.vhash = nullptr; p.vhash_cap = 1024u;
    p.vhash16 = 0; p.q_rem_bits = 0; p.q_key_bits = 32; p.q_inv = 0;
    p.smem_vis = 0; p.smv_limit = 0;
    if (!coop) smv_plan(ix, p);   // short rows, throughput batch: visited tables in shared memory (fewer resident warps)
    p.row_ring = 0;
    {
        // LEANN_CUDA_RING = 1: rows of a hop through a shared-memory ring of bulk async copies (A/B form, measured slower: off)
        static const int ring_env = [] { const char* e = getenv("LEANN_CUDA_RING"); return e ? atoi(e) : 0; }();
        if (!coop && !p.smem_vis && ring_env && ix->d4 <= 32 && graph_search_uses_reg_lists(ix->view(), p)) p.row_ring = 1;
    }
    const int warps_per_sm = graph_search_warps_per_sm(ix->view(), p.ef, p.next_capp, p.nonstrict_term, p.smem_vis, p.row_ring);
    ensure_workspace(ix, nq, warps_per_sm, p.ef);
    p.out_keys = d_keys; p.out_dists = d_dists; p.out_counts = d_counts; p.out_stats = d_stats;
    p.visited = ix->ws.visited; p.epochs = ix->ws.epochs; p.counter = ix->ws.counter;
    p.n_pad = ix->ws.n_pad;
    p.n_warps = (int)std::min<size_t>((size_t)ix->ws.n_warps, (nq + 3) & ~(size_t)3);
    if (p.smem_vis) p.n_warps = std::min(p.n_warps, std::max(4, warps_per_sm * sms));   // the workspace may have been sized by a launch with more resident warps
    p.vhash = ix->ws.large_mode ? ix->ws.vhash : nullptr;
    p.vhash_cap = ix->ws.large_mode ? ix->ws.vhash_cap : 1024u;
    p.pool_locks = ix->ws.pool_locks; p.pool_slots = std::max<uint32_t>(ix->ws.pool_slots, 1u);
    p.coop_ctas = coop ? (int)std::min<size_t>(nq, (size_t)ix->ws.n_warps) : 0;
    if (p.smem_vis == 2) {   // hybrid: the overflow level is a (smaller) q16 table
        uint32_t q16_slots = 0, rem_bits = 0, key_bits = 0;
        if (q16_plan(ix, p.ef, &q16_slots, &rem_bits, &key_bits, p.smv_limit, ix->vhash_mode == 5) && key_bits == p.q_key_bits) {
            ensure_l2_hash(ix, q16_slots / 2);   // words per warp
            p.vhash = ix->ws.vhash; p.vhash_cap = q16_slots; p.vhash16 = 1; p.q_rem_bits = rem_bits;
        } else {
            p.smem_vis = 0; p.smv_limit = 0; p.q_key_bits = 32; p.q_inv = 0;
        }
    }
    // throughput batches only: a latency-bound single traversal pays more for the CAS round trips than it saves (measured +20 %)
    if (p.coop_ctas == 0 && !p.smem_vis) {
        uint32_t q16_slots = 0;
        if (graph_search_uses_reg_lists(ix->view(), p) && q16_plan(ix, p.ef, &q16_slots, &p.q_rem_bits, &p.q_key_bits)) {
            // short rows: bucketed table of 16-bit quotiented entries (graph_device.cuh), L2-resident for all resident warps
            ensure_l2_hash(ix, q16_slots / 2);   // words per warp
            p.vhash = ix->ws.vhash; p.vhash_cap = q16_slots; p.vhash16 = 1; p.q_inv = Q_HASH_INV;
        } else if (!ix->ws.large_mode) {
            if (const uint32_t cap = l2_hash_capacity(ix, p.ef)) {
                ensure_l2_hash(ix, cap);
                p.vhash = ix->ws.vhash; p.vhash_cap = cap;
            }
        }
    }
    launch_graph_search(ix->view(), p, stream);
}

template <typename T>
void ensure_buf(T*& p, size_t& cap, size_t need) {
    if (cap >= need) return;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    LEANN_CUDA_CHECK(cudaMalloc(&p, need * sizeof(T)));
    cap = need;
}

}  // namespace

namespace leann {
void backend_search_device(const leann_cuda_index* ix, const float* d_queries, size_t nq, size_t k, size_t ef,
                           const uint64_t* d_mask, uint64_t* d_keys, float* d_dists, uint32_t* d_counts, cudaStream_t stream) {
    search_device_impl(ix, d_queries, nq, k, ef, d_mask, d_mask ? LEANN_MASK_INLINE : LEANN_MASK_NONE, d_keys, d_dists, d_counts,
                       nullptr, stream);
}
cudaStream_t backend_stream(const leann_cuda_index* ix) {
    ensure_stream(ix);
    return ix->ws.stream;
}
int guard_impl(char* err, size_t errlen, const std::function<void()>& f) {
    auto put = [&](const char* m) { if (err && errlen) { snprintf(err, errlen, "%s", m); } };
    try {
        if (err && errlen) err[0] = 0;
        f();
        return LEANN_OK;
    } catch (const Error& e) {
        put(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        put("out of host memory");
        return LEANN_ERR_INVALID_ARG;
    } catch (const std::exception& e) {
        put(e.what());
        return LEANN_ERR_INVALID_ARG;
    } catch (...) {
        put("unknown error");
        return LEANN_ERR_INVALID_ARG;
    }
}
}  // namespace leann

#define GUARD(...) return leann::guard_impl(err, errlen, [&]() __VA_ARGS__)

extern "C" {

int leann_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
unsigned leann_cuda_compat_flags(void) { return compat::FLAGS; }
const char* leann_cuda_version(void) { return "leann-cuda 0.1.0 (sm_100a)"; }
int leann_cuda_reduction_lanes(size_t dims) { return reduction_lanes(dims); }
size_t leann_cuda_queue_capacity(size_t ef, int masked) { return masked ? std::min<size_t>(4 * ef, 2048) : ef; }

int leann_cuda_open(const char* base_path, int backend, size_t dims, int metric, int device,
                    leann_cuda_index** out, char* err, size_t errlen) {
    GUARD({
        if (!base_path || !out) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        std::string base(base_path);
        if (backend == LEANN_BACKEND_HNSW) {
            std::string file = with_extension(base, "index");
            if (is_faiss_index(file))  // hnsw.rs:24-32
                throw Error(LEANN_ERR_FAISS_FORMAT,
                            "This index was built with Python LEANN (FAISS format).\nRust LEANN uses usearch which has a different binary format.\n\n"
                            "To use this index with Rust LEANN, you need to rebuild it:\n  leann build <name> --docs <path> --force\n\n"
                            "The passages and metadata files are compatible and will be preserved.");
            usearch_probe(file, dims);   // header / format / dimension errors are reported before the device is touched
            require_gpu(device);
            DeviceGuard dg(device);
            bool used_cache = false;
            *out = open_hnsw_streamed(base, dims, device, metric, &used_cache);
            (*out)->layout_cache_used = used_cache;
            // LEANN_CUDA_LAYOUT_CACHE=1: leave <base>.cuda-layout behind after a parse so the next open skips it
            if (!used_cache && getenv("LEANN_CUDA_LAYOUT_CACHE") != nullptr) {
                try { write_layout_cache(*out, base); } catch (const Error&) {}   // read-only index directory: not an error
            }
        } else if (backend == LEANN_BACKEND_VAMANA) {
            diskann_probe(with_extension(base, "diskann"), dims);
            require_gpu(device);
            DeviceGuard dg(device);
            *out = open_vamana_streamed(base, dims, device, metric);
        } else if (backend == LEANN_BACKEND_FLAT) {
            if (dims == 0) throw Error(LEANN_ERR_INVALID_ARG, "embeddings: dimensions must be given (the file has no header)");
            {
                FILE* probe = fopen(with_extension(base, "embeddings").c_str(), "rb");
                if (!probe) throw Error(LEANN_ERR_NOT_FOUND, "Embeddings file not found: " + with_extension(base, "embeddings"));
                fclose(probe);
            }
            require_gpu(device);
            DeviceGuard dg(device);
            *out = open_flat_streamed(base, dims, device, metric);
        } else {
            throw Error(LEANN_ERR_INVALID_ARG, "Unknown backend");  // searcher.rs:98
        }
    });
}

static int flat_from(const float* vectors, bool on_device, size_t n, size_t dims, int metric, int device,
                     leann_cuda_index** out, char* err, size_t errlen) {
    GUARD({
        if (!out || (!vectors && n)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if (dims == 0) throw Error(LEANN_ERR_INVALID_ARG, "dims must be > 0");
        require_gpu(device);
        DeviceGuard dg(device);
        std::unique_ptr<leann_cuda_index> ix(new leann_cuda_index());
        ix->backend = LEANN_BACKEND_FLAT; ix->device = device;
        ix->metric = metric == LEANN_METRIC_DEFAULT ? LEANN_METRIC_DOT_DESC : metric;
        ix->n = n; ix->d = dims;
        upload_vectors(ix.get(), vectors, on_device);
        *out = ix.release();
    });
}
int leann_cuda_flat_from_host(const float* vectors, size_t n, size_t dims, int metric, int device,
                              leann_cuda_index** out, char* err, size_t errlen) {
    return flat_from(vectors, false, n, dims, metric, device, out, err, errlen);
}
int leann_cuda_flat_from_device(const float* d_vectors, size_t n, size_t dims, int metric, int device,
                                leann_cuda_index** out, char* err, size_t errlen) {
    return flat_from(d_vectors, true, n, dims, metric, device, out, err, errlen);
}

int leann_cuda_hnsw_build(const float* vectors, int vectors_on_device, size_t n, size_t dims,
                          size_t graph_degree, size_t complexity, int metric, uint64_t seed, int device,
                          leann_cuda_index** out, char* err, size_t errlen) {
    GUARD({
        if (!out || (!vectors && n)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if (dims == 0 || graph_degree < 2 || 2 * graph_degree > (size_t)MAX_DEG)
            throw Error(LEANN_ERR_INVALID_ARG, "graph_degree must be in 2..64 and dims > 0");
        if (complexity == 0 || complexity > (size_t)MAX_EF) throw Error(LEANN_ERR_INVALID_ARG, "complexity must be in 1..1024");
        require_gpu(device);
        DeviceGuard dg(device);
        std::unique_ptr<leann_cuda_index> ix(new leann_cuda_index());
        ix->backend = LEANN_BACKEND_HNSW; ix->device = device;
        ix->metric = metric == LEANN_METRIC_DEFAULT ? LEANN_METRIC_IP : metric;  // hnsw.rs:112
        ix->n = n; ix->d = dims;
        upload_vectors(ix.get(), vectors, vectors_on_device != 0);
        gpu_hnsw_build(ix.get(), graph_degree, complexity, seed);
        *out = ix.release();
    });
}

int leann_cuda_hnsw_add(leann_cuda_index* ix, const float* vectors, int vectors_on_device, size_t m, uint64_t start_id,
                        size_t complexity, uint64_t seed, char* err, size_t errlen) {
    GUARD({
        if (!ix || (!vectors && m)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        if (ix->backend != LEANN_BACKEND_HNSW) throw Error(LEANN_ERR_INVALID_ARG, "add_to_index is defined for the HNSW backend only (hnsw.rs:142)");
        if (complexity == 0 || complexity > (size_t)MAX_EF) throw Error(LEANN_ERR_INVALID_ARG, "complexity must be in 1..1024");
        if (m == 0) return;
        DeviceGuard dg(ix->device);
        std::lock_guard<std::mutex> lk(ix->mu);
        // padded device copy of the new rows
        float4* rows = dalloc<float4>(m * ix->d4);
        try {
            if (vectors_on_device) {
                launch_pad_rows(vectors, rows, m, (uint32_t)ix->d, ix->d4, nullptr);
            } else {
                float* tmp = dalloc<float>(m * ix->d);
                cudaError_t e = cudaMemcpy(tmp, vectors, m * ix->d * 4, cudaMemcpyHostToDevice);
                if (e == cudaSuccess) launch_pad_rows(tmp, rows, m, (uint32_t)ix->d, ix->d4, nullptr);
                cudaDeviceSynchronize();
                cudaFree(tmp);
                LEANN_CUDA_CHECK(e);
            }
            LEANN_CUDA_CHECK(cudaDeviceSynchronize());
            gpu_hnsw_add(ix, rows, m, start_id, complexity, seed);
        } catch (...) {
            cudaFree(rows);
            throw;
        }
        cudaFree(rows);
        // the traversal workspace is sized by n: drop it, the next search re-creates it
        SearchWorkspace& ws = ix->ws;
        if (ws.visited) cudaFree(ws.visited);
        if (ws.epochs) cudaFree(ws.epochs);
        if (ws.pool_locks) cudaFree(ws.pool_locks);
        ws.visited = nullptr; ws.epochs = nullptr; ws.pool_locks = nullptr; ws.n_warps = 0; ws.warp_cap = 0; ws.n_pad = 0;
        ws.large_mode = false;
    });
}

int leann_cuda_set_visited_hash(leann_cuda_index* ix, size_t capacity) {
    if (!ix || (capacity > 5 && capacity < 1024)) return LEANN_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->vhash_mode = capacity;
    return LEANN_OK;
}

int leann_cuda_vamana_build(const float* vectors, int vectors_on_device, size_t n, size_t dims,
                            size_t graph_degree, size_t complexity, float alpha, int metric,
                            uint64_t seed, int device, leann_cuda_index** out, char* err, size_t errlen) {
    GUARD({
        if (!out || (!vectors && n)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if (dims == 0 || graph_degree < 2 || graph_degree > (size_t)MAX_DEG)
            throw Error(LEANN_ERR_INVALID_ARG, "graph_degree must be in 2..128 and dims > 0");
        if (complexity == 0 || complexity > (size_t)MAX_EF) throw Error(LEANN_ERR_INVALID_ARG, "complexity must be in 1..1024");
        require_gpu(device);
        DeviceGuard dg(device);
        std::unique_ptr<leann_cuda_index> ix(new leann_cuda_index());
        ix->backend = LEANN_BACKEND_VAMANA; ix->device = device;
        ix->metric = metric == LEANN_METRIC_DEFAULT ? LEANN_METRIC_IP_CLAMP : metric;
        ix->n = n; ix->d = dims; ix->distance_name = "DistDot";
        upload_vectors(ix.get(), vectors, vectors_on_device != 0);
        gpu_vamana_build(ix.get(), graph_degree, complexity, alpha, seed);
        *out = ix.release();
    });
}

// Host-only: parses the whole file exactly as leann_cuda_open does (same header pass, same node parser) and reports what
// it holds. No device is touched, so index files can be validated on machines without a GPU (and the reader is covered by the
// CPU test suite).
int leann_cuda_check_index_file(const char* base_path, int backend, size_t dims, uint64_t* info8, char* err, size_t errlen) {
    GUARD({
        if (!base_path || !info8) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        std::string base(base_path);
        auto mix = [](uint64_t h, uint64_t v) { h ^= v; h *= 0x100000001B3ull; return h; };
        uint64_t h = 0xCBF29CE484222325ull;
        if (backend == LEANN_BACKEND_HNSW) {
            const std::string file = with_extension(base, "index");
            if (is_faiss_index(file)) throw Error(LEANN_ERR_FAISS_FORMAT, "This index was built with Python LEANN (FAISS format).");
            HostHnsw g;
            read_usearch_index(file, dims, g);
            for (size_t i = 0; i < g.n; ++i) {      // level-0 lists, then upper lists, in list order; then keys
                for (size_t j = 0; j < g.M0 && g.adj0[i * g.M0 + j] != SENT; ++j) h = mix(h, g.adj0[i * g.M0 + j]);
                h = mix(h, 0xFFFFFFFFFFull);
                for (int l = 1; l <= g.levels[i]; ++l) {
                    const uint32_t* p = &g.adjU[((size_t)g.upper_base[i] + (size_t)(l - 1)) * g.M];
                    for (size_t j = 0; j < g.M && p[j] != SENT; ++j) h = mix(h, p[j]);
                    h = mix(h, 0xFFFFFFFFFFull);
                }
                h = mix(h, g.keys[i]);
            }
            info8[0] = g.n; info8[1] = g.d; info8[2] = g.M; info8[3] = g.M0; info8[4] = (uint64_t)g.max_level; info8[5] = g.entry;
            info8[6] = g.M ? g.adjU.size() / g.M : 0; info8[7] = h;
        } else if (backend == LEANN_BACKEND_VAMANA) {
            HostVamana g;
            read_diskann(with_extension(base, "diskann"), dims, g);
            for (size_t i = 0; i < g.n; ++i) {
                for (size_t j = 0; j < g.R && g.adj[i * g.R + j] != SENT; ++j) h = mix(h, g.adj[i * g.R + j]);
                h = mix(h, 0xFFFFFFFFFFull);
                h = mix(h, i);
            }
            info8[0] = g.n; info8[1] = g.d; info8[2] = g.R; info8[3] = g.R; info8[4] = 0; info8[5] = g.medoid; info8[6] = 0; info8[7] = h;
        } else if (backend == LEANN_BACKEND_FLAT) {
            std::vector<float> v;
            size_t n = 0;
            read_embeddings(with_extension(base, "embeddings"), dims, v, n);
            info8[0] = n; info8[1] = dims; info8[2] = info8[3] = info8[4] = info8[5] = info8[6] = 0; info8[7] = h;
        } else {
            throw Error(LEANN_ERR_INVALID_ARG, "Unknown backend");
        }
    });
}

int leann_cuda_write_layout_cache(const leann_cuda_index* ix, const char* base_path, char* err, size_t errlen) {
    GUARD({
        if (!ix || !base_path) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        DeviceGuard dg(ix->device);
        write_layout_cache(ix, base_path);
    });
}
int leann_cuda_layout_cache_used(const leann_cuda_index* ix) { return ix && ix->layout_cache_used ? 1 : 0; }

int leann_cuda_save(const leann_cuda_index* ix, const char* base_path, char* err, size_t errlen) {
    GUARD({
        if (!ix || !base_path) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        DeviceGuard dg(ix->device);
        std::string base(base_path);
        // un-pad vectors
        std::vector<float> padded(ix->n * ix->d4 * 4), vecs(ix->n * ix->d);
        if (ix->n) LEANN_CUDA_CHECK(cudaMemcpy(padded.data(), ix->vecs, padded.size() * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < ix->n; ++i) memcpy(&vecs[i * ix->d], &padded[i * ix->d4 * 4], ix->d * 4);
        padded.clear(); padded.shrink_to_fit();
        auto download = [](auto& v, const void* src) { if (!v.empty()) LEANN_CUDA_CHECK(cudaMemcpy(v.data(), src, v.size() * sizeof(v[0]), cudaMemcpyDeviceToHost)); };
        if (ix->backend == LEANN_BACKEND_HNSW) {
            HostHnsw h;
            h.n = ix->n; h.d = ix->d; h.M = ix->M; h.M0 = ix->M0; h.max_level = ix->max_level; h.entry = ix->entry;
            h.metric = ix->metric; h.vecs.swap(vecs); h.levels = ix->h_levels;
            h.keys.resize(ix->n); h.adj0.resize(ix->n * ix->M0); h.upper_base.resize(ix->n); h.adjU.resize(ix->n_upper_lists * ix->M);
            download(h.keys, ix->keys); download(h.adj0, ix->adj0); download(h.upper_base, ix->upper_base); download(h.adjU, ix->adjU);
            write_usearch_index(with_extension(base, "index"), h);
        } else if (ix->backend == LEANN_BACKEND_VAMANA) {
            HostVamana h;
            h.n = ix->n; h.d = ix->d; h.R = ix->M0; h.medoid = ix->entry; h.distance_name = ix->distance_name;
            h.vecs.swap(vecs); h.adj.resize(ix->n * ix->M0);
            download(h.adj, ix->adj0);
            write_diskann(with_extension(base, "diskann"), h);
        } else {
            write_embeddings(with_extension(base, "embeddings"), vecs.data(), ix->n, ix->d);
        }
    });
}

size_t leann_cuda_len(const leann_cuda_index* ix) { return ix ? ix->n : 0; }
size_t leann_cuda_dims(const leann_cuda_index* ix) { return ix ? ix->d : 0; }
int leann_cuda_info(const leann_cuda_index* ix, uint64_t* info) {
    if (!ix || !info) return LEANN_ERR_INVALID_ARG;
    info[0] = ix->n; info[1] = ix->d; info[2] = (uint64_t)ix->backend; info[3] = (uint64_t)ix->metric;
    info[4] = ix->M; info[5] = ix->M0; info[6] = (uint64_t)ix->max_level; info[7] = ix->entry;
    return LEANN_OK;
}

int leann_cuda_search_device(const leann_cuda_index* ix, const float* d_queries, size_t nq, size_t k, size_t ef,
                             const uint64_t* d_mask_bits, int mask_mode, uint64_t* d_keys, float* d_dists,
                             uint32_t* d_counts, uint64_t* d_stats, void* cuda_stream, char* err, size_t errlen) {
    GUARD({
        if (!ix) throw Error(LEANN_ERR_INVALID_ARG, "null index");
        DeviceGuard dg(ix->device);
        std::lock_guard<std::mutex> lk(ix->mu);
        search_device_impl(ix, d_queries, nq, k, ef, d_mask_bits, mask_mode, d_keys, d_dists, d_counts, d_stats,
                           (cudaStream_t)cuda_stream);
    });
}

static void search_host_locked(const leann_cuda_index* ix, const float* queries, size_t nq, size_t k, size_t ef,
                               const uint64_t* mask_bits, int mask_mode, uint64_t* keys, float* dists, uint32_t* counts) {
    DeviceGuard dg(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    ensure_stream(ix);   // the traversal workspace is sized once, by search_device_impl, with the kernel's real parameters
    SearchWorkspace& ws = ix->ws;
    ensure_buf(ws.d_queries, ws.cap_q, nq * ix->d);
    size_t out_need = nq * k;
    if (ws.cap_out < out_need) {
        if (ws.d_keys) cudaFree(ws.d_keys);
        if (ws.d_dists) cudaFree(ws.d_dists);
        ws.d_keys = nullptr; ws.d_dists = nullptr; ws.cap_out = 0;
        LEANN_CUDA_CHECK(cudaMalloc(&ws.d_keys, out_need * 8));
        LEANN_CUDA_CHECK(cudaMalloc(&ws.d_dists, out_need * 4));
        ws.cap_out = out_need;
    }
    ensure_buf(ws.d_counts, ws.cap_counts, nq);
    uint32_t* d_counts = ws.d_counts;
    const uint64_t* d_mask = nullptr;
    if (mask_bits && mask_mode != LEANN_MASK_NONE) {
        size_t words = (ix->n + 63) / 64;
        ensure_buf(ws.d_mask, ws.cap_mask, words);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(ws.d_mask, mask_bits, words * 8, cudaMemcpyHostToDevice, ws.stream));
        d_mask = ws.d_mask;
    }
    LEANN_CUDA_CHECK(cudaMemcpyAsync(ws.d_queries, queries, nq * ix->d * 4, cudaMemcpyHostToDevice, ws.stream));
    search_device_impl(ix, ws.d_queries, nq, k, ef, d_mask, mask_mode, ws.d_keys, ws.d_dists, d_counts, nullptr, ws.stream);
    LEANN_CUDA_CHECK(cudaMemcpyAsync(keys, ws.d_keys, out_need * 8, cudaMemcpyDeviceToHost, ws.stream));
    LEANN_CUDA_CHECK(cudaMemcpyAsync(dists, ws.d_dists, out_need * 4, cudaMemcpyDeviceToHost, ws.stream));
    if (counts) LEANN_CUDA_CHECK(cudaMemcpyAsync(counts, d_counts, nq * 4, cudaMemcpyDeviceToHost, ws.stream));
    LEANN_CUDA_CHECK(cudaStreamSynchronize(ws.stream));
}

// One query from one thread, coalescing enabled: join (or lead) a batch. One batch is in flight per handle at a time
// (launches on a handle are serialised anyway): requests that arrive while it runs queue up and the next leader takes all
// of them, so a lone caller pays no waiting time (max_wait_us = 0, the default) and concurrent callers batch naturally.
static void search_coalesced(const leann_cuda_index* ix, const float* query, size_t k, size_t ef, uint64_t* keys, float* dists,
                             uint32_t* count) {
    Coalescer& c = ix->coalescer;
    CoalesceReq r{query, k, ef, keys, dists, count};
    std::unique_lock<std::mutex> lk(c.m);
    c.queue.push_back(&r);
    c.cv_leader.notify_one();
    while (!r.done) {
        if (r.taken || c.leader_active) { c.cv_done.wait(lk); continue; }
        c.leader_active = true;
        if (c.max_wait_us > 0) {
            auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(c.max_wait_us);
            while (c.queue.size() < c.max_batch)
                if (c.cv_leader.wait_until(lk, deadline) == std::cv_status::timeout) break;
        }
        // the batch = queued requests with this thread's (k, ef), in arrival order, at most max_batch (this one included)
        std::vector<CoalesceReq*> batch, rest;
        try {
            batch.reserve(c.queue.size()); rest.reserve(c.queue.size());
            batch.push_back(&r);
            for (CoalesceReq* q : c.queue) {
                if (q == &r) continue;
                ((q->k == r.k && q->ef == r.ef && batch.size() < c.max_batch) ? batch : rest).push_back(q);
            }
        } catch (...) {   // out of memory while forming the batch: run alone
            batch.clear(); rest.clear();
            r.taken = true;
            c.queue.erase(std::find(c.queue.begin(), c.queue.end(), &r));
        }
        if (!batch.empty()) {
            for (CoalesceReq* q : batch) q->taken = true;
            c.queue.swap(rest);
        }
        c.batches++; c.requests += batch.empty() ? 1 : batch.size();
        lk.unlock();
        int rc = 0;
        std::string msg;
        const size_t nb = batch.empty() ? 1 : batch.size(), d = ix->d;
        try {
            if (nb == 1) {
                search_host_locked(ix, query, 1, k, ef, nullptr, LEANN_MASK_NONE, keys, dists, count);
            } else {
                std::vector<float> qbuf(nb * d);
                std::vector<uint64_t> kbuf(nb * k);
                std::vector<float> dbuf(nb * k);
                std::vector<uint32_t> cbuf(nb);
                for (size_t i = 0; i < nb; ++i) memcpy(&qbuf[i * d], batch[i]->query, d * 4);
                search_host_locked(ix, qbuf.data(), nb, k, ef, nullptr, LEANN_MASK_NONE, kbuf.data(), dbuf.data(), cbuf.data());
                for (size_t i = 0; i < nb; ++i) {
                    CoalesceReq* q = batch[i];
                    memcpy(q->keys, &kbuf[i * k], k * 8);
                    memcpy(q->dists, &dbuf[i * k], k * 4);
                    if (q->count) *q->count = cbuf[i];
                }
            }
        } catch (const Error& e) { rc = e.code; msg = e.what(); }
        catch (const std::bad_alloc&) { rc = LEANN_ERR_INVALID_ARG; msg = "out of host memory"; }
        catch (const std::exception& e) { rc = LEANN_ERR_INVALID_ARG; msg = e.what(); }
        catch (...) { rc = LEANN_ERR_INVALID_ARG; msg = "unknown error"; }
        // every request of the batch is completed (with the error code, if any) whatever happened above: a follower
        // left without done = true would wait forever
        lk.lock();
        if (batch.empty()) { r.rc = rc; r.done = true; }
        for (CoalesceReq* q : batch) {
            q->rc = rc;
            if (rc != 0) { try { q->err = msg; } catch (...) {} }
            q->done = true;
        }
        c.leader_active = false;
        c.cv_done.notify_all();
    }
    if (r.rc != 0) throw Error(r.rc, r.err);
}

int leann_cuda_search(const leann_cuda_index* ix, const float* queries, size_t nq, size_t k, size_t ef,
                      const uint64_t* mask_bits, int mask_mode, uint64_t* keys, float* dists, uint32_t* counts,
                      char* err, size_t errlen) {
    GUARD({
        if (!ix) throw Error(LEANN_ERR_INVALID_ARG, "null index");
        if (nq == 0) return;
        if (!queries || !keys || !dists) throw Error(LEANN_ERR_INVALID_ARG, "null buffer");
        if (k == 0) throw Error(LEANN_ERR_INVALID_ARG, "k must be > 0");
        const bool masked = mask_bits && mask_mode != LEANN_MASK_NONE;
        if (nq == 1 && !masked && ix->coalescer.max_batch > 1) {
            search_coalesced(ix, queries, k, ef, keys, dists, counts);
            return;
        }
        search_host_locked(ix, queries, nq, k, ef, mask_bits, mask_mode, keys, dists, counts);
    });
}

int leann_cuda_set_coalescing(leann_cuda_index* ix, size_t max_batch, unsigned max_wait_us) {
    if (!ix) return LEANN_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ix->coalescer.m);
    ix->coalescer.max_batch = max_batch;
    ix->coalescer.max_wait_us = max_wait_us;
    return LEANN_OK;
}
int leann_cuda_workspace_stats(const leann_cuda_index* ix, uint64_t* stats4) {
    if (!ix || !stats4) return LEANN_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ix->mu);
    const SearchWorkspace& ws = ix->ws;
    stats4[0] = ws.reallocs;
    stats4[1] = ws.large_mode ? (uint64_t)ws.pool_slots * ws.n_pad + ws.vhash_words * 4 : (uint64_t)ws.n_warps * ws.n_pad + ws.vhash_words * 4;
    stats4[2] = ws.large_mode ? 1 : 0;
    stats4[3] = (uint64_t)ws.n_warps;
    return LEANN_OK;
}
int leann_cuda_coalescing_stats(const leann_cuda_index* ix, uint64_t* batches, uint64_t* requests) {
    if (!ix) return LEANN_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(ix->coalescer.m);
    if (batches) *batches = ix->coalescer.batches;
    if (requests) *requests = ix->coalescer.requests;
    return LEANN_OK;
}

int leann_cuda_topk_merge_device(const uint64_t* d_keys_in, const float* d_dists_in, size_t n_shards, size_t nq,
                                 size_t k, int descending, uint64_t* d_keys_out, float* d_dists_out,
                                 uint32_t* d_counts_out, void* cuda_stream, char* err, size_t errlen) {
    GUARD({
        if (!d_keys_in || !d_dists_in || !d_keys_out || !d_dists_out) throw Error(LEANN_ERR_INVALID_ARG, "null buffer");
        launch_topk_merge(d_keys_in, d_dists_in, (uint32_t)n_shards, (uint32_t)nq, (uint32_t)k, descending, d_keys_out,
                          d_dists_out, d_counts_out, (cudaStream_t)cuda_stream);
    });
}

void leann_cuda_close(leann_cuda_index* ix) {
    if (!ix) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ix->device);
    cudaFree(ix->vecs); cudaFree(ix->adj0); cudaFree(ix->upper_base); cudaFree(ix->adjU); cudaFree(ix->keys);
    SearchWorkspace& ws = ix->ws;
    cudaFree(ws.visited); cudaFree(ws.epochs); cudaFree(ws.counter); cudaFree(ws.vhash); cudaFree(ws.pool_locks);
    cudaFree(ws.d_queries); cudaFree(ws.d_keys);
    cudaFree(ws.d_dists); cudaFree(ws.d_counts); cudaFree(ws.d_mask);
    if (ws.stream) cudaStreamDestroy(ws.stream);
    if (ix->chain_ev) cudaEventDestroy(ix->chain_ev);
    if (ix->scan_aux.helper) cudaStreamDestroy(ix->scan_aux.helper);
    if (ix->scan_aux.fork) cudaEventDestroy(ix->scan_aux.fork);
    if (ix->scan_aux.join) cudaEventDestroy(ix->scan_aux.join);
    if (ix->scan_pinned) cudaFreeHost(ix->scan_pinned);
    cudaFree(ix->scan_scratch); cudaFree(ix->tc_bf16); cudaFree(ix->tc_norms); cudaFree(ix->tc_xmax);
    cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
    delete ix;
}

}  // extern "C"
