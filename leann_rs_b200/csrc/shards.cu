// shards.cu — the database-sharded layout behind the C ABI (BASELINE north_star: "the database and
// graph are sharded across the GPUs of one box, each GPU searches its shard, and a per-query top-k merge runs over
// NCCL allgather on NVLink"). A `leann_cuda_shards` handle is what BackendType::load_searcher
// (leann-rs src/backend/mod.rs:23-45) would return for an index split over several GPUs: `search` has the contract
// of BackendSearcher::search (src/backend/traits.rs:16-21), keys are global (shard key + the shard's key offset).
//
// Two ways to own the shards:
//   * one host process, several devices  (leann_cuda_shards_open / _from_indexes): one sub-index, stream and result
//     block per device. When every device can map its peers' memory (NVLink / NVSwitch) the exchange step and the
//     merge are ONE kernel: K4p on device 0 waits (stream events) for the shards' searches and reads their result
//     blocks straight out of peer memory over NVLink — no collective launch at all. Otherwise ncclCommInitAll +
//     ncclAllGather + K4.
//   * one process per GPU (torchrun-style; leann_cuda_shards_join): every rank owns one shard and joins an NCCL
//     communicator built from a unique id (leann_cuda_comm_unique_id, broadcast by the caller's own plumbing);
//     per batch: K1/K2 -> ONE ncclAllGather of the packed (keys, dists) block -> K4 on the same stream.
// libnccl is loaded with dlopen at first use (the copy already mapped into the process, e.g. torch's, is preferred),
// so the library itself has no link-time NCCL dependency and single-GPU users never touch it.
#include <dlfcn.h>
#include <math_constants.h>
#include <nccl.h>

#include <algorithm>
#include <memory>

#include "scan_common.h"

namespace leann {
int guard_impl(char* err, size_t errlen, const std::function<void()>& f);
void backend_search_device(const leann_cuda_index* ix, const float* d_queries, size_t nq, size_t k, size_t ef,
                           const uint64_t* d_mask, uint64_t* d_keys, float* d_dists, uint32_t* d_counts, cudaStream_t stream);
}  // namespace leann

using namespace leann;

#define GUARD(...) return leann::guard_impl(err, errlen, [&]() __VA_ARGS__)

namespace {

constexpr int MAX_SHARDS = 16;

// ---- NCCL through dlopen ----------------------------------------------------------------------------
struct Nccl {
    void* h = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

Nccl& nccl() {
    static Nccl n;
    static std::mutex m;
    std::lock_guard<std::mutex> lk(m);
    if (n.h) return n;
    const char* override_path = getenv("LEANN_CUDA_NCCL_LIB");
    void* h = nullptr;
    if (override_path && *override_path) h = dlopen(override_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy this process already uses (e.g. torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) throw Error(LEANN_ERR_NCCL, std::string("libnccl.so.2 cannot be loaded (set LEANN_CUDA_NCCL_LIB): ") + dlerror());
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        if (!p) throw Error(LEANN_ERR_NCCL, std::string("libnccl lacks ") + name);
        return p;
    };
    n.GetVersion = (decltype(n.GetVersion))sym("ncclGetVersion");
    n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
    n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
    n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
    n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
    n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
    n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
    n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    n.h = h;
    return n;
}

#define LEANN_NCCL_CHECK(expr)                                                                              \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess)                                                                              \
            throw Error(LEANN_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(_r));             \
    } while (0)

struct DeviceScope {
    int prev = -1;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        LEANN_CUDA_CHECK(cudaSetDevice(dev));
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

// ---- K4p: per-query merge of G sorted lists read through per-shard pointers -------------------------
// The pointers may name a gathered buffer on this device (after ncclAllGather) or the shards' own result blocks
// in PEER memory (single-process P2P mode: the loads travel over NVLink, the gather and the merge are one kernel).
// Keys become global here: key + offset[s]. Ties: lower shard, then lower rank first (= ascending global key for
// contiguous shards, the order a stable sort over the concatenated database yields).
struct MergeSrc {
    const uint64_t* keys[MAX_SHARDS];
    const float* dists[MAX_SHARDS];
    uint64_t offset[MAX_SHARDS];
};

__device__ __forceinline__ uint32_t order_bits(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

__global__ void __launch_bounds__(128)
topk_merge_ptr_kernel(const MergeSrc src, uint32_t n_shards, uint32_t nq, uint32_t k, int descending,
                      uint64_t* __restrict__ keys_out, float* __restrict__ dists_out, uint32_t* __restrict__ counts_out) {
    extern __shared__ unsigned long long skeys[];   // n2 packed (rank bits, position)
    const uint32_t q = blockIdx.x;
    const uint32_t total = n_shards * k;
    uint32_t n2 = 1;
    while (n2 < total) n2 <<= 1;
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
        unsigned long long v = ~0ull;
        if (i < total) {
            const uint32_t s = i / k, j = i % k;
            const size_t at = (size_t)q * k + j;
            if (src.keys[s][at] != ~0ull) {
                uint32_t ok = order_bits(src.dists[s][at]);
                if (descending) ok = ~ok;
                v = ((unsigned long long)ok << 32) | i;
            }
        }
        skeys[i] = v;
    }
    // bitonic sort, ascending
    for (uint32_t size = 2; size <= n2; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < n2 / 2; t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = skeys[lo], b = skeys[hi];
                if ((a > b) == up) { skeys[lo] = b; skeys[hi] = a; }
            }
        }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {
        const unsigned long long v = j < n2 ? skeys[j] : ~0ull;
        const size_t dst = (size_t)q * k + j;
        if (v != ~0ull) {
            const uint32_t i = (uint32_t)(v & 0xFFFFFFFFull), s = i / k, jj = i % k;
            const size_t at = (size_t)q * k + jj;
            keys_out[dst] = src.keys[s][at] + src.offset[s];
            dists_out[dst] = src.dists[s][at];
        } else {
            keys_out[dst] = ~0ull;
            dists_out[dst] = descending ? -CUDART_INF_F : CUDART_INF_F;
        }
    }
    if (counts_out && threadIdx.x == 0) {
        uint32_t cnt = 0;
        for (uint32_t j = 0; j < k && j < n2; ++j) cnt += skeys[j] != ~0ull;
        counts_out[q] = cnt;
    }
}

void launch_merge_ptr(const MergeSrc& src, uint32_t n_shards, uint32_t nq, uint32_t k, int descending, uint64_t* keys_out,
                      float* dists_out, uint32_t* counts_out, cudaStream_t stream) {
    uint32_t total = n_shards * k, n2 = 1;
    while (n2 < total) n2 <<= 1;
    const size_t smem = (size_t)n2 * 8;
    if (smem > 200 * 1024) throw Error(LEANN_ERR_INVALID_ARG, "sharded search: n_shards * k too large for the merge kernel");
    if (smem > 48 * 1024)
        LEANN_CUDA_CHECK(cudaFuncSetAttribute(topk_merge_ptr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_ptr_kernel<<<nq, 128, smem, stream>>>(src, n_shards, nq, k, descending, keys_out, dists_out, counts_out);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

struct LocalShard {
    leann_cuda_index* ix = nullptr;
    bool owned = false;
    int device = 0;
    int rank = 0;                 // position in the shard order (= NCCL rank)
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    unsigned char* block = nullptr;      // this shard's packed results: [nq][k] u64 keys, then [nq][k] f32 dists
    unsigned char* gathered = nullptr;   // [world][block_bytes] after ncclAllGather
    size_t cap_block = 0, cap_gathered = 0;
    float* d_queries = nullptr; size_t cap_q = 0;
    uint64_t* d_mask = nullptr; size_t cap_mask = 0;
    uint64_t* d_keys = nullptr; float* d_dists = nullptr; uint32_t* d_counts = nullptr; size_t cap_out = 0, cap_cnt = 0;
};

size_t block_bytes_for(size_t nq, size_t k) { return (nq * k * 12 + 255) & ~(size_t)255; }

template <typename T>
void grow(T*& p, size_t& cap, size_t need) {
    if (cap >= need) return;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    LEANN_CUDA_CHECK(cudaMalloc(&p, need * sizeof(T)));
    cap = need;
}

}  // namespace

struct leann_cuda_shards {
    std::vector<LocalShard> loc;   // shards owned by this process (all of them, or one)
    int world = 0;                 // shards in total
    uint64_t offset[MAX_SHARDS] = {0};   // key offset of every shard
    uint64_t lens[MAX_SHARDS] = {0};
    size_t d = 0;
    int descending = 0;
    bool p2p = false;              // single-process: peer-memory fused merge instead of NCCL
    std::mutex mu;
    uint64_t exchanges = 0, exchange_bytes = 0;
};

namespace {

void free_local(LocalShard& s) {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(s.device);
    if (s.comm) { try { nccl().CommDestroy(s.comm); } catch (...) {} }
    cudaFree(s.block); cudaFree(s.gathered); cudaFree(s.d_queries); cudaFree(s.d_mask);
    cudaFree(s.d_keys); cudaFree(s.d_dists); cudaFree(s.d_counts);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.owned && s.ix) leann_cuda_close(s.ix);
    cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
}

void init_local(LocalShard& s) {
    DeviceScope ds(s.device);
    LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    LEANN_CUDA_CHECK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
}

// Peer access between every pair of devices of a single-process handle.
bool enable_p2p(const std::vector<LocalShard>& loc) {
    for (const LocalShard& a : loc)
        for (const LocalShard& b : loc) {
            if (a.device == b.device) continue;
            int ok = 0;
            if (cudaDeviceCanAccessPeer(&ok, a.device, b.device) != cudaSuccess || !ok) { cudaGetLastError(); return false; }
        }
    for (const LocalShard& a : loc) {
        DeviceScope ds(a.device);
        for (const LocalShard& b : loc) {
            if (a.device == b.device) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
            cudaGetLastError();
        }
    }
    return true;
}

leann_cuda_shards* make_single_process(std::vector<LocalShard>&& loc, const uint64_t* key_offsets, int exchange) {
    std::unique_ptr<leann_cuda_shards> sh(new leann_cuda_shards());
    sh->loc = std::move(loc);
    try {
        sh->world = (int)sh->loc.size();
        if (sh->world < 1 || sh->world > MAX_SHARDS) throw Error(LEANN_ERR_INVALID_ARG, "n_shards must be in 1..16");
        uint64_t run = 0;
        for (int i = 0; i < sh->world; ++i) {
            LocalShard& s = sh->loc[i];
            s.rank = i;
            sh->lens[i] = s.ix->n;
            sh->offset[i] = key_offsets ? key_offsets[i] : run;
            run += s.ix->n;
            if (i == 0) { sh->d = s.ix->d; sh->descending = s.ix->metric == LEANN_METRIC_DOT_DESC; }
            else if (s.ix->d != sh->d || (s.ix->metric == LEANN_METRIC_DOT_DESC) != (sh->descending != 0))
                throw Error(LEANN_ERR_DIM_MISMATCH, "shards disagree on dimensions or metric");
            for (int j = 0; j < i; ++j)
                if (sh->loc[j].device == s.device) throw Error(LEANN_ERR_INVALID_ARG, "sharded handle: one shard per device");
            init_local(s);
        }
        // exchange: 0 auto (peer memory when every pair of devices allows it, else NCCL), 1 NCCL, 2 peer memory only
        if (sh->world > 1) {
            if (exchange != 1) sh->p2p = enable_p2p(sh->loc);
            if (exchange == 2 && !sh->p2p) throw Error(LEANN_ERR_CUDA, "peer access between the shard devices is not available");
            if (!sh->p2p) {
                std::vector<int> devs;
                for (const LocalShard& s : sh->loc) devs.push_back(s.device);
                std::vector<ncclComm_t> comms(sh->world);
                LEANN_NCCL_CHECK(nccl().CommInitAll(comms.data(), sh->world, devs.data()));
                for (int i = 0; i < sh->world; ++i) sh->loc[i].comm = comms[i];
            }
        }
    } catch (...) {
        for (LocalShard& s : sh->loc) free_local(s);
        throw;
    }
    return sh.release();
}

// Search on every local shard -> exchange -> merge. Host or device query/result pointers (device form: one local
// shard, everything on the caller's stream).
void shards_search(leann_cuda_shards* sh, const float* queries, bool on_device, size_t nq, size_t k, size_t ef,
                   const uint64_t* const* shard_masks, uint64_t* keys, float* dists, uint32_t* counts, cudaStream_t user_stream) {
    if (nq == 0) return;
    if (!queries || !keys || !dists) throw Error(LEANN_ERR_INVALID_ARG, "null buffer");
    if (k == 0) throw Error(LEANN_ERR_INVALID_ARG, "k must be > 0");
    if (on_device && sh->loc.size() != 1) throw Error(LEANN_ERR_INVALID_ARG, "device-pointer search needs a handle with one local shard");
    std::lock_guard<std::mutex> lk(sh->mu);
    const size_t bb = block_bytes_for(nq, k);
    const size_t L = sh->loc.size();
    // ---- 1. every local shard searches all queries into its packed block ----
    for (size_t l = 0; l < L; ++l) {
        LocalShard& s = sh->loc[l];
        DeviceScope ds(s.device);
        cudaStream_t st = on_device ? user_stream : s.stream;
        grow(s.block, s.cap_block, bb);
        const float* dq = queries;
        if (!on_device) {
            grow(s.d_queries, s.cap_q, nq * sh->d);
            LEANN_CUDA_CHECK(cudaMemcpyAsync(s.d_queries, queries, nq * sh->d * 4, cudaMemcpyHostToDevice, st));
            dq = s.d_queries;
        }
        const uint64_t* dm = nullptr;
        if (shard_masks && shard_masks[l]) {
            if (on_device) dm = shard_masks[l];
            else {
                const size_t words = (s.ix->n + 63) / 64;
                grow(s.d_mask, s.cap_mask, words);
                LEANN_CUDA_CHECK(cudaMemcpyAsync(s.d_mask, shard_masks[l], words * 8, cudaMemcpyHostToDevice, st));
                dm = s.d_mask;
            }
        }
        {
            std::lock_guard<std::mutex> ilk(s.ix->mu);
            backend_search_device(s.ix, dq, nq, k, ef, dm, (uint64_t*)s.block, (float*)(s.block + nq * k * 8), nullptr, st);
        }
        if (!on_device && L > 1) LEANN_CUDA_CHECK(cudaEventRecord(s.done, st));
    }
    // ---- 2. exchange + merge ----
    MergeSrc src{};
    LocalShard& s0 = sh->loc[0];
    cudaStream_t st0 = on_device ? user_stream : s0.stream;
    for (int r = 0; r < sh->world; ++r) src.offset[r] = sh->offset[r];
    if (sh->world == 1) {
        src.keys[0] = (const uint64_t*)s0.block;
        src.dists[0] = (const float*)(s0.block + nq * k * 8);
    } else if (sh->p2p) {
        // one kernel on device 0 gathers (peer loads over NVLink) and merges; it starts when every shard's search is done
        DeviceScope ds(s0.device);
        for (size_t l = 1; l < L; ++l) LEANN_CUDA_CHECK(cudaStreamWaitEvent(st0, sh->loc[l].done, 0));
        for (size_t l = 0; l < L; ++l) {
            src.keys[l] = (const uint64_t*)sh->loc[l].block;
            src.dists[l] = (const float*)(sh->loc[l].block + nq * k * 8);
        }
        sh->exchange_bytes += (uint64_t)(L - 1) * nq * k * 12;
    } else {
        for (size_t l = 0; l < L; ++l) {
            DeviceScope ds(sh->loc[l].device);
            grow(sh->loc[l].gathered, sh->loc[l].cap_gathered, bb * (size_t)sh->world);
        }
        LEANN_NCCL_CHECK(nccl().GroupStart());
        for (size_t l = 0; l < L; ++l) {
            LocalShard& s = sh->loc[l];
            cudaStream_t st = on_device ? user_stream : s.stream;
            ncclResult_t r = nccl().AllGather(s.block, s.gathered, bb, ncclChar, s.comm, st);
            if (r != ncclSuccess) { nccl().GroupEnd(); throw Error(LEANN_ERR_NCCL, std::string("ncclAllGather: ") + nccl().GetErrorString(r)); }
        }
        LEANN_NCCL_CHECK(nccl().GroupEnd());
        for (int r = 0; r < sh->world; ++r) {
            src.keys[r] = (const uint64_t*)(s0.gathered + (size_t)r * bb);
            src.dists[r] = (const float*)(s0.gathered + (size_t)r * bb + nq * k * 8);
        }
        sh->exchange_bytes += (uint64_t)sh->world * bb;
    }
    sh->exchanges++;
    {
        DeviceScope ds(s0.device);
        uint64_t* ok = keys; float* od = dists; uint32_t* oc = counts;
        if (!on_device) {
            if (s0.cap_out < nq * k) {
                cudaFree(s0.d_keys); cudaFree(s0.d_dists);
                s0.d_keys = nullptr; s0.d_dists = nullptr; s0.cap_out = 0;
                LEANN_CUDA_CHECK(cudaMalloc(&s0.d_keys, nq * k * 8));
                LEANN_CUDA_CHECK(cudaMalloc(&s0.d_dists, nq * k * 4));
                s0.cap_out = nq * k;
            }
            grow(s0.d_counts, s0.cap_cnt, nq);
            ok = s0.d_keys; od = s0.d_dists; oc = s0.d_counts;
        }
        launch_merge_ptr(src, (uint32_t)sh->world, (uint32_t)nq, (uint32_t)k, sh->descending, ok, od, oc, st0);
        if (!on_device) {
            LEANN_CUDA_CHECK(cudaMemcpyAsync(keys, ok, nq * k * 8, cudaMemcpyDeviceToHost, st0));
            LEANN_CUDA_CHECK(cudaMemcpyAsync(dists, od, nq * k * 4, cudaMemcpyDeviceToHost, st0));
            if (counts) LEANN_CUDA_CHECK(cudaMemcpyAsync(counts, oc, nq * 4, cudaMemcpyDeviceToHost, st0));
        }
    }
    if (!on_device) {
        // the other ranks of a single-process NCCL exchange also hold the gathered lists; only shard 0's stream carries the answer
        for (size_t l = 0; l < L; ++l) {
            DeviceScope ds(sh->loc[l].device);
            LEANN_CUDA_CHECK(cudaStreamSynchronize(sh->loc[l].stream));
        }
    }
}

}  // namespace

namespace leann {
// ---- internal surface for the sharded hybrid path (text_api.cu) ------------------------------------------------------
void shards_describe(const leann_cuda_shards* sh, int* world, int* rank, int* device, uint64_t* offsets16, size_t* local_shards) {
    *world = sh->world; *rank = sh->loc.empty() ? 0 : sh->loc[0].rank; *device = sh->loc.empty() ? 0 : sh->loc[0].device;
    for (int i = 0; i < MAX_SHARDS; ++i) offsets16[i] = sh->offset[i];
    *local_shards = sh->loc.size();
}
cudaStream_t shards_stream(leann_cuda_shards* sh) { return sh->loc[0].stream; }
// search + exchange + merge of the vector part, device buffers, enqueued on `st` (handles with one local shard)
void shards_vector_search_device(leann_cuda_shards* sh, const float* dq, size_t nq, size_t k, size_t ef, uint64_t* dk, float* dd,
                                 uint32_t* dc, cudaStream_t st) {
    shards_search(sh, dq, true, nq, k, ef, nullptr, dk, dd, dc, st);
}
// recv[r] = rank r's `bytes` (a plain copy for a world of one)
void shards_all_gather(leann_cuda_shards* sh, const void* send, void* recv, size_t bytes, cudaStream_t st) {
    if (sh->world == 1) {
        LEANN_CUDA_CHECK(cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, st));
        return;
    }
    if (sh->loc.size() != 1 || !sh->loc[0].comm) throw Error(LEANN_ERR_INVALID_ARG, "sharded hybrid search needs the process-per-GPU layout (leann_cuda_shards_join)");
    std::lock_guard<std::mutex> lk(sh->mu);
    LEANN_NCCL_CHECK(nccl().AllGather(send, recv, bytes, ncclChar, sh->loc[0].comm, st));
    sh->exchanges++; sh->exchange_bytes += (uint64_t)sh->world * bytes;
}
void shards_merge_lists(const uint64_t* const* keys, const float* const* dists, const uint64_t* offsets, uint32_t g, uint32_t nq, uint32_t k,
                        int descending, uint64_t* ok, float* od, uint32_t* oc, cudaStream_t st) {
    MergeSrc src{};
    for (uint32_t r = 0; r < g; ++r) { src.keys[r] = keys[r]; src.dists[r] = dists[r]; src.offset[r] = offsets[r]; }
    launch_merge_ptr(src, g, nq, k, descending, ok, od, oc, st);
}
}  // namespace leann

extern "C" {

int leann_cuda_shards_open(const char* const* base_paths, size_t n_shards, int backend, size_t dims, int metric,
                           const int* devices, const uint64_t* key_offsets, int exchange, leann_cuda_shards** out,
                           char* err, size_t errlen) {
    GUARD({
        if (!base_paths || !devices || !out) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if (n_shards < 1 || n_shards > (size_t)MAX_SHARDS) throw Error(LEANN_ERR_INVALID_ARG, "n_shards must be in 1..16");
        std::vector<LocalShard> loc(n_shards);
        try {
            for (size_t i = 0; i < n_shards; ++i) {
                char msg[1024];
                leann_cuda_index* ix = nullptr;
                int rc = leann_cuda_open(base_paths[i], backend, dims, metric, devices[i], &ix, msg, sizeof msg);
                if (rc != LEANN_OK) throw Error(rc, std::string("shard ") + std::to_string(i) + ": " + msg);
                loc[i].ix = ix; loc[i].owned = true; loc[i].device = devices[i];
            }
        } catch (...) {
            for (LocalShard& s : loc) if (s.ix) leann_cuda_close(s.ix);
            throw;
        }
        *out = make_single_process(std::move(loc), key_offsets, exchange);
    });
}

int leann_cuda_shards_from_indexes(leann_cuda_index* const* shards, size_t n_shards, const uint64_t* key_offsets,
                                   int take_ownership, int exchange, leann_cuda_shards** out, char* err, size_t errlen) {
    GUARD({
        if (!shards || !out) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if (n_shards < 1 || n_shards > (size_t)MAX_SHARDS) throw Error(LEANN_ERR_INVALID_ARG, "n_shards must be in 1..16");
        std::vector<LocalShard> loc(n_shards);
        for (size_t i = 0; i < n_shards; ++i) {
            if (!shards[i]) throw Error(LEANN_ERR_INVALID_ARG, "null shard");
            loc[i].ix = shards[i]; loc[i].owned = false; loc[i].device = shards[i]->device;
        }
        leann_cuda_shards* sh = make_single_process(std::move(loc), key_offsets, exchange);
        if (take_ownership) for (LocalShard& s : sh->loc) s.owned = true;
        *out = sh;
    });
}

int leann_cuda_comm_unique_id(unsigned char* id, size_t id_bytes, char* err, size_t errlen) {
    GUARD({
        if (!id || id_bytes < sizeof(ncclUniqueId)) throw Error(LEANN_ERR_INVALID_ARG, "unique id buffer must hold 128 bytes");
        ncclUniqueId u;
        LEANN_NCCL_CHECK(nccl().GetUniqueId(&u));
        memcpy(id, &u, sizeof u);
    });
}

int leann_cuda_shards_join(leann_cuda_index* local, int take_ownership, const unsigned char* id, size_t id_bytes, int rank,
                           int n_ranks, uint64_t key_offset, leann_cuda_shards** out, char* err, size_t errlen) {
    GUARD({
        if (!local || !id || !out) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        if (id_bytes < sizeof(ncclUniqueId)) throw Error(LEANN_ERR_INVALID_ARG, "unique id must be 128 bytes");
        if (n_ranks < 1 || n_ranks > MAX_SHARDS || rank < 0 || rank >= n_ranks) throw Error(LEANN_ERR_INVALID_ARG, "rank / n_ranks out of range (max 16 shards)");
        std::unique_ptr<leann_cuda_shards> sh(new leann_cuda_shards());
        sh->loc.resize(1);
        LocalShard& s = sh->loc[0];
        s.ix = local; s.owned = false; s.device = local->device; s.rank = rank;
        sh->world = n_ranks; sh->d = local->d; sh->descending = local->metric == LEANN_METRIC_DOT_DESC;
        try {
            init_local(s);
            DeviceScope ds(s.device);
            if (n_ranks > 1) {
                ncclUniqueId u;
                memcpy(&u, id, sizeof u);
                LEANN_NCCL_CHECK(nccl().CommInitRank(&s.comm, n_ranks, u, rank));
                // every rank learns every shard's key offset and length: one 16-byte all_gather at join time
                uint64_t mine[2] = {key_offset, (uint64_t)local->n};
                uint64_t* d_all = nullptr;
                LEANN_CUDA_CHECK(cudaMalloc(&d_all, (size_t)(n_ranks + 1) * 16));
                LEANN_CUDA_CHECK(cudaMemcpyAsync(d_all + 2 * n_ranks, mine, 16, cudaMemcpyHostToDevice, s.stream));
                ncclResult_t r = nccl().AllGather(d_all + 2 * n_ranks, d_all, 16, ncclChar, s.comm, s.stream);
                std::vector<uint64_t> all(2 * (size_t)n_ranks);
                cudaError_t e = cudaSuccess;
                if (r == ncclSuccess) {
                    e = cudaMemcpyAsync(all.data(), d_all, all.size() * 8, cudaMemcpyDeviceToHost, s.stream);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
                }
                cudaFree(d_all);
                if (r != ncclSuccess) throw Error(LEANN_ERR_NCCL, std::string("ncclAllGather: ") + nccl().GetErrorString(r));
                LEANN_CUDA_CHECK(e);
                for (int i = 0; i < n_ranks; ++i) { sh->offset[i] = all[2 * i]; sh->lens[i] = all[2 * i + 1]; }
            } else {
                sh->offset[0] = key_offset; sh->lens[0] = local->n;
            }
        } catch (...) {
            free_local(s);
            throw;
        }
        s.owned = take_ownership != 0;
        *out = sh.release();
    });
}

size_t leann_cuda_shards_len(const leann_cuda_shards* sh) {
    if (!sh) return 0;
    uint64_t t = 0;
    for (int i = 0; i < sh->world; ++i) t += sh->lens[i];
    return (size_t)t;
}
size_t leann_cuda_shards_count(const leann_cuda_shards* sh) { return sh ? (size_t)sh->world : 0; }
size_t leann_cuda_shards_dims(const leann_cuda_shards* sh) { return sh ? sh->d : 0; }

int leann_cuda_shards_info(const leann_cuda_shards* sh, uint64_t* info4) {
    if (!sh || !info4) return LEANN_ERR_INVALID_ARG;
    info4[0] = (uint64_t)sh->world;
    info4[1] = sh->world == 1 ? 0 : (sh->p2p ? 2 : 1);   // exchange in use: 0 none, 1 NCCL all_gather, 2 peer-memory merge
    info4[2] = sh->exchanges;
    info4[3] = sh->exchange_bytes;
    return LEANN_OK;
}

int leann_cuda_shards_search(leann_cuda_shards* sh, const float* queries, size_t nq, size_t k, size_t ef,
                             const uint64_t* const* shard_masks, uint64_t* keys, float* dists, uint32_t* counts,
                             char* err, size_t errlen) {
    GUARD({
        if (!sh) throw Error(LEANN_ERR_INVALID_ARG, "null handle");
        shards_search(sh, queries, false, nq, k, ef, shard_masks, keys, dists, counts, nullptr);
    });
}

int leann_cuda_shards_search_device(leann_cuda_shards* sh, const float* d_queries, size_t nq, size_t k, size_t ef,
                                    const uint64_t* d_mask_bits, uint64_t* d_keys, float* d_dists, uint32_t* d_counts,
                                    void* cuda_stream, char* err, size_t errlen) {
    GUARD({
        if (!sh) throw Error(LEANN_ERR_INVALID_ARG, "null handle");
        const uint64_t* masks[1] = {d_mask_bits};
        shards_search(sh, d_queries, true, nq, k, ef, d_mask_bits ? masks : nullptr, d_keys, d_dists, d_counts,
                      (cudaStream_t)cuda_stream);
    });
}

void leann_cuda_shards_close(leann_cuda_shards* sh) {
    if (!sh) return;
    for (LocalShard& s : sh->loc) free_local(s);
    delete sh;
}

}  // extern "C"
