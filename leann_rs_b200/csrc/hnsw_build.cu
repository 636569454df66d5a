// hnsw_build.cu — GPU construction of the usearch-format HNSW graph (replaces the sequential
// `index.add(i, v)` loop of hnsw::build_index, leann-rs src/backend/hnsw.rs:96-139, and writes the
// same `.index` file through leann_cuda_save). SURVEY.md §8(f) N1.
//
// Batched insertion: nodes are added in slot order in batches that grow with the graph
// (batch <= size/16). Per batch, three stream-ordered kernels:
//   A  insert_search : one warp per new node — greedy descent, then per level the same beam search
//                      as K1 with ef = expansion_add, usearch's neighbour-selection heuristic
//                      (refine_) down to M links, forward links written, reverse edges emitted;
//   B  reverse_append: one thread per reverse edge — claims a slot in the target list with an
//                      atomic counter; overflowing lists are queued for pruning;
//   C  prune         : one warp per overflowing list — re-ranks (existing + incoming) by distance to
//                      the owner and re-applies the heuristic down to the level's capacity.
// Nodes of one batch do not see each other (they are linked by later batches); everything else
// follows usearch index_gt::add. Level assignment: floor(-ln(U) / ln(M)) as in usearch.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <random>

#include "graph_device.cuh"

namespace leann {

namespace {

constexpr int EXTRA = 16;        // incoming links buffered per list and batch beyond its capacity
constexpr int STAGE = 160;       // staging entries per warp in the builder kernels (>= MAX_DEG + EXTRA)

struct BuildParams {
    uint32_t* adj0; uint32_t* adjU; const uint32_t* upper_base; const uint8_t* levels;
    uint32_t* cnt;      // [n + n_upper] entries per list (may exceed capacity until pruned)
    uint32_t* flag;     // [n + n_upper] queued-for-prune marker
    uint32_t* extra;    // [n + n_upper][EXTRA]
    uint2* edges;       // (target list owner c, new node p | level << 28 in .y? no: see below)
    uint32_t* edge_level;
    uint32_t* edge_count; uint32_t max_edges;
    uint2* work; uint32_t* work_count; uint32_t* work_cursor;
    uint32_t first, count;   // batch = slots [first, first+count)
    uint32_t M, M0, ef_add, n;
    uint32_t next_cap, next_capp;
    uint8_t* visited; uint32_t* epochs; uint32_t* counter; size_t n_pad; int n_warps;
    uint32_t* vhash; uint32_t vhash_cap; uint32_t* pool_locks; uint32_t pool_slots;   // large-index mode (see VisitedSet)
    uint32_t* overflow;  // [0]: edge buffer overflow, [1]: extra overflow (dropped links)
    const uint32_t* order;  // nullable: insertion sequence (node = order[first + i]); identity when null
    int vamana; float alpha;
};

__device__ __forceinline__ void carve(unsigned char* base, uint32_t top_cap, uint32_t next_capp, WarpLists& w) {
    w.top_d = reinterpret_cast<float*>(base);
    w.top_s = reinterpret_cast<uint32_t*>(base + (size_t)top_cap * 4);
    w.next_d = reinterpret_cast<float*>(base + (size_t)top_cap * 8);
    w.next_s = reinterpret_cast<uint32_t*>(base + (size_t)top_cap * 8 + (size_t)next_capp * 4);
    w.st_slot = reinterpret_cast<uint32_t*>(base + (size_t)top_cap * 8 + (size_t)next_capp * 8);
    w.st_dist = reinterpret_cast<float*>(base + (size_t)top_cap * 8 + (size_t)next_capp * 8 + (size_t)STAGE * 4);
}
__host__ __device__ inline size_t build_smem_per_warp(uint32_t top_cap, uint32_t next_capp) {
    return (size_t)top_cap * 8 + (size_t)next_capp * 8 + (size_t)STAGE * 8;
}

// usearch refine_: `top` ascending by distance to the owner; keeps candidate x unless an already
// kept neighbour is closer to x than the owner is. Returns the new size.
// vamana != 0 switches to DiskANN's alpha-robust prune (drop x when alpha * d(kept, x) <= d(owner, x);
// always applied), restating diskann-rs build_index_with_params (diskann.rs:87-100, alpha = 1.2).
template <int LPV, int VPL, int U>
__device__ __forceinline__ int refine_heuristic(const GraphView& g, WarpLists& w, int needed, int lane, int vamana = 0,
                                                float alpha = 1.0f) {
    int total = w.top_size;
    if (total == 0) return 0;
    if (!vamana && total < needed) return total;
    int submitted = 1;
    for (int consumed = 1; consumed < total && submitted < needed; ++consumed) {
        uint32_t x = w.top_s[consumed];
        float xd = w.top_d[consumed];
        float4 qx[VPL];
        {
            const int lig = lane % LPV;
            const float4* row = g.vecs + (size_t)x * g.d4;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                uint32_t idx = (uint32_t)(i * LPV + lig);
                qx[i] = idx < g.d4 ? __ldg(row + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        bool good = true;
        // check kept neighbours in chunks of 8 (closest first): most rejections happen early
        for (int b = 0; b < submitted && good; b += 8) {
            int c = submitted - b < 8 ? submitted - b : 8;
            eval_distances<LPV, VPL, U>(g.vecs, g.d4, g.metric, qx, w.top_s + b, w.st_dist, c, lane);
            bool closer = lane < c && (vamana ? (alpha * w.st_dist[lane] <= xd) : (w.st_dist[lane] < xd));
            if (__any_sync(FULL, closer)) good = false;
            __syncwarp();
        }
        if (good) {
            if (lane == 0) { w.top_d[submitted] = xd; w.top_s[submitted] = x; }
            submitted++;
        }
        __syncwarp();
    }
    return submitted;
}

template <int LPV, int VPL, int U>
__global__ void __launch_bounds__(128)
insert_search_kernel(const GraphView g, const BuildParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + wib;
    if (warp_global >= p.n_warps) return;
    const uint32_t top_cap = (p.ef_add + 31u) & ~31u;
    WarpLists w;
    carve(smem_raw + build_smem_per_warp(top_cap, p.next_capp) * wib, top_cap, p.next_capp, w);
    VisitedSet vs{};
    vs.n_pad = p.n_pad;
    vs.slot = -1;
    if (p.vhash) {   // large-index mode: per-warp hash table, pooled byte maps for the spill
        vs.tbl = p.vhash + (size_t)warp_global * p.vhash_cap;
        vs.cap_mask = p.vhash_cap - 1u;
        vs.shift = 32u - (uint32_t)__ffs((int)p.vhash_cap) + 1u;
        vs.limit = p.vhash_cap / 4u * 3u;
        vs.pool_vis = p.visited; vs.pool_epochs = p.epochs; vs.pool_locks = p.pool_locks; vs.n_slots = p.pool_slots;
    } else {
        vs.vis = p.visited + (size_t)warp_global * p.n_pad;
        vs.epoch_slot = p.epochs + warp_global;
    }

    for (;;) {
        uint32_t bi = 0;
        if (lane == 0) bi = atomicAdd(p.counter, 1u);
        bi = __shfl_sync(FULL, bi, 0);
        if (bi >= p.count) break;
        const uint32_t node = p.order ? p.order[p.first + bi] : p.first + bi;
        const int node_level = p.levels ? p.levels[node] : 0;
        float4 q[VPL];
        {
            const int lig = lane % LPV;
            const float4* row = g.vecs + (size_t)node * g.d4;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                uint32_t idx = (uint32_t)(i * LPV + lig);
                q[i] = idx < g.d4 ? __ldg(row + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        Counters c{0u, 0u, 0u, 0u};
        uint32_t cur = g.entry;
        if (lane == 0) w.st_slot[0] = cur;
        __syncwarp();
        eval_distances<LPV, VPL, U>(g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, 1, lane);
        float cur_d = w.st_dist[0];
        __syncwarp();
        if (g.max_level > node_level) greedy_descend<LPV, VPL, U>(g, q, w, cur, cur_d, g.max_level, node_level, c, lane);
        const int top_level = node_level < g.max_level ? node_level : g.max_level;
        for (int level = top_level; level >= 0; --level) {
            visited_begin(vs, lane);
            LevelAdj adj{g.adj0, g.adjU, g.upper_base, level == 0 ? g.deg0 : g.degU, level};
            beam_level<LPV, VPL, U>(g, adj, q, w, (int)p.ef_add, (int)p.next_cap, (int)p.next_capp - 1, 0, nullptr,
                                    vs, (uint32_t)warp_global, cur, cur_d, c, lane);
            visited_end(vs, lane);
            int kept = refine_heuristic<LPV, VPL, U>(g, w, (int)p.M, lane, p.vamana, p.alpha);
            // forward links of the new node (its rows are pre-filled with SENT)
            uint32_t* myrow = level == 0 ? p.adj0 + (size_t)node * p.M0
                                         : p.adjU + ((size_t)p.upper_base[node] + (uint32_t)(level - 1)) * p.M;
            const uint32_t list_id = level == 0 ? node : p.n + p.upper_base[node] + (uint32_t)(level - 1);
            uint32_t ebase = 0;
            if (lane == 0) {
                p.cnt[list_id] = (uint32_t)kept;
                ebase = atomicAdd(p.edge_count, (uint32_t)kept);
            }
            ebase = __shfl_sync(FULL, ebase, 0);
            for (int j = lane; j < kept; j += 32) {
                uint32_t nb = w.top_s[j];
                myrow[j] = nb;
                if (ebase + j < p.max_edges) {
                    p.edges[ebase + j] = make_uint2(nb, node);
                    p.edge_level[ebase + j] = (uint32_t)level;
                } else {
                    p.overflow[0] = 1u;
                }
            }
            cur = w.top_s[0];
            cur_d = w.top_d[0];
            __syncwarp();
        }
    }
}

__global__ void reverse_append_kernel(const BuildParams p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t ne = *p.edge_count;
    if (ne > p.max_edges) ne = p.max_edges;
    if (i >= ne) return;
    uint2 e = p.edges[i];
    uint32_t c = e.x, node = e.y, level = p.edge_level[i];
    uint32_t list_id = level == 0 ? c : p.n + p.upper_base[c] + (level - 1);
    uint32_t cap = level == 0 ? p.M0 : p.M;
    uint32_t* row = level == 0 ? p.adj0 + (size_t)c * p.M0 : p.adjU + ((size_t)p.upper_base[c] + (level - 1)) * p.M;
    uint32_t slot = atomicAdd(&p.cnt[list_id], 1u);
    if (slot < cap) {
        row[slot] = node;
    } else {
        uint32_t x = slot - cap;
        if (x < (uint32_t)EXTRA) p.extra[(size_t)list_id * EXTRA + x] = node;
        else p.overflow[1] = 1u;
        if (atomicExch(&p.flag[list_id], 1u) == 0u) {
            uint32_t wpos = atomicAdd(p.work_count, 1u);
            p.work[wpos] = make_uint2(c, level);
        }
    }
}

template <int LPV, int VPL, int U>
__global__ void __launch_bounds__(128)
prune_kernel(const GraphView g, const BuildParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    WarpLists w;
    carve(smem_raw + build_smem_per_warp(STAGE, 32) * wib, STAGE, 32, w);
    const uint32_t nwork = *p.work_count;
    for (;;) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(p.work_cursor, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= nwork) break;
        uint2 it = p.work[wi];
        const uint32_t c = it.x, level = it.y;
        const uint32_t list_id = level == 0 ? c : p.n + p.upper_base[c] + (level - 1);
        const uint32_t cap = level == 0 ? p.M0 : p.M;
        uint32_t* row = level == 0 ? p.adj0 + (size_t)c * p.M0 : p.adjU + ((size_t)p.upper_base[c] + (level - 1)) * p.M;
        uint32_t total = p.cnt[list_id];
        uint32_t n_ext = total > cap ? total - cap : 0;
        if (n_ext > (uint32_t)EXTRA) n_ext = EXTRA;
        const int m = (int)(cap + n_ext);
        for (int j = lane; j < m; j += 32)
            w.st_slot[j] = j < (int)cap ? row[j] : p.extra[(size_t)list_id * EXTRA + (j - cap)];
        float4 q[VPL];
        {
            const int lig = lane % LPV;
            const float4* vr = g.vecs + (size_t)c * g.d4;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                uint32_t idx = (uint32_t)(i * LPV + lig);
                q[i] = idx < g.d4 ? __ldg(vr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncwarp();
        eval_distances<LPV, VPL, U>(g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, m, lane);
        w.top_size = 0;
        for (int j = 0; j < m; ++j) {
            float dj = w.st_dist[j];
            uint32_t sj = w.st_slot[j];
            __syncwarp();
            sorted_insert<false, false>(w.top_d, w.top_s, w.top_size, STAGE, 0, 0, dj, sj, lane);
        }
        int kept = refine_heuristic<LPV, VPL, U>(g, w, (int)cap, lane, p.vamana, p.alpha);
        for (int j = lane; j < (int)cap; j += 32) row[j] = j < kept ? w.top_s[j] : SENT;
        if (lane == 0) { p.cnt[list_id] = (uint32_t)kept; p.flag[list_id] = 0u; }
        __syncwarp();
    }
}

template <typename F>
void dispatch_dims(uint32_t d, uint32_t d4, F&& f) {
    if (reduction_lanes(d) == 8) {
        uint32_t vpl = (d4 + 7) / 8;
        if (vpl <= 2) f.template operator()<8, 2, 4>();
        else if (vpl <= 3) f.template operator()<8, 3, 4>();
        else if (vpl <= 4) f.template operator()<8, 4, 4>();
        else f.template operator()<8, 8, 2>();
    } else {
        uint32_t vpl = (d4 + 31) / 32;
        if (vpl <= 3) f.template operator()<32, 3, 8>();
        else if (vpl <= 4) f.template operator()<32, 4, 4>();
        else if (vpl <= 6) f.template operator()<32, 6, 4>();
        else if (vpl <= 8) f.template operator()<32, 8, 2>();
        else if (vpl <= 12) f.template operator()<32, 12, 2>();
        else if (vpl <= 16) f.template operator()<32, 16, 1>();
        else if (vpl <= 32) f.template operator()<32, 32, 1>();
        else throw Error(LEANN_ERR_INVALID_ARG, "dimension above 4096 is not supported");
    }
}

struct LaunchA {
    GraphView g; BuildParams p; cudaStream_t s;
    template <int LPV, int VPL, int U> void operator()() {
        uint32_t top_cap = (p.ef_add + 31u) & ~31u;
        size_t smem = build_smem_per_warp(top_cap, p.next_capp) * 4;
        auto k = insert_search_kernel<LPV, VPL, U>;
        if (smem > 48 * 1024) LEANN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(p.n_warps + 3) / 4, 128, smem, s>>>(g, p);
        LEANN_CUDA_CHECK(cudaGetLastError());
    }
};
struct LaunchC {
    GraphView g; BuildParams p; cudaStream_t s; int blocks;
    template <int LPV, int VPL, int U> void operator()() {
        size_t smem = build_smem_per_warp(STAGE, 32) * 4;
        prune_kernel<LPV, VPL, U><<<blocks, 128, smem, s>>>(g, p);
        LEANN_CUDA_CHECK(cudaGetLastError());
    }
};

template <typename T>
T* dmalloc(size_t count) {
    T* p = nullptr;
    LEANN_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    return p;
}

// Visited workspace of the builders: one byte map per warp when that fits in a quarter of the free memory, else per-warp
// hash tables of 2 * ef_add * degree ids with a pool of byte maps for the spill (the traversal's large-index mode).
void alloc_build_visited(BuildParams& p, int& max_warps, uint32_t MAXB, size_t n, uint32_t ef_add, uint32_t deg, std::vector<void*>& temps,
                         cudaStream_t stream) {
    p.n_pad = (n + 127) & ~(size_t)127;
    size_t free_b = 0, total_b = 0;
    LEANN_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    max_warps = (int)std::min<size_t>((size_t)max_warps, (size_t)MAXB);
    p.vhash = nullptr; p.vhash_cap = 1024; p.pool_locks = nullptr; p.pool_slots = 1;
    const bool force_large = getenv("LEANN_CUDA_FORCE_LARGE_INDEX_MODE") != nullptr;   // test hook: exercise the large-index mode on small data
    if (!force_large && (size_t)max_warps * p.n_pad <= free_b / 4) {
        p.visited = dmalloc<uint8_t>((size_t)max_warps * p.n_pad); temps.push_back(p.visited);
        p.epochs = dmalloc<uint32_t>(max_warps); temps.push_back(p.epochs);
        LEANN_CUDA_CHECK(cudaMemsetAsync(p.visited, 0, (size_t)max_warps * p.n_pad, stream));
        LEANN_CUDA_CHECK(cudaMemsetAsync(p.epochs, 0, (size_t)max_warps * 4, stream));
        return;
    }
    uint32_t cap = 1024;
    while (cap < 2u * ef_add * deg && cap < (1u << 22)) cap <<= 1;
    const uint32_t slots = (uint32_t)std::min<size_t>(64, std::max<size_t>(1, free_b / 8 / p.n_pad));
    p.vhash = dmalloc<uint32_t>((size_t)max_warps * cap); temps.push_back(p.vhash);
    p.vhash_cap = cap;
    p.visited = dmalloc<uint8_t>((size_t)slots * p.n_pad); temps.push_back(p.visited);
    p.epochs = dmalloc<uint32_t>(slots); temps.push_back(p.epochs);
    p.pool_locks = dmalloc<uint32_t>(slots); temps.push_back(p.pool_locks);
    p.pool_slots = slots;
    LEANN_CUDA_CHECK(cudaMemsetAsync(p.visited, 0, (size_t)slots * p.n_pad, stream));
    LEANN_CUDA_CHECK(cudaMemsetAsync(p.epochs, 0, (size_t)slots * 4, stream));
    LEANN_CUDA_CHECK(cudaMemsetAsync(p.pool_locks, 0, (size_t)slots * 4, stream));
}

}  // namespace

namespace {

// entries per list of an existing graph (lists are packed at the front and SENT-padded)
__global__ void count_lists_kernel(const uint32_t* __restrict__ adj0, const uint32_t* __restrict__ adjU, uint32_t n, uint32_t n_upper,
                                   uint32_t M, uint32_t M0, uint32_t* __restrict__ cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + n_upper) return;
    const uint32_t cap = i < n ? M0 : M;
    const uint32_t* row = i < n ? adj0 + (size_t)i * M0 : adjU + (size_t)(i - n) * M;
    uint32_t c = 0;
    while (c < cap && row[c] != SENT) ++c;
    cnt[i] = c;
}

// usearch choose_random_level_ = floor(-ln(U) * 1/ln(connectivity)) for slots [first, first + count)
void draw_levels(size_t count, size_t M, uint64_t seed, std::vector<uint8_t>& out) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> uni(0.0, 1.0);
    const double inv_log = 1.0 / std::log((double)M);
    out.resize(count);
    for (size_t i = 0; i < count; ++i) {
        double u = uni(rng);
        if (u <= 0.0) u = 1e-300;
        int l = (int)(-std::log(u) * inv_log);
        out[i] = (uint8_t)(l > 15 ? 15 : l);
    }
}

// Inserts slots [first, n) into the graph held by `ix` (arrays sized for n nodes, rows of the new slots
// SENT-filled, ix->entry / max_level describing the graph over [0, first)). Batched insertion as described
// at the top of this file; `lv8` holds the level of every slot.
void insert_range(leann_cuda_index* ix, const std::vector<uint8_t>& lv8, size_t first, size_t ef_add) {
    const size_t n = ix->n, M = ix->M, M0 = ix->M0, n_upper = ix->n_upper_lists;
    if (first >= n) return;
    cudaStream_t stream = nullptr;
    LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    std::vector<void*> temps;
    auto cleanup = [&]() { for (void* t : temps) cudaFree(t); cudaStreamDestroy(stream); };
    try {
        const uint32_t MAXB = 16384;
        const size_t n_lists = n + n_upper;
        BuildParams p{};
        p.adj0 = ix->adj0; p.adjU = ix->adjU; p.upper_base = ix->upper_base;
        uint8_t* d_levels = dmalloc<uint8_t>(n); temps.push_back(d_levels);
        LEANN_CUDA_CHECK(cudaMemcpy(d_levels, lv8.data(), n, cudaMemcpyHostToDevice));
        p.levels = d_levels;
        p.cnt = dmalloc<uint32_t>(n_lists); temps.push_back(p.cnt);
        p.flag = dmalloc<uint32_t>(n_lists); temps.push_back(p.flag);
        p.extra = dmalloc<uint32_t>(n_lists * EXTRA); temps.push_back(p.extra);
        count_lists_kernel<<<(unsigned)((n_lists + 255) / 256), 256, 0, stream>>>(ix->adj0, ix->adjU, (uint32_t)n, (uint32_t)n_upper,
                                                                                (uint32_t)M, (uint32_t)M0, p.cnt);
        LEANN_CUDA_CHECK(cudaGetLastError());
        LEANN_CUDA_CHECK(cudaMemsetAsync(p.flag, 0, n_lists * 4, stream));
        p.max_edges = (uint32_t)std::min<size_t>((size_t)MAXB * M * 3, (size_t)0x7FFFFFFF);
        p.edges = dmalloc<uint2>(p.max_edges); temps.push_back(p.edges);
        p.edge_level = dmalloc<uint32_t>(p.max_edges); temps.push_back(p.edge_level);
        p.work = dmalloc<uint2>(p.max_edges); temps.push_back(p.work);
        uint32_t* ctrs = dmalloc<uint32_t>(8); temps.push_back(ctrs);
        LEANN_CUDA_CHECK(cudaMemsetAsync(ctrs, 0, 32, stream));
        p.edge_count = ctrs; p.work_count = ctrs + 1; p.work_cursor = ctrs + 2; p.counter = ctrs + 3; p.overflow = ctrs + 4;
        p.M = (uint32_t)M; p.M0 = (uint32_t)M0; p.ef_add = (uint32_t)ef_add; p.n = (uint32_t)n;
        p.next_cap = (uint32_t)ef_add;
        p.next_capp = 1; while (p.next_capp < p.next_cap) p.next_capp <<= 1;
        p.order = nullptr; p.vamana = 0; p.alpha = 1.0f;
        int max_warps = graph_search_max_warps(ix->device);
        alloc_build_visited(p, max_warps, MAXB, n, p.ef_add, (uint32_t)M0, temps, stream);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);

        size_t inserted = first;
        while (inserted < n) {
            uint32_t b = (uint32_t)std::min<size_t>(std::min<size_t>(n - inserted, MAXB), std::max<size_t>(1, inserted / 16));
            p.first = (uint32_t)inserted; p.count = b;
            p.n_warps = (int)std::min<uint32_t>((uint32_t)max_warps, (b + 3u) & ~3u);
            GraphView g = ix->view();
            LEANN_CUDA_CHECK(cudaMemsetAsync(ctrs, 0, 16, stream));  // edge_count, work_count, work_cursor, counter
            dispatch_dims((uint32_t)ix->d, ix->d4, LaunchA{g, p, stream});
            uint32_t max_e = std::min<uint32_t>(p.max_edges, b * (uint32_t)M * 3u);
            reverse_append_kernel<<<(max_e + 255) / 256, 256, 0, stream>>>(p);
            LEANN_CUDA_CHECK(cudaGetLastError());
            int pblocks = (int)std::min<uint32_t>((uint32_t)sms * 3u, (max_e + 3u) / 4u);
            dispatch_dims((uint32_t)ix->d, ix->d4, LaunchC{g, p, stream, std::max(pblocks, 1)});
            // entry point / top level follow the sequential rule of index_gt::add
            for (size_t i = inserted; i < inserted + b; ++i)
                if ((int)lv8[i] > ix->max_level) { ix->max_level = lv8[i]; ix->entry = (uint32_t)i; }
            inserted += b;
        }
        uint32_t h_over[2] = {0, 0};
        LEANN_CUDA_CHECK(cudaMemcpyAsync(h_over, p.overflow, 8, cudaMemcpyDeviceToHost, stream));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (h_over[0]) throw Error(LEANN_ERR_CUDA, "hnsw build: reverse-edge buffer overflow");
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

// cudaMalloc a larger array, keep the first `old_count` elements, fill the rest with the byte `fill`.
template <typename T>
void grow(T*& ptr, size_t old_count, size_t new_count, int fill) {
    T* np = dmalloc<T>(new_count);
    if (old_count) LEANN_CUDA_CHECK(cudaMemcpy(np, ptr, old_count * sizeof(T), cudaMemcpyDeviceToDevice));
    if (new_count > old_count) LEANN_CUDA_CHECK(cudaMemset(np + old_count, fill, (new_count - old_count) * sizeof(T)));
    cudaFree(ptr);
    ptr = np;
}

}  // namespace

void gpu_hnsw_build(leann_cuda_index* ix, size_t M, size_t ef_add, uint64_t seed) {
    const size_t n = ix->n, M0 = 2 * M;
    ix->M = (uint32_t)M; ix->M0 = (uint32_t)M0;
    ix->identity_keys = true;
    std::vector<uint8_t> lv8;
    draw_levels(n, M, seed, lv8);
    ix->h_levels.resize(n);
    std::vector<uint32_t> upper_base(n);
    size_t n_upper = 0;
    for (size_t i = 0; i < n; ++i) {
        ix->h_levels[i] = (int16_t)lv8[i];
        upper_base[i] = (uint32_t)n_upper;
        n_upper += (size_t)lv8[i];
    }
    ix->n_upper_lists = n_upper;
    ix->adj0 = dmalloc<uint32_t>(n * M0);
    ix->adjU = dmalloc<uint32_t>(n_upper * M);
    ix->upper_base = dmalloc<uint32_t>(n);
    ix->keys = dmalloc<uint64_t>(n);
    LEANN_CUDA_CHECK(cudaMemset(ix->adj0, 0xFF, std::max<size_t>(n * M0, 1) * 4));
    LEANN_CUDA_CHECK(cudaMemset(ix->adjU, 0xFF, std::max<size_t>(n_upper * M, 1) * 4));
    if (n) LEANN_CUDA_CHECK(cudaMemcpy(ix->upper_base, upper_base.data(), n * 4, cudaMemcpyHostToDevice));
    {
        std::vector<uint64_t> keys(n);
        for (size_t i = 0; i < n; ++i) keys[i] = i;
        if (n) LEANN_CUDA_CHECK(cudaMemcpy(ix->keys, keys.data(), n * 8, cudaMemcpyHostToDevice));
    }
    ix->entry = 0;
    ix->max_level = n ? lv8[0] : 0;
    insert_range(ix, lv8, 1, ef_add);
}

// hnsw::add_to_index (leann-rs src/backend/hnsw.rs:142-191): `m` more vectors (already padded
// device rows) are appended to a resident index with keys start_id + i, as `index.add(id, embedding)` does after
// `index.load`. Connectivity comes from the index; the new slots draw their levels from `seed`.
void gpu_hnsw_add(leann_cuda_index* ix, const float4* new_rows, size_t m, uint64_t start_id, size_t ef_add, uint64_t seed) {
    if (m == 0) return;
    const size_t n_old = ix->n, n_new = n_old + m, M = ix->M, M0 = ix->M0;
    if (n_new > 0xFFFFFFF0ull) throw Error(LEANN_ERR_INVALID_ARG, "index would exceed 2^32 slots");
    std::vector<uint8_t> lv_new;
    draw_levels(m, M, seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(n_old + 1)), lv_new);
    std::vector<uint8_t> lv8(n_new);
    for (size_t i = 0; i < n_old; ++i) lv8[i] = (uint8_t)std::min<int>(15, std::max<int>(0, ix->h_levels[i]));
    std::vector<uint32_t> ub_new(m);
    size_t n_upper = ix->n_upper_lists;
    for (size_t i = 0; i < m; ++i) {
        lv8[n_old + i] = lv_new[i];
        ub_new[i] = (uint32_t)n_upper;
        n_upper += (size_t)lv_new[i];
    }
    grow(ix->vecs, n_old * ix->d4, n_new * ix->d4, 0);
    LEANN_CUDA_CHECK(cudaMemcpy(ix->vecs + n_old * ix->d4, new_rows, m * (size_t)ix->d4 * 16, cudaMemcpyDeviceToDevice));
    grow(ix->adj0, n_old * M0, n_new * M0, 0xFF);
    grow(ix->adjU, ix->n_upper_lists * M, n_upper * M, 0xFF);
    grow(ix->upper_base, n_old, n_new, 0);
    LEANN_CUDA_CHECK(cudaMemcpy(ix->upper_base + n_old, ub_new.data(), m * 4, cudaMemcpyHostToDevice));
    {
        std::vector<uint64_t> keys(m);
        for (size_t i = 0; i < m; ++i) keys[i] = start_id + i;
        if (!ix->keys) {   // indexes opened with key == slot keep no key array
            std::vector<uint64_t> old(n_old);
            for (size_t i = 0; i < n_old; ++i) old[i] = i;
            ix->keys = dmalloc<uint64_t>(n_old);
            if (n_old) LEANN_CUDA_CHECK(cudaMemcpy(ix->keys, old.data(), n_old * 8, cudaMemcpyHostToDevice));
        }
        grow(ix->keys, n_old, n_new, 0);
        LEANN_CUDA_CHECK(cudaMemcpy(ix->keys + n_old, keys.data(), m * 8, cudaMemcpyHostToDevice));
    }
    if (start_id != n_old) ix->identity_keys = false;
    ix->h_levels.resize(n_new);
    for (size_t i = 0; i < m; ++i) ix->h_levels[n_old + i] = (int16_t)lv_new[i];
    ix->n = n_new;
    ix->n_upper_lists = n_upper;
    size_t first = n_old;
    if (n_old == 0) { ix->entry = 0; ix->max_level = lv8[0]; first = 1; }
    insert_range(ix, lv8, first, ef_add);
}

namespace {

__global__ void colsum_kernel(const float4* __restrict__ vecs, uint32_t d4, size_t n, float* __restrict__ sums) {
    // grid.x = column groups of float4, grid.y = row slices
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d4) return;
    size_t rows_per = (n + gridDim.y - 1) / gridDim.y;
    size_t r0 = (size_t)blockIdx.y * rows_per, r1 = r0 + rows_per < n ? r0 + rows_per : n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t r = r0; r < r1; ++r) {
        float4 v = vecs[r * d4 + c];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(&sums[c * 4 + 0], acc.x); atomicAdd(&sums[c * 4 + 1], acc.y);
    atomicAdd(&sums[c * 4 + 2], acc.z); atomicAdd(&sums[c * 4 + 3], acc.w);
}

// medoid = row nearest (L2) to the centroid; packed (ordered distance, row) atomicMin.
__global__ void medoid_kernel(const float4* __restrict__ vecs, uint32_t d4, size_t n, const float* __restrict__ sums,
                              unsigned long long* __restrict__ best) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float inv = 1.0f / (float)n;
    float s = 0.f;
    for (uint32_t i = lane; i < d4; i += 32) {
        float4 v = vecs[row * d4 + i];
        float a = v.x - sums[i * 4] * inv, b = v.y - sums[i * 4 + 1] * inv, c = v.z - sums[i * 4 + 2] * inv, d = v.w - sums[i * 4 + 3] * inv;
        s += a * a + b * b + c * c + d * d;
    }
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
    if (lane == 0) atomicMin(best, ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)row);
}

}  // namespace

// Vamana construction (diskann-rs build_index_with_params, diskann.rs:70-105): medoid entry point,
// greedy search with beam L = `complexity`, alpha-robust prune to R, reverse edges with re-prune.
// Batched insertion (ParlayANN-style prefix doubling) replaces the rayon-parallel passes of the crate;
// the file written by leann_cuda_save is the same `.diskann` layout.
void gpu_vamana_build(leann_cuda_index* ix, size_t R, size_t L, float alpha, uint64_t seed) {
    (void)seed;
    const size_t n = ix->n;
    ix->M = (uint32_t)R; ix->M0 = (uint32_t)R;
    ix->max_level = 0; ix->entry = 0; ix->identity_keys = true; ix->n_upper_lists = 0;
    cudaStream_t stream = nullptr;
    LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    std::vector<void*> temps;
    auto cleanup = [&]() { for (void* t : temps) cudaFree(t); cudaStreamDestroy(stream); };
    try {
        ix->adj0 = dmalloc<uint32_t>(n * R);
        LEANN_CUDA_CHECK(cudaMemset(ix->adj0, 0xFF, std::max<size_t>(n * R, 1) * 4));
        if (n <= 1) { cleanup(); return; }
        // ---- medoid ----
        float* sums = dmalloc<float>(ix->d4 * 4); temps.push_back(sums);
        unsigned long long* best = dmalloc<unsigned long long>(1); temps.push_back(best);
        LEANN_CUDA_CHECK(cudaMemsetAsync(sums, 0, (size_t)ix->d4 * 16, stream));
        LEANN_CUDA_CHECK(cudaMemsetAsync(best, 0xFF, 8, stream));
        dim3 cg((ix->d4 + 63) / 64, (unsigned)std::min<size_t>(1024, (n + 255) / 256));
        colsum_kernel<<<cg, 64, 0, stream>>>(ix->vecs, ix->d4, n, sums);
        medoid_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(ix->vecs, ix->d4, n, sums, best);
        unsigned long long h_best = 0;
        LEANN_CUDA_CHECK(cudaMemcpyAsync(&h_best, best, 8, cudaMemcpyDeviceToHost, stream));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(stream));
        const uint32_t medoid = (uint32_t)(h_best & 0xFFFFFFFFull);
        // insertion order: medoid first, then slot order
        std::vector<uint32_t> order(n);
        order[0] = medoid;
        for (size_t i = 0, o = 1; i < n; ++i) if (i != medoid) order[o++] = (uint32_t)i;
        uint32_t* d_order = dmalloc<uint32_t>(n); temps.push_back(d_order);
        LEANN_CUDA_CHECK(cudaMemcpy(d_order, order.data(), n * 4, cudaMemcpyHostToDevice));

        const uint32_t MAXB = 16384;
        BuildParams p{};
        p.adj0 = ix->adj0; p.adjU = nullptr; p.upper_base = nullptr; p.levels = nullptr;
        p.cnt = dmalloc<uint32_t>(n); temps.push_back(p.cnt);
        p.flag = dmalloc<uint32_t>(n); temps.push_back(p.flag);
        p.extra = dmalloc<uint32_t>(n * EXTRA); temps.push_back(p.extra);
        LEANN_CUDA_CHECK(cudaMemset(p.cnt, 0, n * 4));
        LEANN_CUDA_CHECK(cudaMemset(p.flag, 0, n * 4));
        p.max_edges = (uint32_t)std::min<size_t>((size_t)MAXB * R + 1024, (size_t)0x7FFFFFFF);
        p.edges = dmalloc<uint2>(p.max_edges); temps.push_back(p.edges);
        p.edge_level = dmalloc<uint32_t>(p.max_edges); temps.push_back(p.edge_level);
        p.work = dmalloc<uint2>(p.max_edges); temps.push_back(p.work);
        uint32_t* ctrs = dmalloc<uint32_t>(8); temps.push_back(ctrs);
        LEANN_CUDA_CHECK(cudaMemset(ctrs, 0, 32));
        p.edge_count = ctrs; p.work_count = ctrs + 1; p.work_cursor = ctrs + 2; p.counter = ctrs + 3; p.overflow = ctrs + 4;
        p.M = (uint32_t)R; p.M0 = (uint32_t)R; p.ef_add = (uint32_t)std::max(L, R); p.n = (uint32_t)n;
        p.next_cap = p.ef_add;
        p.next_capp = 1; while (p.next_capp < p.next_cap) p.next_capp <<= 1;
        p.order = d_order; p.vamana = 1; p.alpha = alpha;
        int max_warps = graph_search_max_warps(ix->device);
        alloc_build_visited(p, max_warps, MAXB, n, p.ef_add, (uint32_t)R, temps, stream);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
        ix->entry = medoid;  // searches during the build start from the medoid, as they will afterwards
        size_t inserted = 1;
        while (inserted < n) {
            uint32_t b = (uint32_t)std::min<size_t>(std::min<size_t>(n - inserted, MAXB), std::max<size_t>(1, inserted / 16));
            p.first = (uint32_t)inserted; p.count = b;
            p.n_warps = (int)std::min<uint32_t>((uint32_t)max_warps, (b + 3u) & ~3u);
            GraphView g = ix->view();
            LEANN_CUDA_CHECK(cudaMemsetAsync(ctrs, 0, 16, stream));
            dispatch_dims((uint32_t)ix->d, ix->d4, LaunchA{g, p, stream});
            uint32_t max_e = std::min<uint32_t>(p.max_edges, b * (uint32_t)R);
            reverse_append_kernel<<<(max_e + 255) / 256, 256, 0, stream>>>(p);
            LEANN_CUDA_CHECK(cudaGetLastError());
            int pblocks = (int)std::min<uint32_t>((uint32_t)sms * 3u, (max_e + 3u) / 4u);
            dispatch_dims((uint32_t)ix->d, ix->d4, LaunchC{g, p, stream, std::max(pblocks, 1)});
            inserted += b;
        }
        uint32_t h_over[2] = {0, 0};
        LEANN_CUDA_CHECK(cudaMemcpyAsync(h_over, p.overflow, 8, cudaMemcpyDeviceToHost, stream));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (h_over[0]) throw Error(LEANN_ERR_CUDA, "vamana build: reverse-edge buffer overflow");
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace leann
