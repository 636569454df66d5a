// graph_search.cu — K1 / K1f: warp-per-query graph beam search (HNSW upper-level greedy descent +
// level-0 beam; Vamana single-level beam), replacing usearch::Index::search behind
// HnswSearcher::search (leann-rs src/backend/hnsw.rs:79-88) and DiskANN::search_with_dists behind
// DiskAnnSearcher::search (src/backend/diskann.rs:47-62).
//
// Persistent grid: a fixed pool of warps pulls queries from an atomic counter, so long and short
// traversals balance and the visited workspace is sized by the grid, not the batch.
// Bound: HBM gather bandwidth on long rows (0.91 of the copy peak at d = 768); on short rows (0.58 at d = 96) the dependent
// latency of a hop: about 2 300 instructions and five memory round trips paid in sequence by one warp, with 24 warps per SM
// to cover them ("What bounds K1 on short rows": DRAM transactions are not the limit).
// Algorithmic bytes per query =
//     n_dist * d4*16  +  n_hops0 * deg0*4  +  n_hops_upper * degU*4      
// all three counted by the kernel itself (out_stats) and by the oracle.
// Instantiations: <LPV lanes per vector, VPL float4 per lane, U unroll, MINB CTAs per SM> per row length (dispatch_search);
// short rows + diskann-rs stop rule + ef <= 128 + no mask take the register-list form <..., EPL = 4, SINGLE, Q16>.
#include <cstdlib>

#include "graph_device.cuh"

namespace leann {

// EPL > 0: `top` / `next` live in registers (RegList<EPL>, graph_device.cuh); the host picks that instantiation for short
// rows when max(ef, queue capacity) <= 32 * EPL and no mask is set. Shared memory then holds only the staging row.
template <int LPV, int VPL>
constexpr size_t ring_bytes() { return (size_t)RING_STAGES * RING_ROWS * VPL * LPV * 16 + 16 * ((RING_STAGES * 8 + 15) / 16); }

template <int LPV, int VPL, int U, int MINB, int EPL, bool SINGLE = false, bool Q16 = false, bool SMV = false, bool RING = false>
__global__ void __launch_bounds__(128, MINB)
graph_search_kernel(const GraphView g, const SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + warp_in_block;
    if (warp_global >= p.n_warps) return;

    const uint32_t ef_pad = EPL > 0 ? 0u : ((p.ef + 31u) & ~31u);
    const uint32_t ncapp = EPL > 0 ? 0u : p.next_capp;
    constexpr size_t vis_bytes = SMV ? (Q16 ? HYB_BYTES : SMV_BYTES) : 0u;
    const size_t per_warp = (size_t)ef_pad * 8 + (size_t)ncapp * 8 + (size_t)MAX_DEG * 8 + vis_bytes + (RING ? ring_bytes<LPV, VPL>() : 0u);
    unsigned char* base = smem_raw + per_warp * warp_in_block;
    WarpLists w;
    w.top_d = reinterpret_cast<float*>(base);
    w.top_s = reinterpret_cast<uint32_t*>(base + (size_t)ef_pad * 4);
    w.next_d = reinterpret_cast<float*>(base + (size_t)ef_pad * 8);
    w.next_s = reinterpret_cast<uint32_t*>(base + (size_t)ef_pad * 8 + (size_t)ncapp * 4);
    w.st_slot = reinterpret_cast<uint32_t*>(base + (size_t)ef_pad * 8 + (size_t)ncapp * 8);
    w.st_dist = reinterpret_cast<float*>(base + (size_t)ef_pad * 8 + (size_t)ncapp * 8 + (size_t)MAX_DEG * 4);

    VisitedSet vs;
    vs.n_pad = p.n_pad;
    vs.tbl = p.vhash ? p.vhash + (size_t)warp_global * p.vhash_cap : nullptr;
    vs.cap_mask = p.vhash_cap - 1u;
    vs.shift = 32u - (uint32_t)__ffs((int)p.vhash_cap) + 1u;
    vs.limit = p.vhash_cap / 4u * 3u;
    vs.pool_vis = p.visited; vs.pool_epochs = p.epochs; vs.pool_locks = p.pool_locks; vs.n_slots = p.pool_slots;
    vs.vis = p.vhash ? nullptr : p.visited + (size_t)warp_global * p.n_pad;
    vs.epoch_slot = p.vhash ? nullptr : p.epochs + warp_global;
    vs.tag = 0; vs.slot = -1;
    vs.q16 = Q16 && p.vhash != nullptr;
    vs.q_rem_bits = p.q_rem_bits; vs.q_kmask = p.q_key_bits >= 32 ? 0xFFFFFFFFu : ((1u << p.q_key_bits) - 1u); vs.q_inv = p.q_inv;
    vs.q_bmask = p.vhash_cap / 8u - 1u;
    if (vs.q16) {   // 2 bytes per entry
        vs.tbl = p.vhash + (size_t)warp_global * (p.vhash_cap / 2u);
        vs.limit = p.vhash_cap / 8u * 5u;
    }
    if (SMV) {      // two-choice table in this warp's shared memory (after the staging rows)
        vs.smv = true;
        vs.stbl = (uint32_t)__cvta_generic_to_shared(base + (size_t)MAX_DEG * 8);
        vs.s_bmask = (Q16 ? HYB_BUCKETS : SMV_BUCKETS) - 1u;
        vs.s_rem_bits = p.q_key_bits - (Q16 ? 9u : 10u);
        if (Q16) {  // hybrid: first level here, overflow level = the q16 table set up above
            vs.s_limit = p.smv_limit;
        } else {    // stand-alone
            vs.q16 = false; vs.tbl = nullptr; vs.vis = nullptr; vs.epoch_slot = nullptr;
            vs.limit = p.smv_limit;
        }
    }

    RowRing ring{0u, 0u, 0u};
    if constexpr (RING) {   // rows of a hop travel through this warp's shared-memory ring (graph_device.cuh)
        ring.rows = (uint32_t)__cvta_generic_to_shared(base + (size_t)MAX_DEG * 8 + vis_bytes);
        ring.bars = ring.rows + (uint32_t)(RING_STAGES * RING_ROWS * VPL * LPV * 16);
        ring_init(ring, lane);
    }

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(p.counter, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= p.nq) break;

        float4 q[VPL];
        load_query<LPV, VPL>(p.queries + (size_t)qi * g.d, g.d, q, lane);
        visited_begin(vs, lane);
        Counters c{0u, 0u, 0u, 0u};

        // entry point distance
        uint32_t cur = g.entry;
        if (lane == 0) w.st_slot[0] = cur;
        __syncwarp();
        eval_distances<LPV, VPL, U>(g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, 1, lane);
        float cur_d = w.st_dist[0];
        c.n_dist = 1;
        __syncwarp();
        if (g.max_level > 0) greedy_descend<LPV, VPL, U>(g, q, w, cur, cur_d, g.max_level, 0, c, lane);

        LevelAdj adj{g.adj0, g.adjU, g.upper_base, g.deg0, 0};
        int cnt;
        if constexpr (EPL > 0) {
            cnt = beam_level_regs<LPV, VPL, U, EPL, SINGLE, Q16, SMV, RING>(g, adj, q, w, (int)p.ef, (int)p.next_cap, p.nonstrict_term, vs,
                                                    (uint32_t)warp_global, cur, cur_d, c, lane, p.k,
                                                    p.out_keys + (size_t)qi * p.k, p.out_dists + (size_t)qi * p.k, &ring);
            visited_end(vs, lane);
        } else {
            beam_level<LPV, VPL, U, (LPV < 32)>(g, adj, q, w, (int)p.ef, (int)p.next_cap, (int)p.next_capp - 1, p.nonstrict_term,
                                                p.mask, vs, (uint32_t)warp_global, cur, cur_d, c, lane);
            visited_end(vs, lane);
            // results: ascending, truncated to k; tail = UINT64_MAX / +inf
            cnt = w.top_size < (int)p.k ? w.top_size : (int)p.k;
            for (uint32_t i = lane; i < p.k; i += 32) {
                uint64_t key = ~0ull;
                float dd = CUDART_INF_F;
                if ((int)i < cnt) {
                    uint32_t s = w.top_s[i];
                    key = g.keys ? g.keys[s] : (uint64_t)s;
                    dd = w.top_d[i];
                }
                p.out_keys[(size_t)qi * p.k + i] = key;
                p.out_dists[(size_t)qi * p.k + i] = dd;
            }
        }
        if (lane == 0) {
            if (p.out_counts) p.out_counts[qi] = (uint32_t)cnt;
            if (p.out_stats) {
                p.out_stats[(size_t)qi * 4 + 0] = c.n_dist;
                p.out_stats[(size_t)qi * 4 + 1] = c.n_hops0;
                p.out_stats[(size_t)qi * 4 + 2] = c.n_hops_upper;
                p.out_stats[(size_t)qi * 4 + 3] = c.dropped;
            }
        }
        __syncwarp();
    }
}

// Small batches (nq <= two per SM): one CTA of COOP_WARPS warps per query. Warp 0 runs the same traversal as above and
// owns the lists; all warps evaluate the distances of each staged neighbour list (coop_eval / coop_serve), which is where a
// single-warp traversal spends most of a hop at d = 768. Results are bit-identical to the warp-per-query kernel.
// COOP_WARPS = 8 while every query gets an SM to itself (nq <= SM count), 4 up to two queries per SM.
template <int LPV, int VPL, int U, int MINB, int COOP_WARPS>
__global__ void __launch_bounds__(COOP_WARPS * 32, (COOP_WARPS == 4 ? MINB : 1))
graph_search_coop_kernel(const GraphView g, const SearchParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_cmd;
    __shared__ uint32_t s_qi;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t ef_pad = (p.ef + 31u) & ~31u;
    WarpLists w;
    w.top_d = reinterpret_cast<float*>(smem_raw);
    w.top_s = reinterpret_cast<uint32_t*>(smem_raw + (size_t)ef_pad * 4);
    w.next_d = reinterpret_cast<float*>(smem_raw + (size_t)ef_pad * 8);
    w.next_s = reinterpret_cast<uint32_t*>(smem_raw + (size_t)ef_pad * 8 + (size_t)p.next_capp * 4);
    w.st_slot = reinterpret_cast<uint32_t*>(smem_raw + (size_t)ef_pad * 8 + (size_t)p.next_capp * 8);
    w.st_dist = reinterpret_cast<float*>(smem_raw + (size_t)ef_pad * 8 + (size_t)p.next_capp * 8 + (size_t)MAX_DEG * 4);
    VisitedSet vs;
    vs.n_pad = p.n_pad;
    vs.tbl = p.vhash ? p.vhash + (size_t)blockIdx.x * p.vhash_cap : nullptr;
    vs.cap_mask = p.vhash_cap - 1u;
    vs.shift = 32u - (uint32_t)__ffs((int)p.vhash_cap) + 1u;
    vs.limit = p.vhash_cap / 4u * 3u;
    vs.pool_vis = p.visited; vs.pool_epochs = p.epochs; vs.pool_locks = p.pool_locks; vs.n_slots = p.pool_slots;
    vs.vis = p.vhash ? nullptr : p.visited + (size_t)blockIdx.x * p.n_pad;
    vs.epoch_slot = p.vhash ? nullptr : p.epochs + blockIdx.x;
    vs.tag = 0; vs.slot = -1;
    const Coop cp{&s_cmd, warp, COOP_WARPS};

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_qi = atomicAdd(p.counter, 1u);
        __syncthreads();
        const uint32_t qi = s_qi;
        if (qi >= p.nq) break;
        float4 q[VPL];
        load_query<LPV, VPL>(p.queries + (size_t)qi * g.d, g.d, q, lane);
        if (warp != 0) {
            coop_serve<LPV, VPL, U>(cp, g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, lane);
            continue;
        }
        visited_begin(vs, lane);
        Counters c{0u, 0u, 0u, 0u};
        uint32_t cur = g.entry;
        if (lane == 0) w.st_slot[0] = cur;
        __syncwarp();
        coop_eval<LPV, VPL, U>(cp, g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, 1, lane);
        float cur_d = w.st_dist[0];
        c.n_dist = 1;
        __syncwarp();
        if (g.max_level > 0) greedy_descend<LPV, VPL, U>(g, q, w, cur, cur_d, g.max_level, 0, c, lane, cp);
        LevelAdj adj{g.adj0, g.adjU, g.upper_base, g.deg0, 0};
        beam_level<LPV, VPL, U, (LPV < 32)>(g, adj, q, w, (int)p.ef, (int)p.next_cap, (int)p.next_capp - 1, p.nonstrict_term,
                                            p.mask, vs, (uint32_t)blockIdx.x, cur, cur_d, c, lane, cp);
        visited_end(vs, lane);
        coop_finish(cp, lane);
        const int cnt = w.top_size < (int)p.k ? w.top_size : (int)p.k;
        for (uint32_t i = lane; i < p.k; i += 32) {
            uint64_t key = ~0ull;
            float dd = CUDART_INF_F;
            if ((int)i < cnt) {
                uint32_t s = w.top_s[i];
                key = g.keys ? g.keys[s] : (uint64_t)s;
                dd = w.top_d[i];
            }
            p.out_keys[(size_t)qi * p.k + i] = key;
            p.out_dists[(size_t)qi * p.k + i] = dd;
        }
        if (lane == 0) {
            if (p.out_counts) p.out_counts[qi] = (uint32_t)cnt;
            if (p.out_stats) {
                p.out_stats[(size_t)qi * 4 + 0] = c.n_dist;
                p.out_stats[(size_t)qi * 4 + 1] = c.n_hops0;
                p.out_stats[(size_t)qi * 4 + 2] = c.n_hops_upper;
                p.out_stats[(size_t)qi * 4 + 3] = c.dropped;
            }
        }
        __syncwarp();
    }
}

// Lanes cooperating on one distance: 8 for d <= 256 (four vectors side by side), else 32.
int reduction_lanes(size_t dims) { return dims <= 256 ? 8 : 32; }

size_t graph_search_smem_per_warp(uint32_t ef, uint32_t next_capp) {
    uint32_t ef_pad = (ef + 31u) & ~31u;
    return (size_t)ef_pad * 8 + (size_t)next_capp * 8 + (size_t)MAX_DEG * 8;
}

int graph_search_max_warps(int device) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms * 12;  // builder kernels: 3 CTAs x 4 warps per SM (register-bound, see -Xptxas -v)
}

namespace {
// op 0: launch; op 1: report resident warps per SM for this instantiation and shared-memory size.
constexpr int REG_EPL = 4;   // register lists: up to 128 entries
inline bool use_reg_lists(const SearchParams& p) {
    return p.mask == nullptr && p.ef <= 32u * REG_EPL && p.next_cap <= 32u * REG_EPL && p.k <= 32u * REG_EPL;
}

// one list with an "expanded" bit instead of top + next (graph_device.cuh): diskann-rs stop rule, slots below 2^31, and
// the tie rules this form was derived for
inline bool use_single_list(const GraphView& g, const SearchParams& p) {
    return p.nonstrict_term && g.n < 0x80000000u && compat::TOP_NEWCOMER_BEFORE_EQUALS && compat::NEXT_FIFO_AMONG_EQUALS &&
           getenv("LEANN_CUDA_DISABLE_SINGLE_LIST") == nullptr;
}

template <int LPV, int VPL, int U, int MINB, int EPL = 0, bool SINGLE = false, bool Q16 = false, bool SMV = false, bool RING = false>
int launch_t(const GraphView& g, const SearchParams& p, cudaStream_t stream, int op);

// short rows, diskann-rs stop rule: register-list instantiation (one list with an expanded bit; u32 / byte-map or q16
// visited set). The two-list register form (usearch stop rule) exists in graph_device.cuh but is not instantiated: on 1M x 128 /
// 1M x 256 HNSW indexes it measured 3-9 % slower than the shared-memory lists (two sorted inserts per accepted neighbour).
template <int LPV, int VPL, int U, int MINB>
int launch_reg(const GraphView& g, const SearchParams& p, cudaStream_t stream, int op) {
    // The A/B forms below exist for rows of at most 128 floats (the host never selects them beyond that: api.cu smv_plan)
    if constexpr (VPL <= 4) {
        // visited tables in shared memory, stand-alone (graph_device.cuh "smv"): 3 CTAs of 4 warps per SM, so the register
        // budget is no longer 80 per thread; unroll 3 measured best of {2, 3, 4} (profiles/r2_k1_smv_ab.log)
        if (p.smem_vis == 1) return launch_t<LPV, VPL, 3, 3, REG_EPL, true, false, true>(g, p, stream, op);
        // hybrid: shared-memory first level + q16 overflow level, the usual occupancy (op 1 is asked before the q16 table exists)
        if (p.smem_vis == 2) return launch_t<LPV, VPL, U, MINB, REG_EPL, true, true, true>(g, p, stream, op);
    }
    const bool q16 = p.vhash != nullptr && p.vhash16 != 0;
    if constexpr (VPL <= 4) {
        if (p.row_ring) {
            // rows through a shared-memory ring of bulk async copies (graph_device.cuh eval_distances_ring). Bit-identical and OFF
            // by default: 4.88 vs 4.66 ms at L = 100, 2.72 vs 2.61 ms at L = 50 on the 12.5M x 96 shard; with 7 CTAs per SM (72
            // registers) 5.66 / 3.15 ms (profiles/r2_k1_ring_ab.log)
            return q16 ? launch_t<LPV, VPL, U, MINB, REG_EPL, true, true, false, true>(g, p, stream, op)
                       : launch_t<LPV, VPL, U, MINB, REG_EPL, true, false, false, true>(g, p, stream, op);
        }
    }
    return q16 ? launch_t<LPV, VPL, U, MINB, REG_EPL, true, true>(g, p, stream, op) : launch_t<LPV, VPL, U, MINB, REG_EPL, true, false>(g, p, stream, op);
}

template <int LPV, int VPL, int U, int MINB, int EPL, bool SINGLE, bool Q16, bool SMV, bool RING>
int launch_t(const GraphView& g, const SearchParams& p, cudaStream_t stream, int op) {
    if constexpr (EPL == 0 && LPV < 32) {
        // short rows: the register-list instantiation when the lists fit (small batches keep the cooperative kernel)
        const bool off = getenv("LEANN_CUDA_DISABLE_REG_LISTS") != nullptr;   // A/B switch for benchmarks
        if (!off && p.coop_ctas == 0 && use_reg_lists(p) && use_single_list(g, p)) return launch_reg<LPV, VPL, U, MINB>(g, p, stream, op);
    }
    const int warps_per_block = 4;
    size_t smem = (EPL > 0 ? (size_t)MAX_DEG * 8 + (SMV ? (Q16 ? HYB_BYTES : SMV_BYTES) : 0u) + (RING ? ring_bytes<LPV, VPL>() : 0u)
                           : graph_search_smem_per_warp(p.ef, p.next_capp)) * warps_per_block;
    auto kern = graph_search_kernel<LPV, VPL, U, MINB, EPL, SINGLE, Q16, SMV, RING>;
    if (smem > 48 * 1024) LEANN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (op == 1) {
        int blocks_per_sm = 0;
        LEANN_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, warps_per_block * 32, smem));
        return blocks_per_sm * warps_per_block;
    }
    LEANN_CUDA_CHECK(cudaMemsetAsync(p.counter, 0, sizeof(uint32_t), stream));
    if (p.coop_ctas > 0) {
        // small batch: one CTA of COOP_WARPS warps per query (visited slices are indexed by CTA, so coop_ctas <= n_warps)
        const int cw = p.coop_warps == 8 ? 8 : 4;
        if constexpr (EPL > 0) throw Error(LEANN_ERR_INVALID_ARG, "internal: register lists are not used by the cooperative kernel");
        auto ck = cw == 8 ? graph_search_coop_kernel<LPV, VPL, U, MINB, 8> : graph_search_coop_kernel<LPV, VPL, U, MINB, 4>;
        const size_t csmem = graph_search_smem_per_warp(p.ef, p.next_capp);
        if (csmem > 48 * 1024) LEANN_CUDA_CHECK(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
        ck<<<p.coop_ctas, cw * 32, csmem, stream>>>(g, p);
        LEANN_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    int blocks = (p.n_warps + warps_per_block - 1) / warps_per_block;
    kern<<<blocks, warps_per_block * 32, smem, stream>>>(g, p);
    LEANN_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int dispatch_search(const GraphView& g, const SearchParams& p, cudaStream_t stream, int op) {
    if (g.deg0 > (uint32_t)MAX_DEG || g.degU > (uint32_t)MAX_DEG)
        throw Error(LEANN_ERR_INVALID_ARG, "graph degree exceeds MAX_DEG (128)");
    if (p.ef > (uint32_t)MAX_EF) throw Error(LEANN_ERR_INVALID_ARG, "ef exceeds 1024");
    const uint32_t d4 = g.d4;
    if (reduction_lanes(g.d) == 8) {
        // small rows: the traversal is instruction-latency bound, so favour resident warps over unroll depth
        uint32_t vpl = (d4 + 7) / 8;
        if (vpl <= 2) return launch_t<8, 2, 4, 6>(g, p, stream, op);
        if (vpl <= 3) {
            // d <= 96. Register lists (r2): <U = 2, 6 CTAs per SM> measured best of {4,6; 4,5; 3,6; 3,5; 2,6; 2,7; 2,8; 4,4}
            // on the 12.5M x 96 Vamana shard (profiles/r2_k1_tune*.log); shared-memory lists (ef > 128, masks): <4, 6>.
            if (p.coop_ctas == 0 && use_reg_lists(p) && use_single_list(g, p) && getenv("LEANN_CUDA_DISABLE_REG_LISTS") == nullptr)
                return launch_reg<8, 3, 2, 6>(g, p, stream, op);
            return launch_t<8, 3, 4, 6>(g, p, stream, op);   // d = 96: 6 CTAs per SM measured 6 % faster than 5, deeper unrolls slower
        }
        if (vpl <= 4) return launch_t<8, 4, 2, 6>(g, p, stream, op);   // d = 128: +10-15 % over <8,4,4,4>
        return launch_t<8, 8, 2, 4>(g, p, stream, op);                  // d = 256: +6-7 % over <8,8,2,3>
    }
    uint32_t vpl = (d4 + 31) / 32;
    // occupancy / unroll per row length from sweeps on 1M-vector HNSW indexes (d = 384: +13 %, d = 512: +7 % over deeper unrolls)
    if (vpl <= 3) return launch_t<32, 3, 4, 6>(g, p, stream, op);
    if (vpl <= 4) return launch_t<32, 4, 3, 6>(g, p, stream, op);
    if (vpl <= 6) return launch_t<32, 6, 4, 3>(g, p, stream, op);
    if (vpl <= 8) return launch_t<32, 8, 2, 3>(g, p, stream, op);
    if (vpl <= 12) return launch_t<32, 12, 2, 2>(g, p, stream, op);
    if (vpl <= 16) return launch_t<32, 16, 1, 3>(g, p, stream, op);
    if (vpl <= 32) return launch_t<32, 32, 1, 2>(g, p, stream, op);
    throw Error(LEANN_ERR_INVALID_ARG, "dimension above 4096 is not supported");
}
}  // namespace

bool graph_search_uses_reg_lists(const GraphView& g, const SearchParams& p) {
    return reduction_lanes(g.d) == 8 && p.coop_ctas == 0 && use_reg_lists(p) && use_single_list(g, p) &&
           getenv("LEANN_CUDA_DISABLE_REG_LISTS") == nullptr;
}

int graph_search_warps_per_sm(const GraphView& g, uint32_t ef, uint32_t next_capp, int nonstrict_term, int smem_vis, int row_ring) {
    SearchParams p{};
    p.ef = ef; p.next_capp = next_capp; p.nonstrict_term = nonstrict_term; p.smem_vis = smem_vis; p.row_ring = row_ring;
    return dispatch_search(g, p, nullptr, 1);
}

void launch_graph_search(const GraphView& g, const SearchParams& p, cudaStream_t stream) { dispatch_search(g, p, stream, 0); }

}  // namespace leann
