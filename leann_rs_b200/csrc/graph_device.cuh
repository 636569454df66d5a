// graph_device.cuh — warp-cooperative building blocks of the graph traversal kernels (K1/K1f and
// the HNSW builder). One warp owns one query: the query lives in registers; `top` (the bounded
// result list, usearch sorted_buffer_gt) and `next` (the candidate queue) live in shared memory
// (beam_level) or, on short rows under the diskann-rs stop rule, as ONE register-resident list with an
// "expanded" bit per entry (RegList / beam_level_regs). The visited set is exact in all three of its
// representations (VisitedSet): an epoch-tagged byte map in HBM (no clearing between queries), a per-warp
// u32 open-addressing table, or a per-warp bucketed table of 16-bit quotiented entries (q16) that stays
// in L2; the tables move a traversal that outgrows them to a pooled byte map.
//
// Semantics restated from usearch search_for_one_ / search_to_find_in_base_ / search_to_insert_
// (call site leann-rs src/backend/hnsw.rs:85) and diskann-rs search_with_dists (diskann.rs:56);
// see oracle/graph_oracle.cpp for the CPU restatement the tests compare against.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "compat.h"
#include "internal.h"

namespace leann {

constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// L2 eviction priorities of the short-row traversal (LPV == 8), A/B-able at compile time (benchmarks/k1_variants.sh):
//   bit 0: vector rows are loaded evict_first (each row is used once; they should not push the visited tables out of L2)
//   bit 1: the q16 visited tables are read / cleared evict_last (116 MB of tables for 24 warps per SM are meant to stay in L2)
//   bit 2: adjacency rows are loaded evict_first
#ifndef LEANN_K1_L2POL
#define LEANN_K1_L2POL 0
#endif
#ifndef LEANN_K1_SPEC
#define LEANN_K1_SPEC 0   // speculative row prefetch of the register-list traversal (see beam_level_regs)
#endif
__device__ __forceinline__ float4 ldg_stream_evict_first(const float4* p) {
    float4 r;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint32_t ldg_u32_evict_first(const uint32_t* p) {
    uint32_t r;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint4 ldcg_u4_evict_last(const uint4* p) {
    uint4 r;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_u4_evict_last(uint4* p, const uint4 v) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Per-warp shared-memory state.
struct WarpLists {
    float* top_d; uint32_t* top_s;    // ascending, size <= ef
    float* next_d; uint32_t* next_s;  // ascending ring buffer, physical index (head+i)&(capp-1)
    uint32_t* st_slot; float* st_dist;  // staging for one adjacency row (MAX_DEG)
    int top_size, next_size, next_head;
};

// Packed f32x2 arithmetic (sm_100: FFMA2 / FADD2): two independent round-to-nearest operations per instruction, bit for
// bit the results of the scalar __fmaf_rn / __fsub_rn pair. The traversal on short rows is bound by issued instructions
// and memory latency, not by FP throughput; halving the FP instruction count is what these buy.
__device__ __forceinline__ void fma2_rn(float& cx, float& cy, float ax, float ay, float bx, float by) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(ax), "f"(ay));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(bx), "f"(by));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(cx), "f"(cy));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(cx), "=f"(cy) : "l"(c));
}
__device__ __forceinline__ void sub2_rn(float& dx, float& dy, float ax, float ay, float bx, float by) {
    unsigned long long a, b, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(ax), "f"(ay));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(bx), "f"(by));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(dx), "=f"(dy) : "l"(d));
}

// Lane-partial of one distance: float4 index i*LPV+lig, four fma accumulators, (x+y)+(z+w).
template <int VPL>
__device__ __forceinline__ float lane_partial(const float4 (&q)[VPL], const float4 (&x)[VPL], int metric) {
    float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
    if (metric == LEANN_METRIC_L2SQ) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            float tx, ty, tz, tw;
            sub2_rn(tx, ty, q[i].x, q[i].y, x[i].x, x[i].y);
            sub2_rn(tz, tw, q[i].z, q[i].w, x[i].z, x[i].w);
            fma2_rn(ax, ay, tx, ty, tx, ty);
            fma2_rn(az, aw, tz, tw, tz, tw);
        }
    } else {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            fma2_rn(ax, ay, q[i].x, q[i].y, x[i].x, x[i].y);
            fma2_rn(az, aw, q[i].z, q[i].w, x[i].z, x[i].w);
        }
    }
    return __fadd_rn(__fadd_rn(ax, ay), __fadd_rn(az, aw));
}

template <int LPV>
__device__ __forceinline__ float group_reduce(float v) {
#pragma unroll
    for (int off = LPV / 2; off >= 1; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(FULL, v, off));
    return v;
}

__device__ __forceinline__ float finish_distance(float s, int metric) {
    if (metric == LEANN_METRIC_L2SQ) return s;
    float r = __fsub_rn(1.0f, s);
    if (compat::DISTDOT_CLAMP_AT_ZERO && metric == LEANN_METRIC_IP_CLAMP) r = r < 0.0f ? 0.0f : r;
    return r;
}

// Distances from the register-resident query to st_slot[0..cnt) -> st_dist[0..cnt).
// 32/LPV vectors are evaluated side by side, U deep: U*VPL independent 16-byte loads per lane are
// in flight before the first FMA.
// `first` / `stride` (in entries, multiples of the batch size) let several warps share one list: every entry is
// still evaluated by exactly one warp with the same lane mapping, so the bits do not depend on the split.
template <int LPV, int VPL, int U>
__device__ __forceinline__ void eval_distances(const float4* __restrict__ vecs, uint32_t d4, int metric,
                                               const float4 (&q)[VPL], const uint32_t* st_slot,
                                               float* st_dist, int cnt, int lane, int first = 0,
                                               int stride = U * (32 / LPV)) {
    constexpr int GROUPS = 32 / LPV;
    constexpr int BATCH = U * GROUPS;
    const int gid = lane / LPV, lig = lane % LPV;
    for (int b = first; b < cnt; b += stride) {
        float4 x[U][VPL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int j = b + u * GROUPS + gid;
            bool ok = j < cnt;
            uint32_t s = ok ? st_slot[j] : 0u;
            const float4* row = vecs + (size_t)s * d4;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                uint32_t idx = (uint32_t)(i * LPV + lig);
                if (ok && idx < d4) x[u][i] = (LPV == 8 && (LEANN_K1_L2POL & 1)) ? ldg_stream_evict_first(row + idx) : ldg_stream(row + idx);
                else x[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float v = group_reduce<LPV>(lane_partial<VPL>(q, x[u], metric));
            int j = b + u * GROUPS + gid;
            if (j < cnt && lig == 0) st_dist[j] = finish_distance(v, metric);
        }
    }
    __syncwarp();
}

// ---- row ring: the rows of a hop through shared memory (short rows, register-list kernels) -------------------------
// eval_distances holds U * VPL float4 per lane while a batch of U * 32/LPV rows is in flight and pays one L2 / DRAM round
// trip per batch (5-6 per hop at d = 96: the largest single stall of the traversal). Here the rows travel as bulk async
// copies (cp.async.bulk, one instruction per row, issued by as many lanes as there are rows) into a per-warp ring of
// RING_STAGES x RING_ROWS rows; an mbarrier per stage counts the bytes. Stage t + RING_STAGES is requested as soon as stage t
// has been consumed, so after the first stage of a hop the copies overlap the arithmetic, and no registers are tied up by
// rows in flight. The arithmetic (lane mapping, accumulators, reduction tree) is eval_distances', bit for bit.
constexpr int RING_ROWS = 8, RING_STAGES = 2;
struct RowRing {
    uint32_t rows;      // shared-space address of [RING_STAGES][RING_ROWS][row_bytes]
    uint32_t bars;      // shared-space address of RING_STAGES mbarriers (8 bytes each)
    uint32_t phase;     // bit s = parity the next wait on stage s expects
};
__device__ __forceinline__ void ring_init(const RowRing& r, int lane) {
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < RING_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(r.bars + 8u * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
}
template <int LPV, int VPL>
__device__ __forceinline__ void eval_distances_ring(RowRing& r, const float4* __restrict__ vecs, uint32_t d4, int metric,
                                                    const float4 (&q)[VPL], const uint32_t* st_slot, float* st_dist, int cnt,
                                                    int lane) {
    constexpr int GROUPS = 32 / LPV;
    static_assert(RING_ROWS % GROUPS == 0, "a stage holds whole compute steps");
    const int gid = lane / LPV, lig = lane % LPV;
    const uint32_t row_bytes = d4 * 16u, row_stride = (uint32_t)(VPL * LPV) * 16u;
    const int n_stages = (cnt + RING_ROWS - 1) / RING_ROWS;
    auto issue = [&](int t) {
        const int s = t % RING_STAGES, first = t * RING_ROWS;
        const int nrows = min(RING_ROWS, cnt - first);
        const uint32_t bar = r.bars + 8u * (uint32_t)s;
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)nrows * row_bytes) : "memory");
        __syncwarp();
        if (lane < nrows) {
            const float4* src = vecs + (size_t)st_slot[first + lane] * d4;
            const uint32_t dst = r.rows + (uint32_t)(s * RING_ROWS + lane) * row_stride;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                         "r"(row_bytes), "r"(bar)
                         : "memory");
        }
    };
#pragma unroll
    for (int t = 0; t < RING_STAGES; ++t)
        if (t < n_stages) issue(t);
    for (int t = 0; t < n_stages; ++t) {
        const int s = t % RING_STAGES;
        const uint32_t bar = r.bars + 8u * (uint32_t)s, par = (r.phase >> s) & 1u;
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(par)
                : "memory");
        }
        r.phase ^= 1u << s;
#pragma unroll
        for (int step = 0; step < RING_ROWS / GROUPS; ++step) {
            const int slot = step * GROUPS + gid, j = t * RING_ROWS + slot;
            const bool ok = j < cnt;
            float4 x[VPL];
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const uint32_t idx = (uint32_t)(i * LPV + lig);
                if (ok && idx < d4) x[i] = lds_f4(r.rows + (uint32_t)(s * RING_ROWS + slot) * row_stride + idx * 16u);
                else x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float v = group_reduce<LPV>(lane_partial<VPL>(q, x, metric));
            if (ok && lig == 0) st_dist[j] = finish_distance(v, metric);
        }
        __syncwarp();   // every lane has read stage s before it is refilled
        if (t + RING_STAGES < n_stages) issue(t + RING_STAGES);
    }
    __syncwarp();
}

template <int LPV, int VPL>
__device__ __forceinline__ void load_query(const float* __restrict__ qrow, uint32_t d, float4 (&q)[VPL], int lane) {
    const int lig = lane % LPV;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        uint32_t j = (uint32_t)(i * LPV + lig) * 4u;
        q[i].x = j + 0 < d ? qrow[j + 0] : 0.f;
        q[i].y = j + 1 < d ? qrow[j + 1] : 0.f;
        q[i].z = j + 2 < d ? qrow[j + 2] : 0.f;
        q[i].w = j + 3 < d ? qrow[j + 3] : 0.f;
    }
}

// Several warps of one CTA can serve ONE query (small batches: the reference's API is one query per call): warp 0 runs
// the traversal and owns the lists, the other warps only evaluate their share of every staged neighbour list.
// Protocol: warp 0 publishes the list length (or -1 = query finished) and both sides meet at two named barriers.
struct Coop {
    int* cmd;      // shared-memory command word (nullptr: single-warp mode)
    int warp, nwarps;
};
__device__ __forceinline__ void coop_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
template <int LPV, int VPL, int U>
__device__ __forceinline__ void coop_eval(const Coop& cp, const float4* __restrict__ vecs, uint32_t d4, int metric,
                                          const float4 (&q)[VPL], const uint32_t* st_slot, float* st_dist, int cnt, int lane) {
    if (cp.cmd == nullptr) {
        eval_distances<LPV, VPL, U>(vecs, d4, metric, q, st_slot, st_dist, cnt, lane);
        return;
    }
    constexpr int BATCH = U * (32 / LPV);
    if (lane == 0) *cp.cmd = cnt;
    coop_bar(cp.nwarps * 32);
    eval_distances<LPV, VPL, U>(vecs, d4, metric, q, st_slot, st_dist, cnt, lane, 0, cp.nwarps * BATCH);
    coop_bar(cp.nwarps * 32);
}
// helper warps: serve evaluation requests until warp 0 signals the end of the query
template <int LPV, int VPL, int U>
__device__ __forceinline__ void coop_serve(const Coop& cp, const float4* __restrict__ vecs, uint32_t d4, int metric,
                                           const float4 (&q)[VPL], const uint32_t* st_slot, float* st_dist, int lane) {
    constexpr int BATCH = U * (32 / LPV);
    for (;;) {
        coop_bar(cp.nwarps * 32);
        const int cnt = *reinterpret_cast<volatile int*>(cp.cmd);
        if (cnt < 0) break;
        eval_distances<LPV, VPL, U>(vecs, d4, metric, q, st_slot, st_dist, cnt, lane, cp.warp * BATCH, cp.nwarps * BATCH);
        coop_bar(cp.nwarps * 32);
    }
}
__device__ __forceinline__ void coop_finish(const Coop& cp, int lane) {   // warp 0, after the last evaluation of a query
    if (cp.cmd == nullptr) return;
    if (lane == 0) *cp.cmd = -1;
    coop_bar(cp.nwarps * 32);
}

// Sorted insert into an ascending (optionally ring-indexed) array.
//  UPPER == false: position = #elements <  dd  (usearch sorted_buffer_gt: newcomer before equals)
//  UPPER == true : position = #elements <= dd  (FIFO among equals; candidate queue)
// When the array already holds `limit` entries the last one is dropped; pos == limit rejects.
template <bool UPPER, bool RING>
__device__ __forceinline__ bool sorted_insert(float* D, uint32_t* S, int& size, int limit, int head, int mask,
                                              float dd, uint32_t ss, int lane) {
    auto phys = [&](int i) { return RING ? ((head + i) & mask) : i; };
    int pos = 0;
    for (int base = 0; base < size; base += 32) {
        int i = base + lane;
        bool lt = false;
        if (i < size) {
            float v = D[phys(i)];
            lt = UPPER ? (v <= dd) : (v < dd);
        }
        unsigned b = __ballot_sync(FULL, lt);
        pos += __popc(b);
        if (b != FULL) break;
    }
    if (pos == limit) return false;
    int new_size = size + 1 < limit ? size + 1 : limit;
    int last = new_size - 1;
    for (int base = (last >> 5) << 5; base >= 0 && base + 31 > pos; base -= 32) {
        int i = base + lane;
        bool mv = (i > pos) && (i <= last);
        float td = 0.f; uint32_t ts = 0u;
        if (mv) { td = D[phys(i - 1)]; ts = S[phys(i - 1)]; }
        __syncwarp();
        if (mv) { D[phys(i)] = td; S[phys(i)] = ts; }
        __syncwarp();
    }
    if (lane == 0) { D[phys(pos)] = dd; S[phys(pos)] = ss; }
    __syncwarp();
    size = new_size;
    return true;
}

// Adjacency addressing of one level.
struct LevelAdj {
    const uint32_t* adj0; const uint32_t* adjU; const uint32_t* upper_base;
    uint32_t deg; int level;
    __device__ __forceinline__ const uint32_t* row(uint32_t s) const {
        return level == 0 ? adj0 + (size_t)s * deg : adjU + ((size_t)upper_base[s] + (uint32_t)(level - 1)) * deg;
    }
};

struct Counters { uint32_t n_dist, n_hops0, n_hops_upper, dropped; };

// Greedy descent over (from_level .. to_level] — usearch search_for_one_. Every neighbour of the
// current closest node is measured each pass (no visited set); strict `<` keeps the first minimum.
template <int LPV, int VPL, int U>
__device__ __forceinline__ void greedy_descend(const GraphView& g, const float4 (&q)[VPL], WarpLists& w,
                                               uint32_t& cur, float& cur_d, int from_level, int to_level,
                                               Counters& c, int lane, const Coop cp = Coop{nullptr, 0, 1}) {
    for (int level = from_level; level > to_level; --level) {
        bool changed;
        do {
            changed = false;
            const uint32_t* row = g.adjU + ((size_t)g.upper_base[cur] + (uint32_t)(level - 1)) * g.degU;
            int cnt = 0;
            for (uint32_t base = 0; base < g.degU; base += 32) {
                uint32_t j = base + lane;
                uint32_t s = j < g.degU ? __ldg(row + j) : SENT;
                bool ok = s != SENT;
                unsigned b = __ballot_sync(FULL, ok);
                if (ok) w.st_slot[cnt + __popc(b & ((1u << lane) - 1u))] = s;
                cnt += __popc(b);
            }
            __syncwarp();
            c.n_hops_upper++;
            c.n_dist += cnt;
            coop_eval<LPV, VPL, U>(cp, g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, cnt, lane);
            float bd = CUDART_INF_F; int bj = 0x7fffffff;
            for (int j = lane; j < cnt; j += 32) {
                float dj = w.st_dist[j];
                if (dj < bd) { bd = dj; bj = j; }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                float od = __shfl_xor_sync(FULL, bd, off);
                int oj = __shfl_xor_sync(FULL, bj, off);
                if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
            }
            if (bd < cur_d) { cur_d = bd; cur = w.st_slot[bj]; changed = true; }
            __syncwarp();
        } while (changed);
    }
}

// Epoch handling of the visited byte map: tag in [1,255]; the map is wiped when the tag wraps.
__device__ __forceinline__ uint8_t next_epoch(uint32_t* epoch_slot, uint8_t* vis, size_t n_pad, int lane) {
    uint32_t e = *epoch_slot;
    __syncwarp();
    if (e != 0 && e % 255u == 0) {
        uint4* v4 = reinterpret_cast<uint4*>(vis);
        for (size_t i = lane; i < n_pad / 16; i += 32) v4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    if (lane == 0) *epoch_slot = e + 1;
    return (uint8_t)(e % 255u + 1u);
}


// Exact visited set of one traversal.
//  * Byte map in HBM, one epoch tag per node, no clearing between queries: the default, one map per resident warp.
//  * Large indexes (n * resident warps would not fit in memory): an open-addressing table of node ids per warp, sized
//    from ef * degree and independent of n, cleared per query. A traversal that outgrows its table (count > limit) takes
//    one byte map from a small shared pool, replays the table into it and continues there, so the set stays exact for
//    any query; the pool slot goes back at the end of the query.
constexpr uint32_t VIS_EMPTY = 0xFFFFFFFFu;
struct VisitedSet {
    uint8_t* vis; uint8_t tag; uint32_t* epoch_slot; size_t n_pad;       // byte map in use (own map, or a pool slot after a spill)
    uint32_t* tbl; uint32_t shift, cap_mask, limit, count; bool hashed;  // hash table (nullptr: byte map only)
    uint8_t* pool_vis; uint32_t* pool_epochs; uint32_t* pool_locks; uint32_t n_slots; int slot;   // spill pool
    // q16: `tbl` holds 16-bit quotiented entries in buckets of 8 (one 16-byte load tests a key against a whole bucket)
    bool q16 = false; uint32_t q_rem_bits = 0, q_bmask = 0, q_kmask = 0, q_inv = 0;
    // smv: a table in shared memory (two-choice buckets, below); stbl = its shared-space address. With q16 set as well the
    // shared table is the first level and `tbl` the overflow level ("hybrid"): s_count of the `count` members live in shared
    // memory, ovf = the shared table is closed (inserts and unresolved lookups go to `tbl`).
    bool smv = false, ovf = false; uint32_t stbl = 0, s_bmask = 0, s_rem_bits = 0, s_limit = 0, s_count = 0;
};

// ---- q16: bucketed, quotiented visited table --------------------------------------------------------------------------
// The u32 table above is open addressing with linear probing: one atomicCAS per probed slot, and a warp that tests the 64
// neighbours of a hop waits for the LONGEST probe chain among them (3-4 dependent L2 / DRAM round trips at 28 % load; 33 % of
// all stall samples of the d = 96 traversal in profiles/r2_k1_d96_single_ncu_summary.json). Here:
//   * h = (slot * odd) mod 2^B is a bijection on the B-bit slot space; home bucket = top bits of h, remainder = the rest.
//     A 16-bit entry (remainder << 2 | displacement) in bucket `home + displacement` identifies the slot exactly, so the
//     table is exact at 2 bytes per entry: 32 KB per warp for 16384 entries — all resident warps together fit in L2.
//   * one 16-byte load reads the 8 entries of a bucket: a visited neighbour is recognised in one round trip, a fresh one
//     takes one more (a 16-bit atomicCAS on the first empty entry; a lost race retries on the next empty entry of the same
//     snapshot, the racing keys are different neighbours). A key moves to the next bucket only when its bucket is full, at
//     most 3 buckets away; beyond that (or above 62 % load) the traversal moves to a pooled byte map as the u32 table does.
constexpr uint32_t Q_EMPTY = 0xFFFFu;
constexpr uint32_t Q_HASH_MUL = 0x9E3779B1u;
__device__ __forceinline__ uint4 q_bucket(const VisitedSet& v, uint32_t b) {
    if (LEANN_K1_L2POL & 2) return ldcg_u4_evict_last(reinterpret_cast<const uint4*>(v.tbl) + b);
    return __ldcg(reinterpret_cast<const uint4*>(v.tbl) + b);
}
// found: `want` is one of the 8 entries; empties: bit i set = entry i is empty
__device__ __forceinline__ void q_scan(const uint4 w, uint32_t want, bool& found, uint32_t& empties) {
    const uint32_t x[4] = {w.x, w.y, w.z, w.w};
    found = false;
    empties = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t lo = x[i] & 0xFFFFu, hi = x[i] >> 16;
        found |= (lo == want) | (hi == want);
        empties |= (lo == Q_EMPTY ? 1u : 0u) << (2 * i) | (hi == Q_EMPTY ? 1u : 0u) << (2 * i + 1);
    }
}
// One key, one thread (entry point of a traversal; the hop loop uses the interleaved form in beam_level_regs).
// Returns 1 fresh (now inserted), 0 already present, -1 no room within 3 buckets (caller spills).
__device__ __forceinline__ int q_test_and_set(VisitedSet& v, uint32_t s) {
    const uint32_t h = (s * Q_HASH_MUL) & v.q_kmask;
    const uint32_t home = h >> v.q_rem_bits, rem = h & ((1u << v.q_rem_bits) - 1u);
    unsigned short* t16 = reinterpret_cast<unsigned short*>(v.tbl);
    for (uint32_t disp = 0; disp < 4; ++disp) {
        const uint32_t b = (home + disp) & v.q_bmask, want = (rem << 2) | disp;
        bool found;
        uint32_t empties;
        q_scan(q_bucket(v, b), want, found, empties);
        if (found) return 0;
        while (empties) {
            const int e = __ffs((int)empties) - 1;
            empties &= empties - 1;
            const unsigned short old = atomicCAS(t16 + b * 8 + e, (unsigned short)Q_EMPTY, (unsigned short)want);
            if (old == Q_EMPTY) return 1;
            if (old == want) return 0;
        }
    }
    return -1;
}

// ---- smv: the visited table of a traversal in SHARED memory ----------------------------------------------------------
// Short rows make the traversal DRAM-transaction bound, and the q16 tables above still cost a bucket load and a CAS through L2
// (or DRAM: 116 MB of tables share the L2 with the row stream) per neighbour test — about a fifth of the DRAM traffic of a hop.
// A warp can keep (part of) its table in shared memory: buckets of eight 16-bit entries, TWO candidate buckets per key so
// that the table works at 60-80 % load (the single-home form above needs <= 33 %):
//   h = (slot * odd) mod 2^B (a bijection), b1 = top bits, rem = the other <= 15 bits, b2 = b1 ^ g(rem), g != 0;
//   entry = rem << 1 | alt (0xFFFF = empty): alt = 0 in bucket b1, alt = 1 in bucket b2 — the pair (bucket, entry) identifies
//   the slot exactly, a lookup reads the two buckets (two LDS.128) and nothing else. (The one key per 2^15 whose alt entry
//   would equal the empty marker never enters the table: it is treated as "no room".)
// The table is private to the warp, so inserts need no atomics: the lanes of a chunk that picked the same bucket (least loaded
// of their two) are grouped with __match_any_sync, the i-th of a group takes the i-th empty entry of the snapshot, the others
// retry after a __syncwarp.
//   * stand-alone (1024 buckets = 16 KB per warp, 3 CTAs of 4 warps per SM): a key without room, or a table above its load
//     limit, moves the traversal to a pooled byte map as the other table forms do;
//   * hybrid (512 buckets = 8 KB per warp, the usual 6 CTAs per SM): the shared table takes the first ~3000 members of a
//     traversal; once it is closed (limit reached, or a key without room) later members go to the warp's q16 table in global
//     memory, which is cleared only then. A lookup tries shared memory first and the q16 table only after the closure.
constexpr uint32_t SMV_BUCKETS = 1024, SMV_BYTES = SMV_BUCKETS * 16;   // stand-alone
constexpr uint32_t HYB_BUCKETS = 512, HYB_BYTES = HYB_BUCKETS * 16;    // hybrid first level
__device__ __forceinline__ uint32_t smv_alt(uint32_t rem, uint32_t bmask) {
    const uint32_t g = ((rem * 0x5BD1u) >> 3) & bmask;
    return g ? g : 1u;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, const uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)v) : "memory");
}
// One chunk of up to 32 keys (one per lane, SENT = none), executed by the whole warp. Returns per lane: true = the key was
// not in the table and now is. `unresolved` is set for a key that is not in the table and was not inserted: no room, or
// `insert` is false (closed table: the caller looks further).
__device__ __forceinline__ bool smv_test_and_set(const VisitedSet& v, uint32_t s, int lane, bool& unresolved, bool insert) {
    const uint32_t h = (s * Q_HASH_MUL) & v.q_kmask;
    const uint32_t rem = h & ((1u << v.s_rem_bits) - 1u);
    const uint32_t b1 = h >> v.s_rem_bits, b2 = b1 ^ smv_alt(rem, v.s_bmask);
    const uint32_t e1 = rem << 1, e2 = e1 | 1u;
    const unsigned lt = (1u << lane) - 1u;
    bool pend = s != SENT, fresh = false;
    // equal keys inside the chunk (a malformed adjacency row): only the first lane decides, the others are "already seen"
    if (__match_any_sync(FULL, s) & lt) pend = false;
    unresolved = false;
    if (pend && e2 == Q_EMPTY) { unresolved = true; pend = false; }
    for (;;) {
        uint32_t em1 = 0, em2 = 0;
        int tgt = 0;
        if (pend) {
            const uint4 w1 = lds_u4(v.stbl + b1 * 16u), w2 = lds_u4(v.stbl + b2 * 16u);
            bool f1, f2;
            q_scan(w1, e1, f1, em1);
            q_scan(w2, e2, f2, em2);
            if (f1 || f2) pend = false;
            else if (!insert || (em1 | em2) == 0u) { unresolved = true; pend = false; }
            else tgt = __popc(em1) >= __popc(em2) ? 0 : 1;     // least loaded; the home bucket on a tie
        }
        const uint32_t tb = tgt ? b2 : b1;
        const unsigned grp = __match_any_sync(FULL, pend ? tb : (0x80000000u | (uint32_t)lane));
        if (pend) {
            const int rank = __popc(grp & lt);
            uint32_t m = tgt ? em2 : em1;
            if (rank < __popc(m)) {
                for (int r = 0; r < rank; ++r) m &= m - 1u;
                sts_u16(v.stbl + tb * 16u + (uint32_t)(__ffs((int)m) - 1) * 2u, tgt ? e2 : e1);
                fresh = true;
                pend = false;
            }
        }
        __syncwarp();   // orders the stores before the next round's (and the next chunk's) bucket loads
        if (!__any_sync(FULL, pend)) break;
    }
    return fresh;
}
// clears the warp's q16 table in global memory (start of a query; hybrid: when the shared table closes)
__device__ __forceinline__ void q16_clear(const VisitedSet& v, int lane) {
    uint4* t4 = reinterpret_cast<uint4*>(v.tbl);
    const uint32_t n16 = v.q_bmask + 1u;
    if (LEANN_K1_L2POL & 2) {
        for (uint32_t i = lane; i < n16; i += 32) st_u4_evict_last(t4 + i, make_uint4(VIS_EMPTY, VIS_EMPTY, VIS_EMPTY, VIS_EMPTY));
    } else {
        for (uint32_t i = lane; i < n16; i += 32) t4[i] = make_uint4(VIS_EMPTY, VIS_EMPTY, VIS_EMPTY, VIS_EMPTY);
    }
    __syncwarp();
}

// warp: take a byte map from the pool and replay the table into it (the set stays exact for any query)
__device__ __forceinline__ void visited_spill(VisitedSet& v, uint32_t warp_id, int lane) {
    int slot = 0;
    if (lane == 0) {
        uint32_t i = warp_id % v.n_slots;
        while (atomicCAS(v.pool_locks + i, 0u, 1u) != 0u) { i = (i + 1u) % v.n_slots; __nanosleep(64); }
        __threadfence();
        slot = (int)i;
    }
    slot = __shfl_sync(FULL, slot, 0);
    v.slot = slot;
    v.vis = v.pool_vis + (size_t)slot * v.n_pad;
    v.epoch_slot = v.pool_epochs + slot;
    v.tag = next_epoch(v.epoch_slot, v.vis, v.n_pad, lane);
    if (v.smv) {
        for (uint32_t b = lane; b <= v.s_bmask; b += 32) {
            const uint4 w = lds_u4(v.stbl + b * 16u);
            const uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t e = (i & 1) ? (x[i >> 1] >> 16) : (x[i >> 1] & 0xFFFFu);
                if (e != Q_EMPTY) {
                    const uint32_t rem = e >> 1;
                    const uint32_t home = (e & 1u) ? (b ^ smv_alt(rem, v.s_bmask)) : b;
                    const uint32_t h = (home << v.s_rem_bits) | rem;
                    v.vis[(h * v.q_inv) & v.q_kmask] = v.tag;
                }
            }
        }
    }
    if (v.q16 && (!v.smv || v.ovf)) {
        for (uint32_t b = lane; b <= v.q_bmask; b += 32) {
            const uint4 w = q_bucket(v, b);
            const uint32_t x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t e = (i & 1) ? (x[i >> 1] >> 16) : (x[i >> 1] & 0xFFFFu);
                if (e != Q_EMPTY) {
                    const uint32_t home = (b - (e & 3u)) & v.q_bmask;
                    const uint32_t h = (home << v.q_rem_bits) | (e >> 2);
                    v.vis[(h * v.q_inv) & v.q_kmask] = v.tag;
                }
            }
        }
    }
    if (!v.smv && !v.q16) {
        for (uint32_t i = lane; i <= v.cap_mask; i += 32) {
            const uint32_t s = __ldcg(v.tbl + i);
            if (s != VIS_EMPTY) v.vis[s] = v.tag;
        }
    }
    __syncwarp();
    v.hashed = false;
}

__device__ __forceinline__ void visited_begin(VisitedSet& v, int lane) {
    v.count = 0;
    v.slot = -1;
    v.hashed = v.tbl != nullptr || v.smv;
    v.ovf = false;
    v.s_count = 0;
    if (v.smv) {   // (hybrid: the q16 level is cleared when the shared table closes)
        for (uint32_t i = lane; i <= v.s_bmask; i += 32) sts_u4(v.stbl + i * 16u, make_uint4(VIS_EMPTY, VIS_EMPTY, VIS_EMPTY, VIS_EMPTY));
        __syncwarp();
    } else if (v.q16) {
        q16_clear(v, lane);
    } else if (v.hashed) {
        uint4* t4 = reinterpret_cast<uint4*>(v.tbl);
        const uint32_t n16 = v.cap_mask / 4 + 1u;   // 16-byte units to clear
        for (uint32_t i = lane; i < n16; i += 32) t4[i] = make_uint4(VIS_EMPTY, VIS_EMPTY, VIS_EMPTY, VIS_EMPTY);
        __syncwarp();
    } else {
        v.tag = next_epoch(v.epoch_slot, v.vis, v.n_pad, lane);
    }
}
// per lane: true when `s` was not in the set (and now is). v.hashed is warp-uniform. (u32 table / byte map; the q16 table
// goes through q_test_and_set.)
__device__ __forceinline__ bool visited_test_and_set(VisitedSet& v, uint32_t s) {
    if (v.hashed) {
        uint32_t h = (s * 0x9E3779B1u) >> v.shift;
        for (;;) {
            const uint32_t old = atomicCAS(v.tbl + h, VIS_EMPTY, s);
            if (old == VIS_EMPTY) return true;
            if (old == s) return false;
            h = (h + 1u) & v.cap_mask;
        }
    }
    if (v.vis[s] != v.tag) { v.vis[s] = v.tag; return true; }
    return false;
}
// warp: account for `added` new members; when the table passes its limit, move to a pooled byte map
__device__ __forceinline__ void visited_added(VisitedSet& v, uint32_t added, uint32_t warp_id, int lane) {
    v.count += added;
    if (v.hashed && v.count - v.s_count > v.limit) visited_spill(v, warp_id, lane);   // s_count: members held by the hybrid's shared level
}
__device__ __forceinline__ void visited_end(VisitedSet& v, int lane) {
    if (v.slot >= 0) {
        __syncwarp();
        if (lane == 0) { __threadfence(); atomicExch(v.pool_locks + v.slot, 0u); }
        v.slot = -1;
    }
}

// Beam search on one level (usearch search_to_find_in_base_ / search_to_insert_, diskann-rs
// search_with_dists). radius = worst distance in `top` once it holds `ef` entries, +inf before.
//   nonstrict == 0 : stop when cand.d >  radius   (usearch)
//   nonstrict == 1 : stop when cand.d >= radius   (diskann-rs; only reachable when top is full)
// mask: nullable; nodes failing it are traversed but never enter `top` (usearch predicate shape).
template <int LPV, int VPL, int U, bool PREFETCH = false>
__device__ __forceinline__ void beam_level(const GraphView& g, const LevelAdj adj, const float4 (&q)[VPL],
                                           WarpLists& w, int ef, int next_cap, int next_mask, int nonstrict,
                                           const uint64_t* __restrict__ mask, VisitedSet& vs, uint32_t warp_id,
                                           uint32_t start, float start_d, Counters& c, int lane,
                                           const Coop cp = Coop{nullptr, 0, 1}) {
    auto passes = [&](uint32_t s) { return mask == nullptr || ((mask[s >> 6] >> (s & 63u)) & 1ull); };
    w.top_size = 0; w.next_size = 0; w.next_head = 0;
    float radius = CUDART_INF_F;
    sorted_insert<compat::NEXT_FIFO_AMONG_EQUALS, true>(w.next_d, w.next_s, w.next_size, next_cap, w.next_head, next_mask, start_d, start, lane);
    if (lane == 0) visited_test_and_set(vs, start);
    visited_added(vs, 1u, warp_id, lane);
    if (passes(start)) sorted_insert<!compat::TOP_NEWCOMER_BEFORE_EQUALS, false>(w.top_d, w.top_s, w.top_size, ef, 0, 0, start_d, start, lane);
    if (w.top_size == ef) radius = w.top_d[ef - 1];
    __syncwarp();
    while (w.next_size > 0) {
        float cd = w.next_d[w.next_head & next_mask];
        uint32_t cs = w.next_s[w.next_head & next_mask];
        if (nonstrict ? (cd >= radius) : (cd > radius)) break;
        w.next_head = (w.next_head + 1) & next_mask;
        w.next_size--;
        if (adj.level == 0) c.n_hops0++; else c.n_hops_upper++;
        // ---- adjacency row -> unvisited neighbours, list order preserved ----
        const uint32_t* row = adj.row(cs);
        if (PREFETCH && adj.level == 0 && w.next_size > 0) {
            // the most likely next pop is the new head of the queue: pull its adjacency row into L2 now
            const uint32_t nh = w.next_s[w.next_head & next_mask];
            if ((uint32_t)lane * 32u < adj.deg) prefetch_l2(adj.adj0 + (size_t)nh * adj.deg + lane * 32);
        }
        int cnt = 0;
        if (PREFETCH) {
            // short rows (latency-bound): the visited tags of the whole row are requested together, one memory round trip per
            // hop instead of one per 32 neighbours; membership is decided per lane, the compaction keeps list order
            constexpr int NCH = MAX_DEG / 32;
            uint32_t sv[NCH];
            uint8_t tg[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const uint32_t j = (uint32_t)ch * 32u + lane;
                sv[ch] = j < adj.deg ? __ldg(row + j) : SENT;
            }
            bool fr[NCH];
            if (vs.hashed) {
                // table probes of the whole row run as interleaved CAS chains
                uint32_t h[NCH];
                bool act[NCH];
                bool any = false;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) { h[ch] = (sv[ch] * 0x9E3779B1u) >> vs.shift; act[ch] = sv[ch] != SENT; fr[ch] = false; any |= act[ch]; }
                while (any) {
                    uint32_t old[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) if (act[ch]) old[ch] = atomicCAS(vs.tbl + h[ch], VIS_EMPTY, sv[ch]);
                    any = false;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch)
                        if (act[ch]) {
                            if (old[ch] == VIS_EMPTY) { fr[ch] = true; act[ch] = false; }
                            else if (old[ch] == sv[ch]) act[ch] = false;
                            else { h[ch] = (h[ch] + 1u) & vs.cap_mask; any = true; }
                        }
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) tg[ch] = sv[ch] != SENT ? vs.vis[sv[ch]] : vs.tag;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    fr[ch] = tg[ch] != vs.tag;
                    if (fr[ch]) vs.vis[sv[ch]] = vs.tag;
                }
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                if ((uint32_t)ch * 32u >= adj.deg) break;
                const bool fresh = fr[ch];
                unsigned b = __ballot_sync(FULL, fresh);
                if (fresh) w.st_slot[cnt + __popc(b & ((1u << lane) - 1u))] = sv[ch];
                cnt += __popc(b);
            }
        } else {
            for (uint32_t base = 0; base < adj.deg; base += 32) {
                uint32_t j = base + lane;
                uint32_t s = j < adj.deg ? __ldg(row + j) : SENT;
                const bool fresh = s != SENT && visited_test_and_set(vs, s);
                unsigned b = __ballot_sync(FULL, fresh);
                if (fresh) w.st_slot[cnt + __popc(b & ((1u << lane) - 1u))] = s;
                cnt += __popc(b);
            }
        }
        __syncwarp();
        c.n_dist += cnt;
        visited_added(vs, (uint32_t)cnt, warp_id, lane);
        if (PREFETCH) {
            // rows beyond the first register batch are requested from HBM now, so that the later batches of
            // eval_distances find them in L2 (one DRAM latency per hop instead of one per batch)
            constexpr int BATCH = U * (32 / LPV);
            const uint32_t lines = (g.d4 * 16u + 127u) >> 7;
            for (int j = BATCH + lane; j < cnt; j += 32) {
                const char* r = reinterpret_cast<const char*>(g.vecs + (size_t)w.st_slot[j] * g.d4);
                for (uint32_t l = 0; l < lines; ++l) prefetch_l2(r + l * 128u);
            }
        }
        coop_eval<LPV, VPL, U>(cp, g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, cnt, lane);
        // ---- replay the inserts in list order ----
        for (int base = 0; base < cnt; base += 32) {
            int j = base + lane;
            float dj = j < cnt ? w.st_dist[j] : CUDART_INF_F;
            bool maybe = j < cnt && (w.top_size < ef || dj < radius);
            unsigned m = __ballot_sync(FULL, maybe);
            while (m) {
                int l = __ffs(m) - 1;
                m &= m - 1;
                float dd = __shfl_sync(FULL, dj, l);
                if (w.top_size < ef || dd < radius) {
                    uint32_t ss = w.st_slot[base + l];
                    if (w.next_size == next_cap) c.dropped = 1;
                    sorted_insert<compat::NEXT_FIFO_AMONG_EQUALS, true>(w.next_d, w.next_s, w.next_size, next_cap, w.next_head, next_mask, dd, ss, lane);
                    if (passes(ss)) sorted_insert<!compat::TOP_NEWCOMER_BEFORE_EQUALS, false>(w.top_d, w.top_s, w.top_size, ef, 0, 0, dd, ss, lane);
                    radius = w.top_size == ef ? w.top_d[ef - 1] : CUDART_INF_F;
                }
            }
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Register-resident lists (short rows). At d <= 256 a hop moves little data and the traversal is bound by issued
// instructions: the shared-memory sorted inserts above (rank scan + shift loop, two __syncwarp per 32 entries) were
// about two thirds of all instructions of a hop at d = 96. Here `top` and `next` live in registers, EPL entries per
// lane in blocked layout (entry i = lane i / EPL, register i % EPL), ascending, unused entries = +inf:
//   rank    = EPL compares per lane + one ballot + one shuffle,
//   insert  = one shuffle per array (carry from the previous lane) + predicated moves,
//   pop     = one shuffle per array.
// Semantics are those of sorted_insert<UPPER, ...> (usearch sorted_buffer_gt::insert with a bounded size): identical
// results, bit for bit. Used when ef and the queue capacity fit in 32 * EPL entries and no inline mask is set.
template <int EPL>
struct RegList {
    float d[EPL];
    uint32_t s[EPL];
    int size;
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int j = 0; j < EPL; ++j) { d[j] = CUDART_INF_F; s[j] = SENT; }
        size = 0;
    }
    // #entries < dd (UPPER: <= dd). Entries are sorted, so lanes whose EPL entries all count form a prefix.
    template <bool UPPER>
    __device__ __forceinline__ int rank(float dd) const {
        int c = 0;
#pragma unroll
        for (int j = 0; j < EPL; ++j) c += (UPPER ? (d[j] <= dd) : (d[j] < dd)) ? 1 : 0;
        const int full = __popc(__ballot_sync(FULL, c == EPL));
        const int rem = __shfl_sync(FULL, c, full & 31);
        return full == 32 ? 32 * EPL : full * EPL + rem;
    }
    // insert at `pos` (< limit), dropping the last entry when the list already holds `limit`.
    // Written with value selects only: conditional stores into d[j] / s[j] get merged by the compiler into one store
    // through a computed address, which sends the arrays to local memory.
    // DROP = false (single-list mode, where nothing is ever popped): the entry pushed past `limit` is left in place; it is
    // never read (readers respect `size`) and cannot disturb rank() because inserts are gated by d < d[limit - 1].
    template <bool DROP = true>
    __device__ __forceinline__ void insert_at(int pos, float dd, uint32_t ss, int limit, int lane) {
        const int lp = pos / EPL, rp = pos % EPL;
        const float cd = __shfl_up_sync(FULL, d[EPL - 1], 1);
        const uint32_t cs = __shfl_up_sync(FULL, s[EPL - 1], 1);
        const bool after = lane > lp, here = lane == lp;
        const bool full = size == limit;
        const int ll = limit / EPL, rl = limit % EPL;
#pragma unroll
        for (int j = EPL - 1; j >= 0; --j) {
            const bool shift = after || (here && j > rp);
            const float pd = j == 0 ? cd : d[j > 0 ? j - 1 : 0];
            const uint32_t ps = j == 0 ? cs : s[j > 0 ? j - 1 : 0];
            float nd = shift ? pd : d[j];
            uint32_t ns = shift ? ps : s[j];
            const bool ins = here && j == rp;
            nd = ins ? dd : nd;
            ns = ins ? ss : ns;
            const bool drop = DROP && full && limit < 32 * EPL && lane == ll && j == rl;   // the entry pushed past the bound
            d[j] = drop ? CUDART_INF_F : nd;
            s[j] = drop ? SENT : ns;
        }
        size += full ? 0 : 1;
    }
    template <bool UPPER, bool DROP = true>
    __device__ __forceinline__ bool insert(float dd, uint32_t ss, int limit, int lane) {
        const int pos = rank<UPPER>(dd);
        if (pos >= limit) return false;
        insert_at<DROP>(pos, dd, ss, limit, lane);
        return true;
    }
    __device__ __forceinline__ float dist_at(int i) const {
        const int r = i % EPL;
        float v = d[0];
#pragma unroll
        for (int j = 1; j < EPL; ++j) v = (r == j) ? d[j] : v;
        return __shfl_sync(FULL, v, i / EPL);
    }
    __device__ __forceinline__ void pop_front(int lane) {
        const float nd = __shfl_down_sync(FULL, d[0], 1);
        const uint32_t ns = __shfl_down_sync(FULL, s[0], 1);
#pragma unroll
        for (int j = 0; j + 1 < EPL; ++j) { d[j] = d[j + 1]; s[j] = s[j + 1]; }
        d[EPL - 1] = lane == 31 ? CUDART_INF_F : nd;
        s[EPL - 1] = lane == 31 ? SENT : ns;
        size--;
    }
};

// beam_level with register lists: same loop as beam_level<..., PREFETCH = true> without a mask.
//
// SINGLE (diskann-rs stop rule only): one list. Without a mask every queued candidate was inserted into `top` at the same
// moment, so the queue is `top` minus the entries already expanded, plus entries `top` has evicted since. An evicted entry
// has d >= the current worst of a full `top`, and search_with_dists stops as soon as the queue head has d >= that worst: an
// evicted entry can only ever end the search, which the first unexpanded entry of `top` (or its absence) decides just as
// well. The queue therefore becomes one "expanded" bit per entry of `top` (bit 31 of the slot, slots stay below 2^31):
// one sorted insert per accepted neighbour instead of two. Equal distances: the queue is FIFO among equals while `top` puts a
// newcomer first, so the oldest of an equal run is its LAST unexpanded entry — that one is popped. Results are identical to the
// two-list form bit for bit (the randomised parity tests include duplicate rows); the queue can no longer overflow, so the
// `dropped` counter stays 0.
constexpr uint32_t XBIT = 0x80000000u;
template <int LPV, int VPL, int U, int EPL, bool SINGLE, bool Q16, bool SMV = false, bool RING = false>
__device__ __forceinline__ int beam_level_regs(const GraphView& g, const LevelAdj adj, const float4 (&q)[VPL], WarpLists& w,
                                               int ef, int next_cap, int nonstrict, VisitedSet& vs, uint32_t warp_id,
                                               uint32_t start, float start_d, Counters& c, int lane,
                                               uint32_t k, uint64_t* __restrict__ out_keys, float* __restrict__ out_dists,
                                               RowRing* ring = nullptr) {
    RegList<EPL> top, next;
    top.clear();
    if (!SINGLE) next.clear();
    float radius = CUDART_INF_F;
    if (!SINGLE) next.template insert<compat::NEXT_FIFO_AMONG_EQUALS>(start_d, start, next_cap, lane);
    if (SMV && vs.hashed) {
        bool f;
        smv_test_and_set(vs, lane == 0 ? start : SENT, lane, f, true);
        if (__any_sync(FULL, f)) {
            // the one key per 2^15 that the shared table cannot hold: the hybrid opens its q16 level for it, the stand-alone
            // form moves to a byte map
            if (Q16) { vs.ovf = true; q16_clear(vs, lane); if (lane == 0) q_test_and_set(vs, start); }
            else { visited_spill(vs, warp_id, lane); if (lane == 0) vs.vis[start] = vs.tag; }
        } else if (Q16) {
            vs.s_count = 1;
        }
    } else if (lane == 0) {
        if (Q16 && vs.hashed) q_test_and_set(vs, start); else visited_test_and_set(vs, start);   // empty table: cannot fail
    }
    __syncwarp();
    visited_added(vs, 1u, warp_id, lane);
    top.template insert<!compat::TOP_NEWCOMER_BEFORE_EQUALS, !SINGLE>(start_d, start, ef, lane);
    if (top.size == ef) radius = top.dist_at(ef - 1);
    for (;;) {
        float cd;
        uint32_t cs;
        if (SINGLE) {
            // queue head = first unexpanded entry of `top`
            unsigned mine = 0;
#pragma unroll
            for (int j = 0; j < EPL; ++j) mine |= (lane * EPL + j < top.size && !(top.s[j] & XBIT)) ? (1u << j) : 0u;
            const unsigned b = __ballot_sync(FULL, mine != 0);
            if (!b) break;
            const int L = __ffs(b) - 1;
            const int j0 = __ffs(__shfl_sync(FULL, mine, L)) - 1;
            cd = top.dist_at(L * EPL + j0);
            if (top.size >= ef && cd >= radius) break;
            // FIFO among equal distances: the oldest of the run is its last unexpanded entry
            unsigned eq = 0;
#pragma unroll
            for (int j = 0; j < EPL; ++j) eq |= (((mine >> j) & 1u) && top.d[j] == cd) ? (1u << j) : 0u;
            const unsigned be = __ballot_sync(FULL, eq != 0);
            const int H = 31 - __clz(be);
            const int jH = 31 - __clz(__shfl_sync(FULL, eq, H));
            uint32_t sv0 = top.s[0];
#pragma unroll
            for (int j = 1; j < EPL; ++j) sv0 = (j == jH) ? top.s[j] : sv0;
            cs = __shfl_sync(FULL, sv0, H);
#pragma unroll
            for (int j = 0; j < EPL; ++j) top.s[j] |= (lane == H && j == jH) ? XBIT : 0u;
            mine &= (lane == H) ? ~(1u << jH) : ~0u;
            if (adj.level == 0) {
                // the most likely next pop is the next unexpanded entry: pull its adjacency row into L2 now
                const unsigned b2 = __ballot_sync(FULL, mine != 0);
                if (b2) {
                    const int L2 = __ffs(b2) - 1;
                    const int j2 = __ffs(__shfl_sync(FULL, mine, L2)) - 1;
                    uint32_t sn = top.s[0];
#pragma unroll
                    for (int j = 1; j < EPL; ++j) sn = (j == j2) ? top.s[j] : sn;
                    const uint32_t nh = __shfl_sync(FULL, sn, L2);
                    if ((uint32_t)lane * 32u < adj.deg) prefetch_l2(adj.adj0 + (size_t)nh * adj.deg + lane * 32);
                }
            }
        } else {
            if (next.size == 0) break;
            cd = __shfl_sync(FULL, next.d[0], 0);
            cs = __shfl_sync(FULL, next.s[0], 0);
            if (nonstrict ? (cd >= radius) : (cd > radius)) break;
            next.pop_front(lane);
            if (adj.level == 0 && next.size > 0) {
                // the most likely next pop is the new head of the queue: pull its adjacency row into L2 now
                const uint32_t nh = __shfl_sync(FULL, next.s[0], 0);
                if ((uint32_t)lane * 32u < adj.deg) prefetch_l2(adj.adj0 + (size_t)nh * adj.deg + lane * 32);
            }
        }
        if (adj.level == 0) c.n_hops0++; else c.n_hops_upper++;
        const uint32_t* row = adj.row(cs);
        int cnt = 0;
        // the row is handled 64 neighbours at a time (one pass for the usual degrees M0 = 64 / R = 64): the visited tags of a
        // pass are requested together, membership is decided per lane, the compaction keeps list order
        constexpr int NCH = 2;
        for (uint32_t c0 = 0; c0 < adj.deg; c0 += 32u * NCH) {
            uint32_t sv[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const uint32_t j = c0 + (uint32_t)ch * 32u + lane;
                sv[ch] = j < adj.deg ? ((LEANN_K1_L2POL & 4) ? ldg_u32_evict_first(row + j) : __ldg(row + j)) : SENT;
            }
#if LEANN_K1_SPEC > 0
            // Speculative row prefetch: the first LEANN_K1_SPEC neighbours of the row (list order) are the most likely members of the
            // first distance batch; their rows are requested from DRAM now, so that the fetch overlaps the two L2 round trips of
            // the visited test instead of following them. A neighbour that turns out to be visited costs one wasted row.
            if (c0 == 0 && lane < LEANN_K1_SPEC && sv[0] != SENT) {
                const char* r = reinterpret_cast<const char*>(g.vecs + (size_t)sv[0] * g.d4);
                const uint32_t lines = (g.d4 * 16u + 127u) >> 7;
                for (uint32_t l = 0; l < lines; ++l) prefetch_l2(r + l * 128u);
            }
#endif
            bool fr[NCH], go[NCH];
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) { fr[ch] = false; go[ch] = sv[ch] != SENT; }
            const bool hashed0 = vs.hashed;
            bool q16_stage = Q16 && hashed0;
            if (SMV && hashed0) {
                // shared-memory table: the chunks go one after the other (a chunk's stores are visible to the next)
                bool anygo = false;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) { fr[ch] = smv_test_and_set(vs, sv[ch], lane, go[ch], Q16 ? !vs.ovf : true); anygo |= go[ch]; }
                anygo = __any_sync(FULL, anygo);
                if (Q16) {
                    // hybrid: unresolved keys go on to the q16 level; the shared level closes at its limit or when a key found no room
                    if (!vs.ovf) {
#pragma unroll
                        for (int ch = 0; ch < NCH; ++ch) vs.s_count += (uint32_t)__popc(__ballot_sync(FULL, fr[ch]));
                        if (anygo || vs.s_count > vs.s_limit) { vs.ovf = true; q16_clear(vs, lane); }
                    }
                    q16_stage = anygo;
                } else if (anygo) {
                    // both buckets of some key are full: continue this query on a pooled byte map (the table is replayed into it)
                    visited_spill(vs, warp_id, lane);
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        if (go[ch] && vs.vis[sv[ch]] != vs.tag) { vs.vis[sv[ch]] = vs.tag; fr[ch] = true; }
                        __syncwarp();
                    }
                }
            }
            if (q16_stage) {
                // bucketed table: all probes of the pass advance together, one bucket load (or one CAS) per round
                unsigned short* t16 = reinterpret_cast<unsigned short*>(vs.tbl);
                uint32_t bk[NCH], want[NCH], emp[NCH], disp[NCH];
                bool act[NCH], need_load[NCH];
                bool any = false, failed = false;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    const uint32_t h = (sv[ch] * Q_HASH_MUL) & vs.q_kmask;
                    bk[ch] = h >> vs.q_rem_bits;
                    want[ch] = (h & ((1u << vs.q_rem_bits) - 1u)) << 2;
                    disp[ch] = 0; emp[ch] = 0;
                    act[ch] = go[ch]; need_load[ch] = true;
                    any |= act[ch];
                }
                while (any) {
                    uint4 w[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) if (act[ch] && need_load[ch]) w[ch] = q_bucket(vs, bk[ch] & vs.q_bmask);
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch)
                        if (act[ch] && need_load[ch]) {
                            bool found;
                            q_scan(w[ch], want[ch] | disp[ch], found, emp[ch]);
                            need_load[ch] = false;
                            if (found) act[ch] = false;
                        }
                    unsigned short old[NCH];
                    int e[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        e[ch] = -1;
                        if (act[ch] && emp[ch]) {
                            e[ch] = __ffs((int)emp[ch]) - 1;
                            emp[ch] &= emp[ch] - 1;
                            old[ch] = atomicCAS(t16 + (bk[ch] & vs.q_bmask) * 8 + e[ch], (unsigned short)Q_EMPTY, (unsigned short)(want[ch] | disp[ch]));
                        }
                    }
                    any = false;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch)
                        if (act[ch]) {
                            if (e[ch] >= 0) {
                                if (old[ch] == Q_EMPTY) { fr[ch] = true; act[ch] = false; }
                                else if (old[ch] == (want[ch] | disp[ch])) act[ch] = false;   // cannot happen with duplicate-free rows; harmless
                            }
                            if (act[ch] && emp[ch] == 0) {      // bucket full (now): the key lives in, or goes to, the next one
                                if (disp[ch] == 3) { failed = true; act[ch] = false; }
                                else { disp[ch]++; bk[ch]++; need_load[ch] = true; }
                            }
                            any |= act[ch];
                        }
                }
                if (__any_sync(FULL, failed)) {
                    // no room within 3 buckets for some key: continue this query on a pooled byte map. Keys already claimed in
                    // this pass are in the table and therefore in the replay; the ones that failed are decided on the map.
                    visited_spill(vs, warp_id, lane);
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch)
                        if (failed && sv[ch] != SENT && !fr[ch] && vs.vis[sv[ch]] != vs.tag) {
                            // (a lane may reach here for a key that was found in the table; the map agrees, nothing to do)
                            vs.vis[sv[ch]] = vs.tag;
                            fr[ch] = true;
                        }
                    __syncwarp();
                }
            } else if (!Q16 && !SMV && hashed0) {
                uint32_t h[NCH];
                bool act[NCH];
                bool any = false;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) { h[ch] = (sv[ch] * 0x9E3779B1u) >> vs.shift; act[ch] = sv[ch] != SENT; fr[ch] = false; any |= act[ch]; }
                while (any) {
                    uint32_t old[NCH];
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) if (act[ch]) old[ch] = atomicCAS(vs.tbl + h[ch], VIS_EMPTY, sv[ch]);
                    any = false;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch)
                        if (act[ch]) {
                            if (old[ch] == VIS_EMPTY) { fr[ch] = true; act[ch] = false; }
                            else if (old[ch] == sv[ch]) act[ch] = false;
                            else { h[ch] = (h[ch] + 1u) & vs.cap_mask; any = true; }
                        }
                }
            } else if (!hashed0) {
                uint8_t tg[NCH];
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) tg[ch] = sv[ch] != SENT ? vs.vis[sv[ch]] : vs.tag;
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    fr[ch] = tg[ch] != vs.tag;
                    if (fr[ch]) vs.vis[sv[ch]] = vs.tag;
                }
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const bool fresh = fr[ch];
                const unsigned b = __ballot_sync(FULL, fresh);
                if (fresh) w.st_slot[cnt + __popc(b & ((1u << lane) - 1u))] = sv[ch];
                cnt += __popc(b);
            }
        }
        __syncwarp();
        c.n_dist += cnt;
        visited_added(vs, (uint32_t)cnt, warp_id, lane);
        {
            // rows beyond the first batch (ring: beyond the stages requested at once) are pulled into L2 now
            constexpr int BATCH = RING ? RING_ROWS * RING_STAGES : U * (32 / LPV);
            const uint32_t lines = (g.d4 * 16u + 127u) >> 7;
            for (int j = BATCH + lane; j < cnt; j += 32) {
                const char* r = reinterpret_cast<const char*>(g.vecs + (size_t)w.st_slot[j] * g.d4);
                for (uint32_t l = 0; l < lines; ++l) prefetch_l2(r + l * 128u);
            }
        }
        if constexpr (RING) eval_distances_ring<LPV, VPL>(*ring, g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, cnt, lane);
        else eval_distances<LPV, VPL, U>(g.vecs, g.d4, g.metric, q, w.st_slot, w.st_dist, cnt, lane);
        // ---- replay the inserts in list order (usearch loop), lists in registers ----
        for (int base = 0; base < cnt; base += 32) {
            const int j = base + lane;
            const float dj = j < cnt ? w.st_dist[j] : CUDART_INF_F;
            const uint32_t sj = j < cnt ? w.st_slot[j] : SENT;
            unsigned m = __ballot_sync(FULL, j < cnt && (top.size < ef || dj < radius));
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const float dd = __shfl_sync(FULL, dj, l);
                if (top.size < ef || dd < radius) {
                    const uint32_t ss = __shfl_sync(FULL, sj, l);
                    if (!SINGLE) {
                        if (next.size == next_cap) c.dropped = 1;
                        next.template insert<compat::NEXT_FIFO_AMONG_EQUALS>(dd, ss, next_cap, lane);
                    }
                    top.template insert<!compat::TOP_NEWCOMER_BEFORE_EQUALS, !SINGLE>(dd, ss, ef, lane);
                    if (top.size == ef) {
                        radius = top.dist_at(ef - 1);
                        m &= __ballot_sync(FULL, dj < radius);   // later entries that can no longer pass are skipped now
                    }
                }
            }
        }
        __syncwarp();
    }
    // results: ascending, truncated to k; tail = UINT64_MAX / +inf (entry i = lane i / EPL, register i % EPL)
    const int cnt = top.size < (int)k ? top.size : (int)k;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
        const uint32_t i = (uint32_t)(lane * EPL + j);
        if (i < k) {
            const bool ok = (int)i < cnt;
            const uint32_t slot = SINGLE ? (top.s[j] & ~XBIT) : top.s[j];
            out_keys[i] = ok ? (g.keys ? g.keys[slot] : (uint64_t)slot) : ~0ull;
            out_dists[i] = ok ? top.d[j] : CUDART_INF_F;
        }
    }
    return cnt;
}

}  // namespace leann
