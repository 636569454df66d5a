// exact_scan_tc.cu — K2 on the 5th-generation tensor cores: the dense Q·Xᵀ contraction of the exact
// scan (leann-rs src/index/recompute.rs:96-103) as a tcgen05.mma kernel fed by TMA, followed by K2r, the
// fp32 re-rank of the surviving candidates (recompute.rs:137-139 arithmetic, f32 throughout).
//
//   pass 1 (scan_tc_kernel): bf16 copies of Q and X. One CTA pair (two SMs of a TPC, cta_group::2) per work
//        item = (256 queries, group of <= 32 database tiles of 256 rows); UMMA 256x256x16 (kind::f16, f32
//        accumulate in TMEM) issued by one thread of the leader CTA. Each CTA keeps its 128 x d query tile
//        resident in 128B-swizzled shared memory (d <= 448; streamed through the ring above that) and streams
//        its 128-row half of every database tile through a TMA ring of up to 8 stages; the peer's TMA
//        completes on the leader's mbarrier, tcgen05.commit multicast frees ring slots in both CTAs; two
//        256-column TMEM accumulator stages so the epilogue of tile t overlaps the MMAs of tile t+1.
//        Epilogue: tcgen05.ld 32x32b.x32, one thread per query row, survivor bitmask by sign(T_q - score),
//        row ids of the survivors appended in batches of TC_QBUF.
//        T_q = (exact k-th best so far) - eps_q, and eps_q bounds |bf16 score - f32 score| rigorously:
//          |q^.x^ - q.x| = |(q^-q).x + q^.(x^-x)| <= |q^-q| |x| + |q^| |x^-x|          (Cauchy-Schwarz)
//        with the ACTUAL rounding residuals |q^-q| (per query) and max_rows |x^-x|, max_rows |x| measured when
//        the bf16 copies are made (to_bf16_kernel), plus dp8 * 2^-22 |q^||x^| for the f32 accumulation inside the
//        tensor core. (The worst case of the residual is 2^-8 relative per operand, i.e. (2^-7 + 2^-16)|q||x| in
//        total; the round-1 constant 1.25 * 2^-8 was below that and could drop a true top-k row on low-mantissa
//        inputs such as (1 + 2^-8) c — tests/test_exact_gpu.py::test_tc_prefilter_tie_point_rows.)
//   pass 2 (rerank_kernel): exact f32 dot for each survivor -> packed rank key -> select_kernel.
//
// Warp roles (256 threads per CTA): warp 0 = ring producer (TMA), warp 1 = MMA issuer (leader CTA only),
// warp 2 = TMEM allocator, warp 3 = resident query tile producer, warps 4..7 = epilogue (warp w reads TMEM
// lanes 32*(w%4)..+31). Persistent pairs walk a static schedule of (query pair tile, group of database tiles)
// items ordered so that concurrently running pairs read the same database tiles (L2 reuse across query tiles).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "scan_common.h"

namespace leann {

namespace {

constexpr int TC_M = 128;        // queries per CTA; the CTA pair issues UMMA M = 256
constexpr int TC_N = 256;        // database rows per pair tile (UMMA N); each CTA stages TC_N / 2 of them
constexpr int TC_NH = TC_N / 2;
constexpr int TC_K = 64;         // bf16 elements per k-block = 128 B = one swizzle row
constexpr int TC_A_BYTES = TC_M * TC_K * 2;    // 16 KB: one k-block of this CTA's query tile
constexpr int TC_B_BYTES = TC_NH * TC_K * 2;   // 16 KB: one k-block of this CTA's half of the database tile
constexpr int TC_MAX_RES_KB = 7;  // query tile stays resident in shared memory up to 7 k-blocks (d <= 448)
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 256;
constexpr int TC_QBUF = 8;       // survivors (row id + tensor-core score) buffered per query row before one slot reservation
constexpr int TC_BAR_BYTES = 512;
constexpr size_t TC_SMEM_MAX = 225280;  // 220 KB of the 227 KB opt-in limit: the rest lets a select_kernel<true> / rerank CTA of the
                                        // other query half share the SM (launch_exact_scan)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on a barrier of another CTA of the cluster (address from mapa_u32)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA tile load issued by either CTA of the pair: data lands in this CTA's shared memory, the byte count is
// credited to the barrier at `bar_cluster_addr` (the leader CTA's, where the MMA issuer waits).
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives (once all MMAs issued so far retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows at 128 B pitch, 8-row groups 1024 B apart.
// SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO [16,30) unused for K-major
// swizzled layouts, SBO>>4 [32,46) = 64, version [46,48) = 1, layout_type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// InstrDescriptor: c_format F32 (1) [4,6), a/b format BF16 (1) [7,10)/[10,13), K-major both,
// n_dim = N>>3 [17,23), m_dim = M>>4 [24,29). M = 256: the pair's two 128-row halves.
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)((2 * TC_M) >> 4) << 24);

struct TcParams {
    uint32_t nq, r0, r1, kblocks, n_qpairs, n_groups, tiles_total, group, stages;
    const float* thr_dot;          // [nq] strict candidate threshold in dot space (-inf = everything)
    const uint64_t* mask;          // nullable
    uint32_t* cand_ids;            // [nq][cap]
    float* cand_sc;                // [nq][cap] tensor-core score of each survivor (rerank_kernel's approximate cut)
    uint32_t* cand_cnt;            // [nq]
    uint32_t cap;
    uint32_t* overflow;
};

// One CTA pair (two SMs of a TPC) per cluster. Work item = (256-query pair tile, group of database tiles);
// items are dealt round-robin so that concurrently running pairs read the same database tiles out of L2.
//   RESIDENT: the CTA's 128 x d query tile is loaded once per item and stays in shared memory; the ring
//             streams only this CTA's half of each database tile (16 KB per k-block).
//   !RESIDENT (d > 448): query k-blocks travel through the ring beside the database k-blocks.
template <bool RESIDENT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t STAGE_BYTES = RESIDENT ? TC_B_BYTES : TC_A_BYTES + TC_B_BYTES;
    unsigned char* a_res = smem;                                                        // RESIDENT: [kblocks][16 KB]
    unsigned char* ring = smem + (RESIDENT ? (size_t)p.kblocks * TC_A_BYTES : 0);      // [stages][STAGE_BYTES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * STAGE_BYTES);
    uint64_t* full_bar = bars;                            // [8]  leader's copy is the one waited on
    uint64_t* empty_bar = bars + TC_MAX_STAGES;           // [8]  one per CTA (multicast commit)
    uint64_t* tmem_full = bars + 2 * TC_MAX_STAGES;       // [2]  one per CTA (multicast commit)
    uint64_t* tmem_empty = bars + 2 * TC_MAX_STAGES + 2;  // [2]  leader's copy: 8 epilogue warps arrive
    uint64_t* a_full = bars + 2 * TC_MAX_STAGES + 4;      // [7] per resident k-block; leader's copy
    uint64_t* a_empty = a_full + TC_MAX_RES_KB;           // [7] one per CTA (multicast commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + TC_MAX_RES_KB);
    uint32_t* s_qbuf = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(bars) + TC_BAR_BYTES);  // [2][TC_QBUF][128]: ids, scores

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
        for (int kb = 0; kb < TC_MAX_RES_KB; ++kb) { mbar_init(&a_full[kb], 1); mbar_init(&a_empty[kb], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / complete_tx
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t n_items = p.n_qpairs * p.n_groups;

    if (warp == 0) {
        // ===== ring producer (one thread per CTA) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t item = pair; item < n_items; item += n_pairs) {
                const uint32_t qpair = item % p.n_qpairs, grp = item / p.n_qpairs;
                const uint32_t t0 = grp * p.group, t1 = min(t0 + p.group, p.tiles_total);
                const int qrow = (int)(qpair * (2 * TC_M) + rank * TC_M);
                for (uint32_t t = t0; t < t1; ++t) {
                    const int xrow = (int)(p.r0 + t * TC_N + rank * TC_NH);
                    for (uint32_t kb = 0; kb < p.kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * STAGE_BYTES);
                        const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        unsigned char* dst = ring + (size_t)stage * STAGE_BYTES;
                        if (!RESIDENT) {
                            tma_load_2d_pair(dst, &map_q, bar, (int)(kb * TC_K), qrow);
                            dst += TC_A_BYTES;
                        }
                        tma_load_2d_pair(dst, &map_x, bar, (int)(kb * TC_K), xrow);
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ===== resident query tile producer (one thread per CTA) =====
        if (RESIDENT && lane == 0) {
            uint32_t a_phase = 0;
            for (uint32_t item = pair; item < n_items; item += n_pairs) {
                const uint32_t qpair = item % p.n_qpairs;
                const int qrow = (int)(qpair * (2 * TC_M) + rank * TC_M);
                // k-block kb of the tile is replaced as soon as the previous item's last MMAs on it retire,
                // while that item's remaining k-blocks are still running
                for (uint32_t kb = 0; kb < p.kblocks; ++kb) {
                    mbar_wait(&a_empty[kb], a_phase ^ 1u);
                    if (rank == 0) mbar_arrive_expect_tx(&a_full[kb], 2u * TC_A_BYTES);
                    tma_load_2d_pair(a_res + (size_t)kb * TC_A_BYTES, &map_q, mapa_u32(smem_u32(&a_full[kb]), 0), (int)(kb * TC_K), qrow);
                }
                a_phase ^= 1u;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives both SMs =====
        if (rank == 0 && lane == 0) {
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, a_phase = 0;
            for (uint32_t item = pair; item < n_items; item += n_pairs) {
                const uint32_t grp = item / p.n_qpairs;
                const uint32_t t0 = grp * p.group, t1 = min(t0 + p.group, p.tiles_total);
                for (uint32_t t = t0; t < t1; ++t) {
                    mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TC_N;
                    for (uint32_t kb = 0; kb < p.kblocks; ++kb) {
                        if (RESIDENT && t == t0) mbar_wait(&a_full[kb], a_phase);
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t sr = smem_u32(ring + (size_t)stage * STAGE_BYTES);
                        const uint32_t sa = RESIDENT ? smem_u32(a_res + (size_t)kb * TC_A_BYTES) : sr;
                        const uint32_t sb = RESIDENT ? sr : sr + TC_A_BYTES;
                        const uint64_t da = umma_desc(sa), db = umma_desc(sb);
#pragma unroll
                        for (int k = 0; k < TC_K / 16; ++k)  // advance 16 bf16 = 32 B inside the swizzle row
                            tc_mma_bf16_pair(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), TC_IDESC, (kb | (uint32_t)k) ? 1u : 0u);
                        tc_commit_pair(&empty_bar[stage]);   // frees the slot in both CTAs once these MMAs retire
                        if (RESIDENT && t + 1 == t1) tc_commit_pair(&a_empty[kb]);   // query k-block may be replaced
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                    tc_commit_pair(&tmem_full[acc]);         // accumulator halves ready in both CTAs
                    acc ^= 1u;
                    if (acc == 0) acc_phase ^= 1u;
                }
                a_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: one thread per query row of this CTA's half =====
        const uint32_t quad = (uint32_t)(warp & 3);
        uint32_t acc = 0, acc_phase = 0;
        const uint32_t empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0), empty1 = mapa_u32(smem_u32(&tmem_empty[1]), 0);
        for (uint32_t item = pair; item < n_items; item += n_pairs) {
            const uint32_t qpair = item % p.n_qpairs, grp = item / p.n_qpairs;
            const uint32_t t0 = grp * p.group, t1 = min(t0 + p.group, p.tiles_total);
            const uint32_t q = qpair * (2 * TC_M) + rank * TC_M + quad * 32 + lane;
            const float T = q < p.nq ? p.thr_dot[q] : CUDART_INF_F;   // survivor <=> score > T
            // survivors are buffered per thread and appended with ONE slot reservation per TC_QBUF rows
            uint32_t* myq = s_qbuf + (quad * 32 + lane);   // slot i at myq[i * 128], conflict-free
            uint32_t* mys = myq + TC_QBUF * 128;           // the survivor's score bits
            int nbuf = 0;
            auto flush = [&]() {
                if (nbuf == 0) return;
                uint32_t pos = atomicAdd(&p.cand_cnt[q], (uint32_t)nbuf);
                for (int i = 0; i < nbuf; ++i) {
                    if (pos + i < p.cap) {
                        p.cand_ids[(size_t)q * p.cap + pos + i] = myq[i * 128];
                        p.cand_sc[(size_t)q * p.cap + pos + i] = __uint_as_float(mys[i * 128]);
                    } else {
                        *p.overflow = 1u;
                    }
                }
                nbuf = 0;
            };
            for (uint32_t t = t0; t < t1; ++t) {
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t row_base = p.r0 + t * TC_N;
#pragma unroll 1
                for (int c = 0; c < TC_N; c += 32) {
                    uint32_t r[32];
                    tc_ld32(tmem_base + ((quad * 32u) << 16) + acc * TC_N + (uint32_t)c, r);
                    // bit j of m = sign(T - score_j): two instructions per score, no branches
                    uint32_t m = 0;
#pragma unroll
                    for (int j = 31; j >= 0; --j) m = __funnelshift_l(__float_as_uint(T - __uint_as_float(r[j])), m, 1);
                    while (m) {
                        const uint32_t j = (uint32_t)__ffs((int)m) - 1u;
                        m &= m - 1u;
                        const uint32_t row = row_base + (uint32_t)c + j;
                        if (row < p.r1 && (!p.mask || ((p.mask[row >> 6] >> (row & 63u)) & 1ull))) {
                            uint32_t sc = r[0];   // r[j] by selects: a dynamic index would send the 32 registers to local memory
#pragma unroll
                            for (int jj = 1; jj < 32; ++jj) sc = ((uint32_t)jj == j) ? r[jj] : sc;
                            myq[nbuf * 128] = row;
                            mys[nbuf * 128] = sc;
                            if (++nbuf == TC_QBUF) flush();
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc ? empty1 : empty0);
                acc ^= 1u;
                if (acc == 0) acc_phase ^= 1u;
            }
            flush();
        }
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();   // neither CTA leaves (or frees TMEM) while its partner can still signal it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// f32 -> bf16 (round to nearest even) rows padded to dp8 elements. Per row: |x| and the rounding residual |x^ - x|
// (both rounded up: they are only used as upper bounds), optionally stored, optionally folded into two global maxima
// (max_bits[0] = max |x|, max_bits[1] = max |x^ - x|; non-negative floats order as their bit patterns).
// AUG (squared-L2 metric): the float4 group after the data (columns 4*d4 .. 4*d4+2, inside the dp8 padding) carries the
// row term of |q - x|^2 = |q|^2 + |x|^2 - 2 q.x so that the SAME tensor pass yields q.x - |x|^2/2:
//   rows (aug_rows_kernel, once max|x| is known):  -(h/c) split into three bf16 pieces, h = fl(0.5 * sum of squares);
//              an f32 has 24 significant bits and each piece takes 8, so hi + mid + lo == h/c exactly;
//   queries (AUG_QUERY): c, c, c.
// c = 2^floor(log2 max|x|): a power of two (h/c and c * piece are exact) that keeps the augmented columns at the scale of
// the data, so the f32 accumulation term of the error bound stays proportional to |q| |x| whatever the units of the data.
// AUG_SS: `norms` receives the raw sum of squares instead of the norm (input of aug_rows_kernel).
constexpr int AUG_NONE = 0, AUG_SS = 1, AUG_QUERY = 2;
__device__ __forceinline__ float tc_l2_scale(const uint32_t* xmax_bits) {
    const uint32_t e = xmax_bits[0] & 0x7F800000u;
    return (e == 0u || e == 0x7F800000u) ? 1.0f : __uint_as_float(e);
}
__global__ void to_bf16_kernel(const float4* __restrict__ src, uint32_t d4, __nv_bfloat16* __restrict__ dst, uint32_t dp8,
                               size_t n, float* __restrict__ norms, float* __restrict__ resid, uint32_t* __restrict__ max_bits,
                               int aug, const uint32_t* __restrict__ scale_src) {
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    float ss = 0.f, rr = 0.f;
    if (row < n) {
        for (uint32_t i = lane; i < dp8 / 4; i += 32) {
            float4 v = i < d4 ? src[row * d4 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
            const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
            const float ex = fa.x - v.x, ey = fa.y - v.y, ez = fb.x - v.z, ew = fb.y - v.w;   // exact in f32
            rr += ex * ex + ey * ey + ez * ez + ew * ew;
            reinterpret_cast<__nv_bfloat162*>(dst + row * dp8)[i * 2] = a;
            reinterpret_cast<__nv_bfloat162*>(dst + row * dp8)[i * 2 + 1] = b;
        }
    }
    for (int off = 16; off; off >>= 1) {
        ss += __shfl_xor_sync(0xFFFFFFFFu, ss, off);
        rr += __shfl_xor_sync(0xFFFFFFFFu, rr, off);
    }
    if (aug == AUG_QUERY && row < n && (uint32_t)lane == (d4 & 31u)) {   // this lane stored group d4 (zeros) in the loop above
        const float c = tc_l2_scale(scale_src);
        reinterpret_cast<__nv_bfloat162*>(dst + row * dp8)[d4 * 2] = __floats2bfloat162_rn(c, c);
        reinterpret_cast<__nv_bfloat162*>(dst + row * dp8)[d4 * 2 + 1] = __floats2bfloat162_rn(c, 0.f);
    }
    // 1 + 1e-4 covers the f32 rounding of the sums of squares (<= d * 2^-24 relative, d <= 4096) and of sqrtf
    const float nrm = sqrtf(ss) * 1.0001f, res = sqrtf(rr) * 1.0001f;
    if (row < n && lane == 0) {
        if (norms) norms[row] = aug == AUG_SS ? ss : nrm;
        if (resid) resid[row] = res;
    }
    if (max_bits) {
        __shared__ float s_n[32], s_r[32];
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        if (lane == 0) { s_n[w] = row < n ? nrm : 0.f; s_r[w] = row < n ? res : 0.f; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float mn = 0.f, mr = 0.f;
            for (int i = 0; i < nw; ++i) { mn = fmaxf(mn, s_n[i]); mr = fmaxf(mr, s_r[i]); }
            atomicMax(&max_bits[0], __float_as_uint(mn));
            atomicMax(&max_bits[1], __float_as_uint(mr));
        }
    }
}

// Squared-L2 metric, database side: writes -(h/c) in three exact bf16 pieces into columns 4*d4 .. 4*d4+2 of every row.
__global__ void aug_rows_kernel(__nv_bfloat16* __restrict__ dst, uint32_t dp8, uint32_t d4, size_t n, const float* __restrict__ ss,
                                const uint32_t* __restrict__ xmax_bits) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const float h = 0.5f * ss[row] / tc_l2_scale(xmax_bits);      // exact: a division by a power of two
    const __nv_bfloat16 hi = __float2bfloat16_rn(h);
    const float r1 = h - __bfloat162float(hi);                    // exact
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(mid);                  // exact, <= 8 significant bits: lo is exact
    reinterpret_cast<__nv_bfloat162*>(dst + row * dp8)[d4 * 2] = __floats2bfloat162_rn(-__bfloat162float(hi), -__bfloat162float(mid));
    reinterpret_cast<__nv_bfloat162*>(dst + row * dp8)[d4 * 2 + 1] = __floats2bfloat162_rn(-r2, 0.f);
}

// Rigorous bound on |bf16 tensor-core score - f32 score| of query q against any database row (header comment):
// |q^-q| max|x|  +  |q^| max|x^-x|  +  dp8 * 2^-22 |q^| max|x^|, with |q^| <= |q| (1 + 2^-8), |x^| <= |x| (1 + 2^-8).
__device__ __forceinline__ float tc_eps(float qnorm, float qres, const uint32_t* xmax_bits, uint32_t dp8) {
    const float xmax = __uint_as_float(xmax_bits[0]), xres = __uint_as_float(xmax_bits[1]);
    const float qhat = qnorm * 1.00390625f;
    const float acc = (float)dp8 * 2.3841858e-7f * qhat * (xmax * 1.00390625f);
    return (qres * xmax + qhat * xres + acc) * 1.0001f + 1e-30f;
}

// Squared-L2 metric through the dot-product pass. With the augmented columns the tensor core computes
//   s' ~= s* = q.x - |x|^2/2 = (|q|^2 - D*) / 2,     D* = |q - x|^2 (exact arithmetic),
// and a row must survive whenever its f32 distance D (rerank_kernel, scan_tile_kernel) is <= dk, the exact k-th best:
//   * |D - D*| <= g D* with g = (d/8 + 32) 2^-24 <= 3.3e-5 for d <= 4096 (sums of non-negative terms, <= 5 roundings per
//     128 columns and lane + 5 for the butterfly), so D <= dk implies D* <= dk (1 + 1e-4);
//   * |q|^2 >= L = 0.99975 qnorm^2: qnorm = fl(sqrtf(ss) * 1.0001), ss within g of |q|^2;
//   * |s' - s*| <= eps + herr: eps = the bf16 bound of the data columns (as tc_eps: |q^-q| max|x| + |q^| max|x^-x|) plus the
//     f32 accumulation term over the AUGMENTED vectors (|q^_a|^2 = |q^|^2 + 3 c^2, |x^_a|^2 <= |x^|^2 + 1.01 (hmax/c)^2); the
//     augmented products c * piece are exact and c (hi + mid + lo) == h, so herr = |h - |x|^2/2| <= g h <= 5e-5 hmax.
// Survivor <=> s' > T with T = (L - dk (1 + 1e-4)) / 2 - eps - herr, rounded down.
__device__ __forceinline__ float tc_l2_threshold(float dk, float qnorm, float qres, const uint32_t* xmax_bits, uint32_t dp8) {
    const float xmax = __uint_as_float(xmax_bits[0]), xres = __uint_as_float(xmax_bits[1]);
    const float c = tc_l2_scale(xmax_bits);
    const float hmax = 0.5f * xmax * xmax * 1.000001f, hc = hmax / c;
    const float qhat = qnorm * 1.00390625f, xhat = xmax * 1.00390625f;
    const float qa = sqrtf(qhat * qhat + 3.f * c * c) * 1.000001f, xa = sqrtf(xhat * xhat + 1.01f * hc * hc) * 1.000001f;
    const float acc = (float)dp8 * 2.3841858e-7f * qa * xa;
    const float eps = (qres * xmax + qhat * xres + acc) * 1.0001f + 5e-5f * hmax + 1e-30f;
    const float L = qnorm * qnorm * 0.99975f;
    const float T = 0.5f * (L - dk * 1.0001f) - eps;
    return nextafterf(T - fabsf(T) * 1e-6f - (L + dk) * 1e-6f, -CUDART_INF_F);   // the float operations above round either way
}

// thr key (packed rank key of the exact k-th best, or ~0) -> dot-space candidate threshold; and, per query, the slack of
// rerank_kernel's approximate cut: cut_slack = 2 (eps + delta), eps = the bound on |tensor-core score - exact score| used for
// the threshold, delta = 5e-5 |q| max|x| (L2: 5e-5 (|q| + max|x|)^2) for the f32 rounding of the re-rank itself.
__global__ void dot_threshold_kernel(const unsigned long long* __restrict__ thr, const float* __restrict__ qnorm,
                                     const float* __restrict__ qres, const uint32_t* xmax_bits, uint32_t dp8, uint32_t nq, int metric,
                                     float* __restrict__ thr_dot, float* __restrict__ cut_slack) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float xmax = __uint_as_float(xmax_bits[0]);
    if (metric == LEANN_METRIC_L2SQ) {
        // tc_l2_threshold(0) = L / 2 - eps (rounded down): recover a bound on eps from it
        const float L = qnorm[q] * qnorm[q] * 0.99975f;
        const float eps = (0.5f * L - tc_l2_threshold(0.f, qnorm[q], qres[q], xmax_bits, dp8)) * 1.001f;
        const float qx = qnorm[q] + xmax;
        cut_slack[q] = 2.f * (eps + 5e-5f * qx * qx) * 1.001f;
    } else {
        // IP metrics return 1 - dot: two rows must also differ by more than the rounding of that subtraction, or the cut could
        // drop the lower row id of a pair that ties after rounding
        const float ip = metric == LEANN_METRIC_DOT_DESC ? 0.f : 4.8e-7f * (1.0f + qnorm[q] * xmax);
        cut_slack[q] = 2.f * (tc_eps(qnorm[q], qres[q], xmax_bits, dp8) + 5e-5f * qnorm[q] * xmax + ip) * 1.001f;
    }
    unsigned long long key = thr[q];
    if (key == ~0ull) { thr_dot[q] = -CUDART_INF_F; return; }
    uint32_t ok = (uint32_t)(key >> 32);
    if (metric == LEANN_METRIC_L2SQ) { thr_dot[q] = tc_l2_threshold(scan_unorder_f32(ok), qnorm[q], qres[q], xmax_bits, dp8); return; }
    float dotk;
    if (metric == LEANN_METRIC_DOT_DESC) dotk = scan_unorder_f32(~ok);
    else dotk = 1.0f - scan_unorder_f32(ok);   // IP / IP_CLAMP: distance = 1 - dot (clamped at 0: dot >= 1 stays conservative)
    thr_dot[q] = nextafterf(dotk - tc_eps(qnorm[q], qres[q], xmax_bits, dp8) - fabsf(dotk) * 1e-6f, -CUDART_INF_F);   // strict: the kernel keeps score > thr_dot
}

// K2r: exact f32 score of every survivor -> packed rank key in cand[q][i].
// Approximate cut (cand_sc != nullptr): a round keeps every row whose tensor-core score beats (k-th best so far) - eps,
// about (growth - 1) k rows per query, of which at most k can enter the top-k. With a = the k-th largest tensor-core score
// among the survivors, at least k survivors have an exact score >= a - eps, so a survivor below a - 2 eps (minus the f32
// slack of the re-rank, dot_threshold_kernel) cannot be among the k best: it is dropped instead of costing a 4 d-byte row
// read; the kept keys are written compacted and cand_cnt becomes their number. (Clamped IP distances tie at 0 for every
// dot >= 1: rows that may reach 1 are never cut.) One block per query; the k-th largest is a 4-pass radix select over the scores in global memory (a few hundred values; the block must fit beside a
// resident scan_tc CTA, so nothing is staged in shared memory).
__global__ void __launch_bounds__(256)
rerank_kernel(const float4* __restrict__ X, const float4* __restrict__ Q, uint32_t d4, uint32_t nq, int metric,
              const uint32_t* __restrict__ cand_ids, uint32_t* __restrict__ cand_cnt, uint32_t cap,
              unsigned long long* __restrict__ cand, const float* __restrict__ cand_sc, const float* __restrict__ cut_slack, uint32_t k) {
    __shared__ uint32_t s_hist[259];
    __shared__ uint32_t s_kept;
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t n = cand_cnt[q];
    if (n > cap) n = cap;
    if (threadIdx.x == 0) s_kept = 0;
    __syncthreads();
    float cut = -CUDART_INF_F;
    if (cand_sc != nullptr && n > 2 * k) {   // block-uniform
        const float* sc = cand_sc + (size_t)q * cap;
        uint32_t prefix = 0, remaining = k - 1;   // k-th smallest of ~order(score) = k-th largest score
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t v = ~scan_order_f32(sc[i]);
                if (pass == 0 || (v >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(v >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                uint32_t c[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = s_hist[lane * 8 + j]; sum += c[j]; }
                uint32_t incl = sum;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
                    if (lane >= off) incl += t;
                }
                const uint32_t before = incl - sum;
                if (remaining >= before && remaining < incl) {   // exactly one lane
                    uint32_t acc = before;
                    int j = 0;
                    while (j < 7 && remaining >= acc + c[j]) { acc += c[j]; ++j; }
                    s_hist[256] = (uint32_t)lane * 8u + (uint32_t)j;
                    s_hist[257] = acc;
                }
            }
            __syncthreads();
            prefix |= s_hist[256] << shift;
            remaining -= s_hist[257];
            __syncthreads();
        }
        const float a = scan_unorder_f32(~prefix);
        cut = a - cut_slack[q];
        if (metric == LEANN_METRIC_IP_CLAMP) cut = fminf(cut, 1.0f - cut_slack[q]);
    }
    for (uint32_t i = warp; i < n; i += blockDim.x >> 5) {
        if (cand_sc != nullptr && cand_sc[(size_t)q * cap + i] < cut) continue;
        const uint32_t row = cand_ids[(size_t)q * cap + i];
        const float4* x = X + (size_t)row * d4;
        const float4* qq = Q + (size_t)q * d4;
        float ax = 0.f, ay = 0.f, az = 0.f, aw = 0.f;
        if (metric == LEANN_METRIC_L2SQ) {
            for (uint32_t j = lane; j < d4; j += 32) {
                float4 a = qq[j], b = __ldg(&x[j]);
                const float tx = a.x - b.x, ty = a.y - b.y, tz = a.z - b.z, tw = a.w - b.w;
                ax = fmaf(tx, tx, ax); ay = fmaf(ty, ty, ay); az = fmaf(tz, tz, az); aw = fmaf(tw, tw, aw);
            }
        } else {
            for (uint32_t j = lane; j < d4; j += 32) {
                float4 a = qq[j], b = __ldg(&x[j]);
                ax = fmaf(a.x, b.x, ax); ay = fmaf(a.y, b.y, ay); az = fmaf(a.z, b.z, az); aw = fmaf(a.w, b.w, aw);
            }
        }
        float s = (ax + ay) + (az + aw);
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
        if (lane == 0) {
            uint32_t ok;
            if (metric == LEANN_METRIC_DOT_DESC) ok = ~scan_order_f32(s);
            else if (metric == LEANN_METRIC_L2SQ) ok = scan_order_f32(s);
            else {
                float dd = 1.0f - s;
                if (metric == LEANN_METRIC_IP_CLAMP) dd = dd < 0.f ? 0.f : dd;
                ok = scan_order_f32(dd);
            }
            cand[(size_t)q * cap + atomicAdd(&s_kept, 1u)] = ((unsigned long long)ok << 32) | row;   // compacted: select_kernel sees only the kept keys
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cand_cnt[q] = s_kept;
}

// After the first chunk (scored by the f32 tile kernel, a sequential fold): the rows of the running top-k go back into
// the candidate list so that rerank_kernel re-scores them. From then on every key in `best` carries a score in ONE
// arithmetic (rerank_kernel's), whichever chunk the row came from: exact duplicate rows tie exactly and rank by ascending
// row, as the reference's stable sort over enumerate() order does (recompute.rs:106), instead of differing in the last bit.
__global__ void requeue_best_kernel(const unsigned long long* __restrict__ best, uint32_t* __restrict__ best_cnt, uint32_t kpad,
                                    unsigned long long* __restrict__ thr, uint32_t* __restrict__ cand_ids,
                                    uint32_t* __restrict__ cand_cnt, uint32_t cap, uint32_t nq) {
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    const uint32_t n = min(best_cnt[q], cap);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) cand_ids[(size_t)q * cap + i] = (uint32_t)(best[(size_t)q * kpad + i] & 0xFFFFFFFFull);
    __syncthreads();
    if (threadIdx.x == 0) { cand_cnt[q] = n; best_cnt[q] = 0; thr[q] = ~0ull; }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || !p) throw Error(LEANN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

CUtensorMap make_map(const void* base, uint64_t rows, uint32_t dp8, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {dp8, rows};
    cuuint64_t strides[1] = {(cuuint64_t)dp8 * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(LEANN_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

}  // namespace

bool exact_scan_tc_supported(const FlatView& f, uint32_t nq) {
    return (f.metric == LEANN_METRIC_DOT_DESC || f.metric == LEANN_METRIC_IP || f.metric == LEANN_METRIC_IP_CLAMP ||
            f.metric == LEANN_METRIC_L2SQ) && nq >= 1 && f.d >= 64 && f.n >= 16384;
}

// Columns of the bf16 copies: the data rounded up to 8; the squared-L2 form appends one float4 group (-|x|^2/2 in three
// pieces for rows, ones for queries) after the 4-padded data.
uint32_t exact_scan_tc_dp8(uint32_t d, uint32_t d4, int metric) {
    return metric == LEANN_METRIC_L2SQ ? (d4 * 4 + 4 + 7) / 8 * 8 : (d + 7) / 8 * 8;
}


// Builds the bf16 copy of a database (called once per index) with its row norms (f32, kept for the L2 form of the
// tensor path) and the two maxima tc_eps needs: xmax_bits[0] = max |x|, xmax_bits[1] = max |x^ - x|.
void exact_scan_tc_prepare(const float4* vecs, size_t n, uint32_t d4, uint32_t dp8, void* bf16_rows, float* norms,
                           uint32_t* xmax_bits, int metric, cudaStream_t s) {
    LEANN_CUDA_CHECK(cudaMemsetAsync(xmax_bits, 0, 8, s));
    if (!n) return;
    const bool l2 = metric == LEANN_METRIC_L2SQ;
    to_bf16_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(vecs, d4, (__nv_bfloat16*)bf16_rows, dp8, n, norms, nullptr, xmax_bits,
                                                           l2 ? AUG_SS : AUG_NONE, nullptr);
    if (l2) aug_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((__nv_bfloat16*)bf16_rows, dp8, d4, n, norms, xmax_bits);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

// One round over rows [r0, r1): tensor pass -> re-rank. `s.thr` must hold the current exact thresholds;
// on return s.cand / s.cand_cnt hold exact packed keys ready for select_kernel.
void exact_scan_tc_round(const FlatView& f, const TcIndexView& tv, const ScanScratch& s, const TcScratch& ts, uint32_t nq, uint32_t k,
                         uint32_t r0, uint32_t r1, const uint64_t* d_mask, uint32_t cap, int sms, cudaStream_t stream) {
    dot_threshold_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(s.thr, ts.qnorm, ts.qres, tv.xmax_bits, tv.dp8, nq, f.metric, ts.thr_dot, ts.cut_slack);
    CUtensorMap mq = make_map(ts.q_bf16, nq, tv.dp8, TC_M);
    CUtensorMap mx = make_map(tv.x_bf16, f.n, tv.dp8, TC_NH);
    TcParams p;
    p.nq = nq; p.r0 = r0; p.r1 = r1;
    p.kblocks = (tv.dp8 + TC_K - 1) / TC_K;
    p.n_qpairs = (nq + 2 * TC_M - 1) / (2 * TC_M);
    p.tiles_total = (r1 - r0 + TC_N - 1) / TC_N;
    const bool resident = p.kblocks <= (uint32_t)TC_MAX_RES_KB;
    const size_t a_bytes = resident ? (size_t)p.kblocks * TC_A_BYTES : 0;
    const size_t stage_bytes = resident ? TC_B_BYTES : TC_A_BYTES + TC_B_BYTES;
    const size_t fixed = 1024 /*align*/ + TC_BAR_BYTES + (size_t)2 * 128 * TC_QBUF * 4 + a_bytes;
    p.stages = (uint32_t)std::min<size_t>(TC_MAX_STAGES, (TC_SMEM_MAX - fixed) / stage_bytes);
    const size_t smem = fixed + (size_t)p.stages * stage_bytes;
    const uint32_t max_pairs = (uint32_t)std::max(1, sms / 2);
    // database tiles per work item: long enough to amortise the query-tile reload, short enough to balance
    uint64_t want = (uint64_t)p.tiles_total * p.n_qpairs / ((uint64_t)max_pairs * 6u);
    p.group = (uint32_t)std::min<uint64_t>(32u, std::max<uint64_t>(4u, want));
    p.n_groups = (p.tiles_total + p.group - 1) / p.group;
    p.thr_dot = ts.thr_dot; p.mask = d_mask; p.cand_ids = ts.cand_ids; p.cand_sc = ts.cand_sc; p.cand_cnt = s.cand_cnt; p.cap = cap; p.overflow = s.overflow;
    const uint32_t items = p.n_qpairs * p.n_groups;
    const int grid = 2 * (int)std::min<uint32_t>(max_pairs, items);
    static unsigned long long attr_seen = 0;   // cudaFuncSetAttribute is per device
    {
        int dev = 0;
        cudaGetDevice(&dev);
        const unsigned long long bit = 1ull << (dev & 63);
        if (!(attr_seen & bit)) {
            LEANN_CUDA_CHECK(cudaFuncSetAttribute(scan_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_MAX));
            LEANN_CUDA_CHECK(cudaFuncSetAttribute(scan_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_MAX));
            attr_seen |= bit;
        }
    }
    if (resident) scan_tc_kernel<true><<<grid, TC_THREADS, smem, stream>>>(mq, mx, p);
    else scan_tc_kernel<false><<<grid, TC_THREADS, smem, stream>>>(mq, mx, p);
    LEANN_CUDA_CHECK(cudaGetLastError());
    static const bool no_cut = getenv("LEANN_CUDA_SCAN_NO_CUT") != nullptr;   // A/B switch for benchmarks
    rerank_kernel<<<nq, 256, 0, stream>>>(f.vecs, s.qpad, f.d4, nq, f.metric, ts.cand_ids, s.cand_cnt, cap, s.cand,
                                          no_cut ? nullptr : ts.cand_sc, ts.cut_slack, k);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

// Re-scores the running top-k of the first chunk with rerank_kernel (see requeue_best_kernel); the caller runs
// select_kernel afterwards, exactly as after a tensor round.
void exact_scan_tc_canonical_first(const FlatView& f, const ScanScratch& s, const TcScratch& ts, uint32_t nq, uint32_t kpad, uint32_t cap,
                                   cudaStream_t stream) {
    requeue_best_kernel<<<nq, 128, 0, stream>>>(s.best, s.best_cnt, kpad, s.thr, ts.cand_ids, s.cand_cnt, cap, nq);
    rerank_kernel<<<nq, 256, 0, stream>>>(f.vecs, s.qpad, f.d4, nq, f.metric, ts.cand_ids, s.cand_cnt, cap, s.cand, nullptr, nullptr, 0u);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

void exact_scan_tc_queries(const float4* qpad, uint32_t nq, uint32_t d4, uint32_t dp8, int metric, const uint32_t* xmax_bits,
                           const TcScratch& ts, cudaStream_t stream) {
    to_bf16_kernel<<<(nq + 7) / 8, 256, 0, stream>>>(qpad, d4, (__nv_bfloat16*)ts.q_bf16, dp8, nq, ts.qnorm, ts.qres, nullptr,
                                                     metric == LEANN_METRIC_L2SQ ? AUG_QUERY : AUG_NONE, xmax_bits);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

}  // namespace leann
