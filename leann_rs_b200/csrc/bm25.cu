// bm25.cu — K3 / K3f: BM25 over a CSR inverted index resident in HBM, BM25 top-k, and the hybrid
// fusion + post-filter walk of IndexSearcher::search_with_options.
//   Bm25Scorer::score_query   leann-rs src/index/bm25.rs:77-106   (bm25_token_dense_kernel)
//   Bm25Scorer::search        src/index/bm25.rs:109-122           (bm25_query_kernel, top-k part)
//   hybrid_rerank             src/index/bm25.rs:135-170           (hybrid_fuse_kernel)
//   search_with_options glue  src/index/searcher.rs:146-207       (hybrid_fuse_kernel)
// The reference re-tokenises and re-indexes every passage on every hybrid query and probes a hash
// map per (token, document); here the index is built once and a query touches only its postings.
// Arithmetic is f32 with the reference's operation order and no FMA contraction, so scores are
// bit-identical: per document the contributions are added in query-token order (tokens are
// processed one after another; inside one token every posting is a distinct document).
// The per-posting contribution idf*(tf*(K1+1))/(tf+K1*norm) does not depend on the query, so it is
// computed once at build time (host, same f32 operation order) and the kernels only stream and add.
// Bound: HBM (postings stream). Algorithmic bytes per query = sum_t df_t * 8.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>
#include <math_constants.h>

#include "bm25_dev.h"

namespace leann {

namespace {

// Tile geometry of the query kernel; overridable at compile time for A/B builds (benchmarks/k3_variants.sh).
#ifndef LEANN_BM_THREADS
#define LEANN_BM_THREADS 256
#endif
#ifndef LEANN_BM_TILE
#define LEANN_BM_TILE 8192
#endif
#ifndef LEANN_BM_CAP
#define LEANN_BM_CAP 2048
#endif
#ifndef LEANN_BM_MINB
#define LEANN_BM_MINB 3
#endif
constexpr int BM_THREADS = LEANN_BM_THREADS;
constexpr int BM_TILE = LEANN_BM_TILE;     // documents per shared-memory accumulator tile (32 KB of f32)
constexpr int BM_CAP = LEANN_BM_CAP;       // candidate keys held in shared memory between prunes
constexpr int BM_BOUNDS = 2048;   // entries of the (token, tile boundary) -> posting offset table (16 KB)
constexpr int BM_SPARSE = 1024;   // tiles holding fewer postings are collected by re-walking them instead of a full scan
constexpr uint32_t BM_NOT_DENSE = 0xFFFFFFFFu;
constexpr size_t BM_SMEM = (size_t)BM_TILE * 4 + (size_t)BM_CAP * 8 + (size_t)BM_BOUNDS * 8 + (size_t)(BM_BOUNDS / 2) * 4;
static_assert(BM_TILE % (4 * BM_THREADS) == 0, "a tile is a whole number of float4 per thread");

__device__ __forceinline__ uint32_t order_f32(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(uint32_t u) {
    u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
    return __uint_as_float(u);
}

__global__ void bm25_token_dense_kernel(Bm25Dev b, uint32_t term, float* __restrict__ scores) {
    uint64_t p0 = b.term_off[term], p1 = b.term_off[term + 1];
    uint64_t p = p0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    uint32_t doc = b.post_doc[p];
    scores[doc] = __fadd_rn(scores[doc], b.post_score[p]);
}

__device__ void block_sort_n(unsigned long long* keys, const uint32_t n) {   // n: power of two, ascending
    for (uint32_t size = 2; size <= n; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < n / 2; t += blockDim.x) {
                uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                bool up = ((lo & size) == 0);
                unsigned long long a = keys[lo], c = keys[hi];
                if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
            }
        }
    __syncthreads();
}

// One CTA per query (persistent pool, dynamic scheduling). The dense score vector of the reference
// (bm25.rs:77-106) is never materialised in HBM: documents are walked in tiles of BM_TILE, each tile's scores
// are accumulated in shared memory token by token (postings are doc-ascending, so a tile is a contiguous slice
// of every posting list; the slice boundaries come from one parallel binary search per query), then the tile is
// scanned for positives, which feed the running top-K, the positive count and the min/max of hybrid_rerank.
// Every posting is read once, coalesced; no global read-modify-write.
//
// K3d, dense rows: a token whose term has a dense row adds the row's tile float4 by float4 (x + 0.0f == x for the
// non-negative partial sums, so documents outside the posting list keep their bits); a thread owns the same documents in
// every dense step and in the scan, so dense steps need no barrier between them. Three consequences are used:
//   * per document the first two contributions commute exactly (both start from +0.0f): when token 1 is dense and token 0
//     is not, they are processed in the order 1, 0 so that the tile starts with a dense row;
//   * a tile that starts with a dense row is initialised by a plain store of the row: the accumulator is never cleared
//     between the tiles of such a query (once at its end);
//   * when the last token is dense the scan runs on the sums while they are still in registers.
__global__ void __launch_bounds__(BM_THREADS, LEANN_BM_MINB)
bm25_query_kernel(Bm25Dev b, const uint64_t* __restrict__ qtok_off, const uint32_t* __restrict__ qtok_term,
                  uint32_t nq, uint32_t K, const uint64_t* __restrict__ cand_idx,
                  const uint32_t* __restrict__ cand_cnt, uint32_t fk, float* __restrict__ cand_bm,
                  uint64_t* __restrict__ top_idx, float* __restrict__ top_score, uint32_t* __restrict__ top_cnt,
                  float* __restrict__ bmax, float* __restrict__ bmin, uint32_t* __restrict__ qcounter) {
    extern __shared__ __align__(16) unsigned char bm_smem[];
    float* acc = reinterpret_cast<float*>(bm_smem);                                                  // [BM_TILE]
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(bm_smem + (size_t)BM_TILE * 4);   // [BM_CAP]
    unsigned long long* bounds = buf + BM_CAP;                                                         // [T][S + 1] absolute posting offsets
    uint32_t* tokd = reinterpret_cast<uint32_t*>(bounds + BM_BOUNDS);                                  // [T] dense row of the token, or BM_NOT_DENSE
    float4* acc4 = reinterpret_cast<float4*>(acc);
    constexpr int NV = BM_TILE / (4 * BM_THREADS);   // float4 per thread and tile
    __shared__ uint32_t s_cnt, s_q, s_pos, s_minbits;
    __shared__ unsigned long long s_thr;
    const int tid = threadIdx.x;
    for (int i = tid; i < BM_TILE; i += BM_THREADS) acc[i] = 0.0f;

    __shared__ uint32_t s_hist[256], s_sel_bin, s_sel_rem;
    // All threads; keeps the K best (smallest) keys, in no particular order, and tightens the threshold to the K-th best.
    // Radix select, one byte per round from the top: histogram of the keys that still match the prefix, warp 0 finds the bin
    // the K-th key falls in. 24 barriers instead of the 66 stages of a full sort of the buffer (this ran ~6 times per query and
    // was 10 % of the kernel's stall samples).
    auto prune = [&]() {
        __syncthreads();
        const uint32_t c = min(s_cnt, (uint32_t)BM_CAP);   // the scan may reserve past the end
        if (c <= K) return;                                // uniform; nothing to drop, the threshold stays
        constexpr int KPT = BM_CAP / BM_THREADS;
        unsigned long long mine[KPT];
#pragma unroll
        for (int u = 0; u < KPT; ++u) { const uint32_t i = tid + u * BM_THREADS; mine[u] = i < c ? buf[i] : ~0ull; }
        unsigned long long prefix = 0;
        uint32_t remaining = K;
        for (int byte = 7; byte >= 0; --byte) {
            const int sh = byte * 8;
            for (int i = tid; i < 256; i += BM_THREADS) s_hist[i] = 0;
            __syncthreads();
#pragma unroll
            for (int u = 0; u < KPT; ++u)
                if (byte == 7 || (mine[u] >> (sh + 8)) == (prefix >> (sh + 8))) atomicAdd(&s_hist[(uint32_t)(mine[u] >> sh) & 255u], 1u);
            __syncthreads();
            if (tid < 32) {
                uint32_t cnt8[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { cnt8[j] = s_hist[tid * 8 + j]; sum += cnt8[j]; }
                uint32_t incl = sum;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off); if (tid >= off) incl += v; }
                const uint32_t excl = incl - sum;
                if (excl < remaining && remaining <= incl) {
                    uint32_t r = remaining - excl;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (r != 0 && r <= cnt8[j]) { s_sel_bin = tid * 8 + j; s_sel_rem = r; r = 0; }
                        else if (r != 0) r -= cnt8[j];
                    }
                }
            }
            __syncthreads();
            prefix |= (unsigned long long)s_sel_bin << sh;
            remaining = s_sel_rem;
        }
        // prefix is the K-th smallest key (keys are distinct: the document is part of them)
        if (tid == 0) { s_cnt = 0; s_thr = prefix; }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < KPT; ++u) if (mine[u] <= prefix) buf[atomicAdd(&s_cnt, 1u)] = mine[u];
        __syncthreads();
    };
    uint32_t my_pos = 0, my_min = 0xFFFFFFFFu;
    auto collect = [&](uint32_t doc, float v, unsigned long long thr) {   // v != 0
        if (v > 0.0f) {   // bm25.rs:115
            my_pos++;
            uint32_t o = order_f32(v);
            my_min = o < my_min ? o : my_min;
            unsigned long long key = ((unsigned long long)(~o) << 32) | doc;
            if (key <= thr) buf[atomicAdd(&s_cnt, 1u)] = key;
        }
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_q = atomicAdd(qcounter, 1u); s_cnt = 0; s_thr = ~0ull; s_pos = 0; s_minbits = 0xFFFFFFFFu; }
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= nq) break;
        const uint64_t t0 = qtok_off[q];
        const uint32_t T = (uint32_t)(qtok_off[q + 1] - t0);     // <= BM_BOUNDS / 2 (checked by the host)
        const uint32_t nc = cand_idx ? cand_cnt[q] : 0u;
        for (uint32_t j = tid; j < nc; j += BM_THREADS) cand_bm[(size_t)q * fk + j] = 0.0f;   // bm25.rs:160 unwrap_or(0.0) / untouched tiles
        my_pos = 0; my_min = 0xFFFFFFFFu;
        bool zero_seen = false;   // block-uniform: some document of the corpus scores 0.0 (then bm25.rs:153's minimum is 0.0)
        bool zero = false;        // this thread saw one in the current tile
        bool crowded = false;     // this thread's reservation went past half of the candidate buffer: prune before the next tile
        for (uint32_t t = tid; t < T; t += BM_THREADS) tokd[t] = b.dense_of ? __ldg(b.dense_of + qtok_term[t0 + t]) : BM_NOT_DENSE;
        __syncthreads();
        // processing order of the tokens: 1, 0, 2, 3, .. when that puts a dense row first, else 0, 1, 2, ..
        const bool swap01 = T >= 2 && tokd[0] == BM_NOT_DENSE && tokd[1] != BM_NOT_DENSE;
        const bool first_dense = T && tokd[swap01 ? 1 : 0] != BM_NOT_DENSE;   // the accumulator is initialised by a store, never cleared
        const bool fuse_last = T && nc == 0 && tokd[(swap01 && T == 2) ? 0 : T - 1] != BM_NOT_DENSE;   // scan from registers
        const uint32_t n_tiles = (b.n_docs + BM_TILE - 1) / BM_TILE;
        const uint32_t S = T ? max(1u, (uint32_t)BM_BOUNDS / T - 1u) : n_tiles;   // tiles per boundary table

        // One float4 of finished sums: candidates that can still enter the top-K are appended, the statistics of bm25.rs:152-153
        // are kept while no zero score was seen. false: the candidate buffer is full (nothing was consumed; come back after a prune).
        auto consume = [&](const float4 v, const uint32_t doc, const unsigned long long thr, const float thr_f) -> bool {
            const float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            if (m >= thr_f) {   // v >= thr_f is necessary for key <= thr, and for v > 0 (bm25.rs:115)
                const float vv[4] = {v.x, v.y, v.z, v.w};
                unsigned long long key[4];
                uint32_t k = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    key[j] = ((unsigned long long)(~order_f32(vv[j])) << 32) | (doc + j);
                    if (vv[j] > 0.0f && key[j] <= thr) ++k; else key[j] = ~0ull;
                }
                if (k) {
                    uint32_t pos = atomicAdd(&s_cnt, k);
                    if (pos + k > BM_CAP) {   // fill what this reservation holds of the buffer
                        for (; pos < BM_CAP; ++pos) buf[pos] = ~0ull;
                        return false;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (key[j] != ~0ull) buf[pos++] = key[j];
                    crowded |= pos > BM_CAP / 2;
                }
            }
            if (!zero_seen) {
                my_pos += (v.x > 0.0f) + (v.y > 0.0f) + (v.z > 0.0f) + (v.w > 0.0f);
                const float mn4 = fminf(fminf(v.x > 0.0f ? v.x : CUDART_INF_F, v.y > 0.0f ? v.y : CUDART_INF_F),
                                        fminf(v.z > 0.0f ? v.z : CUDART_INF_F, v.w > 0.0f ? v.w : CUDART_INF_F));
                if (mn4 < CUDART_INF_F) { const uint32_t o = order_f32(mn4); my_min = o < my_min ? o : my_min; }
                zero |= (v.x == 0.0f && doc < b.n_docs) || (v.y == 0.0f && doc + 1 < b.n_docs) ||
                        (v.z == 0.0f && doc + 2 < b.n_docs) || (v.w == 0.0f && doc + 3 < b.n_docs);
            }
            return true;
        };

        for (uint32_t sb0 = 0; sb0 < n_tiles && T; sb0 += S) {
            const uint32_t sbt = min(S, n_tiles - sb0);
            // ---- slice boundaries: lower_bound(first doc of tile) in every token's posting list ----
            __syncthreads();
            for (uint32_t e = tid; e < T * (sbt + 1); e += BM_THREADS) {
                const uint32_t t = e / (sbt + 1), i = e % (sbt + 1);
                if (tokd[t] != BM_NOT_DENSE) continue;   // dense rows need no slice boundaries
                const uint32_t term = qtok_term[t0 + t];
                uint64_t lo = b.term_off[term], hi = b.term_off[term + 1];
                const uint64_t target = (uint64_t)(sb0 + i) * BM_TILE;
                if (target >= b.n_docs) lo = hi;
                while (lo < hi) {
                    uint64_t mid = (lo + hi) >> 1;
                    if (b.post_doc[mid] < target) lo = mid + 1; else hi = mid;
                }
                bounds[e] = lo;
            }
            __syncthreads();
            for (uint32_t tile = 0; tile < sbt; ++tile) {
                const uint32_t base = (sb0 + tile) * BM_TILE;
                // ---- accumulate, tokens in query order (duplicates counted again, bm25.rs:81) ----
                uint64_t in_tile = 0;
                int prev = 0;         // 0: nothing written since the last barrier, 1: a posting slice, 2: a dense row
                int resume = 0;       // first float4 of this thread the scan has not consumed yet
                unsigned long long thr = s_thr;   // stable: written by prune only, between barriers
                float thr_f = thr == ~0ull ? __uint_as_float(1u) : unorder_f32(~(uint32_t)(thr >> 32));
                for (uint32_t p = 0; p < T; ++p) {
                    const uint32_t t = (swap01 && p < 2) ? 1u - p : p;
                    const uint32_t dr = tokd[t];
                    if (dr != BM_NOT_DENSE) {
                        in_tile += BM_TILE;   // an upper bound of the term's postings here: always the full scan below
                        const float4* __restrict__ row = reinterpret_cast<const float4*>(b.dense_rows + (size_t)dr * b.n_pad + base);
                        float4 r[NV];
#pragma unroll
                        for (int u = 0; u < NV; ++u) r[u] = __ldg(row + u * BM_THREADS + tid);
                        if (prev == 1) __syncthreads();   // after the loads are issued: the wait overlaps their latency
                        prev = 2;
                        const bool init = p == 0;   // first_dense: whatever the accumulator holds is stale
                        if (!(fuse_last && p == T - 1)) {
#pragma unroll
                            for (int u = 0; u < NV; ++u) {
                                float4 a = r[u];
                                if (!init) {
                                    const float4 o = acc4[u * BM_THREADS + tid];
                                    a.x = __fadd_rn(o.x, a.x); a.y = __fadd_rn(o.y, a.y); a.z = __fadd_rn(o.z, a.z); a.w = __fadd_rn(o.w, a.w);
                                }
                                acc4[u * BM_THREADS + tid] = a;
                            }
                        } else {
                            // last token: the sums are consumed from registers; what does not fit the candidate buffer any more
                            // is parked in the accumulator for the scan below
                            resume = NV;
#pragma unroll
                            for (int u = 0; u < NV; ++u) {
                                float4 a = r[u];
                                const int i4 = u * BM_THREADS + tid;
                                if (!init) {
                                    const float4 o = acc4[i4];
                                    a.x = __fadd_rn(o.x, a.x); a.y = __fadd_rn(o.y, a.y); a.z = __fadd_rn(o.z, a.z); a.w = __fadd_rn(o.w, a.w);
                                }
                                if (resume == NV && !consume(a, base + (uint32_t)i4 * 4u, thr, thr_f)) resume = u;
                                if (resume != NV) acc4[i4] = a;
                                else if (!first_dense) acc4[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
                        continue;
                    }
                    const uint64_t lo = bounds[t * (sbt + 1) + tile], hi = bounds[t * (sbt + 1) + tile + 1];
                    if (lo >= hi) continue;
                    in_tile += hi - lo;
                    // documents inside one token are distinct: plain shared-memory read-modify-write, 8 postings in flight per thread.
                    // The loads of the first chunk are issued BEFORE the barrier that ends the previous token's writes: they do not
                    // depend on it, and the wait for the slowest warp overlaps their latency.
                    const uint32_t* __restrict__ pd = b.post_doc + lo;
                    const float* __restrict__ ps = b.post_score + lo;
                    const uint32_t len = (uint32_t)(hi - lo);
                    for (uint32_t i0 = 0; i0 < len; i0 += 8 * BM_THREADS) {
                        uint32_t d[8]; float sc[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const uint32_t i = i0 + tid + (uint32_t)u * BM_THREADS;
                            const bool ok = i < len;
                            d[u] = ok ? __ldg(pd + i) : 0xFFFFFFFFu;
                            sc[u] = ok ? __ldg(ps + i) : 0.0f;
                        }
                        if (i0 == 0 && prev) __syncthreads();
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (d[u] != 0xFFFFFFFFu) acc[d[u] - base] = __fadd_rn(acc[d[u] - base], sc[u]);
                    }
                    prev = 1;
                }
                const bool fused = fuse_last;   // (a dense last token always runs: the scan from registers happened)
                if (prev == 1 || (prev == 2 && !fused)) __syncthreads();
                const uint32_t docs_here = min((uint32_t)BM_TILE, b.n_docs - base);
                if (in_tile < (uint64_t)docs_here) zero_seen = true;   // fewer postings than documents: some document scores 0.0
                if (in_tile == 0) continue;
                // ---- BM25 score of the vector candidates that live in this tile (bm25.rs:160) ----
                if (nc) {
                    for (uint32_t j = tid; j < nc; j += BM_THREADS) {
                        const uint64_t idx = cand_idx[(size_t)q * fk + j];
                        if (idx >= base && idx < (uint64_t)base + BM_TILE && idx < b.n_docs) cand_bm[(size_t)q * fk + j] = acc[idx - base];
                    }
                    __syncthreads();
                }
                if (in_tile < (uint64_t)BM_SPARSE) {
                    // ---- few postings: collect by re-walking them (the first visit of a document takes and clears its score) ----
                    for (uint32_t t = 0; t < T; ++t) {
                        const uint64_t lo = bounds[t * (sbt + 1) + tile], hi = bounds[t * (sbt + 1) + tile + 1];
                        if (lo >= hi) continue;
                        const uint32_t c = s_cnt;
                        __syncthreads();   // everyone has read the count before anyone appends: the decision is uniform
                        if (c > BM_CAP - BM_SPARSE) prune();
                        const unsigned long long thr2 = s_thr;
                        for (uint64_t p = lo + tid; p < hi; p += BM_THREADS) {
                            const uint32_t doc = __ldg(b.post_doc + p);
                            const float v = acc[doc - base];
                            if (v != 0.0f) { acc[doc - base] = 0.0f; collect(doc, v, thr2); }
                        }
                        __syncthreads();
                    }
                } else {
                    // ---- scan the tile: collect positives, clear the accumulator (unless the next tile overwrites it anyway) ----
                    // Optimistic: no barrier inside a pass. A thread stops at the first float4 whose keys do not fit the candidate
                    // buffer; the pass ends with a vote, a prune, and everyone resumes where they stopped. Once the threshold is
                    // established a tile appends a handful of keys and one pass is all there is.
                    for (int pass = 0;; ++pass) {
                        if (!(fused && pass == 0)) {
                            int it = resume;
                            resume = NV;
                            for (; it < NV; ++it) {
                                const int i4 = it * BM_THREADS + tid;
                                const float4 v = acc4[i4];
                                if (!consume(v, base + (uint32_t)i4 * 4u, thr, thr_f)) { resume = it; break; }
                                if (!first_dense && fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) != 0.0f) acc4[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
                        if (pass == 0 && !zero_seen && __syncthreads_or(zero)) zero_seen = true;
                        zero = false;
                        // one vote in the common case; the thread that made the last reservation knows how full the buffer is
                        if (!__syncthreads_or((resume != NV) | crowded)) break;
                        prune();
                        crowded = false;
                        thr = s_thr;
                        thr_f = thr == ~0ull ? __uint_as_float(1u) : unorder_f32(~(uint32_t)(thr >> 32));
                        if (!__syncthreads_or(resume != NV)) break;
                    }
                }
            }
        }
        if (first_dense) {   // leave the accumulator zero for the next query
            __syncthreads();
#pragma unroll
            for (int u = 0; u < NV; ++u) acc4[u * BM_THREADS + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        atomicAdd(&s_pos, my_pos);
        atomicMin(&s_minbits, my_min);
        prune();
        __syncthreads();
        const uint32_t cnt = min(s_cnt, K);   // (<= K after the prune)
        {   // the answer in order: sort the survivors, padded to a power of two
            uint32_t n2 = 2;
            while (n2 < cnt) n2 <<= 1;
            for (uint32_t i = cnt + tid; i < n2; i += BM_THREADS) buf[i] = ~0ull;
            block_sort_n(buf, n2);
        }
        for (uint32_t j = tid; j < K; j += BM_THREADS) {
            size_t o = (size_t)q * K + j;
            if (j < cnt) {
                unsigned long long key = buf[j];
                top_idx[o] = (uint64_t)(key & 0xFFFFFFFFull);
                top_score[o] = unorder_f32(~(uint32_t)(key >> 32));
            } else {
                top_idx[o] = ~0ull;
                top_score[o] = 0.0f;
            }
        }
        if (tid == 0) {
            top_cnt[q] = cnt;
            // bm25.rs:152-153: max / min over the ENTIRE dense score vector (zeros included)
            float mx = cnt ? unorder_f32(~(uint32_t)(buf[0] >> 32)) : 0.0f;
            float mn = (zero_seen || s_pos < b.n_docs) ? 0.0f : unorder_f32(s_minbits);
            if (b.n_docs == 0) { mx = -CUDART_INF_F; mn = CUDART_INF_F; }
            bmax[q] = mx;
            bmin[q] = mn;
        }
    }
}

// BM25 score of given (query, document) pairs straight from the postings (bm25_scores.get(idx).unwrap_or(0.0), bm25.rs:160): per
// document the contributions of the query's tokens are added in token order, starting from 0.0f, exactly like the dense
// accumulation, so the bits are the same. Lets the vector search and the BM25 top-k kernel run concurrently: neither needs
// the other's output, only this small kernel and the fusion do.
__global__ void __launch_bounds__(64)
bm25_cand_kernel(Bm25Dev b, const uint64_t* __restrict__ qtok_off, const uint32_t* __restrict__ qtok_term, uint32_t nq,
                 const uint64_t* __restrict__ cand_idx, const uint32_t* __restrict__ cand_cnt, uint32_t fk, float* __restrict__ cand_bm) {
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    const uint32_t nc = cand_cnt[q];
    const uint64_t t0 = qtok_off[q], t1 = qtok_off[q + 1];
    for (uint32_t j = threadIdx.x; j < fk; j += blockDim.x) {
        float acc = 0.0f;
        const uint64_t doc = j < nc ? cand_idx[(size_t)q * fk + j] : ~0ull;
        if (doc < b.n_docs) {
            for (uint64_t t = t0; t < t1; ++t) {
                const uint32_t term = qtok_term[t];
                uint64_t lo = b.term_off[term], hi = b.term_off[term + 1];
                const uint64_t end = hi;
                while (lo < hi) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (b.post_doc[mid] < (uint32_t)doc) lo = mid + 1; else hi = mid;
                }
                if (lo < end && b.post_doc[lo] == (uint32_t)doc) acc = __fadd_rn(acc, b.post_score[lo]);
            }
        }
        cand_bm[(size_t)q * fk + j] = acc;
    }
}

// searcher.rs:146-207 + bm25.rs:135-170 for one query per block.
__global__ void __launch_bounds__(128)
hybrid_fuse_kernel(const uint64_t* __restrict__ vkeys, const float* __restrict__ vdists, const uint32_t* __restrict__ vcnt,
                   uint32_t fk, const float* __restrict__ cand_bm, const uint64_t* __restrict__ bm_idx,
                   const float* __restrict__ bm_score, const uint32_t* __restrict__ bm_cnt, uint32_t bm_k,
                   const float* __restrict__ bmax, const float* __restrict__ bmin, int hybrid, float alpha,
                   const uint64_t* __restrict__ mask, uint64_t mask_bits, uint32_t top_k, uint64_t* __restrict__ out_idx,
                   float* __restrict__ out_score, uint32_t* __restrict__ out_cnt, uint32_t nq, uint32_t cap) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint64_t* idx = reinterpret_cast<uint64_t*>(sm);                       // [cap]
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(sm + (size_t)cap * 8);  // [n2 <= 2*cap pow2]
    float* vs = reinterpret_cast<float*>(sm + (size_t)cap * 8 + (size_t)2 * cap * 8);  // [cap]
    float* bs = vs + cap;                                                  // [cap]
    uint8_t* fresh = reinterpret_cast<uint8_t*>(bs + cap);                 // [bm_k]
    __shared__ uint32_t s_len, s_vmin, s_vmax;
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    const int tid = threadIdx.x;
    const uint32_t nv = vcnt[q];
    for (uint32_t i = tid; i < nv; i += blockDim.x) {
        idx[i] = vkeys[(size_t)q * fk + i];
        vs[i] = vdists[(size_t)q * fk + i];
        bs[i] = (hybrid && cand_bm) ? cand_bm[(size_t)q * fk + i] : 0.0f;
    }
    if (tid == 0) { s_len = nv; s_vmin = 0xFFFFFFFFu; s_vmax = 0u; }
    __syncthreads();
    uint32_t len = nv;
    if (hybrid) {
        // BM25 top results not already among the vector results enter with vector score 0.0
        const uint32_t nb = bm_cnt[q];
        for (uint32_t j = tid; j < nb; j += blockDim.x) {
            uint64_t d = bm_idx[(size_t)q * bm_k + j];
            bool present = false;
            for (uint32_t i = 0; i < nv; ++i) if (idx[i] == d) { present = true; break; }
            fresh[j] = present ? 0 : 1;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t l = nv;
            for (uint32_t j = 0; j < nb && l < cap; ++j)
                if (fresh[j]) { idx[l] = bm_idx[(size_t)q * bm_k + j]; vs[l] = 0.0f; bs[l] = bm_score[(size_t)q * bm_k + j]; ++l; }
            s_len = l;
        }
        __syncthreads();
        len = s_len;
        for (uint32_t i = tid; i < len; i += blockDim.x) {
            uint32_t o = order_f32(vs[i]);
            atomicMin(&s_vmin, o);
            atomicMax(&s_vmax, o);
        }
        __syncthreads();
        // fold(NEG_INFINITY, max) / fold(INFINITY, min): an empty list leaves the infinities in place
        const float vmax = len ? unorder_f32(s_vmax) : -CUDART_INF_F;
        const float vmin = len ? unorder_f32(s_vmin) : CUDART_INF_F;
        const float vr = fmaxf(__fsub_rn(vmax, vmin), 1e-6f);
        const float bmn = bmin[q];
        const float br = fmaxf(__fsub_rn(bmax[q], bmn), 1e-6f);
        const float one_minus = __fsub_rn(1.0f, alpha);
        for (uint32_t i = tid; i < len; i += blockDim.x) {
            float nvv = __fdiv_rn(__fsub_rn(vs[i], vmin), vr);
            float nbv = __fdiv_rn(__fsub_rn(bs[i], bmn), br);
            vs[i] = __fadd_rn(__fmul_rn(alpha, nvv), __fmul_rn(one_minus, nbv));
        }
        __syncthreads();
        // stable descending sort: key = (~ordered(score), position)
        uint32_t n2 = 1;
        while (n2 < len) n2 <<= 1;
        for (uint32_t i = tid; i < n2; i += blockDim.x)
            keys[i] = i < len ? (((unsigned long long)(~order_f32(vs[i])) << 32) | i) : ~0ull;
        for (uint32_t size = 2; size <= n2; size <<= 1)
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (uint32_t t = tid; t < n2 / 2; t += blockDim.x) {
                    uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    bool up = ((lo & size) == 0);
                    unsigned long long a = keys[lo], c = keys[hi];
                    if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
                }
            }
        __syncthreads();
    } else {
        for (uint32_t i = tid; i < len; i += blockDim.x) keys[i] = i;
        __syncthreads();
    }
    // post-filter walk (searcher.rs:174-207): take in rank order until top_k pass
    if (tid == 0) {
        uint32_t o = 0;
        for (uint32_t r = 0; r < len && o < top_k; ++r) {
            uint32_t pos = (uint32_t)(keys[r] & 0xFFFFFFFFull);
            uint64_t d = idx[pos];
            if (mask && !(d < mask_bits && ((mask[d >> 6] >> (d & 63ull)) & 1ull))) continue;
            out_idx[(size_t)q * top_k + o] = d;
            out_score[(size_t)q * top_k + o] = vs[pos];
            ++o;
        }
        out_cnt[q] = o;
        for (; o < top_k; ++o) { out_idx[(size_t)q * top_k + o] = ~0ull; out_score[(size_t)q * top_k + o] = 0.0f; }
    }
}

// min / max over a dense f32 vector (for the stand-alone hybrid_rerank entry point)
__global__ void minmax_kernel(const float* __restrict__ v, uint32_t n, uint32_t* __restrict__ out /*[min,max] ordered*/) {
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t o = order_f32(v[i]);
        mn = o < mn ? o : mn;
        mx = o > mx ? o : mx;
    }
    for (int off = 16; off; off >>= 1) {
        uint32_t a = __shfl_xor_sync(0xFFFFFFFFu, mn, off), c = __shfl_xor_sync(0xFFFFFFFFu, mx, off);
        mn = a < mn ? a : mn;
        mx = c > mx ? c : mx;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&out[0], mn); atomicMax(&out[1], mx); }
}
__global__ void minmax_finish_kernel(const uint32_t* in, uint32_t n, float* bmax, float* bmin) {
    if (n == 0) { *bmax = -CUDART_INF_F; *bmin = CUDART_INF_F; return; }
    *bmin = unorder_f32(in[0]);
    *bmax = unorder_f32(in[1]);
}
__global__ void gather_kernel(const float* __restrict__ dense, uint32_t n, const uint64_t* __restrict__ idx, uint32_t m, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = idx[i] < n ? dense[idx[i]] : 0.0f;
}

}  // namespace

// cudaFuncSetAttribute is per device: remember which devices of this process already have it.
static bool first_use_on_device(unsigned long long& seen) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (seen & bit) return false;
    seen |= bit;
    return true;
}

void launch_bm25_dense(const Bm25Dev& b, const uint32_t* terms, const uint64_t* dfs, size_t n_tokens, float* d_scores,
                       cudaStream_t s) {
    for (size_t i = 0; i < n_tokens; ++i) {
        if (dfs[i] == 0) continue;
        unsigned blocks = (unsigned)((dfs[i] + 255) / 256);
        bm25_token_dense_kernel<<<blocks, 256, 0, s>>>(b, terms[i], d_scores);
    }
    LEANN_CUDA_CHECK(cudaGetLastError());
}

int bm25_query_ctas_per_sm() { return LEANN_BM_MINB; }

namespace {
__global__ void dense_of_fill_kernel(uint32_t* __restrict__ dense_of, uint32_t n_terms) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_terms) dense_of[i] = BM_NOT_DENSE;
}
// one launch per dense term: row[doc] = score of the term's posting for doc (the rest of the row stays 0.0f)
__global__ void dense_row_scatter_kernel(Bm25Dev b, uint32_t term, float* __restrict__ row) {
    const uint64_t p0 = b.term_off[term], p1 = b.term_off[term + 1];
    for (uint64_t p = p0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += (uint64_t)gridDim.x * blockDim.x)
        row[b.post_doc[p]] = b.post_score[p];
}
}  // namespace

void bm25_build_dense_rows(leann_cuda_bm25* b) {
    b->n_dense = 0;
    b->is_dense.clear();
    const size_t n = b->host.num_docs, n_terms = b->host.term_off.empty() ? 0 : b->host.term_off.size() - 1;
    double frac = 0.25;
    if (const char* e = getenv("LEANN_CUDA_BM25_DENSE_FRAC")) frac = atof(e);
    size_t max_rows = 64;
    if (const char* e = getenv("LEANN_CUDA_BM25_DENSE_MAX")) max_rows = (size_t)std::max(0, atoi(e));
    if (!(frac > 0.0) || n == 0 || n_terms == 0 || max_rows == 0) return;
    const size_t n_pad = (n + BM_TILE - 1) / BM_TILE * BM_TILE;
    size_t free_b = 0, total_b = 0;
    LEANN_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
    max_rows = std::min(max_rows, free_b / 8 / (n_pad * 4));
    std::vector<std::pair<uint64_t, uint32_t>> cand;   // (df, term)
    for (size_t t = 0; t < n_terms; ++t) {
        const uint64_t df = b->host.term_off[t + 1] - b->host.term_off[t];
        if ((double)df >= frac * (double)n) cand.emplace_back(df, (uint32_t)t);
    }
    std::sort(cand.begin(), cand.end(), [](const std::pair<uint64_t, uint32_t>& x, const std::pair<uint64_t, uint32_t>& y) {
        return x.first != y.first ? x.first > y.first : x.second < y.second;
    });
    if (cand.size() > max_rows) cand.resize(max_rows);
    if (cand.empty()) return;
    LEANN_CUDA_CHECK(cudaMalloc(&b->d_dense_of, n_terms * 4));
    LEANN_CUDA_CHECK(cudaMalloc(&b->d_dense_rows, cand.size() * n_pad * 4));
    LEANN_CUDA_CHECK(cudaMemset(b->d_dense_rows, 0, cand.size() * n_pad * 4));
    dense_of_fill_kernel<<<(unsigned)((n_terms + 255) / 256), 256>>>(b->d_dense_of, (uint32_t)n_terms);
    b->n_pad = (uint32_t)n_pad;
    const Bm25Dev v = b->view();
    for (size_t r = 0; r < cand.size(); ++r) {
        const uint32_t row = (uint32_t)r;
        LEANN_CUDA_CHECK(cudaMemcpy(b->d_dense_of + cand[r].second, &row, 4, cudaMemcpyHostToDevice));
        dense_row_scatter_kernel<<<592, 256>>>(v, cand[r].second, b->d_dense_rows + r * n_pad);
    }
    LEANN_CUDA_CHECK(cudaDeviceSynchronize());
    LEANN_CUDA_CHECK(cudaGetLastError());
    b->n_dense = (uint32_t)cand.size();
    b->is_dense.assign(n_terms, 0);
    for (auto& c : cand) b->is_dense[c.second] = 1;
}

void launch_bm25_query(const Bm25Dev& b, const uint64_t* qtok_off, const uint32_t* qtok_term, uint32_t nq, uint32_t K,
                       int n_ctas, const uint64_t* cand_idx, const uint32_t* cand_cnt, uint32_t fk,
                       float* cand_bm, uint64_t* top_idx, float* top_score, uint32_t* top_cnt, float* bmax, float* bmin,
                       uint32_t* qcounter, cudaStream_t s) {
    if (K == 0 || K > 1024) throw Error(LEANN_ERR_INVALID_ARG, "bm25: top_k must be in 1..1024");
    static unsigned long long seen = 0;
    if (first_use_on_device(seen))
        LEANN_CUDA_CHECK(cudaFuncSetAttribute(bm25_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BM_SMEM));
    LEANN_CUDA_CHECK(cudaMemsetAsync(qcounter, 0, 4, s));
    bm25_query_kernel<<<n_ctas, BM_THREADS, BM_SMEM, s>>>(b, qtok_off, qtok_term, nq, K, cand_idx, cand_cnt, fk, cand_bm,
                                                          top_idx, top_score, top_cnt, bmax, bmin, qcounter);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

size_t bm25_max_query_tokens() { return BM_BOUNDS / 2; }

void launch_bm25_candidates(const Bm25Dev& b, const uint64_t* qtok_off, const uint32_t* qtok_term, uint32_t nq, const uint64_t* cand_idx,
                            const uint32_t* cand_cnt, uint32_t fk, float* cand_bm, cudaStream_t s) {
    if (nq == 0 || fk == 0) return;
    bm25_cand_kernel<<<nq, 64, 0, s>>>(b, qtok_off, qtok_term, nq, cand_idx, cand_cnt, fk, cand_bm);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

void launch_hybrid_fuse(const uint64_t* vkeys, const float* vdists, const uint32_t* vcnt, uint32_t fk, const float* cand_bm,
                        const uint64_t* bm_idx, const float* bm_score, const uint32_t* bm_cnt, uint32_t bm_k,
                        const float* bmax, const float* bmin, int hybrid, float alpha, const uint64_t* mask, uint64_t mask_bits,
                        uint32_t top_k, uint64_t* out_idx, float* out_score, uint32_t* out_cnt, uint32_t nq, cudaStream_t s) {
    if (nq == 0) return;
    uint32_t cap = fk + (hybrid ? bm_k : 0);
    if (cap == 0) cap = 1;
    if (cap > 4096) throw Error(LEANN_ERR_INVALID_ARG, "hybrid: fetch_k too large (top_k <= 400)");
    uint32_t n2 = 1;
    while (n2 < cap) n2 <<= 1;
    size_t smem = (size_t)cap * 8 + (size_t)2 * cap * 8 + (size_t)cap * 8 + bm_k + 16;
    if (smem > 48 * 1024) LEANN_CUDA_CHECK(cudaFuncSetAttribute(hybrid_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hybrid_fuse_kernel<<<nq, 128, smem, s>>>(vkeys, vdists, vcnt, fk, cand_bm, bm_idx, bm_score, bm_cnt, bm_k, bmax, bmin, hybrid,
                                             alpha, mask, mask_bits, top_k, out_idx, out_score, out_cnt, nq, cap);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

namespace {
__global__ void localize_kernel(const uint64_t* __restrict__ g, uint64_t off, uint64_t n_local, size_t count, uint64_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t v = g[i];
    out[i] = (v >= off && v - off < n_local) ? v - off : ~0ull;
}
// Block of rank r (sharded hybrid exchange): [nq][fk] u64 top ids | [nq][fk] f32 top scores | [nq][fk] f32 candidate scores |
// [nq] f32 max | [nq] f32 min. The candidate scores are summed over ranks in rank order (only the owner's is non-zero: exact).
__global__ void shard_reduce_kernel(const unsigned char* __restrict__ gathered, size_t block_bytes, uint32_t g, uint32_t nq, uint32_t fk,
                                    float* __restrict__ cand_bm, float* __restrict__ bmax, float* __restrict__ bmin) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)nq * fk;
    const size_t off_cb = n * 12, off_mx = n * 16, off_mn = n * 16 + (size_t)nq * 4;
    if (i < n) {
        float acc = 0.0f;
        for (uint32_t r = 0; r < g; ++r) acc = __fadd_rn(acc, reinterpret_cast<const float*>(gathered + r * block_bytes + off_cb)[i]);
        cand_bm[i] = acc;
    }
    if (i < nq) {
        float mx = -CUDART_INF_F, mn = CUDART_INF_F;
        for (uint32_t r = 0; r < g; ++r) {
            mx = fmaxf(mx, reinterpret_cast<const float*>(gathered + r * block_bytes + off_mx)[i]);
            mn = fminf(mn, reinterpret_cast<const float*>(gathered + r * block_bytes + off_mn)[i]);
        }
        bmax[i] = mx; bmin[i] = mn;
    }
}
}  // namespace

void launch_localize_candidates(const uint64_t* global_idx, uint64_t doc_offset, uint64_t n_local, size_t count, uint64_t* local_idx, cudaStream_t s) {
    if (!count) return;
    localize_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(global_idx, doc_offset, n_local, count, local_idx);
    LEANN_CUDA_CHECK(cudaGetLastError());
}
void launch_shard_reduce(const unsigned char* gathered, size_t block_bytes, uint32_t g, uint32_t nq, uint32_t fk, float* cand_bm, float* bmax,
                         float* bmin, cudaStream_t s) {
    const size_t n = std::max<size_t>((size_t)nq * fk, nq);
    if (!n) return;
    shard_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(gathered, block_bytes, g, nq, fk, cand_bm, bmax, bmin);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

void launch_dense_minmax_gather(const float* dense, uint32_t n, const uint64_t* idx, uint32_t m, float* cand_bm, float* bmax,
                                float* bmin, uint32_t* scratch2, cudaStream_t s) {
    uint32_t init[2] = {0xFFFFFFFFu, 0u};
    LEANN_CUDA_CHECK(cudaMemcpyAsync(scratch2, init, 8, cudaMemcpyHostToDevice, s));
    if (n) minmax_kernel<<<296, 256, 0, s>>>(dense, n, scratch2);
    minmax_finish_kernel<<<1, 1, 0, s>>>(scratch2, n, bmax, bmin);
    if (m) gather_kernel<<<(m + 255) / 256, 256, 0, s>>>(dense, n, idx, m, cand_bm);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

}  // namespace leann
