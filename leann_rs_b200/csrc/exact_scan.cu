// exact_scan.cu — K2 / K2r: exact brute-force scan with a fused top-k, replacing
// RecomputeSearcher::search's scoring + sort + take(k) (leann-rs src/index/recompute.rs:96-110,
// dot_product :137-139), plus K4 (per-query merge of per-shard top-k lists).
//
// Round structure ("progressive threshold"): the database is visited in growing row chunks. A tile
// kernel computes a 128-query x 128-row block of f32 scores and appends only the scores that can
// still enter the top-k (packed 64-bit key <= the query's current k-th best key) to a per-query
// candidate list; a select kernel then merges candidates into the running top-k and tightens the
// threshold. The full nq x N score matrix is never materialised. Chunks after the first go through the tensor-core
// pass of exact_scan_tc.cu when it applies; large batches are split into two query halves on two streams so that the
// re-rank / select of one half overlap the other half's tensor pass (launch_exact_scan).
// Ordering is total: key = (ordered(score) << 32) | row, so equal scores rank by ascending row —
// exactly what the reference's stable sort over the enumerate() order yields (recompute.rs:106).
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "scan_common.h"

namespace leann {

namespace {

constexpr int TM = 128, TN = 128, TK = 16, LDS_STRIDE = 132;

__device__ __forceinline__ uint32_t order_f32(float f) {  // monotone float -> uint
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unorder_f32(uint32_t u) {
    u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
    return __uint_as_float(u);
}

// One 128x128 tile of scores, K-loop over the padded dimension, threshold epilogue.
// DIRECT (first round: no threshold yet, every row is a candidate): the slot of a row is its offset in the chunk,
// no atomics; masked-out rows leave the ~0 sentinel and the host presets cand_cnt = r1 - r0.
template <int METRIC, bool DIRECT>
__global__ void __launch_bounds__(256)
scan_tile_kernel(const float4* __restrict__ X, const float4* __restrict__ Q, uint32_t d4, uint32_t nq,
                 uint32_t r0, uint32_t r1, const uint64_t* __restrict__ mask,
                 const unsigned long long* __restrict__ thr, unsigned long long* __restrict__ cand,
                 uint32_t* __restrict__ cand_cnt, uint32_t cap, uint32_t* __restrict__ overflow) {
    __shared__ __align__(16) float As[2][TK][LDS_STRIDE];
    __shared__ __align__(16) float Bs[2][TK][LDS_STRIDE];
    __shared__ unsigned long long s_thr[TM];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const uint32_t m0 = blockIdx.y * TM, n0 = r0 + blockIdx.x * TN;
    if (tid < TM) s_thr[tid] = (m0 + tid < nq) ? thr[m0 + tid] : 0ull;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const uint32_t kt = (d4 + 3) / 4;  // TK=16 floats = 4 float4 per k-tile
    float4 ra[2], rb[2];
    auto gload = [&](uint32_t t) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            int idx = tid + u * 256;
            int row = idx >> 2, c4 = idx & 3;
            uint32_t kc = t * 4 + c4;
            uint32_t qa = m0 + row, xb = n0 + row;
            ra[u] = (qa < nq && kc < d4) ? Q[(size_t)qa * d4 + kc] : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[u] = (xb < r1 && kc < d4) ? __ldg(&X[(size_t)xb * d4 + kc]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            int idx = tid + u * 256;
            int row = idx >> 2, c4 = idx & 3;
            As[buf][c4 * 4 + 0][row] = ra[u].x; As[buf][c4 * 4 + 1][row] = ra[u].y;
            As[buf][c4 * 4 + 2][row] = ra[u].z; As[buf][c4 * 4 + 3][row] = ra[u].w;
            Bs[buf][c4 * 4 + 0][row] = rb[u].x; Bs[buf][c4 * 4 + 1][row] = rb[u].y;
            Bs[buf][c4 * 4 + 2][row] = rb[u].z; Bs[buf][c4 * 4 + 3][row] = rb[u].w;
        }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    for (uint32_t t = 0; t < kt; ++t) {
        int buf = t & 1;
        if (t + 1 < kt) gload(t + 1);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (METRIC == LEANN_METRIC_L2SQ) {
                        float df = a[i] - b[j];
                        acc[i][j] = fmaf(df, df, acc[i][j]);
                    } else {
                        acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                    }
                }
        }
        if (t + 1 < kt) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }
    // epilogue: score -> rank key -> threshold test -> append
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int ml = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        uint32_t qi = m0 + ml;
        if (qi >= nq) continue;
        unsigned long long th = s_thr[ml];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int nl = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            uint32_t row = n0 + nl;
            if (row >= r1) continue;
            float s = acc[i][j];
            uint32_t ok;
            if (METRIC == LEANN_METRIC_DOT_DESC) ok = ~order_f32(s);
            else if (METRIC == LEANN_METRIC_L2SQ) ok = order_f32(s);
            else {
                float dd = 1.0f - s;
                if (METRIC == LEANN_METRIC_IP_CLAMP) dd = dd < 0.f ? 0.f : dd;
                ok = order_f32(dd);
            }
            unsigned long long key = ((unsigned long long)ok << 32) | row;
            if (DIRECT) {
                const bool keep = !mask || ((mask[row >> 6] >> (row & 63u)) & 1ull);
                cand[(size_t)qi * cap + (row - r0)] = keep ? key : ~0ull;
                continue;
            }
            if (key > th) continue;
            if (mask && !((mask[row >> 6] >> (row & 63u)) & 1ull)) continue;
            uint32_t pos = atomicAdd(&cand_cnt[qi], 1u);
            if (pos < cap) cand[(size_t)qi * cap + pos] = key;
            else *overflow = 1u;
        }
    }
}

// In-place bitonic sort of n (power of two) 64-bit keys in shared memory, ascending.
__device__ void block_bitonic_sort(unsigned long long* keys, uint32_t n) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < n / 2; t += blockDim.x) {
                uint32_t lo = 2 * t - (t & (stride - 1));
                uint32_t hi = lo + stride;
                bool up = ((lo & size) == 0);
                unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// The (kth+1)-th smallest of n 64-bit keys: MSB-first radix select, 8 bits per pass. `key(i)` reads key i (shared memory
// staging buffer, or global memory in the low-shared-memory variant). All threads of the block call it; hist = 256
// counters + 3 words in shared memory.
// FIRST_PASS .. LAST_PASS of the 8 byte-wise passes (0 = most significant byte). `prefix` / `remaining` carry the state
// between a high-half call (passes 0-3) and an optional low-half call (passes 4-7). After the last executed pass,
// hist[258] = number of keys equal to the prefix on the bytes examined so far.
template <int FIRST_PASS, int LAST_PASS, typename KeyFn>
__device__ void block_radix_select(KeyFn key, uint32_t n, unsigned long long& prefix, uint32_t& remaining, uint32_t* hist) {
    for (int pass = FIRST_PASS; pass <= LAST_PASS; ++pass) {
        const int shift = 56 - 8 * pass;
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long v = key(i);
            if (pass == 0 || (v >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(uint32_t)(v >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t lane = threadIdx.x;
            uint32_t c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; sum += c[j]; }
            uint32_t incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
                if (lane >= (uint32_t)off) incl += t;
            }
            const uint32_t before = incl - sum;
            if (remaining >= before && remaining < incl) {   // exactly one lane
                uint32_t acc = before;
                int j = 0;
                while (j < 7 && remaining >= acc + c[j]) { acc += c[j]; ++j; }
                hist[256] = lane * 8 + (uint32_t)j;
                hist[257] = acc;
                hist[258] = c[j];
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)hist[256] << shift;
        remaining -= hist[257];
        __syncthreads();
    }
}

// Merge (running best) + (candidates) -> new best, dedupe, new threshold. One block per query.
// A row can appear twice (once in best, once re-appended by a chunk re-run after overflow), never more, so the
// 2k smallest keys with multiplicity always hold the k smallest distinct ones: a radix select cuts the working
// set to those before the sort.
// LOWSMEM: the candidates are not staged in shared memory (the radix passes re-read them through L2), so the block needs
// only the sort buffer (4 KB for k <= 255) and can run on an SM whose shared memory belongs to a resident scan_tc CTA —
// the two query halves of a batch overlap this kernel with the other half's tensor pass (launch_exact_scan).
template <bool LOWSMEM>
__global__ void __launch_bounds__(256)
select_kernel(unsigned long long* __restrict__ cand, uint32_t* __restrict__ cand_cnt, uint32_t cap,
              unsigned long long* __restrict__ best, uint32_t* __restrict__ best_cnt, uint32_t k, uint32_t kpad,
              unsigned long long* __restrict__ thr, uint32_t nq, uint32_t sort_cap) {
    extern __shared__ unsigned long long skeys[];   // [sort_cap] sort buffer (+ [cap + kpad] staging unless LOWSMEM)
    __shared__ uint32_t s_hist[259];
    __shared__ uint32_t s_n;
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    uint32_t nc = cand_cnt[q];
    if (nc > cap) nc = cap;
    const uint32_t nb = best_cnt[q];
    if (nc == 0) return;  // nothing new: best/thr unchanged
    uint32_t total = nc + nb;
    unsigned long long* sortbuf = skeys;
    const unsigned long long* gbest = best + (size_t)q * kpad;
    const unsigned long long* gcand = cand + (size_t)q * cap;
    auto gkey = [&](uint32_t i) { return i < nb ? gbest[i] : gcand[i - nb]; };
    if (total > 2 * k + 1 && total > 512) {
        unsigned long long* stage = skeys + sort_cap;
        if (!LOWSMEM) {
            for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) stage[i] = gkey(i);
        }
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        auto skey = [&](uint32_t i) { return LOWSMEM ? gkey(i) : stage[i]; };
        // The cut is the (2k)-th smallest key. Its high half (the score bits) is found in 4 passes; the keys that tie with it
        // on the score are all kept when they fit in the sort buffer beside the strictly better ones (the usual case: one or
        // two keys), otherwise 4 more passes over the row-id half pick exactly the ones needed.
        unsigned long long cut = 0;
        uint32_t remaining = 2 * k - 1;
        block_radix_select<0, 3>(skey, total, cut, remaining, s_hist);
        const uint32_t less = 2 * k - 1 - remaining, ties = s_hist[258];
        __syncthreads();
        if (less + ties <= sort_cap) cut |= 0xFFFFFFFFull;
        else block_radix_select<4, 7>(skey, total, cut, remaining, s_hist);
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            const unsigned long long v = skey(i);
            if (v <= cut && v != ~0ull) {
                uint32_t pos = atomicAdd(&s_n, 1u);
                if (pos < sort_cap) sortbuf[pos] = v;
            }
        }
        __syncthreads();
        total = min(s_n, sort_cap);   // <= sort_cap by construction (2k + 1 after the exact cut; less + ties after the short one)
        uint32_t n2 = 1;
        while (n2 < total) n2 <<= 1;
        for (uint32_t i = total + threadIdx.x; i < n2; i += blockDim.x) sortbuf[i] = ~0ull;
        block_bitonic_sort(sortbuf, n2);
    } else {
        uint32_t n2 = 1;
        while (n2 < total) n2 <<= 1;
        for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) sortbuf[i] = i < total ? gkey(i) : ~0ull;
        block_bitonic_sort(sortbuf, n2);
    }
    // dedupe adjacent equal keys; single thread compaction of at most k survivors.
    if (threadIdx.x == 0) {
        uint32_t o = 0;
        unsigned long long prev = ~0ull;
        for (uint32_t i = 0; i < total && o < k; ++i) {
            unsigned long long v = sortbuf[i];
            if (v == ~0ull) break;
            if (i > 0 && v == prev) continue;
            best[(size_t)q * kpad + o++] = v;
            prev = v;
        }
        best_cnt[q] = o;
        thr[q] = (o == k) ? best[(size_t)q * kpad + k - 1] : ~0ull;
        cand_cnt[q] = 0;
    }
}

__global__ void scan_init_kernel(uint32_t* cand_cnt, uint32_t* best_cnt, unsigned long long* thr, uint32_t nq, uint32_t* overflow,
                                 uint32_t first_rows) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) { cand_cnt[i] = first_rows; best_cnt[i] = 0; thr[i] = ~0ull; }
    if (i == 0) { overflow[0] = 0; overflow[1] = 0xFFFFFFFFu; }
}
// overflow[0]: set by a scoring kernel that had to drop a survivor; overflow[1]: first round where that happened.
__global__ void note_overflow_kernel(uint32_t* overflow, uint32_t round) {
    if (overflow[0] && overflow[1] == 0xFFFFFFFFu) overflow[1] = round;
}
__global__ void scan_reset_overflow_kernel(uint32_t* overflow) { overflow[0] = 0; overflow[1] = 0xFFFFFFFFu; }

__global__ void scan_finish_kernel(const unsigned long long* __restrict__ best, const uint32_t* __restrict__ best_cnt,
                                   uint32_t k, uint32_t kpad, uint32_t nq, int metric, uint64_t* __restrict__ keys,
                                   float* __restrict__ dists, uint32_t* __restrict__ counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * k) return;
    uint32_t q = i / k, j = i % k;
    uint32_t c = best_cnt[q];
    if (j == 0 && counts) counts[q] = c;
    if (j < c) {
        unsigned long long v = best[(size_t)q * kpad + j];
        uint32_t ok = (uint32_t)(v >> 32);
        keys[i] = (uint64_t)(v & 0xFFFFFFFFull);
        dists[i] = metric == LEANN_METRIC_DOT_DESC ? unorder_f32(~ok) : unorder_f32(ok);
    } else {
        keys[i] = ~0ull;
        dists[i] = metric == LEANN_METRIC_DOT_DESC ? -CUDART_INF_F : CUDART_INF_F;
    }
}

__global__ void pad_rows_kernel(const float* __restrict__ src, float4* __restrict__ dst, size_t n, uint32_t d, uint32_t d4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * d4) return;
    size_t r = i / d4;
    uint32_t c = (uint32_t)(i % d4) * 4;
    const float* p = src + r * d;
    float4 v;
    v.x = c + 0 < d ? p[c + 0] : 0.f;
    v.y = c + 1 < d ? p[c + 1] : 0.f;
    v.z = c + 2 < d ? p[c + 2] : 0.f;
    v.w = c + 3 < d ? p[c + 3] : 0.f;
    dst[i] = v;
}

// K4: merge n_shards x k sorted lists per query. One block (128 threads) per query.
__global__ void __launch_bounds__(128)
topk_merge_kernel(const uint64_t* __restrict__ keys_in, const float* __restrict__ dists_in, uint32_t n_shards,
                  uint32_t nq, uint32_t k, int descending, uint64_t* __restrict__ keys_out,
                  float* __restrict__ dists_out, uint32_t* __restrict__ counts_out) {
    extern __shared__ unsigned long long skeys[];  // n2 packed (rank, position) + position -> source index
    const uint32_t q = blockIdx.x;
    const uint32_t total = n_shards * k;
    uint32_t n2 = 1;
    while (n2 < total) n2 <<= 1;
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) {
        unsigned long long v = ~0ull;
        if (i < total) {
            uint32_t s = i / k, j = i % k;
            size_t src = ((size_t)s * nq + q) * k + j;
            if (keys_in[src] != ~0ull) {
                uint32_t ok = order_f32(dists_in[src]);
                if (descending) ok = ~ok;
                v = ((unsigned long long)ok << 32) | i;  // ties: lower shard, then lower rank first
            }
        }
        skeys[i] = v;
    }
    block_bitonic_sort(skeys, n2);
    uint32_t cnt = 0;
    for (uint32_t j = threadIdx.x; j < k; j += blockDim.x) {
        unsigned long long v = j < n2 ? skeys[j] : ~0ull;
        size_t dst = (size_t)q * k + j;
        if (v != ~0ull) {
            uint32_t i = (uint32_t)(v & 0xFFFFFFFFull);
            uint32_t s = i / k, jj = i % k;
            size_t src = ((size_t)s * nq + q) * k + jj;
            keys_out[dst] = keys_in[src];
            dists_out[dst] = dists_in[src];
        } else {
            keys_out[dst] = ~0ull;
            dists_out[dst] = descending ? -CUDART_INF_F : CUDART_INF_F;
        }
    }
    if (counts_out) {
        __syncthreads();
        if (threadIdx.x == 0) {
            for (uint32_t j = 0; j < k && j < n2; ++j) if (skeys[j] != ~0ull) cnt++;
            counts_out[q] = cnt;
        }
    }
}

constexpr uint32_t SCAN_CAP = 4096;

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

void launch_pad_rows(const float* src, float4* dst, size_t n, uint32_t d, uint32_t d4, cudaStream_t stream) {
    size_t total = n * d4;
    if (!total) return;
    pad_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, dst, n, d, d4);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t exact_scan_scratch_bytes(uint32_t d4, uint32_t nq, uint32_t k) {
    uint32_t kpad = (k + 31) & ~31u;
    size_t d4max = d4;
    size_t dp8 = ((size_t)d4 * 4 + 4 + 7) / 8 * 8;   // >= exact_scan_tc_dp8 for every metric
    return align256((size_t)nq * SCAN_CAP * 8) + align256((size_t)nq * 4) + align256((size_t)nq * kpad * 8) +
           align256((size_t)nq * 4) + align256((size_t)nq * 8) + 256 + align256((size_t)nq * d4max * 16) +
           // tensor path: bf16 queries, |q|, dot thresholds, candidate row ids
           align256((size_t)nq * dp8 * 2) + 4 * align256((size_t)nq * 4) + 2 * align256((size_t)nq * SCAN_CAP * 4);
}

// One half (or all) of a query batch: its own slice of every per-query scratch array, its own overflow words and stream.
struct ScanJob {
    uint32_t q0 = 0, nq = 0;
    ScanScratch s{};
    TcScratch ts{};
    cudaStream_t stream = nullptr;
    uint32_t* h_flag = nullptr;   // 2 pinned words
    size_t first = 0;             // first round to (re)run
    bool done = false;
};

void launch_exact_scan(const FlatView& f, const float* d_queries, uint32_t nq, uint32_t k, const uint64_t* d_mask,
                       uint64_t* d_keys, float* d_dists, uint32_t* d_counts, void* scratch, size_t scratch_bytes,
                       cudaStream_t stream, const TcIndexView* tv, int sms, const ScanAux& aux) {
    if (k == 0 || k > 1024) throw Error(LEANN_ERR_INVALID_ARG, "exact scan: k must be in 1..1024");
    if (nq == 0) return;
    const uint32_t kpad = (k + 31) & ~31u;
    unsigned char* p = (unsigned char*)scratch;
    ScanScratch s;
    s.cand = (unsigned long long*)p; p += align256((size_t)nq * SCAN_CAP * 8);
    s.cand_cnt = (uint32_t*)p; p += align256((size_t)nq * 4);
    s.best = (unsigned long long*)p; p += align256((size_t)nq * kpad * 8);
    s.best_cnt = (uint32_t*)p; p += align256((size_t)nq * 4);
    s.thr = (unsigned long long*)p; p += align256((size_t)nq * 8);
    s.overflow = (uint32_t*)p; p += 256;
    s.qpad = (float4*)p; p += align256((size_t)nq * f.d4 * 16);
    TcScratch ts{};
    const bool use_tc = tv != nullptr && exact_scan_tc_supported(f, nq);
    if (use_tc) {
        ts.q_bf16 = p; p += align256((size_t)nq * tv->dp8 * 2);
        ts.qnorm = (float*)p; p += align256((size_t)nq * 4);
        ts.qres = (float*)p; p += align256((size_t)nq * 4);
        ts.thr_dot = (float*)p; p += align256((size_t)nq * 4);
        ts.cut_slack = (float*)p; p += align256((size_t)nq * 4);
        ts.cand_ids = (uint32_t*)p; p += align256((size_t)nq * SCAN_CAP * 4);
        ts.cand_sc = (float*)p; p += align256((size_t)nq * SCAN_CAP * 4);
    }
    if ((size_t)(p - (unsigned char*)scratch) > scratch_bytes) throw Error(LEANN_ERR_INVALID_ARG, "exact scan: scratch too small");

    // Two jobs when the batch is large enough to keep the tensor pipe busy with half of it: the halves run the same rounds
    // on two streams, so that the f32 re-rank and the select of one half (HBM / latency bound, small shared memory)
    // execute while the other half's scan_tc CTAs occupy the tensor cores. Per-query results do not depend on the split.
    const bool split = use_tc && aux.helper != nullptr && nq >= 2048;
    ScanJob jobs[2];
    const uint32_t h = split ? (uint32_t)(((nq / 2) + 255) / 256 * 256) : nq;
    const int n_jobs = split ? 2 : 1;
    for (int j = 0; j < n_jobs; ++j) {
        ScanJob& J = jobs[j];
        J.q0 = j == 0 ? 0 : h;
        J.nq = j == 0 ? h : nq - h;
        J.stream = j == 0 ? stream : aux.helper;
        J.h_flag = aux.h_flags + 2 * j;
        J.s.cand = s.cand + (size_t)J.q0 * SCAN_CAP; J.s.cand_cnt = s.cand_cnt + J.q0; J.s.best = s.best + (size_t)J.q0 * kpad;
        J.s.best_cnt = s.best_cnt + J.q0; J.s.thr = s.thr + J.q0; J.s.overflow = s.overflow + 2 * j; J.s.qpad = s.qpad + (size_t)J.q0 * f.d4;
        if (use_tc) {
            J.ts.q_bf16 = (unsigned char*)ts.q_bf16 + (size_t)J.q0 * tv->dp8 * 2; J.ts.qnorm = ts.qnorm + J.q0; J.ts.qres = ts.qres + J.q0;
            J.ts.thr_dot = ts.thr_dot + J.q0; J.ts.cand_ids = ts.cand_ids + (size_t)J.q0 * SCAN_CAP;
            J.ts.cut_slack = ts.cut_slack + J.q0; J.ts.cand_sc = ts.cand_sc + (size_t)J.q0 * SCAN_CAP;
        }
    }
    if (split) {   // the helper stream starts after everything already queued on the caller's stream
        LEANN_CUDA_CHECK(cudaEventRecord(aux.fork, stream));
        LEANN_CUDA_CHECK(cudaStreamWaitEvent(aux.helper, aux.fork, 0));
    }

    // Round boundaries: the first chunk (f32 tiles, every row is a candidate, cannot overflow) holds SCAN_CAP rows, later
    // chunks grow geometrically so that the expected number of survivors per query and round stays near growth * k.
    // A smaller first chunk (LEANN_CUDA_SCAN_FIRST = rows, A/B) saves tile work but adds a round, and a round costs about 1 ms
    // of launches, re-rank and select for 10 000 queries: 1024 rows measured 13.3 vs 12.4 ms on 1.25M x 384 and 66.9 vs 67.2 ms
    // on 10M x 384 (profiles/r2_k2_first_chunk_ab.log): left at SCAN_CAP.
    uint32_t growth = std::max<uint32_t>(2u, SCAN_CAP / (8u * k));
    {
        static const int growth_env = [] { const char* e = getenv("LEANN_CUDA_SCAN_GROWTH"); return e ? atoi(e) : 0; }();   // A/B
        if (growth_env >= 2) growth = (uint32_t)growth_env;
    }
    uint32_t first_rows = SCAN_CAP;
    {
        static const int first_env = [] { const char* e = getenv("LEANN_CUDA_SCAN_FIRST"); return e ? atoi(e) : 0; }();
        if (first_env >= 128) first_rows = std::min<uint32_t>(SCAN_CAP, std::max<uint32_t>((uint32_t)first_env, k));
    }
    std::vector<std::pair<uint32_t, uint32_t>> rounds;
    for (uint32_t r0 = 0; r0 < f.n;) {
        uint64_t want = r0 == 0 ? first_rows : (uint64_t)r0 * growth;
        uint32_t r1 = (uint32_t)std::min<uint64_t>(f.n, (uint64_t)r0 + want);
        rounds.emplace_back(r0, r1);
        r0 = r1;
    }
    static unsigned long long attr_seen = 0;
    uint32_t sort_cap = 512;   // sort buffer: the 2k + 1 keys a radix select keeps, or a short list sorted whole
    while (sort_cap < 2 * k + 1) sort_cap <<= 1;
    const bool lowsmem = split;
    const size_t sel_smem = lowsmem ? (size_t)sort_cap * 8 : ((size_t)sort_cap + SCAN_CAP + kpad) * 8;
    if (first_use_on_device(attr_seen))
        LEANN_CUDA_CHECK(cudaFuncSetAttribute(select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)4096 + SCAN_CAP + 1024) * 8)));

    auto begin_job = [&](ScanJob& J) {
        scan_init_kernel<<<(J.nq + 255) / 256, 256, 0, J.stream>>>(J.s.cand_cnt, J.s.best_cnt, J.s.thr, J.nq, J.s.overflow,
                                                                   (uint32_t)std::min<uint64_t>(f.n, first_rows));
        launch_pad_rows(d_queries + (size_t)J.q0 * f.d, J.s.qpad, J.nq, f.d, f.d4, J.stream);
        if (use_tc) exact_scan_tc_queries(J.s.qpad, J.nq, f.d4, tv->dp8, f.metric, tv->xmax_bits, J.ts, J.stream);
    };
    auto run_round = [&](ScanJob& J, size_t ri, int attempt) {
        const uint32_t r0 = rounds[ri].first, r1 = rounds[ri].second;
        dim3 grid((r1 - r0 + TN - 1) / TN, (J.nq + TM - 1) / TM);
        if (use_tc && ri > 0) {
            // tensor-core pass + fp32 re-rank for every chunk after the first
            exact_scan_tc_round(f, *tv, J.s, J.ts, J.nq, k, r0, r1, d_mask, SCAN_CAP, sms, J.stream);
        } else {
            const bool direct = ri == 0 && attempt == 0;   // no threshold yet: slot = row offset, cand_cnt preset by scan_init
#define LEANN_TILE(M)                                                                                                            \
    (direct ? scan_tile_kernel<M, true><<<grid, 256, 0, J.stream>>>(f.vecs, J.s.qpad, f.d4, J.nq, r0, r1, d_mask, J.s.thr,        \
                                                                     J.s.cand, J.s.cand_cnt, SCAN_CAP, J.s.overflow)              \
            : scan_tile_kernel<M, false><<<grid, 256, 0, J.stream>>>(f.vecs, J.s.qpad, f.d4, J.nq, r0, r1, d_mask, J.s.thr,       \
                                                                      J.s.cand, J.s.cand_cnt, SCAN_CAP, J.s.overflow))
            switch (f.metric) {
                case LEANN_METRIC_L2SQ: LEANN_TILE(LEANN_METRIC_L2SQ); break;
                case LEANN_METRIC_IP_CLAMP: LEANN_TILE(LEANN_METRIC_IP_CLAMP); break;
                case LEANN_METRIC_DOT_DESC: LEANN_TILE(LEANN_METRIC_DOT_DESC); break;
                default: LEANN_TILE(LEANN_METRIC_IP);
            }
#undef LEANN_TILE
        }
        LEANN_CUDA_CHECK(cudaGetLastError());
        auto select = [&]() {
            if (lowsmem)
                select_kernel<true><<<J.nq, 256, sel_smem, J.stream>>>(J.s.cand, J.s.cand_cnt, SCAN_CAP, J.s.best, J.s.best_cnt, k, kpad, J.s.thr, J.nq, sort_cap);
            else
                select_kernel<false><<<J.nq, 256, sel_smem, J.stream>>>(J.s.cand, J.s.cand_cnt, SCAN_CAP, J.s.best, J.s.best_cnt, k, kpad, J.s.thr, J.nq, sort_cap);
        };
        select();
        if (use_tc && ri == 0 && rounds.size() > 1) {
            // the later chunks are scored by rerank_kernel: re-score the first chunk's top-k the same way (nq * k rows out of
            // L2), so that every score in `best` comes from one arithmetic and exact duplicate rows tie exactly
            exact_scan_tc_canonical_first(f, J.s, J.ts, J.nq, kpad, SCAN_CAP, J.stream);
            select();
        }
        note_overflow_kernel<<<1, 1, 0, J.stream>>>(J.s.overflow, (uint32_t)ri);
        LEANN_CUDA_CHECK(cudaGetLastError());
    };
    // All rounds of every job are enqueued back to back, round by round across the jobs so that the streams alternate;
    // the device records the first round whose candidate list overflowed (overflow[1]). One synchronisation per attempt; a
    // job that overflowed repeats that round and the ones after it with the tightened thresholds (rows already kept are
    // deduplicated by select_kernel).
    for (int j = 0; j < n_jobs; ++j) begin_job(jobs[j]);
    for (int attempt = 0;; ++attempt) {
        bool any = false;
        for (size_t ri = 0; ri < rounds.size(); ++ri)
            for (int j = 0; j < n_jobs; ++j)
                if (!jobs[j].done && ri >= jobs[j].first) { run_round(jobs[j], ri, attempt); any = true; }
        if (!any) break;
        for (int j = 0; j < n_jobs; ++j)
            if (!jobs[j].done) LEANN_CUDA_CHECK(cudaMemcpyAsync(jobs[j].h_flag, jobs[j].s.overflow, 8, cudaMemcpyDeviceToHost, jobs[j].stream));
        for (int j = 0; j < n_jobs; ++j)
            if (!jobs[j].done) LEANN_CUDA_CHECK(cudaStreamSynchronize(jobs[j].stream));
        bool again = false;
        for (int j = 0; j < n_jobs; ++j) {
            ScanJob& J = jobs[j];
            if (J.done) continue;
            if (J.h_flag[1] == 0xFFFFFFFFu) { J.done = true; continue; }
            J.first = J.h_flag[1];
            again = true;
            scan_reset_overflow_kernel<<<1, 1, 0, J.stream>>>(J.s.overflow);
        }
        if (!again) break;
        if (attempt > 64) throw Error(LEANN_ERR_CUDA, "exact scan: candidate overflow did not converge");
    }
    for (int j = 0; j < n_jobs; ++j) {
        ScanJob& J = jobs[j];
        const uint32_t tot = J.nq * k;
        scan_finish_kernel<<<(tot + 255) / 256, 256, 0, J.stream>>>(J.s.best, J.s.best_cnt, k, kpad, J.nq, f.metric, d_keys + (size_t)J.q0 * k,
                                                                     d_dists + (size_t)J.q0 * k, d_counts ? d_counts + J.q0 : nullptr);
        LEANN_CUDA_CHECK(cudaGetLastError());
    }
    if (split) {   // the caller's stream continues after the helper's half
        LEANN_CUDA_CHECK(cudaEventRecord(aux.join, aux.helper));
        LEANN_CUDA_CHECK(cudaStreamWaitEvent(stream, aux.join, 0));
    }
}

void launch_topk_merge(const uint64_t* keys_in, const float* dists_in, uint32_t n_shards, uint32_t nq, uint32_t k,
                       int descending, uint64_t* keys_out, float* dists_out, uint32_t* counts_out, cudaStream_t stream) {
    if (nq == 0 || k == 0) return;
    uint32_t total = n_shards * k, n2 = 1;
    while (n2 < total) n2 <<= 1;
    size_t smem = (size_t)n2 * 8;
    if (smem > 200 * 1024) throw Error(LEANN_ERR_INVALID_ARG, "topk merge: n_shards*k too large");
    if (smem > 48 * 1024) LEANN_CUDA_CHECK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_merge_kernel<<<nq, 128, smem, stream>>>(keys_in, dists_in, n_shards, nq, k, descending, keys_out, dists_out, counts_out);
    LEANN_CUDA_CHECK(cudaGetLastError());
}

}  // namespace leann
