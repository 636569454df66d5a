// bm25_build.cu — GPU construction of the BM25 inverted index: Bm25Scorer::build (leann-rs src/index/bm25.rs:33-74)
// with tokenize (:127-132) and the per-posting score of score_query (:88-100). The reference runs this on every hybrid
// query (searcher.rs:149-150); here it runs once per index, on the device:
//
//   1 tokens   : the passages are concatenated (one '\n' between them) and scanned in chunks; a token is a maximal run
//                of ASCII [A-Za-z0-9] of at least 2 bytes (regex [a-zA-Z0-9]+, to_lowercase, len > 1). Per token: 64-bit
//                hash of the lower-cased bytes, start offset, document id; doc_len by atomics.
//   2 sort     : stable radix sort of (hash -> token index): equal terms become adjacent, documents stay ascending.
//   3 segments : run boundaries of equal hash (term) and equal (hash, document) (posting); tf = run length, df =
//                postings per term; CSR offsets by prefix sums.
//   4 verify   : every token of a run is compared byte for byte with the run's first token: two different terms with
//                one hash make the build start over with another seed (never silently merged).
//   5 dictionary: the term strings go back to the host (query tokens are looked up there), idf is computed on the host
//                with libm logf exactly like f32::ln (bm25.rs:88; CUDA's logf is not bit-compatible), 4 bytes per term.
//   6 scores   : norm = 1 - B + B * (len / avg) and idf * (tf * (K1 + 1)) / (tf + K1 * norm) per posting with
//                round-to-nearest f32 intrinsics in the reference's operation order (no FMA contraction).
//
// For one document-range shard of a larger corpus the same pipeline runs with the corpus-wide N, token count and df
// (Bm25GlobalStats), and steps 1-5 alone produce the shard's statistics blob.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <cmath>

#include "bm25_dev.h"

namespace leann {

namespace {

constexpr int TOK_CHUNK = 4096;   // text bytes per CTA in the tokenizer kernels
constexpr int TOK_THREADS = 256;

__device__ __forceinline__ bool is_alnum(unsigned char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9');
}
__device__ __forceinline__ unsigned char lower(unsigned char c) { return (c >= 'A' && c <= 'Z') ? (unsigned char)(c + 32) : c; }
// token start: first byte of an alnum run that is at least two bytes long
__device__ __forceinline__ bool token_start(const unsigned char* text, uint64_t n, uint64_t i) {
    return is_alnum(text[i]) && (i == 0 || !is_alnum(text[i - 1])) && i + 1 < n && is_alnum(text[i + 1]);
}

__global__ void __launch_bounds__(TOK_THREADS)
count_tokens_kernel(const unsigned char* __restrict__ text, uint64_t n, uint32_t* __restrict__ chunk_count) {
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * TOK_CHUNK;
    uint32_t c = 0;
    for (int k = threadIdx.x; k < TOK_CHUNK; k += TOK_THREADS) {
        const uint64_t i = base + k;
        if (i < n && token_start(text, n, i)) ++c;
    }
    for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) chunk_count[blockIdx.x] = s_cnt;
}

// chunk_base: exclusive prefix sum of chunk_count. Tokens are numbered in text order.
__global__ void __launch_bounds__(TOK_THREADS)
emit_tokens_kernel(const unsigned char* __restrict__ text, uint64_t n, const uint32_t* __restrict__ chunk_base,
                   const uint64_t* __restrict__ doc_off, uint32_t n_docs, uint64_t seed, uint64_t* __restrict__ tok_hash,
                   uint64_t* __restrict__ tok_start, uint32_t* __restrict__ tok_doc, uint32_t* __restrict__ doc_len) {
    __shared__ uint32_t s_warp[TOK_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * TOK_CHUNK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t running = chunk_base[blockIdx.x];
    // the chunk is walked in slices of TOK_THREADS consecutive bytes so that token numbers follow text order
    for (int k0 = 0; k0 < TOK_CHUNK; k0 += TOK_THREADS) {
        const uint64_t i = base + k0 + threadIdx.x;
        const bool st = i < n && token_start(text, n, i);
        const unsigned b = __ballot_sync(0xFFFFFFFFu, st);
        if (lane == 0) s_warp[warp] = __popc(b);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < TOK_THREADS / 32; ++w) {
            const uint32_t c = s_warp[w];
            if (w < warp) before += c;
            total += c;
        }
        if (st) {
            const uint32_t t = running + before + __popc(b & ((1u << lane) - 1u));
            uint64_t h = 0xCBF29CE484222325ull ^ seed;      // FNV-1a over the lower-cased bytes, then a finaliser
            uint64_t j = i;
            while (j < n && is_alnum(text[j])) { h = (h ^ lower(text[j])) * 0x100000001B3ull; ++j; }
            h ^= h >> 32; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
            tok_hash[t] = h;
            tok_start[t] = i;
            // document of the token: last d with doc_off[d] <= i
            uint32_t lo = 0, hi = n_docs;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (doc_off[mid] <= i) lo = mid; else hi = mid;
            }
            tok_doc[t] = lo;
            atomicAdd(&doc_len[lo], 1u);
        }
        running += total;
        __syncthreads();
    }
}

// After the sort: j-th smallest hash, its token. f_term / f_post = 1 where a new term / posting starts.
__global__ void flag_runs_kernel(const uint64_t* __restrict__ hash, const uint32_t* __restrict__ tok, const uint32_t* __restrict__ tok_doc,
                                 uint32_t n_tok, uint32_t* __restrict__ f_term, uint32_t* __restrict__ f_post) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_tok) return;
    const bool nt = j == 0 || hash[j] != hash[j - 1];
    const bool np = nt || tok_doc[tok[j]] != tok_doc[tok[j - 1]];
    f_term[j] = nt ? 1u : 0u;
    f_post[j] = np ? 1u : 0u;
}

// i_term / i_post: inclusive prefix sums of the flags (1-based ids).
__global__ void scatter_runs_kernel(const uint32_t* __restrict__ tok, const uint32_t* __restrict__ tok_doc, const uint32_t* __restrict__ f_term,
                                    const uint32_t* __restrict__ i_term, const uint32_t* __restrict__ f_post, const uint32_t* __restrict__ i_post,
                                    uint32_t n_tok, uint32_t* __restrict__ post_doc, uint32_t* __restrict__ post_begin,
                                    uint32_t* __restrict__ post_term, uint64_t* __restrict__ term_off, uint32_t* __restrict__ term_tok) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_tok) return;
    if (f_post[j]) {
        const uint32_t p = i_post[j] - 1u;
        post_doc[p] = tok_doc[tok[j]];
        post_begin[p] = j;
        post_term[p] = i_term[j] - 1u;
    }
    if (f_term[j]) {
        const uint32_t t = i_term[j] - 1u;
        term_off[t] = (uint64_t)(i_post[j] - 1u);
        term_tok[t] = tok[j];
    }
}

// every token of a term run must spell the run's first token
__global__ void verify_terms_kernel(const unsigned char* __restrict__ text, uint64_t n, const uint32_t* __restrict__ tok,
                                    const uint64_t* __restrict__ tok_start, const uint32_t* __restrict__ f_term,
                                    const uint32_t* __restrict__ i_term, const uint32_t* __restrict__ term_tok, uint32_t n_tok,
                                    uint32_t* __restrict__ mismatch) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_tok || f_term[j]) return;
    uint64_t a = tok_start[tok[j]], b = tok_start[term_tok[i_term[j] - 1u]];
    for (;;) {
        const bool ea = a >= n || !is_alnum(text[a]), eb = b >= n || !is_alnum(text[b]);
        if (ea || eb) { if (ea != eb) *mismatch = 1u; return; }
        if (lower(text[a]) != lower(text[b])) { *mismatch = 1u; return; }
        ++a; ++b;
    }
}

__global__ void term_len_kernel(const unsigned char* __restrict__ text, uint64_t n, const uint32_t* __restrict__ term_tok,
                                const uint64_t* __restrict__ tok_start, uint32_t n_terms, uint32_t* __restrict__ len) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_terms) return;
    uint64_t a = tok_start[term_tok[t]];
    uint32_t l = 0;
    while (a < n && is_alnum(text[a])) { ++a; ++l; }
    len[t] = l;
}
__global__ void term_copy_kernel(const unsigned char* __restrict__ text, const uint32_t* __restrict__ term_tok,
                                 const uint64_t* __restrict__ tok_start, const uint32_t* __restrict__ len,
                                 const uint64_t* __restrict__ str_off, uint32_t n_terms, unsigned char* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_terms) return;
    const uint64_t a = tok_start[term_tok[t]], o = str_off[t];
    for (uint32_t k = 0; k < len[t]; ++k) out[o + k] = lower(text[a + k]);
}
__global__ void iota_kernel(uint32_t* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
__global__ void widen_kernel(const uint32_t* __restrict__ in, uint32_t n, uint64_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// bm25.rs:96-100 per posting, f32 round-to-nearest in the reference's order
__global__ void score_postings_kernel(const uint32_t* __restrict__ post_doc, const uint32_t* __restrict__ post_begin,
                                      const uint32_t* __restrict__ post_term, uint32_t n_post, uint32_t n_tok,
                                      const float* __restrict__ idf, const uint32_t* __restrict__ doc_len, float avg_doc_len,
                                      float* __restrict__ post_score) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_post) return;
    const uint32_t end = p + 1 < n_post ? post_begin[p + 1] : n_tok;
    const float tf = (float)(end - post_begin[p]);
    const float K1 = 1.2f, B = 0.75f;
    const float ratio = __fdiv_rn((float)doc_len[post_doc[p]], avg_doc_len);
    const float norm = __fadd_rn(__fsub_rn(1.0f, B), __fmul_rn(B, ratio));
    const float num = __fmul_rn(idf[post_term[p]], __fmul_rn(tf, __fadd_rn(K1, 1.0f)));
    const float den = __fadd_rn(tf, __fmul_rn(K1, norm));
    post_score[p] = __fdiv_rn(num, den);
}

struct Scratch {   // frees everything it handed out
    std::vector<void*> ptrs;
    ~Scratch() { for (void* p : ptrs) cudaFree(p); }
    template <typename T> T* get(size_t count) {
        void* p = nullptr;
        LEANN_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
        ptrs.push_back(p);
        return reinterpret_cast<T*>(p);
    }
    void release(void* p) {
        for (auto& q : ptrs) if (q == p) { q = nullptr; return; }
    }
};

inline unsigned blocks_for(size_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

// Runs the pipeline. stats_only: stop after the dictionary (per-term df, N, token count) and fill `stats`.
// Otherwise fills b->host {num_docs, total_tokens, avg_doc_len, dict, term_off, idf, n_postings} and the device arrays.
void bm25_build_device(const char* const* docs, const size_t* doc_bytes, size_t n_docs, const Bm25GlobalStats* glob,
                       leann_cuda_bm25* b, Bm25GlobalStats* stats, bool stats_only) {
    // ---- host: concatenate with one separator byte per passage ----
    std::vector<uint64_t> doc_off(n_docs + 1, 0);
    for (size_t d = 0; d < n_docs; ++d) doc_off[d + 1] = doc_off[d] + doc_bytes[d] + 1;
    const uint64_t n_bytes = doc_off[n_docs];
    std::vector<unsigned char> text(std::max<uint64_t>(n_bytes, 1));
    for (size_t d = 0; d < n_docs; ++d) {
        if (doc_bytes[d]) memcpy(text.data() + doc_off[d], docs[d], doc_bytes[d]);
        text[doc_off[d] + doc_bytes[d]] = '\n';
    }
    if (n_bytes / 3 >= 0x7FFFFFF0ull) throw Error(LEANN_ERR_INVALID_ARG, "bm25: corpus too large for one index (shard it by document range)");
    Scratch sc;
    cudaStream_t s = nullptr;   // legacy default stream: the build is a synchronous call
    unsigned char* d_text = sc.get<unsigned char>(n_bytes);
    uint64_t* d_doc_off = sc.get<uint64_t>(n_docs + 1);
    if (n_bytes) LEANN_CUDA_CHECK(cudaMemcpy(d_text, text.data(), n_bytes, cudaMemcpyHostToDevice));
    LEANN_CUDA_CHECK(cudaMemcpy(d_doc_off, doc_off.data(), (n_docs + 1) * 8, cudaMemcpyHostToDevice));
    text.clear(); text.shrink_to_fit();

    // ---- 1 tokens ----
    const size_t n_chunks = (size_t)((n_bytes + TOK_CHUNK - 1) / TOK_CHUNK);
    uint32_t* d_chunk_cnt = sc.get<uint32_t>(n_chunks + 1);
    uint32_t* d_chunk_base = sc.get<uint32_t>(n_chunks + 1);
    LEANN_CUDA_CHECK(cudaMemsetAsync(d_chunk_cnt, 0, (n_chunks + 1) * 4, s));
    if (n_chunks) count_tokens_kernel<<<(unsigned)n_chunks, TOK_THREADS, 0, s>>>(d_text, n_bytes, d_chunk_cnt);
    size_t tmp_bytes = 0;
    void* d_tmp = nullptr;
    auto ensure_tmp = [&](size_t need) {
        if (need <= tmp_bytes) return;
        d_tmp = sc.get<unsigned char>(need);
        tmp_bytes = need;
    };
    {
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, d_chunk_cnt, d_chunk_base, (int)(n_chunks + 1), s);
        ensure_tmp(need);
        cub::DeviceScan::ExclusiveSum(d_tmp, need, d_chunk_cnt, d_chunk_base, (int)(n_chunks + 1), s);
    }
    uint32_t n_tok = 0;
    LEANN_CUDA_CHECK(cudaMemcpy(&n_tok, d_chunk_base + n_chunks, 4, cudaMemcpyDeviceToHost));
    // token numbers are 31-bit (cub item counts): refuse corpora that do not fit instead of wrapping
    if (n_bytes / 3 >= 0x7FFFFFF0ull || n_tok >= 0x7FFFFFF0u || n_docs >= 0xFFFFFFF0ull)
        throw Error(LEANN_ERR_INVALID_ARG, "bm25: corpus too large for one index (shard it by document range)");
    uint32_t* d_doc_len = sc.get<uint32_t>(n_docs);
    LEANN_CUDA_CHECK(cudaMemsetAsync(d_doc_len, 0, std::max<size_t>(n_docs, 1) * 4, s));
    uint64_t* d_hash = sc.get<uint64_t>(n_tok);
    uint64_t* d_hash2 = sc.get<uint64_t>(n_tok);
    uint64_t* d_start = sc.get<uint64_t>(n_tok);
    uint32_t* d_tdoc = sc.get<uint32_t>(n_tok);
    uint32_t* d_tok = sc.get<uint32_t>(n_tok);
    uint32_t* d_tok2 = sc.get<uint32_t>(n_tok);
    uint32_t* d_fterm = sc.get<uint32_t>(n_tok);
    uint32_t* d_fpost = sc.get<uint32_t>(n_tok);
    uint32_t* d_iterm = sc.get<uint32_t>(n_tok);
    uint32_t* d_ipost = sc.get<uint32_t>(n_tok);
    uint32_t* d_mismatch = sc.get<uint32_t>(1);
    uint32_t n_terms = 0, n_post = 0;
    uint32_t* d_post_doc = nullptr; uint32_t* d_post_begin = nullptr; uint32_t* d_post_term = nullptr;
    uint64_t* d_term_off = nullptr; uint32_t* d_term_tok = nullptr;
    for (int attempt = 0;; ++attempt) {
        const uint64_t seed = 0x9E3779B97F4A7C15ull * (uint64_t)attempt;
        if (attempt) LEANN_CUDA_CHECK(cudaMemsetAsync(d_doc_len, 0, std::max<size_t>(n_docs, 1) * 4, s));
        if (n_chunks && n_tok)
            emit_tokens_kernel<<<(unsigned)n_chunks, TOK_THREADS, 0, s>>>(d_text, n_bytes, d_chunk_base, d_doc_off, (uint32_t)n_docs, seed,
                                                                       d_hash, d_start, d_tdoc, d_doc_len);
        LEANN_CUDA_CHECK(cudaGetLastError());
        if (n_tok == 0) break;
        // ---- 2 sort (stable): token order inside a term stays text order, i.e. documents ascending ----
        {
            iota_kernel<<<blocks_for(n_tok), 256, 0, s>>>(d_tok2, n_tok);   // sort values: token numbers 0..n_tok-1
            size_t need = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, need, d_hash, d_hash2, d_tok2, d_tok, (int)n_tok, 0, 64, s);
            ensure_tmp(need);
            cub::DeviceRadixSort::SortPairs(d_tmp, need, d_hash, d_hash2, d_tok2, d_tok, (int)n_tok, 0, 64, s);
        }
        // ---- 3 segments ----
        flag_runs_kernel<<<blocks_for(n_tok), 256, 0, s>>>(d_hash2, d_tok, d_tdoc, n_tok, d_fterm, d_fpost);
        {
            size_t need = 0;
            cub::DeviceScan::InclusiveSum(nullptr, need, d_fterm, d_iterm, (int)n_tok, s);
            ensure_tmp(need);
            cub::DeviceScan::InclusiveSum(d_tmp, need, d_fterm, d_iterm, (int)n_tok, s);
            cub::DeviceScan::InclusiveSum(d_tmp, need, d_fpost, d_ipost, (int)n_tok, s);
        }
        LEANN_CUDA_CHECK(cudaMemcpyAsync(&n_terms, d_iterm + (n_tok - 1), 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaMemcpyAsync(&n_post, d_ipost + (n_tok - 1), 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
        d_post_doc = sc.get<uint32_t>(n_post);
        d_post_begin = sc.get<uint32_t>(n_post);
        d_post_term = sc.get<uint32_t>(n_post);
        d_term_off = sc.get<uint64_t>((size_t)n_terms + 1);
        d_term_tok = sc.get<uint32_t>(n_terms);
        scatter_runs_kernel<<<blocks_for(n_tok), 256, 0, s>>>(d_tok, d_tdoc, d_fterm, d_iterm, d_fpost, d_ipost, n_tok, d_post_doc,
                                                              d_post_begin, d_post_term, d_term_off, d_term_tok);
        const uint64_t n_post64 = n_post;
        LEANN_CUDA_CHECK(cudaMemcpyAsync(d_term_off + n_terms, &n_post64, 8, cudaMemcpyHostToDevice, s));
        // ---- 4 verify ----
        LEANN_CUDA_CHECK(cudaMemsetAsync(d_mismatch, 0, 4, s));
        verify_terms_kernel<<<blocks_for(n_tok), 256, 0, s>>>(d_text, n_bytes, d_tok, d_start, d_fterm, d_iterm, d_term_tok, n_tok, d_mismatch);
        uint32_t mismatch = 0;
        LEANN_CUDA_CHECK(cudaMemcpyAsync(&mismatch, d_mismatch, 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
        LEANN_CUDA_CHECK(cudaGetLastError());
        if (!mismatch) break;
        if (attempt >= 4) throw Error(LEANN_ERR_CUDA, "bm25 build: term hash collisions persisted over 5 seeds");
    }

    // ---- 5 dictionary ----
    std::vector<uint64_t> term_off(n_terms + 1, 0);
    std::vector<uint32_t> term_len(n_terms);
    std::vector<uint64_t> str_off(n_terms + 1, 0);
    std::vector<unsigned char> strs;
    if (n_terms) {
        uint32_t* d_len = sc.get<uint32_t>(n_terms);
        uint64_t* d_len64 = sc.get<uint64_t>((size_t)n_terms + 1);
        uint64_t* d_str_off = sc.get<uint64_t>((size_t)n_terms + 1);
        term_len_kernel<<<blocks_for(n_terms), 256, 0, s>>>(d_text, n_bytes, d_term_tok, d_start, n_terms, d_len);
        LEANN_CUDA_CHECK(cudaMemsetAsync(d_len64, 0, ((size_t)n_terms + 1) * 8, s));
        widen_kernel<<<blocks_for(n_terms), 256, 0, s>>>(d_len, n_terms, d_len64);
        size_t need = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, need, d_len64, d_str_off, (int)(n_terms + 1), s);
        ensure_tmp(need);
        cub::DeviceScan::ExclusiveSum(d_tmp, need, d_len64, d_str_off, (int)(n_terms + 1), s);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(str_off.data(), d_str_off, ((size_t)n_terms + 1) * 8, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaMemcpyAsync(term_off.data(), d_term_off, ((size_t)n_terms + 1) * 8, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
        strs.resize(std::max<uint64_t>(str_off[n_terms], 1));
        unsigned char* d_strs = sc.get<unsigned char>(strs.size());
        term_copy_kernel<<<blocks_for(n_terms), 256, 0, s>>>(d_text, d_term_tok, d_start, d_len, d_str_off, n_terms, d_strs);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(strs.data(), d_strs, strs.size(), cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
    }
    auto term_str = [&](uint32_t t) { return std::string(reinterpret_cast<const char*>(strs.data()) + str_off[t], (size_t)(str_off[t + 1] - str_off[t])); };
    if (stats_only) {
        *stats = Bm25GlobalStats();
        stats->num_docs = n_docs;
        stats->total_tokens = n_tok;
        stats->df.reserve((size_t)n_terms * 2);
        for (uint32_t t = 0; t < n_terms; ++t) stats->df[term_str(t)] = term_off[t + 1] - term_off[t];
        LEANN_CUDA_CHECK(cudaGetLastError());
        return;
    }
    Bm25Host& o = b->host;
    o = Bm25Host();
    o.num_docs = n_docs;
    o.total_tokens = n_tok;
    o.n_postings = n_post;
    o.dict.reserve((size_t)n_terms * 2);
    for (uint32_t t = 0; t < n_terms; ++t) o.dict.emplace(term_str(t), t);
    o.term_off = term_off;
    // bm25.rs:61-65 / :88 on the host (libm logf == f32::ln); over the whole corpus when this is one shard of it
    const uint64_t N_all = glob ? glob->num_docs : (uint64_t)n_docs;
    const uint64_t tokens_all = glob ? glob->total_tokens : (uint64_t)n_tok;
    o.avg_doc_len = N_all > 0 ? (float)tokens_all / (float)N_all : 1.0f;
    o.idf.resize(n_terms);
    const float Nf = (float)N_all;
    for (uint32_t t = 0; t < n_terms; ++t) {
        uint64_t df = term_off[t + 1] - term_off[t];
        if (glob) {
            auto it = glob->df.find(term_str(t));
            if (it != glob->df.end()) df = it->second;
        }
        const float dff = (float)df;
        volatile float a = Nf - dff;
        volatile float num = a + 0.5f;
        volatile float den = dff + 0.5f;
        volatile float r = num / den;
        volatile float sum = r + 1.0f;
        o.idf[t] = logf(sum);
    }
    // ---- 6 scores ----
    float* d_idf = sc.get<float>(n_terms);
    if (n_terms) LEANN_CUDA_CHECK(cudaMemcpyAsync(d_idf, o.idf.data(), (size_t)n_terms * 4, cudaMemcpyHostToDevice, s));
    float* d_score = sc.get<float>(n_post);
    if (n_post)
        score_postings_kernel<<<blocks_for(n_post), 256, 0, s>>>(d_post_doc, d_post_begin, d_post_term, n_post, n_tok, d_idf, d_doc_len,
                                                                  o.avg_doc_len, d_score);
    LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
    LEANN_CUDA_CHECK(cudaGetLastError());
    if (!d_term_off) {   // empty corpus: a single zero offset
        d_term_off = sc.get<uint64_t>(1);
        LEANN_CUDA_CHECK(cudaMemset(d_term_off, 0, 8));
    }
    if (!d_post_doc) d_post_doc = sc.get<uint32_t>(1);
    b->d_term_off = d_term_off; sc.release(d_term_off);
    b->d_post_doc = d_post_doc; sc.release(d_post_doc);
    b->d_post_score = d_score; sc.release(d_score);
    bm25_build_dense_rows(b);
}

}  // namespace leann
