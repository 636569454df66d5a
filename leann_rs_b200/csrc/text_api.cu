// text_api.cu — C ABI of the text path: BM25 handle, filter handle, hybrid search, IndexSearcher.
// Reference interfaces: src/index/bm25.rs, src/index/filter.rs, src/index/searcher.rs,
// src/index/passages.rs (read side only: <base>.passages.jsonl + <base>.passages.idx.json + <base>.ids.txt).
#include <algorithm>
#include <fstream>
#include <functional>
#include <memory>
#include <sstream>
#include <thread>

#include "bm25_dev.h"

using namespace leann;

namespace leann {
int guard_impl(char* err, size_t errlen, const std::function<void()>& f);
// sharded backend internals (shards.cu)
void shards_describe(const leann_cuda_shards* sh, int* world, int* rank, int* device, uint64_t* offsets16, size_t* local_shards);
cudaStream_t shards_stream(leann_cuda_shards* sh);
void shards_vector_search_device(leann_cuda_shards* sh, const float* dq, size_t nq, size_t k, size_t ef, uint64_t* dk, float* dd,
                                 uint32_t* dc, cudaStream_t st);
void shards_all_gather(leann_cuda_shards* sh, const void* send, void* recv, size_t bytes, cudaStream_t st);
void shards_merge_lists(const uint64_t* const* keys, const float* const* dists, const uint64_t* offsets, uint32_t g, uint32_t nq, uint32_t k,
                        int descending, uint64_t* ok, float* od, uint32_t* oc, cudaStream_t st);
}
#define GUARD(...) return leann::guard_impl(err, errlen, [&]() __VA_ARGS__)

struct leann_cuda_filter {
    leann::FilterNode root;
    std::string expr;
};
struct leann_cuda_metacols {
    leann::MetaColumns cols;
};

namespace {

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        cudaError_t e = cudaSetDevice(dev);
        if (e != cudaSuccess) throw Error(LEANN_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e) + " (no CPU fallback)");
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Scoped device scratch of the synchronous host-pointer entry points. Blocks come from a small per-device cache
// (power-of-two buckets, at most 1 GiB kept): steady-state calls make no cudaMalloc / cudaFree, whose implicit
// device synchronisation and page (un)mapping showed up as 100+ ms outliers in batch latency.
class ScratchCache {
public:
    static ScratchCache& get() { static ScratchCache c; return c; }
    void* take(size_t& bytes) {
        size_t b = 256;
        while (b < bytes) b <<= 1;
        bytes = b;
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto& fl = free_[key(dev, b)];
            if (!fl.empty()) { void* p = fl.back(); fl.pop_back(); cached_ -= b; return p; }
        }
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, b);
        if (e != cudaSuccess) { trim(); LEANN_CUDA_CHECK(cudaMalloc(&p, b)); }
        return p;
    }
    void give(void* p, size_t bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (cached_ + bytes <= ((size_t)1 << 30)) { free_[key(dev, bytes)].push_back(p); cached_ += bytes; return; }
        }
        cudaFree(p);
    }
    void trim() {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto& kv : free_) for (void* p : kv.second) cudaFree(p);
        free_.clear();
        cached_ = 0;
        cudaGetLastError();
    }
private:
    static uint64_t key(int dev, size_t b) { return ((uint64_t)dev << 56) | (uint64_t)b; }
    std::mutex mu_;
    std::unordered_map<uint64_t, std::vector<void*>> free_;
    size_t cached_ = 0;
};

struct DevBuf {  // scoped device allocation
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    explicit DevBuf(size_t n) : bytes(std::max<size_t>(n, 16)) { p = ScratchCache::get().take(bytes); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) ScratchCache::get().give(p, bytes); }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

void require_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw Error(LEANN_ERR_CUDA, "no CUDA device available: libleann_cuda has no CPU fallback");
    }
    if (device < 0 || device >= n) throw Error(LEANN_ERR_INVALID_ARG, "device ordinal out of range");
}

template <typename T>
T* upload_vec(const std::vector<T>& v) {
    T* p = nullptr;
    LEANN_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) LEANN_CUDA_CHECK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

void bm25_ensure_ws(const leann_cuda_bm25* b, size_t nq) {
    if (!b->stream) LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    if (!b->d_qcounter) LEANN_CUDA_CHECK(cudaMalloc(&b->d_qcounter, 16));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device);
    (void)nq;
    b->n_ctas = sms * bm25_query_ctas_per_sm();   // persistent pool of the query kernel: 3 CTAs of 64 KB shared memory per SM
    if (b->d_acc) return;
    size_t per = std::max<size_t>(b->host.num_docs, 1) * 4;
    LEANN_CUDA_CHECK(cudaMalloc(&b->d_acc, per));
    LEANN_CUDA_CHECK(cudaMemset(b->d_acc, 0, per));
}

// query texts -> CSR of known term ids (unknown terms contribute nothing: bm25.rs:82-85). Batches are tokenised by a few
// host threads (the dictionary is read-only): at 10k queries this was 8 ms of serial host time on the hybrid path.
void tokenize_queries(const leann_cuda_bm25* b, const char* const* texts, const size_t* bytes, size_t nq,
                      std::vector<uint64_t>& off, std::vector<uint32_t>& terms) {
    off.assign(nq + 1, 0);
    terms.clear();
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t T = nq >= 512 ? std::min<size_t>({(size_t)8, (size_t)hw, nq / 256}) : 1;
    std::vector<std::vector<uint32_t>> part(T);
    std::vector<std::vector<uint32_t>> cnt(T);
    std::vector<std::string> errs(T);
    auto work = [&](size_t t) {
        const size_t lo = nq * t / T, hi = nq * (t + 1) / T;
        std::vector<std::string> toks;
        cnt[t].assign(hi - lo, 0);
        for (size_t i = lo; i < hi; ++i) {
            const size_t before = part[t].size();
            if (texts && texts[i]) {
                tokenize(texts[i], bytes ? bytes[i] : strlen(texts[i]), toks);
                for (auto& tk : toks) {
                    auto it = b->host.dict.find(tk);
                    if (it != b->host.dict.end()) part[t].push_back(it->second);
                }
            }
            cnt[t][i - lo] = (uint32_t)(part[t].size() - before);
            if (part[t].size() - before > bm25_max_query_tokens() && errs[t].empty())
                errs[t] = "bm25: query " + std::to_string(i) + " has more than " + std::to_string(bm25_max_query_tokens()) + " indexed tokens";
        }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (size_t t = 0; t < T; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    for (size_t t = 0; t < T; ++t) if (!errs[t].empty()) throw Error(LEANN_ERR_INVALID_ARG, errs[t]);
    size_t total = 0;
    for (size_t t = 0; t < T; ++t) total += part[t].size();
    terms.reserve(total);
    for (size_t t = 0; t < T; ++t) {
        const size_t lo = nq * t / T;
        for (size_t j = 0; j < cnt[t].size(); ++j) off[lo + j + 1] = off[lo + j] + cnt[t][j];
        terms.insert(terms.end(), part[t].begin(), part[t].end());
    }
}

// The device half of search_with_options for a batch. All host pointers.
void hybrid_search_impl(const leann_cuda_index* ix, const leann_cuda_bm25* bm, const float* queries,
                        const char* const* texts, const size_t* text_bytes, size_t nq, size_t top_k, size_t ef,
                        int hybrid, float alpha, const uint64_t* mask, size_t mask_words, uint64_t* out_idx,
                        float* out_score, uint32_t* out_cnt) {
    if (!ix) throw Error(LEANN_ERR_INVALID_ARG, "null index");
    if (nq == 0) return;
    if (top_k == 0) throw Error(LEANN_ERR_INVALID_ARG, "top_k must be > 0");
    const size_t fk = (mask || hybrid) ? top_k * 5 : top_k;  // searcher.rs:129-133 (decided by the flag, not the text)
    if (hybrid && !texts) hybrid = 0;  // searcher.rs:147: hybrid without query_text leaves the vector order
    if (hybrid && !bm) throw Error(LEANN_ERR_INVALID_ARG, "hybrid search needs a BM25 handle");
    if (hybrid && bm->device != ix->device) throw Error(LEANN_ERR_INVALID_ARG, "BM25 handle lives on another device");
    if (fk > 1024) throw Error(LEANN_ERR_INVALID_ARG, "top_k too large for the fused path (5*top_k <= 1024)");
    DevGuard dg(ix->device);
    std::lock_guard<std::mutex> lk(ix->mu);
    cudaStream_t s = backend_stream(ix);
    DevBuf dq(nq * ix->d * 4), vk(nq * fk * 8), vd(nq * fk * 4), vc(nq * 4);
    DevBuf oi(nq * top_k * 8), os(nq * top_k * 4), oc(nq * 4);
    DevBuf dmask(mask ? mask_words * 8 : 16);
    LEANN_CUDA_CHECK(cudaMemcpyAsync(dq.p, queries, nq * ix->d * 4, cudaMemcpyHostToDevice, s));
    if (mask) LEANN_CUDA_CHECK(cudaMemcpyAsync(dmask.p, mask, mask_words * 8, cudaMemcpyHostToDevice, s));
    backend_search_device(ix, dq.as<float>(), nq, fk, ef, nullptr, vk.as<uint64_t>(), vd.as<float>(), vc.as<uint32_t>(), s);
    std::unique_ptr<DevBuf> qo, qt, cb, bi, bs, bc, bx, bn;
    std::unique_ptr<std::lock_guard<std::mutex>> blk;
    if (hybrid) {
        std::vector<uint64_t> off;
        std::vector<uint32_t> terms;
        tokenize_queries(bm, texts, text_bytes, nq, off, terms);
        blk.reset(new std::lock_guard<std::mutex>(bm->mu));
        bm25_ensure_ws(bm, nq);
        qo.reset(new DevBuf(off.size() * 8)); qt.reset(new DevBuf(terms.size() * 4));
        cb.reset(new DevBuf(nq * fk * 4)); bi.reset(new DevBuf(nq * fk * 8)); bs.reset(new DevBuf(nq * fk * 4));
        bc.reset(new DevBuf(nq * 4)); bx.reset(new DevBuf(nq * 4)); bn.reset(new DevBuf(nq * 4));
        // The BM25 top-k kernel runs on the BM25 handle's own stream, concurrently with the vector search on `s`: neither needs
        // the other's output. The scores of the vector candidates are then read straight from the postings on `s`.
        cudaStream_t sb = bm->stream;
        LEANN_CUDA_CHECK(cudaMemcpyAsync(qo->p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, sb));
        if (!terms.empty()) LEANN_CUDA_CHECK(cudaMemcpyAsync(qt->p, terms.data(), terms.size() * 4, cudaMemcpyHostToDevice, sb));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(sb));  // off/terms are stack-owned host vectors; nothing else is queued on sb
        launch_bm25_query(bm->view(), qo->as<uint64_t>(), qt->as<uint32_t>(), (uint32_t)nq, (uint32_t)fk, bm->n_ctas,
                          nullptr, nullptr, (uint32_t)fk, nullptr, bi->as<uint64_t>(), bs->as<float>(),
                          bc->as<uint32_t>(), bx->as<float>(), bn->as<float>(), bm->d_qcounter, sb);
        if (!bm->ev_join) LEANN_CUDA_CHECK(cudaEventCreateWithFlags(&bm->ev_join, cudaEventDisableTiming));
        LEANN_CUDA_CHECK(cudaEventRecord(bm->ev_join, sb));
        launch_bm25_candidates(bm->view(), qo->as<uint64_t>(), qt->as<uint32_t>(), (uint32_t)nq, vk.as<uint64_t>(), vc.as<uint32_t>(),
                               (uint32_t)fk, cb->as<float>(), s);
        LEANN_CUDA_CHECK(cudaStreamWaitEvent(s, bm->ev_join, 0));
    }
    launch_hybrid_fuse(vk.as<uint64_t>(), vd.as<float>(), vc.as<uint32_t>(), (uint32_t)fk, hybrid ? cb->as<float>() : nullptr,
                       hybrid ? bi->as<uint64_t>() : nullptr, hybrid ? bs->as<float>() : nullptr,
                       hybrid ? bc->as<uint32_t>() : nullptr, (uint32_t)(hybrid ? fk : 0), hybrid ? bx->as<float>() : nullptr,
                       hybrid ? bn->as<float>() : nullptr, hybrid, alpha, mask ? dmask.as<uint64_t>() : nullptr,
                       (uint64_t)mask_words * 64, (uint32_t)top_k, oi.as<uint64_t>(), os.as<float>(), oc.as<uint32_t>(),
                       (uint32_t)nq, s);
    LEANN_CUDA_CHECK(cudaMemcpyAsync(out_idx, oi.p, nq * top_k * 8, cudaMemcpyDeviceToHost, s));
    LEANN_CUDA_CHECK(cudaMemcpyAsync(out_score, os.p, nq * top_k * 4, cudaMemcpyDeviceToHost, s));
    if (out_cnt) LEANN_CUDA_CHECK(cudaMemcpyAsync(out_cnt, oc.p, nq * 4, cudaMemcpyDeviceToHost, s));
    LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
}

std::string read_file(const std::string& path, bool& ok) {
    std::ifstream f(path, std::ios::binary);
    ok = (bool)f;
    if (!f) return {};
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
struct leann_cuda_searcher {
    std::string base, backend_name;
    leann_cuda_index* backend = nullptr;
    leann_cuda_bm25* bm25 = nullptr;       // built once, on first hybrid use
    std::vector<std::string> id_map;       // <base>.ids.txt (searcher.rs:83-92)
    std::string jsonl;                     // whole <base>.passages.jsonl
    std::vector<uint64_t> line_off;        // per ordinal: byte offset of its passage line, or ~0 when missing
    std::vector<uint64_t> exists_mask;     // passages.get(id) succeeds
    MetaColumns meta;                      // typed columns of every passage's metadata (built once at load)
    std::unordered_map<std::string, std::vector<uint64_t>> mask_cache;
    bool honor_complexity = false;
    std::mutex mu;
    int device = 0;

    bool passage(size_t ordinal, Json& out) const {
        if (ordinal >= line_off.size() || line_off[ordinal] == ~0ull) return false;
        size_t o = line_off[ordinal];
        size_t e = jsonl.find('\n', o);
        if (e == std::string::npos) e = jsonl.size();
        std::string err;
        return json_parse(jsonl.data() + o, e - o, out, err) && out.kind == Json::Obj;
    }
};

extern "C" {

// ------------------------------------------------------------------ BM25
int leann_cuda_bm25_build(const char* const* docs, const size_t* doc_bytes, size_t n_docs, int device,
                          leann_cuda_bm25** out, char* err, size_t errlen) {
    GUARD({
        if (!out || (n_docs && (!docs || !doc_bytes))) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        require_device(device);
        if (n_docs > 0xFFFFFFF0ull) throw Error(LEANN_ERR_INVALID_ARG, "too many documents");
        std::unique_ptr<leann_cuda_bm25> b(new leann_cuda_bm25());
        b->device = device;
        DevGuard dg(device);
        bm25_build_device(docs, doc_bytes, n_docs, nullptr, b.get(), nullptr, false);
        *out = b.release();
    });
}
// ---- document-range shards: global statistics exchanged as opaque blobs ----
static int copy_blob(const std::string& blob, unsigned char* out, size_t cap, size_t* needed) {
    if (needed) *needed = blob.size();
    if (!out || cap < blob.size()) return out ? LEANN_ERR_INVALID_ARG : LEANN_OK;   // sizing call: out == NULL
    memcpy(out, blob.data(), blob.size());
    return LEANN_OK;
}
int leann_cuda_bm25_shard_stats(const char* const* docs, const size_t* doc_bytes, size_t n_docs, int device, unsigned char* out,
                                size_t cap, size_t* needed, char* err, size_t errlen) {
    GUARD({
        if (n_docs && (!docs || !doc_bytes)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        // the sizing call (out == NULL) runs the device pipeline and keeps the blob for the fill call that follows on this thread
        thread_local std::string cached;
        thread_local const void* cached_key = nullptr;
        thread_local size_t cached_n = 0;
        if (!(out && cached_key == (const void*)docs && cached_n == n_docs && !cached.empty())) {
            require_device(device);
            DevGuard dg(device);
            Bm25GlobalStats st;
            bm25_build_device(docs, doc_bytes, n_docs, nullptr, nullptr, &st, true);
            cached = st.encode();
            cached_key = docs; cached_n = n_docs;
        }
        int rc = copy_blob(cached, out, cap, needed);
        if (out) { cached.clear(); cached_key = nullptr; }
        if (rc != LEANN_OK) throw Error(rc, "stats buffer too small");
    });
}
int leann_cuda_bm25_stats_merge(const unsigned char* const* blobs, const size_t* blob_bytes, size_t n_blobs, unsigned char* out,
                                size_t cap, size_t* needed, char* err, size_t errlen) {
    GUARD({
        if (n_blobs && (!blobs || !blob_bytes)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        Bm25GlobalStats all;
        for (size_t i = 0; i < n_blobs; ++i) {
            Bm25GlobalStats one;
            if (!one.decode(blobs[i], blob_bytes[i])) throw Error(LEANN_ERR_BAD_FORMAT, "BM25 statistics blob " + std::to_string(i) + " is malformed");
            all.merge(one);
        }
        int rc = copy_blob(all.encode(), out, cap, needed);
        if (rc != LEANN_OK) throw Error(rc, "stats buffer too small");
    });
}
int leann_cuda_bm25_build_sharded(const char* const* docs, const size_t* doc_bytes, size_t n_docs, const unsigned char* global_stats,
                                  size_t stats_bytes, int device, leann_cuda_bm25** out, char* err, size_t errlen) {
    GUARD({
        if (!out || !global_stats || (n_docs && (!docs || !doc_bytes))) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        require_device(device);
        Bm25GlobalStats g;
        if (!g.decode(global_stats, stats_bytes)) throw Error(LEANN_ERR_BAD_FORMAT, "BM25 statistics blob is malformed");
        if (g.num_docs < n_docs) throw Error(LEANN_ERR_INVALID_ARG, "global statistics cover fewer documents than this shard holds");
        std::unique_ptr<leann_cuda_bm25> b(new leann_cuda_bm25());
        b->device = device;
        DevGuard dg(device);
        bm25_build_device(docs, doc_bytes, n_docs, &g, b.get(), nullptr, false);
        *out = b.release();
    });
}
size_t leann_cuda_bm25_len(const leann_cuda_bm25* b) { return b ? b->host.num_docs : 0; }
size_t leann_cuda_bm25_dense_rows(const leann_cuda_bm25* b) { return b ? b->n_dense : 0; }
int leann_cuda_bm25_stats(const leann_cuda_bm25* b, uint64_t* st, float* avg) {
    if (!b || !st) return LEANN_ERR_INVALID_ARG;
    st[0] = b->host.num_docs; st[1] = b->host.idf.size(); st[2] = b->host.n_postings; st[3] = b->host.total_tokens;
    if (avg) *avg = b->host.avg_doc_len;
    return LEANN_OK;
}
size_t leann_cuda_tokenize(const char* text, size_t bytes, char* out, size_t cap) {
    std::vector<std::string> toks;
    tokenize(text, bytes, toks);
    size_t w = 0;
    for (size_t i = 0; i < toks.size(); ++i) {
        for (char c : toks[i]) if (out && w + 1 < cap) out[w++] = c;
        if (out && w + 1 < cap) out[w++] = '\n';
    }
    if (out && cap) out[w < cap ? w : cap - 1] = 0;
    return toks.size();
}
void leann_cuda_bm25_free(leann_cuda_bm25* b) {
    if (!b) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(b->device);
    cudaFree(b->d_term_off); cudaFree(b->d_post_doc); cudaFree(b->d_post_score);
    cudaFree(b->d_dense_of); cudaFree(b->d_dense_rows);
    cudaFree(b->d_acc); cudaFree(b->d_qcounter);
    if (b->ev0) { cudaEventDestroy(b->ev0); cudaEventDestroy(b->ev1); }
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->stream) cudaStreamDestroy(b->stream);
    cudaGetLastError();
    if (prev >= 0) cudaSetDevice(prev);
    delete b;
}

int leann_cuda_bm25_score(const leann_cuda_bm25* b, const char* query, size_t query_bytes, float* scores,
                          char* err, size_t errlen) {
    GUARD({
        if (!b || !query || (!scores && b->host.num_docs)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        DevGuard dg(b->device);
        std::lock_guard<std::mutex> lk(b->mu);
        bm25_ensure_ws(b, 1);
        std::vector<std::string> toks;
        tokenize(query, query_bytes, toks);
        std::vector<uint32_t> terms;
        std::vector<uint64_t> dfs;
        for (auto& t : toks) {
            auto it = b->host.dict.find(t);
            if (it == b->host.dict.end()) continue;
            terms.push_back(it->second);
            dfs.push_back(b->host.term_off[it->second + 1] - b->host.term_off[it->second]);
        }
        size_t n = b->host.num_docs;
        if (n == 0) return;
        // uses CTA slot 0 of the accumulator as the dense score vector, then clears it again
        launch_bm25_dense(b->view(), terms.data(), dfs.data(), terms.size(), b->d_acc, b->stream);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(scores, b->d_acc, n * 4, cudaMemcpyDeviceToHost, b->stream));
        LEANN_CUDA_CHECK(cudaMemsetAsync(b->d_acc, 0, n * 4, b->stream));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(b->stream));
    });
}

int leann_cuda_bm25_search(const leann_cuda_bm25* b, const char* const* queries, const size_t* query_bytes, size_t nq,
                           size_t top_k, uint64_t* idx, float* scores, uint32_t* counts, char* err, size_t errlen) {
    GUARD({
        if (!b || (nq && (!queries || !idx || !scores))) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        if (nq == 0) return;
        DevGuard dg(b->device);
        std::vector<uint64_t> off;
        std::vector<uint32_t> terms;
        tokenize_queries(b, queries, query_bytes, nq, off, terms);
        std::lock_guard<std::mutex> lk(b->mu);
        bm25_ensure_ws(b, nq);
        cudaStream_t s = b->stream;
        DevBuf qo(off.size() * 8), qt(terms.size() * 4), bi(nq * top_k * 8), bs(nq * top_k * 4), bc(nq * 4), bx(nq * 4), bn(nq * 4);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(qo.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, s));
        if (!terms.empty()) LEANN_CUDA_CHECK(cudaMemcpyAsync(qt.p, terms.data(), terms.size() * 4, cudaMemcpyHostToDevice, s));
        if (!b->ev0) { LEANN_CUDA_CHECK(cudaEventCreate(&b->ev0)); LEANN_CUDA_CHECK(cudaEventCreate(&b->ev1)); }
        LEANN_CUDA_CHECK(cudaEventRecord(b->ev0, s));
        launch_bm25_query(b->view(), qo.as<uint64_t>(), qt.as<uint32_t>(), (uint32_t)nq, (uint32_t)top_k, b->n_ctas,
                          nullptr, nullptr, 0, nullptr, bi.as<uint64_t>(), bs.as<float>(), bc.as<uint32_t>(), bx.as<float>(),
                          bn.as<float>(), b->d_qcounter, s);
        LEANN_CUDA_CHECK(cudaEventRecord(b->ev1, s));
        b->last_postings = 0;
        b->last_stream_bytes = 0;
        for (uint32_t t : terms) {
            const uint64_t df = b->host.term_off[t + 1] - b->host.term_off[t];
            b->last_postings += df;
            b->last_stream_bytes += (!b->is_dense.empty() && b->is_dense[t]) ? (uint64_t)b->n_pad * 4 : df * 8;
        }
        LEANN_CUDA_CHECK(cudaMemcpyAsync(idx, bi.p, nq * top_k * 8, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaMemcpyAsync(scores, bs.p, nq * top_k * 4, cudaMemcpyDeviceToHost, s));
        if (counts) LEANN_CUDA_CHECK(cudaMemcpyAsync(counts, bc.p, nq * 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
        LEANN_CUDA_CHECK(cudaEventElapsedTime(&b->last_kernel_ms, b->ev0, b->ev1));
    });
}
uint64_t leann_cuda_bm25_last_batch_bytes(const leann_cuda_bm25* b) { return b ? b->last_stream_bytes : 0; }
int leann_cuda_bm25_last_batch(const leann_cuda_bm25* b, uint64_t* postings, float* kernel_ms) {
    if (!b) return LEANN_ERR_INVALID_ARG;
    if (postings) *postings = b->last_postings;
    if (kernel_ms) *kernel_ms = b->last_kernel_ms;
    return LEANN_OK;
}

int leann_cuda_hybrid_rerank(const uint64_t* idx, const float* vec_scores, size_t n, const float* bm25_scores,
                             size_t n_docs, float alpha, int device, uint64_t* out_idx, float* out_scores,
                             char* err, size_t errlen) {
    GUARD({
        if (n && (!idx || !vec_scores || !out_idx || !out_scores)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        if (n == 0) return;
        if (n > 4096) throw Error(LEANN_ERR_INVALID_ARG, "hybrid_rerank: at most 4096 candidates");
        require_device(device);
        DevGuard dg(device);
        cudaStream_t s = nullptr;
        LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        try {
            DevBuf di(n * 8), dv(n * 4), dc(4), dense(n_docs * 4), cb(n * 4), bx(4), bn(4), sc(8), oi(n * 8), os(n * 4), oc(4);
            uint32_t cnt = (uint32_t)n;
            LEANN_CUDA_CHECK(cudaMemcpyAsync(di.p, idx, n * 8, cudaMemcpyHostToDevice, s));
            LEANN_CUDA_CHECK(cudaMemcpyAsync(dv.p, vec_scores, n * 4, cudaMemcpyHostToDevice, s));
            LEANN_CUDA_CHECK(cudaMemcpyAsync(dc.p, &cnt, 4, cudaMemcpyHostToDevice, s));
            if (n_docs) LEANN_CUDA_CHECK(cudaMemcpyAsync(dense.p, bm25_scores, n_docs * 4, cudaMemcpyHostToDevice, s));
            launch_dense_minmax_gather(dense.as<float>(), (uint32_t)n_docs, di.as<uint64_t>(), (uint32_t)n, cb.as<float>(),
                                       bx.as<float>(), bn.as<float>(), sc.as<uint32_t>(), s);
            uint32_t zero = 0;
            DevBuf bc(4);
            LEANN_CUDA_CHECK(cudaMemcpyAsync(bc.p, &zero, 4, cudaMemcpyHostToDevice, s));
            launch_hybrid_fuse(di.as<uint64_t>(), dv.as<float>(), dc.as<uint32_t>(), (uint32_t)n, cb.as<float>(), di.as<uint64_t>(),
                               dv.as<float>(), bc.as<uint32_t>(), 0, bx.as<float>(), bn.as<float>(), 1, alpha, nullptr, 0,
                               (uint32_t)n, oi.as<uint64_t>(), os.as<float>(), oc.as<uint32_t>(), 1, s);
            LEANN_CUDA_CHECK(cudaMemcpyAsync(out_idx, oi.p, n * 8, cudaMemcpyDeviceToHost, s));
            LEANN_CUDA_CHECK(cudaMemcpyAsync(out_scores, os.p, n * 4, cudaMemcpyDeviceToHost, s));
            LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
        } catch (...) {
            cudaStreamDestroy(s);
            throw;
        }
        cudaStreamDestroy(s);
    });
}

// One document-range shard's part of the hybrid step: BM25 top-k of the shard (ids made global with doc_offset),
// the BM25 score of the (already merged, global) vector candidates this shard owns, and the shard's min / max of
// its dense score vector. The caller reduces over shards: top lists -> top-k merge, cand_bm -> sum (only the owner
// is non-zero), bmax -> max, bmin -> min; then leann_cuda_hybrid_fuse.
int leann_cuda_bm25_search_shard(const leann_cuda_bm25* b, const char* const* queries, const size_t* query_bytes, size_t nq,
                                 size_t top_k, uint64_t doc_offset, const uint64_t* cand_idx, const uint32_t* cand_cnt, size_t fk,
                                 uint64_t* top_idx, float* top_score, uint32_t* top_cnt, float* cand_bm, float* bmax, float* bmin,
                                 char* err, size_t errlen) {
    GUARD({
        if (!b || (nq && (!queries || !top_idx || !top_score || !top_cnt || !bmax || !bmin))) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        if (cand_idx && (!cand_cnt || !cand_bm || fk == 0)) throw Error(LEANN_ERR_INVALID_ARG, "candidate list without counts / output");
        if (nq == 0) return;
        DevGuard dg(b->device);
        std::vector<uint64_t> off;
        std::vector<uint32_t> terms;
        tokenize_queries(b, queries, query_bytes, nq, off, terms);
        std::lock_guard<std::mutex> lk(b->mu);
        bm25_ensure_ws(b, nq);
        cudaStream_t s = b->stream;
        const size_t n_local = b->host.num_docs;
        std::vector<uint64_t> local;
        if (cand_idx) {
            local.resize(nq * fk);
            for (size_t i = 0; i < nq * fk; ++i) {
                const uint64_t g = cand_idx[i];
                local[i] = (g >= doc_offset && g - doc_offset < n_local) ? g - doc_offset : ~0ull;
            }
        }
        DevBuf qo(off.size() * 8), qt(terms.size() * 4), bi(nq * top_k * 8), bs(nq * top_k * 4), bc(nq * 4), bx(nq * 4), bn(nq * 4);
        DevBuf ci(cand_idx ? nq * fk * 8 : 16), cc(cand_idx ? nq * 4 : 16), cb(cand_idx ? nq * fk * 4 : 16);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(qo.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, s));
        if (!terms.empty()) LEANN_CUDA_CHECK(cudaMemcpyAsync(qt.p, terms.data(), terms.size() * 4, cudaMemcpyHostToDevice, s));
        if (cand_idx) {
            LEANN_CUDA_CHECK(cudaMemcpyAsync(ci.p, local.data(), nq * fk * 8, cudaMemcpyHostToDevice, s));
            LEANN_CUDA_CHECK(cudaMemcpyAsync(cc.p, cand_cnt, nq * 4, cudaMemcpyHostToDevice, s));
        }
        launch_bm25_query(b->view(), qo.as<uint64_t>(), qt.as<uint32_t>(), (uint32_t)nq, (uint32_t)top_k, b->n_ctas,
                          cand_idx ? ci.as<uint64_t>() : nullptr, cand_idx ? cc.as<uint32_t>() : nullptr, (uint32_t)fk,
                          cand_idx ? cb.as<float>() : nullptr, bi.as<uint64_t>(), bs.as<float>(), bc.as<uint32_t>(), bx.as<float>(),
                          bn.as<float>(), b->d_qcounter, s);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(top_idx, bi.p, nq * top_k * 8, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaMemcpyAsync(top_score, bs.p, nq * top_k * 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaMemcpyAsync(top_cnt, bc.p, nq * 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaMemcpyAsync(bmax, bx.p, nq * 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaMemcpyAsync(bmin, bn.p, nq * 4, cudaMemcpyDeviceToHost, s));
        if (cand_idx) LEANN_CUDA_CHECK(cudaMemcpyAsync(cand_bm, cb.p, nq * fk * 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
        for (size_t i = 0; i < nq * top_k; ++i) if (top_idx[i] != ~0ull) top_idx[i] += doc_offset;
    });
}

// search_with_options (searcher.rs:123-210) over document-range shards, one process per GPU, everything on the device:
//   vector candidates  = the sharded backend (local K1/K2 + ncclAllGather + merge): identical on every rank, global ids;
//   BM25               = this rank's postings (built with the corpus-wide statistics): local top list, scores of the
//                        candidates this shard owns, max / min of its dense score vector;
//   exchange           = ONE ncclAllGather of the packed per-rank block, then top-list merge + sum / max / min reductions;
//   fusion + post-filter walk on every rank (K3f). No host round trip between the steps.
int leann_cuda_shards_hybrid_search(leann_cuda_shards* sh, const leann_cuda_bm25* bm, const float* queries,
                                    const char* const* query_texts, const size_t* query_text_bytes, size_t nq, size_t top_k,
                                    size_t ef, int hybrid, float alpha, const uint64_t* filter_mask, size_t mask_bits,
                                    uint64_t* idx, float* scores, uint32_t* counts, char* err, size_t errlen) {
    GUARD({
        if (!sh || (nq && (!queries || !idx || !scores))) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        if (nq == 0) return;
        if (top_k == 0) throw Error(LEANN_ERR_INVALID_ARG, "top_k must be > 0");
        int world = 1, rank = 0, device = 0;
        uint64_t offsets[16];
        size_t local_shards = 0;
        shards_describe(sh, &world, &rank, &device, offsets, &local_shards);
        if (local_shards != 1) throw Error(LEANN_ERR_INVALID_ARG, "sharded hybrid search needs a handle with one local shard (leann_cuda_shards_join)");
        const size_t fk = (filter_mask || hybrid) ? top_k * 5 : top_k;   // searcher.rs:129-133
        if (hybrid && !query_texts) hybrid = 0;                          // searcher.rs:147
        if (hybrid && !bm) throw Error(LEANN_ERR_INVALID_ARG, "hybrid search needs this rank's BM25 shard");
        if (hybrid && bm->device != device) throw Error(LEANN_ERR_INVALID_ARG, "BM25 shard lives on another device");
        if (fk > 1024) throw Error(LEANN_ERR_INVALID_ARG, "top_k too large for the fused path (5*top_k <= 1024)");
        DevGuard dg(device);
        cudaStream_t s = shards_stream(sh);
        const size_t d = leann_cuda_shards_dims(sh);
        DevBuf dq(nq * d * 4), vk(nq * fk * 8), vd(nq * fk * 4), vc(nq * 4), oi(nq * top_k * 8), os(nq * top_k * 4), oc(nq * 4);
        const size_t mask_words = filter_mask ? (mask_bits + 63) / 64 : 0;
        DevBuf dmask(filter_mask ? mask_words * 8 : 16);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(dq.p, queries, nq * d * 4, cudaMemcpyHostToDevice, s));
        if (filter_mask) LEANN_CUDA_CHECK(cudaMemcpyAsync(dmask.p, filter_mask, mask_words * 8, cudaMemcpyHostToDevice, s));
        shards_vector_search_device(sh, dq.as<float>(), nq, fk, ef, vk.as<uint64_t>(), vd.as<float>(), vc.as<uint32_t>(), s);
        std::unique_ptr<DevBuf> qo, qt, ci, blk, gat, cb, bi, bs, bc, bx, bn;
        std::unique_ptr<std::lock_guard<std::mutex>> blk_lock;
        if (hybrid) {
            std::vector<uint64_t> off;
            std::vector<uint32_t> terms;
            tokenize_queries(bm, query_texts, query_text_bytes, nq, off, terms);   // host work while the vector search runs
            blk_lock.reset(new std::lock_guard<std::mutex>(bm->mu));
            bm25_ensure_ws(bm, nq);
            const size_t n = nq * fk;
            const size_t block_bytes = (n * 16 + nq * 8 + 255) & ~(size_t)255;
            qo.reset(new DevBuf(off.size() * 8)); qt.reset(new DevBuf(terms.size() * 4)); ci.reset(new DevBuf(n * 8));
            blk.reset(new DevBuf(block_bytes)); gat.reset(new DevBuf(block_bytes * (size_t)world));
            cb.reset(new DevBuf(n * 4)); bi.reset(new DevBuf(n * 8)); bs.reset(new DevBuf(n * 4)); bc.reset(new DevBuf(nq * 4));
            bx.reset(new DevBuf(nq * 4)); bn.reset(new DevBuf(nq * 4));
            unsigned char* B = blk->as<unsigned char>();
            uint64_t* b_idx = reinterpret_cast<uint64_t*>(B);
            float* b_score = reinterpret_cast<float*>(B + n * 8);
            float* b_cand = reinterpret_cast<float*>(B + n * 12);
            float* b_max = reinterpret_cast<float*>(B + n * 16);
            float* b_min = reinterpret_cast<float*>(B + n * 16 + nq * 4);
            // BM25 top list of this shard on the BM25 handle's stream, concurrently with the vector search on `s`
            cudaStream_t sb = bm->stream;
            LEANN_CUDA_CHECK(cudaMemcpyAsync(qo->p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, sb));
            if (!terms.empty()) LEANN_CUDA_CHECK(cudaMemcpyAsync(qt->p, terms.data(), terms.size() * 4, cudaMemcpyHostToDevice, sb));
            LEANN_CUDA_CHECK(cudaStreamSynchronize(sb));   // off / terms are stack-owned
            launch_bm25_query(bm->view(), qo->as<uint64_t>(), qt->as<uint32_t>(), (uint32_t)nq, (uint32_t)fk, bm->n_ctas, nullptr, nullptr,
                              (uint32_t)fk, nullptr, b_idx, b_score, bc->as<uint32_t>(), b_max, b_min, bm->d_qcounter, sb);
            if (!bm->ev_join) LEANN_CUDA_CHECK(cudaEventCreateWithFlags(&bm->ev_join, cudaEventDisableTiming));
            LEANN_CUDA_CHECK(cudaEventRecord(bm->ev_join, sb));
            // scores of the (global) vector candidates this shard owns
            launch_localize_candidates(vk.as<uint64_t>(), offsets[rank], bm->host.num_docs, n, ci->as<uint64_t>(), s);
            launch_bm25_candidates(bm->view(), qo->as<uint64_t>(), qt->as<uint32_t>(), (uint32_t)nq, ci->as<uint64_t>(), vc.as<uint32_t>(),
                                   (uint32_t)fk, b_cand, s);
            LEANN_CUDA_CHECK(cudaStreamWaitEvent(s, bm->ev_join, 0));
            shards_all_gather(sh, B, gat->p, block_bytes, s);
            const uint64_t* keys[16];
            const float* dists[16];
            for (int r = 0; r < world; ++r) {
                keys[r] = reinterpret_cast<const uint64_t*>(gat->as<unsigned char>() + (size_t)r * block_bytes);
                dists[r] = reinterpret_cast<const float*>(gat->as<unsigned char>() + (size_t)r * block_bytes + n * 8);
            }
            shards_merge_lists(keys, dists, offsets, (uint32_t)world, (uint32_t)nq, (uint32_t)fk, 1, bi->as<uint64_t>(), bs->as<float>(),
                               bc->as<uint32_t>(), s);
            launch_shard_reduce(gat->as<unsigned char>(), block_bytes, (uint32_t)world, (uint32_t)nq, (uint32_t)fk, cb->as<float>(),
                                bx->as<float>(), bn->as<float>(), s);
        }
        launch_hybrid_fuse(vk.as<uint64_t>(), vd.as<float>(), vc.as<uint32_t>(), (uint32_t)fk, hybrid ? cb->as<float>() : nullptr,
                           hybrid ? bi->as<uint64_t>() : nullptr, hybrid ? bs->as<float>() : nullptr, hybrid ? bc->as<uint32_t>() : nullptr,
                           (uint32_t)(hybrid ? fk : 0), hybrid ? bx->as<float>() : nullptr, hybrid ? bn->as<float>() : nullptr, hybrid, alpha,
                           filter_mask ? dmask.as<uint64_t>() : nullptr, (uint64_t)mask_words * 64, (uint32_t)top_k, oi.as<uint64_t>(),
                           os.as<float>(), oc.as<uint32_t>(), (uint32_t)nq, s);
        LEANN_CUDA_CHECK(cudaMemcpyAsync(idx, oi.p, nq * top_k * 8, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaMemcpyAsync(scores, os.p, nq * top_k * 4, cudaMemcpyDeviceToHost, s));
        if (counts) LEANN_CUDA_CHECK(cudaMemcpyAsync(counts, oc.p, nq * 4, cudaMemcpyDeviceToHost, s));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
    });
}

// Batched hybrid_rerank + post-filter walk (bm25.rs:135-170, searcher.rs:156-207) on already gathered inputs.
int leann_cuda_hybrid_fuse(const uint64_t* vkeys, const float* vdists, const uint32_t* vcnt, size_t nq, size_t fk,
                           const float* cand_bm, const uint64_t* bm_idx, const float* bm_score, const uint32_t* bm_cnt, size_t bm_k,
                           const float* bmax, const float* bmin, int hybrid, float alpha, const uint64_t* mask, size_t mask_bits,
                           size_t top_k, int device, uint64_t* out_idx, float* out_score, uint32_t* out_cnt, char* err, size_t errlen) {
    GUARD({
        if (nq && (!vkeys || !vdists || !vcnt || !out_idx || !out_score || !out_cnt)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        if (hybrid && (!cand_bm || !bm_idx || !bm_score || !bm_cnt || !bmax || !bmin)) throw Error(LEANN_ERR_INVALID_ARG, "hybrid fuse needs the BM25 inputs");
        if (nq == 0) return;
        if (top_k == 0 || fk == 0) throw Error(LEANN_ERR_INVALID_ARG, "top_k and fetch_k must be > 0");
        require_device(device);
        DevGuard dg(device);
        cudaStream_t s = nullptr;
        LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        try {
            const size_t mw = mask ? (mask_bits + 63) / 64 : 0;
            DevBuf vk(nq * fk * 8), vd(nq * fk * 4), vc(nq * 4), cb(hybrid ? nq * fk * 4 : 16), bi(hybrid ? nq * bm_k * 8 : 16),
                bs(hybrid ? nq * bm_k * 4 : 16), bc(hybrid ? nq * 4 : 16), bx(hybrid ? nq * 4 : 16), bn(hybrid ? nq * 4 : 16),
                dm(mask ? mw * 8 : 16), oi(nq * top_k * 8), os(nq * top_k * 4), oc(nq * 4);
            auto up = [&](DevBuf& d, const void* h, size_t bytes) { LEANN_CUDA_CHECK(cudaMemcpyAsync(d.p, h, bytes, cudaMemcpyHostToDevice, s)); };
            up(vk, vkeys, nq * fk * 8); up(vd, vdists, nq * fk * 4); up(vc, vcnt, nq * 4);
            if (hybrid) {
                up(cb, cand_bm, nq * fk * 4); up(bi, bm_idx, nq * bm_k * 8); up(bs, bm_score, nq * bm_k * 4);
                up(bc, bm_cnt, nq * 4); up(bx, bmax, nq * 4); up(bn, bmin, nq * 4);
            }
            if (mask) up(dm, mask, mw * 8);
            launch_hybrid_fuse(vk.as<uint64_t>(), vd.as<float>(), vc.as<uint32_t>(), (uint32_t)fk, hybrid ? cb.as<float>() : nullptr,
                               hybrid ? bi.as<uint64_t>() : nullptr, hybrid ? bs.as<float>() : nullptr, hybrid ? bc.as<uint32_t>() : nullptr,
                               (uint32_t)(hybrid ? bm_k : 0), hybrid ? bx.as<float>() : nullptr, hybrid ? bn.as<float>() : nullptr, hybrid,
                               alpha, mask ? dm.as<uint64_t>() : nullptr, (uint64_t)mask_bits, (uint32_t)top_k, oi.as<uint64_t>(),
                               os.as<float>(), oc.as<uint32_t>(), (uint32_t)nq, s);
            LEANN_CUDA_CHECK(cudaMemcpyAsync(out_idx, oi.p, nq * top_k * 8, cudaMemcpyDeviceToHost, s));
            LEANN_CUDA_CHECK(cudaMemcpyAsync(out_score, os.p, nq * top_k * 4, cudaMemcpyDeviceToHost, s));
            LEANN_CUDA_CHECK(cudaMemcpyAsync(out_cnt, oc.p, nq * 4, cudaMemcpyDeviceToHost, s));
            LEANN_CUDA_CHECK(cudaStreamSynchronize(s));
        } catch (...) {
            cudaStreamDestroy(s);
            throw;
        }
        cudaStreamDestroy(s);
    });
}

int leann_cuda_hybrid_search(const leann_cuda_index* index, const leann_cuda_bm25* bm25, const float* queries,
                             const char* const* query_texts, const size_t* query_text_bytes, size_t nq, size_t top_k,
                             size_t ef, int hybrid, float alpha, const uint64_t* filter_mask, uint64_t* idx,
                             float* scores, uint32_t* counts, char* err, size_t errlen) {
    GUARD({
        if (!index) throw Error(LEANN_ERR_INVALID_ARG, "null index");
        hybrid_search_impl(index, bm25, queries, query_texts, query_text_bytes, nq, top_k, ef, hybrid, alpha, filter_mask,
                           filter_mask ? (index->n + 63) / 64 : 0, idx, scores, counts);
    });
}

// ------------------------------------------------------------------ filter
int leann_cuda_filter_parse(const char* expr, leann_cuda_filter** out, char* err, size_t errlen) {
    GUARD({
        if (!expr || !out) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        std::unique_ptr<leann_cuda_filter> f(new leann_cuda_filter());
        f->expr = expr;
        if (!filter_parse(f->expr, f->root)) throw Error(LEANN_ERR_PARSE, std::string("cannot parse filter: ") + expr);
        *out = f.release();
    });
}
size_t leann_cuda_filter_describe(const leann_cuda_filter* f, char* out, size_t cap) {
    if (!f) return 0;
    std::string d = filter_describe(f->root);
    if (out && cap) { size_t n = std::min(cap - 1, d.size()); memcpy(out, d.data(), n); out[n] = 0; }
    return d.size();
}
int leann_cuda_filter_matches(const leann_cuda_filter* f, const char* metadata_json, size_t bytes, int* result,
                              char* err, size_t errlen) {
    GUARD({
        if (!f || !metadata_json || !result) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        Json md;
        std::string e;
        if (!json_parse(metadata_json, bytes, md, e)) throw Error(LEANN_ERR_BAD_FORMAT, "metadata is not valid JSON: " + e);
        *result = filter_matches(f->root, md) ? 1 : 0;
    });
}
int leann_cuda_filter_mask(const leann_cuda_filter* f, const char* const* metadata_json, const size_t* bytes, size_t n,
                           uint64_t* mask_bits, char* err, size_t errlen) {
    GUARD({
        if (!f || (n && (!metadata_json || !bytes || !mask_bits))) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        size_t words = (n + 63) / 64;
        for (size_t w = 0; w < words; ++w) mask_bits[w] = 0;
        Json md;
        std::string e;
        for (size_t i = 0; i < n; ++i) {
            if (!json_parse(metadata_json[i], bytes[i], md, e)) throw Error(LEANN_ERR_BAD_FORMAT, "metadata " + std::to_string(i) + " is not valid JSON: " + e);
            if (filter_matches(f->root, md)) mask_bits[i >> 6] |= 1ull << (i & 63);
        }
    });
}
void leann_cuda_filter_free(leann_cuda_filter* f) { delete f; }

// ------------------------------------------------------------------ metadata columns (host only)
int leann_cuda_metacols_build(const char* const* metadata_json, const size_t* bytes, size_t n, leann_cuda_metacols** out,
                              char* err, size_t errlen) {
    GUARD({
        if (!out || (n && (!metadata_json || !bytes))) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        std::unique_ptr<leann_cuda_metacols> mc(new leann_cuda_metacols());
        mc->cols.resize(n);
        Json md;
        std::string e;
        for (size_t i = 0; i < n; ++i) {
            if (!metadata_json[i]) continue;   // passage without metadata: every field missing
            if (!json_parse(metadata_json[i], bytes[i], md, e)) throw Error(LEANN_ERR_BAD_FORMAT, "metadata " + std::to_string(i) + " is not valid JSON: " + e);
            mc->cols.add_row(i, md);
        }
        *out = mc.release();
    });
}
int leann_cuda_metacols_mask(const leann_cuda_metacols* mc, const leann_cuda_filter* f, uint64_t* mask_bits, char* err, size_t errlen) {
    GUARD({
        if (!mc || !f || (mc->cols.n && !mask_bits)) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        std::vector<uint64_t> m;
        mc->cols.eval(f->root, m);
        for (size_t w = 0; w < m.size(); ++w) mask_bits[w] = m[w];
    });
}
size_t leann_cuda_metacols_len(const leann_cuda_metacols* mc) { return mc ? mc->cols.n : 0; }
size_t leann_cuda_metacols_fields(const leann_cuda_metacols* mc) { return mc ? mc->cols.cols.size() : 0; }
void leann_cuda_metacols_free(leann_cuda_metacols* mc) { delete mc; }

// ------------------------------------------------------------------ IndexSearcher
int leann_cuda_searcher_load(const char* base_path, const char* backend_name, size_t dims, int device,
                             leann_cuda_searcher** out, char* err, size_t errlen) {
    GUARD({
        if (!base_path || !backend_name || !out) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        *out = nullptr;
        std::unique_ptr<leann_cuda_searcher> s(new leann_cuda_searcher());
        s->base = base_path; s->backend_name = backend_name; s->device = device;
        int backend;
        if (s->backend_name == "hnsw") backend = LEANN_BACKEND_HNSW;
        else if (s->backend_name == "diskann") backend = LEANN_BACKEND_VAMANA;
        else if (s->backend_name == "flat") backend = LEANN_BACKEND_FLAT;
        else throw Error(LEANN_ERR_INVALID_ARG, "Unknown backend: " + s->backend_name);  // searcher.rs:98
        // PassageStore::open (passages.rs:47-59)
        bool ok = false;
        std::string idx_txt = read_file(with_extension(s->base, "passages.idx.json"), ok);
        if (!ok) throw Error(LEANN_ERR_NOT_FOUND, "cannot read " + with_extension(s->base, "passages.idx.json"));
        Json offsets;
        std::string e;
        if (!json_parse(idx_txt.data(), idx_txt.size(), offsets, e) || offsets.kind != Json::Obj)
            throw Error(LEANN_ERR_BAD_FORMAT, "passages.idx.json is not a JSON object: " + e);
        s->jsonl = read_file(with_extension(s->base, "passages.jsonl"), ok);
        if (!ok) throw Error(LEANN_ERR_NOT_FOUND, "cannot read " + with_extension(s->base, "passages.jsonl"));
        std::unordered_map<std::string, uint64_t> off;
        off.reserve(offsets.obj.size() * 2);
        for (auto& kv : offsets.obj) if (kv.second.kind == Json::Num) off[kv.first] = (uint64_t)kv.second.num;
        // ids.txt -> id_map (searcher.rs:83-92). The HashMap-order fallback of the reference is
        // nondeterministic; without ids.txt the JSONL order is used.
        std::string ids = read_file(with_extension(s->base, "ids.txt"), ok);
        if (ok) {
            size_t p = 0;
            while (p < ids.size()) {  // str::lines(): split on \n, strip one trailing \r, no empty tail
                size_t nl = ids.find('\n', p);
                std::string line = ids.substr(p, nl == std::string::npos ? std::string::npos : nl - p);
                if (!line.empty() && line.back() == '\r') line.pop_back();
                s->id_map.push_back(line);
                if (nl == std::string::npos) break;
                p = nl + 1;
            }
        } else {
            std::vector<std::pair<uint64_t, std::string>> byoff;
            for (auto& kv : off) byoff.emplace_back(kv.second, kv.first);
            std::sort(byoff.begin(), byoff.end());
            for (auto& kv : byoff) s->id_map.push_back(kv.second);
        }
        s->line_off.assign(s->id_map.size(), ~0ull);
        s->exists_mask.assign((s->id_map.size() + 63) / 64, 0);
        for (size_t i = 0; i < s->id_map.size(); ++i) {
            auto it = off.find(s->id_map[i]);
            if (it == off.end() || it->second >= s->jsonl.size()) continue;
            s->line_off[i] = it->second;
            Json p;
            if (s->passage(i, p)) {
                s->exists_mask[i >> 6] |= 1ull << (i & 63);
                if (const Json* md = p.get("metadata")) s->meta.add_row(i, *md);
            } else s->line_off[i] = ~0ull;
        }
        s->meta.resize(s->id_map.size());
        int rc = leann_cuda_open(base_path, backend, dims, LEANN_METRIC_DEFAULT, device, &s->backend, err, errlen);
        if (rc != LEANN_OK) throw Error(rc, err ? std::string(err) : std::string("backend open failed"));
        *out = s.release();
    });
}
size_t leann_cuda_searcher_len(const leann_cuda_searcher* s) { return s && s->backend ? s->backend->n : 0; }
size_t leann_cuda_searcher_id(const leann_cuda_searcher* s, uint64_t idx, char* out, size_t cap) {
    if (!s) return 0;
    std::string id = idx < s->id_map.size() ? s->id_map[idx] : std::to_string(idx);  // searcher.rs:180-184
    if (out && cap) { size_t n = std::min(cap - 1, id.size()); memcpy(out, id.data(), n); out[n] = 0; }
    return id.size();
}
void leann_cuda_searcher_close(leann_cuda_searcher* s) {
    if (!s) return;
    if (s->bm25) leann_cuda_bm25_free(s->bm25);
    if (s->backend) leann_cuda_close(s->backend);
    delete s;
}

int leann_cuda_searcher_search(const leann_cuda_searcher* cs, const float* queries, const char* const* query_texts,
                               const size_t* query_text_bytes, size_t nq, size_t top_k, size_t complexity,
                               const char* filter_expr, int hybrid, float alpha, uint64_t* idx, float* scores,
                               uint32_t* counts, char* err, size_t errlen) {
    GUARD({
        leann_cuda_searcher* s = const_cast<leann_cuda_searcher*>(cs);
        if (!s || !s->backend) throw Error(LEANN_ERR_INVALID_ARG, "null searcher");
        const size_t n_ord = std::max(s->id_map.size(), s->backend->n);
        const size_t words = (n_ord + 63) / 64;
        // the walk of searcher.rs:174-207 drops ordinals whose passage cannot be loaded, then applies the filter
        std::vector<uint64_t> mask(words, 0);
        const std::vector<uint64_t>* use = nullptr;
        bool have_filter = filter_expr && filter_expr[0];
        {
            std::lock_guard<std::mutex> lk(s->mu);
            if (have_filter) {
                auto it = s->mask_cache.find(filter_expr);
                if (it == s->mask_cache.end()) {
                    FilterNode root;
                    if (!filter_parse(filter_expr, root)) throw Error(LEANN_ERR_PARSE, std::string("cannot parse filter: ") + filter_expr);
                    // column-wise over the side-car (rows whose passage is missing are cleared by exists_mask below)
                    std::vector<uint64_t> m;
                    s->meta.eval(root, m);
                    m.resize(words, 0);
                    it = s->mask_cache.emplace(filter_expr, std::move(m)).first;
                }
                use = &it->second;
            }
            if (hybrid && query_texts && !s->bm25) {
                // get_all_texts (searcher.rs:213-224): missing passages contribute an empty document
                std::vector<std::string> texts(s->id_map.size());
                Json p;
                for (size_t i = 0; i < s->id_map.size(); ++i)
                    if (s->passage(i, p)) { const Json* t = p.get("text"); if (t && t->kind == Json::Str) texts[i] = t->str; }
                std::vector<const char*> ptrs(texts.size());
                std::vector<size_t> lens(texts.size());
                for (size_t i = 0; i < texts.size(); ++i) { ptrs[i] = texts[i].data(); lens[i] = texts[i].size(); }
                int rc = leann_cuda_bm25_build(ptrs.data(), lens.data(), texts.size(), s->device, &s->bm25, err, errlen);
                if (rc != LEANN_OK) throw Error(rc, err ? std::string(err) : std::string("bm25 build failed"));
            }
        }
        // mask = exists AND filter; it is passed whenever a filter is set or some passage is missing
        bool all_exist = true;
        for (size_t i = 0; i < s->id_map.size() && all_exist; ++i) all_exist = (s->exists_mask[i >> 6] >> (i & 63)) & 1ull;
        if (s->backend->n > s->id_map.size()) all_exist = false;  // ordinals beyond ids.txt have no passage
        const uint64_t* mask_ptr = nullptr;
        if (have_filter || !all_exist) {
            for (size_t w = 0; w < s->exists_mask.size(); ++w) mask[w] = s->exists_mask[w] & (use ? (*use)[w] : ~0ull);
            mask_ptr = mask.data();
        }
        // fetch_k is 5k only when a filter or hybrid was REQUESTED (searcher.rs:129); a mask that exists
        // only because passages are missing must not widen it.
        size_t ef = complexity;
        if (s->backend->backend == LEANN_BACKEND_HNSW && !s->honor_complexity) ef = 64;  // hnsw.rs:49,83
        if (!have_filter && !hybrid && mask_ptr) {
            // plain search with holes: fetch top_k, then drop missing ones (may return < k, as the reference does)
            std::vector<uint64_t> k0(nq * top_k);
            std::vector<float> d0(nq * top_k);
            std::vector<uint32_t> c0(nq);
            int rc = leann_cuda_search(s->backend, queries, nq, top_k, ef, nullptr, LEANN_MASK_NONE, k0.data(), d0.data(), c0.data(), err, errlen);
            if (rc != LEANN_OK) throw Error(rc, err ? std::string(err) : std::string("search failed"));
            for (size_t q = 0; q < nq; ++q) {
                uint32_t o = 0;
                for (uint32_t j = 0; j < c0[q]; ++j) {
                    uint64_t d = k0[q * top_k + j];
                    if (d < words * 64 && ((mask[d >> 6] >> (d & 63)) & 1ull)) { idx[q * top_k + o] = d; scores[q * top_k + o] = d0[q * top_k + j]; ++o; }
                }
                if (counts) counts[q] = o;
                for (; o < top_k; ++o) { idx[q * top_k + o] = ~0ull; scores[q * top_k + o] = 0.0f; }
            }
            return;
        }
        hybrid_search_impl(s->backend, s->bm25, queries, (hybrid ? query_texts : nullptr), query_text_bytes, nq, top_k, ef,
                           hybrid, alpha, mask_ptr, mask_ptr ? words : 0, idx, scores, counts);
        (void)n_ord;
    });
}

// IndexSearcher::bm25_search (searcher.rs:228-246): BM25-only top-k passage ordinals (the caller maps
// them to texts through the passage store, as the reference does).
int leann_cuda_searcher_bm25_search(const leann_cuda_searcher* cs, const char* query, size_t query_bytes, size_t top_k,
                                    uint64_t* idx, float* scores, uint32_t* count, char* err, size_t errlen) {
    GUARD({
        leann_cuda_searcher* s = const_cast<leann_cuda_searcher*>(cs);
        if (!s || !query || !idx || !scores || !count) throw Error(LEANN_ERR_INVALID_ARG, "null argument");
        {
            std::lock_guard<std::mutex> lk(s->mu);
            if (!s->bm25) {
                std::vector<std::string> texts(s->id_map.size());
                Json p;
                for (size_t i = 0; i < s->id_map.size(); ++i)
                    if (s->passage(i, p)) { const Json* t = p.get("text"); if (t && t->kind == Json::Str) texts[i] = t->str; }
                std::vector<const char*> ptrs(texts.size());
                std::vector<size_t> lens(texts.size());
                for (size_t i = 0; i < texts.size(); ++i) { ptrs[i] = texts[i].data(); lens[i] = texts[i].size(); }
                int rc = leann_cuda_bm25_build(ptrs.data(), lens.data(), texts.size(), s->device, &s->bm25, err, errlen);
                if (rc != LEANN_OK) throw Error(rc, err ? std::string(err) : std::string("bm25 build failed"));
            }
        }
        const char* qs[1] = {query};
        size_t ql[1] = {query_bytes};
        int rc = leann_cuda_bm25_search(s->bm25, qs, ql, 1, top_k, idx, scores, count, err, errlen);
        if (rc != LEANN_OK) throw Error(rc, err ? std::string(err) : std::string("bm25 search failed"));
    });
}

int leann_cuda_searcher_set_honor_complexity(leann_cuda_searcher* s, int on) {
    if (!s) return LEANN_ERR_INVALID_ARG;
    s->honor_complexity = on != 0;
    return LEANN_OK;
}

}  // extern "C"
