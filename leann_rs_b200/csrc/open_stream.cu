// open_stream.cu — the load half of leann_cuda_open: replaces `index.load(path)` of HnswSearcher::load
// (leann-rs src/backend/hnsw.rs:53-55; usearch reads the whole file into RAM) and the mmap of DiskAnnSearcher::load
// (src/backend/diskann.rs:34-37) by a file -> HBM pipeline:
//   * the vectors block (3.07 GB of the 3.35 GB 1M x 768 file) never exists in pageable host memory: a few host threads
//     pread() slices of it into two pinned staging buffers while the previous buffer's cudaMemcpyAsync is in flight;
//   * the usearch node block (variable-length records) is parsed by a pool of threads into the fixed-stride adjacency the
//     kernels read, concurrently with the vector stream;
//   * `<base>.cuda-layout` (optional): the parsed adjacency, written once and bound to the `.index` file by
//     (size, mtime, hash of its headers); when valid it is streamed like the vectors and the node block is not read at all.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <memory>
#include <thread>

#include "internal.h"

namespace leann {

namespace {

int io_threads() {
    static int n = 0;
    if (!n) {
        const char* e = getenv("LEANN_CUDA_IO_THREADS");
        n = e ? atoi(e) : (int)std::min<unsigned>(8u, std::max<unsigned>(1u, std::thread::hardware_concurrency()));
        if (n < 1) n = 1;
    }
    return n;
}

template <typename F>
void parallel_ranges(size_t total, int threads, F f) {   // f(lo, hi) on disjoint ranges; first exception is rethrown
    threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, total ? total : 1));
    if (threads == 1) { f((size_t)0, total); return; }
    std::vector<std::thread> pool;
    std::exception_ptr err;
    std::mutex m;
    const size_t per = (total + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const size_t lo = std::min(total, (size_t)t * per), hi = std::min(total, lo + per);
        if (lo == hi) break;
        pool.emplace_back([&, lo, hi] {
            try { f(lo, hi); } catch (...) { std::lock_guard<std::mutex> lk(m); if (!err) err = std::current_exception(); }
        });
    }
    for (auto& t : pool) t.join();
    if (err) std::rethrow_exception(err);
}

struct File {
    int fd = -1;
    std::string path;
    explicit File(const std::string& p) : fd(open(p.c_str(), O_RDONLY)), path(p) {
        if (fd < 0) throw Error(LEANN_ERR_NOT_FOUND, "cannot open " + p + ": " + strerror(errno));
    }
    ~File() { if (fd >= 0) close(fd); }
    void read_at(void* dst, size_t off, size_t bytes, const char* what) const {
        unsigned char* p = (unsigned char*)dst;
        while (bytes) {
            ssize_t r = pread(fd, p, bytes, (off_t)off);
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) throw Error(LEANN_ERR_BAD_FORMAT, path + ": truncated while reading " + what);
            p += r; off += (size_t)r; bytes -= (size_t)r;
        }
    }
    void read_parallel(void* dst, size_t off, size_t bytes, const char* what) const {
        parallel_ranges(bytes, bytes >= ((size_t)8 << 20) ? io_threads() : 1,
                        [&](size_t lo, size_t hi) { read_at((unsigned char*)dst + lo, off + lo, hi - lo, what); });
    }
};

// Two pinned staging buffers; chunk c is read from the file while chunk c - 1 travels to the device.
struct Staging {
    static constexpr size_t CHUNK = (size_t)32 << 20;
    unsigned char* buf[2] = {nullptr, nullptr};
    cudaEvent_t free_ev[2] = {nullptr, nullptr};
    int next = 0;
    Staging() {
        for (int i = 0; i < 2; ++i) {
            LEANN_CUDA_CHECK(cudaMallocHost(&buf[i], CHUNK));
            LEANN_CUDA_CHECK(cudaEventCreateWithFlags(&free_ev[i], cudaEventDisableTiming));
        }
    }
    ~Staging() {
        for (int i = 0; i < 2; ++i) {
            if (free_ev[i]) { cudaEventSynchronize(free_ev[i]); cudaEventDestroy(free_ev[i]); }
            if (buf[i]) cudaFreeHost(buf[i]);
        }
    }
    // file[off, off + bytes) -> device `dst` (same layout), on `stream`
    void copy(const File& f, size_t off, size_t bytes, void* dst, cudaStream_t stream, const char* what) {
        for (size_t pos = 0; pos < bytes; pos += CHUNK) {
            const size_t len = std::min(CHUNK, bytes - pos);
            const int b = next;
            next ^= 1;
            LEANN_CUDA_CHECK(cudaEventSynchronize(free_ev[b]));
            f.read_parallel(buf[b], off + pos, len, what);
            LEANN_CUDA_CHECK(cudaMemcpyAsync((unsigned char*)dst + pos, buf[b], len, cudaMemcpyHostToDevice, stream));
            LEANN_CUDA_CHECK(cudaEventRecord(free_ev[b], stream));
        }
    }
    // rows of `d` floats in the file -> rows padded to d4 float4 on the device (d % 4 != 0)
    void copy_rows_padded(const File& f, size_t off, size_t n, uint32_t d, uint32_t d4, float4* dst, cudaStream_t stream, const char* what) {
        const size_t rows_per = std::max<size_t>(1, CHUNK / ((size_t)d * 4));
        float* tmp = nullptr;
        LEANN_CUDA_CHECK(cudaMalloc(&tmp, rows_per * d * 4));
        try {
            for (size_t r0 = 0; r0 < n; r0 += rows_per) {
                const size_t m = std::min(rows_per, n - r0);
                copy(f, off + r0 * d * 4, m * d * 4, tmp, stream, what);   // one chunk: tmp is reused in stream order
                launch_pad_rows(tmp, dst + r0 * d4, m, d, d4, stream);
            }
            LEANN_CUDA_CHECK(cudaStreamSynchronize(stream));
        } catch (...) {
            cudaStreamSynchronize(stream);
            cudaFree(tmp);
            throw;
        }
        cudaFree(tmp);
    }
    void vectors(const File& f, size_t off, size_t n, size_t d, uint32_t d4, float4* dst, cudaStream_t stream) {
        if (n == 0) return;
        if (d % 4 == 0) copy(f, off, n * d * 4, dst, stream, "vectors");
        else copy_rows_padded(f, off, n, (uint32_t)d, d4, dst, stream, "vectors");
    }
};

template <typename T>
T* dmalloc(size_t count) {
    T* p = nullptr;
    LEANN_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    return p;
}

// ---- <base>.cuda-layout ----------------------------------------------------------------------------------------
struct LayoutHead {   // 128 bytes
    char magic[8];            // "LEANNCL1"
    uint64_t index_size, index_mtime_ns, index_head_hash;
    uint64_t n, d, M, M0, max_level, entry, n_upper, identity_keys;
    uint64_t reserved[4];
};
static_assert(sizeof(LayoutHead) == 128, "layout head");

size_t layout_bytes(const LayoutHead& h) {
    return sizeof(LayoutHead) + h.n * 2 + h.n * 4 + h.n * h.M0 * 4 + h.n_upper * h.M * 4 + h.n * 8;
}

bool layout_valid(const std::string& path, const UsearchPlan& pl, LayoutHead& h) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return false;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    const bool ok = fread(&h, 1, sizeof h, f) == sizeof h;
    fclose(f);
    if (!ok || memcmp(h.magic, "LEANNCL1", 8) != 0) return false;
    if (h.index_size != pl.file_size || (int64_t)h.index_mtime_ns != pl.mtime_ns || h.index_head_hash != pl.head_hash) return false;
    if (h.n != pl.n || h.d != pl.d || h.M != pl.M || h.M0 != pl.M0 || (int64_t)h.max_level != pl.max_level || h.entry != pl.entry) return false;
    return (size_t)st.st_size == layout_bytes(h);
}

}  // namespace

void write_layout_cache(const leann_cuda_index* ix, const std::string& base) {
    if (ix->backend != LEANN_BACKEND_HNSW) throw Error(LEANN_ERR_INVALID_ARG, "the device-layout cache is defined for the HNSW backend (the .diskann and .embeddings layouts are already flat)");
    const std::string index_file = with_extension(base, "index");
    const UsearchPlan pl = usearch_probe(index_file, ix->d);
    if (pl.n != ix->n || pl.M != ix->M || pl.M0 != ix->M0 || (uint64_t)pl.max_level != (uint64_t)ix->max_level || pl.entry != ix->entry)
        throw Error(LEANN_ERR_INVALID_ARG, "the .index file at this base path is not the one this handle was loaded from / saved to");
    LayoutHead h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "LEANNCL1", 8);
    h.index_size = pl.file_size; h.index_mtime_ns = (uint64_t)pl.mtime_ns; h.index_head_hash = pl.head_hash;
    h.n = ix->n; h.d = ix->d; h.M = ix->M; h.M0 = ix->M0; h.max_level = (uint64_t)ix->max_level; h.entry = ix->entry;
    h.n_upper = ix->n_upper_lists; h.identity_keys = ix->identity_keys ? 1 : 0;
    std::vector<uint32_t> upper(ix->n), adj0(ix->n * ix->M0), adjU(ix->n_upper_lists * ix->M);
    std::vector<uint64_t> keys(ix->n);
    auto down = [](auto& v, const void* src) { if (!v.empty()) LEANN_CUDA_CHECK(cudaMemcpy(v.data(), src, v.size() * sizeof(v[0]), cudaMemcpyDeviceToHost)); };
    down(upper, ix->upper_base); down(adj0, ix->adj0); down(adjU, ix->adjU); down(keys, ix->keys);
    const std::string path = with_extension(base, "cuda-layout"), tmp = path + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) throw Error(LEANN_ERR_NOT_FOUND, "cannot create " + tmp + ": " + strerror(errno));
    auto put = [&](const void* p, size_t bytes) { return bytes == 0 || fwrite(p, 1, bytes, f) == bytes; };
    bool ok = put(&h, sizeof h) && put(ix->h_levels.data(), ix->n * 2) && put(upper.data(), upper.size() * 4) &&
              put(adj0.data(), adj0.size() * 4) && put(adjU.data(), adjU.size() * 4) && put(keys.data(), keys.size() * 8);
    ok = (fflush(f) == 0) && ok;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) { remove(tmp.c_str()); throw Error(LEANN_ERR_BAD_FORMAT, "write failed: " + path); }
}

// HnswSearcher::load (hnsw.rs:18-75) after the FAISS sniff: headers verified, blocks streamed. `used_cache` reports
// whether <base>.cuda-layout supplied the adjacency.
leann_cuda_index* open_hnsw_streamed(const std::string& base, size_t dims, int device, int metric, bool* used_cache) {
    const std::string file = with_extension(base, "index");
    const UsearchPlan pl = usearch_probe(file, dims);
    std::unique_ptr<leann_cuda_index> ix(new leann_cuda_index());
    ix->backend = LEANN_BACKEND_HNSW; ix->device = device;
    ix->metric = metric == LEANN_METRIC_DEFAULT ? pl.metric : metric;
    ix->n = pl.n; ix->d = pl.d; ix->d4 = (uint32_t)((pl.d + 3) / 4);
    ix->M = (uint32_t)pl.M; ix->M0 = (uint32_t)pl.M0; ix->max_level = (int)pl.max_level; ix->entry = (uint32_t)pl.entry;
    File f(file);
    cudaStream_t stream;
    LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } sg{stream};
    try {
        ix->vecs = dmalloc<float4>(pl.n * ix->d4);
        ix->adj0 = dmalloc<uint32_t>(pl.n * pl.M0);
        ix->upper_base = dmalloc<uint32_t>(pl.n);
        ix->keys = dmalloc<uint64_t>(pl.n);
        Staging st;
        LayoutHead lh;
        const std::string cache = with_extension(base, "cuda-layout");
        const bool cached = getenv("LEANN_CUDA_NO_LAYOUT_CACHE") == nullptr && layout_valid(cache, pl, lh);
        if (used_cache) *used_cache = cached;
        if (cached) {
            // adjacency straight from the cache file; the node block of the .index is not read
            File cf(cache);
            ix->h_levels.resize(pl.n);
            size_t off = sizeof(LayoutHead);
            cf.read_parallel(ix->h_levels.data(), off, pl.n * 2, "layout levels"); off += pl.n * 2;
            ix->n_upper_lists = lh.n_upper;
            ix->adjU = dmalloc<uint32_t>(lh.n_upper * pl.M);
            st.copy(cf, off, pl.n * 4, ix->upper_base, stream, "layout upper_base"); off += pl.n * 4;
            st.copy(cf, off, pl.n * pl.M0 * 4, ix->adj0, stream, "layout adj0"); off += pl.n * pl.M0 * 4;
            st.copy(cf, off, lh.n_upper * pl.M * 4, ix->adjU, stream, "layout adjU"); off += lh.n_upper * pl.M * 4;
            ix->identity_keys = lh.identity_keys != 0;
            st.copy(cf, off, pl.n * 8, ix->keys, stream, "layout keys");
            st.vectors(f, pl.vec_off, pl.n, pl.d, ix->d4, ix->vecs, stream);
            LEANN_CUDA_CHECK(cudaStreamSynchronize(stream));
        } else {
            // node block -> fixed-stride adjacency on a worker thread, while this thread streams the vectors
            std::vector<uint32_t> upper, adj0, adjU;
            std::vector<uint64_t> keys;
            std::exception_ptr perr;
            std::thread parser([&] {
                try {
                    ix->h_levels.resize(pl.n);
                    f.read_parallel(ix->h_levels.data(), pl.levels_off, pl.n * 2, "levels");
                    std::vector<uint64_t> node_off;
                    size_t n_upper = 0;
                    usearch_layout(pl, file, ix->h_levels.data(), upper, node_off, n_upper);
                    ix->n_upper_lists = n_upper;
                    std::unique_ptr<unsigned char[]> nodes(new unsigned char[std::max<size_t>(node_off[pl.n], 1)]);
                    f.read_parallel(nodes.get(), pl.nodes_off, node_off[pl.n], "node links");
                    keys.resize(pl.n);
                    adj0.resize(pl.n * pl.M0);
                    adjU.resize(n_upper * pl.M);
                    parallel_ranges(pl.n, io_threads(), [&](size_t lo, size_t hi) {
                        usearch_parse_nodes(pl, file, nodes.get(), node_off.data(), ix->h_levels.data(), upper.data(), lo, hi, keys.data(),
                                            adj0.data(), adjU.data());
                    });
                } catch (...) { perr = std::current_exception(); }
            });
            std::exception_ptr verr;
            try { st.vectors(f, pl.vec_off, pl.n, pl.d, ix->d4, ix->vecs, stream); } catch (...) { verr = std::current_exception(); }
            parser.join();
            if (perr) std::rethrow_exception(perr);
            if (verr) std::rethrow_exception(verr);
            ix->adjU = dmalloc<uint32_t>(adjU.size());
            auto up = [&](void* dst, const void* src, size_t bytes) { if (bytes) LEANN_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream)); };
            up(ix->adj0, adj0.data(), adj0.size() * 4);
            up(ix->upper_base, upper.data(), upper.size() * 4);
            up(ix->adjU, adjU.data(), adjU.size() * 4);
            up(ix->keys, keys.data(), keys.size() * 8);
            ix->identity_keys = true;
            for (size_t i = 0; i < pl.n; ++i) if (keys[i] != i) { ix->identity_keys = false; break; }
            LEANN_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
    } catch (...) {
        cudaStreamSynchronize(stream);
        leann_cuda_close(ix.release());
        throw;
    }
    return ix.release();
}

__global__ void adjacency_check_kernel(const uint32_t* __restrict__ adj, size_t count, uint32_t n, unsigned int* __restrict__ bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int b = 0;
    for (; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t s = adj[i];
        b |= (s != SENT && s >= n) ? 1u : 0u;
    }
    if (b) atomicOr(bad, 1u);
}

// DiskAnnSearcher::load (diskann.rs:21-43): the file already holds fixed-degree adjacency, both blocks are streamed;
// neighbour ids are range-checked on the device.
leann_cuda_index* open_vamana_streamed(const std::string& base, size_t dims, int device, int metric) {
    const std::string file = with_extension(base, "diskann");
    const DiskannPlan pl = diskann_probe(file, dims);
    std::unique_ptr<leann_cuda_index> ix(new leann_cuda_index());
    ix->backend = LEANN_BACKEND_VAMANA; ix->device = device;
    ix->metric = metric == LEANN_METRIC_DEFAULT ? LEANN_METRIC_IP_CLAMP : metric;   // diskann.rs:16,36 hard-wires DistDot
    ix->n = pl.n; ix->d = pl.d; ix->d4 = (uint32_t)((pl.d + 3) / 4);
    ix->M = (uint32_t)pl.R; ix->M0 = (uint32_t)pl.R; ix->max_level = 0; ix->entry = pl.medoid; ix->distance_name = pl.distance_name;
    ix->identity_keys = true;
    File f(file);
    cudaStream_t stream;
    LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } sg{stream};
    unsigned int* d_bad = nullptr;
    try {
        ix->vecs = dmalloc<float4>(pl.n * ix->d4);
        ix->adj0 = dmalloc<uint32_t>(pl.n * pl.R);
        Staging st;
        st.vectors(f, pl.vec_off, pl.n, pl.d, ix->d4, ix->vecs, stream);
        st.copy(f, pl.adj_off, pl.n * pl.R * 4, ix->adj0, stream, "adjacency");
        d_bad = dmalloc<unsigned int>(1);
        LEANN_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, 4, stream));
        if (pl.n) adjacency_check_kernel<<<592, 256, 0, stream>>>(ix->adj0, pl.n * pl.R, (uint32_t)pl.n, d_bad);
        unsigned int bad = 0;
        LEANN_CUDA_CHECK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, stream));
        LEANN_CUDA_CHECK(cudaStreamSynchronize(stream));
        cudaFree(d_bad);
        d_bad = nullptr;
        if (bad) throw Error(LEANN_ERR_BAD_FORMAT, file + ": neighbour id out of range");
    } catch (...) {
        cudaStreamSynchronize(stream);
        cudaFree(d_bad);
        leann_cuda_close(ix.release());
        throw;
    }
    return ix.release();
}

// Raw f32 `.embeddings` (index/embeddings.rs:21-36: count = file length / bytes per embedding) for the exact scan.
leann_cuda_index* open_flat_streamed(const std::string& base, size_t dims, int device, int metric) {
    if (dims == 0) throw Error(LEANN_ERR_INVALID_ARG, "embeddings: dimensions must be given (the file has no header)");
    const std::string file = with_extension(base, "embeddings");
    struct stat sb;
    if (stat(file.c_str(), &sb) != 0) throw Error(LEANN_ERR_NOT_FOUND, "Embeddings file not found: " + file);
    std::unique_ptr<leann_cuda_index> ix(new leann_cuda_index());
    ix->backend = LEANN_BACKEND_FLAT; ix->device = device;
    ix->metric = metric == LEANN_METRIC_DEFAULT ? LEANN_METRIC_DOT_DESC : metric;
    ix->n = (size_t)sb.st_size / (dims * 4); ix->d = dims; ix->d4 = (uint32_t)((dims + 3) / 4);
    File f(file);
    cudaStream_t stream;
    LEANN_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } sg{stream};
    try {
        ix->vecs = dmalloc<float4>(ix->n * ix->d4);
        Staging st;
        st.vectors(f, 0, ix->n, dims, ix->d4, ix->vecs, stream);
        LEANN_CUDA_CHECK(cudaStreamSynchronize(stream));
    } catch (...) {
        cudaStreamSynchronize(stream);
        leann_cuda_close(ix.release());
        throw;
    }
    return ix.release();
}

}  // namespace leann
