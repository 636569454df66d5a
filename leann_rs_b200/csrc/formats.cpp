// formats.cpp — readers/writers of the reference's on-disk index formats, converting to the
// fixed-stride, sentinel-padded adjacency layout the kernels read from HBM.
//   usearch dense `.index`   written hnsw.rs:134, read hnsw.rs:55        
//   diskann-rs `.diskann`    written diskann.rs:94-99, read diskann.rs:34-37   (Appendix A.3)
//   raw f32 `.embeddings`    index/embeddings.rs:21-36,126-147
// Both graph formats are third-party serialisations recalled from upstream; every header field and
// the exact file-size equation are verified so a mismatch fails loudly instead of mis-reading.
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <fstream>

#include "internal.h"

namespace leann {

namespace {
#pragma pack(push, 1)
struct UsearchDenseHead {  // index_dense_head_t, 64 bytes
    char magic[7];
    uint16_t version_major, version_minor, version_patch;
    uint8_t kind_metric, kind_scalar, kind_key, kind_slot;
    uint64_t count_present, count_deleted, dimensions;
    uint8_t multi;
    uint8_t reserved[22];
};
struct UsearchGraphHead {  // index_serialized_header_t, 40 bytes
    uint64_t size, connectivity, connectivity_base, max_level, entry_slot;
};
#pragma pack(pop)
static_assert(sizeof(UsearchDenseHead) == 64, "dense head");
static_assert(sizeof(UsearchGraphHead) == 40, "graph head");

constexpr uint8_t KIND_F32 = 11, KIND_U64 = 14, KIND_U32 = 15;

size_t file_size(const std::string& p) {
    struct stat st;
    if (stat(p.c_str(), &st) != 0) return (size_t)-1;
    return (size_t)st.st_size;
}

struct Reader {
    FILE* f;
    std::string path;
    size_t pos = 0, size;
    Reader(const std::string& p) : f(fopen(p.c_str(), "rb")), path(p), size(file_size(p)) {}
    ~Reader() { if (f) fclose(f); }
    void read(void* dst, size_t bytes, const char* what) {
        if (pos + bytes > size || fread(dst, 1, bytes, f) != bytes)
            throw Error(LEANN_ERR_BAD_FORMAT, path + ": truncated while reading " + what);
        pos += bytes;
    }
};
// Writes go to `<path>.tmp`; every fwrite is checked, the file is flushed to disk and renamed over the target only
// when complete, so a short write (ENOSPC, crash) never truncates the only copy of an index (hnsw::add_to_index
// overwrites `.index` in place, hnsw.rs:183).
struct Writer {
    FILE* f;
    std::string path, tmp;
    explicit Writer(const std::string& p) : f(nullptr), path(p), tmp(p + ".tmp") {
        f = fopen(tmp.c_str(), "wb");
        if (!f) throw Error(LEANN_ERR_NOT_FOUND, "cannot create " + tmp + ": " + strerror(errno));
    }
    ~Writer() { if (f) { fclose(f); remove(tmp.c_str()); } }
    void write(const void* src, size_t size, size_t count) {
        if (count == 0 || size == 0) return;
        if (fwrite(src, size, count, f) != count) {
            const std::string why = strerror(errno);
            throw Error(LEANN_ERR_BAD_FORMAT, "write failed: " + tmp + ": " + why);
        }
    }
    void commit() {
        bool ok = fflush(f) == 0 && fsync(fileno(f)) == 0;
        ok = (fclose(f) == 0) && ok;
        f = nullptr;
        if (!ok || rename(tmp.c_str(), path.c_str()) != 0) {
            const std::string why = strerror(errno);
            remove(tmp.c_str());
            throw Error(LEANN_ERR_BAD_FORMAT, "write failed: " + path + ": " + why);
        }
    }
};
}  // namespace

std::string with_extension(const std::string& base, const std::string& ext) {
    // Rust Path::with_extension: replace everything after the last '.' of the file name.
    size_t slash = base.find_last_of('/');
    size_t dot = base.find_last_of('.');
    size_t name0 = slash == std::string::npos ? 0 : slash + 1;
    if (dot == std::string::npos || dot <= name0) return base + "." + ext;  // no extension (leading dot = hidden file)
    return base.substr(0, dot) + "." + ext;
}

bool is_faiss_index(const std::string& index_file) {
    FILE* f = fopen(index_file.c_str(), "rb");
    if (!f) return false;
    unsigned char h[4];
    bool ok = fread(h, 1, 4, f) == 4;
    fclose(f);
    if (!ok) return false;
    if (h[0] == 'I' && h[1] == 'x') return true;
    if (memcmp(h, "CSR\0", 4) == 0 || memcmp(h, "HNSW", 4) == 0) return true;
    return false;
}

uint64_t fnv1a64(const void* data, size_t bytes, uint64_t h) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 0x100000001B3ull; }
    return h;
}

void usearch_layout(const UsearchPlan& pl, const std::string& path, const int16_t* levels, std::vector<uint32_t>& upper_base,
                    std::vector<uint64_t>& node_off, size_t& n_upper);

// Header pass over a usearch `.index`: every header field is verified, the byte ranges of the three blocks are returned.
// The vectors block and the node block are NOT read here (leann_cuda_open streams them, open_stream.cu).
UsearchPlan usearch_probe(const std::string& path, size_t dims) {
    Reader r(path);
    if (!r.f) throw Error(LEANN_ERR_NOT_FOUND, "Index file not found: " + path + "\nRun 'leann build' to create an index first.");
    UsearchPlan pl;
    pl.file_size = r.size;
    struct stat st;
    pl.mtime_ns = stat(path.c_str(), &st) == 0 ? (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec : 0;
    uint32_t rc[2];
    r.read(rc, 8, "matrix shape");
    size_t rows = rc[0], cols = rc[1];
    if (cols == 0 || cols % 4 != 0) throw Error(LEANN_ERR_BAD_FORMAT, path + ": vector byte width is not a multiple of 4 (not an f32 usearch index)");
    if (8 + rows * cols + sizeof(UsearchDenseHead) + sizeof(UsearchGraphHead) > r.size)
        throw Error(LEANN_ERR_BAD_FORMAT, path + ": vectors block larger than file (incompatible format / header)");
    pl.vec_off = 8; pl.vec_bytes = rows * cols;
    if (fseek(r.f, (long)(8 + rows * cols), SEEK_SET) != 0) throw Error(LEANN_ERR_BAD_FORMAT, path + ": seek failed");
    r.pos = 8 + rows * cols;
    UsearchDenseHead h;
    r.read(&h, sizeof h, "dense head");
    if (memcmp(h.magic, "usearch", 7) != 0) throw Error(LEANN_ERR_BAD_FORMAT, path + ": bad magic (not a usearch index)");
    if (h.version_major != 2) throw Error(LEANN_ERR_BAD_FORMAT, path + ": unsupported usearch version " + std::to_string(h.version_major));
    if (h.kind_scalar != KIND_F32 || h.kind_key != KIND_U64 || h.kind_slot != KIND_U32)
        throw Error(LEANN_ERR_BAD_FORMAT, path + ": header scalar/key/slot kinds are not f32/u64/u32");
    if (h.kind_metric != 'i' && h.kind_metric != 'e' && h.kind_metric != 'c')
        throw Error(LEANN_ERR_BAD_FORMAT, path + ": header metric kind unsupported");
    if (h.dimensions * 4 != cols) throw Error(LEANN_ERR_BAD_FORMAT, path + ": header dimensions disagree with the vectors block");
    if (dims && h.dimensions != dims)
        throw Error(LEANN_ERR_DIM_MISMATCH, path + ": index has " + std::to_string(h.dimensions) + " dimensions, expected " + std::to_string(dims));
    if (h.multi) throw Error(LEANN_ERR_BAD_FORMAT, path + ": multi-vector indexes are not produced by leann (hnsw.rs:50)");
    UsearchGraphHead gh;
    r.read(&gh, sizeof gh, "graph header");
    if (gh.size != rows) throw Error(LEANN_ERR_BAD_FORMAT, path + ": graph size != vector rows");
    if (gh.connectivity == 0 || gh.connectivity_base < gh.connectivity || gh.connectivity_base > (uint64_t)MAX_DEG)
        throw Error(LEANN_ERR_BAD_FORMAT, path + ": connectivity out of range (max " + std::to_string(MAX_DEG) + ")");
    if (rows && gh.entry_slot >= rows) throw Error(LEANN_ERR_BAD_FORMAT, path + ": entry slot out of range");
    pl.n = rows; pl.d = h.dimensions; pl.M = gh.connectivity; pl.M0 = gh.connectivity_base;
    pl.max_level = (int64_t)gh.max_level; pl.entry = gh.entry_slot;
    pl.metric = h.kind_metric == 'e' ? LEANN_METRIC_L2SQ : LEANN_METRIC_IP;
    pl.levels_off = r.pos;
    pl.nodes_off = pl.levels_off + rows * 2;
    if (pl.nodes_off > r.size) throw Error(LEANN_ERR_BAD_FORMAT, path + ": truncated while reading levels");
    pl.head_hash = fnv1a64(&gh, sizeof gh, fnv1a64(&h, sizeof h, fnv1a64(rc, 8, 0xCBF29CE484222325ull)));
    // the level table is small (2 bytes per node): verify the file-size equation here, before any device work
    {
        std::vector<int16_t> levels(rows);
        r.read(levels.data(), rows * 2, "levels");
        std::vector<uint32_t> upper;
        std::vector<uint64_t> node_off;
        size_t n_upper = 0;
        usearch_layout(pl, path, levels.data(), upper, node_off, n_upper);
    }
    return pl;
}

// Levels table -> per-node upper-list base and byte offset inside the node block; verifies the file-size equation.
void usearch_layout(const UsearchPlan& pl, const std::string& path, const int16_t* levels, std::vector<uint32_t>& upper_base,
                    std::vector<uint64_t>& node_off, size_t& n_upper) {
    upper_base.assign(pl.n, 0);
    node_off.resize(pl.n + 1);
    n_upper = 0;
    uint64_t off = 0;
    for (size_t i = 0; i < pl.n; ++i) {
        if (levels[i] < 0 || levels[i] > pl.max_level) throw Error(LEANN_ERR_BAD_FORMAT, path + ": node level out of range");
        upper_base[i] = (uint32_t)n_upper;
        node_off[i] = off;
        n_upper += (size_t)levels[i];
        off += 10 + 4 * ((1 + pl.M0) + (uint64_t)levels[i] * (1 + pl.M));
    }
    node_off[pl.n] = off;
    if (n_upper > 0xFFFFFFF0ull) throw Error(LEANN_ERR_BAD_FORMAT, path + ": too many upper-level lists");
    if (pl.nodes_off + off != pl.file_size) {
        if (pl.nodes_off + off > pl.file_size) throw Error(LEANN_ERR_BAD_FORMAT, path + ": truncated while reading node links");
        throw Error(LEANN_ERR_BAD_FORMAT, path + ": file size equation violated (" + std::to_string(pl.file_size - pl.nodes_off - off) + " trailing bytes)");
    }
}

// Nodes [i0, i1) of the node block -> fixed-stride adjacency (SENT padded, list order kept) + keys. Thread-safe on disjoint ranges.
void usearch_parse_nodes(const UsearchPlan& pl, const std::string& path, const unsigned char* nodes, const uint64_t* node_off,
                         const int16_t* levels, const uint32_t* upper_base, size_t i0, size_t i1, uint64_t* keys, uint32_t* adj0,
                         uint32_t* adjU) {
    const size_t rows = pl.n, M0 = pl.M0, M = pl.M;
    for (size_t i = i0; i < i1; ++i) {
        const unsigned char* p = nodes + node_off[i];
        const int lv = levels[i];
        memcpy(&keys[i], p, 8);
        int16_t lv2;
        memcpy(&lv2, p + 8, 2);
        if (lv2 != lv) throw Error(LEANN_ERR_BAD_FORMAT, path + ": node level disagrees with the level table");
        uint32_t cnt;
        const unsigned char* l = p + 10;
        memcpy(&cnt, l, 4);
        if (cnt > M0) throw Error(LEANN_ERR_BAD_FORMAT, path + ": neighbour count exceeds connectivity_base");
        uint32_t* dst = adj0 + i * M0;
        memcpy(dst, l + 4, (size_t)cnt * 4);
        for (uint32_t j = 0; j < cnt; ++j)
            if (dst[j] >= rows) throw Error(LEANN_ERR_BAD_FORMAT, path + ": neighbour slot out of range");
        for (size_t j = cnt; j < M0; ++j) dst[j] = SENT;
        l += 4 * (1 + M0);
        for (int lev = 1; lev <= lv; ++lev, l += 4 * (1 + M)) {
            memcpy(&cnt, l, 4);
            if (cnt > M) throw Error(LEANN_ERR_BAD_FORMAT, path + ": neighbour count exceeds connectivity");
            uint32_t* du = adjU + ((size_t)upper_base[i] + (size_t)(lev - 1)) * M;
            memcpy(du, l + 4, (size_t)cnt * 4);
            for (uint32_t j = 0; j < cnt; ++j)
                if (du[j] >= rows) throw Error(LEANN_ERR_BAD_FORMAT, path + ": neighbour slot out of range");
            for (size_t j = cnt; j < M; ++j) du[j] = SENT;
        }
    }
}

// Whole-file reader into host memory (tools and tests; leann_cuda_open streams instead).
void read_usearch_index(const std::string& path, size_t dims, HostHnsw& g) {
    UsearchPlan pl = usearch_probe(path, dims);
    Reader r(path);
    if (!r.f) throw Error(LEANN_ERR_NOT_FOUND, "Index file not found: " + path);
    g.n = pl.n; g.d = pl.d; g.M = pl.M; g.M0 = pl.M0; g.max_level = pl.max_level; g.entry = pl.entry; g.metric = pl.metric;
    fseek(r.f, (long)pl.vec_off, SEEK_SET);
    r.pos = pl.vec_off;
    g.vecs.resize(pl.n * pl.d);
    r.read(g.vecs.data(), pl.vec_bytes, "vectors");
    fseek(r.f, (long)pl.levels_off, SEEK_SET);
    r.pos = pl.levels_off;
    g.levels.resize(pl.n);
    r.read(g.levels.data(), pl.n * 2, "levels");
    std::vector<uint64_t> node_off;
    size_t n_upper = 0;
    usearch_layout(pl, path, g.levels.data(), g.upper_base, node_off, n_upper);
    std::vector<unsigned char> nodes(node_off[pl.n]);
    r.read(nodes.data(), nodes.size(), "node links");
    g.keys.resize(pl.n);
    g.adj0.resize(pl.n * pl.M0);
    g.adjU.resize(n_upper * pl.M);
    usearch_parse_nodes(pl, path, nodes.data(), node_off.data(), g.levels.data(), g.upper_base.data(), 0, pl.n, g.keys.data(),
                        g.adj0.data(), g.adjU.data());
}

void write_usearch_index(const std::string& path, const HostHnsw& g) {
    Writer w(path);
    uint32_t rc[2] = {(uint32_t)g.n, (uint32_t)(g.d * 4)};
    w.write(rc, 4, 2);
    w.write(g.vecs.data(), 4, g.n * g.d);
    UsearchDenseHead h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "usearch", 7);
    h.version_major = 2; h.version_minor = 23; h.version_patch = 0;  // Cargo.lock:4381-4388
    h.kind_metric = g.metric == LEANN_METRIC_L2SQ ? 'e' : 'i';
    h.kind_scalar = KIND_F32; h.kind_key = KIND_U64; h.kind_slot = KIND_U32;
    h.count_present = g.n; h.count_deleted = 0; h.dimensions = g.d; h.multi = 0;
    w.write(&h, sizeof h, 1);
    UsearchGraphHead gh{g.n, g.M, g.M0, (uint64_t)g.max_level, g.entry};
    w.write(&gh, sizeof gh, 1);
    w.write(g.levels.data(), 2, g.n);
    std::vector<uint32_t> node;
    for (size_t i = 0; i < g.n; ++i) {
        int lv = g.levels[i];
        node.assign((1 + g.M0) + (size_t)lv * (1 + g.M), 0);
        uint32_t c = 0;
        for (size_t j = 0; j < g.M0; ++j) { uint32_t s = g.adj0[i * g.M0 + j]; if (s != SENT) node[1 + c++] = s; }
        node[0] = c;
        for (int l = 1; l <= lv; ++l) {
            uint32_t* p = &node[(1 + g.M0) + (size_t)(l - 1) * (1 + g.M)];
            const uint32_t* src = &g.adjU[((size_t)g.upper_base[i] + (l - 1)) * g.M];
            uint32_t cu = 0;
            for (size_t j = 0; j < g.M; ++j) if (src[j] != SENT) p[1 + cu++] = src[j];
            p[0] = cu;
        }
        uint64_t key = g.keys.empty() ? (uint64_t)i : g.keys[i];
        int16_t lv16 = (int16_t)lv;
        w.write(&key, 8, 1);
        w.write(&lv16, 2, 1);
        w.write(node.data(), 4, node.size());
    }
    w.commit();
}

// bincode 1 (fixint, little endian): usize -> u64, String -> u64 length + bytes. Header pass only.
DiskannPlan diskann_probe(const std::string& path, size_t dims) {
    Reader r(path);
    if (!r.f) throw Error(LEANN_ERR_NOT_FOUND, "DiskANN index not found: " + path + "\nRun 'leann build' with --backend-name diskann to create an index first.");
    DiskannPlan g;
    uint64_t meta_len;
    r.read(&meta_len, 8, "meta length");
    if (meta_len < 52 || meta_len > 4096) throw Error(LEANN_ERR_BAD_FORMAT, path + ": implausible metadata length");
    std::vector<unsigned char> m(meta_len);
    r.read(m.data(), meta_len, "metadata");
    auto u64 = [&](size_t o) { uint64_t v; memcpy(&v, &m[o], 8); return v; };
    g.d = u64(0); g.n = u64(8); g.R = u64(16);
    memcpy(&g.medoid, &m[24], 4);
    uint64_t voff = u64(28), aoff = u64(36), nlen = u64(44);
    if (52 + nlen != meta_len) throw Error(LEANN_ERR_BAD_FORMAT, path + ": metadata length disagrees with its string field");
    g.distance_name.assign((const char*)&m[52], nlen);
    if (g.R == 0 || g.R > (size_t)MAX_DEG) throw Error(LEANN_ERR_BAD_FORMAT, path + ": max_degree out of range");
    if (dims && g.d != dims) throw Error(LEANN_ERR_DIM_MISMATCH, path + ": index has " + std::to_string(g.d) + " dimensions, expected " + std::to_string(dims));
    if (voff < 8 + meta_len || aoff != voff + (uint64_t)g.n * g.d * 4 || r.size != aoff + (uint64_t)g.n * g.R * 4)
        throw Error(LEANN_ERR_BAD_FORMAT, path + ": file size equation violated");
    if (g.n && g.medoid >= g.n) throw Error(LEANN_ERR_BAD_FORMAT, path + ": medoid out of range");
    g.vec_off = voff; g.adj_off = aoff; g.file_size = r.size;
    return g;
}

void read_diskann(const std::string& path, size_t dims, HostVamana& g) {
    DiskannPlan pl = diskann_probe(path, dims);
    Reader r(path);
    g.n = pl.n; g.d = pl.d; g.R = pl.R; g.medoid = pl.medoid; g.distance_name = pl.distance_name;
    fseek(r.f, (long)pl.vec_off, SEEK_SET);
    r.pos = pl.vec_off;
    g.vecs.resize(g.n * g.d);
    r.read(g.vecs.data(), g.n * g.d * 4, "vectors");
    g.adj.resize(g.n * g.R);
    r.read(g.adj.data(), g.n * g.R * 4, "adjacency");
    for (uint32_t s : g.adj)
        if (s != SENT && s >= g.n) throw Error(LEANN_ERR_BAD_FORMAT, path + ": neighbour id out of range");
}

void write_diskann(const std::string& path, const HostVamana& g) {
    std::vector<unsigned char> m;
    auto p64 = [&](uint64_t v) { for (int i = 0; i < 8; ++i) m.push_back((unsigned char)(v >> (8 * i))); };
    const uint64_t voff = 1u << 20, aoff = voff + (uint64_t)g.n * g.d * 4;
    std::string name = g.distance_name.empty() ? "DistDot" : g.distance_name;
    p64(g.d); p64(g.n); p64(g.R);
    for (int i = 0; i < 4; ++i) m.push_back((unsigned char)(g.medoid >> (8 * i)));
    p64(voff); p64(aoff); p64(name.size());
    m.insert(m.end(), name.begin(), name.end());
    Writer w(path);
    uint64_t ml = m.size();
    w.write(&ml, 8, 1);
    w.write(m.data(), 1, m.size());
    std::vector<unsigned char> z(voff - 8 - m.size(), 0);
    w.write(z.data(), 1, z.size());
    w.write(g.vecs.data(), 4, g.n * g.d);
    w.write(g.adj.data(), 4, g.n * g.R);
    w.commit();
}

void read_embeddings(const std::string& path, size_t dims, std::vector<float>& out, size_t& n) {
    if (dims == 0) throw Error(LEANN_ERR_INVALID_ARG, "embeddings: dimensions must be given (the file has no header)");
    Reader r(path);
    if (!r.f) throw Error(LEANN_ERR_NOT_FOUND, "Embeddings file not found: " + path);
    n = r.size / (dims * 4);  // embeddings.rs:26-27: count = len / bytes_per_embedding
    out.resize(n * dims);
    r.read(out.data(), n * dims * 4, "embeddings");
}

void write_embeddings(const std::string& path, const float* v, size_t n, size_t dims) {
    Writer w(path);
    w.write(v, 4, n * dims);
    w.commit();
}

}  // namespace leann
