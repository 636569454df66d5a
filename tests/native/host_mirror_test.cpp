// Exercises the C++ host mirror (leann_rs_b200/host/leann_cuda.hpp) the way the reference's own unit tests
// exercise the Rust types: bm25.rs:264-329 and filter.rs:446-551 assertions, plus one BackendSearcher::search
// call whose result is printed for the Python side to compare. Usage: host_mirror_test <base_path> <dims> <query.f32> [scratch_base]
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../leann_rs_b200/host/leann_cuda.hpp"

#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "REQUIRE failed line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    using namespace leann;
    if (argc < 4) return 2;
    try {
        // --- bm25.rs:264-280 test_bm25_search_top_k
        auto scorer = Bm25Scorer::build({"apple banana", "apple cherry", "banana cherry", "apple apple apple"});
        auto r = scorer.search("apple", 2);
        REQUIRE(r.size() == 2 && r[0].first == 3 && r[1].first == 0);
        auto dense = scorer.score_query("apple");
        REQUIRE(dense.size() == 4 && dense[2] == 0.0f && dense[3] == r[0].second && dense[0] == dense[1]);
        REQUIRE(Bm25Scorer::build({"hello world"}).search("xyz", 5).empty());            // bm25.rs:255-262
        // --- bm25.rs:282-329 hybrid_rerank
        auto h = hybrid_rerank({{0, 0.9f}, {1, 0.8f}, {2, 0.7f}}, {0.5f, 0.9f, 0.3f}, 0.5f);
        REQUIRE(h.size() == 3 && h[0].first == 1 && h[1].first == 0 && h[2].first == 2 && h[2].second == 0.0f);
        REQUIRE(hybrid_rerank({{0, 0.9f}, {1, 0.5f}}, {0.1f, 0.9f}, 1.0f)[0].first == 0);
        REQUIRE(hybrid_rerank({{0, 0.9f}, {1, 0.5f}}, {0.1f, 0.9f}, 0.0f)[0].first == 1);
        // --- filter.rs:446-551
        const char* md = "{\"source\": \"main.rs\", \"type\": \"code\", \"lines\": 100}";
        REQUIRE(MetadataFilter::parse("source:*.rs")->matches(md));
        REQUIRE(MetadataFilter::parse("type=code,lines>50")->matches(md));
        REQUIRE(!MetadataFilter::parse("type=code,lines>200")->matches(md));
        REQUIRE(MetadataFilter::parse("type in [code,text,doc]")->matches(md));
        REQUIRE(!MetadataFilter::parse("type not_in [code,text]")->matches(md));
        REQUIRE(MetadataFilter::parse("type=text OR type=code")->matches(md));
        REQUIRE(!MetadataFilter::parse("missing?")->matches(md));
        REQUIRE(!MetadataFilter::parse("nonsense").has_value());
        // --- BackendType::load_searcher + BackendSearcher::search (traits.rs:16-21)
        size_t d = (size_t)atol(argv[2]);
        auto s = load_searcher(BackendType::Hnsw, argv[1], d);
        REQUIRE(!s->is_empty());
        std::vector<float> q(d);
        {
            FILE* f = fopen(argv[3], "rb");   // raw f32 query written by the caller
            REQUIRE(f && fread(q.data(), 4, d, f) == d);
            fclose(f);
        }
        auto res = s->search(q, 5, 999);   // complexity ignored for HNSW (hnsw.rs:83)
        REQUIRE(res.first.size() == 5 && res.second.size() == 5);
        for (size_t i = 1; i < 5; ++i) REQUIRE(res.second[i - 1] <= res.second[i]);
        printf("KEYS");
        for (auto k : res.first) printf(" %llu", (unsigned long long)k);
        printf("\n");
        // --- the sharded BackendSearcher over the same file as a one-shard index: identical answer through the merge kernel
        {
            ShardedSearcher sh({argv[1]}, LEANN_BACKEND_HNSW, d, {0}, 64);
            REQUIRE(sh.len() == s->len() && sh.shards() == 1);
            auto rs = sh.search(q, 5, 999);
            REQUIRE(rs.first == res.first && rs.second == res.second);
        }
        // --- hnsw::build_index + hnsw::add_to_index (hnsw.rs:96-191) on a scratch base path, then load and search
        if (argc > 4) {
            const std::string base = argv[4];
            std::vector<std::vector<float>> first, more;
            for (int i = 0; i < 300; ++i) {
                std::vector<float> v(8);
                for (int j = 0; j < 8; ++j) v[j] = std::sin(0.37f * (float)(i + 1) * (float)(j + 1));
                (i < 200 ? first : more).push_back(v);
            }
            hnsw::build_index(first, {}, base, 8, 8, 32);
            REQUIRE(HnswSearcher::load(base, 8)->len() == 200);
            hnsw::add_to_index(more, base, 8, 200);
            auto grown = HnswSearcher::load(base, 8);
            REQUIRE(grown->len() == 300);
            auto hit = grown->search(more[42], 1, 64);            // an appended vector finds itself under its key start_id + i
            REQUIRE(hit.first.size() == 1 && hit.first[0] == 242);
            try { hnsw::add_to_index({{1.0f, 2.0f}}, base, 8, 300); REQUIRE(false); } catch (const Error& e) { REQUIRE(e.code == LEANN_ERR_DIM_MISMATCH); }
            // --- MetadataColumns == MetadataFilter::mask
            std::vector<std::string> metas = {"{\"lines\": 5, \"type\": \"code\"}", "{\"lines\": 500}", "{}", "{\"type\": \"doc\", \"lines\": 7}"};
            MetadataColumns cols(metas);
            for (const char* expr : {"lines>6", "type=code", "type!=code", "lines>=5,type?", "type=doc OR lines<6"}) {
                auto f = MetadataFilter::parse(expr);
                REQUIRE(f.has_value() && cols.mask(*f) == f->mask(metas));
            }
        }
        try {
            HnswSearcher::load("/nonexistent/documents.leann", d);
            REQUIRE(false);
        } catch (const Error& e) { REQUIRE(e.code == LEANN_ERR_NOT_FOUND); }
    } catch (const std::exception& e) {
        fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    printf("OK\n");
    return 0;
}
