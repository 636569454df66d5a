// text.h — host-side text structures: tokenizer, BM25 inverted index, metadata filter, mini JSON.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace leann {

// ---- JSON value (serde_json::Value stand-in for filter evaluation) --------------------------------
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0.0;
    bool is_int = false;   // number written without fraction/exponent (serde i64/u64)
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json* get(const std::string& key) const;  // Value::get(&str): objects only
};
bool json_parse(const char* s, size_t n, Json& out, std::string& err);
std::string json_dump(const Json& v);

// ---- index/bm25.rs:127-132 ---------------------------------------------------------------------------
void tokenize(const char* text, size_t n, std::vector<std::string>& out);

// ---- index/filter.rs ---------------------------------------------------------------------------------------
enum class FilterOp { Eq, Ne, Gt, Gte, Lt, Lte, In, NotIn, Contains, StartsWith, EndsWith, Exists };
struct FilterNode {
    enum Kind { Condition, And, Or } kind = Condition;
    std::string field;
    FilterOp op = FilterOp::Eq;
    Json value;
    std::vector<FilterNode> children;
};
bool filter_parse(const std::string& s, FilterNode& out);             // MetadataFilter::parse
bool filter_matches(const FilterNode& f, const Json& metadata);       // MetadataFilter::matches
std::string filter_describe(const FilterNode& f);

// ---- columnar side-car of the passages' metadata -------------------------------------------
// Every dotted field path a filter can reach (filter.rs:376-388 splits on '.') becomes one typed column;
// a parsed filter is then evaluated column-wise into an N-bit mask without touching JSON again. The per-row
// semantics are exactly those of filter_matches (the condition code is shared).
struct MetaColumn {
    enum : uint8_t { Missing = 0, Null = 1, False = 2, True = 3, Num = 4, Str = 5, Other = 6 };
    std::vector<uint8_t> kind;        // per row
    std::vector<double> num;          // per row, valid when kind == Num (empty while the column holds no number)
    std::vector<uint32_t> sid;        // per row, dictionary id, valid when kind == Str
    std::vector<std::string> dict;
    std::unordered_map<std::string, uint32_t> dict_ids;
};
struct MetaColumns {
    size_t n = 0;
    std::unordered_map<std::string, MetaColumn> cols;
    void resize(size_t rows);
    void add_row(size_t row, const Json& metadata);                 // metadata of passage `row` (any JSON value)
    void eval(const FilterNode& f, std::vector<uint64_t>& mask) const;   // mask: ceil(n/64) words
};

// ---- index/bm25.rs:17-74: inverted (CSR) form of Bm25Scorer ------------------------------------------------
struct Bm25Host {
    size_t num_docs = 0;
    uint64_t total_tokens = 0;
    uint64_t n_postings = 0;
    float avg_doc_len = 1.0f;
    std::unordered_map<std::string, uint32_t> dict;  // term -> term id
    std::vector<uint64_t> term_off;   // n_terms + 1
    std::vector<float> idf;           // per term, f32 exactly as bm25.rs:88 (libm logf on the host)
    // device only: post_doc (ascending doc id inside a term) and post_score = idf * (tf * (K1 + 1)) / (tf + K1 * norm),
    // query independent, f32 as bm25.rs:88-100
};
// Corpus-wide statistics of a document-range sharded corpus: N, total tokens and df per term are
// global, so idf (bm25.rs:88), avg_doc_len (:61-65) and every per-posting score equal the unsharded ones bit for bit.
struct Bm25GlobalStats {
    uint64_t num_docs = 0, total_tokens = 0;
    std::unordered_map<std::string, uint64_t> df;
    std::string encode() const;                          // blob exchanged between ranks
    bool decode(const unsigned char* p, size_t n);
    void merge(const Bm25GlobalStats& other);
};
// The index itself is built on the device (bm25_build.cu); the host keeps the dictionary, the term offsets and idf.

}  // namespace leann
