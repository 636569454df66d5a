// text.cpp — host side of the text path:
//   tokenize            index/bm25.rs:127-132   regex [a-zA-Z0-9]+, lower-case, drop byte length <= 1
//   bm25_build_host     index/bm25.rs:33-74     same statistics as Bm25Scorer::build, stored inverted (CSR)
//   filter_parse        index/filter.rs:52-134, 137-316, 420-439
//   filter_matches      index/filter.rs:319-418
//   json_parse          serde_json::Value stand-in (metadata documents of index/passages.rs:12-17)
#include "text.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace leann {

// ================================= JSON =================================
namespace {
struct JP {
    const char* p; const char* e; std::string err;
    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool fail(const char* m) { if (err.empty()) err = m; return false; }
    static void utf8(std::string& o, uint32_t c) {
        if (c < 0x80) o += (char)c;
        else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 63)); }
        else if (c < 0x10000) { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
        else { o += (char)(0xF0 | (c >> 18)); o += (char)(0x80 | ((c >> 12) & 63)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
    }
    bool hex4(uint32_t& v) {
        if (e - p < 4) return fail("bad \\u escape");
        v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else return fail("bad \\u escape");
        }
        return true;
    }
    bool str(std::string& o) {
        if (p >= e || *p != '"') return fail("expected string");
        ++p;
        while (p < e && *p != '"') {
            if (*p == '\\') {
                if (++p >= e) return fail("bad escape");
                char c = *p++;
                switch (c) {
                    case '"': o += '"'; break; case '\\': o += '\\'; break; case '/': o += '/'; break;
                    case 'b': o += '\b'; break; case 'f': o += '\f'; break; case 'n': o += '\n'; break;
                    case 'r': o += '\r'; break; case 't': o += '\t'; break;
                    case 'u': {
                        uint32_t c1;
                        if (!hex4(c1)) return false;
                        if (c1 >= 0xD800 && c1 < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            p += 2;
                            uint32_t c2;
                            if (!hex4(c2)) return false;
                            c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
                        }
                        utf8(o, c1);
                        break;
                    }
                    default: return fail("bad escape");
                }
            } else {
                o += *p++;
            }
        }
        if (p >= e) return fail("unterminated string");
        ++p;
        return true;
    }
    bool val(Json& v, int depth) {
        if (depth > 128) return fail("nesting too deep");
        ws();
        if (p >= e) return fail("unexpected end");
        char c = *p;
        if (c == '{') {
            ++p; v.kind = Json::Obj; ws();
            if (p < e && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!str(k)) return false;
                ws();
                if (p >= e || *p != ':') return fail("expected ':'");
                ++p;
                Json child;
                if (!val(child, depth + 1)) return false;
                // serde_json maps keep the last value of a duplicated key
                bool rep = false;
                for (auto& kv : v.obj) if (kv.first == k) { kv.second = std::move(child); rep = true; break; }
                if (!rep) v.obj.emplace_back(std::move(k), std::move(child));
                ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p; v.kind = Json::Arr; ws();
            if (p < e && *p == ']') { ++p; return true; }
            for (;;) {
                Json child;
                if (!val(child, depth + 1)) return false;
                v.arr.push_back(std::move(child));
                ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') { v.kind = Json::Str; return str(v.str); }
        if (e - p >= 4 && !memcmp(p, "true", 4)) { p += 4; v.kind = Json::Bool; v.b = true; return true; }
        if (e - p >= 5 && !memcmp(p, "false", 5)) { p += 5; v.kind = Json::Bool; v.b = false; return true; }
        if (e - p >= 4 && !memcmp(p, "null", 4)) { p += 4; v.kind = Json::Null; return true; }
        if (c == '-' || (c >= '0' && c <= '9')) {
            const char* s = p;
            bool isint = true;
            if (*p == '-') ++p;
            while (p < e && *p >= '0' && *p <= '9') ++p;
            if (p < e && *p == '.') { isint = false; ++p; while (p < e && *p >= '0' && *p <= '9') ++p; }
            if (p < e && (*p == 'e' || *p == 'E')) { isint = false; ++p; if (p < e && (*p == '+' || *p == '-')) ++p; while (p < e && *p >= '0' && *p <= '9') ++p; }
            std::string t(s, p);
            v.kind = Json::Num; v.is_int = isint; v.num = strtod(t.c_str(), nullptr);
            return true;
        }
        return fail("unexpected character");
    }
};

void dump_str(std::string& o, const std::string& s) {
    o += '"';
    for (unsigned char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '\n') o += "\\n";
        else if (c == '\r') o += "\\r";
        else if (c == '\t') o += "\\t";
        else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o += (char)c;
    }
    o += '"';
}
void dump(std::string& o, const Json& v) {
    switch (v.kind) {
        case Json::Null: o += "null"; break;
        case Json::Bool: o += v.b ? "true" : "false"; break;
        case Json::Num: {
            char b[40];
            if (v.is_int && std::fabs(v.num) < 9.2e18) snprintf(b, sizeof b, "%lld", (long long)v.num);
            else snprintf(b, sizeof b, "%.17g", v.num);
            o += b;
            break;
        }
        case Json::Str: dump_str(o, v.str); break;
        case Json::Arr: o += '['; for (size_t i = 0; i < v.arr.size(); ++i) { if (i) o += ','; dump(o, v.arr[i]); } o += ']'; break;
        case Json::Obj: o += '{'; for (size_t i = 0; i < v.obj.size(); ++i) { if (i) o += ','; dump_str(o, v.obj[i].first); o += ':'; dump(o, v.obj[i].second); } o += '}'; break;
    }
}
}  // namespace

const Json* Json::get(const std::string& key) const {
    if (kind != Obj) return nullptr;
    for (auto& kv : obj) if (kv.first == key) return &kv.second;
    return nullptr;
}
bool json_parse(const char* s, size_t n, Json& out, std::string& err) {
    JP jp{s, s + n, {}};
    out = Json();
    if (!jp.val(out, 0)) { err = jp.err; return false; }
    jp.ws();
    if (jp.p != jp.e) { err = "trailing characters"; return false; }
    return true;
}
std::string json_dump(const Json& v) { std::string o; dump(o, v); return o; }

// ================================= tokenizer =================================
void tokenize(const char* text, size_t n, std::vector<std::string>& out) {
    out.clear();
    size_t i = 0;
    auto alnum = [](unsigned char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9'); };
    while (i < n) {
        while (i < n && !alnum((unsigned char)text[i])) ++i;
        size_t s = i;
        while (i < n && alnum((unsigned char)text[i])) ++i;
        if (i - s > 1) {  // .filter(|s| s.len() > 1)
            std::string t(text + s, i - s);
            for (auto& c : t) if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
            out.push_back(std::move(t));
        }
    }
}

// ================================= BM25 statistics blobs =================================
void Bm25GlobalStats::merge(const Bm25GlobalStats& o) {
    num_docs += o.num_docs;
    total_tokens += o.total_tokens;
    for (auto& kv : o.df) df[kv.first] += kv.second;
}
std::string Bm25GlobalStats::encode() const {
    std::string out;
    auto put64 = [&](uint64_t v) { out.append(reinterpret_cast<const char*>(&v), 8); };
    put64(0x314D42534E41454Cull);  // "LEANSBM1"
    put64(num_docs); put64(total_tokens); put64(df.size());
    std::vector<const std::pair<const std::string, uint64_t>*> order;
    order.reserve(df.size());
    for (auto& kv : df) order.push_back(&kv);
    std::sort(order.begin(), order.end(), [](auto* a, auto* b) { return a->first < b->first; });   // deterministic bytes
    for (auto* kv : order) {
        uint32_t len = (uint32_t)kv->first.size();
        out.append(reinterpret_cast<const char*>(&len), 4);
        out.append(kv->first);
        put64(kv->second);
    }
    return out;
}
bool Bm25GlobalStats::decode(const unsigned char* p, size_t n) {
    *this = Bm25GlobalStats();
    size_t o = 0;
    auto get64 = [&](uint64_t& v) { if (o + 8 > n) return false; memcpy(&v, p + o, 8); o += 8; return true; };
    uint64_t magic = 0, nt = 0;
    if (!get64(magic) || magic != 0x314D42534E41454Cull || !get64(num_docs) || !get64(total_tokens) || !get64(nt)) return false;
    for (uint64_t i = 0; i < nt; ++i) {
        uint32_t len = 0;
        if (o + 4 > n) return false;
        memcpy(&len, p + o, 4); o += 4;
        if (o + len > n) return false;
        std::string term(reinterpret_cast<const char*>(p + o), len); o += len;
        uint64_t v = 0;
        if (!get64(v)) return false;
        df[term] += v;
    }
    return o == n;
}

// ================================= filter =================================
namespace {
// Rust str::trim (filter.rs:53,60,122,138,151,156,168,173): strips chars with the Unicode White_Space property — U+0009..000D,
// U+0020, U+0085, U+00A0, U+1680, U+2000..200A, U+2028, U+2029, U+202F, U+205F, U+3000 — and nothing else (not U+001C..001F).
// Returns the byte length of the white-space char that starts at s[i], 0 if there is none.
size_t ws_char_len(const std::string& s, size_t i) {
    const size_t n = s.size();
    const unsigned char c = (unsigned char)s[i];
    if ((c >= 0x09 && c <= 0x0D) || c == 0x20) return 1;
    if (c == 0xC2 && i + 1 < n) {
        const unsigned char d = (unsigned char)s[i + 1];
        return (d == 0x85 || d == 0xA0) ? 2 : 0;
    }
    if (i + 2 < n) {
        const unsigned char d = (unsigned char)s[i + 1], e = (unsigned char)s[i + 2];
        if (c == 0xE1) return (d == 0x9A && e == 0x80) ? 3 : 0;                                   // U+1680
        if (c == 0xE2) {
            if (d == 0x80) return ((e >= 0x80 && e <= 0x8A) || e == 0xA8 || e == 0xA9 || e == 0xAF) ? 3 : 0;   // U+2000..200A, 2028, 2029, 202F
            return (d == 0x81 && e == 0x9F) ? 3 : 0;                                              // U+205F
        }
        if (c == 0xE3) return (d == 0x80 && e == 0x80) ? 3 : 0;                                   // U+3000
    }
    return 0;
}
std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    for (size_t l; a < b && (l = ws_char_len(s, a)) != 0;) a += l;
    while (b > a) {
        size_t st = b - 1;   // start of the last char: step back over at most 3 continuation bytes
        for (int k = 0; k < 3 && st > a && ((unsigned char)s[st] & 0xC0) == 0x80; ++k) --st;
        const size_t l = ws_char_len(s, st);
        if (l == 0 || st + l != b) break;
        b = st;
    }
    return s.substr(a, b - a);
}
std::vector<std::string> split(const std::string& s, const std::string& sep) {
    std::vector<std::string> out;
    size_t pos = 0;
    for (;;) {
        size_t f = s.find(sep, pos);
        if (f == std::string::npos) { out.push_back(s.substr(pos)); break; }
        out.push_back(s.substr(pos, f - pos));
        pos = f + sep.size();
    }
    return out;
}
bool parse_i64(const std::string& s, long long& v) {  // Rust i64::from_str: [+-]?[0-9]+ with overflow check
    if (s.empty()) return false;
    size_t i = 0;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
    if (i >= s.size()) return false;
    unsigned long long acc = 0;
    const unsigned long long lim = neg ? 9223372036854775808ull : 9223372036854775807ull;
    for (; i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') return false;
        unsigned d = s[i] - '0';
        if (acc > (lim - d) / 10) return false;
        acc = acc * 10 + d;
    }
    v = neg ? (long long)(0ull - acc) : (long long)acc;
    return true;
}
bool parse_f64(const std::string& s, double& v) {  // Rust f64::from_str grammar; non-finite rejected by Number::from_f64
    if (s.empty()) return false;
    size_t i = 0;
    if (s[i] == '+' || s[i] == '-') ++i;
    std::string rest = s.substr(i);
    std::string low = rest;
    for (auto& c : low) c = (char)tolower((unsigned char)c);
    if (low == "inf" || low == "infinity" || low == "nan") return false;  // parses, but from_f64 -> None
    size_t digits = 0, j = i;
    while (j < s.size() && isdigit((unsigned char)s[j])) { ++j; ++digits; }
    if (j < s.size() && s[j] == '.') { ++j; while (j < s.size() && isdigit((unsigned char)s[j])) { ++j; ++digits; } }
    if (digits == 0) return false;
    if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
        ++j;
        if (j < s.size() && (s[j] == '+' || s[j] == '-')) ++j;
        size_t ed = 0;
        while (j < s.size() && isdigit((unsigned char)s[j])) { ++j; ++ed; }
        if (ed == 0) return false;
    }
    if (j != s.size()) return false;
    v = strtod(s.c_str(), nullptr);
    return std::isfinite(v);
}
Json parse_value(const std::string& s) {  // filter.rs:420-439
    Json v;
    long long iv;
    if (parse_i64(s, iv)) { v.kind = Json::Num; v.num = (double)iv; v.is_int = true; return v; }
    double dv;
    if (parse_f64(s, dv)) { v.kind = Json::Num; v.num = dv; v.is_int = false; return v; }
    if (s == "true") { v.kind = Json::Bool; v.b = true; return v; }
    if (s == "false") { v.kind = Json::Bool; v.b = false; return v; }
    v.kind = Json::Str; v.str = s;
    return v;
}
Json jstr(const std::string& s) { Json v; v.kind = Json::Str; v.str = s; return v; }
bool cond(FilterNode& out, const std::string& field, FilterOp op, Json value) {
    out = FilterNode();
    out.kind = FilterNode::Condition; out.field = field; out.op = op; out.value = std::move(value);
    return true;
}
bool contains(const std::string& s, const char* t) { return s.find(t) != std::string::npos; }
bool splitn2(const std::string& s, const std::string& sep, std::string& a, std::string& b) {
    size_t f = s.find(sep);
    if (f == std::string::npos) return false;
    a = s.substr(0, f); b = s.substr(f + sep.size());
    return true;
}

bool parse_single(const std::string& in, FilterNode& out) {  // filter.rs:137-316
    std::string s = trim(in);
    if (!s.empty() && s.back() == '?') return cond(out, s.substr(0, s.size() - 1), FilterOp::Exists, Json());
    for (int pass = 0; pass < 2; ++pass) {
        const char* key = pass == 0 ? " in [" : " not_in [";
        size_t klen = strlen(key);
        size_t idx = s.find(key);
        if (idx != std::string::npos) {
            std::string field = trim(s.substr(0, idx));
            std::string rest = s.substr(idx + klen);
            size_t end = rest.find(']');
            if (end != std::string::npos) {
                Json arr; arr.kind = Json::Arr;
                for (auto& v : split(rest.substr(0, end), ",")) arr.arr.push_back(parse_value(trim(v)));
                return cond(out, field, pass == 0 ? FilterOp::In : FilterOp::NotIn, std::move(arr));
            }
        }
    }
    std::string a, b;
    if (contains(s, "~")) { splitn2(s, "~", a, b); return cond(out, a, FilterOp::Contains, jstr(b)); }
    if (contains(s, "^") && !contains(s, ">=")) { splitn2(s, "^", a, b); return cond(out, a, FilterOp::StartsWith, jstr(b)); }
    if (contains(s, "$")) { splitn2(s, "$", a, b); return cond(out, a, FilterOp::EndsWith, jstr(b)); }
    if (contains(s, "!=")) { splitn2(s, "!=", a, b); return cond(out, a, FilterOp::Ne, parse_value(b)); }
    if (contains(s, ">=")) { splitn2(s, ">=", a, b); return cond(out, a, FilterOp::Gte, parse_value(b)); }
    if (contains(s, "<=")) { splitn2(s, "<=", a, b); return cond(out, a, FilterOp::Lte, parse_value(b)); }
    if (contains(s, ">")) { splitn2(s, ">", a, b); return cond(out, a, FilterOp::Gt, parse_value(b)); }
    if (contains(s, "<")) { splitn2(s, "<", a, b); return cond(out, a, FilterOp::Lt, parse_value(b)); }
    if (contains(s, "=")) splitn2(s, "=", a, b);
    else if (contains(s, ":")) splitn2(s, ":", a, b);
    else return false;
    const std::string& value = b;
    if (contains(value, "*")) {
        bool st = value.front() == '*', en = value.back() == '*';
        if (st && en && value.size() > 2) return cond(out, a, FilterOp::Contains, jstr(value.substr(1, value.size() - 2)));
        if (st) return cond(out, a, FilterOp::EndsWith, jstr(value.substr(1)));
        if (en) return cond(out, a, FilterOp::StartsWith, jstr(value.substr(0, value.size() - 1)));
    }
    return cond(out, a, FilterOp::Eq, parse_value(value));
}

const Json* nested(const Json& md, const std::string& path) {  // filter.rs:376-388
    const Json* cur = &md;
    for (auto& part : split(path, ".")) {
        cur = cur->get(part);
        if (!cur) return nullptr;
    }
    return cur;
}
bool values_equal(const Json& a, const Json& b) {  // filter.rs:390-400
    if (a.kind == Json::Str && b.kind == Json::Str) return a.str == b.str;
    if (a.kind == Json::Num && b.kind == Json::Num) return std::fabs(a.num - b.num) < 2.220446049250313e-16;
    if (a.kind == Json::Bool && b.kind == Json::Bool) return a.b == b.b;
    if (a.kind == Json::Null && b.kind == Json::Null) return true;
    return false;
}
int compare_values(const Json& a, const Json& b) {  // filter.rs:402-418
    if (a.kind == Json::Num && b.kind == Json::Num) return a.num < b.num ? -1 : (a.num > b.num ? 1 : 0);
    if (a.kind == Json::Str && b.kind == Json::Str) { int c = a.str.compare(b.str); return c < 0 ? -1 : (c > 0 ? 1 : 0); }
    return 0;
}
// filter.rs:329-373 on an already resolved field value (nullptr = field missing)
bool cond_on_value(const FilterNode& c, const Json* fv) {
    auto pat = [&]() -> const std::string& { static const std::string empty; return c.value.kind == Json::Str ? c.value.str : empty; };
    switch (c.op) {
        case FilterOp::Exists: return fv != nullptr;
        case FilterOp::Eq: return fv && values_equal(*fv, c.value);
        case FilterOp::Ne: return !fv || !values_equal(*fv, c.value);
        case FilterOp::Gt: return fv && compare_values(*fv, c.value) > 0;
        case FilterOp::Gte: return fv && compare_values(*fv, c.value) >= 0;
        case FilterOp::Lt: return fv && compare_values(*fv, c.value) < 0;
        case FilterOp::Lte: return fv && compare_values(*fv, c.value) <= 0;
        case FilterOp::In:
            if (c.value.kind != Json::Arr) return false;
            if (!fv) return false;
            for (auto& it : c.value.arr) if (values_equal(*fv, it)) return true;
            return false;
        case FilterOp::NotIn:
            if (c.value.kind != Json::Arr) return true;
            if (!fv) return true;
            for (auto& it : c.value.arr) if (values_equal(*fv, it)) return false;
            return true;
        case FilterOp::Contains: return fv && fv->kind == Json::Str && fv->str.find(pat()) != std::string::npos;
        case FilterOp::StartsWith: return fv && fv->kind == Json::Str && fv->str.compare(0, pat().size(), pat()) == 0 && fv->str.size() >= pat().size();
        case FilterOp::EndsWith:
            return fv && fv->kind == Json::Str && fv->str.size() >= pat().size() &&
                   fv->str.compare(fv->str.size() - pat().size(), pat().size(), pat()) == 0;
    }
    return false;
}
bool cond_matches(const FilterNode& c, const Json& md) { return cond_on_value(c, nested(md, c.field)); }
const char* op_name(FilterOp op) {
    switch (op) {
        case FilterOp::Eq: return "eq"; case FilterOp::Ne: return "ne"; case FilterOp::Gt: return "gt";
        case FilterOp::Gte: return "gte"; case FilterOp::Lt: return "lt"; case FilterOp::Lte: return "lte";
        case FilterOp::In: return "in"; case FilterOp::NotIn: return "notin"; case FilterOp::Contains: return "contains";
        case FilterOp::StartsWith: return "startswith"; case FilterOp::EndsWith: return "endswith"; case FilterOp::Exists: return "exists";
    }
    return "?";
}
}  // namespace

bool filter_parse(const std::string& in, FilterNode& out) {  // filter.rs:52-134
    std::string s = trim(in);
    if (contains(s, " OR ")) {
        std::vector<FilterNode> fs;
        for (auto& p : split(s, " OR ")) { FilterNode n; if (filter_parse(trim(p), n)) fs.push_back(std::move(n)); }
        if (fs.size() > 1) { out = FilterNode(); out.kind = FilterNode::Or; out.children = std::move(fs); return true; }
        if (fs.size() == 1) { out = std::move(fs[0]); return true; }
        return false;
    }
    bool has_and = contains(s, " AND ");
    bool has_comma = false;
    {
        int depth = 0;
        for (char c : s) {
            if (c == '[') depth++;
            else if (c == ']') depth--;
            else if (c == ',' && depth == 0) { has_comma = true; break; }
        }
    }
    if (has_and || has_comma) {
        std::vector<std::string> parts;
        if (has_and) parts = split(s, " AND ");
        else {
            std::string cur;
            int depth = 0;
            for (char c : s) {
                if (c == '[') { depth++; cur += c; }
                else if (c == ']') { depth--; cur += c; }
                else if (c == ',' && depth == 0) { parts.push_back(cur); cur.clear(); }
                else cur += c;
            }
            if (!cur.empty()) parts.push_back(cur);
        }
        std::vector<FilterNode> fs;
        for (auto& p : parts) { FilterNode n; if (parse_single(trim(p), n)) fs.push_back(std::move(n)); }
        if (fs.size() > 1) { out = FilterNode(); out.kind = FilterNode::And; out.children = std::move(fs); return true; }
        if (fs.size() == 1) { out = std::move(fs[0]); return true; }
        return false;
    }
    return parse_single(s, out);
}

bool filter_matches(const FilterNode& f, const Json& md) {  // filter.rs:319-325
    switch (f.kind) {
        case FilterNode::Condition: return cond_matches(f, md);
        case FilterNode::And: for (auto& c : f.children) if (!filter_matches(c, md)) return false; return true;
        case FilterNode::Or: for (auto& c : f.children) if (filter_matches(c, md)) return true; return false;
    }
    return false;
}

std::string filter_describe(const FilterNode& f) {
    std::string o;
    if (f.kind == FilterNode::Condition) {
        o += "{\"field\":"; dump_str(o, f.field);
        o += ",\"op\":\""; o += op_name(f.op); o += "\",\"value\":"; o += json_dump(f.value); o += "}";
        return o;
    }
    o += f.kind == FilterNode::And ? "{\"and\":[" : "{\"or\":[";
    for (size_t i = 0; i < f.children.size(); ++i) { if (i) o += ','; o += filter_describe(f.children[i]); }
    o += "]}";
    return o;
}

// ---- MetaColumns ---------------------------------------------------------------------------------------------
void MetaColumns::resize(size_t rows) {
    n = rows;
    for (auto& kv : cols) {
        kv.second.kind.resize(n, MetaColumn::Missing);
        kv.second.sid.resize(n, 0);
        if (!kv.second.num.empty()) kv.second.num.resize(n, 0.0);
    }
}

namespace {
void put_value(MetaColumns& mc, const std::string& path, size_t row, const Json& v) {
    MetaColumn& c = mc.cols[path];
    if (c.kind.size() < mc.n) { c.kind.resize(mc.n, MetaColumn::Missing); c.sid.resize(mc.n, 0); }
    switch (v.kind) {
        case Json::Null: c.kind[row] = MetaColumn::Null; break;
        case Json::Bool: c.kind[row] = v.b ? MetaColumn::True : MetaColumn::False; break;
        case Json::Num:
            if (c.num.size() < mc.n) c.num.resize(mc.n, 0.0);
            c.kind[row] = MetaColumn::Num; c.num[row] = v.num; break;
        case Json::Str: {
            auto it = c.dict_ids.find(v.str);
            if (it == c.dict_ids.end()) { it = c.dict_ids.emplace(v.str, (uint32_t)c.dict.size()).first; c.dict.push_back(v.str); }
            c.kind[row] = MetaColumn::Str; c.sid[row] = it->second; break;
        }
        default: c.kind[row] = MetaColumn::Other;
    }
}
// Every path `nested` can resolve: keys joined by '.', keys that themselves contain '.' are unreachable
// (filter.rs:376 splits the field on '.'), duplicate keys resolve as Json::get does.
void flatten(MetaColumns& mc, const std::string& prefix, bool top, size_t row, const Json& obj) {
    for (auto& kv : obj.obj) {
        if (kv.first.find('.') != std::string::npos) continue;
        const Json* v = obj.get(kv.first);
        const std::string path = top ? kv.first : prefix + "." + kv.first;
        put_value(mc, path, row, *v);
        if (v->kind == Json::Obj) flatten(mc, path, false, row, *v);
    }
}
}  // namespace

void MetaColumns::add_row(size_t row, const Json& metadata) {
    if (row >= n) resize(row + 1);
    if (metadata.kind == Json::Obj) flatten(*this, "", true, row, metadata);
}

void MetaColumns::eval(const FilterNode& f, std::vector<uint64_t>& mask) const {
    const size_t words = (n + 63) / 64;
    if (f.kind != FilterNode::Condition) {
        const bool is_and = f.kind == FilterNode::And;
        mask.assign(words, is_and ? ~0ull : 0ull);
        std::vector<uint64_t> m;
        for (auto& ch : f.children) {
            eval(ch, m);
            for (size_t w = 0; w < words; ++w) mask[w] = is_and ? (mask[w] & m[w]) : (mask[w] | m[w]);
        }
        if (words && (n & 63)) mask[words - 1] &= (1ull << (n & 63)) - 1ull;
        return;
    }
    mask.assign(words, 0ull);
    // a field path with an empty segment in the middle ("a..b") or one the flattening never produced is missing everywhere
    auto it = cols.find(f.field);
    const bool when_missing = cond_on_value(f, nullptr);
    if (it == cols.end()) {
        if (when_missing) {
            for (size_t w = 0; w < words; ++w) mask[w] = ~0ull;
            if (words && (n & 63)) mask[words - 1] &= (1ull << (n & 63)) - 1ull;
        }
        return;
    }
    const MetaColumn& c = it->second;
    // one evaluation per distinct non-numeric value, one per row for numbers
    Json probe;
    bool by_kind[7];
    by_kind[MetaColumn::Missing] = when_missing;
    probe.kind = Json::Null; by_kind[MetaColumn::Null] = cond_on_value(f, &probe);
    probe.kind = Json::Bool; probe.b = false; by_kind[MetaColumn::False] = cond_on_value(f, &probe);
    probe.b = true; by_kind[MetaColumn::True] = cond_on_value(f, &probe);
    probe = Json(); probe.kind = Json::Arr; by_kind[MetaColumn::Other] = cond_on_value(f, &probe);   // arrays/objects: only kind matters
    std::vector<uint8_t> str_ok(c.dict.size());
    probe = Json(); probe.kind = Json::Str;
    for (size_t i = 0; i < c.dict.size(); ++i) { probe.str = c.dict[i]; str_ok[i] = cond_on_value(f, &probe) ? 1 : 0; }
    probe = Json(); probe.kind = Json::Num;
    const size_t rows = std::min(n, c.kind.size());
    for (size_t i = 0; i < rows; ++i) {
        bool ok;
        const uint8_t k = c.kind[i];
        if (k == MetaColumn::Str) ok = str_ok[c.sid[i]] != 0;
        else if (k == MetaColumn::Num) { probe.num = c.num[i]; ok = cond_on_value(f, &probe); }
        else ok = by_kind[k];
        if (ok) mask[i >> 6] |= 1ull << (i & 63);
    }
    if (when_missing) for (size_t i = rows; i < n; ++i) mask[i >> 6] |= 1ull << (i & 63);
}

}  // namespace leann
