// json_min.h -- a small DOM JSON parser (RFC 8259) for the glTF loader.  No third-party JSON library is
// available in this image; the reference gets its JSON through the `gltf` crate (Cargo.toml:22-24).
#pragma once
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace rtjson {

struct Value;
typedef std::shared_ptr<Value> ValuePtr;

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;  // insertion order kept

    bool is_object() const { return kind == Object; }
    bool is_array() const { return kind == Array; }
    bool is_number() const { return kind == Number; }
    const Value* get(const char* key) const {
        if (kind != Object) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    bool has(const char* key) const { return get(key) != nullptr; }
    size_t size() const { return kind == Array ? arr.size() : (kind == Object ? obj.size() : 0); }
    const Value& at(size_t i) const {
        if (kind != Array || i >= arr.size()) throw std::runtime_error("json: array index out of range");
        return *arr[i];
    }
    double number(const char* key, double dflt) const { const Value* v = get(key); return (v && v->kind == Number) ? v->num : dflt; }
    long long integer(const char* key, long long dflt) const { const Value* v = get(key); return (v && v->kind == Number) ? (long long)v->num : dflt; }
    // Non-negative integral number below 2^53, or -1 when the key is absent (absent_ok) -- anything else (negative, fractional, NaN,
    // huge, not a number) is -2: the caller rejects the file instead of casting an out-of-range double (undefined behaviour).
    long long index(const char* key) const { const Value* v = get(key); return v ? v->as_index() : -1; }
    long long as_index() const {
        if (kind != Number || !(num >= 0.0) || !(num < 9007199254740992.0) || num != (double)(long long)num) return -2;
        return (long long)num;
    }
    std::string string(const char* key, const std::string& dflt) const { const Value* v = get(key); return (v && v->kind == String) ? v->str : dflt; }
};

class Parser {
public:
    explicit Parser(const std::string& text) : s_(text), p_(0) {}
    ValuePtr parse() {
        ValuePtr v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t p_;

    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("json: ") + what + " at byte " + std::to_string(p_)); }
    void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_; }
    bool lit(const char* w) { size_t n = std::strlen(w); if (s_.compare(p_, n, w) == 0) { p_ += n; return true; } return false; }

    int depth_ = 0;
    struct DepthGuard { int& d; explicit DepthGuard(int& x) : d(x) { ++d; } ~DepthGuard() { --d; } };
    ValuePtr value() {
        DepthGuard guard(depth_);
        if (depth_ > 256) fail("JSON nested deeper than 256 levels");      // recursion bound: a hostile file must not overflow the stack
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        ValuePtr v = std::make_shared<Value>();
        char c = s_[p_];
        if (c == '{') {
            v->kind = Value::Object; ++p_; ws();
            if (p_ < s_.size() && s_[p_] == '}') { ++p_; return v; }
            for (;;) {
                ws();
                if (p_ >= s_.size() || s_[p_] != '"') fail("expected object key");
                std::string k = string_body();
                ws();
                if (p_ >= s_.size() || s_[p_] != ':') fail("expected ':'");
                ++p_;
                v->obj.emplace_back(k, value());
                ws();
                if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
                if (p_ < s_.size() && s_[p_] == '}') { ++p_; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v->kind = Value::Array; ++p_; ws();
            if (p_ < s_.size() && s_[p_] == ']') { ++p_; return v; }
            for (;;) {
                v->arr.push_back(value());
                ws();
                if (p_ < s_.size() && s_[p_] == ',') { ++p_; continue; }
                if (p_ < s_.size() && s_[p_] == ']') { ++p_; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v->kind = Value::String; v->str = string_body();
        } else if (lit("true")) { v->kind = Value::Bool; v->b = true; }
        else if (lit("false")) { v->kind = Value::Bool; v->b = false; }
        else if (lit("null")) { v->kind = Value::Null; }
        else {
            const char* begin = s_.c_str() + p_;
            char* end = nullptr;
            double d = std::strtod(begin, &end);   // correctly rounded decimal -> f64
            if (end == begin) fail("unexpected character");
            p_ += (size_t)(end - begin);
            v->kind = Value::Number; v->num = d;
        }
        return v;
    }

    static void put_utf8(std::string& out, unsigned cp) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    }
    unsigned hex4() {
        if (p_ + 4 > s_.size()) fail("bad \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = s_[p_++]; v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string string_body() {
        std::string out;
        ++p_;  // opening quote
        for (;;) {
            if (p_ >= s_.size()) fail("unterminated string");
            char c = s_[p_++];
            if (c == '"') break;
            if (c != '\\') { out.push_back(c); continue; }
            if (p_ >= s_.size()) fail("unterminated escape");
            char e = s_[p_++];
            switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF && p_ + 1 < s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
                        p_ += 2;
                        unsigned lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    put_utf8(out, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
        return out;
    }
};

inline ValuePtr parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace rtjson
