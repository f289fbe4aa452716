// json_min.h -- a small recursive-descent JSON reader for scene files.
// The reference reads scenes with nlohmann::json 2.0.7 (include/nlohmann/json.hpp:4) and only ever
// uses: object lookup by key (at / find), arrays, strings, booleans and numbers converted to float
// (scene.cpp:189-246, main.cpp:65-69).  Numbers are parsed with strtod, like nlohmann's parser, so
// `(float)value` rounds identically.
#ifndef MCRT_JSON_MIN_H
#define MCRT_JSON_MIN_H

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mcrt_json {

struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    bool is_array() const { return type == Array; }
    bool is_object() const { return type == Object; }
    const Value* find(const std::string& key) const
    {
        if (type != Object) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const Value& at(const std::string& key) const
    {
        const Value* v = find(key);
        if (!v) throw std::out_of_range("key '" + key + "' not found");
        return *v;
    }
    const Value& at(size_t i) const
    {
        if (type != Array || i >= arr.size()) throw std::out_of_range("array index out of range");
        return arr[i];
    }
    double number() const
    {
        if (type == Number) return num;
        if (type == Bool) return b ? 1.0 : 0.0;
        throw std::domain_error("type must be number");
    }
    float as_float() const { return (float)number(); }
    bool as_bool() const
    {
        if (type == Bool) return b;
        if (type == Number) return num != 0.0;
        throw std::domain_error("type must be boolean");
    }
    const std::string& as_string() const
    {
        if (type != String) throw std::domain_error("type must be string");
        return str;
    }
};

class Parser {
public:
    explicit Parser(const std::string& text) : s_(text), p_(0) {}
    Value parse()
    {
        Value v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t p_;
    [[noreturn]] void fail(const char* what) const
    {
        throw std::invalid_argument(std::string("parse error at byte ") + std::to_string(p_) + ": " + what);
    }
    void ws()
    {
        while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) p_++;
    }
    Value value()
    {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        const char c = s_[p_];
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') { Value v; v.type = Value::String; v.str = string(); return v; }
        if (c == 't' || c == 'f' || c == 'n') return literal();
        return number();
    }
    Value literal()
    {
        Value v;
        if (s_.compare(p_, 4, "true") == 0) { v.type = Value::Bool; v.b = true; p_ += 4; }
        else if (s_.compare(p_, 5, "false") == 0) { v.type = Value::Bool; v.b = false; p_ += 5; }
        else if (s_.compare(p_, 4, "null") == 0) { v.type = Value::Null; p_ += 4; }
        else fail("invalid literal");
        return v;
    }
    Value number()
    {
        const char* begin = s_.c_str() + p_;
        char* end = nullptr;
        const double d = strtod(begin, &end);
        if (end == begin) fail("invalid number");
        p_ += (size_t)(end - begin);
        Value v; v.type = Value::Number; v.num = d;
        return v;
    }
    std::string string()
    {
        std::string out;
        p_++;   // opening quote
        while (true) {
            if (p_ >= s_.size()) fail("unterminated string");
            const char c = s_[p_++];
            if (c == '"') break;
            if (c == '\\') {
                if (p_ >= s_.size()) fail("bad escape");
                const char e = s_[p_++];
                switch (e) {
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case '/': out += '/'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'n': out += '\n'; break;
                    case 'r': out += '\r'; break;
                    case 't': out += '\t'; break;
                    case 'u': {
                        if (p_ + 4 > s_.size()) fail("bad \\u escape");
                        const unsigned cp = (unsigned)strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
                        p_ += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: fail("bad escape");
                }
            } else out += c;
        }
        return out;
    }
    Value array()
    {
        Value v; v.type = Value::Array;
        p_++;
        ws();
        if (p_ < s_.size() && s_[p_] == ']') { p_++; return v; }
        while (true) {
            v.arr.push_back(value());
            ws();
            if (p_ >= s_.size()) fail("unterminated array");
            if (s_[p_] == ',') { p_++; continue; }
            if (s_[p_] == ']') { p_++; break; }
            fail("expected ',' or ']'");
        }
        return v;
    }
    Value object()
    {
        Value v; v.type = Value::Object;
        p_++;
        ws();
        if (p_ < s_.size() && s_[p_] == '}') { p_++; return v; }
        while (true) {
            ws();
            if (p_ >= s_.size() || s_[p_] != '"') fail("expected string key");
            std::string k = string();
            ws();
            if (p_ >= s_.size() || s_[p_] != ':') fail("expected ':'");
            p_++;
            v.obj.emplace_back(std::move(k), value());
            ws();
            if (p_ >= s_.size()) fail("unterminated object");
            if (s_[p_] == ',') { p_++; continue; }
            if (s_[p_] == '}') { p_++; break; }
            fail("expected ',' or '}'");
        }
        return v;
    }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace mcrt_json
#endif
