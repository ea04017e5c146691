// Minimal JSON DOM parser for the scene files (stands in for serde_json, which the reference's
// loader uses: src/world/mod.rs:46-49).  Objects keep insertion order; unknown keys are simply
// never looked up, matching serde's default of ignoring them.
#pragma once
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
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
    std::vector<std::pair<std::string, ValuePtr>> obj;

    const Value* find(const std::string& key) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj)
            if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    const Value& at(const std::string& key) const {
        const Value* v = find(key);
        if (!v) throw std::runtime_error("missing field `" + key + "`");
        return *v;
    }
    double as_number() const {
        if (kind != Number) throw std::runtime_error("expected a number");
        return num;
    }
    const std::string& as_string() const {
        if (kind != String) throw std::runtime_error("expected a string");
        return str;
    }
    bool as_bool() const {
        if (kind != Bool) throw std::runtime_error("expected a boolean");
        return b;
    }
};

class Parser {
public:
    explicit Parser(const char* text) : p_(text) {}
    ValuePtr parse() {
        ValuePtr v = value();
        ws();
        if (*p_) fail("trailing characters");
        return v;
    }

private:
    const char* p_;
    [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("JSON: ") + what); }
    void ws() {
        while (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r') p_++;
    }
    ValuePtr value() {
        ws();
        auto v = std::make_shared<Value>();
        char c = *p_;
        if (c == '{') {
            v->kind = Value::Object;
            p_++;
            ws();
            if (*p_ == '}') { p_++; return v; }
            for (;;) {
                ws();
                if (*p_ != '"') fail("expected object key");
                std::string key = string();
                ws();
                if (*p_ != ':') fail("expected ':'");
                p_++;
                v->obj.emplace_back(key, value());
                ws();
                if (*p_ == ',') { p_++; continue; }
                if (*p_ == '}') { p_++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v->kind = Value::Array;
            p_++;
            ws();
            if (*p_ == ']') { p_++; return v; }
            for (;;) {
                v->arr.push_back(value());
                ws();
                if (*p_ == ',') { p_++; continue; }
                if (*p_ == ']') { p_++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v->kind = Value::String;
            v->str = string();
        } else if (!strncmp(p_, "true", 4)) {
            v->kind = Value::Bool; v->b = true; p_ += 4;
        } else if (!strncmp(p_, "false", 5)) {
            v->kind = Value::Bool; v->b = false; p_ += 5;
        } else if (!strncmp(p_, "null", 4)) {
            p_ += 4;
        } else if (c == '-' || (c >= '0' && c <= '9')) {
            char* end = nullptr;
            v->kind = Value::Number;
            v->num = strtod(p_, &end);  // correctly rounded, like serde_json's float parsing
            if (end == p_) fail("bad number");
            p_ = end;
        } else {
            fail("unexpected character");
        }
        return v;
    }
    std::string string() {
        std::string s;
        p_++;  // opening quote
        while (*p_ && *p_ != '"') {
            if (*p_ == '\\') {
                p_++;
                switch (*p_) {
                    case 'n': s += '\n'; break;
                    case 't': s += '\t'; break;
                    case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break;
                    case 'u': {  // keep BMP code points as UTF-8
                        unsigned cp = 0;
                        for (int i = 1; i <= 4; i++) {
                            char h = p_[i];
                            cp <<= 4;
                            if (h >= '0' && h <= '9') cp |= h - '0';
                            else if (h >= 'a' && h <= 'f') cp |= h - 'a' + 10;
                            else if (h >= 'A' && h <= 'F') cp |= h - 'A' + 10;
                            else fail("bad \\u escape");
                        }
                        p_ += 4;
                        if (cp < 0x80) s += (char)cp;
                        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
                        else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: s += *p_;
                }
                p_++;
            } else {
                s += *p_++;
            }
        }
        if (*p_ != '"') fail("unterminated string");
        p_++;
        return s;
    }
};

inline ValuePtr parse(const std::string& text) { return Parser(text.c_str()).parse(); }

}  // namespace rtjson
