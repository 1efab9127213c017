// Minimal JSON reader for the metric config files (scripts/*.json).  The reference parses
// them with nlohmann::json (/root/reference/content_manager.cpp:76, /root/reference/metric.hpp:359).
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mini_json {

struct Value {
    enum Type { Null, Bool, Number, String, Array, Object } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;  // insertion ordered
    const Value* find(const std::string& key) const {
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s) {}
    Value parse() {
        // tolerate a UTF-8 BOM
        if (s_.size() >= 3 && (unsigned char)s_[0] == 0xEF && (unsigned char)s_[1] == 0xBB && (unsigned char)s_[2] == 0xBF) p_ = 3;
        Value v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

private:
    const std::string& s_;
    size_t p_ = 0;
    [[noreturn]] void fail(const std::string& m) { throw std::runtime_error("JSON parse error at offset " + std::to_string(p_) + ": " + m); }
    void ws() { while (p_ < s_.size() && isspace((unsigned char)s_[p_])) p_++; }
    Value value() {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        char c = s_[p_];
        Value v;
        if (c == '{') {
            v.type = Value::Object;
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
                v.obj.emplace_back(k, value());
                ws();
                if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
                if (p_ < s_.size() && s_[p_] == '}') { p_++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.type = Value::Array;
            p_++;
            ws();
            if (p_ < s_.size() && s_[p_] == ']') { p_++; return v; }
            while (true) {
                v.arr.push_back(value());
                ws();
                if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
                if (p_ < s_.size() && s_[p_] == ']') { p_++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.type = Value::String;
            v.str = string();
        } else if (s_.compare(p_, 4, "true") == 0) { v.type = Value::Bool; v.b = true; p_ += 4; }
        else if (s_.compare(p_, 5, "false") == 0) { v.type = Value::Bool; v.b = false; p_ += 5; }
        else if (s_.compare(p_, 4, "null") == 0) { p_ += 4; }
        else {
            char* end;
            v.type = Value::Number;
            v.num = strtod(s_.c_str() + p_, &end);
            if (end == s_.c_str() + p_) fail("unexpected character");
            p_ = end - s_.c_str();
        }
        return v;
    }
    std::string string() {
        std::string out;
        p_++;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if (c == '\\' && p_ < s_.size()) {
                char e = s_[p_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        if (p_ + 4 > s_.size()) fail("bad \\u escape");
                        unsigned cp = (unsigned)strtoul(s_.substr(p_, 4).c_str(), nullptr, 16);
                        p_ += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += e;
                }
            } else out += c;
        }
        if (p_ >= s_.size()) fail("unterminated string");
        p_++;
        return out;
    }
};

inline Value parse(const std::string& s) { return Parser(s).parse(); }

}  // namespace mini_json
