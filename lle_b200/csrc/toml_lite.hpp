// A small TOML reader for the v2 map format (src/core/parsing/toml/*.rs; schema in resources/lle_toml_schema.json).
//
// The reference deserialises with the `toml` crate (Cargo.toml: toml = "0.9", third-party, not vendored).  The v2 map
// schema only uses: comments, bare and quoted keys, integers, booleans, basic / literal strings (single- and multi-line),
// arrays (multi-line, trailing comma), inline tables, [table] and [[array-of-tables]] headers.  Floats, dates and dotted
// keys never appear in it; a document using them is reported as "not TOML" so that, like the reference
// (parsing/mod.rs:14-21), the caller falls back to the v1 grammar.
#pragma once
#include <cctype>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace lle {
namespace toml {

struct SyntaxError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Value {
    enum Kind { Int, Bool, Str, Array, Table } kind = Table;
    int64_t i = 0;
    bool b = false;
    std::string s;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> tbl;  // insertion order

    const Value* get(const std::string& key) const {
        for (const auto& kv : tbl)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    Value* get(const std::string& key) {
        for (auto& kv : tbl)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

class Parser {
  public:
    explicit Parser(const std::string& text) : s_(text) {}

    Value parse() {
        Value root;
        Value* cur = &root;
        for (;;) {
            skip_ws_comments_newlines();
            if (eof()) break;
            if (peek() == '[') {
                ++p_;
                const bool is_array = peek() == '[';
                if (is_array) ++p_;
                skip_ws();
                const std::string name = parse_key();
                skip_ws();
                expect(']');
                if (is_array) expect(']');
                end_of_line();
                Value* slot = root.get(name);
                if (is_array) {
                    if (!slot) {
                        Value arr;
                        arr.kind = Value::Array;
                        root.tbl.emplace_back(name, arr);
                        slot = &root.tbl.back().second;
                    }
                    if (slot->kind != Value::Array) throw SyntaxError("redefinition of " + name);
                    slot->arr.emplace_back();
                    cur = &slot->arr.back();
                } else {
                    if (slot) throw SyntaxError("duplicate table " + name);
                    root.tbl.emplace_back(name, Value());
                    cur = &root.tbl.back().second;
                }
                continue;
            }
            const std::string key = parse_key();
            skip_ws();
            expect('=');
            skip_ws();
            Value v = parse_value();
            end_of_line();
            if (cur->get(key)) throw SyntaxError("duplicate key " + key);
            cur->tbl.emplace_back(key, std::move(v));
        }
        return root;
    }

  private:
    const std::string& s_;
    size_t p_ = 0;

    bool eof() const { return p_ >= s_.size(); }
    char peek() const { return eof() ? '\0' : s_[p_]; }
    void expect(char c) {
        if (peek() != c) throw SyntaxError(std::string("expected '") + c + "'");
        ++p_;
    }
    void skip_ws() {
        while (!eof() && (s_[p_] == ' ' || s_[p_] == '\t')) ++p_;
    }
    void skip_comment() {
        if (peek() == '#')
            while (!eof() && s_[p_] != '\n') ++p_;
    }
    void skip_ws_comments_newlines() {
        for (;;) {
            skip_ws();
            skip_comment();
            if (!eof() && (s_[p_] == '\n' || s_[p_] == '\r')) { ++p_; continue; }
            break;
        }
    }
    void end_of_line() {
        skip_ws();
        skip_comment();
        if (eof()) return;
        if (s_[p_] == '\r') ++p_;
        if (eof()) return;
        if (s_[p_] != '\n') throw SyntaxError("expected end of line");
        ++p_;
    }
    std::string parse_key() {
        if (peek() == '"' || peek() == '\'') return parse_string().s;
        std::string k;
        while (!eof() && (std::isalnum((unsigned char)s_[p_]) || s_[p_] == '_' || s_[p_] == '-')) k.push_back(s_[p_++]);
        if (k.empty()) throw SyntaxError("expected a key");
        if (peek() == '.') throw SyntaxError("dotted keys are not supported");
        return k;
    }
    Value parse_string() {
        Value v;
        v.kind = Value::Str;
        const char q = s_[p_];
        const bool multi = s_.compare(p_, 3, std::string(3, q)) == 0;
        p_ += multi ? 3 : 1;
        if (multi) {  // a newline right after the opening delimiter is trimmed
            if (peek() == '\r') ++p_;
            if (peek() == '\n') ++p_;
        }
        for (;;) {
            if (eof()) throw SyntaxError("unterminated string");
            if (multi ? s_.compare(p_, 3, std::string(3, q)) == 0 : s_[p_] == q) {
                p_ += multi ? 3 : 1;
                // up to two extra quotes may sit right before a multi-line closing delimiter
                while (multi && peek() == q) { v.s.push_back(q); ++p_; }
                return v;
            }
            char c = s_[p_++];
            if (!multi && c == '\n') throw SyntaxError("newline in a single-line string");
            if (q == '"' && c == '\\') {
                if (eof()) throw SyntaxError("bad escape");
                char e = s_[p_++];
                switch (e) {
                    case 'n': v.s.push_back('\n'); break;
                    case 't': v.s.push_back('\t'); break;
                    case 'r': v.s.push_back('\r'); break;
                    case '"': v.s.push_back('"'); break;
                    case '\\': v.s.push_back('\\'); break;
                    case '\n':  // line-ending backslash: skip the whitespace that follows
                        while (!eof() && std::isspace((unsigned char)s_[p_])) ++p_;
                        break;
                    default: throw SyntaxError("unsupported escape");
                }
            } else {
                v.s.push_back(c);
            }
        }
    }
    Value parse_value() {
        const char c = peek();
        if (c == '"' || c == '\'') return parse_string();
        if (c == '[') {
            ++p_;
            Value v;
            v.kind = Value::Array;
            for (;;) {
                skip_ws_comments_newlines();
                if (peek() == ']') { ++p_; return v; }
                v.arr.push_back(parse_value());
                skip_ws_comments_newlines();
                if (peek() == ',') { ++p_; continue; }
                skip_ws_comments_newlines();
                expect(']');
                return v;
            }
        }
        if (c == '{') {
            ++p_;
            Value v;
            v.kind = Value::Table;
            skip_ws();
            if (peek() == '}') { ++p_; return v; }
            for (;;) {
                skip_ws();
                const std::string key = parse_key();
                skip_ws();
                expect('=');
                skip_ws();
                Value item = parse_value();
                if (v.get(key)) throw SyntaxError("duplicate key " + key);
                v.tbl.emplace_back(key, std::move(item));
                skip_ws();
                if (peek() == ',') { ++p_; continue; }
                expect('}');
                return v;
            }
        }
        if (s_.compare(p_, 4, "true") == 0) { p_ += 4; Value v; v.kind = Value::Bool; v.b = true; return v; }
        if (s_.compare(p_, 5, "false") == 0) { p_ += 5; Value v; v.kind = Value::Bool; v.b = false; return v; }
        if (c == '+' || c == '-' || std::isdigit((unsigned char)c)) {
            size_t q = p_;
            if (s_[q] == '+' || s_[q] == '-') ++q;
            int64_t n = 0;
            size_t digits = 0;
            while (q < s_.size() && (std::isdigit((unsigned char)s_[q]) || s_[q] == '_')) {
                if (s_[q] != '_') { n = n * 10 + (s_[q] - '0'); ++digits; if (n > (int64_t)1 << 40) throw SyntaxError("integer too large"); }
                ++q;
            }
            if (digits == 0) throw SyntaxError("bad number");
            if (q < s_.size() && (s_[q] == '.' || s_[q] == 'e' || s_[q] == 'E' || s_[q] == ':' || s_[q] == 'T'))
                throw SyntaxError("floats and dates are not supported");
            Value v;
            v.kind = Value::Int;
            v.i = s_[p_] == '-' ? -n : n;
            p_ = q;
            return v;
        }
        throw SyntaxError("unexpected character in a value");
    }
};

}  // namespace toml
}  // namespace lle
