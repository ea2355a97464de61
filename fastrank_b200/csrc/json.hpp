// json.hpp -- the small JSON reader/writer behind the JSON-over-C-ABI surface.
//
// The reference marshals everything through serde_json (json_api.rs:13-34, ffi.rs:76-80).
// What matters for a drop-in: objects keep insertion order, u64 integers survive exactly
// (the `seed` field, coordinate_ascent.rs:17), floats are written in shortest round-trip
// form and always look like floats ("1.0", not "1"), non-finite floats become null.
#pragma once
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace frb {
namespace json {

struct Value;
using Member = std::pair<std::string, Value>;

struct Value {
    enum Kind { Null, Bool, Int, UInt, Double, String, Array, Object };
    Kind kind = Null;
    bool b = false;
    int64_t i = 0;
    uint64_t u = 0;
    double d = 0.0;
    std::string s;
    std::vector<Value> arr;
    std::vector<Member> obj;

    Value() = default;
    static Value null() { return Value(); }
    static Value boolean(bool v) {
        Value x;
        x.kind = Bool;
        x.b = v;
        return x;
    }
    static Value integer(int64_t v) {
        Value x;
        x.kind = Int;
        x.i = v;
        return x;
    }
    static Value uinteger(uint64_t v) {
        Value x;
        x.kind = UInt;
        x.u = v;
        return x;
    }
    static Value number(double v) {
        Value x;
        x.kind = Double;
        x.d = v;
        return x;
    }
    static Value string(std::string v) {
        Value x;
        x.kind = String;
        x.s = std::move(v);
        return x;
    }
    static Value array() {
        Value x;
        x.kind = Array;
        return x;
    }
    static Value object() {
        Value x;
        x.kind = Object;
        return x;
    }

    bool is_null() const { return kind == Null; }
    bool is_number() const { return kind == Int || kind == UInt || kind == Double; }
    double as_double() const {
        if (kind == Int) return (double)i;
        if (kind == UInt) return (double)u;
        return d;
    }
    const Value *find(const std::string &key) const {
        for (const Member &m : obj)
            if (m.first == key) return &m.second;
        return nullptr;
    }
    Value &set(const std::string &key, Value v) {
        obj.emplace_back(key, std::move(v));
        return obj.back().second;
    }
    void push(Value v) { arr.push_back(std::move(v)); }
};

class ParseError : public std::runtime_error {
   public:
    explicit ParseError(const std::string &m) : std::runtime_error(m) {}
};

class Parser {
   public:
    explicit Parser(const std::string &text) : p_(text.data()), end_(text.data() + text.size()), begin_(text.data()) {}
    Value parse_document() {
        Value v = parse_value(0);
        skip_ws();
        if (p_ != end_) error("trailing characters");
        return v;
    }

   private:
    const char *p_, *end_, *begin_;

    [[noreturn]] void error(const std::string &what) const {
        size_t line = 1, col = 1;
        for (const char *q = begin_; q < p_ && q < end_; ++q) {
            if (*q == '\n') {
                ++line;
                col = 1;
            } else {
                ++col;
            }
        }
        throw ParseError(what + " at line " + std::to_string(line) + " column " + std::to_string(col));
    }
    void skip_ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
    }
    bool consume(const char *lit) {
        size_t n = strlen(lit);
        if ((size_t)(end_ - p_) >= n && memcmp(p_, lit, n) == 0) {
            p_ += n;
            return true;
        }
        return false;
    }
    Value parse_value(int depth) {
        if (depth > 4096) error("recursion limit exceeded");
        skip_ws();
        if (p_ >= end_) error("EOF while parsing a value");
        char c = *p_;
        if (c == '{') return parse_object(depth);
        if (c == '[') return parse_array(depth);
        if (c == '"') return Value::string(parse_string());
        if (consume("null")) return Value::null();
        if (consume("true")) return Value::boolean(true);
        if (consume("false")) return Value::boolean(false);
        if (c == '-' || (c >= '0' && c <= '9')) return parse_number();
        if (consume("NaN")) return Value::number(NAN);  // python json.dumps emits these
        if (consume("Infinity")) return Value::number(INFINITY);
        error("expected value");
    }
    Value parse_number() {
        const char *start = p_;
        bool is_float = false;
        if (*p_ == '-') {
            ++p_;
            if (consume("Infinity")) return Value::number(-INFINITY);
        }
        while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-')) {
            if (*p_ == '.' || *p_ == 'e' || *p_ == 'E') is_float = true;
            ++p_;
        }
        std::string tok(start, p_);
        if (tok.empty() || tok == "-") error("invalid number");
        if (!is_float) {
            if (tok[0] == '-') {
                int64_t v = 0;
                auto r = std::from_chars(tok.data(), tok.data() + tok.size(), v);
                if (r.ec == std::errc() && r.ptr == tok.data() + tok.size()) return Value::integer(v);
            } else {
                uint64_t v = 0;
                auto r = std::from_chars(tok.data(), tok.data() + tok.size(), v);
                if (r.ec == std::errc() && r.ptr == tok.data() + tok.size()) return Value::uinteger(v);
            }
        }
        char *endp = nullptr;
        double d = strtod(tok.c_str(), &endp);
        if (endp != tok.c_str() + tok.size()) error("invalid number");
        return Value::number(d);
    }
    static void append_utf8(std::string &out, uint32_t cp) {
        if (cp < 0x80) {
            out.push_back((char)cp);
        } else if (cp < 0x800) {
            out.push_back((char)(0xC0 | (cp >> 6)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else if (cp < 0x10000) {
            out.push_back((char)(0xE0 | (cp >> 12)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        } else {
            out.push_back((char)(0xF0 | (cp >> 18)));
            out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back((char)(0x80 | (cp & 0x3F)));
        }
    }
    uint32_t parse_hex4() {
        if (end_ - p_ < 4) error("EOF in \\u escape");
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) {
            char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else error("invalid \\u escape");
        }
        return v;
    }
    std::string parse_string() {
        ++p_;  // opening quote
        std::string out;
        for (;;) {
            if (p_ >= end_) error("EOF while parsing a string");
            char c = *p_++;
            if (c == '"') break;
            if (c != '\\') {
                out.push_back(c);
                continue;
            }
            if (p_ >= end_) error("EOF in escape");
            char e = *p_++;
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
                    uint32_t cp = parse_hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        p_ += 2;
                        uint32_t lo = parse_hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    append_utf8(out, cp);
                    break;
                }
                default: error("invalid escape");
            }
        }
        return out;
    }
    Value parse_array(int depth) {
        ++p_;
        Value v = Value::array();
        skip_ws();
        if (p_ < end_ && *p_ == ']') {
            ++p_;
            return v;
        }
        for (;;) {
            v.arr.push_back(parse_value(depth + 1));
            skip_ws();
            if (p_ >= end_) error("EOF while parsing a list");
            if (*p_ == ',') {
                ++p_;
                continue;
            }
            if (*p_ == ']') {
                ++p_;
                return v;
            }
            error("expected `,` or `]`");
        }
    }
    Value parse_object(int depth) {
        ++p_;
        Value v = Value::object();
        skip_ws();
        if (p_ < end_ && *p_ == '}') {
            ++p_;
            return v;
        }
        for (;;) {
            skip_ws();
            if (p_ >= end_ || *p_ != '"') error("key must be a string");
            std::string key = parse_string();
            skip_ws();
            if (p_ >= end_ || *p_ != ':') error("expected `:`");
            ++p_;
            Value val = parse_value(depth + 1);
            v.obj.emplace_back(std::move(key), std::move(val));
            skip_ws();
            if (p_ >= end_) error("EOF while parsing an object");
            if (*p_ == ',') {
                ++p_;
                continue;
            }
            if (*p_ == '}') {
                ++p_;
                return v;
            }
            error("expected `,` or `}`");
        }
    }
};

inline Value parse(const std::string &text) { return Parser(text).parse_document(); }

inline void write_string(std::string &out, const std::string &s) {
    out.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            default:
                if (c < 0x20) {
                    char buf[8];
                    snprintf(buf, sizeof buf, "\\u%04x", c);
                    out += buf;
                } else {
                    out.push_back((char)c);
                }
        }
    }
    out.push_back('"');
}

inline void write_double(std::string &out, double v) {
    if (!std::isfinite(v)) {
        out += "null";
        return;
    }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    std::string tok(buf, r.ptr);
    if (tok.find_first_of(".eE") == std::string::npos) tok += ".0";
    out += tok;
}

inline void write(std::string &out, const Value &v) {
    switch (v.kind) {
        case Value::Null: out += "null"; break;
        case Value::Bool: out += v.b ? "true" : "false"; break;
        case Value::Int: out += std::to_string(v.i); break;
        case Value::UInt: out += std::to_string(v.u); break;
        case Value::Double: write_double(out, v.d); break;
        case Value::String: write_string(out, v.s); break;
        case Value::Array: {
            out.push_back('[');
            for (size_t k = 0; k < v.arr.size(); ++k) {
                if (k) out.push_back(',');
                write(out, v.arr[k]);
            }
            out.push_back(']');
            break;
        }
        case Value::Object: {
            out.push_back('{');
            for (size_t k = 0; k < v.obj.size(); ++k) {
                if (k) out.push_back(',');
                write_string(out, v.obj[k].first);
                out.push_back(':');
                write(out, v.obj[k].second);
            }
            out.push_back('}');
            break;
        }
    }
}

inline std::string dump(const Value &v) {
    std::string out;
    write(out, v);
    return out;
}

}  // namespace json
}  // namespace frb
