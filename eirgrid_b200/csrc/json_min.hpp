// json_min.hpp — minimal JSON reader/writer for the three files the path touches:
// settlements.json, coastline_points.json (inputs of the reference's loaders) and the weights
// checkpoints (SerializableWeights, ai/learning/serialization.rs:37-51). Host-only.
#pragma once
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <string>
#include <vector>
#include <utility>
#include <stdexcept>

namespace egjson {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;

  const Value* get(const char* key) const {
    if (kind != Object) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool is_null() const { return kind == Null; }
};

class Parser {
 public:
  explicit Parser(const std::string& text) : s_(text), i_(0), depth_(0) {}
  Value parse() {
    Value v = value();
    ws();
    if (i_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t i_;
  int depth_;  // nesting of the value being parsed; serde_json's recursion limit is 128
  struct Nest {
    Parser& p;
    explicit Nest(Parser& q) : p(q) { if (++p.depth_ > 128) p.fail("recursion limit exceeded"); }
    ~Nest() { p.depth_--; }
  };
  [[noreturn]] void fail(const char* what) const {
    throw std::runtime_error(std::string("JSON parse error at byte ") + std::to_string(i_) + ": " + what);
  }
  void ws() {
    while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) i_++;
  }
  Value value() {
    ws();
    if (i_ >= s_.size()) fail("unexpected end");
    char c = s_[i_];
    Value v;
    if (c == '{') {
      Nest nest(*this);
      v.kind = Value::Object;
      i_++;
      ws();
      if (i_ < s_.size() && s_[i_] == '}') { i_++; return v; }
      for (;;) {
        ws();
        if (i_ >= s_.size() || s_[i_] != '"') fail("expected key");
        std::string k = string();
        ws();
        if (i_ >= s_.size() || s_[i_] != ':') fail("expected ':'");
        i_++;
        v.obj.emplace_back(std::move(k), value());
        ws();
        if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
        if (i_ < s_.size() && s_[i_] == '}') { i_++; break; }
        fail("expected ',' or '}'");
      }
    } else if (c == '[') {
      Nest nest(*this);
      v.kind = Value::Array;
      i_++;
      ws();
      if (i_ < s_.size() && s_[i_] == ']') { i_++; return v; }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
        if (i_ < s_.size() && s_[i_] == ']') { i_++; break; }
        fail("expected ',' or ']'");
      }
    } else if (c == '"') {
      v.kind = Value::String;
      v.str = string();
    } else if (c == 't' && s_.compare(i_, 4, "true") == 0) {
      v.kind = Value::Bool; v.b = true; i_ += 4;
    } else if (c == 'f' && s_.compare(i_, 5, "false") == 0) {
      v.kind = Value::Bool; v.b = false; i_ += 5;
    } else if (c == 'n' && s_.compare(i_, 4, "null") == 0) {
      i_ += 4;
    } else {
      // JSON number grammar only (strtod alone would also take "NaN", "inf", hex floats and a leading '+')
      size_t j = i_;
      if (j < s_.size() && s_[j] == '-') j++;
      const size_t int_start = j;
      while (j < s_.size() && s_[j] >= '0' && s_[j] <= '9') j++;
      if (j == int_start) fail("expected value");
      if (j < s_.size() && s_[j] == '.') {
        const size_t frac_start = ++j;
        while (j < s_.size() && s_[j] >= '0' && s_[j] <= '9') j++;
        if (j == frac_start) fail("bad number");
      }
      if (j < s_.size() && (s_[j] == 'e' || s_[j] == 'E')) {
        j++;
        if (j < s_.size() && (s_[j] == '+' || s_[j] == '-')) j++;
        const size_t exp_start = j;
        while (j < s_.size() && s_[j] >= '0' && s_[j] <= '9') j++;
        if (j == exp_start) fail("bad number");
      }
      v.num = std::strtod(s_.substr(i_, j - i_).c_str(), nullptr);  // correctly rounded
      v.kind = Value::Number;
      i_ = j;
    }
    return v;
  }
  std::string string() {
    std::string out;
    i_++;  // opening quote
    while (i_ < s_.size() && s_[i_] != '"') {
      char c = s_[i_++];
      if (c == '\\') {
        if (i_ >= s_.size()) fail("bad escape");
        char e = s_[i_++];
        switch (e) {
          case 'n': out.push_back('\n'); break;
          case 't': out.push_back('\t'); break;
          case 'r': out.push_back('\r'); break;
          case 'b': out.push_back('\b'); break;
          case 'f': out.push_back('\f'); break;
          case 'u': {
            if (i_ + 4 > s_.size()) fail("bad \\u escape");
            unsigned cp = (unsigned)std::strtoul(s_.substr(i_, 4).c_str(), nullptr, 16);
            i_ += 4;
            if (cp < 0x80) out.push_back((char)cp);
            else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
            else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
            break;
          }
          default: out.push_back(e); break;
        }
      } else {
        out.push_back(c);
      }
    }
    if (i_ >= s_.size()) fail("unterminated string");
    i_++;  // closing quote
    return out;
  }
};

inline bool read_file(const char* path, std::string* out) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out->resize((size_t)(n > 0 ? n : 0));
  size_t got = n > 0 ? std::fread(&(*out)[0], 1, (size_t)n, f) : 0;
  std::fclose(f);
  return got == (size_t)(n > 0 ? n : 0);
}

// A double as serde_json prints it (ryu's `format64`): the shortest digit string that parses back to exactly `v`, laid out as
//   digits followed by ".0" for integral values below 1e16      1234e7  -> 12340000000.0
//   a decimal point inside the digits                           1234e-2 -> 12.34
//   leading zeros for values down to 1e-5                       1234e-8 -> 0.00001234
//   scientific otherwise, no '+' and no padding in the exponent 1e-6, 1.234e-7, 1.2e16
// NaN and the infinities become null like serde_json's.
inline std::string fmt_double(double v) {
  if (v != v || v == HUGE_VAL || v == -HUGE_VAL) return "null";
  if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
  char buf[40];
  // std::to_chars without a precision: the shortest digits that round-trip (libstdc++ implements it with Ryu, like serde_json)
  const std::to_chars_result tc = std::to_chars(buf, buf + sizeof(buf) - 1, v, std::chars_format::scientific);
  *tc.ptr = 0;
  // buf = [-]d[.ddd]e[+-]xx
  std::string text(buf), digits, out;
  const bool negative = text[0] == '-';
  const size_t epos = text.find('e');
  for (size_t i = negative ? 1 : 0; i < epos; i++)
    if (text[i] != '.') digits += text[i];
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int exp10 = std::atoi(text.c_str() + epos + 1);
  const int length = (int)digits.size();
  const int kk = exp10 + 1;           // position of the decimal point relative to the first digit
  const int k = kk - length;          // value = digits * 10^k
  if (negative) out += '-';
  if (k >= 0 && kk <= 16) {
    out += digits + std::string((size_t)k, '0') + ".0";
  } else if (kk > 0 && kk <= 16) {
    out += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
  } else if (kk > -5 && kk <= 0) {
    out += "0." + std::string((size_t)(-kk), '0') + digits;
  } else {
    out += digits.substr(0, 1);
    if (length > 1) out += "." + digits.substr(1);
    out += "e" + std::to_string(kk - 1);
  }
  return out;
}

}  // namespace egjson
