// A small recursive-descent JSON reader shared by the rig description (RIG_JSON.md; the reference uses folly::parseJson,
// Camera.cpp:243-254) and the ISP configuration (res/config/isp/*.json; the reference uses the vendored supereasyjson, numbers
// through strtod like here): objects, arrays, numbers, strings, literals.
#pragma once
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace s360 {
namespace jsonmini {
struct Json {
  enum Kind { Null, Number, String, Array, Object, Bool } kind = Null;
  double number = 0;
  std::string text;
  std::vector<std::shared_ptr<Json>> items;
  std::vector<std::pair<std::string, std::shared_ptr<Json>>> members;
  const Json* find(const char* key) const {
    for (auto& m : members)
      if (m.first == key) return m.second.get();
    return nullptr;
  }
  const Json& at(const char* key) const {
    const Json* j = find(key);
    if (!j) throw std::runtime_error(std::string("json: missing field '") + key + "'");
    return *j;
  }
};

class JsonReader {
 public:
  explicit JsonReader(const std::string& s) : s_(s) {}
  std::shared_ptr<Json> read() {
    auto v = value();
    skip();
    if (pos_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t pos_ = 0;
  [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("json: ") + what + " at offset " + std::to_string(pos_)); }
  void skip() { while (pos_ < s_.size() && std::strchr(" \t\r\n", s_[pos_])) ++pos_; }
  char peek() { skip(); if (pos_ >= s_.size()) fail("unexpected end"); return s_[pos_]; }
  void expect(char c) { if (peek() != c) fail("unexpected character"); ++pos_; }
  std::string str() {
    expect('"');
    std::string out;
    while (pos_ < s_.size() && s_[pos_] != '"') {
      char c = s_[pos_++];
      if (c == '\\' && pos_ < s_.size()) {
        char e = s_[pos_++];
        switch (e) { case 'n': c = '\n'; break; case 't': c = '\t'; break; case 'r': c = '\r'; break; default: c = e; }
      }
      out.push_back(c);
    }
    if (pos_ >= s_.size()) fail("unterminated string");
    ++pos_;
    return out;
  }
  std::shared_ptr<Json> value() {
    auto j = std::make_shared<Json>();
    char c = peek();
    if (c == '{') {
      j->kind = Json::Object;
      ++pos_;
      if (peek() == '}') { ++pos_; return j; }
      for (;;) {
        std::string k = str();
        expect(':');
        j->members.emplace_back(k, value());
        char d = peek();
        ++pos_;
        if (d == '}') break;
        if (d != ',') fail("expected ',' or '}'");
      }
    } else if (c == '[') {
      j->kind = Json::Array;
      ++pos_;
      if (peek() == ']') { ++pos_; return j; }
      for (;;) {
        j->items.push_back(value());
        char d = peek();
        ++pos_;
        if (d == ']') break;
        if (d != ',') fail("expected ',' or ']'");
      }
    } else if (c == '"') {
      j->kind = Json::String;
      j->text = str();
    } else if (!s_.compare(pos_, 4, "true")) { j->kind = Json::Bool; j->number = 1; pos_ += 4; }
    else if (!s_.compare(pos_, 5, "false")) { j->kind = Json::Bool; pos_ += 5; }
    else if (!s_.compare(pos_, 4, "null")) { pos_ += 4; }
    else {
      const char* b = s_.c_str() + pos_;
      char* e = nullptr;
      j->kind = Json::Number;
      j->number = std::strtod(b, &e);
      if (e == b) fail("bad token");
      pos_ += (size_t)(e - b);
    }
    return j;
  }
};

}  // namespace jsonmini
}  // namespace s360
