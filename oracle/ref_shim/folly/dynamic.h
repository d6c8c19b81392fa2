// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// Minimal stand-in for folly::dynamic / folly::parseJson / folly::readFile (absent from this image): the JSON
// value type the reference's Camera (render/Camera.h:20-24, Camera.cpp:44-83, 243-262, 291-310) reads rigs with.
#pragma once
#include <cstdint>
#include <initializer_list>
#include <map>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace folly {

class dynamic {
 public:
  enum Type { NULLT, BOOL, INT64, DOUBLE, STRING, ARRAY, OBJECT };
  struct ObjectMaker;

  dynamic() : type_(NULLT) {}
  dynamic(std::nullptr_t) : type_(NULLT) {}
  dynamic(bool b) : type_(BOOL), i_(b) {}
  dynamic(int v) : type_(INT64), i_(v) {}
  dynamic(long v) : type_(INT64), i_(v) {}
  dynamic(long long v) : type_(INT64), i_(v) {}
  dynamic(unsigned v) : type_(INT64), i_(v) {}
  dynamic(unsigned long v) : type_(INT64), i_((int64_t)v) {}
  dynamic(float v) : type_(DOUBLE), d_(v) {}
  dynamic(double v) : type_(DOUBLE), d_(v) {}
  dynamic(const char* s) : type_(STRING), s_(s) {}
  dynamic(const std::string& s) : type_(STRING), s_(s) {}
  template <class It>
  dynamic(It first, It last) : type_(ARRAY) { for (; first != last; ++first) a_.emplace_back(*first); }
  dynamic(ObjectMaker&& m);

  template <class... Args>
  static dynamic array(Args&&... args) {
    dynamic d;
    d.type_ = ARRAY;
    (void)std::initializer_list<int>{(d.a_.emplace_back(std::forward<Args>(args)), 0)...};
    return d;
  }
  struct ObjectMaker {
    std::map<std::string, std::shared_ptr<dynamic>> o;
    ObjectMaker() {}
    ObjectMaker(const char* k, dynamic v) { o[k] = std::make_shared<dynamic>(std::move(v)); }
    ObjectMaker&& operator()(const char* k, dynamic v) && { o[k] = std::make_shared<dynamic>(std::move(v)); return std::move(*this); }
    ObjectMaker& operator()(const char* k, dynamic v) & { o[k] = std::make_shared<dynamic>(std::move(v)); return *this; }
  };
  // `dynamic::object("k", v)("k2", v2)` and `dynamic::object` (as a callable, Camera.cpp:292)
  struct ObjectFn {
    ObjectMaker operator()() const { return ObjectMaker(); }
    ObjectMaker operator()(const char* k, dynamic v) const { return ObjectMaker(k, std::move(v)); }
  };
  static constexpr ObjectFn object{};

  Type type() const { return type_; }
  bool isObject() const { return type_ == OBJECT; }
  bool isArray() const { return type_ == ARRAY; }
  bool isString() const { return type_ == STRING; }
  bool isNumber() const { return type_ == INT64 || type_ == DOUBLE; }

  const std::string& getString() const { need(STRING); return s_; }
  std::string asString() const;
  double asDouble() const {
    if (type_ == DOUBLE) return d_;
    if (type_ == INT64 || type_ == BOOL) return (double)i_;
    if (type_ == STRING) return std::stod(s_);
    throw std::runtime_error("folly::dynamic: not a number");
  }
  int64_t asInt() const {
    if (type_ == INT64 || type_ == BOOL) return i_;
    if (type_ == DOUBLE) return (int64_t)d_;
    if (type_ == STRING) return std::stoll(s_);
    throw std::runtime_error("folly::dynamic: not a number");
  }
  bool asBool() const { return asInt() != 0; }

  std::size_t size() const {
    if (type_ == ARRAY) return a_.size();
    if (type_ == OBJECT) return o_.size();
    if (type_ == STRING) return s_.size();
    throw std::runtime_error("folly::dynamic: size() of a scalar");
  }
  std::size_t count(const std::string& k) const { need(OBJECT); return o_.count(k); }
  const dynamic& operator[](const std::string& k) const {
    need(OBJECT);
    auto it = o_.find(k);
    if (it == o_.end()) throw std::out_of_range("folly::dynamic: missing key " + k);
    return *it->second;
  }
  const dynamic& operator[](const char* k) const { return (*this)[std::string(k)]; }
  dynamic& operator[](const std::string& k) {
    need(OBJECT);
    auto& p = o_[k];
    if (!p) p = std::make_shared<dynamic>();
    return *p;
  }
  dynamic& operator[](const char* k) { return (*this)[std::string(k)]; }
  const dynamic& operator[](std::size_t i) const { need(ARRAY); return a_.at(i); }
  dynamic& operator[](std::size_t i) { need(ARRAY); return a_.at(i); }
  const dynamic& operator[](int i) const { return (*this)[(std::size_t)i]; }
  dynamic& operator[](int i) { return (*this)[(std::size_t)i]; }
  void push_back(dynamic v) { need(ARRAY); a_.push_back(std::move(v)); }

  // iterating an array (Camera.cpp:256)
  std::vector<dynamic>::const_iterator begin() const { need(ARRAY); return a_.begin(); }
  std::vector<dynamic>::const_iterator end() const { need(ARRAY); return a_.end(); }
  const std::map<std::string, std::shared_ptr<dynamic>>& items() const { need(OBJECT); return o_; }

  friend std::ostream& operator<<(std::ostream& os, const dynamic& d);

 private:
  void need(Type t) const { if (type_ != t) throw std::runtime_error("folly::dynamic: unexpected type"); }
  Type type_;
  int64_t i_ = 0;
  double d_ = 0;
  std::string s_;
  std::vector<dynamic> a_;
  std::map<std::string, std::shared_ptr<dynamic>> o_;
  friend dynamic parseJson(const std::string&);
  friend std::string toJsonImpl(const dynamic&, int, bool);
};
inline dynamic::dynamic(ObjectMaker&& m) : type_(OBJECT), o_(std::move(m.o)) {}

dynamic parseJson(const std::string& text);
std::string toJson(const dynamic& d);
std::string toPrettyJson(const dynamic& d);
bool readFile(const char* path, std::string& out);
bool writeFile(const std::string& data, const char* path);

}  // namespace folly
