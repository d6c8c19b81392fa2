// ORACLE — TEST INFRASTRUCTURE ONLY (see ocv.h header).
//
// Restates the camera model and rig description of the reference:
//   Camera            surround360_render/source/render/Camera.h:133-140,221-295
//                     surround360_render/source/render/Camera.cpp:16-83,147-170
//   RigDescription    surround360_render/source/render/RigDescription.cpp:18-78
//   approximateFov    surround360_render/source/test/TestRenderStereoPanorama.cpp:75-97
// Eigen (AngleAxis/Quaternion) and folly::json are third-party and absent; their
// published algorithms are restated inline. Pinned by Camera::unitTest's known answers
// (Camera.cpp:291-374) in tests/test_oracle_camera.py.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

// ---------------- minimal JSON (RIG_JSON.md format) ----------------
struct JVal {
  enum T { NUL, NUM, STR, ARR, OBJ, BOOL } t = NUL;
  double num = 0;
  std::string str;
  std::vector<JVal> arr;
  std::map<std::string, JVal> obj;
  bool has(const std::string& k) const { return t == OBJ && obj.count(k); }
  const JVal& operator[](const std::string& k) const {
    auto it = obj.find(k);
    if (it == obj.end()) throw std::runtime_error("json: missing key " + k);
    return it->second;
  }
  const JVal& operator[](size_t i) const { return arr.at(i); }
};
struct JParser {
  const std::string& s;
  size_t i = 0;
  explicit JParser(const std::string& s_) : s(s_) {}
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i; }
  JVal parse() {
    ws();
    if (i >= s.size()) throw std::runtime_error("json: unexpected end");
    JVal v;
    char ch = s[i];
    if (ch == '{') {
      v.t = JVal::OBJ;
      ++i;
      ws();
      if (s[i] == '}') { ++i; return v; }
      for (;;) {
        ws();
        JVal k = parse();
        if (k.t != JVal::STR) throw std::runtime_error("json: key must be string");
        ws();
        if (s[i++] != ':') throw std::runtime_error("json: expected :");
        v.obj[k.str] = parse();
        ws();
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == '}') { ++i; break; }
        throw std::runtime_error("json: expected , or }");
      }
    } else if (ch == '[') {
      v.t = JVal::ARR;
      ++i;
      ws();
      if (s[i] == ']') { ++i; return v; }
      for (;;) {
        v.arr.push_back(parse());
        ws();
        if (s[i] == ',') { ++i; continue; }
        if (s[i] == ']') { ++i; break; }
        throw std::runtime_error("json: expected , or ]");
      }
    } else if (ch == '"') {
      v.t = JVal::STR;
      ++i;
      while (i < s.size() && s[i] != '"') {
        if (s[i] == '\\' && i + 1 < s.size()) ++i;
        v.str.push_back(s[i++]);
      }
      ++i;
    } else if (!s.compare(i, 4, "true")) { v.t = JVal::BOOL; v.num = 1; i += 4; }
    else if (!s.compare(i, 5, "false")) { v.t = JVal::BOOL; i += 5; }
    else if (!s.compare(i, 4, "null")) { i += 4; }
    else {
      char* e = nullptr;
      v.t = JVal::NUM;
      v.num = strtod(s.c_str() + i, &e);
      if (e == s.c_str() + i) throw std::runtime_error("json: bad token");
      i = e - s.c_str();
    }
    return v;
  }
};

// ---------------- tiny fixed-size linear algebra (double) ----------------
struct V2 { double x = 0, y = 0; };
struct V3 {
  double x = 0, y = 0, z = 0;
  V3() {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  V3 operator+(const V3& o) const { return {x + o.x, y + o.y, z + o.z}; }
  V3 operator-(const V3& o) const { return {x - o.x, y - o.y, z - o.z}; }
  V3 operator*(double s) const { return {x * s, y * s, z * s}; }
  V3 operator-() const { return {-x, -y, -z}; }
  double dot(const V3& o) const { return x * o.x + y * o.y + z * o.z; }
  V3 cross(const V3& o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }
  double norm() const { return std::sqrt(dot(*this)); }
};
struct M3 {
  double m[3][3];
  V3 row(int r) const { return {m[r][0], m[r][1], m[r][2]}; }
  V3 mul(const V3& v) const {  // column-accumulated product
    return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
            m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
  }
  V3 tmul(const V3& v) const {
    return {m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z, m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
            m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z};
  }
};
struct Ray {
  V3 origin, dir;
  V3 pointAt(double t) const { return origin + dir * t; }
  double distance(const V3& p) const {  // Eigen::ParametrizedLine::distance
    V3 diff = p - origin;
    V3 r = diff - dir * dir.dot(diff);
    return std::sqrt(r.dot(r));
  }
};

// Eigen::AngleAxis<double>(Matrix3).toRotationMatrix() — re-unitarisation, Camera.cpp:24-28
static inline M3 reunitarize(const M3& a) {
  double q[4];  // x y z w
  double t = a.m[0][0] + a.m[1][1] + a.m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (a.m[2][1] - a.m[1][2]) * t;
    q[1] = (a.m[0][2] - a.m[2][0]) * t;
    q[2] = (a.m[1][0] - a.m[0][1]) * t;
  } else {
    int i = 0;
    if (a.m[1][1] > a.m[0][0]) i = 1;
    if (a.m[2][2] > a.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(a.m[i][i] - a.m[j][j] - a.m[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (a.m[k][j] - a.m[j][k]) * t;
    q[j] = (a.m[j][i] + a.m[i][j]) * t;
    q[k] = (a.m[k][i] + a.m[i][k]) * t;
  }
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  double angle;
  V3 ax;
  if (n != 0) {
    angle = 2 * std::atan2(n, std::fabs(q[3]));
    if (q[3] < 0) n = -n;
    ax = V3(q[0] / n, q[1] / n, q[2] / n);
  } else {
    angle = 0;
    ax = V3(1, 0, 0);
  }
  double s = std::sin(angle), c = std::cos(angle);
  V3 sa = ax * s, c1 = ax * (1 - c);
  M3 r;
  double tmp;
  tmp = c1.x * ax.y; r.m[0][1] = tmp - sa.z; r.m[1][0] = tmp + sa.z;
  tmp = c1.x * ax.z; r.m[0][2] = tmp + sa.y; r.m[2][0] = tmp - sa.y;
  tmp = c1.y * ax.z; r.m[1][2] = tmp - sa.x; r.m[2][1] = tmp + sa.x;
  r.m[0][0] = c1.x * ax.x + c;
  r.m[1][1] = c1.y * ax.y + c;
  r.m[2][2] = c1.z * ax.z + c;
  return r;
}

struct Camera {
  enum Type { FTHETA = 0, RECTILINEAR = 1 } type = FTHETA;
  V3 position;
  M3 rotation;
  V2 resolution, principal, distortion, focal;
  double fovThreshold = -1;
  std::string id, group;
  static constexpr double kNearInfinity = 1e6;

  V3 backward() const { return rotation.row(2); }
  V3 forward() const { return -backward(); }
  V3 up() const { return rotation.row(1); }
  V3 right() const { return rotation.row(0); }

  void setRotation(const V3& fwd, const V3& upv, const V3& rightv) {
    if (!(rightv.cross(upv).dot(fwd) < 0)) throw std::runtime_error("rotation must be right-handed");
    M3 r;
    V3 b = -fwd;
    r.m[2][0] = b.x; r.m[2][1] = b.y; r.m[2][2] = b.z;
    r.m[1][0] = upv.x; r.m[1][1] = upv.y; r.m[1][2] = upv.z;
    r.m[0][0] = rightv.x; r.m[0][1] = rightv.y; r.m[0][2] = rightv.z;
    // CHECK(rotation.isUnitary(0.001)), Camera.cpp:24-26 (Eigen: unit columns, |n - 1| <= min(n, 1) tol; orthogonal columns, |dot| <= tol)
    const double tol = 0.001;
    for (int i = 0; i < 3; ++i) {
      const double n = r.m[0][i] * r.m[0][i] + r.m[1][i] * r.m[1][i] + r.m[2][i] * r.m[2][i];
      if (!(std::fabs(n - 1.0) <= (std::fabs(n) < 1.0 ? std::fabs(n) : 1.0) * tol)) throw std::runtime_error("camera rotation is not close to unitary");
      for (int j = 0; j < i; ++j) {
        const double d = r.m[0][i] * r.m[0][j] + r.m[1][i] * r.m[1][j] + r.m[2][i] * r.m[2][j];
        if (!(std::fabs(d) <= tol)) throw std::runtime_error("camera rotation is not close to unitary");
      }
    }
    rotation = reunitarize(r);
  }
  void setFov(double fov) {
    double c = std::cos(fov);
    fovThreshold = c * std::fabs(c);
  }
  double getFov() const {
    return fovThreshold < 0 ? std::acos(-std::sqrt(-fovThreshold)) : std::acos(std::sqrt(fovThreshold));
  }
  void setDefaultFov() { fovThreshold = type == FTHETA ? -1 : 0; }
  bool isDefaultFov() const { return type == FTHETA ? fovThreshold == -1 : fovThreshold == 0; }

  double distortFactor(double r2) const { return 1 + r2 * (distortion.x + r2 * distortion.y); }
  double distort(double r) const { return distortFactor(r * r) * r; }
  double undistort(double d) const {
    if (distortion.x == 0 && distortion.y == 0) return d;
    double r0 = d;
    const double smidgen = 1.0 / kNearInfinity;
    for (int step = 0; step < 10; ++step) {
      double d0 = distort(r0);
      if (std::fabs(d0 - d) < smidgen) break;
      double r1 = r0 + smidgen;
      double d1 = distort(r1);
      double derivative = (d1 - d0) / smidgen;
      r0 -= (d0 - d) / derivative;
    }
    return r0;
  }
  V2 cameraToSensor(const V3& cam) const {
    if (type == FTHETA) {
      double norm = std::sqrt(cam.x * cam.x + cam.y * cam.y);
      double r = std::atan2(norm, -cam.z);
      double s = distort(r) / norm;
      return {s * cam.x, s * cam.y};
    }
    V2 planar = {cam.x / -cam.z, cam.y / -cam.z};
    double f = distortFactor(planar.x * planar.x + planar.y * planar.y);
    return {f * planar.x, f * planar.y};
  }
  V2 pixel(const V3& rig) const {
    V3 cam = rotation.mul(rig - position);
    V2 s = cameraToSensor(cam);
    return {focal.x * s.x + principal.x, focal.y * s.y + principal.y};
  }
  V3 sensorToCamera(const V2& sensor) const {
    double sq = sensor.x * sensor.x + sensor.y * sensor.y;
    if (sq == 0) return V3(0, 0, -1);
    double norm = std::sqrt(sq);
    double r = undistort(norm);
    double angle = type == FTHETA ? r : std::atan(r);
    double s = std::sin(angle) / norm;
    return V3(s * sensor.x, s * sensor.y, -std::cos(angle));
  }
  Ray rig(const V2& px) const {
    V2 sensor = {(px.x - principal.x) / focal.x, (px.y - principal.y) / focal.y};
    V3 unit = sensorToCamera(sensor);
    return Ray{position, rotation.tmul(unit)};
  }
  V3 rigNearInfinity(const V2& px) const { return rig(px).pointAt(kNearInfinity); }
  bool isBehind(const V3& r) const { return backward().dot(r - position) >= 0; }
  bool isOutsideFov(const V3& r) const {
    if (fovThreshold == -1) return false;
    if (fovThreshold == 0) return isBehind(r);
    V3 v = r - position;
    double dot = -backward().dot(v);
    return dot * std::fabs(dot) <= fovThreshold * v.dot(v);
  }
  bool sees(const V3& r) const {
    if (isOutsideFov(r)) return false;
    V2 p = pixel(r);
    return 0 <= p.x && p.x < resolution.x && 0 <= p.y && p.y < resolution.y;
  }
  // Camera.h:201-212
  static float approximateUsablePixelsRadius(const Camera& c) {
    const double fov = c.getFov();
    const double kStep = 2 * M_PI / 10.0;
    double result = std::sqrt(c.resolution.x * c.resolution.x + c.resolution.y * c.resolution.y);
    for (double a = 0; a < 2 * M_PI; a += kStep) {
      V3 ortho = c.right() * std::cos(a) + c.up() * std::sin(a);
      V3 direction = c.forward() * std::cos(fov) + ortho * std::sin(fov);
      V2 px = c.pixel(c.position + direction);
      double dx = px.x - c.resolution.x / 2.0, dy = px.y - c.resolution.y / 2.0;
      result = std::min(result, std::sqrt(dx * dx + dy * dy));
    }
    return (float)result;
  }

  static V3 vec3(const JVal& j) { return V3(j[0].num, j[1].num, j[2].num); }
  static V2 vec2(const JVal& j) { return V2{j[0].num, j[1].num}; }
  static Camera fromJson(const JVal& j) {  // Camera.cpp:44-83
    Camera c;
    if (!(j["version"].num >= 1.0)) throw std::runtime_error("camera version < 1");
    c.id = j["id"].str;
    const std::string& ty = j["type"].str;
    if (ty == "FTHETA") c.type = FTHETA;
    else if (ty == "RECTILINEAR") c.type = RECTILINEAR;
    else throw std::runtime_error("unknown camera type " + ty);
    c.position = vec3(j["origin"]);
    c.setRotation(vec3(j["forward"]), vec3(j["up"]), vec3(j["right"]));
    c.resolution = vec2(j["resolution"]);
    if (j.has("principal")) c.principal = vec2(j["principal"]);
    else c.principal = {c.resolution.x / 2, c.resolution.y / 2};
    if (j.has("distortion")) c.distortion = vec2(j["distortion"]);
    if (j.has("fov")) c.setFov(j["fov"].num);
    else c.setDefaultFov();
    c.focal = vec2(j["focal"]);
    if (j.has("group")) c.group = j["group"].str;
    return c;
  }
};

// TRSP:75-97
static inline float approximateFov(const Camera& cam, bool vertical) {
  V2 a = cam.principal, b = cam.principal;
  if (vertical) { a.y = 0; b.y = cam.resolution.y; }
  else { a.x = 0; b.x = cam.resolution.x; }
  V3 f = cam.forward();
  return (float)std::acos(std::max(cam.rig(a).dir.dot(f), cam.rig(b).dir.dot(f)));
}
static inline float approximateFov(const std::vector<Camera>& rig, bool vertical) {
  float r = 0;
  for (const auto& c : rig) r = std::max(r, approximateFov(c, vertical));
  return r;
}

struct Rig {  // RigDescription.cpp:18-78
  std::vector<Camera> rig, side;
  static Rig fromString(const std::string& json) {
    JParser p(json);
    JVal root = p.parse();
    Rig r;
    for (const JVal& c : root["cameras"].arr) r.rig.push_back(Camera::fromJson(c));
    for (const Camera& c : r.rig)
      if (c.group.find("side") != std::string::npos) r.side.push_back(c);
    if (r.side.empty()) throw std::runtime_error("rig has no side cameras");
    return r;
  }
  static Rig load(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("could not read JSON file: " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return fromString(ss.str());
  }
  static double axisDist(const Camera& c) { return c.rig(c.principal).distance(V3(0, 0, 0)); }
  const Camera& findByDirection(const V3& d, double maxAxisDist = 1.0) const {
    const Camera* best = nullptr;
    for (const Camera& c : rig)
      if (!best || best->forward().dot(d) < c.forward().dot(d))
        if (axisDist(c) <= maxAxisDist) best = &c;
    if (!best) throw std::runtime_error("no camera found by direction");
    return *best;
  }
  const Camera& findLargestAxisDist() const {
    const Camera* best = &rig.back();
    for (const Camera& c : rig)
      if (axisDist(c) > axisDist(*best)) best = &c;
    return *best;
  }
  float ringRadius() const { return (float)side[0].position.norm(); }
};

}  // namespace orc
