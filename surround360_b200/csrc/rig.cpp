// Rig JSON parsing and the host camera model. See rig.h for the reference citations.
#include "rig.h"
#include "json_mini.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

namespace s360 {

namespace {
using jsonmini::Json;
using jsonmini::JsonReader;
void readVec(const Json& j, double* out, size_t n) {
  if (j.kind != Json::Array || j.items.size() != n) throw std::runtime_error("rig json: bad vector");
  for (size_t i = 0; i < n; ++i) out[i] = j.items[i]->number;
}
Vec3 readVec3(const Json& j) { double v[3]; readVec(j, v, 3); return {v[0], v[1], v[2]}; }
}  // namespace

// ------------------------------------------------------------------------------------
// Camera
// ------------------------------------------------------------------------------------
// Camera::setRotation(forward, up, right): rows right/up/-forward, then the rotation is
// replaced by AngleAxis(rotation).toRotationMatrix() (Camera.cpp:24-28). Eigen's
// matrix -> quaternion -> angle-axis -> matrix round trip is restated here.
void Camera::setRotation(Vec3 fwd, Vec3 upv, Vec3 rightv) {
  if (!(dot(cross(rightv, upv), fwd) < 0)) throw std::runtime_error("rotation must be right-handed");
  double m[3][3] = {{rightv.x, rightv.y, rightv.z}, {upv.x, upv.y, upv.z}, {-fwd.x, -fwd.y, -fwd.z}};
  // CHECK(rotation.isUnitary(0.001)) (Camera.cpp:24-26), with Eigen's definition: every COLUMN has unit squared norm within
  // |n - 1| <= min(n, 1) * tol, and distinct columns have |dot| <= tol
  const double tol = 0.001;
  for (int i = 0; i < 3; ++i) {
    const double n = m[0][i] * m[0][i] + m[1][i] * m[1][i] + m[2][i] * m[2][i];
    if (!(std::fabs(n - 1.0) <= std::min(std::fabs(n), 1.0) * tol)) throw std::runtime_error("camera rotation is not close to unitary");
    for (int j = 0; j < i; ++j) {
      const double d = m[0][i] * m[0][j] + m[1][i] * m[1][j] + m[2][i] * m[2][j];
      if (!(std::fabs(d) <= tol)) throw std::runtime_error("camera rotation is not close to unitary");
    }
  }
  double q[4];  // x y z w
  double tr = m[0][0] + m[1][1] + m[2][2];
  if (tr > 0) {
    double t = std::sqrt(tr + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
  }
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  double angle = 0, ax = 1, ay = 0, az = 0;
  if (n != 0) {
    angle = 2 * std::atan2(n, std::fabs(q[3]));
    if (q[3] < 0) n = -n;
    ax = q[0] / n; ay = q[1] / n; az = q[2] / n;
  }
  const double s = std::sin(angle), c = std::cos(angle);
  const double sx = s * ax, sy = s * ay, sz = s * az, cx = (1 - c) * ax, cy = (1 - c) * ay, cz = (1 - c) * az;
  double t;
  t = cx * ay; rotation[1] = t - sz; rotation[3] = t + sz;
  t = cx * az; rotation[2] = t + sy; rotation[6] = t - sy;
  t = cy * az; rotation[5] = t - sx; rotation[7] = t + sx;
  rotation[0] = cx * ax + c;
  rotation[4] = cy * ay + c;
  rotation[8] = cz * az + c;
}
void Camera::setFov(double radians) {
  if (!(radians <= M_PI / 2 || type == Type::FTHETA)) throw std::runtime_error("fov > pi/2 needs an FTHETA camera");
  const double c = std::cos(radians);
  fovThreshold = c * std::fabs(c);
}
double Camera::getFov() const {
  return fovThreshold < 0 ? std::acos(-std::sqrt(-fovThreshold)) : std::acos(std::sqrt(fovThreshold));
}
void Camera::setDefaultFov() { fovThreshold = type == Type::FTHETA ? -1 : 0; }
DevCamera Camera::toDevice() const {
  DevCamera d{};
  d.pos[0] = position.x; d.pos[1] = position.y; d.pos[2] = position.z;
  for (int i = 0; i < 9; ++i) d.rot[i] = rotation[i];
  for (int i = 0; i < 2; ++i) { d.principal[i] = principal[i]; d.distortion[i] = distortion[i]; d.focal[i] = focal[i]; }
  d.type = (int)type;
  return d;
}
void Camera::pixel(const Vec3& rig, double& px, double& py) const { cameraPixel(toDevice(), rig.x, rig.y, rig.z, px, py); }
Vec3 Camera::rigDirection(double px, double py) const {
  // Camera::rig -> sensorToCamera (Camera.h:143-150, 263-286), undistort by Newton (Camera.h:229-248)
  const double sx = (px - principal[0]) / focal[0], sy = (py - principal[1]) / focal[1];
  const double sq = sx * sx + sy * sy;
  double ux = 0, uy = 0, uz = -1;
  if (sq != 0) {
    const double norm = std::sqrt(sq);
    double r = norm;
    if (distortion[0] != 0 || distortion[1] != 0) {
      auto distort = [&](double v) { return (1 + v * v * (distortion[0] + v * v * distortion[1])) * v; };
      const double smidgen = 1.0 / 1e6;
      for (int step = 0; step < 10; ++step) {
        const double d0 = distort(r);
        if (std::fabs(d0 - norm) < smidgen) break;
        const double d1 = distort(r + smidgen);
        r -= (d0 - norm) / ((d1 - d0) / smidgen);
      }
    }
    const double angle = type == Type::FTHETA ? r : std::atan(r);
    const double k = std::sin(angle) / norm;
    ux = k * sx; uy = k * sy; uz = -std::cos(angle);
  }
  // rotation^T * unit
  return {rotation[0] * ux + rotation[3] * uy + rotation[6] * uz, rotation[1] * ux + rotation[4] * uy + rotation[7] * uz,
          rotation[2] * ux + rotation[5] * uy + rotation[8] * uz};
}

float Camera::approximateUsablePixelsRadius() const {
  const double fov = getFov();
  const double kStep = 2 * M_PI / 10.0;
  double result = std::sqrt(resolution[0] * resolution[0] + resolution[1] * resolution[1]);
  const Vec3 r = right(), u = up(), f = forward();
  for (double a = 0; a < 2 * M_PI; a += kStep) {
    const Vec3 ortho = r * std::cos(a) + u * std::sin(a);
    const Vec3 direction = f * std::cos(fov) + ortho * std::sin(fov);
    double px, py;
    pixel(position + direction, px, py);
    const double dx = px - resolution[0] / 2.0, dy = py - resolution[1] / 2.0;
    result = std::min(result, std::sqrt(dx * dx + dy * dy));
  }
  return (float)result;
}

// ------------------------------------------------------------------------------------
// Rig
// ------------------------------------------------------------------------------------
void Rig::parse(const std::string& jsonText) {
  JsonReader reader(jsonText);
  auto root = reader.read();
  const Json& cams = root->at("cameras");
  if (cams.kind != Json::Array) throw std::runtime_error("rig json: 'cameras' is not an array");
  cameras.clear();
  side.clear();
  for (auto& cj : cams.items) {
    const Json& j = *cj;
    Camera c;
    if (!(j.at("version").number >= 1.0)) throw std::runtime_error("rig json: camera version < 1");
    c.id = j.at("id").text;
    const std::string& ty = j.at("type").text;
    if (ty == "FTHETA") c.type = Camera::Type::FTHETA;
    else if (ty == "RECTILINEAR") c.type = Camera::Type::RECTILINEAR;
    else throw std::runtime_error("rig json: unknown camera type " + ty);
    c.position = readVec3(j.at("origin"));
    c.setRotation(readVec3(j.at("forward")), readVec3(j.at("up")), readVec3(j.at("right")));
    readVec(j.at("resolution"), c.resolution, 2);
    if (j.find("principal")) readVec(j.at("principal"), c.principal, 2);
    else { c.principal[0] = c.resolution[0] / 2; c.principal[1] = c.resolution[1] / 2; }
    if (j.find("distortion")) readVec(j.at("distortion"), c.distortion, 2);
    if (j.find("fov")) c.setFov(j.at("fov").number);
    else c.setDefaultFov();
    readVec(j.at("focal"), c.focal, 2);
    if (j.find("group")) c.group = j.at("group").text;
    cameras.push_back(c);
  }
  for (size_t i = 0; i < cameras.size(); ++i)
    if (cameras[i].group.find("side") != std::string::npos) side.push_back((int)i);
  if (side.empty()) throw std::runtime_error("rig has no side cameras");
}

static double distCamAxisToRigCenter(const Camera& c) {
  // ParametrizedLine(position, direction).distance(origin)
  const Vec3 d = c.rigDirection(c.principal[0], c.principal[1]);
  const Vec3 diff = Vec3{0, 0, 0} - c.position;
  const Vec3 r = diff - d * dot(d, diff);
  return std::sqrt(dot(r, r));
}
int Rig::findCameraByDirection(Vec3 dir, double maxDist) const {
  int best = -1;
  for (size_t i = 0; i < cameras.size(); ++i)
    if (best < 0 || dot(cameras[best].forward(), dir) < dot(cameras[i].forward(), dir))
      if (distCamAxisToRigCenter(cameras[i]) <= maxDist) best = (int)i;
  if (best < 0) throw std::runtime_error("no camera found by direction");
  return best;
}
int Rig::findLargestDistCamAxisToRigCenter() const {
  int best = (int)cameras.size() - 1;
  for (size_t i = 0; i < cameras.size(); ++i)
    if (distCamAxisToRigCenter(cameras[i]) > distCamAxisToRigCenter(cameras[best])) best = (int)i;
  return best;
}
float Rig::getRingRadius() const {
  const Vec3& p = cameras[side[0]].position;
  return (float)std::sqrt(dot(p, p));
}
float Rig::approximateFov(bool vertical) const {
  float result = 0;
  for (int idx : side) {
    const Camera& c = cameras[idx];
    double ax = c.principal[0], ay = c.principal[1], bx = ax, by = ay;
    if (vertical) { ay = 0; by = c.resolution[1]; }
    else { ax = 0; bx = c.resolution[0]; }
    const Vec3 f = c.forward();
    const float v = (float)std::acos(std::max(dot(c.rigDirection(ax, ay), f), dot(c.rigDirection(bx, by), f)));
    result = std::max(result, v);
  }
  return result;
}

// Binary layout: u32 magic, u32 nCameras, then per camera: 24 doubles + u32 type + u32 isSide
// + u16 idLen + id + u16 groupLen + group.
std::vector<uint8_t> Rig::serialize() const {
  std::vector<uint8_t> out;
  auto put = [&](const void* p, size_t n) { out.insert(out.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
  const uint32_t magic = 0x52333630u, n = (uint32_t)cameras.size();
  put(&magic, 4); put(&n, 4);
  for (const Camera& c : cameras) {
    double d[24] = {c.position.x, c.position.y, c.position.z};
    for (int i = 0; i < 9; ++i) d[3 + i] = c.rotation[i];
    d[12] = c.resolution[0]; d[13] = c.resolution[1]; d[14] = c.principal[0]; d[15] = c.principal[1];
    d[16] = c.distortion[0]; d[17] = c.distortion[1]; d[18] = c.focal[0]; d[19] = c.focal[1]; d[20] = c.fovThreshold;
    put(d, sizeof(d));
    const uint32_t ty = (uint32_t)c.type; put(&ty, 4);
    const uint16_t il = (uint16_t)c.id.size(), gl = (uint16_t)c.group.size();
    put(&il, 2); put(c.id.data(), il); put(&gl, 2); put(c.group.data(), gl);
  }
  return out;
}
void Rig::deserialize(const uint8_t* buf, size_t len) {
  size_t pos = 0;
  auto get = [&](void* p, size_t n) { if (pos + n > len) throw std::runtime_error("rig blob truncated"); std::memcpy(p, buf + pos, n); pos += n; };
  uint32_t magic, n;
  get(&magic, 4); get(&n, 4);
  if (magic != 0x52333630u) throw std::runtime_error("rig blob: bad magic");
  cameras.clear(); side.clear();
  for (uint32_t i = 0; i < n; ++i) {
    Camera c; double d[24]; get(d, sizeof(d));
    c.position = {d[0], d[1], d[2]};
    for (int k = 0; k < 9; ++k) c.rotation[k] = d[3 + k];
    c.resolution[0] = d[12]; c.resolution[1] = d[13]; c.principal[0] = d[14]; c.principal[1] = d[15];
    c.distortion[0] = d[16]; c.distortion[1] = d[17]; c.focal[0] = d[18]; c.focal[1] = d[19]; c.fovThreshold = d[20];
    uint32_t ty; get(&ty, 4); c.type = (Camera::Type)ty;
    uint16_t il, gl; get(&il, 2); c.id.resize(il); get(&c.id[0], il); get(&gl, 2); c.group.resize(gl); get(&c.group[0], gl);
    if (c.group.find("side") != std::string::npos) side.push_back((int)i);
    cameras.push_back(c);
  }
  if (side.empty()) throw std::runtime_error("rig has no side cameras");
}

}  // namespace s360
