// Host side of the camera model: rig JSON (RIG_JSON.md) -> cameras -> the tables the
// kernels consume. Mirrors the reference interface for this path:
//   Camera / Camera::loadRig      surround360_render/source/render/Camera.{h,cpp}
//   RigDescription                surround360_render/source/render/RigDescription.{h,cpp}
//   approximateFov                surround360_render/source/test/TestRenderStereoPanorama.cpp:75-97
// Names follow the reference (forward/up/right, getFov, findCameraByDirection, ...).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "common.cuh"

namespace s360 {

struct Vec3 {
  double x = 0, y = 0, z = 0;
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

struct Camera {
  enum class Type { FTHETA = 0, RECTILINEAR = 1 };
  Type type = Type::FTHETA;
  Vec3 position;
  double rotation[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // rows: right, up, backward
  double resolution[2] = {0, 0}, principal[2] = {0, 0}, distortion[2] = {0, 0}, focal[2] = {0, 0};
  double fovThreshold = -1;
  std::string id, group;

  Vec3 right() const { return {rotation[0], rotation[1], rotation[2]}; }
  Vec3 up() const { return {rotation[3], rotation[4], rotation[5]}; }
  Vec3 backward() const { return {rotation[6], rotation[7], rotation[8]}; }
  Vec3 forward() const { return backward() * -1.0; }
  void setRotation(Vec3 forward, Vec3 up, Vec3 right);  // re-unitarises like Camera.cpp:16-29
  void setFov(double radians);
  double getFov() const;
  void setDefaultFov();
  DevCamera toDevice() const;
  void pixel(const Vec3& rig, double& px, double& py) const;  // Camera.h:133-140
  // Camera::rig(pixel): origin = position, returns the (unnormalised) direction
  Vec3 rigDirection(double px, double py) const;
  // Camera::approximateUsablePixelsRadius (Camera.h:201-212)
  float approximateUsablePixelsRadius() const;
};

struct Rig {
  std::vector<Camera> cameras;   // RigDescription::rig
  std::vector<int> side;         // indices of RigDescription::rigSideOnly
  void parse(const std::string& jsonText);  // throws std::runtime_error
  int findCameraByDirection(Vec3 dir, double distCamAxisToRigCenterMax = 1.0) const;
  int findLargestDistCamAxisToRigCenter() const;
  float getRingRadius() const;
  float approximateFov(bool vertical) const;  // over the side cameras
  int topCamera() const { return findCameraByDirection({0, 0, 1}); }
  int bottomCamera() const { return findCameraByDirection({0, 0, -1}); }
  // binary (de)serialisation for the start-up broadcast
  std::vector<uint8_t> serialize() const;
  void deserialize(const uint8_t* buf, size_t len);
};

}  // namespace s360
