// oracle/compat/pcl/point_types.h — the PCL point structs the reference uses, with PCL's memory layouts
// (pcl/impl/point_types.hpp: 16-byte xyz block, colour bytes b,g,r,a at offset 16).  TEST INFRASTRUCTURE.
#pragma once
#include <Eigen/Eigen>
#include <cstdint>
namespace pcl
{
// a writable view of three consecutive floats (Eigen::Map<Vector3f> in PCL)
struct Vector3fMap
{
  float* p;
  explicit Vector3fMap (float* q) : p (q) {}
  operator Eigen::Vector3f () const { return Eigen::Vector3f (p[0], p[1], p[2]); }
  Vector3fMap& operator= (const Eigen::Vector3f& v) { p[0] = v (0); p[1] = v (1); p[2] = v (2); return *this; }
  Vector3fMap& operator= (const Vector3fMap& o) { p[0] = o.p[0]; p[1] = o.p[1]; p[2] = o.p[2]; return *this; }
  Eigen::Vector3f normalized () const { return Eigen::Vector3f (p[0], p[1], p[2]).normalized (); }
  Eigen::Vector3f operator+ (const Eigen::Vector3f& o) const { return Eigen::Vector3f (p[0], p[1], p[2]) + o; }
};
struct Vector3fMapConst
{
  const float* p;
  explicit Vector3fMapConst (const float* q) : p (q) {}
  operator Eigen::Vector3f () const { return Eigen::Vector3f (p[0], p[1], p[2]); }
  Eigen::Vector3f normalized () const { return Eigen::Vector3f (p[0], p[1], p[2]).normalized (); }
  Eigen::Vector3f operator+ (const Eigen::Vector3f& o) const { return Eigen::Vector3f (p[0], p[1], p[2]) + o; }
};
inline Eigen::Vector3f operator* (const Eigen::Affine3f& t, const Vector3fMap& m) { return t * static_cast<Eigen::Vector3f> (m); }
inline Eigen::Vector3f operator* (const Eigen::Affine3f& t, const Vector3fMapConst& m) { return t * static_cast<Eigen::Vector3f> (m); }

#define ORC_XYZ union { float data[4]; struct { float x, y, z; }; }; \
  Vector3fMap getVector3fMap () { return Vector3fMap (data); } \
  Vector3fMapConst getVector3fMap () const { return Vector3fMapConst (data); }
#define ORC_NORMAL union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; }; \
  Vector3fMap getNormalVector3fMap () { return Vector3fMap (data_n); } \
  Vector3fMapConst getNormalVector3fMap () const { return Vector3fMapConst (data_n); }
#define ORC_RGB union { union { struct { std::uint8_t b, g, r, a; }; float rgb; }; std::uint32_t rgba; };

struct PointXYZ
{
  ORC_XYZ
  PointXYZ () { x = y = z = 0.f; data[3] = 1.f; }
  PointXYZ (float _x, float _y, float _z) { x = _x; y = _y; z = _z; data[3] = 1.f; }
};
struct PointXYZRGBA
{
  ORC_XYZ
  union { ORC_RGB float data_c[4]; };
  PointXYZRGBA () { x = y = z = 0.f; data[3] = 1.f; r = g = b = 0; a = 255; }
};
struct PointXYZRGB
{
  ORC_XYZ
  union { ORC_RGB float data_c[4]; };
  PointXYZRGB () { x = y = z = 0.f; data[3] = 1.f; r = g = b = 0; a = 255; }
};
struct Normal
{
  ORC_NORMAL
  union { struct { float curvature; }; float data_c[4]; };
  Normal () { normal_x = normal_y = normal_z = data_n[3] = 0.f; curvature = 0.f; }
};
struct PointNormal
{
  ORC_XYZ
  ORC_NORMAL
  union { struct { float curvature; }; float data_c[4]; };
  PointNormal () { x = y = z = 0.f; data[3] = 1.f; normal_x = normal_y = normal_z = data_n[3] = 0.f; curvature = 0.f; }
};
struct PointXYZRGBNormal
{
  ORC_XYZ
  ORC_NORMAL
  union { struct { ORC_RGB float curvature; }; float data_c[4]; };
  PointXYZRGBNormal () { x = y = z = 0.f; data[3] = 1.f; normal_x = normal_y = normal_z = data_n[3] = 0.f; r = g = b = 0; a = 255; curvature = 0.f; }
};
struct Intensity { float intensity; };
static_assert (sizeof (PointXYZ) == 16 && sizeof (PointXYZRGBA) == 32 && sizeof (PointNormal) == 48 && sizeof (PointXYZRGBNormal) == 48, "PCL layouts");
}
