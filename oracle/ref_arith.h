// oracle/ref_arith.h — TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Arithmetic conventions of the un-vendored third-party code the reference calls
// (Eigen 3.3/3.4 fixed-size expressions, PCL >= 1.10 common/transforms, PCL filters/
// frustum_culling, PCL surface/marching_cubes).  None of those sources exist in this
// container, so every function here is a RESTATEMENT FROM THE
// PUBLISHED SOURCES AS RECALLED; each one names the upstream construct it follows.
// The product library (cpu_tsdf_b200/csrc) implements the same conventions
// independently; parity tests compare the two.
//
// All helpers are written so that no FMA contraction can occur (the oracle is built
// with -ffp-contract=off; the reference's default x86-64 build has no FMA either).
#pragma once
#include <cmath>
#include <cstdint>

namespace ref_arith {

// Eigen redux_novec_unroller<Func,Derived,0,3>: sum of three terms is a0 + (a1 + a2)
// (Eigen/src/Core/Redux.h: HalfLength = Length/2 = 1).
template <typename T> inline T sum3 (T a0, T a1, T a2) { return a0 + (a1 + a2); }

// Eigen dot / squaredNorm of fixed 3-vectors (Dot.h -> cwiseProduct().sum()).
template <typename T> inline T dot3 (const T* a, const T* b)
{ return sum3<T> (a[0] * b[0], a[1] * b[1], a[2] * b[2]); }

// Eigen cross product (Geometry/OrthoMethods.h).
inline void cross3 (const float* a, const float* b, float* o)
{
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// 4x4 row-major affine transform.  Eigen::Transform<S,3,Affine> * Vector3:
// transform_right_product_impl<...,2,1>::run = translation + (linear * v) with the
// lazy coefficient product of a fixed 3x3 (Transform.h).
template <typename T> inline void affine_mul (const T* m, const T* v, T* o)
{
  T r0 = m[3]  + sum3<T> (m[0] * v[0], m[1] * v[1], m[2]  * v[2]);
  T r1 = m[7]  + sum3<T> (m[4] * v[0], m[5] * v[1], m[6]  * v[2]);
  T r2 = m[11] + sum3<T> (m[8] * v[0], m[9] * v[1], m[10] * v[2]);
  o[0] = r0; o[1] = r1; o[2] = r2;
}

// Linear part only (Matrix3 * Vector3, lazy coefficient product).
template <typename T> inline void linear_mul (const T* m, const T* v, T* o)
{
  T r0 = sum3<T> (m[0] * v[0], m[1] * v[1], m[2]  * v[2]);
  T r1 = sum3<T> (m[4] * v[0], m[5] * v[1], m[6]  * v[2]);
  T r2 = sum3<T> (m[8] * v[0], m[9] * v[1], m[10] * v[2]);
  o[0] = r0; o[1] = r1; o[2] = r2;
}

// Eigen::Transform<double,3,Affine>::inverse() (Transform.h, Affine mode): general 3x3
// inverse by cofactors (Inverse.h compute_inverse_size3_helper), translation =
// -(inv * t).  In/out are 4x4 row-major.
inline void affine_inverse (const double* m, double* o)
{
  auto M = [&] (int r, int c) { return m[r * 4 + c]; };
  auto cof = [&] (int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M (i1, j1) * M (i2, j2) - M (i1, j2) * M (i2, j1);
  };
  double c00 = cof (0, 0), c10 = cof (1, 0), c20 = cof (2, 0);
  double det = sum3<double> (c00 * M (0, 0), c10 * M (1, 0), c20 * M (2, 0));
  double invdet = 1.0 / det;
  double inv[9];
  inv[0] = c00 * invdet; inv[1] = c10 * invdet; inv[2] = c20 * invdet;
  inv[3] = cof (0, 1) * invdet; inv[4] = cof (1, 1) * invdet; inv[5] = cof (2, 1) * invdet;
  inv[6] = cof (0, 2) * invdet; inv[7] = cof (1, 2) * invdet; inv[8] = cof (2, 2) * invdet;
  double t[3] = { m[3], m[7], m[11] };
  for (int r = 0; r < 3; ++r)
  {
    o[r * 4 + 0] = inv[r * 3 + 0]; o[r * 4 + 1] = inv[r * 3 + 1]; o[r * 4 + 2] = inv[r * 3 + 2];
    o[r * 4 + 3] = -sum3<double> (inv[r * 3 + 0] * t[0], inv[r * 3 + 1] * t[1], inv[r * 3 + 2] * t[2]);
  }
  o[12] = 0; o[13] = 0; o[14] = 0; o[15] = 1;
}

// pcl::transformPoint (point, Affine3f) — pcl/common/impl/transforms.hpp (>= 1.10),
// detail::Transformer<float>::se3 on SSE2: p0 + (p1 + (p2 + c3)) with pk = ck * src[k].
inline void pcl_transform_point_f (const float* m, const float* v, float* o)
{
  float r0 = m[0] * v[0] + (m[1] * v[1] + (m[2]  * v[2] + m[3]));
  float r1 = m[4] * v[0] + (m[5] * v[1] + (m[6]  * v[2] + m[7]));
  float r2 = m[8] * v[0] + (m[9] * v[1] + (m[10] * v[2] + m[11]));
  o[0] = r0; o[1] = r1; o[2] = r2;
}

// pcl::transformPointCloudWithNormals (cloud, cloud, Affine3d) — generic
// detail::Transformer<double>: float(tf(0,0)*p0 + tf(0,1)*p1 + tf(0,2)*p2 + tf(0,3)),
// left to right in double.
inline void pcl_transform_se3_d (const double* m, const float* v, float* o)
{
  double p0 = v[0], p1 = v[1], p2 = v[2];
  float r0 = static_cast<float> (m[0] * p0 + m[1] * p1 + m[2]  * p2 + m[3]);
  float r1 = static_cast<float> (m[4] * p0 + m[5] * p1 + m[6]  * p2 + m[7]);
  float r2 = static_cast<float> (m[8] * p0 + m[9] * p1 + m[10] * p2 + m[11]);
  o[0] = r0; o[1] = r1; o[2] = r2;
}
inline void pcl_transform_so3_d (const double* m, const float* v, float* o)
{
  double p0 = v[0], p1 = v[1], p2 = v[2];
  float r0 = static_cast<float> (m[0] * p0 + m[1] * p1 + m[2]  * p2);
  float r1 = static_cast<float> (m[4] * p0 + m[5] * p1 + m[6]  * p2);
  float r2 = static_cast<float> (m[8] * p0 + m[9] * p1 + m[10] * p2);
  o[0] = r0; o[1] = r1; o[2] = r2;
}

// Eigen Vector3f::normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z)  (Dot.h).
inline void normalize3 (float* v)
{
  float z = sum3<float> (v[0] * v[0], v[1] * v[1], v[2] * v[2]);
  if (z > 0.f)
  {
    float n = std::sqrt (z);
    v[0] /= n; v[1] /= n; v[2] /= n;
  }
}

// pcl::FrustumCulling<PointT>::applyFilter plane set-up (pcl/filters/impl/
// frustum_culling.hpp).  cam is the 4x4 row-major float camera pose with columns
// (view, up, right, T).  planes[6][4] = l, r, t, b, f, n in evaluation order.
struct Frustum { float pl[6][4]; };

inline Frustum pcl_frustum_planes (const float* cam, float hfov_deg, float vfov_deg, float np_dist, float fp_dist)
{
  float view[3] = { cam[0], cam[4], cam[8] }, up[3] = { cam[1], cam[5], cam[9] };
  float right[3] = { cam[2], cam[6], cam[10] }, T[3] = { cam[3], cam[7], cam[11] };
  float vfov_rad = float (vfov_deg * M_PI / 180);
  float hfov_rad = float (hfov_deg * M_PI / 180);
  float np_h = float (2 * std::tan (vfov_rad / 2) * np_dist);
  float np_w = float (2 * std::tan (hfov_rad / 2) * np_dist);
  float fp_h = float (2 * std::tan (vfov_rad / 2) * fp_dist);
  float fp_w = float (2 * std::tan (hfov_rad / 2) * fp_dist);
  float fp_c[3], fp_tl[3], fp_tr[3], fp_bl[3], fp_br[3], np_c[3], np_tr[3], np_bl[3], np_br[3];
  for (int k = 0; k < 3; ++k)
  {
    fp_c[k] = T[k] + view[k] * fp_dist;
    fp_tl[k] = (fp_c[k] + (up[k] * fp_h / 2)) - (right[k] * fp_w / 2);
    fp_tr[k] = (fp_c[k] + (up[k] * fp_h / 2)) + (right[k] * fp_w / 2);
    fp_bl[k] = (fp_c[k] - (up[k] * fp_h / 2)) - (right[k] * fp_w / 2);
    fp_br[k] = (fp_c[k] - (up[k] * fp_h / 2)) + (right[k] * fp_w / 2);
    np_c[k] = T[k] + view[k] * np_dist;
    np_tr[k] = (np_c[k] + (up[k] * np_h / 2)) + (right[k] * np_w / 2);
    np_bl[k] = (np_c[k] - (up[k] * np_h / 2)) - (right[k] * np_w / 2);
    np_br[k] = (np_c[k] - (up[k] * np_h / 2)) + (right[k] * np_w / 2);
  }
  auto sub = [] (const float* a, const float* b, float* o) { for (int k = 0; k < 3; ++k) o[k] = a[k] - b[k]; };
  Frustum F;
  float e0[3], e1[3], a[3], b[3], c[3], d[3];
  float *pl_l = F.pl[0], *pl_r = F.pl[1], *pl_t = F.pl[2], *pl_b = F.pl[3], *pl_f = F.pl[4], *pl_n = F.pl[5];
  sub (fp_bl, fp_br, e0); sub (fp_tr, fp_br, e1); cross3 (e0, e1, pl_f);
  pl_f[3] = -dot3<float> (fp_c, pl_f);
  sub (np_tr, np_br, e0); sub (np_bl, np_br, e1); cross3 (e0, e1, pl_n);
  pl_n[3] = -dot3<float> (np_c, pl_n);
  sub (fp_bl, T, a); sub (fp_br, T, b); sub (fp_tr, T, c); sub (fp_tl, T, d);
  cross3 (b, c, pl_r); cross3 (d, a, pl_l); cross3 (c, d, pl_t); cross3 (a, b, pl_b);
  pl_r[3] = -dot3<float> (T, pl_r);
  pl_l[3] = -dot3<float> (T, pl_l);
  pl_t[3] = -dot3<float> (T, pl_t);
  pl_b[3] = -dot3<float> (T, pl_b);
  return F;
}

// Vector4f(pt,1).dot(plane) with Eigen's SSE2 predux<Packet4f>: (a0+a2)+(a1+a3).
inline float plane_dot (const float* pl, float x, float y, float z)
{
  return (x * pl[0] + z * pl[2]) + (y * pl[1] + 1.0f * pl[3]);
}
inline bool pcl_frustum_contains (const Frustum& F, float x, float y, float z)
{
  for (int i = 0; i < 6; ++i)
    if (!(plane_dot (F.pl[i], x, y, z) <= 0)) return false;
  return true;
}

} // namespace ref_arith
