// host_math.h — host-side pose/frustum arithmetic of the engine (double/float, libm).
//
// These are the per-call scalar computations the reference does once per frame on the host
// before its parallel loops (Eigen::Affine3d::inverse, hpp:54; the pcl::FrustumCulling plane
// set-up, tsdf_volume_octree.cpp:632-646).  They stay on the host because they use libm
// (atan/tan), whose results the device's math library does not reproduce bit for bit.
#pragma once
#include <cmath>

namespace b2host {

// a0 + (a1 + a2): Eigen's unrolled 3-term reduction (Redux.h)
template <typename T> inline T sum3 (T a0, T a1, T a2) { return a0 + (a1 + a2); }

// Eigen::Transform<double,3,Affine>::inverse(): cofactor inverse of the linear part,
// translation = -(inv * t).  4x4 row-major in, rows 0..2 (12 doubles) out.
inline void affine_inverse (const double* m, double* o12)
{
  auto M = [&] (int r, int c) { return m[r * 4 + c]; };
  auto cof = [&] (int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return M (i1, j1) * M (i2, j2) - M (i1, j2) * M (i2, j1);
  };
  double c00 = cof (0, 0), c10 = cof (1, 0), c20 = cof (2, 0);
  double det = sum3<double> (c00 * M (0, 0), c10 * M (1, 0), c20 * M (2, 0));
  double invdet = 1.0 / det;
  double inv[9] = { c00 * invdet, c10 * invdet, c20 * invdet,
                    cof (0, 1) * invdet, cof (1, 1) * invdet, cof (2, 1) * invdet,
                    cof (0, 2) * invdet, cof (1, 2) * invdet, cof (2, 2) * invdet };
  double t[3] = { m[3], m[7], m[11] };
  for (int r = 0; r < 3; ++r)
  {
    o12[r * 4 + 0] = inv[r * 3 + 0]; o12[r * 4 + 1] = inv[r * 3 + 1]; o12[r * 4 + 2] = inv[r * 3 + 2];
    o12[r * 4 + 3] = -sum3<double> (inv[r * 3 + 0] * t[0], inv[r * 3 + 1] * t[1], inv[r * 3 + 2] * t[2]);
  }
}

inline void cross3 (const float* a, const float* b, float* o)
{
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
inline float dot3 (const float* a, const float* b) { return sum3<float> (a[0] * b[0], a[1] * b[1], a[2] * b[2]); }

// Six frustum planes (l, r, t, b, f, n) of pcl::FrustumCulling::applyFilter for the camera of
// getFrustumCulledVoxels (tsdf_volume_octree.cpp:632-646): pose * cam2robot has columns
// (z_cam, -y_cam, x_cam, t); FOV scaled by 1.1; near/far = sensor distance bounds.
inline void frustum_planes (const double* pose, int image_width, int image_height, double fx, double fy,
                            float np_dist, float fp_dist, float (*pl)[4])
{
  float view[3], up[3], right[3], T[3];
  for (int r = 0; r < 3; ++r)
  {
    view[r] = static_cast<float> (pose[r * 4 + 2]);
    up[r] = -static_cast<float> (pose[r * 4 + 1]);
    right[r] = static_cast<float> (pose[r * 4 + 0]);
    T[r] = static_cast<float> (pose[r * 4 + 3]);
  }
  float hfov = static_cast<float> (1.1 * 2 * std::fabs (std::atan (0.5 * image_width / fx) * 180 / M_PI));
  float vfov = static_cast<float> (1.1 * 2 * std::fabs (std::atan (0.5 * image_height / fy) * 180 / M_PI));
  float vfov_rad = float (vfov * M_PI / 180);
  float hfov_rad = float (hfov * M_PI / 180);
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
  float e0[3], e1[3], a[3], b[3], c[3], d[3];
  float *pl_l = pl[0], *pl_r = pl[1], *pl_t = pl[2], *pl_b = pl[3], *pl_f = pl[4], *pl_n = pl[5];
  sub (fp_bl, fp_br, e0); sub (fp_tr, fp_br, e1); cross3 (e0, e1, pl_f);
  pl_f[3] = -dot3 (fp_c, pl_f);
  sub (np_tr, np_br, e0); sub (np_bl, np_br, e1); cross3 (e0, e1, pl_n);
  pl_n[3] = -dot3 (np_c, pl_n);
  sub (fp_bl, T, a); sub (fp_br, T, b); sub (fp_tr, T, c); sub (fp_tl, T, d);
  cross3 (b, c, pl_r); cross3 (d, a, pl_l); cross3 (c, d, pl_t); cross3 (a, b, pl_b);
  pl_r[3] = -dot3 (T, pl_r);
  pl_l[3] = -dot3 (T, pl_l);
  pl_t[3] = -dot3 (T, pl_t);
  pl_b[3] = -dot3 (T, pl_b);
}

} // namespace b2host
