// organize.cuh — z-buffer re-organisation of unorganised clouds (the front end of the
// reference's `integrate` program, src/prog/integrate.cpp:548-635).
//
// The reference walks the points in order and keeps, per pixel, the point with the smallest z
// (`isnan(old.z) || old.z > pt.z`, :603-606): the winner is the minimum of (z, input index) in
// lexicographic order.  z is positive for every accepted point (reprojectPoint requires z > 0,
// :216-222), so its IEEE bit pattern orders like the value and one 64-bit atomicMin per point
// on the key (z bits << 32 | index) reproduces the sequential result exactly and in any order.
//
//   k_org_clear   zkey[pixel] = ~0
//   k_org_zmin    one thread per point: units, zero->NaN, world->camera, project, atomicMin
//   k_org_gather  one thread per pixel: recompute the winner's point and write it (or the
//                 default point with z = NaN)
#pragma once
#include "tsdf_core.cuh"
#include <cstring>

namespace b2 {

struct OrgParams
{
  float fx, fy, cx, cy;          // float, as the program's globals (integrate.cpp:63-68)
  int width, height;
  float cloud_units;
  int zero_nans, has_tf;
  double tf[12];                 // poses[i].inverse() rows 0..2 (:570-571)
};

// float -> int the way cvttss2si does it: NaN / out of range -> INT_MIN
B2_HD int to_int_x86_f (float v)
{
  if (!(v >= -2147483648.f && v < 2147483648.f)) return INT_MIN;
  return (int) v;
}

// :550-571 — the per-point preprocessing, in the program's order
B2_HD void org_prepare (const OrgParams& o, float& x, float& y, float& z)
{
  if (o.cloud_units != 1.f) { x = fmul (x, o.cloud_units); y = fmul (y, o.cloud_units); z = fmul (z, o.cloud_units); }
  if (o.zero_nans && x == 0.f && y == 0.f && z == 0.f) x = y = z = nanf ("");
  if (o.has_tf)
  {
    // pcl::transformPointCloud<PointT, double>: float (m0*x + m1*y + m2*z + m3), left to right in double
    const double* m = o.tf;
    double p0 = x, p1 = y, p2 = z;
    x = (float) dadd (dadd (dadd (dmul (m[0], p0), dmul (m[1], p1)), dmul (m[2],  p2)), m[3]);
    y = (float) dadd (dadd (dadd (dmul (m[4], p0), dmul (m[5], p1)), dmul (m[6],  p2)), m[7]);
    z = (float) dadd (dadd (dadd (dmul (m[8], p0), dmul (m[9], p1)), dmul (m[10], p2)), m[11]);
  }
}

// reprojectPoint (integrate.cpp:216-222): float arithmetic, truncating conversion; -1 = rejected
B2_HD int org_pixel (const OrgParams& o, float x, float y, float z)
{
  int u = to_int_x86_f (fadd (fdiv (fmul (x, o.fx), z), o.cx));
  int v = to_int_x86_f (fadd (fdiv (fmul (y, o.fy), z), o.cy));
  if (is_nan (z) || !(z > 0.f) || u < 0 || u >= o.width || v < 0 || v >= o.height) return -1;
  return v * o.width + u;
}

B2_HD unsigned long long org_key (float z, unsigned int j)
{
#ifdef __CUDA_ARCH__
  return ((unsigned long long) __float_as_uint (z) << 32) | j;
#else
  unsigned int b; memcpy (&b, &z, 4);
  return ((unsigned long long) b << 32) | j;
#endif
}

B2_HD void org_load (const unsigned char* pts, size_t stride, int xyz_off, size_t j, float& x, float& y, float& z)
{
  const float* q = (const float*) (pts + j * stride + xyz_off);
  x = q[0]; y = q[1]; z = q[2];
}

// one pixel of the organized cloud (integrate.cpp:596-607 result)
B2_HD void org_emit (const OrgParams& o, const unsigned char* pts, size_t stride, int xyz_off, int rgba_off,
                     unsigned long long key, unsigned char* out, size_t out_stride, int out_rgba_off)
{
  float x = 0.f, y = 0.f, z = nanf ("");
  unsigned char c[4] = { 0, 0, 0, 255 };                  // default PointXYZRGBA: b,g,r = 0, a = 255
  if (key != ~0ull)
  {
    size_t j = (size_t) (key & 0xffffffffull);
    org_load (pts, stride, xyz_off, j, x, y, z);
    org_prepare (o, x, y, z);
    if (rgba_off >= 0) { const unsigned char* s = pts + j * stride + rgba_off; c[0] = s[0]; c[1] = s[1]; c[2] = s[2]; c[3] = s[3]; }
  }
  for (size_t b = 0; b < out_stride / 4; ++b) ((unsigned int*) out)[b] = 0u;   // padding bytes are zero
  float* q = (float*) out;
  q[0] = x; q[1] = y; q[2] = z;
  if (out_stride >= 32) q[3] = 1.f;
  if (out_rgba_off >= 0) { unsigned char* d = out + out_rgba_off; d[0] = c[0]; d[1] = c[1]; d[2] = c[2]; d[3] = c[3]; }
}

#ifdef __CUDACC__
__global__ void k_org_clear (unsigned long long* __restrict__ zkey, int npix, unsigned long long* __restrict__ n_filled)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix) zkey[i] = ~0ull;
  if (i == 0) *n_filled = 0ull;
}

__global__ void k_org_zmin (OrgParams o, const unsigned char* __restrict__ pts, size_t n, size_t stride, int xyz_off,
                            unsigned long long* __restrict__ zkey)
{
  for (size_t j = (size_t) blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t) gridDim.x * blockDim.x)
  {
    float x, y, z;
    org_load (pts, stride, xyz_off, j, x, y, z);
    org_prepare (o, x, y, z);
    int pix = org_pixel (o, x, y, z);
    if (pix < 0) continue;
    unsigned long long key = org_key (z, (unsigned int) j);
    if (key < zkey[pix]) atomicMin (&zkey[pix], key);     // the plain read filters most losers
  }
}

__global__ void k_org_gather (OrgParams o, const unsigned char* __restrict__ pts, size_t stride, int xyz_off, int rgba_off,
                              const unsigned long long* __restrict__ zkey, unsigned char* __restrict__ out, size_t out_stride,
                              int out_rgba_off, unsigned long long* __restrict__ n_filled)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int npix = o.width * o.height;
  bool filled = false;
  if (i < npix)
  {
    unsigned long long key = zkey[i];
    filled = key != ~0ull;
    org_emit (o, pts, stride, xyz_off, rgba_off, key, out + (size_t) i * out_stride, out_stride, out_rgba_off);
  }
  unsigned m = __ballot_sync (0xffffffffu, filled);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd (n_filled, (unsigned long long) __popc (m));
}
#endif

} // namespace b2
