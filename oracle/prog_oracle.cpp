// oracle/prog_oracle.cpp — CPU restatement of the data-parallel pieces of the reference's `integrate`
// program (src/prog/integrate.cpp) that sit either side of the volumetric path (SURVEY.md §8(f) rows
// 2-3).  TEST INFRASTRUCTURE ONLY, like the rest of oracle/.
//
// PARITY UNPINNED for this file: integrate.cpp is a program (main() + static helpers) that needs
// boost::program_options, pcl::io, pcl::search::KdTree (FLANN) and pcl::EuclideanClusterExtraction,
// none of which exist here, so it cannot be compiled into oracle/_ref, and the reference has no
// tests for it.  The loops below follow the cited lines statement by statement.
#include "ref_arith.h"

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct PointRGBA { float x, y, z, pad; uint8_t b, g, r, a; uint8_t pad2[12]; };   // pcl::PointXYZRGBA (32 bytes)

// C++ float -> int conversion on x86-64 (cvttss2si): NaN / out of range gives INT_MIN
inline int to_int (float v)
{
  if (!(v >= -2147483648.f && v < 2147483648.f)) return INT_MIN;
  return static_cast<int> (v);
}

} // namespace

extern "C" {

// integrate.cpp:548-607 for one cloud.  intr = {fx, fy, cx, cy} as the program's float globals (:63-68).
// out: width*height pcl::PointXYZRGBA (32 bytes each).  Returns the number of pixels that received a point.
int64_t orc_organize (const void* points, size_t n, size_t stride, int xyz_off, int rgba_off,
                      const float* intr, int width, int height, float cloud_units, int zero_nans,
                      const double* world_to_camera /* 4x4 row-major or NULL */, void* out)
{
  const float fx = intr[0], fy = intr[1], cx = intr[2], cy = intr[3];
  // loadPCDFile into PointXYZRGBA: fields the file lacks keep the point's defaults (rgba = 0,0,0,255)
  std::vector<PointRGBA> cloud (n);
  const unsigned char* base = static_cast<const unsigned char*> (points);
  for (size_t j = 0; j < n; ++j)
  {
    PointRGBA& p = cloud[j];
    std::memset (&p, 0, sizeof (p));
    const float* q = reinterpret_cast<const float*> (base + j * stride + xyz_off);
    p.x = q[0]; p.y = q[1]; p.z = q[2]; p.pad = 1.f; p.a = 255;
    if (rgba_off >= 0) { const unsigned char* c = base + j * stride + rgba_off; p.b = c[0]; p.g = c[1]; p.r = c[2]; p.a = c[3]; }
  }
  // :550-559
  if (cloud_units != 1)
    for (size_t j = 0; j < n; ++j) { cloud[j].x *= cloud_units; cloud[j].y *= cloud_units; cloud[j].z *= cloud_units; }
  // :561-568
  if (zero_nans)
    for (size_t j = 0; j < n; ++j)
      if (cloud[j].x == 0 && cloud[j].y == 0 && cloud[j].z == 0)
        cloud[j].x = cloud[j].y = cloud[j].z = std::numeric_limits<float>::quiet_NaN ();
  // :570-571  pcl::transformPointCloud (*cloud, *cloud, poses[i].inverse ()) with an Affine3d
  if (world_to_camera)
    for (size_t j = 0; j < n; ++j)
    {
      float v[3] = { cloud[j].x, cloud[j].y, cloud[j].z }, o[3];
      ref_arith::pcl_transform_se3_d (world_to_camera, v, o);
      cloud[j].x = o[0]; cloud[j].y = o[1]; cloud[j].z = o[2];
    }
  // :573, :596-598: default-constructed points, z = NaN
  PointRGBA* org = static_cast<PointRGBA*> (out);
  const size_t npix = static_cast<size_t> (width) * height;
  for (size_t i = 0; i < npix; ++i)
  {
    std::memset (&org[i], 0, sizeof (PointRGBA));
    org[i].pad = 1.f; org[i].a = 255;
    org[i].z = std::numeric_limits<float>::quiet_NaN ();
  }
  int64_t filled = 0;
  // :599-627
  for (size_t j = 0; j < n; ++j)
  {
    const PointRGBA& pt = cloud[j];
    // reprojectPoint, :216-222 (float arithmetic: the operands are all float)
    int u = to_int ((pt.x * fx / pt.z) + cx);
    int v = to_int ((pt.y * fy / pt.z) + cy);
    if (!(!std::isnan (pt.z) && pt.z > 0 && u >= 0 && u < width && v >= 0 && v < height)) continue;
    PointRGBA& pt_old = org[static_cast<size_t> (v) * width + u];              // (*cloud_organized) (u, v)
    if (std::isnan (pt_old.z) || (pt_old.z > pt.z))
    {
      if (std::isnan (pt_old.z)) ++filled;
      pt_old = pt;
    }
  }
  return filled;
}

} // extern "C"
