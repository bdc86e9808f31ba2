// oracle/compat/pcl/common/transforms.h — pcl::transformPoint / transformPointCloud[WithNormals] with the
// conventions documented in oracle/ref_arith.h.  TEST INFRASTRUCTURE.
#pragma once
#include "../point_cloud.h"
#include <cmath>
namespace pcl
{
template <typename PointT> inline PointT transformPoint (const PointT& point, const Eigen::Affine3f& transform)
{
  PointT ret = point;
  ref_arith::pcl_transform_point_f (transform.matrix ().a, point.data, ret.data);
  return ret;
}
template <typename PointT> inline void transformPointCloud (const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Affine3d& t)
{
  if (&in != &out) out = in;
  for (size_t i = 0; i < out.points.size (); ++i)
  {
    if (!in.is_dense && (!std::isfinite (out.points[i].x) || !std::isfinite (out.points[i].y) || !std::isfinite (out.points[i].z))) continue;
    float q[3];
    ref_arith::pcl_transform_se3_d (t.matrix ().a, out.points[i].data, q);
    out.points[i].x = q[0]; out.points[i].y = q[1]; out.points[i].z = q[2];
  }
}
template <typename PointT> inline void transformPointCloudWithNormals (const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Affine3d& t)
{
  if (&in != &out) out = in;
  for (size_t i = 0; i < out.points.size (); ++i)
  {
    if (!in.is_dense && (!std::isfinite (out.points[i].x) || !std::isfinite (out.points[i].y) || !std::isfinite (out.points[i].z))) continue;
    float q[3], n[3];
    ref_arith::pcl_transform_se3_d (t.matrix ().a, out.points[i].data, q);
    ref_arith::pcl_transform_so3_d (t.matrix ().a, out.points[i].data_n, n);
    out.points[i].x = q[0]; out.points[i].y = q[1]; out.points[i].z = q[2];
    out.points[i].normal_x = n[0]; out.points[i].normal_y = n[1]; out.points[i].normal_z = n[2];
  }
}
}
