// oracle/compat/pcl/filters/frustum_culling.h — pcl::FrustumCulling as restated in oracle/ref_arith.h.
#pragma once
#include "../point_cloud.h"
#include <vector>
namespace pcl
{
template <typename PointT> class FrustumCulling
{
public:
  explicit FrustumCulling (bool = false) : hfov_ (60.f), vfov_ (60.f), np_ (0.1f), fp_ (5.f) {}
  void setCameraPose (const Eigen::Matrix4f& m) { pose_ = m; }
  void setHorizontalFOV (float h) { hfov_ = h; }
  void setVerticalFOV (float v) { vfov_ = v; }
  void setNearPlaneDistance (float d) { np_ = d; }
  void setFarPlaneDistance (float d) { fp_ = d; }
  void setInputCloud (const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
  void filter (std::vector<int>& indices)
  {
    indices.clear ();
    ref_arith::Frustum F = ref_arith::pcl_frustum_planes (pose_.a, hfov_, vfov_, np_, fp_);
    for (size_t i = 0; i < input_->points.size (); ++i)
    {
      const PointT& p = input_->points[i];
      if (ref_arith::pcl_frustum_contains (F, p.x, p.y, p.z)) indices.push_back (static_cast<int> (i));
    }
  }
private:
  Eigen::Matrix4f pose_;
  float hfov_, vfov_, np_, fp_;
  typename PointCloud<PointT>::ConstPtr input_;
};
}
