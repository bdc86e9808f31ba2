// oracle/compat/pcl/search/kdtree.h — stand-in for pcl::search::KdTree (FLANN kd-tree, absent here) for the verbatim build of
// the reference's program-side functions (oracle/Makefile refprog).  TEST INFRASTRUCTURE.
//
// radiusSearch follows the published behaviour of pcl::KdTreeFLANN::radiusSearch [recalled, see oracle/ref_arith.h]:
// FLANN's L2_Simple<float> accumulates (dx*dx + dy*dy) + dz*dz in float over the first three fields (DefaultPointRepresentation
// caps the dimension at 3, so PointNormal searches on xyz); a point is returned when dist < float (radius * radius) computed in
// double; results are sorted by (distance, index); k_sqr_distances holds the squared distances.  The search itself is a
// uniform bucket grid — only a search structure, membership and order are decided by the rules above.
#pragma once
#include <boost/shared_ptr.hpp>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <utility>
#include <vector>
#include "../point_cloud.h"
namespace pcl { namespace search {
template <typename PointT> class KdTree
{
public:
  typedef boost::shared_ptr<KdTree<PointT>> Ptr;
  explicit KdTree (bool sorted = true) : sorted_ (sorted), cell_ (0.f) {}
  void setInputCloud (const typename pcl::PointCloud<PointT>::ConstPtr& c) { cloud_ = c; cell_ = 0.f; buckets_.clear (); }
  int radiusSearch (int index, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances, unsigned int = 0) const
  { return radiusSearch (cloud_->points[index], radius, k_indices, k_sqr_distances); }
  int radiusSearch (const PointT& q, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances, unsigned int = 0) const
  {
    k_indices.clear (); k_sqr_distances.clear ();
    if (!std::isfinite (q.x) || !std::isfinite (q.y) || !std::isfinite (q.z)) return 0;
    const float r2 = static_cast<float> (radius * radius);
    build (static_cast<float> (radius));
    std::vector<std::pair<float, int>> found;
    const int64_t cx = c1 (q.x), cy = c1 (q.y), cz = c1 (q.z);
    for (int64_t x = cx - 1; x <= cx + 1; ++x) for (int64_t y = cy - 1; y <= cy + 1; ++y) for (int64_t z = cz - 1; z <= cz + 1; ++z)
    {
      auto it = buckets_.find (key (x, y, z));
      if (it == buckets_.end ()) continue;
      for (int j : it->second)
      {
        const PointT& b = cloud_->points[j];
        const float dx = q.x - b.x, dy = q.y - b.y, dz = q.z - b.z;
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < r2) found.push_back (std::make_pair (d, j));
      }
    }
    std::sort (found.begin (), found.end ());
    for (auto& f : found) { k_indices.push_back (f.second); k_sqr_distances.push_back (f.first); }
    return static_cast<int> (found.size ());
  }
private:
  int64_t c1 (float x) const { return static_cast<int64_t> (std::floor (static_cast<double> (x) / cell_)); }
  static uint64_t key (int64_t x, int64_t y, int64_t z)
  { return (static_cast<uint64_t> (x & 0x1fffff) << 42) | (static_cast<uint64_t> (y & 0x1fffff) << 21) | static_cast<uint64_t> (z & 0x1fffff); }
  void build (float radius) const
  {
    const float want = std::max (radius * 1.01f, 1e-6f);
    if (cell_ == want) return;
    cell_ = want; buckets_.clear ();
    for (size_t i = 0; i < cloud_->points.size (); ++i)
    {
      const PointT& p = cloud_->points[i];
      if (!std::isfinite (p.x) || !std::isfinite (p.y) || !std::isfinite (p.z)) continue;
      buckets_[key (c1 (p.x), c1 (p.y), c1 (p.z))].push_back (static_cast<int> (i));
    }
  }
  bool sorted_;
  typename pcl::PointCloud<PointT>::ConstPtr cloud_;
  mutable float cell_;
  mutable std::unordered_map<uint64_t, std::vector<int>> buckets_;
};
} }
