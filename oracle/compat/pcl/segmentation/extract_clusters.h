// oracle/compat/pcl/segmentation/extract_clusters.h — pcl::EuclideanClusterExtraction restated [recalled from PCL's
// extract_clusters.hpp]: points are seeds in index order; a cluster grows breadth first through radiusSearch (tolerance) over
// unprocessed points; it is kept when min <= size <= max (defaults 1 / INT_MAX), its indices sorted and made unique; the
// clusters are finally ordered by decreasing size.  TEST INFRASTRUCTURE for the verbatim build of cleanupMesh.
#pragma once
#include <algorithm>
#include <climits>
#include <vector>
#include "../search/kdtree.h"
namespace pcl
{
struct PointIndices { std::vector<int> indices; };
template <typename PointT> class EuclideanClusterExtraction
{
public:
  EuclideanClusterExtraction () : tol_ (0), min_ (1), max_ (INT_MAX) {}
  void setInputCloud (const typename PointCloud<PointT>::ConstPtr& c) { cloud_ = c; }
  void setSearchMethod (const typename search::KdTree<PointT>::Ptr& t) { tree_ = t; }
  void setClusterTolerance (double t) { tol_ = t; }
  void setMinClusterSize (int n) { min_ = n; }
  void setMaxClusterSize (int n) { max_ = n; }
  void extract (std::vector<PointIndices>& clusters)
  {
    clusters.clear ();
    const size_t n = cloud_->points.size ();
    std::vector<bool> processed (n, false);
    std::vector<int> nn; std::vector<float> nd;
    for (size_t i = 0; i < n; ++i)
    {
      if (processed[i]) continue;
      std::vector<int> q; size_t sq = 0;
      q.push_back (static_cast<int> (i)); processed[i] = true;
      while (sq < q.size ())
      {
        if (!tree_->radiusSearch (q[sq], tol_, nn, nd)) { sq++; continue; }
        for (size_t j = 0; j < nn.size (); ++j)
        {
          if (nn[j] == -1 || processed[nn[j]]) continue;
          q.push_back (nn[j]); processed[nn[j]] = true;
        }
        sq++;
      }
      if (static_cast<int> (q.size ()) >= min_ && static_cast<int> (q.size ()) <= max_)
      {
        PointIndices r; r.indices = q;
        std::sort (r.indices.begin (), r.indices.end ());
        r.indices.erase (std::unique (r.indices.begin (), r.indices.end ()), r.indices.end ());
        clusters.push_back (r);
      }
    }
    std::stable_sort (clusters.begin (), clusters.end (), [] (const PointIndices& a, const PointIndices& b) { return a.indices.size () > b.indices.size (); });
  }
private:
  typename PointCloud<PointT>::ConstPtr cloud_;
  typename search::KdTree<PointT>::Ptr tree_;
  double tol_; int min_, max_;
};
}
