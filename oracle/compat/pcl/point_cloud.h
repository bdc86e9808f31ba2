#pragma once
#include <boost/shared_ptr.hpp>
#include <cstdint>
#include <vector>
#include "point_types.h"
namespace pcl
{
template <typename PointT> class PointCloud
{
public:
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PointCloud () : width (0), height (0), is_dense (true) {}
  PointCloud (std::uint32_t w, std::uint32_t h) : points (static_cast<size_t> (w) * h), width (w), height (h), is_dense (true) {}
  const PointT& operator() (size_t column, size_t row) const { return points[row * width + column]; }
  PointT& operator() (size_t column, size_t row) { return points[row * width + column]; }
  const PointT& at (size_t n) const { return points.at (n); }
  PointT& at (size_t n) { return points.at (n); }
  const PointT& operator[] (size_t n) const { return points[n]; }
  PointT& operator[] (size_t n) { return points[n]; }
  size_t size () const { return points.size (); }
  void push_back (const PointT& pt) { points.push_back (pt); width = static_cast<std::uint32_t> (points.size ()); height = 1; }
  void resize (size_t n) { points.resize (n); if (width * height != n) { width = static_cast<std::uint32_t> (n); height = 1; } }
  std::vector<PointT> points;
  std::uint32_t width, height;
  bool is_dense;
};
}
