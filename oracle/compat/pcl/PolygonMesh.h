#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "point_cloud.h"
namespace pcl
{
struct PCLPointCloud2 { std::uint32_t width = 0, height = 0, point_step = 0; std::vector<std::uint8_t> data; };
struct Vertices { std::vector<std::uint32_t> vertices; };
struct PolygonMesh { PCLPointCloud2 cloud; std::vector<Vertices> polygons; };
// the leading sizeof (PointT) bytes of every record (the callers convert xyz-first clouds to PointXYZ)
template <typename PointT> inline void fromPCLPointCloud2 (const PCLPointCloud2& m, PointCloud<PointT>& c)
{
  const size_t n = m.point_step ? m.data.size () / m.point_step : 0;
  c.points.assign (n, PointT ());
  for (size_t i = 0; i < n; ++i) std::memcpy (&c.points[i], m.data.data () + i * m.point_step, sizeof (PointT) < m.point_step ? sizeof (PointT) : m.point_step);
  c.width = static_cast<std::uint32_t> (n); c.height = 1;
}
template <typename PointT> inline void toPCLPointCloud2 (const PointCloud<PointT>& c, PCLPointCloud2& m)
{
  m.width = static_cast<std::uint32_t> (c.points.size ()); m.height = 1; m.point_step = sizeof (PointT);
  m.data.resize (c.points.size () * sizeof (PointT));
  if (!c.points.empty ()) std::memcpy (m.data.data (), c.points.data (), m.data.size ());
}
}
