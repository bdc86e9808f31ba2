#pragma once
#include "../point_cloud.h"
namespace pcl
{
template <typename PointT> inline void copyPointCloud (const PointCloud<PointT>& in, PointCloud<PointT>& out) { out = in; }
}
