// oracle/compat/pcl/surface/marching_cubes.h — the part of pcl::MarchingCubes the reference's adaptor
// relies on (members and createSurface/interpolateEdge/getBoundingBox; pcl/surface/impl/marching_cubes.hpp,
// PCL >= 1.9 names), restated.  Case tables: oracle/mc_tables.h.  TEST INFRASTRUCTURE.
#pragma once
#include "../PolygonMesh.h"
#include "../../../mc_tables.h"
#include <vector>
namespace pcl
{
template <typename PointNT> class MarchingCubes
{
public:
  MarchingCubes () : percentage_extend_grid_ (0.f), iso_level_ (0.f), res_x_ (32), res_y_ (32), res_z_ (32) {}
  virtual ~MarchingCubes () {}
  void setIsoLevel (float iso) { iso_level_ = iso; }
  void setGridResolution (int x, int y, int z) { res_x_ = x; res_y_ = y; res_z_ = z; }
  void setPercentageExtendGrid (float p) { percentage_extend_grid_ = p; }
  void setInputCloud (const typename PointCloud<PointNT>::ConstPtr& c) { input_ = c; }
  void reconstruct (PolygonMesh& output) { performReconstruction (output); }
protected:
  std::vector<float> grid_;
  float percentage_extend_grid_, iso_level_;
  int res_x_, res_y_, res_z_;
  Eigen::Array3f upper_boundary_, lower_boundary_, size_voxel_;
  typename PointCloud<PointNT>::ConstPtr input_;

  virtual void voxelizeData () = 0;
  virtual void performReconstruction (PolygonMesh& output) = 0;

  void getBoundingBox ()
  {
    float lo[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, hi[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
    for (const auto& p : input_->points)
      for (int k = 0; k < 3; ++k) { lo[k] = p.data[k] < lo[k] ? p.data[k] : lo[k]; hi[k] = p.data[k] > hi[k] ? p.data[k] : hi[k]; }
    lower_boundary_ = Eigen::Array3f (lo[0], lo[1], lo[2]);
    upper_boundary_ = Eigen::Array3f (hi[0], hi[1], hi[2]);
    Eigen::Array3f d = upper_boundary_ - lower_boundary_;
    float e = 0.5f * percentage_extend_grid_;
    Eigen::Array3f ext (e * d[0], e * d[1], e * d[2]);
    lower_boundary_ = lower_boundary_ - ext;
    upper_boundary_ = upper_boundary_ + ext;
  }
  void interpolateEdge (const float* p1, const float* p2, float val_p1, float val_p2, float* output)
  {
    const float mu = (iso_level_ - val_p1) / (val_p2 - val_p1);
    for (int k = 0; k < 3; ++k) output[k] = p1[k] + mu * (p2[k] - p1[k]);
  }
  void createSurface (const std::vector<float>& leaf_node, const Eigen::Vector3i& index_3d, PointCloud<PointNT>& cloud)
  {
    int cubeindex = 0;
    for (int k = 0; k < 8; ++k) if (leaf_node[k] < iso_level_) cubeindex |= (1 << k);
    if (mc_tables::edge_table[cubeindex] == 0) return;
    float center[3], p[8][3];
    for (int k = 0; k < 3; ++k) center[k] = lower_boundary_[k] + size_voxel_[k] * static_cast<float> (index_3d[k]);
    for (int i = 0; i < 8; ++i)
    {
      p[i][0] = center[0]; p[i][1] = center[1]; p[i][2] = center[2];
      if (i & 0x4) p[i][1] = static_cast<float> (center[1] + size_voxel_[1]);
      if (i & 0x2) p[i][2] = static_cast<float> (center[2] + size_voxel_[2]);
      if ((i & 0x1) ^ ((i >> 1) & 0x1)) p[i][0] = static_cast<float> (center[0] + size_voxel_[0]);
    }
    float vl[12][3];
    for (int e = 0; e < 12; ++e)
      if (mc_tables::edge_table[cubeindex] & (1 << e))
      {
        int a = mc_tables::edge_corners[e][0], b = mc_tables::edge_corners[e][1];
        interpolateEdge (p[a], p[b], leaf_node[a], leaf_node[b], vl[e]);
      }
    for (int i = 0; mc_tables::tri_table[cubeindex][i] != -1; i += 3)
      for (int j = 0; j < 3; ++j)
      {
        PointNT q;
        const float* s = vl[mc_tables::tri_table[cubeindex][i + j]];
        q.x = s[0]; q.y = s[1]; q.z = s[2];
        cloud.push_back (q);
      }
  }
};
}
