// oracle/ref_prog_capi.cpp — C API over the reference's OWN program-side functions (src/prog/integrate.cpp): meshToFaceCloud,
// flattenVertices, cleanupMesh, reprojectPoint (lines 63-222) and the per-cloud preparation + z-buffer re-organisation inside
// main() (lines 559-635).  TEST INFRASTRUCTURE.
//
// integrate.cpp is a program: its main() needs boost::program_options, boost::filesystem, pcl::io and a PCD directory, so the
// file cannot be compiled as a whole here.  Instead oracle/Makefile (target `refprog`) cuts the two line ranges out of the
// reference's source where it lies (sed -n, into oracle/_ref/, never into the repo) and this file #includes them: the
// functions are compiled from the reference's text, and the block from main() is compiled inside a function that declares
// the variables it reads under the names main() gives them.  PCL's KdTree / EuclideanClusterExtraction are the compat
// stand-ins of oracle/compat (recalled library behaviour, documented there).
#include <cpu_tsdf/tsdf_volume_octree.h>
#include <pcl/PolygonMesh.h>
#include <pcl/common/io.h>
#include <pcl/common/transforms.h>
#include <pcl/console/print.h>
#include <pcl/search/kdtree.h>
#include <pcl/segmentation/extract_clusters.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <sys/types.h>

#include "_ref/integrate_fns.inc"           // integrate.cpp:63-222, verbatim

namespace {
void to_mesh (const float* verts, size_t nverts, const int32_t* tris, size_t ntris, pcl::PolygonMesh& mesh)
{
  pcl::PointCloud<pcl::PointXYZ> c;
  for (size_t i = 0; i < nverts; ++i) c.push_back (pcl::PointXYZ (verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]));
  pcl::toPCLPointCloud2 (c, mesh.cloud);
  mesh.polygons.resize (ntris);
  for (size_t i = 0; i < ntris; ++i) { mesh.polygons[i].vertices.resize (3); for (int k = 0; k < 3; ++k) mesh.polygons[i].vertices[k] = static_cast<std::uint32_t> (tris[3 * i + k]); }
}
void from_mesh (const pcl::PolygonMesh& mesh, float* out_verts, size_t* out_nverts, int32_t* out_tris, size_t* out_ntris)
{
  pcl::PointCloud<pcl::PointXYZ> c;
  pcl::fromPCLPointCloud2 (mesh.cloud, c);
  *out_nverts = c.size (); *out_ntris = mesh.polygons.size ();
  for (size_t i = 0; i < c.size (); ++i) { out_verts[3 * i] = c[i].x; out_verts[3 * i + 1] = c[i].y; out_verts[3 * i + 2] = c[i].z; }
  for (size_t i = 0; i < mesh.polygons.size (); ++i) for (int k = 0; k < 3; ++k) out_tris[3 * i + k] = static_cast<int32_t> (mesh.polygons[i].vertices[k]);
}

// integrate.cpp:559-635 inside a function with main()'s variable names
// main() maps world-frame clouds with poses[i].inverse () (:570-571).  The C API receives the world -> camera matrix itself (as the
// engine and the restatement do), so `poses[i]` is an object whose inverse () IS that matrix: the block's text stays as written and
// no second matrix inversion (compat-layer arithmetic) slips between the two sides of the comparison
struct GivenInverse { Eigen::Affine3d w2c; Eigen::Affine3d inverse () const { return w2c; } };
int organize_block (pcl::PointCloud<pcl::PointXYZRGBA>::Ptr cloud, float cloud_units, bool zero_nans, bool world_frame,
                    std::vector<GivenInverse>& poses, bool organized, bool verbose, pcl::PointCloud<pcl::PointXYZRGBA>::Ptr& result)
{
  const size_t i = 0;
  {
#include "_ref/integrate_organize.inc"      // integrate.cpp:559-635, verbatim (ends with cloud_organized filled)
    result = cloud_organized;
  }
  return 0;
}
} // namespace

extern "C" {

void orc_flatten_vertices (const float* verts, size_t nverts, const int32_t* tris, size_t ntris, float min_dist,
                           float* out_verts, size_t* out_nverts, int32_t* out_tris, size_t* out_ntris)
{
  pcl::PolygonMesh mesh; to_mesh (verts, nverts, tris, ntris, mesh);
  flattenVertices (mesh, min_dist);
  from_mesh (mesh, out_verts, out_nverts, out_tris, out_ntris);
}

void orc_cleanup_mesh (const float* verts, size_t nverts, const int32_t* tris, size_t ntris, float face_dist, int min_neighbors,
                       float* out_verts, size_t* out_nverts, int32_t* out_tris, size_t* out_ntris)
{
  pcl::PolygonMesh mesh; to_mesh (verts, nverts, tris, ntris, mesh);
  cleanupMesh (mesh, face_dist, min_neighbors);
  from_mesh (mesh, out_verts, out_nverts, out_tris, out_ntris);
}

// same contract as oracle/prog_oracle.cpp::orc_organize: out = height x width pcl::PointXYZRGBA (32 bytes), returns the filled pixels
int64_t orc_organize (const void* points, size_t n, size_t stride, int xyz_off, int rgba_off,
                      const float* intr, int width, int height, float cloud_units, int zero_nans,
                      const double* world_to_camera /* 4x4 row-major or NULL */, void* out)
{
  width_ = width; height_ = height;
  focal_length_x_ = intr[0]; focal_length_y_ = intr[1]; principal_point_x_ = intr[2]; principal_point_y_ = intr[3];
  pcl::PointCloud<pcl::PointXYZRGBA>::Ptr cloud (new pcl::PointCloud<pcl::PointXYZRGBA>);
  const unsigned char* base = static_cast<const unsigned char*> (points);
  for (size_t k = 0; k < n; ++k)
  {
    pcl::PointXYZRGBA pt;
    const float* f = reinterpret_cast<const float*> (base + k * stride + xyz_off);
    pt.x = f[0]; pt.y = f[1]; pt.z = f[2];
    if (rgba_off >= 0) std::memcpy (&pt.rgba, base + k * stride + rgba_off, 4);
    cloud->push_back (pt);
  }
  std::vector<GivenInverse> poses (1);
  bool world_frame = world_to_camera != nullptr;
  if (world_frame) for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) poses[0].w2c.matrix () (r, c) = world_to_camera[4 * r + c];
  pcl::PointCloud<pcl::PointXYZRGBA>::Ptr result;
  if (organize_block (cloud, cloud_units, zero_nans != 0, world_frame, poses, false, false, result)) return -1;
  std::memcpy (out, result->points.data (), result->points.size () * sizeof (pcl::PointXYZRGBA));
  int64_t filled = 0;
  for (size_t k = 0; k < result->points.size (); ++k) filled += !std::isnan (result->points[k].z);
  return filled;
}

} // extern "C"
