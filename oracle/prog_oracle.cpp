// oracle/prog_oracle.cpp — CPU restatement of the data-parallel pieces of the reference's `integrate`
// program (src/prog/integrate.cpp) that sit either side of the volumetric path (SURVEY.md §8(f) rows
// 2-3).  TEST INFRASTRUCTURE ONLY, like the rest of oracle/.
//
// PINNED (round 2): integrate.cpp is a program (main() + static helpers) that cannot be compiled as a whole here, but its
// functions can: oracle/Makefile `refprog` cuts lines 63-222 (meshToFaceCloud, flattenVertices, cleanupMesh, reprojectPoint)
// and 559-635 (cloud preparation + z-buffer organisation inside main()) out of the reference's source where it lies and
// compiles them verbatim (oracle/ref_prog_capi.cpp); tests/test_ref_prog_pin.py requires this restatement to agree with that
// build bit for bit.  What stays recalled is library behaviour only (FLANN radius search order, EuclideanClusterExtraction),
// stated in oracle/compat/pcl/search/kdtree.h and .../segmentation/extract_clusters.h.
#include "ref_arith.h"

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <limits>
#include <unordered_map>
#include <vector>

namespace {

struct PointRGBA { float x, y, z, pad; uint8_t b, g, r, a; uint8_t pad2[12]; };   // pcl::PointXYZRGBA (32 bytes)

// C++ float -> int conversion on x86-64 (cvttss2si): NaN / out of range gives INT_MIN
inline int to_int (float v)
{
  if (!(v >= -2147483648.f && v < 2147483648.f)) return INT_MIN;
  return static_cast<int> (v);
}

// Stand-in for pcl::search::KdTree<PointT>::radiusSearch (FLANN kd-tree, absent here): returns the
// indices j != query with squared distance < radius^2.  Conventions [recalled]: FLANN's L2_Simple<float>
// accumulates (dx*dx + dy*dy) + dz*dz in float; RadiusResultSet keeps dist < radius^2 where
// radius^2 = float (double (radius) * double (radius)) (KdTreeFLANN::radiusSearch).  The query point itself
// (distance 0, returned first by the sorted search and skipped by the callers' `j = 1` / processed flag) is
// left out.  A uniform bucket grid is only the search structure; membership is decided by the test above.
struct RadiusSearch
{
  const float* pts; size_t n; float cell, r2;
  std::unordered_map<uint64_t, std::vector<int> > buckets;
  static int64_t c1 (float x, float cell) { return static_cast<int64_t> (std::floor (static_cast<double> (x) / cell)); }
  static uint64_t key (int64_t x, int64_t y, int64_t z)
  { return (static_cast<uint64_t> (x & 0x1fffff) << 42) | (static_cast<uint64_t> (y & 0x1fffff) << 21) | static_cast<uint64_t> (z & 0x1fffff); }
  RadiusSearch (const float* p, size_t n_, float radius) : pts (p), n (n_)
  {
    cell = std::max (radius * 1.01f, 1e-6f);
    r2 = radius > 0.f ? static_cast<float> (static_cast<double> (radius) * static_cast<double> (radius)) : 0.f;   // no radius, no neighbours
    buckets.reserve (n * 2);
    for (size_t i = 0; i < n; ++i)
      buckets[key (c1 (p[3 * i], cell), c1 (p[3 * i + 1], cell), c1 (p[3 * i + 2], cell))].push_back (static_cast<int> (i));
  }
  void search (int i, std::vector<int>& out) const
  {
    out.clear ();
    const float* q = pts + 3 * static_cast<size_t> (i);
    int64_t cx = c1 (q[0], cell), cy = c1 (q[1], cell), cz = c1 (q[2], cell);
    for (int64_t x = cx - 1; x <= cx + 1; ++x)
      for (int64_t y = cy - 1; y <= cy + 1; ++y)
        for (int64_t z = cz - 1; z <= cz + 1; ++z)
        {
          auto it = buckets.find (key (x, y, z));
          if (it == buckets.end ()) continue;
          for (int j : it->second)
          {
            if (j == i) continue;
            const float* b = pts + 3 * static_cast<size_t> (j);
            float dx = q[0] - b[0], dy = q[1] - b[1], dz = q[2] - b[2];
            float d = dx * dx + dy * dy + dz * dz;
            if (d < r2) out.push_back (j);
          }
        }
  }
};

} // namespace

extern "C" {

// flattenVertices, integrate.cpp:103-150.  out_verts has room for nverts xyz, out_tris for ntris triples.
void orc_flatten_vertices (const float* verts, size_t nverts, const int32_t* tris, size_t ntris, float min_dist,
                           float* out_verts, size_t* out_nverts, int32_t* out_tris, size_t* out_ntris)
{
  RadiusSearch vert_tree (verts, nverts, min_dist);
  // :109-127 Find duplicates
  std::vector<int> vertex_remap (nverts, -1);
  int idx = 0;
  std::vector<int> neighbors;
  size_t nnew = 0;
  for (size_t i = 0; i < nverts; i++)
  {
    if (vertex_remap[i] >= 0) continue;
    vertex_remap[i] = idx;
    vert_tree.search (static_cast<int> (i), neighbors);
    // :121-125: every returned neighbour passes `dists[j] < min_dist` (a squared distance below radius^2 <= 1e-8
    // compared with 1e-4), so all of them are (re)assigned
    for (size_t j = 0; j < neighbors.size (); j++) vertex_remap[neighbors[j]] = idx;
    out_verts[3 * nnew] = verts[3 * i]; out_verts[3 * nnew + 1] = verts[3 * i + 1]; out_verts[3 * nnew + 2] = verts[3 * i + 2];
    ++nnew;
    idx++;
  }
  // :128-147
  size_t face_idx = 0;
  for (size_t i = 0; i < ntris; i++)
  {
    int32_t v[3];
    for (int j = 0; j < 3; j++) v[j] = vertex_remap[tris[3 * i + j]];
    if (v[0] == v[1] || v[1] == v[2] || v[2] == v[0]) continue;              // "Degenerate face"
    out_tris[3 * face_idx] = v[0]; out_tris[3 * face_idx + 1] = v[1]; out_tris[3 * face_idx + 2] = v[2];
    ++face_idx;
  }
  *out_nverts = nnew; *out_ntris = face_idx;
}

// cleanupMesh, integrate.cpp:152-214, with pcl::EuclideanClusterExtraction (pcl/segmentation/impl/
// extract_clusters.hpp, extractEuclideanClusters [recalled]: breadth-first growth over radiusSearch, clusters
// with min_pts (default 1) <= size <= max_pts are returned).
void orc_cleanup_mesh (const float* verts, size_t nverts, const int32_t* tris, size_t ntris, float face_dist, int min_neighbors,
                       float* out_verts, size_t* out_nverts, int32_t* out_tris, size_t* out_ntris)
{
  // meshToFaceCloud, :70-101: p_new.getVector3fMap () = (v0 + v1 + v2) / 3.  (Vector3f arithmetic: float)
  std::vector<float> faces (3 * ntris);
  for (size_t i = 0; i < ntris; ++i)
    for (int k = 0; k < 3; ++k)
      faces[3 * i + k] = ((verts[3 * tris[3 * i] + k] + verts[3 * tris[3 * i + 1] + k]) + verts[3 * tris[3 * i + 2] + k]) / 3.f;
  RadiusSearch face_tree (faces.data (), ntris, face_dist);
  std::vector<std::vector<int> > clusters;
  {
    std::vector<bool> processed (ntris, false);
    std::vector<int> nn;
    for (size_t i = 0; i < ntris; ++i)
    {
      if (processed[i]) continue;
      std::vector<int> seed_queue;
      size_t sq_idx = 0;
      seed_queue.push_back (static_cast<int> (i));
      processed[i] = true;
      while (sq_idx < seed_queue.size ())
      {
        face_tree.search (seed_queue[sq_idx], nn);
        for (int j : nn)
        {
          if (processed[j]) continue;
          seed_queue.push_back (j);
          processed[j] = true;
        }
        sq_idx++;
      }
      if (seed_queue.size () >= 1 && seed_queue.size () <= static_cast<size_t> (min_neighbors)) clusters.push_back (seed_queue);
    }
  }
  // :170-183
  std::vector<size_t> faces_to_remove;
  for (auto& c : clusters) for (int j : c) faces_to_remove.push_back (static_cast<size_t> (j));
  std::sort (faces_to_remove.begin (), faces_to_remove.end ());
  std::vector<bool> erased (ntris, false);
  for (size_t f : faces_to_remove) erased[f] = true;                           // polygons.erase, back to front
  std::vector<int32_t> polys;
  for (size_t i = 0; i < ntris; ++i) if (!erased[i]) { polys.push_back (tris[3 * i]); polys.push_back (tris[3 * i + 1]); polys.push_back (tris[3 * i + 2]); }
  // :184-213 Remove all vertices with no face
  std::vector<bool> has_face (nverts, false);
  for (int32_t v : polys) has_face[v] = true;
  std::vector<size_t> get_new_idx (nverts);
  size_t cur_idx = 0;
  for (size_t i = 0; i < nverts; i++)
    if (has_face[i])
    {
      out_verts[3 * cur_idx] = verts[3 * i]; out_verts[3 * cur_idx + 1] = verts[3 * i + 1]; out_verts[3 * cur_idx + 2] = verts[3 * i + 2];
      get_new_idx[i] = cur_idx++;
    }
  for (size_t i = 0; i < polys.size (); ++i) out_tris[i] = static_cast<int32_t> (get_new_idx[polys[i]]);
  *out_nverts = cur_idx; *out_ntris = polys.size () / 3;
}

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
