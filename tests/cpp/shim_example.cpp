// tests/cpp/shim_example.cpp — the reference's README usage (README.md:26-50) written against the C++ shim.
// Built (and, on a GPU box, run) by tests/test_cpp_shim.py.
#include "../../include/cpu_tsdf_b200/tsdf_volume_octree.h"
#include <cmath>
#include <cstdio>
using namespace cpu_tsdf_b200;

int main ()
{
  TSDFVolumeOctree::Ptr tsdf (new TSDFVolumeOctree (0, 14));
  if (!tsdf->ok ()) { std::printf ("no device: %d (%s)\n", tsdf->status (), tsdf->lastError ()); return tsdf->status () == B200TSDF_ENODEVICE ? 3 : 1; }
  tsdf->setGridSize (3.f, 3.f, 3.f);
  tsdf->setResolution (256, 256, 256);
  tsdf->setIntegrateColor (false);
  tsdf->setCameraIntrinsics (525., 525., 319.5, 239.5);
  tsdf->reset ();
  if (tsdf->status () != 0) { std::printf ("reset: %s\n", tsdf->lastError ()); return 1; }
  // a fronto-parallel wall 1 m in front of the camera
  PointCloud<PointXYZ> cloud; cloud.width = 640; cloud.height = 480; cloud.points.resize (640 * 480);
  for (int v = 0; v < 480; ++v) for (int u = 0; u < 640; ++u)
  {
    PointXYZ& p = cloud.points[v * 640 + u];
    p.z = 1.0f; p.x = (u - 319.5f) / 525.f; p.y = (v - 239.5f) / 525.f;
  }
  Affine3d pose; pose.m[11] = -1.0;               // camera at z = -1 looking down +z: wall at z = 0
  for (int i = 0; i < 3; ++i) if (!tsdf->integrateCloud (cloud, PointCloud<PointXYZ> (), pose)) { std::printf ("integrate: %s\n", tsdf->lastError ()); return 1; }
  float d = 0; PointXYZ q; q.x = 0.01f; q.y = 0.02f; q.z = -0.012f;
  bool in = tsdf->getFxn (q, d);
  std::printf ("getFxn in=%d d=%f (expect ~ +0.4 = 12 mm / 30 mm in front of the wall)\n", in, d);
  if (!in || std::fabs (d - 0.4f) > 0.15f) return 1;
  PointCloud<PointNormal>::Ptr ray = tsdf->renderView (pose, 4);
  const PointNormal& c = ray->points[(ray->height / 2) * ray->width + ray->width / 2];
  std::printf ("renderView centre depth %f normal z %f\n", c.z, c.normal_z);
  if (std::fabs (c.z - 1.0f) > 0.01f) return 1;
  MarchingCubesTSDFOctree mc;
  mc.setInputTSDF (tsdf);
  mc.setMinWeight (2);
  mc.setColorByRGB (false);
  TriangleSoup mesh;
  if (!mc.reconstruct (mesh) || mesh.xyz.empty ()) return 1;
  std::printf ("mesh: %zu triangles\n", mesh.xyz.size () / 9);
  tsdf->save ("/tmp/shim_example.vol");
  // the rest of the surface a reference user compiles against: coloured render, global transform round trip through a .vol,
  // instantiateFromFile, direct TSDF values, the colour-mode / random-split setters
  Affine3d gt; gt.m[3] = 0.25; gt.m[7] = -0.5;
  tsdf->setGlobalTransform (gt);
  tsdf->save ("/tmp/shim_example_gt.vol");
  TSDFInterface::Ptr back = TSDFInterface::instantiateFromFile ("/tmp/shim_example_gt.vol", 0, 14);
  if (!back->ok () || back->getGlobalTransform ().m[3] != 0.25 || back->getGlobalTransform ().m[7] != -0.5) { std::printf ("instantiateFromFile / getGlobalTransform: %s\n", back->lastError ()); return 1; }
  PointCloud<PointXYZRGBNormal>::Ptr col = back->renderColoredView (pose, 4);
  const PointXYZRGBNormal& cc = col->points[(col->height / 2) * col->width + col->width / 2];
  std::printf ("renderColoredView centre depth %f rgb %d %d %d (a colourless volume renders 127)\n", cc.z, cc.r, cc.g, cc.b);
  if (std::fabs (cc.z - 1.0f) > 0.01f || cc.r != 127 || cc.g != 127 || cc.b != 127) return 1;
  bool valid = true;
  float tv = back->getTSDFValue (0.01f, 0.02f, -0.012f, &valid);
  std::printf ("getTSDFValue %f valid %d\n", tv, valid);
  if (!valid || std::fabs (tv - 0.4f) > 0.15f) return 1;
  tsdf->setNumRandomSplts (4);                        // refused loudly, not ignored
  if (tsdf->integrateCloud (cloud, PointCloud<PointXYZ> (), pose) || tsdf->status () != B200TSDF_EINVAL) return 1;
  tsdf->setNumRandomSplts (1);
  tsdf->setColorMode ("LAB"); tsdf->reset ();
  if (tsdf->status () != B200TSDF_EINVAL) { std::printf ("LAB was not refused\n"); return 1; }
  tsdf->setColorMode ("RGBNormalized"); tsdf->setIntegrateColor (true); tsdf->reset ();
  if (tsdf->status () != 0) { std::printf ("RGBNormalized: %s\n", tsdf->lastError ()); return 1; }
  return 0;
}
