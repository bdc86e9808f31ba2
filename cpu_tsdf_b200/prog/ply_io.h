// ply_io.h — PLY writer for triangle meshes, standing in for pcl::io::savePLYFile / savePLYFileBinary on a
// pcl::PolygonMesh (src/prog/integrate.cpp:707-710, src/prog/tsdf2mesh.cpp:68).  Header layout as PCL writes it
// [recalled]: float x y z (+ uchar red green blue), faces as `property list uchar int vertex_indices`.
#pragma once
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

namespace b200prog {

struct Mesh
{
  std::vector<float> xyz;              // 3 per vertex
  std::vector<std::uint8_t> rgb;       // 3 per vertex or empty
  std::vector<std::int32_t> tris;      // 3 per face
  std::size_t nverts () const { return xyz.size () / 3; }
  std::size_t ntris () const { return tris.size () / 3; }
};

inline std::string save_ply (const std::string& path, const Mesh& m, bool binary)
{
  std::ofstream f (path, std::ios::binary);
  if (!f) return "cannot write " + path;
  const bool color = m.rgb.size () == m.xyz.size () && !m.rgb.empty ();
  f << "ply\nformat " << (binary ? "binary_little_endian" : "ascii") << " 1.0\ncomment PCL generated\n";
  f << "element vertex " << m.nverts () << "\nproperty float x\nproperty float y\nproperty float z\n";
  if (color) f << "property uchar red\nproperty uchar green\nproperty uchar blue\n";
  f << "element face " << m.ntris () << "\nproperty list uchar int vertex_indices\nend_header\n";
  if (binary)
  {
    for (std::size_t i = 0; i < m.nverts (); ++i)
    {
      f.write ((const char*) &m.xyz[3 * i], 12);
      if (color) f.write ((const char*) &m.rgb[3 * i], 3);
    }
    for (std::size_t i = 0; i < m.ntris (); ++i)
    {
      const unsigned char three = 3;
      f.write ((const char*) &three, 1);
      f.write ((const char*) &m.tris[3 * i], 12);
    }
  }
  else
  {
    char buf[160];
    for (std::size_t i = 0; i < m.nverts (); ++i)
    {
      std::snprintf (buf, sizeof buf, "%.9g %.9g %.9g", m.xyz[3 * i], m.xyz[3 * i + 1], m.xyz[3 * i + 2]);
      f << buf;
      if (color) f << ' ' << (int) m.rgb[3 * i] << ' ' << (int) m.rgb[3 * i + 1] << ' ' << (int) m.rgb[3 * i + 2];
      f << '\n';
    }
    for (std::size_t i = 0; i < m.ntris (); ++i)
      f << "3 " << m.tris[3 * i] << ' ' << m.tris[3 * i + 1] << ' ' << m.tris[3 * i + 2] << '\n';
  }
  return f ? "" : "write failed: " + path;
}

} // namespace b200prog
