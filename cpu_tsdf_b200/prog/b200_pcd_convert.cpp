// b200_pcd_convert — re-encode a PCD file (ascii | binary | binary_compressed) or dump its points.
//   b200_pcd_convert in.pcd out.pcd <ascii|binary|binary_compressed>
//   b200_pcd_convert in.pcd out.bin --dump        16-byte records {float x, y, z; uint8 b, g, r, a}, then prints "W H N color"
// Small utility around prog/pcd_io.h (the reader the `b200_integrate` front end uses).  Host only, no GPU.
#include "pcd_io.h"

#include <cstdio>
#include <string>

int main (int argc, char** argv)
{
  if (argc < 4) { std::fprintf (stderr, "usage: %s in.pcd out.pcd <ascii|binary|binary_compressed>  |  in.pcd out.bin --dump\n", argv[0]); return 1; }
  b200prog::Cloud cloud;
  std::string e = b200prog::load_pcd (argv[1], cloud);
  if (!e.empty ()) { std::fprintf (stderr, "%s\n", e.c_str ()); return 1; }
  std::string mode = argv[3];
  if (mode == "--dump")
  {
    FILE* f = std::fopen (argv[2], "wb");
    if (!f) { std::fprintf (stderr, "cannot write %s\n", argv[2]); return 1; }
    if (cloud.size ()) std::fwrite (cloud.points.data (), 16, cloud.size (), f);
    std::fclose (f);
    std::printf ("%u %u %zu %d\n", cloud.width, cloud.height, cloud.size (), cloud.has_color ? 1 : 0);
    return 0;
  }
  e = b200prog::save_pcd (argv[2], cloud, mode, cloud.has_color);
  if (!e.empty ()) { std::fprintf (stderr, "%s\n", e.c_str ()); return 1; }
  return 0;
}
