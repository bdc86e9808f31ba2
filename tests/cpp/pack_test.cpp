// tests/cpp/pack_test.cpp — host packing (cpu_tsdf_b200/csrc/host_pack.h): the 16-byte pixels are the source's x, y, z and colour
// word bit for bit, for the pcl::PointXYZRGBA layout (SSE path) and for other strides / offsets / alignments (generic path), and
// the fork/join pool runs every job exactly once (run, and begin / help / end with the caller waiting on per-chunk counters the
// way b200tsdf_integrate_batch_rows does).  Prints "OK" on success.
#include "../../cpu_tsdf_b200/csrc/host_pack.h"
#include <cstdio>
#include <cstdlib>
#include <random>

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::fprintf (stderr, "FAIL line %d: %s\n", __LINE__, #c); ++fails; } } while (0)

static void check_pack (size_t stride, int xyz_off, int rgba_off, size_t n, size_t misalign, std::mt19937& rng)
{
  std::vector<unsigned char> in (n * stride + 64), out (n * 16 + 64, 0xEE);
  for (auto& b : in) b = (unsigned char) rng ();           // every bit pattern, NaNs included
  unsigned char* ip = in.data () + misalign; unsigned char* op = out.data () + ((16 - (reinterpret_cast<uintptr_t> (out.data ()) & 15)) & 15);
  if (misalign == 0) ip += (16 - (reinterpret_cast<uintptr_t> (ip) & 15)) & 15;
  b2host::pack_points16 (ip, stride, xyz_off, rgba_off, n, op);
  for (size_t i = 0; i < n; ++i)
  {
    CHECK (std::memcmp (op + 16 * i, ip + i * stride + xyz_off, 12) == 0);
    uint32_t c = 0; if (rgba_off >= 0) std::memcpy (&c, ip + i * stride + rgba_off, 4);
    CHECK (std::memcmp (op + 16 * i + 12, &c, 4) == 0);
  }
  CHECK (op[16 * n] == 0xEE);
}

int main ()
{
  std::mt19937 rng (7);
  for (size_t n : { (size_t) 0, (size_t) 1, (size_t) 7, (size_t) 1000 })
  {
    check_pack (32, 0, 16, n, 0, rng);      // pcl::PointXYZRGBA, aligned: SSE path
    check_pack (32, 0, -1, n, 0, rng);      // colour ignored
    check_pack (32, 0, 16, n, 4, rng);      // unaligned source: generic path
    check_pack (48, 4, 28, n, 0, rng);      // PointXYZRGBNormal-like stride
    check_pack (20, 0, 16, n, 0, rng);
  }
  for (int threads : { 1, 2, 5 })
  {
    b2host::PackPool pool (threads, threads == 5 ? "0-1" : "");
    CHECK (pool.threads () == threads);
    for (int rep = 0; rep < 50; ++rep)
    {
      const int njobs = rep % 7 == 0 ? 1 : 1 + (int) (rng () % 200);
      std::vector<std::atomic<int>> hit (njobs);
      for (auto& h : hit) h.store (0);
      std::function<void (int)> job = [&] (int j) { hit[j].fetch_add (1); };
      pool.run (njobs, job);
      for (auto& h : hit) CHECK (h.load () == 1);
    }
    // begin / help / end with in-order chunk hand-off
    for (int rep = 0; rep < 50; ++rep)
    {
      const int nchunks = 1 + (int) (rng () % 8), per = 1 + (int) (rng () % 16);
      std::vector<std::atomic<int>> left (nchunks), hit (nchunks * per);
      for (auto& l : left) l.store (per);
      for (auto& h : hit) h.store (0);
      std::function<void (int)> job = [&] (int j) { hit[j].fetch_add (1); left[j / per].fetch_sub (1, std::memory_order_release); };
      pool.begin (nchunks * per, job);
      for (int c = 0; c < nchunks; ++c)
      {
        while (left[c].load (std::memory_order_acquire) > 0) if (!pool.help ()) std::this_thread::yield ();
        for (int j = c * per; j < (c + 1) * per; ++j) CHECK (hit[j].load () == 1);     // the chunk is complete when its counter says so
      }
      pool.end ();
      pool.end ();                                                                      // idempotent
      for (auto& h : hit) CHECK (h.load () == 1);
    }
  }
  // stage schedule of b200tsdf_integrate_batch_rows: covers [0, n) with increasing boundaries, stages never longer than the chunk,
  // the last stage at most 2 frames (when n > 2), never more than n stages
  for (int chunk = 1; chunk <= 32; ++chunk)
    for (int n = 1; n <= 32; ++n)
    {
      int cb[33];
      const int k = b2host::stage_schedule (n, chunk, cb);
      CHECK (k >= 1 && k <= n && cb[0] == 0 && cb[k] == n);
      for (int c = 0; c < k; ++c) CHECK (cb[c + 1] > cb[c] && cb[c + 1] - cb[c] <= std::max (chunk, 1));
      if (n > 2) CHECK (cb[k] - cb[k - 1] <= 2);
    }
  { int cb[33]; CHECK (b2host::stage_schedule (32, 8, cb) == 6 && cb[1] == 8 && cb[3] == 24 && cb[4] == 28 && cb[5] == 30); }
  if (fails) return 1;
  std::puts ("OK");
  return 0;
}
