// tools/microbench/pack_numa.cu — is the host packing (host_pack.h) bound by the memory bandwidth of ONE socket?  Inputs and staging are
// placed on chosen NUMA nodes (mmap + mbind + cudaHostRegister) and the packing rate is measured, alone and with a concurrent H2D
// stream of the staging (what b200tsdf_integrate_batch_rows does).
// Build: nvcc -O3 -std=c++17 -Xcompiler -pthread tools/microbench/pack_numa.cu -o tools/microbench/pack_numa
// Run:   pack_numa <threads> <in_node> <out_node> [h2d=0|1]
#include "../../cpu_tsdf_b200/csrc/host_pack.h"
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

static void* alloc_on_node (size_t bytes, int node)
{
  void* p = mmap (nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) { perror ("mmap"); exit (1); }
  if (node >= 0)
  {
    unsigned long mask = 1ul << node;
    if (syscall (SYS_mbind, p, bytes, 2 /* MPOL_BIND */, &mask, sizeof (mask) * 8, 0) != 0) perror ("mbind");
  }
  else if (node == -2)
  {
    unsigned long mask = 3ul;
    if (syscall (SYS_mbind, p, bytes, 3 /* MPOL_INTERLEAVE */, &mask, sizeof (mask) * 8, 0) != 0) perror ("mbind");
  }
  memset (p, 1, bytes);
  if (cudaHostRegister (p, bytes, cudaHostRegisterDefault) != cudaSuccess) { fprintf (stderr, "cudaHostRegister failed\n"); exit (1); }
  return p;
}

int main (int argc, char** argv)
{
  const int T = argc > 1 ? atoi (argv[1]) : 16, in_node = argc > 2 ? atoi (argv[2]) : -1, out_node = argc > 3 ? atoi (argv[3]) : -1, h2d = argc > 4 ? atoi (argv[4]) : 0;
  const size_t npts = 640 * 480, nf = 32;
  unsigned char* in = (unsigned char*) alloc_on_node (2 * nf * npts * 32, in_node);
  unsigned char* out = (unsigned char*) alloc_on_node (2 * nf * npts * 16, out_node);
  unsigned char* dev = nullptr; cudaMalloc (&dev, nf * npts * 16);
  cudaStream_t cs; cudaStreamCreateWithFlags (&cs, cudaStreamNonBlocking);
  b2host::PackPool pool (T);
  double best = 1e9, best_total = 1e9;
  for (int rep = 0; rep < 14; ++rep)
  {
    const unsigned char* src = in + (size_t) (rep & 1) * nf * npts * 32; unsigned char* dst = out + (size_t) (rep & 1) * nf * npts * 16;
    const int nb = 32;
    std::vector<std::atomic<int>> left (nf);
    for (auto& l : left) l.store (nb);
    std::function<void (int)> job = [&] (int j)
    {
      const int i = j / nb, b = j % nb; const size_t p0 = npts * b / nb, p1 = npts * (b + 1) / nb;
      b2host::pack_points16 (src + (i * npts + p0) * 32, 32, 0, 16, p1 - p0, dst + (i * npts + p0) * 16);
      left[i].fetch_sub (1, std::memory_order_release);
    };
    auto t0 = std::chrono::steady_clock::now ();
    pool.begin ((int) nf * nb, job);
    for (size_t i = 0; i < nf; ++i)
    {
      while (left[i].load (std::memory_order_acquire) > 0) if (!pool.help ()) std::this_thread::yield ();
      if (h2d) cudaMemcpyAsync (dev + i * npts * 16, dst + i * npts * 16, npts * 16, cudaMemcpyHostToDevice, cs);
    }
    pool.end ();
    const double ms = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0).count ();
    cudaStreamSynchronize (cs);
    const double ms_total = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0).count ();
    if (rep >= 2) { best = std::min (best, ms); best_total = std::min (best_total, ms_total); }
  }
  printf ("T=%d in_node=%d out_node=%d h2d=%d: pack 32 frames best %.2f ms (%.1f GB/s read), incl. H2D drain %.2f ms -> %.0f frames/s\n", T, in_node, out_node, h2d, best,
          nf * npts * 32 / best / 1e6, best_total, 32e3 / best_total);
  return 0;
}
