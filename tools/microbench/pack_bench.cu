// tools/microbench/pack_bench.cu — host packing throughput (host_pack.h) on the GPU box's cores: threads x affinity x kernel variant.
// Build: nvcc -O3 -std=c++17 -Xcompiler -pthread,-mavx2 tools/microbench/pack_bench.cu -o gpurun_out/pack_bench
#include "../../cpu_tsdf_b200/csrc/host_pack.h"
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <sched.h>
#include <immintrin.h>

static void pack_avx2 (const unsigned char* in, size_t n, unsigned char* out)
{
  const __m256i keep = _mm256_set_epi32 (0, -1, -1, -1, 0, -1, -1, -1);
  size_t i = 0;
  for (; i + 2 <= n; i += 2)
  {
    const unsigned char* p = in + 32 * i;
    _mm_prefetch ((const char*) (p + 1024), _MM_HINT_NTA);
    __m256i a = _mm256_set_m128i (_mm_load_si128 ((const __m128i*) (p + 32)), _mm_load_si128 ((const __m128i*) p));
    __m256i c = _mm256_set_m128i (_mm_slli_si128 (_mm_cvtsi32_si128 (*(const int*) (p + 48)), 12), _mm_slli_si128 (_mm_cvtsi32_si128 (*(const int*) (p + 16)), 12));
    _mm256_stream_si256 ((__m256i*) (out + 16 * i), _mm256_or_si256 (_mm256_and_si256 (a, keep), c));
  }
  if (i < n) b2host::pack_points16 (in + 32 * i, 32, 0, 16, n - i, out + 16 * i);
  _mm_sfence ();
}

int main (int argc, char** argv)
{
  const int T = argc > 1 ? atoi (argv[1]) : 8, variant = argc > 2 ? atoi (argv[2]) : 0, chunk = argc > 3 ? atoi (argv[3]) : 8;
  const char* cpus = argc > 4 ? argv[4] : nullptr;         // "lo-hi" range the whole process is confined to (inherited by the pool)
  if (cpus)
  {
    int lo, hi; sscanf (cpus, "%d-%d", &lo, &hi);
    cpu_set_t set; CPU_ZERO (&set); for (int c = lo; c <= hi; ++c) CPU_SET (c, &set);
    sched_setaffinity (0, sizeof (set), &set);
  }
  const size_t npts = 640 * 480;
  std::vector<unsigned char*> in (64);
  for (auto& p : in) { cudaHostAlloc (&p, npts * 32, cudaHostAllocDefault); memset (p, 1, npts * 32); }
  unsigned char* out; cudaHostAlloc (&out, 32 * npts * 16, cudaHostAllocDefault); memset (out, 0, 32 * npts * 16);
  b2host::PackPool pool (T);
  double best = 1e9;
  for (int rep = 0; rep < 12; ++rep)
  {
    auto t0 = std::chrono::steady_clock::now ();
    for (int c0 = 0; c0 < 32; c0 += chunk)
    {
      const int nb = std::max (1, std::min ((4 * T + chunk - 1) / chunk, (int) npts / 4096));
      std::function<void (int)> job = [&] (int j)
      {
        const int i = c0 + j / nb, b = j % nb; const size_t p0 = (npts * b / nb) & ~(size_t) 1, p1 = b + 1 == nb ? npts : ((npts * (b + 1) / nb) & ~(size_t) 1);
        const unsigned char* src = in[(i + 32 * (rep & 1))] + p0 * 32; unsigned char* dst = out + i * npts * 16 + p0 * 16;
        if (variant == 1) pack_avx2 (src, p1 - p0, dst); else b2host::pack_points16 (src, 32, 0, 16, p1 - p0, dst);
      };
      pool.run (chunk * nb, job);
    }
    const double ms = std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0).count ();
    if (rep >= 2) best = std::min (best, ms);
  }
  printf ("T=%d variant=%d chunk=%d cpus=%s: 32 frames best %.2f ms  read %.1f GB/s\n", T, variant, chunk, cpus ? cpus : "all", best, 32 * npts * 32 / best / 1e6);
  return 0;
}
