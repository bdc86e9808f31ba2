// cpu_tsdf_b200/csrc/host_pack.h — host side of the batched upload: a pcl::PointXYZRGBA is 32 bytes of which the fusion reads 16
// (x, y, z and the packed colour; integrate hpp:64-80 reads nothing else), and the PCIe link is what bounds the end-to-end rate
// of one GPU.  The caller's rows are therefore packed to 16-byte pixels {x, y, z, bgra} by a small pool of host
// threads straight into pinned staging, and only the packed pixels cross the link.  The copy is bit-preserving (no arithmetic).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace b2host
{

// fork/join over job indices: run (n, fn) calls fn (j) for every j in [0, n) on the pool's threads and on the calling thread
class PackPool
{
public:
  explicit PackPool (int threads)
  {
    for (int i = 0; i < threads - 1; ++i) th_.emplace_back ([this] { worker (); });     // the caller is the last worker
  }
  ~PackPool ()
  {
    { std::lock_guard<std::mutex> g (mu_); stop_ = true; }
    cv_.notify_all ();
    for (auto& t : th_) t.join ();
  }
  PackPool (const PackPool&) = delete;
  PackPool& operator= (const PackPool&) = delete;
  int threads () const { return (int) th_.size () + 1; }

  void run (int njobs, const std::function<void (int)>& fn)
  {
    if (njobs <= 0) return;
    {
      std::lock_guard<std::mutex> g (mu_);
      fn_ = &fn; njobs_ = njobs; next_.store (0, std::memory_order_relaxed); left_.store (njobs, std::memory_order_relaxed);
      ++gen_;
    }
    cv_.notify_all ();
    drain ();
    // every job has been run and no worker is still inside this generation's loop (fn must outlive them)
    std::unique_lock<std::mutex> g (mu_);
    done_.wait (g, [this] { return left_.load (std::memory_order_acquire) == 0 && active_ == 0; });
    fn_ = nullptr; njobs_ = 0;
  }

private:
  void drain ()
  {
    for (;;)
    {
      const int j = next_.fetch_add (1, std::memory_order_relaxed);
      if (j >= njobs_) break;
      (*fn_) (j);
      left_.fetch_sub (1, std::memory_order_release);
    }
  }
  void worker ()
  {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> g (mu_);
    for (;;)
    {
      cv_.wait (g, [&] { return stop_ || gen_ != seen; });
      if (stop_) return;
      seen = gen_;
      if (!fn_) continue;                                   // that generation is already over
      ++active_;
      g.unlock ();
      drain ();
      g.lock ();
      --active_;
      done_.notify_all ();
    }
  }

  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void (int)>* fn_ = nullptr;
  int njobs_ = 0, active_ = 0;
  std::atomic<int> next_ { 0 }, left_ { 0 };
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// n points of `stride` bytes -> n 16-byte pixels {x, y, z, colour word}; rgba_off < 0: colour word 0
inline void pack_points16 (const unsigned char* in, size_t stride, int xyz_off, int rgba_off, size_t n, unsigned char* out)
{
#if defined(__SSE2__)
  if (stride == 32 && xyz_off == 0 && (rgba_off == 16 || rgba_off < 0) && (reinterpret_cast<uintptr_t> (in) & 15) == 0 && (reinterpret_cast<uintptr_t> (out) & 15) == 0)
  {
    // the pcl::PointXYZRGBA layout: one aligned 16-byte load, the colour word spliced into the fourth lane, a streaming store
    const __m128i keep = _mm_set_epi32 (0, -1, -1, -1);
    for (size_t i = 0; i < n; ++i)
    {
      const __m128i a = _mm_load_si128 (reinterpret_cast<const __m128i*> (in + 32 * i));
      __m128i v = _mm_and_si128 (a, keep);
      if (rgba_off >= 0)
      {
        int c; std::memcpy (&c, in + 32 * i + 16, 4);
        v = _mm_or_si128 (v, _mm_slli_si128 (_mm_cvtsi32_si128 (c), 12));
      }
      _mm_stream_si128 (reinterpret_cast<__m128i*> (out + 16 * i), v);
    }
    _mm_sfence ();
    return;
  }
#endif
  for (size_t i = 0; i < n; ++i)
  {
    const unsigned char* p = in + i * stride;
    std::memcpy (out + 16 * i, p + xyz_off, 12);
    uint32_t c = 0;
    if (rgba_off >= 0) std::memcpy (&c, p + rgba_off, 4);
    std::memcpy (out + 16 * i + 12, &c, 4);
  }
}

} // namespace b2host
