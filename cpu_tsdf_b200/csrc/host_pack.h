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
#include <pthread.h>
#include <sched.h>
#include <string>
#include <cstdlib>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace b2host
{

// fork/join over job indices: fn (j) is called for every j in [0, n), in ascending order of j, on the pool's threads and on the
// calling thread
class PackPool
{
public:
  // cpus: optional Linux cpulist ("0-31,64-95") the pool's own threads are confined to (the caller's thread is left alone)
  explicit PackPool (int threads, const std::string& cpus = std::string ())
  {
    cpu_set_t set; CPU_ZERO (&set);
    bool confine = false;
    for (size_t at = 0; at < cpus.size ();)
    {
      char* e = nullptr;
      const long lo = std::strtol (cpus.c_str () + at, &e, 10);
      if (e == cpus.c_str () + at) break;
      long hi = lo;
      if (*e == '-') { const char* b = e + 1; hi = std::strtol (b, &e, 10); if (e == b) break; }
      for (long c = lo; c <= hi && c < CPU_SETSIZE; ++c) if (c >= 0) { CPU_SET ((int) c, &set); confine = true; }
      at = (size_t) (e - cpus.c_str ());
      if (at < cpus.size () && cpus[at] == ',') ++at; else break;
    }
    for (int i = 0; i < threads - 1; ++i)
    {
      try { th_.emplace_back ([this] { worker (); }); }                                 // the caller is the last worker
      catch (...) { break; }                                                            // (thread limit reached: a smaller pool)
      if (confine) pthread_setaffinity_np (th_.back ().native_handle (), sizeof (set), &set);
    }
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

  // begin () hands the jobs to the pool and returns; help () runs one pending job on the calling thread (false: none left to
  // take); end () drains on the calling thread and returns once every job has been run and no worker still holds `fn`.
  void begin (int njobs, const std::function<void (int)>& fn)
  {
    if (njobs <= 0) return;
    {
      std::lock_guard<std::mutex> g (mu_);
      fn_ = &fn; njobs_ = njobs; next_.store (0, std::memory_order_relaxed); left_.store (njobs, std::memory_order_relaxed);
      ++gen_;
    }
    cv_.notify_all ();
  }
  bool help ()
  {
    if (!fn_) return false;
    const int j = next_.fetch_add (1, std::memory_order_relaxed);
    if (j >= njobs_) return false;
    (*fn_) (j);
    left_.fetch_sub (1, std::memory_order_release);
    return true;
  }
  void end ()
  {
    if (!fn_) return;
    drain ();
    std::unique_lock<std::mutex> g (mu_);
    done_.wait (g, [this] { return left_.load (std::memory_order_acquire) == 0 && active_ == 0; });
    fn_ = nullptr; njobs_ = 0;
  }
  void run (int njobs, const std::function<void (int)>& fn)
  {
    if (njobs == 1) { fn (0); return; }                    // nothing to share: no wake-up
    begin (njobs, fn); end ();
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

// Pipeline stages of a batch of n frames: stages of `chunk` frames while more than `chunk` remain, then the remainder halved down to
// 2 frames, so that little is left to upload and fuse once the last frame has been packed / uploaded.  cbeg[0 .. return] are the
// stage boundaries (cbeg[0] = 0, cbeg[return] = n); at most n stages.
inline int stage_schedule (int n, int chunk, int* cbeg)
{
  int k = 0, at = 0;
  if (chunk < 1) chunk = 1;
  while (n - at > chunk) { cbeg[k++] = at; at += chunk; }
  while (at < n) { const int rem = n - at, t = rem > 2 ? (rem + 1) / 2 : rem; cbeg[k++] = at; at += t; }
  cbeg[k] = n;
  return k;
}

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
