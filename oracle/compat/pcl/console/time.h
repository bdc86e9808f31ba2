#pragma once
#include <chrono>
namespace pcl { namespace console {
class TicToc
{
public:
  void tic () { t0_ = std::chrono::steady_clock::now (); }
  double toc () const { return std::chrono::duration<double, std::milli> (std::chrono::steady_clock::now () - t0_).count (); }
private:
  std::chrono::steady_clock::time_point t0_;
};
} }
