// oracle/compat — minimal stand-in so the reference's sources compile verbatim (TEST INFRASTRUCTURE).
#pragma once
#include <memory>
namespace boost { template <typename T> using shared_ptr = std::shared_ptr<T>; }
