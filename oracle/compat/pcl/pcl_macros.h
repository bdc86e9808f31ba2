// oracle/compat — TEST INFRASTRUCTURE (see oracle/compat/Eigen/Eigen)
#pragma once
#include <cmath>
#include <cstdint>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
