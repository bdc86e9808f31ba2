#pragma once
#include <cstdio>
#define PCL_INFO(...)  do { if (std::getenv ("ORACLE_REF_VERBOSE")) std::printf (__VA_ARGS__); } while (0)
#define PCL_WARN(...)  std::fprintf (stderr, __VA_ARGS__)
#define PCL_ERROR(...) std::fprintf (stderr, __VA_ARGS__)
#include <cstdlib>
