// oracle/compat — replaces include/eigen_extensions/eigen_extensions.h for the oracle/_ref build: only the two
// functions the library calls (tsdf_volume_octree.cpp:242, :269), same text format
// (eigen_extensions.h:249-294: "% rows cols" then the matrix at precision 16, columns aligned as Eigen's
// default IOFormat does).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <Eigen/Eigen>
#include <algorithm>
#include <cstdio>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>

namespace eigen_extensions
{
template <class S, int T, int U> void serializeASCII (const Eigen::Mat<S, T, U>& mat, std::ostream& strm)
{
  int old_precision = strm.precision ();
  strm.precision (16);
  strm << "% " << mat.rows () << " " << mat.cols () << std::endl;
  std::string cell[T * U]; size_t width = 0;
  for (int i = 0; i < T * U; ++i) { char b[64]; std::snprintf (b, sizeof (b), "%.16g", (double) mat.a[i]); cell[i] = b; width = std::max (width, cell[i].size ()); }
  for (int r = 0; r < T; ++r)
  {
    for (int c = 0; c < U; ++c) { if (c) strm << " "; strm << std::string (width - cell[r * U + c].size (), ' ') << cell[r * U + c]; }
    strm << "\n";
  }
  strm.flush ();
  strm.precision (old_precision);
}
template <class S, int T, int U> void deserializeASCII (std::istream& strm, Eigen::Mat<S, T, U>* mat)
{
  std::string line = "";
  while (line.length () == 0) getline (strm, line);
  std::istringstream iss (line.substr (1));
  int rows, cols; iss >> rows; iss >> cols;
  for (int y = 0; y < rows; ++y)
  {
    getline (strm, line);
    std::istringstream is2 (line);
    for (int x = 0; x < cols; ++x)
    {
      std::string token; is2 >> token;
      if (token[0] == 'n') mat->coeffRef (y, x) = std::numeric_limits<S>::quiet_NaN ();
      else { std::istringstream buf (token); buf >> mat->coeffRef (y, x); }
    }
  }
}
}
