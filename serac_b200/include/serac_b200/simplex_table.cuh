// serac_b200: device cache of the simplex shape-function tables (finite_element.hpp: make_simplex_table).
#pragma once

#include <cuda_runtime.h>

#include <mutex>

#include "serac_b200/finite_element.hpp"

namespace serac {
namespace b200 {

/// device copy of make_simplex_table(g, p, Q).tab (weights trail the table); built once per (g, p, Q) per process
inline const double* simplex_table_dev(Geometry g, int p, int Q)
{
  static std::mutex    mtx;
  static const double* cache[3][4][MAX_Q + 1] = {};
  if (!uses_dense_table(g) || p < 0 || p > 3 || Q < 1 || Q > MAX_Q) return nullptr;
  const int                   gi = (g == Geometry::TRIANGLE) ? 0 : (g == Geometry::TETRAHEDRON) ? 1 : 2;
  std::lock_guard<std::mutex> lock(mtx);
  if (!cache[gi][p][Q]) {
    SimplexTable t = make_simplex_table(g, p, Q);
    double*      d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * size_t(t.doubles())) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, t.tab, sizeof(double) * size_t(t.doubles()), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(d);
      return nullptr;
    }
    cache[gi][p][Q] = d;
  }
  return cache[gi][p][Q];
}

}  // namespace b200
}  // namespace serac
