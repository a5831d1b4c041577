// serac_b200 internal helpers shared by the library's translation units (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "serac_b200.h"

namespace sb {

void set_error(const std::string& msg);

#define SB_CHECK_CUDA(expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::sb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                          \
      return SB_ERR_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

#define SB_REQUIRE(cond, msg)       \
  do {                              \
    if (!(cond)) {                  \
      ::sb::set_error(msg);         \
      return SB_ERR_INVALID;        \
    }                               \
  } while (0)

/// a named, compiled-in integrand (registry_*.cu)
struct NamedIntegrand {
  const char*     name;
  sb_kernel_table table;
  /// writes the functor built from `params` into `out`; returns bytes written or -1 on bad parameters
  long (*make_blob)(const double* params, int num_params, void* out, size_t capacity);
  const char* param_doc;
};

std::vector<NamedIntegrand>& registry();

struct Registrar {
  Registrar(const NamedIntegrand& n) { registry().push_back(n); }
};

}  // namespace sb
