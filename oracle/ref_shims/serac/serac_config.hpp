// TEST INFRASTRUCTURE (oracle/): hand-written equivalent of the reference's generated serac_config.hpp
// (src/serac/serac_config.hpp.in:61-66) with the default byVDIM ordering (cmake/SeracBasics.cmake:66).
#pragma once
#define SERAC_USE_MFEM
#define SERAC_USE_VDIM_ORDERING
#ifndef SERAC_REPO_DIR
#define SERAC_REPO_DIR "/root/reference"
#endif
#include "mfem.hpp"
namespace serac {
constexpr mfem::Ordering::Type ordering = mfem::Ordering::byVDIM;
}
