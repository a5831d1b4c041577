// TEST INFRASTRUCTURE (oracle/): overlay found before the reference's functional.hpp.  The reference's Functional CLASS
// needs mfem::ParFiniteElementSpace + HypreParMatrix (MPI + hypre, absent here); everything else in that header is its
// include list, which is what the reference's material headers (physics/materials/solid_material.hpp:15) want.  This
// file includes the reference's own sub-headers from where they lie and adds the one small type they expect.
#pragma once
#include "mfem.hpp"
#include "serac/serac_config.hpp"
#include "serac/infrastructure/logger.hpp"
#include "serac/numerics/functional/tensor.hpp"
#include "serac/numerics/functional/quadrature.hpp"
#include "serac/numerics/functional/finite_element.hpp"
#include "serac/numerics/functional/integral.hpp"
#include "serac/numerics/functional/dof_numbering.hpp"
#include "serac/numerics/functional/differentiate_wrt.hpp"
#include "serac/numerics/functional/element_restriction.hpp"
#include "serac/numerics/functional/domain.hpp"
#include <array>
#include <vector>
namespace serac {
template <int... i>
struct DependsOn {
};
}  // namespace serac
namespace serac {
// forward declaration so that shape_aware_functional.hpp (which only *names* Functional inside a class template the
// oracle never instantiates) parses
template <typename T, ExecutionSpace exec = ExecutionSpace::CPU>
class Functional;
}  // namespace serac
