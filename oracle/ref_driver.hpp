// TEST INFRASTRUCTURE ONLY (oracle/).  Shared declarations of the reference-oracle driver: a thin C API over the
// REFERENCE's own Functional building blocks (ElementRestriction, GeometricFactors, Integral and its kernels,
// GradientAssemblyLookupTables, material models), compiled from /root/reference where they lie.  The only logic restated
// here is the ~60 lines of orchestration that live inside the Functional class itself (functional.hpp:406-498,586-652),
// which cannot be compiled without MPI + hypre.  Product code never includes or links this.
#pragma once

#include <functional>
#include <map>
#include <memory>
#include <string>

#include "serac/numerics/functional/functional.hpp"  // the overlay: reference sub-headers
#include "serac/numerics/functional/shape_aware_functional.hpp"
#include "serac/physics/materials/solid_material.hpp"
#include "serac/physics/materials/thermal_material.hpp"
#include "serac/physics/materials/green_saint_venant_thermoelastic.hpp"

namespace ref {

struct IntegralHolder {
  std::unique_ptr<serac::Integral> integral;
  std::shared_ptr<void>            qdata;            // keeps QuadratureData<State> alive
  void*                            qdata_raw = nullptr;  // &qdata[CUBE or SQUARE](0,0)
  size_t                           state_bytes = 0;
  size_t                           qdata_count = 0;
  int                              num_trials = 0;
};

using Factory = std::function<IntegralHolder(mfem::Mesh& mesh, const double* params, int nparams)>;
std::map<std::string, Factory>& factories();

/// the domain the next factory call integrates over: EntireDomain(mesh), or the elements selected by the mask installed
/// through ref_integral_create_subset (Domain::ofElements, F/domain.cpp:231-306)
serac::Domain domain_for(mfem::Mesh& mesh);

struct Registrar {
  Registrar(const std::string& name, Factory f) { factories()[name] = f; }
};

}  // namespace ref
