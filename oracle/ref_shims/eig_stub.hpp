// TEST INFRASTRUCTURE (oracle/): mfem::DenseMatrixEigensystem only exists with LAPACK; the reference mentions it in
// eigenvalues() (tuple_tensor_dual_functions.hpp:737-748), which is off the Functional path.  Declaration-only stub.
#pragma once
#include "mfem.hpp"
#ifndef MFEM_USE_LAPACK
namespace mfem {
class DenseMatrixEigensystem {
public:
  DenseMatrixEigensystem(DenseMatrix&) {}
  void    Eval() {}
  Vector& Eigenvalues() { return ev_; }
  double  Eigenvalue(int) { return 0.0; }
  Vector& Eigenvector(int) { return ev_; }

private:
  Vector ev_;
};
}  // namespace mfem
#endif
#ifndef MFEM_USE_MPI
// the serial MFEM build only forward-declares ParFiniteElementSpace; the reference's legacy DofNumbering helpers
// (dof_numbering.hpp:88-239, unused by Functional) mention it by value.  At one rank it IS a FiniteElementSpace.
namespace mfem {
class ParFiniteElementSpace : public FiniteElementSpace {
public:
  using FiniteElementSpace::FiniteElementSpace;
};
}  // namespace mfem
#endif
