"""GPU: the mfem-typed boundary (serac_b200/functional_mfem.hpp: serac::Functional over mfem::FiniteElementSpace /
mfem::Vector, Gradient as an mfem::Operator, assemble -> mfem::SparseMatrix, ShapeAwareFunctional).  The driver
tests/cpp/functional_mfem_test.cu is the reference's own unit tests ported with minimal edits (functional_comparisons.cpp
against MFEM's integrators, functional_basic_h1_vector.cpp + check_gradient.hpp), linked statically against the
reference's serial mfem, which is test infrastructure only."""
import os
import subprocess

import pytest

from harness import ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "tests", "cpp", "functional_mfem_test")


@pytest.mark.parametrize("suite", ["comparisons", "l2", "gradients", "scalar", "domains", "qoi", "nonlinear", "boundary", "shape"])
def test_reference_unit_tests_through_the_mfem_typed_functional(suite):
    if not os.path.exists(BIN):
        pytest.skip("tests/cpp/functional_mfem_test not built (needs oracle/_ref/mfem-install: make -C tests/cpp)")
    out = subprocess.run([BIN, suite], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "FAIL" not in out.stdout and " 0 failure(s)" in out.stdout, \
        out.stdout[-3000:] + out.stderr[-2000:]
    assert out.stdout.count("ok  ") >= 4
