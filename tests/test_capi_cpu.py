"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and its host-side (no GPU needed)
entry points agree with the oracle.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from harness import ROOT, boundary_golden_names, golden_names, load_golden
from oracle import functional_oracle as fo
from serac_b200 import capi


def header_symbols():
    import glob
    syms = set()
    for path in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"typedef[^;{]*\(\*\s*sb_[a-zA-Z0-9_]+\)[^;]*;", "", text)  # function-pointer typedefs are not symbols
        syms |= set(re.findall(r"\b(sb_[a-zA-Z0-9_]+)\s*\(", text))
    return sorted(syms)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(capi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(lib, s), f"libserac_b200.so does not export {s}"
    assert set(syms) <= set(capi.SIGNATURES)  # and the ctypes binding covers the header


def test_version_and_registry():
    lib = capi.lib()
    assert lib.sb_version() >= 100
    names = capi.named_integrands()
    assert "solid_neohookean_hex_p2" in names and "solid_linear_isotropic_hex_p1" in names


@pytest.mark.parametrize("geom", [capi.SEGMENT, capi.TRIANGLE, capi.SQUARE, capi.TETRAHEDRON, capi.CUBE])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_lexicographic_permutation_matches_oracle(geom, p):
    assert np.array_equal(capi.lexicographic_permutation(geom, p), fo.lexicographic_permutation(geom, p))


@pytest.mark.parametrize("name", golden_names())
def test_dof_info_matches_reference(name):
    prob = load_golden(name)
    info = capi.h1_dof_info(int(prob["geom"]), int(prob["p"]), prob["native_dofs"])
    assert np.array_equal(info, prob["dof_info"])


@pytest.mark.parametrize("name", boundary_golden_names())
def test_boundary_face_dof_info_matches_reference(name):
    """GetFaceDofs for H1 (element_restriction.cpp:366-371): native face dofs through the face geometry's permutation"""
    prob = load_golden(name)
    fg, p = int(prob["face_geom"]), int(prob["p"])
    assert np.array_equal(capi.h1_dof_info(fg, p, prob["face_native_dofs"]), prob["face_dof_info"])
    assert np.array_equal(capi.h1_dof_info(fg, 1, prob["shape_face_native_dofs"]), prob["shape_face_dof_info"])
    assert np.array_equal(capi.h1_dof_info(fg, int(prob["mesh_order"]), prob["mesh_face_native_dofs"]),
                          prob["mesh_face_dof_info"])
    assert "solid_pressure_hex_p2" in capi.named_integrands() and "solid_traction_hex_p1" in capi.named_integrands()


def test_unknown_integrand_is_an_error():
    tp = C.POINTER(capi.KernelTable)()
    rc = capi.lib().sb_named_kernel_table(b"no_such_integrand", C.byref(tp))
    assert rc == 3 and b"no_such_integrand" in capi.lib().sb_last_error()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    info = np.zeros((1, 8), dtype=np.uint64)
    with pytest.raises(capi.SeracB200Error):
        capi.ElementRestriction(info, 3, 8)  # needs device memory: must fail loudly, not fall back


def test_abi_headers_are_plain_c():
    """the drop-in boundary is a C ABI: both headers must compile as C99 (no C++ or torch types in the signatures)"""
    import glob
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    for path in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        res = subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", path],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_solver_option_structs_match_the_header():
    """ctypes mirrors of sb_cg_options / sb_cg_result have the C layout (checked against a tiny C program)"""
    import shutil
    import subprocess
    import tempfile
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "serac_b200_solvers.h"\n'
           'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(sb_cg_options), offsetof(sb_cg_options, max_iter),'
           ' offsetof(sb_cg_options, check_every), sizeof(sb_cg_result), offsetof(sb_cg_result, initial_norm),'
           ' offsetof(sb_cg_result, final_norm)); return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "layout.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "layout")
        subprocess.check_call([gcc, "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    want = [C.sizeof(capi.CgOptions), capi.CgOptions.max_iter.offset, capi.CgOptions.check_every.offset,
            C.sizeof(capi.CgResult), capi.CgResult.initial_norm.offset, capi.CgResult.final_norm.offset]
    assert got == want
