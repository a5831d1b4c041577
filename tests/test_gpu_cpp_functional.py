"""-m gpu: the C++ host side (serac_b200/include/serac_b200/functional.hpp, serac::Functional<test(trials...)> with the
reference's API) driven by tests/cpp/functional_test on problems exported from the golden fixtures; its outputs are
compared here with the reference's (tolerance 1e-12 relative L2, CSR pattern bit-exact).  The `user` case runs a
q-function that exists only in the test driver and is checked there against finite differences and its own matrix."""
import os
import subprocess

import numpy as np
import pytest

from harness import ROOT, load_golden, rel_l2
from oracle import functional_oracle as fo

pytestmark = pytest.mark.gpu
TOL = 1e-12
EXE = os.path.join(ROOT, "tests", "cpp", "functional_test")
MASK = np.uint64(0x0000FFFFFFFFFFFF)


def export(prob, d, boundary=False):
    """raw arrays + meta.txt the driver reads"""
    geom, p, dim = int(prob["geom"]), int(prob["p"]), int(prob["dim"])
    mesh_order = int(prob["mesh_order"])
    meta = dict(geom=geom, p=p, dim=dim, ne=int(prob["ne"]), num_nodes=int(prob["num_nodes"]),
                shape_num_nodes=int(prob["shape_num_nodes"]), mesh_order=mesh_order,
                mesh_ordering=int(prob["mesh_ordering"]), mesh_num_nodes=int(prob["mesh_num_nodes"]))
    if boundary:
        meta["nf"] = int(prob["nf"])
    with open(os.path.join(d, "meta.txt"), "w") as f:
        for k, v in meta.items():
            f.write(f"{k} {v}\n")
    arrays = dict(native_dofs=prob["native_dofs"].astype(np.int32), params=prob["params"].astype(np.float64),
                  mesh_nodes=prob["mesh_nodes"], U=prob["U"], A=prob["A"], S=prob["S"], dU=prob["dU"])
    if boundary:
        arrays.update(face_native_dofs=prob["face_native_dofs"].astype(np.int32),
                      shape_face_native_dofs=prob["shape_face_native_dofs"].astype(np.int32),
                      mesh_face_native_dofs=prob["mesh_face_native_dofs"].astype(np.int32), dS=prob["dS"])
        # the fixture's mesh-nodes space has the shape space's numbering (both H1 p=1 on the same mesh)
        assert mesh_order == 1
        arrays["shape_native_dofs"] = arrays["mesh_native_dofs"] = recover_native(prob, geom)
    else:
        arrays["shape_native_dofs"] = prob["shape_native_dofs"].astype(np.int32)
        # native dofs of the mesh-nodes space: undo the lexicographic permutation of its dof_info table
        lex = fo.lexicographic_permutation(geom, mesh_order)
        native = np.empty(prob["mesh_dof_info"].shape, dtype=np.int32)
        native[:, lex] = (prob["mesh_dof_info"] & MASK).astype(np.int32)
        arrays["mesh_native_dofs"] = native
    if "domain_elements" in prob:
        arrays["domain_elements"] = prob["domain_elements"].astype(np.int32)
        with open(os.path.join(d, "meta.txt"), "a") as f:
            f.write(f"domain_elements {len(prob['domain_elements'])}\n")
    for k, a in arrays.items():
        np.ascontiguousarray(a).tofile(os.path.join(d, k + ".bin"))


def recover_native(prob, geom):
    """boundary fixtures do not store the p = 1 element dofs: the vertices are the first 8 (hex) native dofs of every
    element of any H1 space on the mesh"""
    nv = {5: 8, 4: 4}[geom]
    return np.ascontiguousarray(prob["native_dofs"][:, :nv]).astype(np.int32)


def run(case, d):
    assert os.path.exists(EXE), "tests/cpp/functional_test is not built (python __graft_entry__.py)"
    res = subprocess.run([EXE, case, d], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "launches: 0" not in res.stdout
    return res.stdout


def out(d, name, dtype=np.float64):
    return np.fromfile(os.path.join(d, f"out_{name}.bin"), dtype=dtype)


@pytest.mark.parametrize("name", ["solid_neohookean_hex_p2_patch", "solid_neohookean_tet_p2_cart",
                                  "solid_neohookean_hex_p2_subdomain"])
def test_solid_functional(name, tmp_path, cuda):
    prob = load_golden(name)
    d = str(tmp_path)
    export(prob, d)
    run("solid", d)
    assert rel_l2(out(d, "r_plain"), prob["r"]) < TOL
    assert rel_l2(out(d, "r"), prob["r"]) < TOL
    assert rel_l2(out(d, "dr"), prob["dr"]) < TOL
    assert rel_l2(out(d, "dr_da"), prob["dr_da"]) < TOL
    assert np.array_equal(out(d, "row_ptr", np.int32), prob["row_ptr"])
    assert np.array_equal(out(d, "col_ind", np.int32), prob["col_ind"])
    assert rel_l2(out(d, "values"), prob["values"]) < TOL


def test_j2_state_functional(tmp_path, cuda):
    prob = load_golden("solid_j2smallstrain_linear_hex_p2_cart")
    d = str(tmp_path)
    export(prob, d)
    run("j2", d)
    assert rel_l2(out(d, "r"), prob["r"]) < TOL
    assert rel_l2(out(d, "dr"), prob["dr"]) < TOL
    assert rel_l2(out(d, "qdata1"), np.asarray(prob["qdata1"]).ravel()) < TOL
    want = fo.scatter_add(prob["dof_info"] & MASK, prob["r_E_after_commit"], np.zeros(len(prob["U"])), 3)
    assert rel_l2(out(d, "r_after_commit"), want) < TOL


def test_j2_finite_strain_default_state_and_checkpoint_restart(tmp_path, cuda):
    """J2::State{} is the virgin material (Fpinv = I, solid_material.hpp:273); the state buffer survives a
    to_host / from_host round trip into a fresh Functional (QuadratureData checkpointing)"""
    prob = load_golden("solid_j2_linear_hex_p2_cart")
    q0 = np.asarray(prob["qdata0"]).reshape(-1, 10)
    assert np.array_equal(q0[:, :9], np.tile(np.eye(3).ravel(), (len(q0), 1))) and not q0[:, 9].any()
    d = str(tmp_path)
    export(prob, d)
    run("j2_finite", d)
    assert rel_l2(out(d, "r"), prob["r"]) < TOL
    assert rel_l2(out(d, "dr"), prob["dr"]) < TOL
    assert rel_l2(out(d, "qdata1"), np.asarray(prob["qdata1"]).ravel()) < TOL
    want = fo.scatter_add(prob["dof_info"] & MASK, prob["r_E_after_commit"], np.zeros(len(prob["U"])), 3)
    assert rel_l2(out(d, "r_after_commit"), want) < TOL
    assert np.array_equal(out(d, "r_after_restart"), out(d, "r_after_commit"))  # same state, same bits


def test_boundary_pressure_functional(tmp_path, cuda):
    prob = load_golden("bdr_solid_pressure_hex_p2_cart")
    d = str(tmp_path)
    export(prob, d, boundary=True)
    run("pressure", d)
    assert rel_l2(out(d, "r"), prob["r"]) < TOL
    assert rel_l2(out(d, "dr"), prob["dr"]) < TOL
    assert rel_l2(out(d, "dr_dshape"), prob["dr_dshape"]) < TOL


def test_user_defined_qfunction(tmp_path, cuda):
    prob = load_golden("solid_neohookean_hex_p2_patch")
    d = str(tmp_path)
    export(prob, d)
    stdout = run("user", d)
    assert "PASS" in stdout


def test_qoi_functional(tmp_path, cuda):
    """Functional<double(H1<2,3>)> (functional_qoi.inl) with a q-function defined in the C++ driver"""
    prob = load_golden("qoi_quadratic_hex_p2_cart")
    d = str(tmp_path)
    fake = dict(prob)
    # the QOI fixture has one trial field; the driver's loader also wants the shape space and the other inputs
    lex = fo.lexicographic_permutation(int(prob["geom"]), int(prob["mesh_order"]))
    native = np.empty(prob["mesh_dof_info"].shape, dtype=np.int32)
    native[:, lex] = (prob["mesh_dof_info"] & MASK).astype(np.int32)
    fake.update(shape_native_dofs=native, shape_num_nodes=prob["mesh_num_nodes"], A=prob["U"], S=np.zeros(1))
    export(fake, d)
    run("qoi", d)
    v = out(d, "value")
    assert abs(v[0] - prob["value"]) < TOL * abs(prob["value"]) and v[0] == v[1]
    assert abs(out(d, "dvalue")[0] - prob["dvalue"]) < TOL * abs(prob["dvalue"])
    assert rel_l2(out(d, "gradient"), prob["gradient"]) < TOL


def test_boundary_qoi_functional(tmp_path, cuda):
    """Functional<double(H1<2,3>)>::AddBoundaryIntegral with a q-function defined in the C++ driver"""
    prob = load_golden("bqoi_quadratic_hex_p2_cart")
    d = str(tmp_path)
    fake = dict(prob)
    nv = recover_native(prob, int(prob["geom"]))
    fake.update(shape_num_nodes=prob["mesh_num_nodes"], A=prob["U"], S=np.zeros(1), dS=np.zeros(1),
                shape_face_native_dofs=prob["mesh_face_native_dofs"])
    export(fake, d, boundary=True)
    run("boundary_qoi", d)
    assert abs(out(d, "value")[0] - prob["value"]) < TOL * abs(prob["value"])
    assert abs(out(d, "dvalue")[0] - prob["dvalue"]) < TOL * abs(prob["dvalue"])
    assert rel_l2(out(d, "gradient"), prob["gradient"]) < TOL
