"""CPU: face restrictions of discontinuous (L2) spaces -- FaceType::BOUNDARY and FaceType::INTERIOR
(F/element_restriction.cpp:66-215, 304-355).  The product derives the side / orientation tables from the reference vertices
of the geometries (capi.cu: sb_l2_face_dof_info, host only); the fixture tests/golden/l2faces_tables.npz holds, for hex,
tet, quad and tri meshes at p = 0..3, the inputs mfem provides per (face, adjacent element) pair and the reference's own
tables (oracle/make_golden.py: reference_l2_face_tables).  Bit-exact."""
import numpy as np
import pytest

from harness import load_golden

G = load_golden("l2faces_tables")
CASES = [str(c) for c in G["cases"]]


@pytest.mark.parametrize("case", CASES)
def test_dg_face_table_bit_exact(case):
    from serac_b200 import capi
    fg, p = int(G[case + "_face_geom"]), int(G[case + "_p"])
    out = capi.l2_face_dof_info(fg, p, G[case + "_elem_geom"], G[case + "_elem_dofs"], G[case + "_offsets"],
                                G[case + "_side"], G[case + "_orientation"])
    ref = G[case + "_dof_info"]
    # interior faces: the reference stores the two elements of a face in one row of 2 x nodes-per-face entries
    assert np.array_equal(out.reshape(ref.shape), ref)


def test_orientations_are_exercised():
    seen = {3: set(), 2: set(), 1: set()}
    for case in CASES:
        seen[int(G[case + "_face_geom"])].update(int(o) for o in G[case + "_orientation"])
    assert len(seen[3]) >= 3 and len(seen[2]) >= 4 and seen[1] == {0, 1}


def test_bad_side_is_rejected():
    from serac_b200 import capi
    case = CASES[0]
    side = np.array(G[case + "_side"]).copy()
    side[0] = 17
    with pytest.raises(Exception):
        capi.l2_face_dof_info(int(G[case + "_face_geom"]), int(G[case + "_p"]), G[case + "_elem_geom"],
                              G[case + "_elem_dofs"], G[case + "_offsets"], side, G[case + "_orientation"])
