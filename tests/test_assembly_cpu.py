"""CPU: the sparsity pattern and the index tables of the fused assembly (serac_b200/csrc/assembly.cu) against the
reference's GradientAssemblyLookupTables output and assembled values in tests/golden (written by oracle/make_golden.py
from the reference's own kernels).  Host entry points only -- no device, no arithmetic of the product path."""
import numpy as np
import pytest

from harness import golden_names, load_golden, multifield_golden_names
from serac_b200 import capi


def _table(info):
    return (np.asarray(info, dtype=np.uint64) & np.uint64(0x0000FFFFFFFFFFFF)).astype(np.int32)


def brute_force_pattern(tt, ct, NT, t_ord, kt, ck, NK, k_ord):
    """F/dof_numbering.hpp:290-351 literally: every (row, col) pair of every element, sorted and uniqued"""
    row_of = (lambda n, a: n * ct + a) if t_ord == capi.BYVDIM else (lambda n, a: a * NT + n)
    col_of = (lambda m, b: m * ck + b) if k_ord == capi.BYVDIM else (lambda m, b: b * NK + m)
    pairs = set()
    for e in range(tt.shape[0]):
        for n in tt[e]:
            for a in range(ct):
                for m in kt[e]:
                    for b in range(ck):
                        pairs.add((row_of(int(n), a), col_of(int(m), b)))
    pairs = sorted(pairs)
    rp = np.zeros(NT * ct + 1, dtype=np.int32)
    for r, _ in pairs:
        rp[r + 1] += 1
    return np.cumsum(rp).astype(np.int32), np.array([c for _, c in pairs], dtype=np.int32)


@pytest.mark.parametrize("name", golden_names())
def test_pattern_is_bit_exact(name):
    prob = load_golden(name)
    if "row_ptr" not in prob:
        pytest.skip("fixture has no assembled matrix")
    tt = _table(prob["dof_info"])
    c, nn = int(prob["vdim"]), int(prob["num_nodes"])
    rp, ci = capi.csr_pattern_host(tt, c, nn, tt, c, nn)
    assert np.array_equal(rp, prob["row_ptr"]) and np.array_equal(ci, prob["col_ind"])


@pytest.mark.parametrize("t_ord", [capi.BYVDIM, capi.BYNODES])
@pytest.mark.parametrize("k_ord", [capi.BYVDIM, capi.BYNODES])
def test_pattern_orderings_and_rectangular(t_ord, k_ord):
    """test space H1 p=2 (3 components) x trial space H1 p=1 (2 components) on the 7-hex patch, every ordering pair"""
    prob = load_golden("solid_neohookean_hex_p2_patch")
    tt, kt = _table(prob["dof_info"]), _table(prob["shape_dof_info"])
    NT, NK = int(prob["num_nodes"]), int(prob["shape_num_nodes"])
    rp, ci = capi.csr_pattern_host(tt, 3, NT, kt, 2, NK, t_ord, k_ord)
    rp_ref, ci_ref = brute_force_pattern(tt, 3, NT, t_ord, kt, 2, NK, k_ord)
    assert np.array_equal(rp, rp_ref) and np.array_equal(ci, ci_ref)


@pytest.mark.parametrize("acc_cap,trial_blocks", [(1600, 1), (400, 2), (800, 3)])
@pytest.mark.parametrize("name", golden_names())
def test_assembly_plan_reproduces_reference_values(name, acc_cap, trial_blocks):
    """the rows-as-units walk over the reference's element matrices gives the reference's CSR values: same sums, each in
    ascending element order (F/functional.hpp:610-644), for whole meshes and sub-domains; a small accumulator cap cuts
    the groups and must change nothing"""
    prob = load_golden(name)
    if "K_e" not in prob or "values" not in prob:
        pytest.skip("fixture has no assembled matrix")
    tt = _table(prob["dof_info"])
    c, nn = int(prob["vdim"]), int(prob["num_nodes"])
    K_e, dom = prob["K_e"], None
    if "domain_elements" in prob:
        dom = prob["domain_elements"].astype(np.int32)
        K_e = K_e[dom] if K_e.shape[0] == tt.shape[0] else K_e
    need = int(np.diff(prob["row_ptr"]).max())
    values, ng = capi.debug_assembly_plan_host(tt, c, nn, tt, c, nn, K_e, prob["row_ptr"], len(prob["values"]),
                                               elements=dom, acc_cap=max(acc_cap, need),
                                               trial_blocks=min(trial_blocks, c))
    assert ng >= 1
    ref = prob["values"]
    assert np.max(np.abs(values - ref)) <= 1e-15 * max(1.0, np.max(np.abs(ref))) * 8
    if dom is None and acc_cap == 1600:
        assert np.array_equal(values, ref)  # same order of additions as the reference's loop: same bits


def test_plan_with_bynodes_orderings():
    """byNODES on either side only moves entries inside the rows"""
    prob = load_golden("solid_neohookean_hex_p2_patch")
    tt = _table(prob["dof_info"])
    nn, c = int(prob["num_nodes"]), 3
    rp0, ci0 = capi.csr_pattern_host(tt, c, nn, tt, c, nn)
    dense = np.zeros((nn * c, nn * c))
    for r in range(nn * c):
        dense[r, ci0[rp0[r]:rp0[r + 1]]] = prob["values"][rp0[r]:rp0[r + 1]]
    for t_ord in (capi.BYVDIM, capi.BYNODES):
        for k_ord in (capi.BYVDIM, capi.BYNODES):
            rp, ci = capi.csr_pattern_host(tt, c, nn, tt, c, nn, t_ord, k_ord)
            vals, _ = capi.debug_assembly_plan_host(tt, c, nn, tt, c, nn, prob["K_e"], rp, len(ci), test_ordering=t_ord,
                                                    trial_ordering=k_ord)
            row_perm = np.arange(nn * c) if t_ord == capi.BYVDIM else \
                np.array([(r % nn) * c + r // nn for r in range(nn * c)])
            col_perm = np.arange(nn * c) if k_ord == capi.BYVDIM else \
                np.array([(q % nn) * c + q // nn for q in range(nn * c)])
            for r in range(nn * c):
                got = vals[rp[r]:rp[r + 1]]
                want = dense[row_perm[r], col_perm[ci[rp[r]:rp[r + 1]]]]
                assert np.allclose(got, want, rtol=0, atol=1e-13 * np.abs(dense).max())
