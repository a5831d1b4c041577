"""TEST INFRASTRUCTURE ONLY (oracle/): numpy restatement of the reference stack's linear algebra for a Newton step --
the checker for serac_b200/csrc/krylov.cu.  Pinned against the reference's own mfem (oracle/ref_solvers.cpp ->
oracle/_ref/libserac_ref.so) through tests/golden/krylov_*.npz (tests/test_oracle_krylov.py).  Never imported by the
product package.

  eliminate_rows_cols  HypreParMatrix::EliminateRowsCols, mfem/linalg/hypre_parcsr.cpp:476-489
                       (serial twin: SparseMatrix::EliminateRowCol(rc, Ae, DIAG_ONE), mfem/linalg/sparsemat.cpp:2284-2380)
  eliminate_bc         HypreParMatrix::EliminateBC, mfem/linalg/hypre.cpp:2422-2478
  jacobi_dinv          OperatorJacobiSmoother::Setup (dinv = damping / diag), mfem/linalg/solvers.cpp:262-302
  cg                   CGSolver::Mult, mfem/linalg/solvers.cpp:719-897
"""
import numpy as np


def csr_rows(row_ptr):
    """row index of every nonzero"""
    row_ptr = np.asarray(row_ptr)
    return np.repeat(np.arange(len(row_ptr) - 1), np.diff(row_ptr))


def spmv(row_ptr, col_ind, values, x):
    """y = A x (SparseMatrix::AddMult, mfem/linalg/sparsemat.cpp:755-800); numpy's row sums, so the last bits may differ
    from mfem's left-to-right sums"""
    prod = np.concatenate([values * x[col_ind], [0.0]])
    nonempty = np.diff(row_ptr) > 0
    y = np.zeros(len(row_ptr) - 1)
    y[nonempty] = np.add.reduceat(prod, np.asarray(row_ptr[:-1])[nonempty])
    return y


def diagonal(row_ptr, col_ind, values):
    rows = csr_rows(row_ptr)
    d = np.zeros(len(row_ptr) - 1)
    on = rows == col_ind
    d[rows[on]] = values[on]
    return d


def eliminate_rows_cols(row_ptr, col_ind, values, ess):
    """A -> [A_ii 0; 0 I] (returned first), Ae = [0 A_ib; A_bi A_bb - I] (returned second), both in A's pattern
    (hypre_parcsr.cpp:476-489)"""
    n = len(row_ptr) - 1
    mark = np.zeros(n, dtype=bool)
    mark[np.asarray(ess, dtype=np.int64)] = True
    rows = csr_rows(row_ptr)
    elim = mark[rows] | mark[col_ind]
    keep = np.where(rows == col_ind, 1.0, 0.0)
    A = np.where(elim, keep, values)
    Ae = np.where(elim, values - keep, 0.0)
    return A, Ae


def eliminate_bc(row_ptr, col_ind, values, values_e, ess, X, B):
    """b -= Ae x; b(r) = A(r,r) x(r) for r in ess (hypre.cpp:2426-2451)"""
    B = B - spmv(row_ptr, col_ind, values_e, X)
    d = diagonal(row_ptr, col_ind, values)
    ess = np.asarray(ess, dtype=np.int64)
    B[ess] = d[ess] * X[ess]
    return B


def jacobi_dinv(diag, damping=1.0):
    return damping / diag


def cg(apply_A, b, dinv=None, rel_tol=1e-12, abs_tol=0.0, max_iter=1000, x0=None):
    """CGSolver::Mult, statement by statement (solvers.cpp:719-897).  apply_A: callable x -> A x.
    Returns x, dict(converged, final_iter, initial_norm, final_norm)."""
    b = np.asarray(b, dtype=np.float64)
    if x0 is not None:  # iterative_mode (:725-729)
        x = np.array(x0, dtype=np.float64)
        r = b - apply_A(x)
    else:  # :732-733
        x = np.zeros_like(b)
        r = b.copy()
    if dinv is not None:  # :736-744
        z = dinv * r
        d = z.copy()
    else:
        d = r.copy()
    nom0 = nom = float(d @ r)  # :745
    info = dict(converged=False, final_iter=0, initial_norm=np.sqrt(nom0) if nom0 >= 0 else nom0, final_norm=0.0)
    if nom < 0.0:  # :755-767
        info.update(initial_norm=nom, final_norm=nom)
        return x, info
    r0 = max(nom * rel_tol * rel_tol, abs_tol * abs_tol)  # :768
    if nom <= r0:  # :769-775
        info.update(converged=True, final_norm=np.sqrt(nom))
        return x, info
    z = apply_A(d)  # :777
    den = float(z @ d)  # :778
    if den == 0.0:  # :787-793
        info.update(final_norm=np.sqrt(nom))
        return x, info
    info["final_iter"] = max_iter  # :798
    i = 1
    betanom = nom
    while True:
        alpha = nom / den  # :801
        x = x + alpha * d  # :802
        r = r - alpha * z  # :803
        if dinv is not None:  # :805-813
            z = dinv * r
            betanom = float(r @ z)
        else:
            betanom = float(r @ r)
        if betanom < 0.0:  # :815-825
            info.update(converged=False, final_iter=i)
            break
        if betanom <= r0:  # :835-840
            info.update(converged=True, final_iter=i)
            break
        i += 1
        if i > max_iter:  # :842-845
            break
        beta = betanom / nom  # :847
        d = (z if dinv is not None else r) + beta * d  # :848-855
        z = apply_A(d)  # :856
        den = float(d @ z)  # :857
        if den == 0.0:  # :866-870
            info["final_iter"] = i
            break
        nom = betanom  # :872
    info["final_norm"] = np.sqrt(betanom)  # :894
    return x, info
