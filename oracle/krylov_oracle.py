"""TEST INFRASTRUCTURE ONLY (oracle/): numpy restatement of the reference stack's linear algebra for a Newton step --
the checker for serac_b200/csrc/krylov.cu.  Pinned against the reference's own mfem (oracle/ref_solvers.cpp ->
oracle/_ref/libserac_ref.so) through tests/golden/krylov_*.npz (tests/test_oracle_krylov.py).  Never imported by the
product package.

  eliminate_rows_cols  HypreParMatrix::EliminateRowsCols, mfem/linalg/hypre_parcsr.cpp:476-489
                       (serial twin: SparseMatrix::EliminateRowCol(rc, Ae, DIAG_ONE), mfem/linalg/sparsemat.cpp:2284-2380)
  eliminate_bc         HypreParMatrix::EliminateBC, mfem/linalg/hypre.cpp:2422-2478
  jacobi_dinv          OperatorJacobiSmoother::Setup (dinv = damping / diag), mfem/linalg/solvers.cpp:262-302
  cg                   CGSolver::Mult, mfem/linalg/solvers.cpp:719-897
  gmres                GMRESSolver::Mult, mfem/linalg/solvers.cpp:981-1165 (+ plane rotations / Update, :931-980)
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


def _generate_rotation(dx, dy):  # solvers.cpp:931-951
    if dy == 0.0:
        return 1.0, 0.0
    if abs(dy) > abs(dx):
        t = dx / dy
        sn = 1.0 / np.sqrt(1.0 + t * t)
        return t * sn, sn
    t = dy / dx
    cs = 1.0 / np.sqrt(1.0 + t * t)
    return cs, t * cs


def _apply_rotation(dx, dy, cs, sn):  # :953-958
    return cs * dx + sn * dy, -sn * dx + cs * dy


def _update(x, k, H, s, v):  # :960-980
    y = s[:k + 1].copy()
    for i in range(k, -1, -1):
        y[i] /= H[i, i]
        for j in range(i - 1, -1, -1):
            y[j] -= H[j, i] * y[i]
    for i in range(k + 1):
        x = x + y[i] * v[i]
    return x


def gmres(apply_A, b, dinv=None, kdim=50, rel_tol=1e-12, abs_tol=0.0, max_iter=1000, x0=None):
    """GMRESSolver::Mult, statement by statement (solvers.cpp:981-1165): restarted GMRES(kdim), modified Gram-Schmidt,
    left preconditioner.  Returns x, dict(converged, final_iter, initial_norm, final_norm)."""
    b = np.asarray(b, dtype=np.float64)
    m = kdim
    M = (lambda v: dinv * v) if dinv is not None else (lambda v: v)
    if x0 is not None:  # :995-1026
        x = np.array(x0, dtype=np.float64)
        r = M(b - apply_A(x))
    else:
        x = np.zeros_like(b)
        r = M(b.copy())
    beta = float(np.sqrt(r @ r))  # :1027
    info = dict(converged=False, final_iter=0, initial_norm=beta, final_norm=0.0)
    final_norm = max(rel_tol * beta, abs_tol)  # :1030
    if beta <= final_norm:  # :1032-1039
        info.update(converged=True, final_norm=beta)
        return x, info
    H = np.zeros((m + 1, m))
    cs, sn = np.zeros(m + 1), np.zeros(m + 1)
    v = [None] * (m + 1)
    j = 1
    while j <= max_iter:  # :1053
        v[0] = r / beta  # :1056
        s = np.zeros(m + 1)
        s[0] = beta
        i = 0
        while i < m and j <= max_iter:  # :1059
            w = M(apply_A(v[i]))  # :1061-1069
            for k in range(i + 1):  # :1071-1075
                H[k, i] = float(w @ v[k])
                w = w - H[k, i] * v[k]
            H[i + 1, i] = float(np.sqrt(w @ w))  # :1077
            v[i + 1] = w / H[i + 1, i]  # :1080
            for k in range(i):  # :1082-1085
                H[k, i], H[k + 1, i] = _apply_rotation(H[k, i], H[k + 1, i], cs[k], sn[k])
            cs[i], sn[i] = _generate_rotation(H[i, i], H[i + 1, i])
            H[i, i], H[i + 1, i] = _apply_rotation(H[i, i], H[i + 1, i], cs[i], sn[i])
            s[i], s[i + 1] = _apply_rotation(s[i], s[i + 1], cs[i], sn[i])
            resid = abs(s[i + 1])  # :1091
            if resid <= final_norm:  # :1094-1101
                x = _update(x, i, H, s, v)
                info.update(converged=True, final_iter=j, final_norm=resid)
                return x, info
            i += 1
            j += 1
        x = _update(x, i - 1, H, s, v)  # :1117
        r = M(b - apply_A(x))  # :1119-1128
        beta = float(np.sqrt(r @ r))
        if beta <= final_norm:  # :1131-1137
            info.update(converged=True, final_iter=j, final_norm=beta)
            return x, info
    info.update(converged=False, final_iter=max_iter, final_norm=beta)  # :1140-1142
    return x, info
