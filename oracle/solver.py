"""Oracle solver: deal.II SolverCG recurrences, Multigrid V-cycle with edge matrices,
Ifpack point relaxation (Jacobi / lexicographic SSOR / processor-block SSOR), coarse CG.

Restates src/step-50.cc:938-1017 and the deal.II 9.0 / Trilinos classes it wires together
(SURVEY.md section 3.2, Appendix A8).  TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import math

import numpy as np
import scipy.sparse as sp


class NoConvergence(RuntimeError):
    """SolverControl::NoConvergence."""

    def __init__(self, last_step, last_residual):
        super().__init__(f"Iterative method reported convergence failure in step {last_step}. "
                         f"The residual in the last step was {last_residual}.")
        self.last_step, self.last_residual = last_step, last_residual


def cg(A, b, x0=None, tol=1e-10, max_it=1000, precond=None, history=None):
    """deal.II SolverCG::solve.  Returns (x, iterations, initial residual, last residual)."""
    x = np.zeros_like(b) if x0 is None else x0.copy()
    if x0 is not None and np.any(x):
        g = A @ x - b
    else:
        g = -b.copy()
    res = math.sqrt(g @ g)
    res0 = res
    it = 0
    if res <= tol:
        return x, it, res0, res
    if precond is not None:
        h = precond(g)
        d = -h
        gh = g @ h
    else:
        d = -g
        gh = res * res
    while True:
        it += 1
        h = A @ d
        alpha = gh / (d @ h)
        x += alpha * d
        g += alpha * h
        res = math.sqrt(g @ g)
        if history is not None:
            history.append(res)
        if res <= tol:
            break
        if it >= max_it:
            raise NoConvergence(it, res)
        if precond is not None:
            h = precond(g)
            beta = gh
            gh = g @ h
            beta = gh / beta
            d = beta * d - h
        else:
            beta = gh
            gh = res * res
            beta = gh / beta
            d = beta * d - g
    return x, it, res0, res


# ------------------------------------------------------------------ smoothers
def jacobi(A, omega):
    """TrilinosWrappers::PreconditionJacobi = Ifpack point relaxation 'Jacobi', 1 sweep, zero start."""
    dinv = omega / A.diagonal()
    return lambda r: dinv * r


def ssor(A, omega, blocks=None):
    """TrilinosWrappers::PreconditionSSOR = Ifpack 'symmetric Gauss-Seidel', 1 sweep, zero start:
    forward then backward y_i += omega (r_i - sum_j a_ij y_j) / a_ii over the rank-local rows in local
    order, off-rank y = 0 (overlap 0).  `blocks`: optional array rank-of-row emulating MPI ranks."""
    A = A.tocsr()
    if blocks is not None:
        keep = blocks[np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))] == blocks[A.indices]
        A = sp.csr_matrix((A.data * keep, A.indices, A.indptr), shape=A.shape)
        A.eliminate_zeros()
    dg = A.diagonal()
    L = sp.tril(A, -1).tocsr()
    U = sp.triu(A, 1).tocsr()
    Lo = (L * omega + sp.diags(dg)).tocsr()
    Up = (U * omega + sp.diags(dg)).tocsr()
    from scipy.sparse.linalg import spsolve_triangular

    def apply(r):
        # forward: (D + w L) y = w r
        y = spsolve_triangular(Lo, omega * r, lower=True)
        # backward: y_i += w (r_i - sum_j a_ij y_j)/a_ii with updated upper part
        # (D + w U) z = w r + (1-w) D y - w L y
        return spsolve_triangular(Up, omega * r + (1.0 - omega) * dg * y - omega * (L @ y), lower=False)

    return apply


def ssor_reference_loops(A, omega):
    """Literal double loop of Ifpack_PointRelaxation::ApplyInverseSGS (tiny cases, validates `ssor`)."""
    A = A.tocsr()
    dg = A.diagonal()

    def apply(r):
        y = np.zeros_like(r)
        n = len(r)
        for rng in (range(n), range(n - 1, -1, -1)):
            for i in rng:
                s = A.data[A.indptr[i]:A.indptr[i + 1]] @ y[A.indices[A.indptr[i]:A.indptr[i + 1]]]
                y[i] += omega * (r[i] - s) / dg[i]
        return y

    return apply


class Multigrid:
    """deal.II Multigrid::cycle (V) + PreconditionMG::vmult with MGTransferPrebuilt and edge matrices."""

    def __init__(self, ops, dofs, smoother="ssor", omega=0.5, steps=2, coarse_tol=1e-10, coarse_max=1000,
                 ssor_blocks=None):
        self.ops, self.dofs = ops, dofs
        self.nl = len(ops.A)
        self.steps = steps
        self.coarse_tol, self.coarse_max = coarse_tol, coarse_max
        self.coarse_its = []
        self.sm = [None]
        for l in range(1, self.nl):
            if smoother == "jacobi":
                self.sm.append(jacobi(ops.A[l], omega))
            elif smoother == "ssor":
                self.sm.append(ssor(ops.A[l], omega, None if ssor_blocks is None else ssor_blocks[l]))
            else:
                raise ValueError(smoother)

    def smooth(self, l, u, rhs, zero_start):
        A, S = self.ops.A[l], self.sm[l]
        for s in range(self.steps):
            if zero_start and s == 0:
                u = S(rhs)
            else:
                u = u + S(rhs - A @ u)
        return u

    def vmult(self, src):
        d = self.dofs
        defect = []
        for l in range(self.nl):
            v = np.zeros(d.level_n[l])
            v[d.copy_level[l]] = src[d.copy_global[l]]
            defect.append(v)
        sol = [None] * self.nl

        # run the recursion keeping per-level solutions for copy_from_mg
        def rec(l):
            ops = self.ops
            if l == 0:
                x, it, _, _ = cg(ops.A[0], defect[0], tol=self.coarse_tol, max_it=self.coarse_max)
                self.coarse_its.append(it)
                sol[0] = x
                return
            u = self.smooth(l, None, defect[l], True)
            t = defect[l] - (ops.A[l] @ u + ops.I[l] @ u)
            defect[l - 1] = defect[l - 1] + ops.P[l - 1].T @ t
            rec(l - 1)
            u = u + ops.P[l - 1] @ sol[l - 1]
            defect[l] = defect[l] - ops.I[l].T @ u
            sol[l] = self.smooth(l, u, defect[l], False)

        rec(self.nl - 1)
        dst = np.zeros(d.n)
        for l in range(self.nl):
            dst[d.copy_global[l]] = sol[l][d.copy_level[l]]
        return dst


def pcg_gmg(system_A, b, x0, mg, rel_tol=1e-8, max_it=500):
    """solve() of the reference with PreconditionerType == GMG (src/step-50.cc:942, 991)."""
    tol = rel_tol * math.sqrt(b @ b)
    return cg(system_A, b, x0=x0, tol=tol, max_it=max_it, precond=mg.vmult)


def pcg_jacobi(system_A, b, x0, rel_tol=1e-8, max_it=500, omega=0.6):
    """solve() with PreconditionerType == Jacobi (src/step-50.cc:996-1006)."""
    tol = rel_tol * math.sqrt(b @ b)
    return cg(system_A, b, x0=x0, tol=tol, max_it=max_it, precond=jacobi(system_A, omega))
