"""Oracle assembly: system matrix + rhs with constraints, level / interface matrices, transfers.

Restates src/step-50.cc:735-833 (`assemble_system`), :835-933 (`assemble_multigrid`) and
deal.II's `ConstraintMatrix::distribute_local_to_global`, `make_sparsity_pattern`,
`MGTools::make_sparsity_pattern`, `MGTransferPrebuilt::build_matrices`
(SURVEY.md Appendix A4, A5, A7).  TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import numpy as np
import scipy.sparse as sp

from . import fe


def _coo(n_rows, n_cols, cell_rows, cell_cols, cell_vals):
    """Sum element matrices: cell_rows (nc, r), cell_cols (nc, c), cell_vals (nc, r, c) or (r, c)."""
    nc, r = cell_rows.shape
    c = cell_cols.shape[1]
    rows = np.repeat(cell_rows, c, axis=1).ravel()
    cols = np.tile(cell_cols, (1, r)).ravel()
    vals = np.broadcast_to(cell_vals, (nc, r, c)).ravel()
    return sp.coo_matrix((vals, (rows, cols)), shape=(n_rows, n_cols)).tocsr()


def embed(pattern, values):
    """CSR with the entries of `pattern` (explicit zeros kept) and the values of `values`."""
    pattern = pattern.tocsr()
    pattern.sort_indices()
    values = values.tocoo()
    n = pattern.shape[1]
    rows = np.repeat(np.arange(pattern.shape[0], dtype=np.int64), np.diff(pattern.indptr))
    pkey = rows * n + pattern.indices
    vkey = values.row.astype(np.int64) * n + values.col
    pos = np.searchsorted(pkey, vkey)
    assert (pkey[np.minimum(pos, len(pkey) - 1)] == vkey).all(), "value outside pattern"
    data = np.zeros(len(pkey))
    np.add.at(data, pos, values.data)
    return sp.csr_matrix((data, pattern.indices.copy(), pattern.indptr.copy()), shape=pattern.shape)


def cell_matrices(forest, l, idx=None, coef_fn=None):
    """Cell stiffness matrices of level-l cells; (8, 8) if the coefficient is 1, else (nc, 8, 8)."""
    dim = forest.dim
    h = forest.h(l)
    if coef_fn is None:
        return fe.stiffness(h, dim)
    G, pts = fe.stiffness_q(dim)
    ijk = forest.ijk[l] if idx is None else forest.ijk[l][idx]
    org = forest.lo + ijk * h
    xq = org[:, None, :] + h * pts[None, :, :]
    return fe.stiffness(h, dim, coef_fn(xq))


class System:
    """Active-mesh system: matrix (constraints condensed, deal.II style), pattern, rhs."""

    def __init__(self, forest, dofs, coef_fn=None):
        self.forest, self.dofs = forest, dofs
        n = dofs.n
        blocks, absdiag = [], np.zeros(n)
        ones = []
        for l in range(forest.n_levels):
            cd = dofs.cell_dofs[l]
            if len(cd) == 0:
                continue
            K = cell_matrices(forest, l, dofs.active_cells[l], coef_fn)
            blocks.append(_coo(n, n, cd, cd, K))
            ones.append(_coo(n, n, cd, cd, np.ones((cd.shape[1], cd.shape[1]))))
            diag = np.broadcast_to(np.abs(np.diagonal(K, axis1=-2, axis2=-1)), cd.shape)
            np.add.at(absdiag, cd.ravel(), diag.ravel())
        self.A_raw = sum(blocks[1:], blocks[0]).tocsr()
        B = sum(ones[1:], ones[0]).tocsr()
        self.absdiag = absdiag
        # x = T x~ + g^ : hanging rows interpolate their parents
        nh = ~dofs.hanging
        idn = np.nonzero(nh)[0]
        self.T = sp.coo_matrix(
            (np.concatenate([np.ones(len(idn)), dofs.hang_vals]),
             (np.concatenate([idn, dofs.hang_rows]), np.concatenate([idn, dofs.hang_cols]))), shape=(n, n)).tocsr()
        free = ~dofs.constrained
        Df = sp.diags(free.astype(float))
        Z = dofs.constrained
        A1 = (self.T.T @ self.A_raw @ self.T).tocsr()
        self.A = (Df @ A1 @ Df + sp.diags(absdiag * Z)).tocsr()
        self.A.eliminate_zeros()
        # sparsity: per cell {dofs}^2 U {resolved dofs}^2 (make_sparsity_pattern(dof, dsp, constraints, true))
        Tf = (self.T @ Df).tocsr()
        Tf.data[:] = 1.0
        pat = (B + Tf.T @ B @ Tf).tocsr()
        pat.data[:] = 1.0
        self.pattern = pat
        self.A_stored = embed(pat, self.A)

    def rhs(self, f_raw, g):
        """f_raw: unconstrained load vector; g: Dirichlet values on dofs.dirichlet (0 elsewhere)."""
        Z = self.dofs.constrained
        ghat = self.T @ g
        b = self.T.T @ (f_raw - self.A_raw @ ghat)
        b[Z] = 0.0
        return b

    def distribute(self, x, g):
        """constraints.distribute (src/step-50.cc:1016)."""
        xt = x.copy()
        xt[self.dofs.constrained] = 0.0
        return self.T @ (xt + g)


def load_vector(forest, dofs, dens, nq):
    """cell_rhs(i) = sum_q phi_i(x_q) rho_q JxW_q (src/step-50.cc:813-820); dens: per level (n_active_l, nq^dim)."""
    dim = forest.dim
    pts, wts = fe.tensor_rule(nq, dim)
    N = fe.shape_values(pts, dim)
    f = np.zeros(dofs.n)
    for l in range(forest.n_levels):
        cd = dofs.cell_dofs[l]
        if len(cd) == 0:
            continue
        jxw = wts * forest.h(l) ** dim
        np.add.at(f, cd.ravel(), ((dens[l] * jxw) @ N).ravel())
    return f


class LevelOps:
    """mg_matrices, mg_interface_matrices, prolongation matrices and copy indices."""

    def __init__(self, forest, dofs, coef_fn=None):
        self.forest, self.dofs = forest, dofs
        dim = forest.dim
        self.A, self.I, self.P, self.A_stored, self.pattern = [], [], [], [], []
        for l in range(forest.n_levels):
            n = dofs.level_n[l]
            cd = dofs.level_cell_dofs[l]
            K = cell_matrices(forest, l, None, coef_fn)
            raw = _coo(n, n, cd, cd, K)
            pat = _coo(n, n, cd, cd, np.ones((cd.shape[1], cd.shape[1])))
            pat.data[:] = 1.0
            absdiag = np.zeros(n)
            np.add.at(absdiag, cd.ravel(), np.broadcast_to(np.abs(np.diagonal(K, axis1=-2, axis2=-1)), cd.shape).ravel())
            edge, bd = dofs.level_edge[l], dofs.level_boundary[l]
            Z = edge | bd
            Dn = sp.diags((~Z).astype(float))
            A = (Dn @ raw @ Dn + sp.diags(absdiag * Z)).tocsr()
            A.eliminate_zeros()
            # interface: rows on the refinement edge, columns not, neither on the boundary (src/step-50.cc:896-920)
            E = sp.diags((edge & ~bd).astype(float))
            C = sp.diags((~edge & ~bd).astype(float))
            I = (E @ raw @ C).tocsr()
            I.eliminate_zeros()
            self.A.append(A)
            self.I.append(I)
            self.pattern.append(pat)
            self.A_stored.append(embed(pat, A))
        # prolongation level l -> l+1 (rows: level l+1 dofs, cols: level l dofs)
        VO = forest.VO
        nv = 1 << dim
        for l in range(forest.n_levels - 1):
            par = np.nonzero(forest.child0[l] >= 0)[0]
            pd = dofs.level_cell_dofs[l][par]  # (np, 8)
            rows, cols, vals = [], [], []
            for c in range(nv):
                chd = dofs.level_cell_dofs[l + 1][forest.child0[l][par] + c]  # (np, 8)
                for v in range(nv):
                    p = VO[c] + VO[v]  # position in the parent, half units
                    for w in range(nv):
                        wt = np.prod([(p[d] / 2.0) if VO[w, d] else (1.0 - p[d] / 2.0) for d in range(dim)])
                        if wt != 0.0:
                            rows.append(chd[:, v])
                            cols.append(pd[:, w])
                            vals.append(np.full(len(par), wt))
            rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
            pair = rows * dofs.level_n[l] + cols
            _, keep = np.unique(pair, return_index=True)  # `set`, not add
            rows, cols, vals = rows[keep], cols[keep], vals[keep]
            vals = np.where(dofs.level_boundary[l][cols], 0.0, vals)  # boundary columns zeroed
            P = sp.coo_matrix((vals, (rows, cols)), shape=(dofs.level_n[l + 1], dofs.level_n[l])).tocsr()
            P.eliminate_zeros()
            self.P.append(P)
