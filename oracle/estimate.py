"""Oracle error indicator, marking, solution transfer, energies.

Restates src/step-50.cc:1020-1090 (`estimate_error_and_mark_cells`: KellyErrorEstimator with
Strategy::cell_diameter, face rule QGauss<dim-1>(2), plus h_K^2 * int (lap u_h + 4 pi rho)^2; the
Q1 Laplacian on cubes is 0), :1095-1121 (`refine_grid`: SolutionTransfer interpolation then
`constraints.set_zero`), :1310-1420 (`postprocess_electrostatic_energy`) and :1423-1461
(`postprocess_error_in_energy_norm`).  `Vector<float>` storage is mimicked.
TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import math

import numpy as np
from scipy.special import erf, erfc

from . import fe


def _normal_derivative(U, h, a, side_pts, VO):
    """U: (n, 2^dim) nodal values; derivative along axis a at tangential unit coords side_pts (n, nq, dim-1)
    (constant along a for Q1).  Returns (n, nq)."""
    dim = VO.shape[1]
    others = [d for d in range(dim) if d != a]
    out = np.zeros(side_pts.shape[:2])
    for v in range(1 << dim):
        w = np.full(side_pts.shape[:2], (1.0 if VO[v, a] else -1.0) / h)
        for k, d in enumerate(others):
            t = side_pts[:, :, k]
            w = w * (t if VO[v, d] else (1.0 - t))
        out += U[:, v][:, None] * w
    return out


def kelly_plus_residual(forest, dofs, u, dens, nq_rhs, residual_term=True):
    """Per-level float32 indicators eta_K on active cells (u: solution after constraints.distribute)."""
    dim, VO = forest.dim, forest.VO
    fpts, fw = fe.tensor_rule(2, dim - 1)  # QGauss<dim-1>(degree+1)
    _, wq = fe.tensor_rule(nq_rhs, dim)
    nl = forest.n_levels
    # face integrals per active cell and face, double
    FI = [np.zeros((len(dofs.active_cells[l]), 2 * dim)) for l in range(nl)]
    act_pos = []  # map cell index -> position in active list
    for l in range(nl):
        pos = np.full(forest.n_cells(l), -1, dtype=np.int64)
        pos[dofs.active_cells[l]] = np.arange(len(dofs.active_cells[l]))
        act_pos.append(pos)
    for l in range(nl):
        act = dofs.active_cells[l]
        if len(act) == 0:
            continue
        h = forest.h(l)
        U = u[dofs.cell_dofs[l]]
        n = forest.cells_per_axis(l)
        for a in range(dim):
            others = [d for d in range(dim) if d != a]
            for side in (0, 1):
                face = 2 * a + side
                nb = forest.ijk[l][act].copy()
                nb[:, a] += 1 if side else -1
                inside = (nb[:, a] >= 0) & (nb[:, a] < n)
                nidx = forest.lookup(l, nb)
                jxw = fw * h ** (dim - 1)
                pts = np.broadcast_to(fpts[None, :, :], (len(act),) + fpts.shape)
                own = _normal_derivative(U, h, a, pts, VO)
                # (A) same-level active neighbour
                same = nidx >= 0
                same[same] = forest.child0[l][nidx[same]] < 0
                if same.any():
                    Un = u[dofs.cell_dofs[l][act_pos[l][nidx[same]]]]
                    oth = _normal_derivative(Un, h, a, pts[same], VO)
                    FI[l][same, face] = (((own[same] - oth) ** 2) * jxw).sum(1)
                # (B) coarser neighbour: this face is a subface of the coarse cell's face
                coarse = inside & (nidx < 0)
                if coarse.any():
                    assert l > 0
                    cidx = forest.lookup(l - 1, nb[coarse] >> 1)
                    assert (cidx >= 0).all() and (forest.child0[l - 1][cidx] < 0).all()
                    cpos = act_pos[l - 1][cidx]
                    Uc = u[dofs.cell_dofs[l - 1][cpos]]
                    sub = (forest.ijk[l][act[coarse]][:, others] - 2 * forest.ijk[l - 1][cidx][:, others])
                    cpts = (sub[:, None, :] + pts[coarse]) / 2.0
                    oth = _normal_derivative(Uc, 2 * h, a, cpts, VO)
                    I = (((own[coarse] - oth) ** 2) * jxw).sum(1)
                    FI[l][coarse, face] = I
                    # coarse side: sum over its subfaces, in subface (= child index) order
                    np.add.at(FI[l - 1], (cpos, np.full(len(cpos), 2 * a + (1 - side))), I)
    eta = []
    for l in range(nl):
        h = forest.h(l)
        diam = math.sqrt(dim * h * h)
        err = np.zeros(len(dofs.active_cells[l]), dtype=np.float32)
        for face in range(2 * dim):  # Vector<float> accumulation, face by face
            err = (err.astype(np.float64) + FI[l][:, face] * diam).astype(np.float32)
        kelly = np.sqrt(err.astype(np.float64)).astype(np.float32)
        jxw = wq * h ** dim
        res = (((4.0 * math.pi * dens[l]) ** 2) * jxw).sum(1) if len(err) else np.zeros(0)
        if not residual_term:  # the build that produced the cluster logs marked with the Kelly part only
            res = res * 0.0
        eta.append(np.sqrt(kelly.astype(np.float64) ** 2 + diam ** 2 * res).astype(np.float32))
    return eta


def mark(forest, dofs, eta, fraction=0.6):
    """threshold = 0.6 * max eta (double); refine where eta >= threshold (GridRefinement::refine)."""
    mx = max((float(e.max()) for e in eta if len(e)), default=0.0)
    threshold = fraction * mx
    flags = []
    for l in range(forest.n_levels):
        fl = np.zeros(forest.n_cells(l), dtype=bool)
        if mx > 0.0:
            fl[dofs.active_cells[l]] = eta[l].astype(np.float64) >= threshold
        flags.append(fl)
    return threshold, flags


def transfer_solution(old_forest_res, old_dofs, u_old, forest, dofs):
    """SolutionTransfer::interpolate onto the refined mesh followed by constraints.set_zero:
    persisting vertices keep their value, new vertices get the multilinear interpolant of the
    (already refined) parent; constrained entries are zeroed (src/step-50.cc:1118-1119)."""
    dim, VO = forest.dim, forest.VO
    shift = forest.resolution() - old_forest_res
    old_xyz = old_dofs.xyz << shift
    x = np.full(dofs.n, np.nan)
    new_of_old = dofs.lookup_dof(forest.vertex_key(old_xyz))
    ok = new_of_old >= 0
    x[new_of_old[ok]] = u_old[ok]
    # children of cells that were active in the old mesh: interpolate from the parent's vertices
    for l in range(forest.n_levels - 1):
        par = np.nonzero(forest.child0[l] >= 0)[0]
        if len(par) == 0:
            continue
        pv = forest.vertex_coords(l, par)  # (np, 8, dim) at new resolution
        pd = dofs.lookup_dof(forest.vertex_key(pv))  # all corners are dofs of the children
        Up = x[pd]
        half = 1 << (forest.resolution() - l - 1)
        import itertools
        for t in itertools.product((0, 1, 2), repeat=dim):
            if all(c != 1 for c in t):
                continue
            p = pv[:, 0, :] + np.array(t) * half
            d = dofs.lookup_dof(forest.vertex_key(p))
            val = np.zeros(len(par))
            for v in range(1 << dim):
                w = np.prod([(t[k] / 2.0) if VO[v, k] else (1.0 - t[k] / 2.0) for k in range(dim)])
                if w != 0.0:
                    val += w * Up[:, v]
            new = np.isnan(x[d]) & ~np.isnan(val)
            x[d[new]] = val[new]
    assert not np.isnan(x).any()
    x[dofs.constrained] = 0.0
    return x


def locate(forest, dofs, X):
    """Active cell (level, index) around points X and their unit-cell coordinates."""
    n = len(X)
    fpos = (X - forest.lo) / forest.H
    ijk = np.clip(np.floor(fpos).astype(np.int64), 0, forest.reps - 1)
    lev = np.zeros(n, dtype=np.int64)
    idx = forest.lookup(0, ijk)
    for l in range(forest.n_levels - 1):
        m = lev == l
        if m.any():
            m[m] = forest.child0[l][idx[m]] >= 0
        if not m.any():
            continue
        h = forest.h(l + 1)
        cijk = np.clip(np.floor((X[m] - forest.lo) / h).astype(np.int64), 2 * forest.ijk[l][idx[m]],
                       2 * forest.ijk[l][idx[m]] + 1)
        idx[m] = forest.lookup(l + 1, cijk)
        lev[m] = l + 1
    xi = np.zeros_like(X)
    for l in range(forest.n_levels):
        m = lev == l
        if m.any():
            h = forest.h(l)
            xi[m] = (X[m] - (forest.lo + forest.ijk[l][idx[m]] * h)) / h
    return lev, idx, xi


def point_values(forest, dofs, u, X):
    lev, idx, xi = locate(forest, dofs, X)
    out = np.zeros(len(X))
    for l in range(forest.n_levels):
        m = lev == l
        if not m.any():
            continue
        pos = np.full(forest.n_cells(l), -1, dtype=np.int64)
        pos[dofs.active_cells[l]] = np.arange(len(dofs.active_cells[l]))
        N = fe.shape_values(xi[m], forest.dim)
        out[m] = (N * u[dofs.cell_dofs[l][pos[idx[m]]]]).sum(1)
    return out


def electrostatic_energy(forest, dofs, u, pos, q, r_c):
    """The five printed energies (src/step-50.cc:1315-1418); u after constraints.distribute."""
    n = len(q)
    analytic = short = 0.0
    for i in range(n):
        r = np.sqrt(((pos[i + 1:] - pos[i]) ** 2).sum(1))
        analytic += float((q[i] * q[i + 1:] / r).sum())
        short += float((q[i] * q[i + 1:] * erfc(r / r_c) / r).sum())
    fe_part = float((0.5 * q * point_values(forest, dofs, u, pos)).sum())
    self_e = float((q * q / (math.sqrt(math.pi) * r_c)).sum())
    return dict(analytic=analytic, short=short, fe=fe_part, self=self_e, total=short + fe_part - self_e)


def energy_norm_error(forest, dofs, u, pos, q, r_c):
    """sqrt(int |grad u_h - grad u_exact|^2), QGauss(2) (src/step-50.cc:1424-1461)."""
    dim = forest.dim
    pts, wts = fe.tensor_rule(2, dim)
    G = fe.shape_grads(pts, dim)
    inv_c = 1.0 / (math.sqrt(math.pi) * r_c)
    tot = 0.0
    for l in range(forest.n_levels):
        act = dofs.active_cells[l]
        if len(act) == 0:
            continue
        h = forest.h(l)
        org = forest.lo + forest.ijk[l][act] * h
        U = u[dofs.cell_dofs[l]]
        gh = np.einsum("cv,qvd->cqd", U, G) / h
        xq = org[:, None, :] + h * pts[None, :, :]
        ga = np.zeros_like(gh)
        for k in range(len(q)):
            dvec = xq - pos[k]
            r = np.sqrt((dvec ** 2).sum(-1))
            f = q[k] * ((2.0 * r * np.exp(-(r / r_c) ** 2) * inv_c - erf(r / r_c)) / r ** 2)
            ga += f[..., None] * dvec / r[..., None]
        tot += float((((gh - ga) ** 2).sum(-1) * (wts * h ** dim)).sum())
    return math.sqrt(tot)
