"""CPU tests: the C++ host substrate (ministep: mesh, DoF numbering, constraints, sparsity, assembled
operators, error indicator, solution transfer, .prm grammar) against the oracle.  North star: DoF
numbering, sparsity and maps bit-exact; values to rounding."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_prm
from helpers import oracle_cycle
import hostlib

TWO = dict(bc="Exact", atom="atom_n1_2.data", nq=4)


@pytest.fixture(scope="module")
def pair():
    """Oracle advanced to cycle 3 of the 2-atom golden case and a ministep forest refined with the same flags."""
    P = oracle_cycle(make_prm(cycles=4, **TWO), 3)
    f = P.forest
    M = hostlib.Ministep(f.reps, f.lo, f.hi)
    prev = None
    for flags in P.flag_history[:3]:
        prev = M
        M.refine(flags)
    M.build()
    return P, M


def test_forest_cells_and_child_order(pair):
    P, M = pair
    assert M.n_levels == P.forest.n_levels
    for l in range(M.n_levels):
        assert np.array_equal(M.get("ijk", l).reshape(-1, 3), P.forest.ijk[l])
        assert np.array_equal(M.get("parent", l), P.forest.parent[l])
        assert np.array_equal(M.get("child0", l), P.forest.child0[l])


def test_dof_numbering_constraints_and_maps_bit_exact(pair):
    P, M = pair
    d = P.dofs
    assert np.array_equal(M.get("dof_xyz").reshape(-1, 3), d.xyz)
    for name in ("boundary", "hanging", "dirichlet", "constrained"):
        assert np.array_equal(M.get(name).astype(bool), getattr(d, name)), name
    hp = np.zeros(d.n + 1, dtype=np.int64)
    np.add.at(hp, d.hang_rows + 1, 1)
    assert np.array_equal(M.get("hang_rowptr"), np.cumsum(hp))
    assert np.array_equal(M.get("hang_col"), d.hang_cols)
    assert np.array_equal(M.get("hang_val"), d.hang_vals)
    assert np.array_equal(M.get("level_n"), d.level_n)
    for l in range(M.n_levels):
        assert np.array_equal(M.get("active_cells", l), d.active_cells[l])
        assert np.array_equal(M.get("cell_dofs", l).reshape(-1, 8), d.cell_dofs[l])
        assert np.array_equal(M.get("level_cell_dofs", l).reshape(-1, 8), d.level_cell_dofs[l])
        assert np.array_equal(M.get("level_edge", l).astype(bool), d.level_edge[l])
        assert np.array_equal(M.get("level_boundary", l).astype(bool), d.level_boundary[l])
        assert np.array_equal(M.get("copy_global", l), d.copy_global[l])
        assert np.array_equal(M.get("copy_level", l), d.copy_level[l])


def _same_csr(host, ref, tol=1e-13):
    rp, col, val = host
    ref = ref.tocsr()
    ref.sort_indices()
    assert np.array_equal(rp, ref.indptr), "row pointers differ"
    assert np.array_equal(col, ref.indices), "sparsity differs"
    scale = np.abs(ref.data).max()
    assert np.abs(val - ref.data).max() <= tol * scale


def test_system_matrix_sparsity_bit_exact_values_to_rounding(pair):
    P, M = pair
    _same_csr(M.csr("sys"), P.system.A_stored)


def test_level_interface_and_transfer_operators(pair):
    P, M = pair
    for l in range(M.n_levels):
        _same_csr(M.csr("A", l), P.ops.A_stored[l])
        rp, col, val = M.csr("I", l)
        n = P.dofs.level_n[l]
        I = sp.csr_matrix((val, col, rp), shape=(n, n))
        diff = (I - P.ops.I[l])
        assert (abs(diff).max() if diff.nnz else 0.0) < 1e-14
        if l + 1 < M.n_levels:
            rp, col, val = M.csr("P", l)
            Pm = sp.csr_matrix((val, col, rp), shape=P.ops.P[l].shape)
            Pm.sort_indices()
            ref = P.ops.P[l].copy()
            ref.sort_indices()
            assert np.array_equal(Pm.indptr, ref.indptr) and np.array_equal(Pm.indices, ref.indices)
            assert np.array_equal(Pm.data, ref.data)


def test_error_indicator_threshold_and_flags_bit_exact(pair):
    from oracle import estimate
    P, M = pair
    eta = estimate.kelly_plus_residual(P.forest, P.dofs, P.u, P.dens, P.nq_rhs)
    thr_o, flags_o = estimate.mark(P.forest, P.dofs, eta)
    thr = M.error_indicator(P.u, np.concatenate(P.dens), P.nq_rhs)
    assert thr == thr_o
    for l in range(M.n_levels):
        assert np.array_equal(M.get("eta", l), eta[l])
        assert np.array_equal(M.get("flags", l).astype(bool), flags_o[l])


def test_solution_transfer_and_distribute(pair):
    from oracle import estimate
    P, M = pair
    # refine once more with the oracle's flags of cycle 3 and transfer
    eta = estimate.kelly_plus_residual(P.forest, P.dofs, P.u, P.dens, P.nq_rhs)
    _, flags = estimate.mark(P.forest, P.dofs, eta)
    import copy
    f2 = copy.deepcopy(P.forest)
    f2.refine(flags)
    from oracle.dofs import DoFs
    d2 = DoFs(f2)
    x_o = estimate.transfer_solution(P.forest.resolution(), P.dofs, P.u, f2, d2)
    M2 = hostlib.Ministep(P.forest.reps, P.forest.lo, P.forest.hi)
    for fl in P.flag_history[:3]:
        M2.refine(fl)
    M2.refine(flags)
    M2.build(matrices=False)
    x_h = M2.transfer_from(M, P.forest.resolution(), P.u)
    assert np.abs(x_h - x_o).max() <= 1e-15 * np.abs(x_o).max()
    assert np.abs(M.distribute(P.g, P.x) - P.u).max() <= 1e-15 * np.abs(P.u).max()


def test_step16_uniform_hierarchy_with_coefficient():
    P = oracle_cycle(make_prm("Step16", nref=3, atom="x"), 0, smoother="jacobi")
    M = hostlib.Ministep(1, 0.0, 1.0)
    M.refine_global(3)
    M.build(step16=True)
    assert np.array_equal(M.get("level_n"), P.dofs.level_n)
    _same_csr(M.csr("sys"), P.system.A_stored)
    for l in range(4):
        _same_csr(M.csr("A", l), P.ops.A_stored[l])


def test_gauss_rules_match_numpy():
    from oracle import fe
    for n in (1, 2, 3, 5, 8):
        p, w = hostlib.gauss(n)
        po, wo = fe.gauss_unit(n)
        assert np.abs(p - po).max() < 1e-15 and np.abs(w - wo).max() < 1e-15


def test_prm_grammar_defaults_and_errors():
    echo = hostlib.check_prm(make_prm(cycles=5, atom="a.data"))
    assert echo.split()[:5] == ["0", "1", "0.25", "10", "5"]
    assert echo.split()[5:] == ["1", "GMG", "SSOR"]
    assert hostlib.check_prm("").split() == ["-1", "1", "0.25", "1", "2", "0", "GMG", "SSOR"]  # declared defaults
    with pytest.raises(hostlib.HostError):
        hostlib.check_prm("subsection Problem Selection\n set Problem = two charges\nend\n")  # prmtest.prm:27 is invalid
    with pytest.raises(hostlib.HostError):
        hostlib.check_prm("set No such key = 1\n")
    with pytest.raises(hostlib.HostError):
        hostlib.check_prm("subsection Geometry\n set Mesh size = 0.5\n")


def test_indicator_topology_reproduces_the_sequential_indicator(pair):
    """The face topology handed to gmg_error_indicator (ministep indicator_topology: same-level neighbour / fine side /
    coarse side with its four fine neighbours) carries everything the indicator needs: a gather formulation over it in
    numpy reproduces the sequential face loop of error_indicator (src/step-50.cc:1020-1090)."""
    P, M = pair
    nl = M.n_levels
    nb = M.get("topo_face_nb").reshape(-1, 6)
    kind = M.get("topo_face_kind").reshape(-1, 6)
    hang = M.get("topo_hang_children").reshape(-1, 4)
    h = np.concatenate([np.full(len(M.get("active_cells", l)), P.forest.h(l)) for l in range(nl)])
    dofs = np.concatenate([M.get("cell_dofs", l).reshape(-1, 8) for l in range(nl)])
    assert len(nb) == len(h) and (kind & 3 == 1).any() and (kind & 3 == 2).any()
    assert np.array_equal(np.sort(hang, axis=1), hang)                       # children in ascending active order
    fine_side = np.argwhere(kind & 3 == 1)
    assert len(fine_side) == 4 * len(hang)                                   # every hanging subface is seen from both sides
    rng = np.random.default_rng(4)
    u = rng.standard_normal(len(M.get("boundary")))
    gp, gw = hostlib.gauss(2)

    def dn(U, hh, a, s, t):
        o0, o1 = (1 if a == 0 else 0), (1 if a == 2 else 2)
        r = 0.0
        for v in range(8):
            w = (1.0 if (v >> a) & 1 else -1.0) / hh
            w *= s if (v >> o0) & 1 else 1.0 - s
            w *= t if (v >> o1) & 1 else 1.0 - t
            r += U[v] * w
        return r

    def face(U, hh, Un, hn, a, coarse, s0, s1):
        I = 0.0
        for t1 in range(2):
            for t0 in range(2):
                s, t = gp[t0], gp[t1]
                oth = dn(Un, hn, a, (s0 + s) / 2, (s1 + t) / 2) if coarse else dn(Un, hn, a, s, t)
                I += (dn(U, hh, a, s, t) - oth) ** 2 * gw[t0] * gw[t1] * hh * hh
        return I

    eta = np.zeros(len(h))
    sample = rng.choice(len(h), size=min(len(h), 1500), replace=False)  # (pure-Python loops: a sample of the cells)
    sample = np.unique(np.concatenate([sample, fine_side[:200, 0], np.argwhere(kind & 3 == 2)[:100, 0]]))
    for c in sample:
        U, err = u[dofs[c]], 0.0
        for f in range(6):
            a, n_, k = f >> 1, nb[c, f], kind[c, f]
            if n_ < 0:
                continue
            if k & 3 == 0:
                I = face(U, h[c], u[dofs[n_]], h[c], a, False, 0, 0)
            elif k & 3 == 1:
                I = face(U, h[c], u[dofs[n_]], 2 * h[c], a, True, (k >> 2) & 1, (k >> 3) & 1)
            else:
                I = 0.0
                for ch in hang[n_]:
                    ck = kind[ch, f ^ 1]
                    assert ck & 3 == 1 and nb[ch, f ^ 1] == c             # the child points back at this coarse cell
                    I += face(u[dofs[ch]], h[ch], U, 2 * h[ch], a, True, (ck >> 2) & 1, (ck >> 3) & 1)
            err += I * np.sqrt(3.0) * h[c]
        eta[c] = np.sqrt(err)
    M.error_indicator(u, np.zeros(0), 2, False)
    ref = np.concatenate([M.get("eta", l) for l in range(nl)]).astype(np.float64)
    assert np.allclose(eta[sample], ref[sample], rtol=2e-6, atol=1e-12)
