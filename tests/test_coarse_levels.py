"""`Coarse levels below the base mesh = k`, CPU part: the base lattice built as k global refinements of a
lattice 2^k times coarser IS the reference's base lattice (same cells, same edge length bits, same matrices up to the dof
numbering); the extra levels below it are what the device's V-cycle then recurses into."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import make_prm
import hostlib


def by_coordinates(M, level=None):
    xyz = (M.get("dof_xyz") if level is None else M.get("level_xyz", level)).reshape(-1, 3)
    return np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0])), xyz


def permuted(M, prefix, perm, level=0):
    rp, col, val = M.csr(prefix, level)
    n = len(rp) - 1
    A = sp.csr_matrix((val, col, rp), shape=(n, n))
    # keep the explicit zeros: compare patterns through a matrix of ones
    S = sp.csr_matrix((np.ones_like(val), col, rp), shape=(n, n))
    return A[perm][:, perm].toarray(), S[perm][:, perm].toarray()


@pytest.mark.parametrize("reps,k", [(8, 1), (12, 2), (8, 3)])
def test_coarsened_base_lattice_is_the_reference_base_lattice(reps, k):
    lo, hi = -1.5, 2.5
    std = hostlib.Ministep(reps, lo, hi)
    std.build()
    crs = hostlib.Ministep(reps >> k, lo, hi)
    crs.refine_global(k)
    crs.build()
    assert crs.n_levels == k + 1 and crs.n_cells(k) == reps ** 3
    # same edge length, bit for bit
    assert crs.h0 / (1 << k) == std.h0
    # same cells
    a = {tuple(c) for c in std.get("ijk", 0).reshape(-1, 3)}
    b = {tuple(c) for c in crs.get("ijk", k).reshape(-1, 3)}
    assert a == b
    # same system matrix and same top-level matrix up to the numbering (pattern incl. explicit zeros; values to rounding:
    # the cell order, hence the summation order, differs)
    p_std, xyz_std = by_coordinates(std)
    p_crs, xyz_crs = by_coordinates(crs)
    assert np.array_equal(xyz_std[p_std], xyz_crs[p_crs])
    A1, S1 = permuted(std, "sys", p_std)
    A2, S2 = permuted(crs, "sys", p_crs)
    assert np.array_equal(S1, S2) and np.allclose(A1, A2, rtol=0, atol=1e-15)
    q_std, _ = by_coordinates(std, 0)
    q_crs, _ = by_coordinates(crs, k)
    L1, T1 = permuted(std, "A", q_std, 0)
    L2, T2 = permuted(crs, "A", q_crs, k)
    assert np.array_equal(T1, T2) and np.allclose(L1, L2, rtol=0, atol=1e-15)
    # the levels below: (reps / 2^j + 1)^3 dofs, no refinement edges, prolongations between all of them
    level_n = crs.get("level_n")
    assert list(level_n) == [((reps >> (k - j)) + 1) ** 3 for j in range(k + 1)]
    for j in range(k + 1):
        assert crs.get("level_edge", j).sum() == 0
    for j in range(k):
        rp = crs.get("P_rowptr", j)
        assert len(rp) - 1 == level_n[j + 1]


def test_prm_key_is_declared_and_validated():
    ok = "subsection Geometry\n set Coarse levels below the base mesh = 2\nend\n"
    hostlib.check_prm(make_prm(extra=ok))
    with pytest.raises(hostlib.HostError):
        hostlib.check_prm(make_prm(extra=ok.replace("= 2", "= two")))


def test_oracle_and_host_agree_on_the_coarsened_hierarchy_and_the_solution_is_the_reference_one():
    """The oracle restates the option too: with 2 coarse levels below the base mesh of the 2-atom golden case it produces,
    on cycle 1 (hanging nodes present), the forest / numbering / constraints ministep produces (bit-exact), the load
    vector of the reference hierarchy (norms to rounding) and the same solution (to the CG tolerance)."""
    from helpers import oracle_cycle
    two = dict(bc="Exact", atom="atom_n1_2.data", nq=4)
    ref = oracle_cycle(make_prm(cycles=2, **two), 1, smoother="jacobi")
    extra = "subsection Geometry\n set Coarse levels below the base mesh = 2\nend\n"
    P = oracle_cycle(make_prm(cycles=2, extra=extra, **two), 1, smoother="jacobi")
    f, d = P.forest, P.dofs
    assert f.n_levels == ref.forest.n_levels + 2 and list(d.level_n[:3]) == [12 ** 3, 23 ** 3, 45 ** 3]
    assert list(d.level_n[2:]) == list(ref.dofs.level_n)
    # host substrate on the same flags
    M = hostlib.Ministep(f.reps, f.lo, f.hi)
    M.refine_global(2)
    for flags in P.flag_history[:1]:
        M.refine(flags)
    M.build()
    assert M.n_levels == f.n_levels
    for l in range(M.n_levels):
        assert np.array_equal(M.get("ijk", l).reshape(-1, 3), f.ijk[l])
        assert np.array_equal(M.get("cell_dofs", l).reshape(-1, 8), d.cell_dofs[l])
        assert np.array_equal(M.get("level_cell_dofs", l).reshape(-1, 8), d.level_cell_dofs[l])
        assert np.array_equal(M.get("copy_global", l), d.copy_global[l])
    assert np.array_equal(M.get("dof_xyz").reshape(-1, 3), d.xyz)
    assert np.array_equal(M.get("hanging").astype(bool), d.hanging)
    # same problem as the reference hierarchy
    a, b = P.records[-1], ref.records[-1]
    assert a["n_active_cells"] == b["n_active_cells"] and a["n_dofs"] == b["n_dofs"]
    for k in ("rhs_l1", "rhs_l2", "rhs_linf", "mat_l1", "mat_linf", "mat_frob"):
        assert abs(a[k] - b[k]) <= 1e-12 * abs(b[k]), k
    for k in ("sol_l1", "sol_l2", "sol_linf"):
        assert abs(a[k] - b[k]) <= 2e-7 * abs(b[k]), k
    assert abs(a["its"] - b["its"]) <= 2 and max(a["coarse_its"]) < max(b["coarse_its"])
    # the same cells carry the same atom lists (sets of atoms per active cell, compared through cell centres)
    def lists_by_centre(Q):
        out = {}
        for l in range(Q.forest.n_levels):
            ptr, idx = Q.lists[l]
            for c in Q.dofs.active_cells[l]:
                key = tuple(np.round((Q.forest.lo + (Q.forest.ijk[l][c] + 0.5) * Q.forest.h(l)) * 64).astype(int))
                out[key] = tuple(idx[ptr[c]:ptr[c + 1]])
        return out
    assert lists_by_centre(P) == lists_by_centre(ref)
