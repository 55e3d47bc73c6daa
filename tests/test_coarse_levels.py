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
