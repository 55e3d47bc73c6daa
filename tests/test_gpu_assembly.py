"""Device-side matrix assembly (gmg_assemble_matrix) on the GPU: the CSR built by the kernels is the
CSR of the host assembly -- row pointer, columns and value bits -- and LaplaceProblem prints the same numbers with
`Matrix assembly = Device` as with the matrices assembled on the host."""
import numpy as np
import pytest

from conftest import make_prm
from helpers import pkg
import hostlib
from test_assembly_emulation import random_forest

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("seed,reps,cycles,frac", [(1, 4, 3, 0.15), (2, 6, 2, 0.3), (3, 3, 4, 0.1), (5, 5, 0, 0.0),
                                                   (6, 12, 2, 0.1)])
def test_device_assembly_reproduces_host_csr_bit_for_bit(seed, reps, cycles, frac):
    capi = pkg().capi
    M = random_forest(seed, reps, cycles, frac)
    K = hostlib.unit_stiffness()
    gmg = capi.Gmg()
    gmg.set_num_levels(M.n_levels)
    # system matrix: hanging nodes and Dirichlet rows condensed
    n, dofs, h, uh, flags, hang = hostlib.assembly_inputs(M, 0)
    gmg.assemble_matrix(capi.GMG_SYSTEM, 0, n, dofs, h, flags, K, hang=hang, uniform_h=uh)
    rp, col, val = gmg.raw_matrix(capi.GMG_SYSTEM, 0, n)
    ref = M.csr("sys")
    assert np.array_equal(rp, ref[0]) and np.array_equal(col, ref[1]) and same_bits(val, ref[2])
    # level 0: boundary dofs eliminated
    n, dofs, h, uh, flags, hang = hostlib.assembly_inputs(M, 1, 0)
    gmg.assemble_matrix(capi.GMG_LEVEL, 0, n, dofs, h, flags, K, hang=hang, uniform_h=uh)
    rp, col, val = gmg.raw_matrix(capi.GMG_LEVEL, 0, n)
    ref = M.csr("A", 0)
    assert np.array_equal(rp, ref[0]) and np.array_equal(col, ref[1]) and same_bits(val, ref[2])
    gmg.close()


def test_rows_wider_than_the_small_kernel_and_argument_checks():
    """A synthetic cell complex whose first dof sits in 12 cells (85 columns): handled by the wide-row kernel; the
    reference is the cell loop in numpy (np.add.at adds duplicates in index order = cell order)."""
    capi = pkg().capi
    K = hostlib.unit_stiffness()
    n_cells, h = 12, 0.5
    dofs = np.zeros((n_cells, 8), dtype=np.int32)
    for c in range(n_cells):
        dofs[c, 1:] = 1 + 7 * c + np.arange(7)
    dofs[:, [0, 3]] = dofs[:, [3, 0]]  # dof 0 is local vertex 3 of every cell
    n = 1 + 7 * n_cells
    flags = np.zeros(n, dtype=np.uint8)
    flags[[5, 20]] = 1
    dense = np.zeros((n, n))
    mask = np.zeros((n, n), dtype=bool)
    rows = np.repeat(dofs, 8, axis=1).ravel()
    cols = np.tile(dofs, (1, 8)).ravel()
    vals = np.tile((h * K).ravel(), n_cells)
    mask[rows, cols] = True
    keep = (flags[rows] == 0) & (flags[cols] == 0)
    np.add.at(dense, (rows[keep], cols[keep]), vals[keep])
    diag = (rows == cols) & (flags[rows] == 1)
    np.add.at(dense, (rows[diag], cols[diag]), np.abs(vals[diag]))
    gmg = capi.Gmg()
    gmg.set_num_levels(1)
    gmg.assemble_matrix(capi.GMG_SYSTEM, 0, n, dofs, None, flags, K, uniform_h=h)
    rp, col, val = gmg.raw_matrix(capi.GMG_SYSTEM, 0, n)
    assert rp[1] - rp[0] == 85 and rp[-1] == mask.sum()
    for i in range(n):
        assert np.array_equal(col[rp[i]:rp[i + 1]], np.nonzero(mask[i])[0])
        assert same_bits(val[rp[i]:rp[i + 1]], dense[i, mask[i]])
    bad = dofs.copy()
    bad[3, 2] = n
    with pytest.raises(capi.GmgError):
        gmg.assemble_matrix(capi.GMG_SYSTEM, 0, n, bad, None, flags, K, uniform_h=h)
    with pytest.raises(capi.GmgError):
        gmg.assemble_matrix(capi.GMG_LEVEL, 1, n, dofs, None, flags, K, uniform_h=h)
    gmg.close()


def test_laplace_problem_prints_the_same_numbers_with_device_assembly():
    """Three adaptive cycles of the 2-atom golden case (hanging nodes from cycle 1 on) with `Matrix assembly = Device`
    against the host-assembled run: same meshes, iteration counts (outer and coarse) and matrix norms; vector norms to
    1e-10 (the load vector is accumulated with atomics: its last bits vary from run to run in either mode)."""
    base = dict(cycles=3, bc="Exact", atom="atom_n1_2.data", nq=4)
    _, recs_h = hostlib.run_problem(make_prm(**base))
    extra = "subsection Solver input data\n set Matrix assembly = Device\nend\n"
    _, recs_d = hostlib.run_problem(make_prm(**base, extra=extra))
    assert len(recs_d) == len(recs_h) == 3
    for d, h in zip(recs_d, recs_h):
        for k in ("n_active_cells", "n_dofs", "n_dofs_level", "its", "n_flagged", "coarse_its"):
            assert d[k] == h[k], k
        for k in ("mat_l1", "mat_linf", "mat_frob"):
            assert abs(d[k] - h[k]) <= 1e-14 * abs(h[k]), k
        for k in ("rhs_l1", "rhs_l2", "rhs_linf", "start", "sol_l1", "sol_l2", "sol_linf", "threshold"):
            assert abs(d[k] - h[k]) <= 1e-10 * abs(h[k]), k
        assert abs(d["conv"] - h["conv"]) <= 1e-3 * abs(h["conv"])
