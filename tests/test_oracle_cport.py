"""The plain-C oracle (oracle/csrc/oracle.c: CPU baseline) against the numpy oracle that is pinned to the goldens."""
import numpy as np
import pytest

from conftest import make_prm
from helpers import oracle_cycle, rel_l2
from oracle import cport, fe, solver


@pytest.fixture(scope="module")
def P3():
    return oracle_cycle(make_prm(cycles=4, bc="Exact", atom="atom_n1_2.data", nq=4), 3)


def hierarchy(P):
    return cport.Hierarchy(P.system.A, P.ops.A, [None] + P.ops.I[1:], P.ops.P, P.dofs.copy_global, P.dofs.copy_level)


@pytest.mark.parametrize("sm,gold_its", [("ssor", 6), ("jacobi", None)])
def test_c_pcg_matches_numpy_oracle(P3, sm, gold_its):
    H = hierarchy(P3)
    r = H.pcg(P3.b, P3.x0, smoother=sm)
    mg = solver.Multigrid(P3.ops, P3.dofs, sm, 0.5, 2)
    xo, ito, r0o, ro = solver.pcg_gmg(P3.system.A, P3.b, P3.x0, mg)
    assert r["rc"] == 0 and r["its"] == ito and (gold_its is None or ito == gold_its)
    assert r["coarse_its"] == list(mg.coarse_its)
    assert abs(r["res0"] - r0o) < 1e-12 and rel_l2(r["x"], xo) < 1e-10


def test_block_ssor_emulates_rank_dependence(P3):
    """Processor-block SSOR: results depend on the number of blocks (as the 1/3/7-rank goldens do) but stay within +-2."""
    H = hierarchy(P3)
    its = [H.pcg(P3.b, P3.x0, smoother="ssor", n_blocks=nb)["its"] for nb in (1, 3, 7, 20)]
    assert its[0] == 6 and max(its) - min(its) <= 2


def test_c_cg_spmv_and_rhs_path(P3):
    P = P3
    rng = np.random.default_rng(0)
    x = rng.standard_normal(P.dofs.n)
    assert rel_l2(cport.spmv(P.system.A, x), P.system.A @ x) < 1e-14
    b0 = rng.standard_normal(P.dofs.level_n[0]) * ~P.dofs.level_boundary[0]
    xc, it, res, rc = cport.cg(P.ops.A[0], b0)
    xo, ito, _, _ = solver.cg(P.ops.A[0], b0)
    assert rc == 0 and it == ito and rel_l2(xc, xo) < 1e-9
    f = P.forest
    ptr, idx = cport.bin_atoms(f.reps, f.lo, f.H, P.pos, P.cutoff * P.r_c)
    assert np.array_equal(ptr, P.lists0[0]) and np.array_equal(idx, P.lists0[1])
    from test_gpu_rhs import active_cell_arrays
    lo, h, dofs, base = active_cell_arrays(P)
    pts, wts = fe.tensor_rule(P.nq_rhs, 3)
    rho = cport.density(lo, h, base, P.lists0[0], P.lists0[1], P.pos, P.charges, pts, P.r_c)
    assert rel_l2(rho, np.concatenate(P.dens)) < 1e-13
    d = P.dofs
    hp = np.zeros(d.n + 1, dtype=np.int64)
    np.add.at(hp, d.hang_rows + 1, 1)
    b = cport.load_vector(rho, h, dofs, fe.shape_values(pts, 3), wts, d.n, np.cumsum(hp), d.hang_cols, d.hang_vals,
                          d.constrained, kref=fe.stiffness(1.0, 3), ghat=P.system.T @ P.g)
    assert rel_l2(b, P.b) < 1e-13
