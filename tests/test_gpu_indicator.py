"""GPU parity test of the refinement indicator (src/step-50.cc:1020-1090): gmg_error_indicator
against the sequential host restatement (ministep error_indicator, itself pinned against the oracle and the reference's
golden thresholds in test_host_vs_oracle.py / test_oracle_goldens.py).  The device gathers the face integrals in the
same order with unfused arithmetic: the float32 indicators must be BIT-identical, so the marked cells are too."""
import numpy as np
import pytest

from conftest import make_prm
from helpers import oracle_cycle, pkg
import hostlib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refined():
    """The 2-atom golden case advanced three adaptive cycles (4 levels, hanging faces in all three directions)."""
    P = oracle_cycle(make_prm(cycles=4, bc="Exact", atom="atom_n1_2.data", nq=4), 3)
    f = P.forest
    M = hostlib.Ministep(f.reps, f.lo, f.hi)
    for flags in P.flag_history[:3]:
        M.refine(flags)
    M.build(matrices=False)
    return P, M


@pytest.mark.parametrize("residual", [True, False])
def test_indicator_bit_identical_to_sequential_restatement(refined, residual):
    P, M = refined
    capi = pkg().capi
    f = P.forest
    nq = 2
    gp, gw = hostlib.gauss(nq)
    cell_h = np.concatenate([np.full(len(M.get("active_cells", l)), f.h(l)) for l in range(M.n_levels)])
    cell_dofs = np.concatenate([M.get("cell_dofs", l).reshape(-1, 8) for l in range(M.n_levels)])
    nc, n_dofs = len(cell_h), len(M.get("boundary"))
    weights = np.array([gw[x] * gw[y] * gw[z] for z in range(nq) for y in range(nq) for x in range(nq)])
    shape = np.zeros((nq ** 3, 8))
    for q, (z, y, x) in enumerate([(z, y, x) for z in range(nq) for y in range(nq) for x in range(nq)]):
        p = (gp[x], gp[y], gp[z])
        for v in range(8):
            shape[q, v] = np.prod([p[k] if (v >> k) & 1 else 1.0 - p[k] for k in range(3)])
    rng = np.random.default_rng(11)
    rho = rng.standard_normal((nc, nq ** 3))
    u = rng.standard_normal(n_dofs)
    g = capi.Gmg()
    g.assemble_rhs(rho, cell_h, cell_dofs, shape, weights, n_dofs, M.get("hang_rowptr"), M.get("hang_col"), M.get("hang_val"),
                   M.get("constrained"))
    gp2, gw2 = hostlib.gauss(2)
    eta, mx = g.error_indicator(M.get("topo_face_nb"), M.get("topo_face_kind"), M.get("topo_hang_children"), u, rho, residual,
                                gp2, gw2)
    marked, thr_dev = g.mark_cells(nc)
    g.close()
    thr = M.error_indicator(u, rho, nq, residual)
    ref = np.concatenate([M.get("eta", l) for l in range(M.n_levels)])
    # marking on the device (gmg_mark_cells) = the host's mark_cells: same threshold bits, same flags
    ref_flags = np.concatenate([M.get("flags", l)[M.get("active_cells", l)] for l in range(M.n_levels)])
    assert thr_dev == thr
    assert np.array_equal(marked.astype(bool), ref_flags.astype(bool)) and marked.any() and not marked.all()
    assert ref.dtype == np.float32 and eta.dtype == np.float32
    assert (M.get("topo_face_kind") & 3 == 2).any() and (M.get("topo_face_kind") & 3 == 1).any()  # hanging faces are covered
    assert np.array_equal(eta.view(np.uint32), ref.view(np.uint32))
    # same indicators, same maximum: the threshold 0.6 * max (src/step-50.cc:1084) marks the same cells
    assert mx == ref.max() and abs(thr - 0.6 * float(mx)) <= 1e-12 * thr


def test_solution_transfer_bit_identical_to_sequential_restatement():
    """gmg_transfer_solution (SolutionTransfer::interpolate + set_zero, src/step-50.cc:1110-1119) against ministep's
    transfer_solution across every refinement step of the 2-atom golden hierarchy: the same doubles, bit for bit
    (shared edge / face points take the value of the first refined cell, as in the sequential loop)."""
    P = oracle_cycle(make_prm(cycles=4, bc="Exact", atom="atom_n1_2.data", nq=4), 3)
    f = P.forest
    capi = pkg().capi
    g = capi.Gmg()
    old = hostlib.Ministep(f.reps, f.lo, f.hi)
    old.build(matrices=False)
    rng = np.random.default_rng(5)
    history = []
    for step, flags in enumerate(P.flag_history[:3]):
        new = hostlib.Ministep(f.reps, f.lo, f.hi)
        for fl in history + [flags]:
            new.refine(fl)
        new.build(matrices=False)
        history.append(flags)
        u_old = rng.standard_normal(len(old.get("boundary")))
        old_res = old.n_levels - 1
        ref = new.transfer_from(old, old_res, u_old)
        co, cn, pp, pd = new.transfer_tables(old, old_res)
        assert len(pd) > 0
        x = g.transfer_solution(u_old, len(ref), co, cn, pp, pd, new.get("constrained"))
        assert np.array_equal(x.view(np.uint64), ref.view(np.uint64)), step
        assert np.abs(x).max() > 0
        old = new
    g.close()


def test_refinement_entry_points_reject_bad_hand_overs():
    """Error behaviour of the two new entry points: an index table that points outside the new vector, or marking
    without indicators on the device, is GMG_EINVAL with a message -- never a wrong answer or an out-of-bounds access."""
    capi = pkg().capi
    g = capi.Gmg()
    with pytest.raises(capi.GmgError):
        g.mark_cells(10)
    u_old = np.arange(4, dtype=float)
    pass_ptr = np.array([0, 1], dtype=np.int64)
    pd = -np.ones((1, 27), dtype=np.int32)
    pd[0, 0] = 99  # outside [-1, n_new)
    with pytest.raises(capi.GmgError):
        g.transfer_solution(u_old, 4, [0, 1, 2, 3], [0, 1, 2, 3], pass_ptr, pd, np.zeros(4, dtype=np.uint8))
    with pytest.raises(capi.GmgError):
        g.transfer_solution(u_old, 4, [0, 1, 2, 7], [0, 1, 2, 3], np.array([0], dtype=np.int64), np.zeros((0, 27), dtype=np.int32),
                            np.zeros(4, dtype=np.uint8))
    # a dof nobody provides a value for
    with pytest.raises(capi.GmgError):
        g.transfer_solution(u_old, 5, [0, 1, 2, 3], [0, 1, 2, 3], np.array([0], dtype=np.int64), np.zeros((0, 27), dtype=np.int32),
                            np.zeros(5, dtype=np.uint8))
    x = g.transfer_solution(u_old, 4, [0, 1, 2, 3], [3, 2, 1, 0], np.array([0], dtype=np.int64), np.zeros((0, 27), dtype=np.int32),
                            np.array([0, 0, 0, 1], dtype=np.uint8))
    assert np.array_equal(x, [3.0, 2.0, 1.0, 0.0])  # (constrained dof 3 zeroed: it held u_old[0] = 0 anyway)
    g.close()
