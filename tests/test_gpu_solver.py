"""GPU parity tests of the solve path, through the C ABI, against the CPU oracle on the same inputs.

Tolerances (fp64 everywhere): single operator applications 1e-13 relative L2 (different summation
order / FMA only); Krylov results 1e-9 relative (north star: solution norms to 1e-9, iteration
counts +-2 for Jacobi; the level-scheduled lexicographic SSOR reproduces the reference's counts)."""
import math

import numpy as np
import pytest

from conftest import make_prm
from helpers import device_solve_cycle, hand_over, oracle_cycle, pkg, rel_l2

pytestmark = pytest.mark.gpu

TWO_ATOMS = dict(bc="Exact", atom="atom_n1_2.data", nq=4)


@pytest.fixture(scope="module")
def capi():
    return pkg().capi


@pytest.fixture(scope="module")
def P3():
    """2-atom golden case advanced to cycle 3 (levels 91125 / 1553 / 250, hanging nodes, Exact BC)."""
    return oracle_cycle(make_prm(cycles=4, **TWO_ATOMS), 3)


@pytest.fixture(scope="module")
def lattice8():
    return oracle_cycle(make_prm(atom="atom_n1_8.data"), 0)


def test_library_targets_sm100(capi):
    assert capi.load_library().gmg_compiled_arch() == 100


def test_spmv_all_operators(capi, P3):
    rng = np.random.default_rng(0)
    g = capi.Gmg()
    hand_over(P3, g)
    x = rng.standard_normal(P3.dofs.n)
    assert rel_l2(g.spmv(capi.GMG_SYSTEM, 0, x, P3.dofs.n), P3.system.A @ x) < 1e-13
    for l in range(len(P3.ops.A)):
        xl = rng.standard_normal(P3.dofs.level_n[l])
        assert rel_l2(g.spmv(capi.GMG_LEVEL, l, xl, len(xl)), P3.ops.A[l] @ xl) < 1e-13
        if l + 1 < len(P3.ops.A):
            assert rel_l2(g.spmv(capi.GMG_PROLONG, l, xl, P3.dofs.level_n[l + 1]), P3.ops.P[l] @ xl) < 1e-13
    g.close()


def test_matrix_and_vector_norms(capi, P3, goldens):
    g = capi.Gmg()
    hand_over(P3, g)
    gold = goldens["gaussian_charges_mpirun1"][0]["cycles"][3]
    l1, linf, frob = g.matrix_norms(capi.GMG_SYSTEM)
    assert abs(l1 - gold["mat_l1"]) < 1e-9 and abs(linf - gold["mat_linf"]) < 1e-9
    assert abs(frob - gold["mat_frob"]) < 2e-8
    n1, n2, ninf = g.vector_norms(P3.b)
    assert abs(n1 - gold["rhs_l1"]) < 1e-8 and abs(n2 - gold["rhs_l2"]) < 1e-10 and abs(ninf - gold["rhs_linf"]) < 1e-11
    g.close()


def test_coarse_cg_lattice8_matches_reference_log(capi, lattice8, goldens):
    """Persistent cooperative CG on level 0 (91125 rows): 97 iterations, final residual 7.192166e-11
    (cluster log SSOR_run.o876223), solution equal to the oracle's."""
    from oracle import solver
    P = lattice8
    g = capi.Gmg()
    hand_over(P, g)
    x, its, res = g.cg_solve(capi.GMG_LEVEL, 0, P.b, 1000, 1e-10)
    xo, ito, _, reso = solver.cg(P.ops.A[0], P.b, tol=1e-10, max_it=1000)
    gold = goldens["cluster_ssor_run"][0]["cycles"][0]
    assert its == ito == 97
    assert abs(res - gold["conv"]) <= 1e-5 * gold["conv"]
    assert rel_l2(x, xo) < 1e-9
    assert abs(np.linalg.norm(x) - gold["sol_l2"]) < 1e-9
    # NoConvergence when the iteration limit is hit (SolverControl semantics)
    with pytest.raises(capi.NoConvergence):
        g.cg_solve(capi.GMG_LEVEL, 0, P.b, 10, 1e-10)
    g.close()


def test_compressed_coarse_matrix_is_lossless(capi, lattice8):
    """The compressed level-0 formats -- CSELL (value dictionary + 16-bit column offsets, 4 bytes per entry) and the
    row-pattern dictionary (4 bytes per row) -- decode to the same fp64 numbers in the same order: the coarse CG
    takes the same iterations and the SpMV is bit-identical in all three formats."""
    P = lattice8
    out = {}
    rng = np.random.default_rng(5)
    xr = rng.standard_normal(P.ops.A[0].shape[0])
    for mode in (0, 1, 2):
        g = capi.Gmg()
        g.set_compression(mode)
        hand_over(P, g)
        tr = g.matrix_traffic(capi.GMG_LEVEL, 0)
        assert tr["format"] == mode
        out[mode] = g.cg_solve(capi.GMG_LEVEL, 0, P.b, 1000, 1e-10) + (tr,)
        out[mode] += (g.spmv(capi.GMG_LEVEL, 0, xr, len(xr)),)
        g.close()
    assert out[0][1] == out[1][1] == out[2][1] == 97
    # (the kernels may run different cooperative grids, i.e. group the dot-product partials differently)
    for m in (1, 2):
        assert rel_l2(out[m][0], out[0][0]) < 1e-11 and abs(out[m][2] - out[0][2]) <= 1e-6 * out[0][2]
        assert np.array_equal(out[m][4], out[0][4])  # same entries, same order, same FMA chain: same bits
    assert out[1][3]["cg_iter_bytes"] < 0.55 * out[0][3]["cg_iter_bytes"]
    assert out[2][3]["cg_iter_bytes"] < 0.35 * out[0][3]["cg_iter_bytes"]  # ~ the 88 n bytes of the vectors


def test_window_kernel_with_two_rows_per_lane_is_the_default_and_changes_nothing(capi, lattice8, goldens, monkeypatch):
    """The coarse-grid CG's window kernel walks the dominant pattern of a Q1 lattice with two consecutive rows per lane
    (XPAIR, csrc/pattern_win2.cuh: kernel ids 7 / 8) unless GMG_WIN2_VARIANT forces the one-row-per-lane loop (5 / 6): the
    rows are summed in the same entry order, so both take the cluster log's 97 iterations to the same residual and
    solution.  "global": the layout of levels too large for shared memory (h and the row codes in global memory: ids 6 / 8)."""
    P = lattice8
    gold = goldens["cluster_ssor_run"][0]["cycles"][0]
    out = {}
    for variant, layout in ((None, "smem"), ("0", "smem"), ("3", "smem"), (None, "global"), ("0", "global")):
        if variant is None:
            monkeypatch.delenv("GMG_WIN2_VARIANT", raising=False)
        else:
            monkeypatch.setenv("GMG_WIN2_VARIANT", variant)
        monkeypatch.setenv("GMG_WIN_GLOBAL_CODES", "1" if layout == "global" else "0")
        g = capi.Gmg()
        hand_over(P, g)
        kid = g.coarse_kernel(capi.GMG_LEVEL, 0)
        assert kid == (5 if variant == "0" else 7) + (1 if layout == "global" else 0), (variant, layout, kid)
        out[variant, layout] = g.cg_solve(capi.GMG_LEVEL, 0, P.b, 1000, 1e-10)
        g.close()
    ref = out["0", "smem"]
    for key, (x, its, res) in out.items():
        assert its == ref[1] == 97, key
        assert abs(res - ref[2]) <= 1e-9 * ref[2], key
        assert rel_l2(x, ref[0]) < 1e-13, key
    assert abs(out[None, "smem"][2] - gold["conv"]) <= 1e-5 * gold["conv"]


@pytest.mark.parametrize("kind", ["jacobi", "lex_ssor"])
def test_smoother_steps_match_oracle(capi, P3, kind):
    from oracle import solver
    rng = np.random.default_rng(1)
    g = capi.Gmg()
    hand_over(P3, g, kind)
    for l in range(1, len(P3.ops.A)):
        A = P3.ops.A[l]
        S = solver.jacobi(A, 0.5) if kind == "jacobi" else solver.ssor(A, 0.5)
        rhs = rng.standard_normal(A.shape[0])
        u = S(rhs)
        u = u + S(rhs - A @ u)
        assert rel_l2(g.smooth(l, rhs, np.zeros_like(rhs), True), u) < 1e-12
        u0 = rng.standard_normal(A.shape[0])
        v = u0 + S(rhs - A @ u0)
        v = v + S(rhs - A @ v)
        assert rel_l2(g.smooth(l, rhs, u0, False), v) < 1e-12
    g.close()


def _greedy_colors(A):
    A = A.tocsr()
    color = -np.ones(A.shape[0], dtype=int)
    for r in range(A.shape[0]):
        used = set()
        for k in range(A.indptr[r], A.indptr[r + 1]):
            c = A.indices[k]
            if c != r and A.data[k] != 0.0 and color[c] >= 0:
                used.add(color[c])
        pick = 0
        while pick in used:
            pick += 1
        color[r] = pick
    return color


def test_multicolour_ssor_is_colour_ordered_ssor(capi, P3):
    """The multicolour smoother equals SSOR in the colour-major ordering (same greedy colouring)."""
    rng = np.random.default_rng(2)
    g = capi.Gmg()
    hand_over(P3, g, "mc_ssor")
    l = 1
    A = P3.ops.A_stored[l].tocsr()
    color = _greedy_colors(A)
    dg = A.diagonal()
    rhs = rng.standard_normal(A.shape[0])
    u = np.zeros_like(rhs)
    order = list(range(color.max() + 1))
    for _ in range(2):
        for c in order + order[::-1]:
            rows = np.nonzero(color == c)[0]
            u[rows] += 0.5 * (rhs[rows] - A[rows] @ u) / dg[rows]
    assert rel_l2(g.smooth(l, rhs, np.zeros_like(rhs), True), u) < 1e-12
    g.close()


@pytest.mark.parametrize("kind", ["mc_ssor", "lex_ssor"])
def test_persistent_smoother_and_graphs_do_not_change_results(capi, P3, kind):
    """One cooperative launch per smoothing call (grid.sync between colours) and CUDA-graph replay of the fine-level
    sweeps are scheduling changes only: bit-identical to one launch per colour."""
    rng = np.random.default_rng(7)
    src = rng.standard_normal(P3.dofs.n)
    src[P3.dofs.constrained] = 0.0
    rhs = rng.standard_normal(P3.dofs.level_n[1])
    out = []
    for persistent, graphs in ((False, False), (True, False), (True, True)):
        g = capi.Gmg()
        g.set_persistent_smoother(persistent)
        g.set_graphs(graphs)
        hand_over(P3, g, kind)
        sm = g.smooth(1, rhs, np.zeros_like(rhs), True)
        v1 = g.vcycle(src)
        v2 = g.vcycle(src)  # second call replays the captured graphs
        assert np.array_equal(v1, v2)
        out.append((sm, v1))
        g.close()
    for sm, v in out[1:]:
        assert np.array_equal(sm, out[0][0]) and np.array_equal(v, out[0][1])


@pytest.mark.parametrize("kind", ["jacobi", "lex_ssor"])
def test_vcycle_matches_oracle(capi, P3, kind):
    from oracle import solver
    rng = np.random.default_rng(3)
    g = capi.Gmg()
    hand_over(P3, g, kind)
    src = rng.standard_normal(P3.dofs.n)
    src[P3.dofs.constrained] = 0.0
    mg = solver.Multigrid(P3.ops, P3.dofs, "jacobi" if kind == "jacobi" else "ssor", 0.5, 2)
    ref = mg.vmult(src)
    out = g.vcycle(src)
    assert rel_l2(out, ref) < 1e-9
    assert abs(g.last_coarse_iterations()[0] - mg.coarse_its[0]) <= 1
    g.close()


@pytest.mark.parametrize("cycle", [1, 2, 3])
def test_pcg_lexicographic_ssor_reproduces_reference_iterations(capi, goldens, cycle):
    """Level-scheduled lexicographic SSOR = Ifpack's sweep: the golden's iteration counts 6, 7, 6 and
    solution norms (tests/gaussian-charges.mpirun=1.output)."""
    P = oracle_cycle(make_prm(cycles=4, **TWO_ATOMS), cycle)
    gold = goldens["gaussian_charges_mpirun1"][0]["cycles"][cycle]
    r = device_solve_cycle(P, "lex_ssor")
    assert r["its"] == gold["its"] == r["oracle_its"]
    assert abs(r["res0"] - gold["start"]) < 1e-9
    assert abs(r["res"] - gold["conv"]) <= 1e-4 * gold["conv"]
    x = r["x"]
    assert abs(np.abs(x).sum() - gold["sol_l1"]) <= 1e-9 * gold["sol_l1"]
    assert abs(np.linalg.norm(x) - gold["sol_l2"]) <= 1e-9 * gold["sol_l2"]
    assert abs(np.abs(x).max() - gold["sol_linf"]) <= 1e-9 * gold["sol_linf"]
    assert r["rel_solution_error"] < 1e-9
    assert r["coarse_its"] == r["oracle_coarse_its"]


def test_pcg_jacobi_smoother_iterations_within_two(capi, P3):
    r = device_solve_cycle(P3, "jacobi")
    assert abs(r["its"] - r["oracle_its"]) <= 2 and r["its"] == r["oracle_its"]
    assert r["rel_solution_error"] < 1e-9


@pytest.mark.parametrize("kind", ["mc_ssor", "chebyshev"])
def test_pcg_other_smoothers_converge_to_the_same_solution(capi, P3, kind):
    """Multicolour SSOR / Chebyshev are different preconditioners: same solution (to the CG tolerance),
    iteration count reported, not matched."""
    r = device_solve_cycle(P3, kind)
    ref = device_solve_cycle(P3, "lex_ssor")
    tol = 1e-8 * math.sqrt(P3.b @ P3.b)
    res = np.linalg.norm(P3.system.A @ r["x"] - P3.b)
    assert res <= 1.01 * tol
    assert r["its"] <= ref["its"] + 4
    # energy-norm-small difference between two converged iterates
    assert rel_l2(r["x"], ref["x"]) < 1e-6


def test_pcg_plain_jacobi_preconditioner(capi, lattice8):
    """PreconditionerType == Jacobi (src/step-50.cc:996-1006): PCG with 0.6 D^-1."""
    from oracle import solver
    P = lattice8
    g = capi.Gmg()
    hand_over(P, g)
    tol = 1e-8 * math.sqrt(P.b @ P.b)
    x, its, r0, r1 = g.pcg_solve_jacobi(P.b, np.zeros_like(P.b), 0.6, 500, tol)
    xo, ito, r0o, r1o = solver.pcg_jacobi(P.system.A, P.b, None)
    assert abs(its - ito) <= 1 and rel_l2(x, xo) < 1e-8
    g.close()


def test_step16_uniform_hierarchy_jacobi(capi, goldens):
    """tests_3D/step-16.mpirun=1.output cycle 0: 8 iterations, 6.0837e-10 (5 uniform levels)."""
    P = oracle_cycle(make_prm("Step16", nref=4, atom="x"), 0, smoother="jacobi")
    gold = goldens["step16_3d"][0]["cycles"][0]
    r = device_solve_cycle(P, "jacobi")
    assert r["its"] == gold["its"] == 8
    assert abs(r["res"] - gold["conv"]) <= 1e-3 * gold["conv"]
    assert abs(np.linalg.norm(r["x"]) - gold["sol_l2"]) < 1e-4


def _chain_matrix(n, period, eliminated):
    """1D Laplacian chains (-1, 2, -1).  eliminated=True: every `period`-th row is an eliminated Dirichlet row (identity,
    the neighbours' couplings to it are dropped) -- the FE situation, the zeroed-operand set of pattern_win.cuh is
    conflict-free.  eliminated=False: the coupling between rows i-1 and i is cut instead, so the column a shortened row
    skips IS needed by its other neighbour: the conflict case (only exact dominant rows may use the windows)."""
    import scipy.sparse as sp
    main = np.full(n, 2.0)
    lo = np.full(n - 1, -1.0)   # entry (i+1, i)
    cut = np.arange(period, n - 1, period)
    if eliminated:
        lo[cut - 1] = 0.0       # (i, i-1)
        lo[cut] = 0.0           # (i+1, i)
        main[cut] = 1.0
    else:
        lo[cut - 1] = 0.0
    A = sp.diags([lo, main, lo], [-1, 0, 1], format="csr")
    A.eliminate_zeros()
    return A


@pytest.mark.parametrize("eliminated", [True, False])
@pytest.mark.parametrize("windows", ["1", "0", "global-codes"])
def test_row_pattern_cg_on_chains_with_shortened_rows(capi, eliminated, windows, monkeypatch):
    """The row-pattern coarse CG (TMA-window kernel and its L1-gather fallback) on matrices whose shortened rows are
    sub-sequences of the dominant row: same iteration count and solution as the plain format, with a right-hand side
    that is NOT zero on the eliminated rows (their operand entries live in the side vector of the window kernel)."""
    import scipy.sparse.linalg as spla
    monkeypatch.setenv("GMG_CG_WIN", "0" if windows == "0" else "1")
    monkeypatch.setenv("GMG_WIN_GLOBAL_CODES", "1" if windows == "global-codes" else "0")  # the > 35 k rows / SM variant
    expected_kernel = {"1": 3, "0": 2, "global-codes": 4}[windows]
    n = 150001
    A = _chain_matrix(n, 97, eliminated)
    rng = np.random.default_rng(2)
    b = rng.standard_normal(n)
    out = {}
    for mode in (0, 2):
        g = capi.Gmg()
        g.set_compression(mode)
        g.set_num_levels(1)
        g.set_matrix(capi.GMG_SYSTEM, 0, A)
        g.set_matrix(capi.GMG_LEVEL, 0, A)
        g.set_copy_indices(0, np.arange(n), np.arange(n))
        g.setup()
        assert g.matrix_traffic(capi.GMG_LEVEL, 0)["format"] == mode
        assert g.coarse_kernel(capi.GMG_LEVEL, 0) == (expected_kernel if mode == 2 else 0)
        out[mode] = g.cg_solve(capi.GMG_LEVEL, 0, b, 1000, 1e-10)
        y = g.spmv(capi.GMG_LEVEL, 0, b, n)
        assert np.array_equal(y, out.setdefault("y", y))  # the SpMV is bit-identical across the formats
        g.close()
    assert out[0][1] == out[2][1]
    assert rel_l2(out[2][0], out[0][0]) < 1e-11
    assert rel_l2(out[2][0], spla.spsolve(A.tocsc(), b)) < 1e-8
