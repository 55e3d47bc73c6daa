"""Pin the CPU oracle against the goldens of the reference's OTHER regression tests:
`tests/test_with_optimal_parameters`, `tests_rhs_rc_variation/rc_variation`, `tests_3D` / `tests_2D`
`gaussian-charges`, and the 3- / 7-rank outputs of `tests/gaussian-charges`.

Those tests override `run()` and build their mesh as `GridGenerator::hyper_cube(left, right)` +
`refine_global(n)` (tests/test_with_optimal_parameters.cc:95-101) instead of the lattice of
src/step-50.cc:1500-1527.  The same forest is reached through the product's `Coarse levels below the base mesh`
key: a lattice of 2^n cells per direction without vacuum whose n coarser levels end in ONE root cell is the
hyper cube refined n times, and the atom lists are made on the cells that are active in cycle 0, as
`rhs_assembly_optimization` does (src/step-50.cc:262-293).  The golden files were written by the build that smoothed
with Jacobi(0.5) x 2 and marked with the Kelly estimator alone (the same build as the cluster logs, DESIGN.md
section 2): with those two settings every printed digit is reproduced."""
import numpy as np
import pytest

from conftest import make_prm
from oracle import problem
from test_oracle_goldens import check_cycle, printed_tol


def hyper_cube_prm(left, right, nref, cycles, cutoff=3.5, flag="true", atom="atom_2.data", dim=3, extra=""):
    """The .prm of a reference test that meshes with hyper_cube(left, right) + refine_global(nref)."""
    return make_prm(left=left, right=right, h=(right - left) / (1 << nref), vacuum=0, cycles=cycles, cutoff=cutoff,
                    flag=flag, atom=atom, dim=dim, bc="Homogeneous",
                    extra=f"subsection Geometry\n set Coarse levels below the base mesh = {nref}\nend\n" + extra)


@pytest.mark.parametrize("run,flag", [(0, "true"), (1, "false")])
def test_with_optimal_parameters_eight_cycles_every_digit(goldens, run, flag):
    """tests/test_with_optimal_parameters.mpirun=1.output: 2 atoms in [-5, 5]^3, 16^3 cells, 8 adaptive cycles, first
    with and then without the atom lists: cells, DoFs per level, starting residuals, iteration counts
    7,7,7,9,9,10,9,9, convergence values and the three solution norms of every cycle."""
    gold = goldens["optimal_parameters"][run]
    assert gold["n_atoms"] == 2 and len(gold["cycles"]) == 8
    P = problem.from_prm_string(hyper_cube_prm(-5, 5, 4, 8, flag=flag), smoother="jacobi", indicator="kelly")
    recs = P.run(energy_gate=0)
    for rec, g in zip(recs, gold["cycles"]):
        check_cycle(rec, g)
    assert [r["its"] for r in recs] == [7, 7, 7, 9, 9, 10, 9, 9]
    assert recs[-1]["n_active_cells"] == 22786


def test_rc_variation_rhs_norms(goldens):
    """tests_rhs_rc_variation/rc_variation.mpirun=1.output: || b ||_2 and || b ||_inf of 2 atoms on 16^3 cells of
    [-2.5, 2.5]^3.  The file holds the same two numbers for all 17 cutoff parameters 2.0 ... 6.0: the sum over all
    atoms.  Without lists every digit is reproduced; with lists the largest cutoff (3 = 6.0 r_c: exp(-36)) agrees
    to the printed digits, and the smallest (1 = 2.0 r_c) must differ: the truncation is real."""
    runs = goldens["rc_variation"]
    assert len(runs) == 17
    g = runs[0]["cycles"][0]
    assert all(r["cycles"][0]["rhs_l2_digits"] == g["rhs_l2_digits"] for r in runs)

    def rhs_norms(cutoff, flag):
        P = problem.from_prm_string(hyper_cube_prm(-2.5, 2.5, 4, 1, cutoff=cutoff, flag=flag), smoother="jacobi")
        rec = P.run(stop_after_solve_of_last_cycle=True)[0]
        assert rec["n_active_cells"] == g["n_active_cells"] and rec["n_dofs_level"] == g["n_dofs_level"]
        return rec["rhs_l2"], rec["rhs_linf"]

    for cutoff, flag in ((2.0, "false"), (6.0, "true")):
        l2, linf = rhs_norms(cutoff, flag)
        assert abs(l2 - g["rhs_l2"]) <= printed_tol(g["rhs_l2_digits"])
        assert abs(linf - g["rhs_linf"]) <= printed_tol(g["rhs_linf_digits"])
    l2, _ = rhs_norms(2.0, "true")
    assert 1e-5 < abs(l2 - g["rhs_l2"]) < 1e-3


@pytest.mark.parametrize("key,dim", [("gaussian_function_3d", 3), ("gaussian_function_2d", 2)])
def test_gaussian_function_without_atoms(goldens, key, dim):
    """tests_3D / tests_2D gaussian-charges.mpirun=1.output: no atom file, the right-hand side is
    GaussianCharges::RightHandSide (include/step_50.h:321-329); 8 adaptive cycles, every printed number of every cycle
    (iteration counts 7,7,9,9,9,9,10,10 in 3D, 6,6,7,7,8,8,8,8 in 2D).  The build behind these files started the solve
    of every cycle from zero (its starting residual of cycle 1 is || b ||, not the residual of an interpolated
    solution): `initial_guess="zero"`; with the shipped source's transferred guess the meshes and solution norms are the
    same and the starting residuals smaller."""
    gold = goldens[key][0]["cycles"]
    assert len(gold) == 8
    prm = hyper_cube_prm(-2.5, 2.5, 4, 8, cutoff=3, atom="", dim=dim)
    recs = problem.from_prm_string(prm, smoother="jacobi", indicator="kelly", initial_guess="zero").run()
    for rec, g in zip(recs, gold):
        check_cycle(rec, g, conv_rel=1e-4)
    if dim == 2:
        shipped = problem.from_prm_string(prm, smoother="jacobi", indicator="kelly").run()
        for rec, old, g in zip(shipped, recs, gold):
            assert rec["n_dofs_level"] == g["n_dofs_level"]
            assert abs(rec["sol_l2"] - g["sol_l2"]) <= printed_tol(g["sol_l2_digits"]) + 1e-7 * g["sol_l2"]
            assert rec["cycle"] == 0 or rec["start"] < old["start"]


@pytest.mark.parametrize("key,dim", [("step16_3d", 3), ("step16_2d", 2)])
def test_step16_all_adaptive_cycles(goldens, key, dim):
    """tests_3D / tests_2D step-16.mpirun=1.output: all 5 adaptive cycles of the Step16 problem (coefficient jump,
    f = 10) -- cells, DoFs per level, starting residuals, iteration counts 8,11,10,14,14 / 7,10,11,13,13, convergence
    values, solution norms -- with the settings of the build that wrote them (Jacobi smoothing, Kelly marking, zero
    initial guess).  test_oracle_goldens.py pins cycle 0 with the defaults."""
    gold = goldens[key][0]["cycles"]
    assert len(gold) == 5
    P = problem.from_prm_string(make_prm("Step16", dim=dim, nref=4, cycles=5, atom="x"), smoother="jacobi",
                                indicator="kelly", initial_guess="zero")
    recs = P.run()
    for rec, g in zip(recs, gold):
        check_cycle(rec, g, conv_rel=1e-4)


def test_step16_mesh_with_atom_right_hand_side(goldens):
    """tests/step-16.mpirun=1.output: `Problem = Step16` (unit cube refined 4 times, coefficient jump, homogeneous BC) run
    WITH an atom file, so the right-hand side is the Gaussian charge density of the 2 atoms of tests/atom_2.data summed
    over all atoms (that build had no atom lists): every printed number of all 5 adaptive cycles (8,10,11,10,14
    iterations; Jacobi smoothing, Kelly marking, zero initial guess)."""
    gold = goldens["step16_with_atoms"][0]
    assert gold["n_atoms"] == 2 and len(gold["cycles"]) == 5
    P = problem.from_prm_string(make_prm("Step16", nref=4, cycles=5, atom="atom_2.data", flag="false", cutoff=3),
                                smoother="jacobi", indicator="kelly", initial_guess="zero")
    recs = P.run(energy_gate=0)
    for rec, g in zip(recs, gold["cycles"]):
        check_cycle(rec, g, conv_rel=1e-4)


@pytest.mark.parametrize("key", ["gaussian_charges_mpirun3", "gaussian_charges_mpirun7"])
def test_rank_independent_numbers_of_the_parallel_goldens(goldens, key):
    """tests/gaussian-charges.mpirun=3 / 7.output against the 1-rank file the oracle reproduces digit by digit
    (test_oracle_goldens.py): meshes, rhs / matrix norms, thresholds and energies do not depend on the rank count; the
    processor-block SSOR changes iteration counts by at most one and the solution within the CG tolerance."""
    one = goldens["gaussian_charges_mpirun1"][0]["cycles"]
    par = goldens[key][0]["cycles"]
    assert len(par) == len(one) == 6
    for a, b in zip(one, par):
        assert a["n_active_cells"] == b["n_active_cells"] and a["n_dofs_level"] == b["n_dofs_level"]
        for k in ("rhs_l1", "rhs_l2", "rhs_linf", "mat_l1", "mat_linf", "start", "threshold", "energy_analytic",
                  "energy_short", "energy_self"):
            # the last printed digit may round the other way (the starting residual of cycle 4 does)
            assert abs(a[k] - b[k]) <= 2.1 * printed_tol(a[k + "_digits"]), (a["cycle"], k)
        assert abs(a["mat_frob"] - b["mat_frob"]) <= 2e-10 * a["mat_frob"]
        assert abs(a["its"] - b["its"]) <= 1
        for k in ("sol_l1", "sol_l2", "sol_linf", "energy_fe", "energy_total", "energy_norm_error"):
            assert abs(a[k] - b[k]) <= 1e-8 * abs(a[k]), (a["cycle"], k)


def test_host_builds_the_same_hyper_cube_hierarchy_as_the_oracle():
    """The product's host substrate on the forest of test_with_optimal_parameters, refined with the oracle's flags of
    the first four cycles (hanging nodes, 2:1 balance across patch levels): cells, numbering, level numbering,
    copy indices and hanging-node flags are the oracle's, bit for bit -- the GPU run of the same golden
    (tests/test_gpu_reference_tests.py) then differs from the oracle only in floating-point summation orders."""
    import numpy as np
    import hostlib
    P = problem.from_prm_string(hyper_cube_prm(-5, 5, 4, 5), smoother="jacobi", indicator="kelly")
    P.run(stop_after_solve_of_last_cycle=True, energy_gate=0)
    f, d = P.forest, P.dofs
    M = hostlib.Ministep(1, -5.0, 5.0)
    M.refine_global(4)
    for flags in P.flag_history:
        M.refine(flags)
    M.build()
    assert M.n_levels == f.n_levels == 7
    assert list(M.get("level_n")) == list(d.level_n)
    for l in range(M.n_levels):
        assert np.array_equal(M.get("ijk", l).reshape(-1, 3), f.ijk[l])
        assert np.array_equal(M.get("cell_dofs", l).reshape(-1, 8), d.cell_dofs[l])
        assert np.array_equal(M.get("level_cell_dofs", l).reshape(-1, 8), d.level_cell_dofs[l])
        assert np.array_equal(M.get("copy_global", l), d.copy_global[l])
    assert np.array_equal(M.get("dof_xyz").reshape(-1, 3), d.xyz)
    assert np.array_equal(M.get("hanging").astype(bool), d.hanging)


def cell_data_transfer_case():
    """Mesh and atoms of tests/cell_data_transfer_test.prm: hyper_cube(0, 1) in 2D refined once (4 cells), the 3 atoms of
    tests/atom_3.data, cutoff 3.5 r_c with r_c = 0.1; returns (forest, positions, cutoff)."""
    import os
    from conftest import GOLDEN
    from oracle import lammps
    from oracle.mesh import Forest
    pos, q, _ = lammps.read(os.path.join(GOLDEN, "atom_3.data"))
    assert np.array_equal(pos[:, 2], np.zeros(3)) and list(q) == [1.0, -1.0, 1.0]
    f = Forest(1, 0.0, 1.0, 2)
    f.refine_global(1)
    return f, pos[:, :2], 3.5 * 0.1


def literal_lists(f, level, pos, cutoff):
    """src/step-50.cc:262-293 word for word: a cell lists atom i iff one of its vertices is closer than the cutoff."""
    verts = f.real_coords(f.vertex_coords(level))  # (cells, 2^dim, dim)
    ptr, idx = [0], []
    for c in range(len(verts)):
        for i in range(len(pos)):
            if any(np.sqrt(((pos[i] - v) ** 2).sum()) < cutoff for v in verts[c]):
                idx.append(i)
        ptr.append(len(idx))
    return np.array(ptr, dtype=np.int64), np.array(idx, dtype=np.int64)


def test_cell_data_transfer_atom_lists(goldens):
    """tests/cell_data_transfer_test.mpirun=1 / 3.output: the ONLY golden that prints atom lists.  Cycle 1 prints the
    list of each of the 4 cells of the first mesh (the binning rule, including the atom that sits 0.35 - 2e-17 from a
    vertex with the cutoff at 0.35 + 3e-17), cycle 2 the lists of their 16 children (every child carries its parent's
    list, src/step-50.cc:441-449) in the order the cells are visited."""
    from oracle import rhs
    f, pos, cutoff = cell_data_transfer_case()
    lists1 = literal_lists(f, 1, pos, cutoff)
    f.refine([np.zeros(1, dtype=bool), np.ones(4, dtype=bool)])
    assert f.n_cells(2) == 16
    lists = rhs.inherit_lists(f, lists1, base_level=1)
    for key in ("cell_data_transfer", "cell_data_transfer_mpirun3"):
        cyc = goldens[key][0]["cycles"]
        assert goldens[key][0]["n_atoms"] == 3
        for level, c in ((1, cyc[1]), (2, cyc[2])):
            ptr, idx = lists[level]
            centres = f.lo + (f.ijk[level] + 0.5) * f.h(level)
            assert len(c["cell_lists"]) == f.n_cells(level)
            mine = {tuple(centres[k]): list(idx[ptr[k]:ptr[k + 1]]) for k in range(f.n_cells(level))}
            assert mine == {tuple(gl["center"]): gl["atoms"] for gl in c["cell_lists"]}, (key, level)
            if key == "cell_data_transfer":  # one rank: also the visiting order (parent by parent, children in z-order)
                assert [tuple(gl["center"]) for gl in c["cell_lists"]] == list(mine)


def test_multi_rank_goldens_covered_by_the_one_rank_reproduction(goldens):
    """The 3- and 7-rank outputs of the Jacobi-smoothed reference tests carry, line by line, the numbers of their
    1-rank outputs (checked when the goldens were transcribed, tests/golden/make_reference_goldens.py): reproducing the
    1-rank files digit by digit -- the tests above and in test_oracle_goldens.py -- reproduces these 16 files as well."""
    same = set(goldens["same_numbers_as_one_rank"])
    for stem in ("tests_3D/step-16", "tests_2D/step-16", "tests_3D/gaussian-charges", "tests_2D/gaussian-charges",
                 "tests/step-16", "tests/test_with_optimal_parameters"):
        assert {"%s.mpirun=%d.output" % (stem, r) for r in (3, 7)} <= same
    assert "tests_rhs_rc_variation/rc_variation.mpirun=3.output" in same
    assert len(same) == 16
