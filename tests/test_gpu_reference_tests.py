"""GPU runs of the reference's other regression tests through the drop-in `LaplaceProblem` (host class + CUDA library):
`tests/test_with_optimal_parameters`, `tests_rhs_rc_variation/rc_variation`, `tests_3D/gaussian-charges`.
Mesh and settings as in tests/test_oracle_reference_tests.py (hyper cube refined 4 times = `Coarse levels below the base
mesh = 4` on a 16^3 lattice without vacuum; Jacobi(0.5) x 2 smoothing, Kelly marking: the build that wrote the goldens)."""
import pytest

import hostlib
from test_oracle_goldens import check_cycle, printed_tol
from test_oracle_reference_tests import hyper_cube_prm

pytestmark = pytest.mark.gpu

OLD_BUILD = ("subsection Solver input data\n set Smoother = Jacobi\nend\n"
             "subsection Misc\n set Refinement indicator = Kelly\n set Energy postprocessing atom limit = 0\nend\n")


@pytest.mark.parametrize("run,flag", [(0, "true"), (1, "false")])
def test_with_optimal_parameters_eight_cycles_every_digit(goldens, run, flag):
    """All 8 adaptive cycles of tests/test_with_optimal_parameters.mpirun=1.output on the device: cells, DoFs per
    level, starting residuals, Jacobi-smoothed iteration counts 7,7,7,9,9,10,9,9 and solution norms to the printed
    digits, convergence values to 1e-3 (the tail of the Krylov recurrence follows the summation order)."""
    gold = goldens["optimal_parameters"][run]
    text, recs = hostlib.run_problem(hyper_cube_prm(-5, 5, 4, 8, flag=flag, extra=OLD_BUILD))
    assert len(recs) == 8
    assert ("Rhs assembly optimization ENABLED" if flag == "true" else "Without rhs assembly optimization") in text
    for rec, g in zip(recs, gold["cycles"]):
        check_cycle(rec, g, conv_rel=1e-3)
    assert "(by level: 8, 27, 125, 729, 4913, 981, 622)" in text


def test_rc_variation_rhs_norms(goldens):
    """tests_rhs_rc_variation: the load vector of 2 atoms on 16^3 cells of [-2.5, 2.5]^3 summed over all atoms (no
    lists) and through lists with the largest cutoff of the sweep (6.0 r_c ... the printed digits are the same)."""
    g = goldens["rc_variation"][0]["cycles"][0]
    for cutoff, flag in ((2.0, "false"), (6.0, "true")):
        _, recs = hostlib.run_problem(hyper_cube_prm(-2.5, 2.5, 4, 1, cutoff=cutoff, flag=flag, extra=OLD_BUILD))
        assert recs[0]["n_dofs_level"] == g["n_dofs_level"]
        assert abs(recs[0]["rhs_l2"] - g["rhs_l2"]) <= printed_tol(g["rhs_l2_digits"])
        assert abs(recs[0]["rhs_linf"] - g["rhs_linf"]) <= printed_tol(g["rhs_linf_digits"])


ZERO_GUESS = "subsection Misc\n set Initial guess = Zero\nend\n"


def test_gaussian_function_without_atoms(goldens):
    """tests_3D/gaussian-charges.mpirun=1.output: no atom file (`GaussianCharges::RightHandSide`), 8 adaptive cycles,
    every printed number of every cycle (iteration counts 7,7,9,9,9,9,10,10) with `Initial guess = Zero` (the build
    behind the file, see the CPU twin of this test)."""
    gold = goldens["gaussian_function_3d"][0]["cycles"]
    text, recs = hostlib.run_problem(hyper_cube_prm(-2.5, 2.5, 4, 8, cutoff=3, atom="", flag="false",
                                                    extra=OLD_BUILD + ZERO_GUESS))
    assert "Unable to open the file." in text and len(recs) == 8
    for rec, g in zip(recs, gold):
        check_cycle(rec, g, conv_rel=1e-3)


def test_step16_all_adaptive_cycles(goldens):
    """tests_3D/step-16.mpirun=1.output through `LaplaceProblem`: all 5 adaptive cycles (8,11,10,14,14 iterations)."""
    from conftest import make_prm
    gold = goldens["step16_3d"][0]["cycles"]
    text, recs = hostlib.run_problem(make_prm("Step16", nref=4, cycles=5, atom="null", extra=OLD_BUILD + ZERO_GUESS))
    assert len(recs) == 5
    for rec, g in zip(recs, gold):
        check_cycle(rec, g, conv_rel=1e-3)


def test_step16_mesh_with_atom_right_hand_side(goldens):
    """tests/step-16.mpirun=1.output through `LaplaceProblem`: the Step16 problem with the Gaussian charge density of two
    atoms as right-hand side (no lists), all 5 adaptive cycles (8,10,11,10,14 iterations)."""
    from conftest import make_prm
    gold = goldens["step16_with_atoms"][0]["cycles"]
    text, recs = hostlib.run_problem(make_prm("Step16", nref=4, cycles=5, atom="atom_2.data", flag="false", cutoff=3,
                                              extra=OLD_BUILD + ZERO_GUESS))
    assert "Number of atoms: 2" in text and len(recs) == 5
    for rec, g in zip(recs, gold):
        check_cycle(rec, g, conv_rel=1e-3)


def test_binning_reproduces_the_cell_data_transfer_lists(goldens):
    """tests/cell_data_transfer_test.mpirun=1.output (the only golden that prints atom lists) through `gmg_bin_atoms`:
    the 2D cells become 3D cells in the plane of the atoms (z = 0: the four vertices at z = h are never the nearest), so
    the device's nearest-vertex test must reproduce the printed lists -- including atom 0, whose distance to a vertex
    differs from the cutoff in the last bit (0.85 - 0.5 against 3.5 * 0.1)."""
    import numpy as np
    from helpers import pkg
    from test_oracle_reference_tests import cell_data_transfer_case
    f, pos2, cutoff = cell_data_transfer_case()
    pos = np.concatenate([pos2, np.zeros((3, 1))], axis=1)
    g = pkg().capi.Gmg()
    try:
        for level, cyc in ((1, goldens["cell_data_transfer"][0]["cycles"][1]),):
            lo = np.concatenate([f.lo + f.ijk[level] * f.h(level), np.zeros((f.n_cells(level), 1))], axis=1)
            ptr, idx = g.bin_atoms(lo, np.full(f.n_cells(level), f.h(level)), pos, cutoff)
            assert [list(idx[ptr[k]:ptr[k + 1]]) for k in range(f.n_cells(level))] == [c["atoms"] for c in cyc["cell_lists"]]
    finally:
        g.close()
