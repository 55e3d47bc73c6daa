"""GPU end-to-end tests of the drop-in: the C++ `LaplaceProblem` (host) + CUDA library reproduce the
reference's golden stdout, cycle by cycle, like the reference's own regression tests."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, make_prm
from helpers import pkg
import hostlib
from test_oracle_goldens import check_cycle

pytestmark = pytest.mark.gpu


def test_two_atoms_six_cycles_reproduce_reference_golden(goldens):
    """tests/gaussian-charges.cc: all printed numbers of all 6 adaptive cycles (SSOR = lexicographic,
    level-scheduled on the device; iteration counts 1,6,7,6,7,7)."""
    gold = goldens["gaussian_charges_mpirun1"][0]
    text, recs = hostlib.run_problem(make_prm(cycles=6, bc="Exact", atom="atom_n1_2.data", nq=4))
    assert len(recs) == 6
    for rec, g in zip(recs, gold["cycles"]):
        check_cycle(rec, g)
    assert recs[0]["coarse_its"] == [112]
    assert "Threshold value for refinement:\t4.4445997238e+00" in text
    assert "   Starting value 0.9219052997" in text
    assert "   CG converged in 6 iterations." in text


@pytest.mark.parametrize("smoother,expected", [("Jacobi", None), ("MulticolourSSOR", None), ("Chebyshev", None)])
def test_other_smoothers_same_solution(goldens, smoother, expected):
    gold = goldens["gaussian_charges_mpirun1"][0]["cycles"]
    extra = f"subsection Solver input data\n set Smoother = {smoother}\nend\n"
    _, recs = hostlib.run_problem(make_prm(cycles=3, bc="Exact", atom="atom_n1_2.data", nq=4, extra=extra))
    for rec, g in zip(recs, gold):
        # the mesh sequence, rhs and matrices do not depend on the smoother; the solution agrees to the CG tolerance
        assert rec["n_dofs_level"] == g["n_dofs_level"]
        assert abs(rec["rhs_l2"] - g["rhs_l2"]) < 1e-9
        assert abs(rec["sol_l2"] - g["sol_l2"]) < 1e-6 * g["sol_l2"]
        assert abs(rec["its"] - g["its"]) <= 3
        assert abs(rec["threshold"] - g["threshold"]) < 1e-6 * g["threshold"]


def test_main_executable_lattice_8_atoms_reproduces_cluster_log(tmp_path, goldens):
    """`main file.prm` on atom_n1_8 (configs[0]) with the calibrated cluster parameters and the
    Kelly-only marking of the build behind the logs: cell / DoF counts and starting residuals of all 5
    cycles of SSOR_run.o876223; iteration counts within +-2 (the log ran 20-rank block SSOR)."""
    prm = tmp_path / "gaussian-charges.prm"
    extra = "subsection Misc\n set Refinement indicator = Kelly\nend\n"
    prm.write_text(make_prm(cycles=5, atom=os.path.join(GOLDEN, "atom_n1_8.data"), extra=extra))
    exe = os.path.join(os.path.dirname(pkg().capi.LIB_PATH), "main")
    out = subprocess.run([exe, str(prm)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    g = goldens["cluster_ssor_run"][0]["cycles"]
    assert "Number of atoms: 8" in out.stdout
    start = [float(x) for x in re.findall(r"Starting value (\S+)", out.stdout)]
    its = [int(x) for x in re.findall(r"CG converged in (\d+) iterations", out.stdout)]
    cells = [int(x) for x in re.findall(r"Number of active cells:\s+(\d+)", out.stdout)]
    levels = re.findall(r"\(by level: ([\d, ]+)\)", out.stdout)
    l2 = [float(x) for x in re.findall(r"L2 solution norm (\S+)", out.stdout)]
    for c in range(5):
        assert cells[c] == g[c]["n_active_cells"]
        assert [int(t) for t in levels[c].split(",")] == g[c]["n_dofs_level"]
        assert abs(start[c] - g[c]["start"]) < 2e-9 + 1e-6 * (c == 0)
        assert abs(its[c] - g[c]["its"]) <= 2
        assert abs(l2[c] - g[c]["sol_l2"]) < 2e-7 * g[c]["sol_l2"]
    assert its[0] == 1


@pytest.mark.parametrize("assembly", ["Host", "Device"])
def test_64k_atoms_five_cycles_reproduce_cluster_log(tmp_path, goldens, assembly):
    """BASELINE config 4 at full size (atom_n20_64000, the SSOR_64k_atoms.o876224 run): cell / DoF counts per level,
    ||b||_2 of cycle 0, starting residuals and solution norms of all 5 cycles to the printed digits; CG iteration
    counts within +-2 (the log ran 20-rank block SSOR; here: multicolour SSOR).  assembly = Device: system and level-0
    matrices of every cycle assembled on the GPU (gmg_assemble_matrix) instead of on the host."""
    P = pkg()
    pos, q = P.lattice.nacl_lattice(20)
    atom = tmp_path / "atom_n20_64000.data"
    P.lattice.write_lammps(str(atom), pos, q)
    text, recs = hostlib.run_problem(P.lattice.cluster_prm(str(atom), 20, cycles=5, smoother="MulticolourSSOR",
                                                             assembly=assembly))
    g = goldens["cluster_ssor_64k"][0]["cycles"]
    assert goldens["cluster_ssor_64k"][0]["n_atoms"] == 64000 and len(recs) == 5
    for c, (rec, gold) in enumerate(zip(recs, g)):
        assert rec["n_active_cells"] == gold["n_active_cells"]
        assert rec["n_dofs_level"] == gold["n_dofs_level"]
        digits = gold["start_digits"]
        tol = 0.5000001 * 10.0 ** (-len(digits.split(".")[1])) + 2e-9
        assert abs(rec["start"] - gold["start"]) <= tol, (c, rec["start"], digits)
        assert abs(rec["its"] - gold["its"]) <= 2
        for k in ("sol_l1", "sol_l2", "sol_linf"):
            assert abs(rec[k] - gold[k]) <= 2e-7 * gold[k], (c, k, rec[k], gold[k])
    assert recs[0]["its"] == 1 and recs[0]["coarse_its"] == [243]
    assert abs(recs[0]["conv"] - g[0]["conv"]) <= 1e-4 * g[0]["conv"]


def test_missing_parameter_file_and_bad_device_fail_loudly(tmp_path):
    exe = os.path.join(os.path.dirname(pkg().capi.LIB_PATH), "main")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode != 0 and "Invalid inputs" in out.stderr
    extra = "subsection Solver input data\n set GPU device = 99\nend\n"
    with pytest.raises(hostlib.HostError, match="no CPU fallback"):
        hostlib.run_problem(make_prm(cycles=1, extra=extra))


def test_coarse_levels_below_the_base_mesh_same_problem_fewer_coarse_iterations(tmp_path):
    """`Coarse levels below the base mesh = 2` on atom_n1_8 with the cluster parameters (base lattice 44^3
    = 4 x 11^3): same meshes, load vectors and solutions as the reference hierarchy (the dof numbering differs, so norms
    agree to rounding / to the CG tolerance); the coarse-grid CG now runs on 12^3 instead of 45^3 dofs."""
    P = pkg()
    atom = os.path.join(GOLDEN, "atom_n1_8.data")
    _, ref = hostlib.run_problem(P.lattice.cluster_prm(atom, 1, cycles=3, smoother="MulticolourSSOR"))
    extra = "subsection Geometry\n set Coarse levels below the base mesh = 2\nend\n"
    _, mg = hostlib.run_problem(P.lattice.cluster_prm(atom, 1, cycles=3, smoother="MulticolourSSOR") + extra)
    assert len(mg) == len(ref) == 3
    for a, b in zip(mg, ref):
        assert a["n_active_cells"] == b["n_active_cells"] and a["n_dofs"] == b["n_dofs"]
        assert a["n_dofs_level"][:2] == [12 ** 3, 23 ** 3] and a["n_dofs_level"][2:] == b["n_dofs_level"]
        for k in ("rhs_l1", "rhs_l2", "rhs_linf", "mat_l1", "mat_linf", "mat_frob"):
            assert abs(a[k] - b[k]) <= 1e-12 * abs(b[k]), k
        assert abs(a["start"] - b["start"]) <= 1e-9 * abs(b["start"]) + 1e-12
        for k in ("sol_l1", "sol_l2", "sol_linf"):
            assert abs(a[k] - b[k]) <= 2e-7 * abs(b[k]), k
        assert abs(a["threshold"] - b["threshold"]) <= 1e-5 * abs(b["threshold"])
        assert a["its"] <= b["its"] + 6
        assert max(a["coarse_its"]) < max(b["coarse_its"])
    out = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "coarse_levels.json"), "w") as f:
            import json
            json.dump({"reference_hierarchy": [{k: r[k] for k in ("its", "coarse_its", "solve_seconds")} for r in ref],
                       "two_coarse_levels": [{k: r[k] for k in ("its", "coarse_its", "solve_seconds")} for r in mg]}, f)
