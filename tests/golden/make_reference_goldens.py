"""Transcribe the reference's golden stdout files into tests/golden/reference_goldens.json.

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_reference_goldens.py
Every number is kept as the printed string's float together with file and line.  Also writes the
small LAMMPS atom files the golden runs use (regenerated from their published coordinates by
oracle.lammps.write, not copied)."""
import json
import os
import re
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

KEYS = [
    (r"Number of active cells:\s+(\d+)", "n_active_cells", int),
    (r"Number of degrees of freedom:\s+(\d+) \(by level: ([\d, ]+)\)", "n_dofs", None),
    (r"L1 rhs norm (\S+)", "rhs_l1", float), (r"L2 rhs norm (\S+)", "rhs_l2", float),
    (r"LInfinity rhs norm (\S+)", "rhs_linf", float),
    (r"L1 Matrix norm (\S+)", "mat_l1", float), (r"LInfinity Matrix norm (\S+)", "mat_linf", float),
    (r"Frobenius Matrix norm (\S+)", "mat_frob", float),
    (r"Starting value (\S+)", "start", float), (r"CG converged in (\d+) iterations", "its", int),
    (r"Convergence value (\S+)", "conv", float),
    (r"L1 solution norm (\S+)", "sol_l1", float), (r"L2 solution norm (\S+)", "sol_l2", float),
    (r"LInfinity solution norm (\S+)", "sol_linf", float),
    (r"Threshold value for refinement:\s+(\S+)", "threshold", float),
    (r"Total analytical electrostatic energy :\s+(\S+)", "energy_analytic", float),
    (r"Short-ranged energy contribution :\s+(\S+)", "energy_short", float),
    (r"FE solution long-ranged energy contribution :\s+(\S+)", "energy_fe", float),
    (r"Self energy contribution :\s+(\S+)", "energy_self", float),
    (r"Total electrostatic energy with split in short- and long-ranged :\s+(\S+)", "energy_total", float),
    (r"Error in FE solution in energy norm:\s+(\S+)", "energy_norm_error", float),
]


def parse(relpath):
    runs, run, cyc = [], None, None
    with open(os.path.join(REF, relpath)) as f:
        for no, line in enumerate(f, 1):
            if line.startswith("Problem type is:"):
                run = dict(file=relpath, line=no, cycles=[])
                runs.append(run)
                continue
            if run is None:
                continue
            m = re.match(r"Number of atoms: (\d+)", line)
            if m:
                run["n_atoms"] = int(m.group(1))
            m = re.match(r"Running with \w+ on (\d+) MPI", line)
            if m:
                run["ranks"] = int(m.group(1))
            m = re.match(r"Cycle (\d+):", line)
            if m:
                cyc = dict(cycle=int(m.group(1)), line=no)
                run["cycles"].append(cyc)
                continue
            if cyc is None:
                continue
            # tests/cell_data_transfer_test: the atom list every cell hands to its children, printed while the
            # refinement of this cycle packs the parents' data ("cell with center x y has n values:" + the list)
            m = re.match(r"cell with center (.+) has (\d+) values:", line)
            if m:
                cyc.setdefault("cell_lists", []).append(dict(center=[float(t) for t in m.group(1).split()],
                                                             n=int(m.group(2)), atoms=None, line=no))
                continue
            if cyc.get("cell_lists") and cyc["cell_lists"][-1]["atoms"] is None:
                cyc["cell_lists"][-1]["atoms"] = [int(t) for t in line.split()]
                assert len(cyc["cell_lists"][-1]["atoms"]) == cyc["cell_lists"][-1]["n"]
                continue
            for pat, key, conv in KEYS:
                m = re.search(pat, line)
                if m:
                    if key == "n_dofs":
                        cyc["n_dofs"] = int(m.group(1))
                        cyc["n_dofs_level"] = [int(t) for t in m.group(2).split(",")]
                    else:
                        cyc[key] = conv(m.group(1))
                        cyc[key + "_digits"] = m.group(1)
                    break
    return runs


FILES = {
    "gaussian_charges_mpirun1": "tests/gaussian-charges.mpirun=1.output",
    "gaussian_charges_mpirun3": "tests/gaussian-charges.mpirun=3.output",
    "gaussian_charges_mpirun7": "tests/gaussian-charges.mpirun=7.output",
    "step16_3d": "tests_3D/step-16.mpirun=1.output",
    "step16_with_atoms": "tests/step-16.mpirun=1.output",
    "step16_2d": "tests_2D/step-16.mpirun=1.output",
    "gaussian_function_3d": "tests_3D/gaussian-charges.mpirun=1.output",
    "gaussian_function_2d": "tests_2D/gaussian-charges.mpirun=1.output",
    "optimal_parameters": "tests/test_with_optimal_parameters.mpirun=1.output",
    "rc_variation": "tests_rhs_rc_variation/rc_variation.mpirun=1.output",
    "cell_data_transfer": "tests/cell_data_transfer_test.mpirun=1.output",
    "cell_data_transfer_mpirun3": "tests/cell_data_transfer_test.mpirun=3.output",
    "cluster_ssor_run": "Cluster runs output and postprocessing/SSOR_run.o876223",
    "cluster_ssor_64k": "Cluster runs output and postprocessing/SSOR_64k_atoms.o876224",
    "cluster_without_opti": "Cluster runs output and postprocessing/without_opti.o875054",
}

# Reference tests whose 3- and 7-rank outputs carry the same numbers as the 1-rank output (they smooth with Jacobi, which
# does not depend on the partition): compared line by line here, wall-clock lines and blank lines aside; the list of
# files found identical is stored so that the tests can state which multi-rank goldens the 1-rank reproduction covers.
RANK_INDEPENDENT = ["tests_3D/step-16", "tests_2D/step-16", "tests_3D/gaussian-charges", "tests_2D/gaussian-charges",
                    "tests/step-16", "tests/test_with_optimal_parameters", "tests_rhs_rc_variation/rc_variation",
                    "tests_rc_variation/rc_variation"]


def numbers_only(relpath):
    with open(os.path.join(REF, relpath)) as f:
        return [l.rstrip() for l in f if l.strip() and not l.startswith("Elapsed wall time")]


def same_as_one_rank():
    out = []
    for stem in RANK_INDEPENDENT:
        one = numbers_only(stem + ".mpirun=1.output")
        for ranks in (3, 7):
            rel = "%s.mpirun=%d.output" % (stem, ranks)
            if os.path.exists(os.path.join(REF, rel)):
                assert numbers_only(rel) == one, rel
                out.append(rel)
    return out


if __name__ == "__main__":
    out = {k: parse(v) for k, v in FILES.items()}
    out["same_numbers_as_one_rank"] = same_as_one_rank()
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(out, f, indent=1)
    from oracle import lammps
    import numpy as np
    for src, dst in (("tests/atom_n1_2.data", "atom_n1_2.data"), ("tests/atom_2.data", "atom_2.data"),
                     ("tests/atom_3.data", "atom_3.data"),
                     ("atom/atom_n1_8.data", "atom_n1_8.data")):
        pos, q, _ = lammps.read(os.path.join(REF, src))
        lammps.write(os.path.join(HERE, dst), pos, q)
        p2, q2, _ = lammps.read(os.path.join(HERE, dst))
        assert np.array_equal(pos, p2) and np.array_equal(q, q2)
    print({k: [len(r["cycles"]) for r in v] for k, v in out.items() if k != "same_numbers_as_one_rank"})
    print(out["same_numbers_as_one_rank"])
