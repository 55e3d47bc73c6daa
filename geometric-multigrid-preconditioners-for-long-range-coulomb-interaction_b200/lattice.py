"""Synthetic inputs of the benchmark configs: the NaCl rock-salt lattices of the reference's `atom/` directory
(spacing 0.5, unit cells ordered x slowest / z fastest, 8-atom basis, charge +1 where 2(x+y+z) is even), written
in LAMMPS "full" format, and the calibrated cluster parameter file (SURVEY.md sections 0.8 and 8d)."""
import numpy as np

BASIS = np.array([(0, 0, 0), (.5, 0, 0), (.5, .5, 0), (0, .5, 0), (.5, 0, .5), (0, 0, .5), (0, .5, .5), (.5, .5, .5)])


def nacl_lattice(n):
    g = np.arange(n)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    pos = (cells[:, None, :] + BASIS[None, :, :]).reshape(-1, 3).astype(np.float64)
    q = np.where(np.rint(2 * pos.sum(1)).astype(np.int64) % 2 == 0, 1.0, -1.0)
    return pos, q


def write_lammps(path, pos, q):
    n = len(q)
    hi = float(np.ceil(pos.max() + 0.5)) if n else 1.0
    with open(path, "w") as f:
        f.write("LAMMPS Description\n\n")
        f.write(f"     {n}  atoms\n     0  bonds\n     0  angles\n     0  dihedrals\n     0  impropers\n\n")
        f.write("     2  atom types\n\n")
        for ax in "xyz":
            f.write(f"  0.0 {hi:.1f} {ax}lo {ax}hi\n")
        f.write("\nMasses\n\n      1\t\t22.989\n      2 \t35.453\n\nAtoms # full\n\n")
        lines = [f"{i + 1} {i + 1} {1 if q[i] > 0 else 2} {q[i]:.1f} {pos[i, 0]:.1f} {pos[i, 1]:.1f} {pos[i, 2]:.1f}\n"
                 for i in range(n)]
        f.write("".join(lines))


def cluster_prm(atom_file, n, cycles=5, smoother="MulticolourSSOR", indicator="Kelly", nq_param=1, cutoff=3.5,
                vacuum=10, mesh_size=0.25, device=0, assembly="Host", coarse_levels=0, energy_limit=300, energy_norm_limit=300,
                flag="true", n_gpus=1):
    """The parameter file of the reference's cluster runs for the n^3-unit-cell lattice (domain [0, n]).  The builds behind
    the cluster logs printed neither energies above 300 atoms nor the energy-norm error (no such line in the logs), hence the
    two atom limits."""
    return f"""
subsection Geometry
  set Number of global refinement = 0
  set Domain limit left = 0
  set Domain limit right = {n}
  set Mesh size = {mesh_size}
  set Vacuum repetitions = {vacuum}
  set Coarse levels below the base mesh = {coarse_levels}
end
subsection Misc
  set Number of Adaptive Refinement = {cycles}
  set smoothing length = 0.5
  set Nonzero Density radius parameter around each charge = {cutoff}
  set Flag for RHS evaluation optimization = {flag}
  set Quadrature points for RHS function = {nq_param}
  set Output time summary table = false
  set Refinement indicator = {indicator}
  set Energy postprocessing atom limit = {energy_limit}
  set Energy norm error atom limit = {energy_norm_limit}
end
set Polynomial degree = 1
subsection Solver input data
  set Preconditioner = GMG
  set Smoother = {smoother}
  set GPU device = {device}
  set Matrix assembly = {assembly}
end
subsection Problem Selection
  set Problem = GaussianCharges
  set Dimension = 3
  set Boundary conditions selection = Homogeneous
end
subsection Lammps data
  set Lammps input file = {atom_file}
end
"""
