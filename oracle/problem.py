"""Oracle mirror of `Step50::LaplaceProblem<dim>` (include/step_50.h:111-202): the same
setup / assemble / solve / estimate / refine cycle as `run()` (src/step-50.cc:1463-1573),
returning per-cycle records and the reference's stdout lines.

TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import math
import os

import numpy as np
from scipy.special import erf

from . import assemble, estimate, lammps, prm as prm_mod, rhs, solver
from .dofs import DoFs
from .mesh import Forest


def step16_coefficient(xq):
    """Step16::Coefficient (include/step_50.h:246-254): 5 inside |p| < 0.5, else 1."""
    return np.where((xq ** 2).sum(-1) < 0.25, 5.0, 1.0)


def gaussian_rhs_function(xq, r_c):
    """GaussianCharges::RightHandSide (include/step_50.h:321-329), used when no atom file is read."""
    c = (xq ** 2).sum(-1) / (r_c * r_c)
    return (8.0 * np.exp(-4.0 * c) - np.exp(-c)) / (r_c ** 3 * math.pi ** 1.5)


class LaplaceProblem:
    def __init__(self, params, smoother="ssor", omega=0.5, smoothing_steps=2, ssor_ranks=1, verbose=False,
                 base_dir=None, indicator="kelly+residual", initial_guess="transfer"):
        # indicator="kelly", initial_guess="zero": the older builds behind some of the golden files (tests/ docstrings)
        self.indicator = indicator
        self.initial_guess = initial_guess
        p = self.p = params
        g = lambda s, k: p[(s, k)]
        self.dim = g("Problem Selection", "Dimension")
        self.problem = g("Problem Selection", "Problem")
        self.bc = g("Problem Selection", "Boundary conditions selection")
        self.precond = g("Solver input data", "Preconditioner")
        self.n_global = g("Geometry", "Number of global refinement")
        self.left, self.right = g("Geometry", "Domain limit left"), g("Geometry", "Domain limit right")
        self.mesh_h = g("Geometry", "Mesh size")
        self.vacuum = g("Geometry", "Vacuum repetitions")
        self.coarse_levels = p.get(("Geometry", "Coarse levels below the base mesh"), 0)
        self.n_cycles = g("Misc", "Number of Adaptive Refinement")
        self.r_c = g("Misc", "smoothing length")
        self.cutoff = g("Misc", "Nonzero Density radius parameter around each charge")
        self.flag_rhs = g("Misc", "Flag for RHS evaluation optimization")
        self.nq_rhs = g("", "Polynomial degree") + g("Misc", "Quadrature points for RHS function")
        self.atom_file = g("Lammps data", "Lammps input file")
        if base_dir and not os.path.isabs(self.atom_file):
            self.atom_file = os.path.join(base_dir, self.atom_file)
        self.smoother, self.omega, self.steps, self.ssor_ranks = smoother, omega, smoothing_steps, ssor_ranks
        self.verbose = verbose
        self.lines = []
        self.records = []
        self.flag_history = []
        self.out("Problem type is:   " + self.problem)
        self.out("Preconditioner :    " + self.precond)
        self.out("Rhs assembly optimization ENABLED" if self.flag_rhs else "Without rhs assembly optimization")

    def out(self, s):
        self.lines.append(s)
        if self.verbose:
            print(s, flush=True)

    # ----------------------------------------------------------------- phases
    def read_lammps_input_file(self):
        self.lammps = False
        if self.dim == 3 and os.path.isfile(self.atom_file):
            self.pos, self.charges, self.types = lammps.read(self.atom_file)
            self.lammps = True
            self.out(f"Number of atoms: {len(self.charges)}")
        elif self.dim == 3:
            self.out("Unable to open the file.")
        else:
            self.out("\nReading of Lammps input file implemented for 3D only\n")

    def make_mesh(self):
        if self.problem == "Step16":
            self.forest = Forest(1, self.left, self.right, self.dim)
            self.forest.refine_global(self.n_global)
        else:
            a = 2 * self.mesh_h
            N = (self.right - self.left) / a
            M = self.vacuum
            reps = int(2 * (N + 2 * M))
            k = self.coarse_levels  # the product's optional deeper hierarchy: level k is the reference's base lattice
            assert reps % (1 << k) == 0
            self.forest = Forest(reps >> k, self.left - M * a, self.right + M * a, self.dim)
            if k:
                self.forest.refine_global(k)

    def base_level(self):
        return 0 if self.problem == "Step16" else self.coarse_levels

    def boundary_values(self):
        d = self.dofs
        g = np.zeros(d.n)
        bd = np.nonzero(d.dirichlet)[0]
        if self.problem != "GaussianCharges" or self.bc == "Homogeneous" or not self.lammps:
            if self.bc == "Exact" and self.problem == "GaussianCharges":
                raise NotImplementedError
            return g
        x = d.real_coords()[bd]
        if self.bc == "Exact":
            val = np.zeros(len(bd))
            for k in range(len(self.charges)):
                r = np.sqrt(((x - self.pos[k]) ** 2).sum(1))
                safe = np.where(r < 1e-10, 1.0, r)
                val += np.where(r < 1e-10, self.charges[k] * 2.0 / (math.sqrt(math.pi) * self.r_c),
                                self.charges[k] * erf(safe / self.r_c) / safe)
            g[bd] = val
        else:  # Inhomogeneous: dipole term only, quadrupole is zeroed (src/step-50.cc:623-624)
            pm = (self.charges[:, None] * self.pos).sum(0)
            r = np.sqrt((x ** 2).sum(1))
            g[bd] = (x @ pm) / r ** 3
        return g

    def setup_system(self, cycle):
        self.dofs = DoFs(self.forest)
        if cycle == 0 and self.flag_rhs and self.lammps:
            self.lists0 = rhs.bin_atoms_base(self.forest, self.pos, self.cutoff * self.r_c,
                                             base_level=self.base_level())
        if self.lammps:
            self.lists = rhs.inherit_lists(self.forest, self.lists0, self.base_level()) if self.flag_rhs else None
            self.dens = rhs.charge_densities(self.forest, self.dofs, self.pos, self.charges, self.r_c, self.nq_rhs,
                                             self.lists)
        self.g = self.boundary_values()

    def assemble_system(self):
        coef = step16_coefficient if self.problem == "Step16" else None
        self.system = assemble.System(self.forest, self.dofs, coef)
        f, d = self.forest, self.dofs
        if not self.lammps:
            from . import fe
            pts, _ = fe.tensor_rule(self.nq_rhs, self.dim)
            self.dens = []
            for l in range(f.n_levels):
                org = f.lo + f.ijk[l][d.active_cells[l]] * f.h(l)
                xq = org[:, None, :] + f.h(l) * pts[None, :, :]
                self.dens.append(np.full(xq.shape[:2], 10.0) if self.problem == "Step16"
                                 else gaussian_rhs_function(xq, self.r_c))
        self.f_raw = assemble.load_vector(f, d, self.dens, self.nq_rhs)
        self.b = self.system.rhs(self.f_raw, self.g)

    def assemble_multigrid(self):
        coef = step16_coefficient if self.problem == "Step16" else None
        self.ops = assemble.LevelOps(self.forest, self.dofs, coef)

    def solve(self, x0):
        A, b = self.system.A, self.b
        rec = self.rec
        self.x0 = np.zeros(self.dofs.n) if x0 is None else x0.copy()
        rec["rhs_l1"], rec["rhs_l2"], rec["rhs_linf"] = float(np.abs(b).sum()), float(np.sqrt(b @ b)), float(np.abs(b).max())
        rec["mat_l1"] = float(abs(A).sum(0).max())
        rec["mat_linf"] = float(abs(A).sum(1).max())
        rec["mat_frob"] = float(math.sqrt((A.data ** 2).sum()))
        if self.precond == "GMG":
            blocks = None
            if self.ssor_ranks > 1:
                blocks = [np.minimum(np.arange(n) * self.ssor_ranks // max(n, 1), self.ssor_ranks - 1)
                          for n in self.dofs.level_n]
            self.mg = solver.Multigrid(self.ops, self.dofs, self.smoother, self.omega, self.steps, ssor_blocks=blocks)
            x, it, res0, res = solver.pcg_gmg(A, b, x0, self.mg)
            rec["coarse_its"] = list(self.mg.coarse_its)
        else:
            x, it, res0, res = solver.pcg_jacobi(A, b, x0)
        rec.update(start=res0, its=it, conv=res, sol_l1=float(np.abs(x).sum()), sol_l2=float(np.sqrt(x @ x)),
                   sol_linf=float(np.abs(x).max()))
        e = lambda v: "%.10e" % v
        for name, key in (("L1 rhs norm", "rhs_l1"), ("L2 rhs norm", "rhs_l2"), ("LInfinity rhs norm", "rhs_linf"),
                          ("L1 Matrix norm", "mat_l1"), ("LInfinity Matrix norm", "mat_linf"),
                          ("Frobenius Matrix norm", "mat_frob")):
            self.out(f"   {name} {e(rec[key])}")
        self.out("   Starting value %.10f" % res0)
        self.out(f"   CG converged in {it} iterations.")
        self.out(f"   Convergence value {e(res)}")
        for name, key in (("L1 solution norm", "sol_l1"), ("L2 solution norm", "sol_l2"),
                          ("LInfinity solution norm", "sol_linf")):
            self.out(f"   {name} {e(rec[key])}")
        self.x = x
        self.u = self.system.distribute(x, self.g)

    # -------------------------------------------------------------------- run
    def run(self, stop_after_solve_of_last_cycle=False, energy_gate=300):
        self.out("Running with Trilinos on 1 MPI rank(s)...")
        self.out(f"Dimension:\t{self.dim}")
        self.read_lammps_input_file()
        x0 = None
        for cycle in range(self.n_cycles):
            self.rec = rec = dict(cycle=cycle)
            self.out(f"Cycle {cycle}:")
            if cycle == 0:
                self.make_mesh()
            else:
                old_res, old_dofs, u_old = self.forest.resolution(), self.dofs, self.u
                self.forest.refine(self.flags)
                self.setup_system(cycle)
                x0 = estimate.transfer_solution(old_res, old_dofs, u_old, self.forest, self.dofs)
                if self.initial_guess == "zero":
                    x0 = None
            rec["n_active_cells"] = self.forest.n_active_cells()
            self.out(f"   Number of active cells:       {rec['n_active_cells']}")
            if cycle == 0:
                self.setup_system(cycle)
            rec["n_dofs"] = self.dofs.n
            rec["n_dofs_level"] = list(self.dofs.level_n)
            self.out(f"   Number of degrees of freedom: {self.dofs.n} (by level: "
                     + ", ".join(str(n) for n in self.dofs.level_n) + ")")
            self.assemble_system()
            if self.precond == "GMG":
                self.assemble_multigrid()
            self.solve(x0)
            self.records.append(rec)
            last = cycle == self.n_cycles - 1
            if last and stop_after_solve_of_last_cycle:
                break
            eta = estimate.kelly_plus_residual(self.forest, self.dofs, self.u, self.dens, self.nq_rhs,
                                               residual_term=self.indicator == "kelly+residual")
            rec["threshold"], self.flags = estimate.mark(self.forest, self.dofs, eta)
            self.eta = eta
            self.flag_history.append([f.copy() for f in self.flags])
            rec["n_flagged"] = int(sum(f.sum() for f in self.flags))
            self.out("Threshold value for refinement:\t%.10e" % rec["threshold"])
            if self.lammps and len(self.charges) < energy_gate:
                en = estimate.electrostatic_energy(self.forest, self.dofs, self.u, self.pos, self.charges, self.r_c)
                rec["energy"] = en
                self.out("\nTotal analytical electrostatic energy :   %.10e" % en["analytic"])
                self.out("Short-ranged energy contribution :  %.10e" % en["short"])
                self.out("FE solution long-ranged energy contribution :    %.10e" % en["fe"])
                self.out("Self energy contribution : %.10e" % en["self"])
                self.out("Total electrostatic energy with split in short- and long-ranged : %.10e" % en["total"])
                self.out("Absolute Error between both energies :\t%.10e\n" % abs(abs(en["analytic"]) - abs(en["total"])))
                rec["energy_norm_error"] = estimate.energy_norm_error(self.forest, self.dofs, self.u, self.pos,
                                                                      self.charges, self.r_c)
                self.out("Error in FE solution in energy norm:  %.10e" % rec["energy_norm_error"])
        return self.records


def from_prm_string(text, **kw):
    return LaplaceProblem(prm_mod.parse_string(text), **kw)


def from_prm_file(path, **kw):
    return LaplaceProblem(prm_mod.parse_file(path), base_dir=os.path.dirname(os.path.abspath(path)), **kw)
