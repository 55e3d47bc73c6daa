"""CPU arm of the benchmark: the hot path (Gaussian-charge RHS + MG-PCG solve) executed by the oracle's
plain-C restatement (oracle/csrc/oracle.c) on the host cores.  Mesh, DoF numbering and matrix assembly
(host-side inputs of the path, not the thing measured) come from the host library's `ministep`.

TEST INFRASTRUCTURE -- only bench.py's `cpu_baseline` / `--impl reference` legs and tests use this.
"""
import time

import numpy as np

from . import cport, fe


def cell_arrays(get, n_levels, lo, H, base_level=0, reps_base=None):
    """Flatten active cells (level by level): lower corners, edge lengths, dofs, base-lattice ancestor (= atom list id).
    base_level > 0 (`Coarse levels below the base mesh`): the base lattice is level `base_level` of the forest; its cells
    are mapped to the lexicographic (x fastest) index the atom lists use."""
    los, hs, dofs, base = [], [], [], []
    parents = [get("parent", l) for l in range(n_levels)]
    for l in range(n_levels):
        act = get("active_cells", l)
        ijk = get("ijk", l).reshape(-1, 3)[act]
        h = H / (1 << l)
        los.append(lo + ijk * h)
        hs.append(np.full(len(act), h))
        dofs.append(get("cell_dofs", l).reshape(-1, 8))
        anc = act.copy()
        if l < base_level:
            assert len(act) == 0
        for k in range(l, base_level, -1):
            anc = parents[k][anc]
        if base_level > 0 and len(anc):
            ijk_b = get("ijk", base_level).reshape(-1, 3)[anc].astype(np.int64)
            anc = (ijk_b[:, 2] * reps_base + ijk_b[:, 1]) * reps_base + ijk_b[:, 0]
        base.append(anc)
    return np.concatenate(los), np.concatenate(hs), np.concatenate(dofs), np.concatenate(base).astype(np.int32)


def hierarchy_from(get, n_levels):
    csr = lambda pre, l=0: (get(pre + "_rowptr", l), get(pre + "_col", l), get(pre + "_val", l))
    A = [csr("A", l) for l in range(n_levels)]
    I = [None] + [csr("I", l) for l in range(1, n_levels)]
    P = [csr("P", l) for l in range(n_levels - 1)]
    return cport.Hierarchy(csr("sys"), A, I, P, [get("copy_global", l) for l in range(n_levels)],
                           [get("copy_level", l) for l in range(n_levels)])


class CpuStep:
    """One hot-path step on the CPU for a given mesh state: densities + load vector + MG-PCG."""

    def __init__(self, get, n_levels, lo, H, pos, charges, lists, nq, r_c=0.5, smoother="ssor", n_blocks=None,
                 base_level=0, reps_base=None):
        self.H = hierarchy_from(get, n_levels)
        self.cell_lo, self.cell_h, self.dofs, self.base = cell_arrays(get, n_levels, lo, H, base_level, reps_base)
        self.pts, self.wts = fe.tensor_rule(nq, 3)
        self.shape = fe.shape_values(self.pts, 3)
        self.pos, self.charges, self.lists = pos, charges, lists
        self.hang = (get("hang_rowptr"), get("hang_col"), get("hang_val"))
        self.constrained = get("constrained")
        self.n_dofs = len(self.constrained)
        self.r_c, self.smoother = r_c, smoother
        self.n_blocks = n_blocks or cport.max_threads()

    def rhs(self):
        self.rho = cport.density(self.cell_lo, self.cell_h, self.base, self.lists[0], self.lists[1], self.pos,
                                 self.charges, self.pts, self.r_c)
        return cport.load_vector(self.rho, self.cell_h, self.dofs, self.shape, self.wts, self.n_dofs, *self.hang,
                                 self.constrained)

    def run(self, x0):
        t0 = time.perf_counter()
        b = self.rhs()
        t1 = time.perf_counter()
        r = self.H.pcg(b, x0, smoother=self.smoother, n_blocks=self.n_blocks)
        t2 = time.perf_counter()
        r.update(b=b, rhs_seconds=t1 - t0, solve_seconds=t2 - t1, seconds=t2 - t0)
        return r


def adaptive_run_on_cpu(hostapi, pos, charges, n, cycles, nq=2, r_c=0.5, cutoff=3.5, mesh_size=0.25, vacuum=10,
                        kelly_only=True, smoother="ssor", n_blocks=None, log=None, coarse_levels=0, solve_last=False):
    """The reference's run() loop on the CPU up to (not including) the solve of the last cycle: returns the CpuStep
    of the last cycle and its transferred initial guess.  Homogeneous Dirichlet values (the lattices are neutral
    with zero dipole)."""
    a = 2 * mesh_size
    reps = int(2 * (n / a + 2 * vacuum))
    lo, hi = 0.0 - vacuum * a, n + vacuum * a
    k = coarse_levels  # `Coarse levels below the base mesh`: the base lattice is level k of a 2^k times coarser forest
    assert reps % (1 << k) == 0

    def fresh():
        F = hostapi.Ministep(reps >> k, lo, hi)
        if k:
            F.refine_global(k)
        return F
    M = fresh()
    H0 = (hi - lo) / reps
    lists = cport.bin_atoms(reps, lo, H0, pos, cutoff * r_c)
    records = []
    x0 = None
    step = None
    for c in range(cycles):
        M.build()
        step = CpuStep(M.get, M.n_levels, lo, M.h0, pos, charges, lists, nq, r_c, smoother, n_blocks, k, reps)
        if x0 is None:
            x0 = np.zeros(step.n_dofs)
        if log:
            log(f"cpu arm: cycle {c}: {step.n_dofs} dofs, levels {M.get('level_n').tolist()}")
        if c == cycles - 1 and not solve_last:
            return step, x0
        r = step.run(x0)
        records.append({key: r[key] for key in r if key not in ("x", "b")})
        if c == cycles - 1:
            return step, x0, records
        u = M.distribute(np.zeros(step.n_dofs), r["x"])
        M.error_indicator(u, step.rho, nq, residual_term=not kelly_only)
        flags = [M.get("flags", l) for l in range(M.n_levels)]
        old, old_res = M, M.n_levels - 1
        M = fresh()
        # replay the refinement history on a fresh forest, then add this cycle's flags
        for fl in getattr(old, "_history", []):
            M.refine(fl)
        M.refine(flags)
        M._history = getattr(old, "_history", []) + [flags]
        M.build(matrices=False)
        x0 = M.transfer_from(old, old_res, u)
    return step, x0
