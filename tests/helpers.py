"""Shared test plumbing: run the CPU oracle up to a refinement cycle, hand its assembled operators to
the CUDA library through the C ABI, and compare.  (The oracle is the checker, never the product.)"""
import functools
import importlib
import math

import numpy as np

PKG = "geometric-multigrid-preconditioners-for-long-range-coulomb-interaction_b200"


def pkg():
    return importlib.import_module(PKG)


@functools.lru_cache(maxsize=8)
def oracle_cycle(prm_text, cycle, smoother="ssor"):
    """Oracle LaplaceProblem advanced to `cycle` (0-based) with that cycle solved."""
    from oracle import prm as prm_mod, problem
    params = prm_mod.parse_string(prm_text)
    params[("Misc", "Number of Adaptive Refinement")] = cycle + 1
    P = problem.LaplaceProblem(params, smoother=smoother)
    P.run(stop_after_solve_of_last_cycle=True)
    return P


def hand_over(P, gmg, smoother="jacobi", omega=0.5, steps=2):
    capi = pkg().capi
    kind = {"jacobi": capi.SMOOTHER_JACOBI, "chebyshev": capi.SMOOTHER_CHEBYSHEV, "mc_ssor": capi.SMOOTHER_MC_SSOR,
            "lex_ssor": capi.SMOOTHER_LEX_SSOR}[smoother]
    capi.hand_over_hierarchy(gmg, P.system.A_stored, P.ops.A_stored, P.ops.I, P.ops.P, P.dofs.copy_global,
                             P.dofs.copy_level)
    gmg.set_smoother(kind, omega, steps)
    gmg.set_coarse(1000, 1e-10)
    gmg.setup()


def device_solve_cycle(P, smoother="jacobi", device=0):
    """Solve the oracle's current cycle on the GPU (same matrices, rhs and initial guess) and compare."""
    from oracle import solver
    capi = pkg().capi
    gmg = capi.Gmg(device)
    try:
        hand_over(P, gmg, smoother)
        tol = 1e-8 * math.sqrt(P.b @ P.b)
        x, its, res0, res = gmg.pcg_solve(P.b, P.x0, 500, tol)
        coarse = gmg.last_coarse_iterations()
    finally:
        gmg.close()
    o_sm = {"jacobi": "jacobi", "lex_ssor": "ssor"}.get(smoother)
    out = dict(its=its, res0=res0, res=res, x=x, coarse_its=coarse)
    if o_sm is not None:
        mg = solver.Multigrid(P.ops, P.dofs, o_sm, 0.5, 2)
        xo, ito, r0o, ro = solver.pcg_gmg(P.system.A, P.b, P.x0, mg)
        out.update(oracle_its=ito, oracle_res0=r0o, oracle_res=ro, oracle_coarse_its=list(mg.coarse_its),
                   rel_solution_error=float(np.linalg.norm(x - xo) / np.linalg.norm(xo)), x_oracle=xo)
    return out


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
