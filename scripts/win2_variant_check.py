"""Developer probe: the coarse-grid CG of the n-atom lattice's level-0 operator with two variants of the window kernel
(GMG_WIN2_VARIANT, read by gmg_create): same iteration count, same solution, time per inner iteration of each.
python scripts/win2_variant_check.py 20 0 3"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import scipy.sparse as sp
from helpers import pkg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
variants = [int(v) for v in sys.argv[2:]] or [0, 3]
M = pkg().hostapi.Ministep(2 * (2 * n + 20), -5.0, n + 5.0); M.build()
rp, col, val = M.csr("A", 0)
A = sp.csr_matrix((val, col, rp)); N = A.shape[0]
boundary = M.get("level_boundary", 0).astype(bool)
capi = pkg().capi
rng = np.random.default_rng(0); b = rng.standard_normal(N); b[boundary] = 0
out = {}
for v in variants:
    os.environ["GMG_WIN2_VARIANT"] = str(v)
    g = capi.Gmg(); g.set_num_levels(1)
    g.set_matrix(capi.GMG_SYSTEM, 0, A); g.set_matrix(capi.GMG_LEVEL, 0, A)
    g.set_copy_indices(0, np.arange(N), np.arange(N)); g.setup()
    kid = g.coarse_kernel(capi.GMG_LEVEL, 0)
    x, its, res = g.cg_solve(capi.GMG_LEVEL, 0, b, 1000, 1e-10)
    r = np.linalg.norm(A @ x - b)
    xd = g.vec_alloc(N); yd = g.vec_alloc(N); g.vec_upload(xd, b)
    best = 1e9
    for rep in range(3):
        g.coarse_profile(True)
        try: g.cg_solve_dev(capi.GMG_LEVEL, 0, xd, yd, 200, 1e-300)
        except capi.NoConvergence: pass
        g.synchronize(); p = g.coarse_profile(True)
        best = min(best, 1e3 * p["ms"] / max(p["iterations"], 1))
    out[v] = (x, its, res, r)
    print("variant %d (kernel id %d): rows %d, %d iterations, residual %.6e (true %.3e), %.2f us / inner iteration" % (
        v, kid, N, its, res, r, best), flush=True)
    g.close()
v0 = variants[0]
for v in variants[1:]:
    d = np.linalg.norm(out[v][0] - out[v0][0]) / np.linalg.norm(out[v0][0])
    print("variant %d vs %d: iterations %d / %d, rel. solution difference %.3e, residual ratio %.9f" % (
        v, v0, out[v][1], out[v0][1], d, out[v][2] / out[v0][2]))
    assert out[v][1] == out[v0][1] and d < 1e-9, "variants disagree"
print("ok")
