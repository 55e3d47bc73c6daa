"""Developer probe (not the bench): time level-0 SpMV and the persistent coarse CG on an n-atom lattice
level-0 operator assembled by the host library (ministep).  python scripts/perf_probe.py 20"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from helpers import pkg
import scipy.sparse as sp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = 2 * (2 * n + 20)
t = time.time()
M = pkg().hostapi.Ministep(reps, -5.0, n + 5.0); M.build()
rp, col, val = M.csr("A", 0)
A = sp.csr_matrix((val, col, rp))
boundary = M.get("level_boundary", 0).astype(bool)
print("assembled", A.shape, A.nnz, "in %.1fs" % (time.time() - t), flush=True)
capi = pkg().capi
g = capi.Gmg()
g.set_num_levels(1)
g.set_matrix(capi.GMG_SYSTEM, 0, A)
g.set_matrix(capi.GMG_LEVEL, 0, A)
g.set_copy_indices(0, np.arange(A.shape[0]), np.arange(A.shape[0]))
N = A.shape[0]
rng = np.random.default_rng(0)
b = rng.standard_normal(N); b[boundary] = 0
xs = {}
MODES = [int(c) for c in os.environ.get('PROBE_MODES', '012')]
for comp in MODES:
    g.set_compression(comp)
    t = time.time(); g.setup(); print("setup %.2fs" % (time.time() - t))
    tr = g.matrix_traffic(capi.GMG_LEVEL, 0)
    print("compressed", comp, tr)
    x = g.vec_alloc(N); y = g.vec_alloc(N)
    g.vec_upload(x, rng.standard_normal(N))
    for _ in range(5): g.spmv_dev(capi.GMG_LEVEL, 0, x, y)
    g.synchronize(); t = time.time()
    R = 50
    for _ in range(R): g.spmv_dev(capi.GMG_LEVEL, 0, x, y)
    g.synchronize(); dt = (time.time() - t) / R
    print("spmv %.3f ms  %.0f GB/s CSR-equivalent" % (dt * 1e3, tr["csr_spmv_bytes"] / dt / 1e9))
    g.vec_upload(x, b)
    g.coarse_profile(True)
    g.debug_cg_phases(1)
    for rep in range(2):
        try:
            g.cg_solve_dev(capi.GMG_LEVEL, 0, x, y, 200, 1e-300)
        except capi.NoConvergence:
            pass
    p = g.coarse_profile(True)
    per = p["ms"] / p["iterations"]
    print("cg: %d its in %.2f ms -> %.4f ms/it, %.0f GB/s of the stored format, %.0f GB/s CSR-equivalent" % (
        p["iterations"], p["ms"], per, tr["cg_iter_bytes"] / per / 1e6, tr["csr_cg_iter_bytes"] / per / 1e6))
    ph = g.debug_cg_phases(0)
    if ph[6] > 0:
        print("phases (us/it, block 0): spmv %.1f  bar1 %.1f  update %.1f  bar2 %.1f  direction %.1f  bar3 %.1f" % tuple(ph[:6] / ph[6] / 1e3))
    xs[comp] = g.vec_download(y, N)
    g.vec_free(x); g.vec_free(y)
for m in MODES[1:]:
    print("mode", m, "vs", MODES[0], "bit-identical:", np.array_equal(xs[MODES[0]], xs[m]), "rel", np.linalg.norm(xs[m] - xs[MODES[0]]) / np.linalg.norm(xs[MODES[0]]))
