"""Developer probe: wall/device time of coarse-CG launches on the n-atom lattice level-0 operator for several
iteration limits (separates per-launch overhead from per-iteration cost).  python scripts/cg_launch_probe.py 20"""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from helpers import pkg
import scipy.sparse as sp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
M = pkg().hostapi.Ministep(2 * (2 * n + 20), -5.0, n + 5.0); M.build()
rp, col, val = M.csr("A", 0)
A = sp.csr_matrix((val, col, rp)); N = A.shape[0]
boundary = M.get("level_boundary", 0).astype(bool)
capi = pkg().capi
g = capi.Gmg(); g.set_num_levels(1)
g.set_matrix(capi.GMG_SYSTEM, 0, A); g.set_matrix(capi.GMG_LEVEL, 0, A)
g.set_copy_indices(0, np.arange(N), np.arange(N)); g.setup()
rng = np.random.default_rng(0); b = rng.standard_normal(N); b[boundary] = 0
x = g.vec_alloc(N); y = g.vec_alloc(N); g.vec_upload(x, b)
BLK = int(os.environ.get("PROBE_BLOCK", "74"))
for its in (1, 1, 200, 200):
    g.coarse_profile(True); g.debug_cg_phases(BLK + 1)
    t = time.time()
    try: g.cg_solve_dev(capi.GMG_LEVEL, 0, x, y, its, 1e-300)
    except capi.NoConvergence: pass
    g.synchronize(); wall = time.time() - t
    p = g.coarse_profile(True); ph = g.debug_cg_phases(0); blk = g.debug_cg_blocks()[:, :148] / max(ph[6], 1) / 1e3
    print("max_it %4d: events %.3f ms, iterations %d, block %d us/it: tail %.1f bar1 %.1f update %.1f bar2 %.1f dir %.1f bar3 %.1f | tiles: issue %.1f wait %.1f dom %.1f other %.1f sync %.1f rem %.1f" % (
        (its, p["ms"], p["iterations"], BLK) + tuple(ph[:6] / max(ph[6], 1) / 1e3) + tuple(ph[8:14] / max(ph[6], 1) / 1e3)))
    for name, row in zip(("spmv", "update", "direction"), blk):
        o = np.argsort(row)
        print("   %-9s per block us/it: min %.1f median %.1f max %.1f; slowest blocks %s (%s)" % (
            name, row.min(), np.median(row), row.max(), o[-5:][::-1], np.round(row[o[-5:][::-1]], 1)))
