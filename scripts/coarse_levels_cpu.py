"""Developer probe (CPU only, no GPU needed): what do coarse levels below the base mesh (prm key
`Coarse levels below the base mesh`) do to the iteration counts?  Runs the adaptive loop with the oracle's C port
(processor-block SSOR with one block per thread, like the reference's MPI ranks) on ministep's hierarchies.

    python scripts/coarse_levels_cpu.py [atoms_n=20] [cycles=5] [k,k,...=0,1,2,3]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from oracle import cpu_arm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ks = [int(t) for t in (sys.argv[3] if len(sys.argv) > 3 else "0,1,2,3").split(",")]
P = bench.pkg()
pos, q = P.lattice.nacl_lattice(n)
out = {}
for k in ks:
    t0 = time.time()
    _, _, recs = cpu_arm.adaptive_run_on_cpu(P.hostapi, pos, q, n, cycles, coarse_levels=k, solve_last=True,
                                             log=lambda s: print(s, file=sys.stderr, flush=True))
    out[k] = [{"its": r["its"], "coarse_its": r["coarse_its"], "solve_seconds": round(r["solve_seconds"], 3),
               "res": r["res"]} for r in recs]
    print(f"k={k}: outer its {[r['its'] for r in recs]}, coarse its per cycle {[sum(r['coarse_its']) for r in recs]}, "
          f"solve s {[round(r['solve_seconds'], 2) for r in recs]}  ({time.time() - t0:.0f} s)", flush=True)
print(json.dumps(out))
