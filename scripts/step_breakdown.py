"""Developer probe: where one device-resident bench step spends its time (CUDA events, no profiler):
RHS, V-cycle down sweeps / coarse solves / up sweeps, the rest (system-matrix SpMVs, PCG vector kernels, host syncs).
python scripts/step_breakdown.py [n]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
class A: atoms_n = n
P = bench.pkg()
path, pos, q = bench.write_atoms(A)
B = P.hostapi.BenchProblem(P.lattice.cluster_prm(path, n, cycles=5))
g = B.gmg
for _ in range(3):
    B.step_device()
g.synchronize()
g.debug_vcycle_profile(True)
K = 5
t = time.time()
for _ in range(K):
    B.step_device()
g.synchronize()
wall = (time.time() - t) / K * 1e3
p = g.debug_vcycle_profile(False)
print("step %.2f ms wall; per step: V-cycle down %.2f ms, coarse CG %.2f ms, up %.2f ms (%d V-cycles per step); rest %.2f ms" % (
    wall, p["down_ms"] / K, p["coarse_ms"] / K, p["up_ms"] / K, p["vcycles"] // K,
    wall - (p["down_ms"] + p["coarse_ms"] + p["up_ms"]) / K))
