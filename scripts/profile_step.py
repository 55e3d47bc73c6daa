"""One device-resident hot-path step of the bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (set-up cycles run unprofiled).  python scripts/profile_step.py [n] [smoother]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sm = sys.argv[2] if len(sys.argv) > 2 else "MulticolourSSOR"
class A: atoms_n = n
P = bench.pkg()
path, pos, q = bench.write_atoms(A)
B = P.hostapi.BenchProblem(P.lattice.cluster_prm(path, n, cycles=5, smoother=sm))
for _ in range(2):
    B.step_device()
B.gmg.synchronize()
torch.cuda.synchronize()
torch.cuda.profiler.start()
its, res = B.step_device()
B.gmg.synchronize()
torch.cuda.profiler.stop()
print("step ok: outer its", its, "res", res, "coarse", B.gmg.last_coarse_iterations())
