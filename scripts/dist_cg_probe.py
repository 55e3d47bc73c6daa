"""Developer probe (multi-GPU): per-phase timers of the distributed coarse CG on rank 0.
torchrun --nproc-per-node N scripts/dist_cg_probe.py [n] [block]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 74
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
class A: atoms_n = n
P = bench.pkg()
path, pos, q = bench.write_atoms(A)
def connect(gmg):
    def gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    P.capi.connect_ranks(gmg, rank, world, gather)
    dist.barrier()
B = P.hostapi.BenchProblem(P.lattice.cluster_prm(path, n, cycles=5, device=local), connect=connect)
for _ in range(3):
    B.step_device()
B.gmg.synchronize(); dist.barrier()
B.gmg.debug_cg_phases(blk + 1)
K = 3
for _ in range(K):
    B.step_device()
B.gmg.synchronize()
ph = B.gmg.debug_cg_phases(0)
if rank == 0:
    c = max(ph[6], 1)
    print("N=%d block %d, %d iterations: spmv %.1f  allsum1 %.1f  update %.1f  allsum2 %.1f  dir %.1f  halo %.1f  barrier %.1f us/it; total %.1f" % (
        (world, blk, int(ph[6])) + tuple(ph[i] / c / 1e3 for i in (0, 1, 2, 3, 4, 7, 5)) + (sum(ph[i] for i in (0, 1, 2, 3, 4, 5, 7)) / c / 1e3,)))
dist.barrier()
dist.destroy_process_group()
