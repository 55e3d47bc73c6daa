"""Developer probe (2 GPUs): NVLink round-trip latency of the LL words the multi-GPU kernels use.
torchrun --nproc-per-node 2 scripts/p2p_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")
P = bench.pkg()
g = P.capi.Gmg(rank)
def gather(b):
    out = [None] * world
    dist.all_gather_object(out, b)
    return out
P.capi.connect_ranks(g, rank, world, gather, comm_bytes=1 << 20)
dist.barrier()
for mode in (0, 1, 0):
    dist.barrier()
    us = g.pingpong(20000, mode)
    if rank == 0:
        print("LL ping-pong mode %d: %.2f us per round trip (%.2f us one way)" % (mode, us, us / 2), flush=True)
dist.barrier()
