"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2`): the row-partitioned solve over NVLink peer memory gives
the single-GPU result -- same outer iteration count, same coarse iteration counts, solution to 1e-9 relative."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import pkg

pytestmark = pytest.mark.gpu


def _prm(P, device, smoother):
    return P.lattice.cluster_prm(os.path.join(GOLDEN, "atom_n1_8.data"), 1, cycles=3, smoother=smoother, device=device)


def _worker(rank, world, port, smoother, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    P = pkg()

    def connect(gmg):
        def gather(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        P.capi.connect_ranks(gmg, rank, world, gather, comm_bytes=64 << 20)
        dist.barrier()

    B = P.hostapi.BenchProblem(_prm(P, rank, smoother), connect=connect)
    out = []
    for _ in range(2):
        dist.barrier()
        its, res = B.step_device()
        out.append((its, res, B.gmg.last_coarse_iterations()))
    x = B.download_x()
    if rank == 0:
        q.put((out, x))
    dist.barrier()
    B.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,smoother", [(2, "Jacobi"), (2, "MulticolourSSOR")])
def test_partitioned_solve_matches_single_gpu(world, smoother):
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    P = pkg()
    B = P.hostapi.BenchProblem(_prm(P, 0, smoother))
    its1, res1 = B.step_device()
    coarse1 = B.gmg.last_coarse_iterations()
    x1 = B.download_x()
    B.close()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, smoother, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, x = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for its, res, coarse in out:
        assert its == its1
        assert coarse == coarse1
        assert abs(res - res1) <= 1e-6 * res1
    assert np.linalg.norm(x - x1) <= 1e-9 * np.linalg.norm(x1)
