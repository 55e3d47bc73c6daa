"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2`): the row-partitioned solve over NVLink peer memory gives
the single-GPU result -- same outer iteration count, same coarse iteration counts, solution to 1e-9 relative."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, make_prm
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


@pytest.mark.parametrize("world", [2, 4])
def test_main_np_reproduces_single_gpu_stdout(tmp_path, world):
    """`main -np N file.prm` (the reference: `mpirun -np N main file.prm`, src/main.cc:8): LaplaceProblem::run() itself on
    N GPUs, one forked process per GPU, system matrix and level 0 row-partitioned.  Rank 0's stdout carries the numbers
    of the one-GPU run: mesh / DoF lines identical, norms of b and of the matrix identical (the same host assembly),
    iteration counts equal, solution norms to 1e-9 relative."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    extra = "subsection Misc\n set Refinement indicator = Kelly\nend\nsubsection Solver input data\n set Smoother = MulticolourSSOR\nend\n"
    prm = tmp_path / "gaussian-charges.prm"
    prm.write_text(make_prm(cycles=3, atom=os.path.join(GOLDEN, "atom_n1_8.data"), extra=extra))
    exe = os.path.join(os.path.dirname(pkg().capi.LIB_PATH), "main")
    one = subprocess.run([exe, str(prm)], capture_output=True, text=True, timeout=900)
    assert one.returncode == 0, one.stderr
    many = subprocess.run([exe, "-np", str(world), str(prm)], capture_output=True, text=True, timeout=900)
    assert many.returncode == 0, many.stderr
    assert f"on {world} GPU(s)" in many.stdout and "on 1 GPU(s)" in one.stdout
    assert many.stdout.count("Cycle ") == 3  # only rank 0 prints

    def numbers(text, label):
        return [float(x) for x in re.findall(label + r"\s+(\S+)", text)]
    for label in ("Number of active cells:", "L2 rhs norm", "Frobenius Matrix norm", "Starting value", "CG converged in"):
        assert numbers(one.stdout, label) == numbers(many.stdout, label), label
    for label in ("L1 solution norm", "L2 solution norm", "LInfinity solution norm"):
        a, b = numbers(one.stdout, label), numbers(many.stdout, label)
        assert len(a) == 3 and len(b) == 3
        assert all(abs(x - y) <= 1e-9 * abs(x) for x, y in zip(a, b)), (label, a, b)
