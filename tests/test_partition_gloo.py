"""CPU tests of the N > 1 host logic: the row partitioner of the CUDA library (csrc/partition.cc) drives a
2-process (gloo) emulation of the distributed SpMV / halo exchange / all-reduce and must reproduce the global
result.  (The device-side communication itself needs >= 2 GPUs: tests/test_gpu_multi.py.)"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import pkg


def slab_owner(n_side, world):
    k = np.arange(n_side ** 3) // (n_side * n_side)
    return np.minimum(k * world // n_side, world - 1).astype(np.int32)


def laplace27(n):
    e = np.ones(n)
    T = sp.diags([e[:-1], 2 * e, e[:-1]], [-1, 0, 1])
    M = sp.diags([e[:-1] / 6, 4 * e / 6, e[:-1] / 6], [-1, 0, 1])
    return (sp.kron(sp.kron(M, M), T) + sp.kron(sp.kron(M, T), M) + sp.kron(sp.kron(T, M), M)).tocsr()


def test_partition_maps_cover_and_match_single_process():
    capi = pkg().capi
    A = laplace27(9)
    for world in (2, 3, 4):
        owner = slab_owner(9, world)
        parts = [capi.partition_probe(r, world, A, owner) for r in range(world)]
        assert sum(p["n_owned"] for p in parts) == A.shape[0]
        x = np.random.default_rng(0).standard_normal(A.shape[0])
        # extended vectors: [lower-rank halo | owned | higher-rank halo]
        ext = [np.concatenate([np.zeros(p["n_halo_lo"]), x[p["owned_global"]], np.zeros(p["n_halo"] - p["n_halo_lo"])])
               for p in parts]
        # emulate the pushes: r -> q
        for r, p in enumerate(parts):
            off = 0
            for q in range(world):
                c = p["send_count"][q]
                idx = p["send_idx"][off:off + c]
                off += c
                ext[q][p["send_dst_base"][q] + np.arange(c)] = ext[r][p["n_halo_lo"] + idx]
        y = np.zeros(A.shape[0])
        for p, xe in zip(parts, ext):
            lo = p["n_halo_lo"]
            assert np.array_equal(np.concatenate([xe[:lo], xe[lo + p["n_owned"]:]]), x[p["halo_global"]])
            L = sp.csr_matrix((p["val"], p["col"] + lo, p["rowptr"]), shape=(p["n_owned"], p["n_owned"] + p["n_halo"]))
            y[p["owned_global"]] = L @ xe
        assert np.abs(y - A @ x).max() < 1e-13


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    capi = pkg().capi
    A = laplace27(8)
    owner = slab_owner(8, world)
    p = capi.partition_probe(rank, world, A, owner)
    lo, no = p["n_halo_lo"], p["n_owned"]
    L = sp.csr_matrix((p["val"], p["col"] + lo, p["rowptr"]), shape=(no, no + p["n_halo"]))
    rng = np.random.default_rng(1)
    b = rng.standard_normal(A.shape[0])
    # distributed CG (deal.II recurrences), halos by send/recv, dots by all_reduce
    x = np.zeros(p["n_owned"])
    g = -b[p["owned_global"]]
    d = np.concatenate([np.zeros(lo), -g, np.zeros(p["n_halo"] - lo)])  # [lower halo | owned | upper halo]

    def halo(v):
        off, reqs, bufs = 0, [], {}
        for qq in range(world):
            c = int(p["send_count"][qq])
            if c:
                reqs.append(dist.isend(torch.from_numpy(v[lo + p["send_idx"][off:off + c]].copy()), qq))
            off += c
        other = [parts_n for parts_n in range(world) if parts_n != rank]
        k = 0  # position in the halo list (ordered by owner)
        for qq in range(world):
            cnt = int((owner[p["halo_global"]] == qq).sum())
            if cnt:
                t = torch.zeros(cnt, dtype=torch.float64)
                dist.recv(t, qq)
                pos = k if qq < rank else no + k
                v[pos:pos + cnt] = t.numpy()
                k += cnt
        for r in reqs:
            r.wait()

    def allsum(s):
        t = torch.tensor([s], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t.item())

    gh = allsum(g @ g)
    its = 0
    while np.sqrt(gh) > 1e-10 and its < 500:
        its += 1
        halo(d)
        h = L @ d
        alpha = gh / allsum(d[lo:lo + no] @ h)
        x += alpha * d[lo:lo + no]
        g += alpha * h
        new = allsum(g @ g)
        d[lo:lo + no] = (new / gh) * d[lo:lo + no] - g
        gh = new
    full = np.zeros(A.shape[0])
    full[p["owned_global"]] = x
    t = torch.from_numpy(full)
    dist.all_reduce(t)
    if rank == 0:
        q.put((its, float(np.abs(A @ t.numpy() - b).max())))
    dist.destroy_process_group()


def test_two_process_distributed_cg_over_gloo():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    its, err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert its < 100 and err < 1e-9
