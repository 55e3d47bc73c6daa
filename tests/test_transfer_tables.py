"""CPU check of the index tables behind the device solution transfer (gmg_transfer_solution, SURVEY 8f N1;
src/step-50.cc:1110-1119): a numpy restatement of what the kernels do with the tables (copy; pass by pass: claim by the
lowest refined cell, trilinear fill with unfused sums in vertex order; constrained dofs to zero) must give ministep's
transfer_solution bit for bit.  The GPU test (tests/test_gpu_indicator.py) runs the kernels themselves."""
import numpy as np

from conftest import make_prm
from helpers import oracle_cycle
import hostlib


def emulate(u_old, n_new, copy_old, copy_new, pass_ptr, parent_dofs, constrained):
    x = np.zeros(n_new)
    stamp = np.full(n_new, np.iinfo(np.int32).max, dtype=np.int64)
    x[copy_new] = u_old[copy_old]
    stamp[copy_new] = -1
    corner = [2 * (v & 1) + 6 * ((v >> 1) & 1) + 18 * ((v >> 2) & 1) for v in range(8)]
    for l in range(len(pass_ptr) - 1):
        owner = {}
        cells = range(int(pass_ptr[l]), int(pass_ptr[l + 1]))
        ready = []
        for p in cells:
            q = parent_dofs[p]
            c = q[corner]
            if (c < 0).any() or (stamp[c] >= l).any():
                continue
            ready.append(p)
            for t in range(27):
                tt = (t % 3, (t // 3) % 3, t // 9)
                if 1 not in tt:
                    continue
                d = q[t]
                if d >= 0 and stamp[d] >= l:
                    owner[d] = min(owner.get(d, p), p)
        for p in ready:
            q = parent_dofs[p]
            U = x[q[corner]]
            for t in range(27):
                tt = (t % 3, (t // 3) % 3, t // 9)
                d = q[t]
                if 1 not in tt or d < 0 or owner.get(d) != p:
                    continue
                val = 0.0
                for v in range(8):
                    w = 1.0
                    for k in range(3):
                        w *= tt[k] / 2.0 if (v >> k) & 1 else 1.0 - tt[k] / 2.0
                    if w != 0.0:
                        val = val + w * U[v]
                x[d] = val
        for d in owner:
            stamp[d] = l
    assert (stamp != np.iinfo(np.int32).max).all()
    x[np.asarray(constrained, dtype=bool)] = 0.0
    return x


def test_tables_reproduce_sequential_transfer_bit_for_bit():
    P = oracle_cycle(make_prm(cycles=4, bc="Exact", atom="atom_n1_2.data", nq=4), 3)
    f = P.forest
    old = hostlib.Ministep(f.reps, f.lo, f.hi)
    old.build(matrices=False)
    rng = np.random.default_rng(5)
    history = []
    for flags in P.flag_history[:3]:
        new = hostlib.Ministep(f.reps, f.lo, f.hi)
        for fl in history + [flags]:
            new.refine(fl)
        new.build(matrices=False)
        history.append(flags)
        u_old = rng.standard_normal(len(old.get("boundary")))
        old_res = old.n_levels - 1
        ref = new.transfer_from(old, old_res, u_old)
        co, cn, pp, pd = new.transfer_tables(old, old_res)
        assert len(co) == len(u_old)  # refinement only: every old dof survives
        assert len(pp) == new.n_levels and pp[-1] == len(pd) > 0
        x = emulate(u_old, len(ref), co, cn, pp, pd, new.get("constrained"))
        assert np.array_equal(x.view(np.uint64), ref.view(np.uint64))
        old = new
