"""ctypes binding of oracle/csrc/oracle.c (the plain-C restatement: CPU baseline + large-size checks).
TEST INFRASTRUCTURE -- see oracle/__init__.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None

_pd = C.POINTER(C.c_double)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "csrc", "oracle.c")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", HERE])
        _lib = C.CDLL(LIB)
        _lib.orc_bin_atoms.restype = C.c_int64
    return _lib


def max_threads():
    return lib().orc_max_threads()


def set_threads(n):
    """OpenMP threads of the port (bench.py: all cores of the affinity mask, whatever OMP_NUM_THREADS the launcher set)."""
    lib().orc_set_threads(int(n))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _csr(m):
    m = m.tocsr()
    return (np.ascontiguousarray(m.indptr, dtype=np.int64), np.ascontiguousarray(m.indices, dtype=np.int32), _f64(m.data))


def _ptr_array(arrays, ctype):
    arr = (C.c_void_p * len(arrays))()
    for i, a in enumerate(arrays):
        arr[i] = a.ctypes.data if a is not None else None
    return arr


class Hierarchy:
    """Keeps contiguous copies of the operators so repeated solves do not re-marshal them."""

    def __init__(self, system_A, level_A, level_I, prolong, copy_global, copy_level):
        self.nl = len(level_A)
        self.S = _csr(system_A) if not isinstance(system_A, tuple) else system_A
        conv = lambda m: m if isinstance(m, tuple) else _csr(m)
        self.A = [conv(m) for m in level_A]
        empty = (np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0))
        self.I = [empty if (l == 0 or level_I[l] is None) else conv(level_I[l]) for l in range(self.nl)]
        self.P = [conv(prolong[l]) if l + 1 < self.nl else empty for l in range(self.nl)]
        self.level_n = np.array([len(a[0]) - 1 for a in self.A], dtype=np.int32)
        self.cg = [np.ascontiguousarray(c, dtype=np.int32) for c in copy_global]
        self.cl = [np.ascontiguousarray(c, dtype=np.int32) for c in copy_level]
        self.n_copy = np.array([len(c) for c in self.cg], dtype=np.int32)
        self.n_sys = len(self.S[0]) - 1

    def pcg(self, b, x0, smoother="ssor", omega=0.5, steps=2, n_blocks=1, rel_tol=1e-8, max_it=500, coarse_tol=1e-10,
            coarse_max=1000):
        b = _f64(b)
        x = _f64(x0).copy() if x0 is not None else np.zeros(self.n_sys)
        tol = rel_tol * float(np.sqrt(b @ b))
        its, n_coarse = C.c_int(0), C.c_int(0)
        r0, r1 = C.c_double(0), C.c_double(0)
        coarse = np.zeros(4096, dtype=np.int32)
        pa = lambda lst, k: _ptr_array([t[k] for t in lst], None)
        rc = lib().orc_pcg_gmg(
            C.c_int(self.nl), C.c_int(self.n_sys), self.S[0].ctypes.data_as(_pi64), self.S[1].ctypes.data_as(_pi32),
            self.S[2].ctypes.data_as(_pd), self.level_n.ctypes.data_as(_pi32), pa(self.A, 0), pa(self.A, 1), pa(self.A, 2),
            pa(self.I, 0), pa(self.I, 1), pa(self.I, 2), pa(self.P, 0), pa(self.P, 1), pa(self.P, 2),
            self.n_copy.ctypes.data_as(_pi32), _ptr_array(self.cg, None), _ptr_array(self.cl, None),
            C.c_int({"jacobi": 0, "ssor": 1}[smoother]), C.c_double(omega), C.c_int(steps), C.c_int(n_blocks),
            b.ctypes.data_as(_pd), x.ctypes.data_as(_pd), C.c_double(tol), C.c_int(max_it), C.c_double(coarse_tol),
            C.c_int(coarse_max), C.byref(its), C.byref(r0), C.byref(r1), coarse.ctypes.data_as(_pi32), C.c_int(len(coarse)),
            C.byref(n_coarse))
        return dict(x=x, its=its.value, res0=r0.value, res=r1.value, rc=rc,
                    coarse_its=coarse[:min(n_coarse.value, len(coarse))].tolist())


def cg(A, b, tol=1e-10, max_it=1000):
    rp, col, val = A if isinstance(A, tuple) else _csr(A)
    b = _f64(b)
    x = np.zeros_like(b)
    its, res = C.c_int(0), C.c_double(0)
    rc = lib().orc_cg(C.c_int(len(b)), rp.ctypes.data_as(_pi64), col.ctypes.data_as(_pi32), val.ctypes.data_as(_pd),
                      b.ctypes.data_as(_pd), x.ctypes.data_as(_pd), C.c_double(tol), C.c_int(max_it), C.byref(its),
                      C.byref(res))
    return x, its.value, res.value, rc


def spmv(A, x):
    rp, col, val = A if isinstance(A, tuple) else _csr(A)
    x = _f64(x)
    y = np.zeros(len(rp) - 1)
    lib().orc_spmv(C.c_int(len(y)), rp.ctypes.data_as(_pi64), col.ctypes.data_as(_pi32), val.ctypes.data_as(_pd),
                   x.ctypes.data_as(_pd), y.ctypes.data_as(_pd))
    return y


def bin_atoms(reps, lo, H, pos, radius):
    pos = _f64(pos)
    nc = reps ** 3
    ptr = np.zeros(nc + 1, dtype=np.int64)
    args = [C.c_int(reps), C.c_double(lo), C.c_double(H), C.c_int(len(pos)), pos.ctypes.data_as(_pd), C.c_double(radius),
            ptr.ctypes.data_as(_pi64)]
    total = lib().orc_bin_atoms(*args, None)
    idx = np.zeros(max(total, 1), dtype=np.int32)
    lib().orc_bin_atoms(*args, idx.ctypes.data_as(_pi32))
    return ptr, idx[:total]


def density(cell_lo, cell_h, list_of_cell, list_ptr, list_atoms, pos, q, qpts, r_c):
    cell_lo, cell_h, pos, q, qpts = _f64(cell_lo), _f64(cell_h), _f64(pos), _f64(q), _f64(qpts)
    loc = np.ascontiguousarray(list_of_cell, dtype=np.int32)
    lp = np.ascontiguousarray(list_ptr if list_ptr is not None else [0], dtype=np.int64)
    la = np.ascontiguousarray(list_atoms if list_atoms is not None else [0], dtype=np.int32)
    rho = np.zeros((len(cell_h), len(qpts)))
    lib().orc_density(C.c_int(len(cell_h)), cell_lo.ctypes.data_as(_pd), cell_h.ctypes.data_as(_pd),
                      loc.ctypes.data_as(_pi32), lp.ctypes.data_as(_pi64), la.ctypes.data_as(_pi32), C.c_int(len(q)),
                      pos.ctypes.data_as(_pd), q.ctypes.data_as(_pd), C.c_int(len(qpts)), qpts.ctypes.data_as(_pd),
                      C.c_double(r_c), rho.ctypes.data_as(_pd))
    return rho


def load_vector(rho, cell_h, cell_dofs, shape, weights, n_dofs, hang_ptr, hang_col, hang_val, constrained, kref=None,
                ghat=None):
    rho, cell_h, shape, weights = _f64(rho), _f64(cell_h), _f64(shape), _f64(weights)
    cd = np.ascontiguousarray(cell_dofs, dtype=np.int32)
    hp = np.ascontiguousarray(hang_ptr, dtype=np.int64)
    hc = np.ascontiguousarray(hang_col if len(hang_col) else [0], dtype=np.int32)
    hv = _f64(hang_val if len(hang_val) else [0.0])
    cons = np.ascontiguousarray(constrained, dtype=np.uint8)
    b = np.zeros(n_dofs)
    k = _f64(kref) if kref is not None else None
    g = _f64(ghat) if ghat is not None else None
    lib().orc_load_vector(C.c_int(len(cell_h)), rho.ctypes.data_as(_pd), cell_h.ctypes.data_as(_pd),
                          cd.ctypes.data_as(_pi32), C.c_int(len(weights)), shape.ctypes.data_as(_pd),
                          weights.ctypes.data_as(_pd), k.ctypes.data_as(_pd) if k is not None else None,
                          g.ctypes.data_as(_pd) if g is not None else None, C.c_int(n_dofs), hp.ctypes.data_as(_pi64),
                          hc.ctypes.data_as(_pi32), hv.ctypes.data_as(_pd), cons.ctypes.data_as(C.POINTER(C.c_uint8)),
                          b.ctypes.data_as(_pd))
    return b


def pair_energies(pos, q, r_c):
    """(analytic, short-ranged) pair sums of postprocess_electrostatic_energy (src/step-50.cc:1315-1345)."""
    pos, q = _f64(pos), _f64(q)
    out = np.zeros(2)
    lib().orc_pair_energies(C.c_int(len(q)), pos.ctypes.data_as(_pd), q.ctypes.data_as(_pd), C.c_double(r_c),
                            out.ctypes.data_as(_pd))
    return float(out[0]), float(out[1])
