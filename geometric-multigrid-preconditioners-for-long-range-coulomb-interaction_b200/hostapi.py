"""ctypes binding of the host library (libstep50_b200.so, host/capi_host.cc): ministep pieces (host only),
LaplaceProblem runs and the bench hooks (need a B200).  Plumbing for tests and bench.py."""
import ctypes as C
import json
import os

import numpy as np

from . import capi

_DT = {0: np.int32, 1: np.int64, 2: np.float64, 3: np.uint8, 4: np.float32}
_lib = None
_ms_lib = None


def _declare_ms(L):
    L.ms_create.restype = C.c_void_p
    L.ms_create.argtypes = [C.c_int, C.c_double, C.c_double]
    L.ms_destroy.argtypes = [C.c_void_p]
    L.ms_refine.argtypes = [C.c_void_p, C.c_void_p]
    L.ms_refine_global.argtypes = [C.c_void_p, C.c_int]
    L.ms_build.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ms_n_levels.argtypes = [C.c_void_p]
    L.ms_n_cells.restype = C.c_int64
    L.ms_n_cells.argtypes = [C.c_void_p, C.c_int]
    L.ms_get.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L.ms_assemble_emulate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    L.ms_unit_stiffness.argtypes = [C.c_void_p]
    L.ms_error_indicator.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.ms_transfer.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ms_transfer_tables.argtypes = [C.c_void_p, C.c_int, C.c_void_p] + [C.c_void_p] * 6
    L.ms_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ms_locate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
    L.ms_gauss.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    L.ms_last_error.restype = C.c_char_p


def ms_lib():
    """libministep_b200.so: ministep alone (mesh, numbering, host assembly, indicator, transfer).  Links no CUDA
    library, so the CPU arm of bench.py and the CPU tests never map the product's kernels."""
    global _ms_lib
    if _ms_lib is None:
        path = os.path.join(os.path.dirname(capi.LIB_PATH), "libministep_b200.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run __graft_entry__.build()")
        L = C.CDLL(path)
        _declare_ms(L)
        _ms_lib = L
    return _ms_lib


def lib():
    """libstep50_b200.so: LaplaceProblem + bench hooks (links libgmg_b200.so, needs a B200 to run)."""
    global _lib
    if _lib is None:
        path = os.path.join(os.path.dirname(capi.LIB_PATH), "libstep50_b200.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run __graft_entry__.build()")
        L = C.CDLL(path)
        _declare_ms(L)
        L.step50_check_prm.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.step50_run_string.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.step50_free.argtypes = [C.c_void_p]
        L.step50_bench_create.restype = C.c_void_p
        L.step50_bench_create.argtypes = [C.c_char_p]
        L.step50_bench_destroy.argtypes = [C.c_void_p]
        L.step50_bench_finish_setup.argtypes = [C.c_void_p]
        L.step50_bench_set_device_assembly.argtypes = [C.c_void_p, C.c_int]
        L.step50_bench_download_x.argtypes = [C.c_void_p, C.c_void_p]
        L.step50_bench_download_b.argtypes = [C.c_void_p, C.c_void_p]
        L.step50_bench_time_binning.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.step50_bench_gmg.restype = C.c_void_p
        L.step50_bench_gmg.argtypes = [C.c_void_p]
        L.step50_bench_info.argtypes = [C.c_void_p, C.c_void_p]
        L.step50_bench_step_device.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.step50_bench_step_host.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.step50_bench_get.argtypes = L.ms_get.argtypes
        L.step50_bench_mesh.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        _lib = L
    return _lib


class HostError(RuntimeError):
    pass


def _ck(rc):
    if rc != 0:
        raise HostError(lib().ms_last_error().decode())


def _ck_ms(rc):
    if rc != 0:
        raise HostError(ms_lib().ms_last_error().decode())


def _fetch(fn, handle, name, level, ck=None):
    ptr, n, dt = C.c_void_p(), C.c_int64(), C.c_int()
    (ck or _ck)(fn(handle, name.encode(), level, C.byref(ptr), C.byref(n), C.byref(dt)))
    dtype = np.dtype(_DT[dt.value])
    if n.value == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n.value * dtype.itemsize)).from_address(ptr.value)
    return np.frombuffer(buf, dtype=dtype).copy()


class Ministep:
    def __init__(self, reps, lo, hi):
        self.L = ms_lib()
        self.p = self.L.ms_create(reps, lo, hi)
        self.h0 = (hi - lo) / reps  # edge of a base cell (Forest::H)

    def __del__(self):
        if getattr(self, "p", None):
            self.L.ms_destroy(self.p)
            self.p = None

    def refine(self, flags_per_level):
        flat = np.ascontiguousarray(np.concatenate([np.asarray(f, dtype=np.uint8) for f in flags_per_level]))
        _ck_ms(self.L.ms_refine(self.p, flat.ctypes.data))

    def refine_global(self, times):
        _ck_ms(self.L.ms_refine_global(self.p, times))

    def build(self, step16=False, matrices=True):
        _ck_ms(self.L.ms_build(self.p, int(step16), int(matrices)))

    @property
    def n_levels(self):
        return self.L.ms_n_levels(self.p)

    def n_cells(self, l):
        return self.L.ms_n_cells(self.p, l)

    def get(self, name, level=0):
        return _fetch(self.L.ms_get, self.p, name, level, _ck_ms)

    def csr(self, prefix, level=0):
        import scipy.sparse as sp
        rp, col, val = self.get(prefix + "_rowptr", level), self.get(prefix + "_col", level), self.get(prefix + "_val", level)
        return rp, col, val

    def assemble_emulate(self, which, level=0):
        """Sequential emulation of the device-side assembly against the host assembly: (differing words, longest row)."""
        nd, mr = C.c_int64(-1), C.c_int(0)
        _ck_ms(self.L.ms_assemble_emulate(self.p, which, level, C.byref(nd), C.byref(mr)))
        return nd.value, mr.value

    def error_indicator(self, u, rho, nq, residual_term=True):
        u = np.ascontiguousarray(u, dtype=np.float64)
        rho = np.ascontiguousarray(rho, dtype=np.float64).ravel()
        thr = C.c_double()
        _ck_ms(self.L.ms_error_indicator(self.p, u.ctypes.data, len(rho), rho.ctypes.data, nq, int(residual_term), C.byref(thr)))
        return thr.value

    def transfer_from(self, old, old_res, u_old):
        u_old = np.ascontiguousarray(u_old, dtype=np.float64)
        out = np.zeros(len(self.get("boundary")))
        _ck_ms(self.L.ms_transfer(self.p, old_res, old.p, u_old.ctypes.data, out.ctypes.data))
        return out

    def transfer_tables(self, old, old_res):
        """Index tables of the device solution transfer: (copy_old, copy_new, pass_ptr, parent_dofs[.., 27])."""
        nc, npass = C.c_int64(), C.c_int64()
        co, cn, pd = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        pp = C.POINTER(C.c_int64)()
        _ck_ms(self.L.ms_transfer_tables(self.p, int(old_res), old.p, C.byref(nc), C.byref(co), C.byref(cn), C.byref(npass),
                                         C.byref(pp), C.byref(pd)))
        take = lambda ptr, n, dt: np.ctypeslib.as_array(ptr, shape=(max(n, 1),)).astype(dt)[:n].copy()
        pass_ptr = take(pp, npass.value + 1, np.int64)
        return (take(co, nc.value, np.int32), take(cn, nc.value, np.int32), pass_ptr,
                take(pd, 27 * int(pass_ptr[-1]), np.int32).reshape(-1, 27))

    def distribute(self, g, x):
        g = np.ascontiguousarray(g, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64).copy()
        _ck_ms(self.L.ms_distribute(self.p, g.ctypes.data, x.ctypes.data))
        return x


def unit_stiffness():
    """Q1 Laplace cell matrix of the unit cube (8 x 8), the host's bits."""
    K = np.zeros((8, 8))
    ms_lib().ms_unit_stiffness(K.ctypes.data)
    return K


def assembly_inputs(M, which, level=0):
    """Inputs of Gmg.assemble_matrix from a built Ministep: (n_rows, cell_dofs, cell_h or None, uniform_h, flags, hang)."""
    if which == 0:
        dofs, hs = [], []
        for l in range(M.n_levels):
            cd = M.get("cell_dofs", l).reshape(-1, 8)
            dofs.append(cd)
            hs.append(np.full(len(cd), M.h0 / (1 << l)))
        hanging, dirichlet = M.get("hanging"), M.get("dirichlet")
        flags = np.where(hanging != 0, 2, np.where(dirichlet != 0, 1, 0)).astype(np.uint8)
        hang = (M.get("hang_rowptr"), M.get("hang_col"), M.get("hang_val")) if hanging.any() else None
        return len(flags), np.concatenate(dofs), np.concatenate(hs), 0.0, flags, hang
    cd = M.get("level_cell_dofs", level).reshape(-1, 8)
    flags = ((M.get("level_edge", level) != 0) | (M.get("level_boundary", level) != 0)).astype(np.uint8)
    return len(flags), cd, None, M.h0 / (1 << level), flags, None


def gauss(n):
    p, w = np.zeros(n), np.zeros(n)
    ms_lib().ms_gauss(n, p.ctypes.data, w.ctypes.data)
    return p, w


def check_prm(text):
    out = C.c_void_p()
    rc = lib().step50_check_prm(text.encode(), C.byref(out))
    if rc != 0:
        raise HostError(lib().ms_last_error().decode())
    s = C.cast(out, C.c_char_p).value.decode()
    lib().step50_free(out)
    return s


def run_problem(prm_text):
    """LaplaceProblem::run() from a parameter string: (stdout text, list of per-cycle records).  Needs a B200."""
    so, js = C.c_void_p(), C.c_void_p()
    rc = lib().step50_run_string(prm_text.encode(), C.byref(so), C.byref(js))
    if rc != 0:
        raise HostError(lib().ms_last_error().decode())
    text = C.cast(so, C.c_char_p).value.decode()
    recs = json.loads(C.cast(js, C.c_char_p).value.decode())
    lib().step50_free(so)
    lib().step50_free(js)
    return text, recs


class BenchProblem:
    """LaplaceProblem advanced to the last refinement cycle, hierarchy on the device, ready to repeat the hot path."""

    def __init__(self, prm_text, connect=None):
        """connect(gmg): optional callable that joins this rank's device context to its peers (multi-GPU) before the
        hierarchy of the last cycle is handed over."""
        self.L = lib()
        self.p = self.L.step50_bench_create(prm_text.encode())
        if not self.p:
            raise HostError(self.L.ms_last_error().decode())
        self.gmg = capi.Gmg.__new__(capi.Gmg)  # view of the problem's own device context (not owned)
        self.gmg.lib = capi.load_library()
        self.gmg.h = C.c_void_p(self.L.step50_bench_gmg(self.p))
        self.gmg.close = lambda: None
        if connect is not None:
            connect(self.gmg)
        _ck(self.L.step50_bench_finish_setup(self.p))
        info = np.zeros(16, dtype=np.int64)
        self.L.step50_bench_info(self.p, info.ctypes.data)
        self.n_dofs, self.n_cells, self.n_levels = int(info[0]), int(info[1]), int(info[2])
        self.level_n = [int(v) for v in info[3:3 + self.n_levels]]
        self.n_atoms, self.n_pairs, self.sys_nnz, self.nq = int(info[11]), int(info[12]), int(info[13]), int(info[14])
        self.n_pairs_active = int(info[15])

    def close(self):
        if self.p:
            self.L.step50_bench_destroy(self.p)
            self.p = None

    def step_device(self):
        its, res = C.c_int(0), C.c_double(0)
        _ck(self.L.step50_bench_step_device(self.p, C.byref(its), C.byref(res)))
        return its.value, res.value

    def step_host(self, with_hierarchy=True):
        its, res = C.c_int(0), C.c_double(0)
        _ck(self.L.step50_bench_step_host(self.p, int(with_hierarchy), C.byref(its), C.byref(res)))
        return its.value, res.value

    def set_device_assembly(self, on):
        """step_host(True) hands the hierarchy over with the system / level-0 matrices assembled on the device."""
        _ck(self.L.step50_bench_set_device_assembly(self.p, int(on)))

    def download_x(self):
        """Solution of the last step_device() (device -> host)."""
        out = np.zeros(self.n_dofs)
        _ck(self.L.step50_bench_download_x(self.p, out.ctypes.data))
        return out

    def download_b(self):
        """Load vector of the last step_device() (device -> host)."""
        out = np.zeros(self.n_dofs)
        _ck(self.L.step50_bench_download_b(self.p, out.ctypes.data))
        return out

    def time_binning(self):
        """rhs_assembly_optimization() once more through its host-buffer entry point: (ms, cell-atom pairs)."""
        ms, pairs = C.c_double(0), C.c_int64(0)
        _ck(self.L.step50_bench_time_binning(self.p, C.byref(ms), C.byref(pairs)))
        return ms.value, pairs.value

    def get(self, name, level=0):
        return _fetch(self.L.step50_bench_get, self.p, name, level)

    def mesh(self):
        lo, H, reps = C.c_double(0), C.c_double(0), C.c_int(0)
        self.L.step50_bench_mesh(self.p, C.byref(lo), C.byref(H), C.byref(reps))
        return lo.value, H.value, reps.value
