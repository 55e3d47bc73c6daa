"""ctypes binding of include/gmg_b200.h (the C ABI of the CUDA library).

Plumbing only: loads `lib/libgmg_b200.so`, declares every exported symbol and offers a thin
`Gmg` handle class used by the tests, `bench.py` and `__graft_entry__.smoke()`.  There is no
fallback: if the library is missing or no B200 is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libgmg_b200.so")

GMG_SYSTEM, GMG_LEVEL, GMG_EDGE, GMG_PROLONG = 0, 1, 2, 3
SMOOTHER_JACOBI, SMOOTHER_CHEBYSHEV, SMOOTHER_MC_SSOR, SMOOTHER_LEX_SSOR = 0, 1, 2, 3
GMG_ENOCONVERGENCE = -4

_h = C.c_void_p
_i, _i64, _d = C.c_int, C.c_int64, C.c_double
_pd = C.POINTER(C.c_double)
_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pu8 = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); every symbol declared in include/gmg_b200.h
SIGNATURES = {
    "gmg_create": (_i, [_i, C.POINTER(_h)]),
    "gmg_destroy": (_i, [_h]),
    "gmg_last_error": (C.c_char_p, [_h]),
    "gmg_set_stream": (_i, [_h, C.c_void_p]),
    "gmg_synchronize": (_i, [_h]),
    "gmg_compiled_arch": (_i, []),
    "gmg_set_num_levels": (_i, [_h, _i]),
    "gmg_set_matrix": (_i, [_h, _i, _i, C.c_int32, C.c_int32, _pi64, _pi32, _pd]),
    "gmg_assemble_matrix": (_i, [_h, _i, _i, C.c_int32, _i64, _pi32, _pd, _d, _pu8, _pi64, _pi32, _pd, _pd]),
    "gmg_raw_matrix_get": (_i, [_h, _i, _i, _pi64, _pi64, _pi32, _pd]),
    "gmg_set_copy_indices": (_i, [_h, _i, C.c_int32, _pi32, _pi32]),
    "gmg_set_smoother": (_i, [_h, _i, _d, _i]),
    "gmg_set_coarse": (_i, [_h, _i, _d]),
    "gmg_set_level_coloring": (_i, [_h, _i, C.c_int32, _pi32]),
    "gmg_set_graphs": (_i, [_h, _i]),
    "gmg_set_persistent_smoother": (_i, [_h, _i]),
    "gmg_set_drop_tolerance": (_i, [_h, _d]),
    "gmg_set_compression": (_i, [_h, _i]),
    "gmg_setup": (_i, [_h]),
    "gmg_pcg_solve": (_i, [_h, _pd, _pd, _i, _d, C.POINTER(_i), _pd, _pd]),
    "gmg_pcg_solve_jacobi": (_i, [_h, _pd, _pd, _d, _i, _d, C.POINTER(_i), _pd, _pd]),
    "gmg_vcycle_apply": (_i, [_h, _pd, _pd]),
    "gmg_spmv": (_i, [_h, _i, _i, _pd, _pd]),
    "gmg_cg_solve": (_i, [_h, _i, _i, _pd, _pd, _i, _d, C.POINTER(_i), _pd]),
    "gmg_smooth": (_i, [_h, _i, _pd, _pd, _i]),
    "gmg_matrix_norms": (_i, [_h, _i, _i, _pd]),
    "gmg_vector_norms": (_i, [_h, _i64, _pd, _pd]),
    "gmg_last_coarse_iterations": (_i, [_h, _pi32, _i, C.POINTER(_i)]),
    "gmg_vec_alloc": (_i, [_h, _i64, C.POINTER(C.c_void_p)]),
    "gmg_vec_free": (_i, [_h, C.c_void_p]),
    "gmg_vec_upload": (_i, [_h, C.c_void_p, _pd, _i64]),
    "gmg_vec_download": (_i, [_h, _pd, C.c_void_p, _i64]),
    "gmg_vec_copy_dev": (_i, [_h, C.c_void_p, C.c_void_p, _i64]),
    "gmg_transfer_bytes": (_i, [_h, _i, _pi64, _pi64]),
    "gmg_pcg_solve_dev": (_i, [_h, C.c_void_p, C.c_void_p, _i, _d, C.POINTER(_i), _pd, _pd]),
    "gmg_vcycle_apply_dev": (_i, [_h, C.c_void_p, C.c_void_p]),
    "gmg_spmv_dev": (_i, [_h, _i, _i, C.c_void_p, C.c_void_p]),
    "gmg_cg_solve_dev": (_i, [_h, _i, _i, C.c_void_p, C.c_void_p, _i, _d, C.POINTER(_i), _pd]),
    "gmg_matrix_traffic": (_i, [_h, _i, _i, _pd]),
    "gmg_pair_energies": (_i, [_h, _d, _pd]),
    "gmg_energy_norm_error": (_i, [_h, _pd, C.c_int32, _d, _pd, _pd, _pd]),
    "gmg_error_indicator": (_i, [_h, _i, _pi32, _pu8, _i, _pi32, _pd, _i, _pd, _i, _pd, _pd, C.POINTER(C.c_float),
                                 C.POINTER(C.c_float)]),
    "gmg_mark_cells": (_i, [_h, _i, _d, _pu8, _pd]),
    "gmg_transfer_solution": (_i, [_h, _i, _pd, _i, _i, _pi32, _pi32, _i, _pi64, _pi32, _pu8, _pd]),
    "gmg_debug_cg_phases": (_i, [_h, _i, _pd]),
    "gmg_debug_cg_blocks": (_i, [_h, _pd]),
    "gmg_coarse_kernel": (_i, [_h, _i, _i, C.POINTER(_i)]),
    "gmg_debug_vcycle_profile": (_i, [_h, _i, _pd]),
    "gmg_coarse_profile": (_i, [_h, _i, _pd, _pi64, _pi64]),
    "gmg_launch_count": (_i64, [_h]),
    "gmg_dist_init": (_i, [_h, _i, _i, _i64, C.c_void_p]),
    "gmg_dist_connect": (_i, [_h, C.c_void_p]),
    "gmg_set_ownership": (_i, [_h, _i, _i, C.c_int32, _pi32]),
    "gmg_dist_rank": (_i, [_h, C.POINTER(_i), C.POINTER(_i)]),
    "gmg_partition_probe": (_i, [_i, _i, C.c_int32, _pi64, _pi32, _pd, _pi32, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(_pi64), C.POINTER(_pi32), C.POINTER(_pd), C.POINTER(_pi32), C.POINTER(_pi32),
                                 C.POINTER(_pi32), C.POINTER(_pi32), C.POINTER(_pi32)]),
    "gmg_free_host": (None, [C.c_void_p]),
    "gmg_dist_pingpong": (_i, [_h, _i, _i, _pd]),
    "gmg_bin_atoms": (_i, [_h, C.c_int32, _pd, _pd, C.c_int32, _pd, _d, _pi64, _pi32]),
    "gmg_set_atom_lists": (_i, [_h, C.c_int32, _pi64, _pi32]),
    "gmg_set_atoms": (_i, [_h, C.c_int32, _pd, _pd]),
    "gmg_charge_density": (_i, [_h, C.c_int32, _pd, _pd, _pi32, C.c_int32, _pd, _d, _pd]),
    "gmg_assemble_rhs": (_i, [_h, C.c_int32, _pd, _pd, _pi32, C.c_int32, _pd, _pd, _pd, _pd, C.c_int32, _pi64, _pi32,
                              _pd, _pu8, _pd]),
    "gmg_point_values": (_i, [_h, C.c_int32, _pi32, _pd, _pd, C.c_int32, _pd]),
    "gmg_rhs_step_dev": (_i, [_h, C.c_void_p]),
}

_lib = None


def load_library(path=None):
    """dlopen the CUDA library and attach the signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GmgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gmg error {code}: {msg}")
        self.code = code


class NoConvergence(GmgError):
    """SolverControl::NoConvergence of the reference (src/step-50.cc:942, 962)."""


def _pd_of(a):
    return a.ctypes.data_as(_pd)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Gmg:
    """One device context (one GPU).  Mirrors the handle-based C ABI one to one."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = _h()
        rc = self.lib.gmg_create(int(device), C.byref(self.h))
        if rc != 0:
            raise GmgError(rc, "gmg_create failed: no B200 (sm_100) CUDA device available; there is no CPU fallback")

    def close(self):
        if self.h:
            self.lib.gmg_destroy(self.h)
            self.h = _h()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.gmg_last_error(self.h).decode()
            raise (NoConvergence if rc == GMG_ENOCONVERGENCE else GmgError)(rc, msg)

    # ---- hierarchy
    def set_num_levels(self, n):
        self._ck(self.lib.gmg_set_num_levels(self.h, n))

    def set_matrix(self, which, level, csr):
        csr = csr.tocsr()
        rp = np.ascontiguousarray(csr.indptr, dtype=np.int64)
        col = _i32(csr.indices)
        val = _f64(csr.data)
        self._ck(self.lib.gmg_set_matrix(self.h, which, level, csr.shape[0], csr.shape[1], rp.ctypes.data_as(_pi64),
                                         col.ctypes.data_as(_pi32), _pd_of(val)))

    def assemble_matrix(self, which, level, n_rows, cell_dofs, cell_h, row_flags, k_ref, hang=None, uniform_h=0.0):
        """gmg_assemble_matrix: cell_h an array or None (uniform_h); hang = (rowptr, col, val) or None."""
        cell_dofs = _i32(cell_dofs).reshape(-1, 8)
        flags = np.ascontiguousarray(row_flags, dtype=np.uint8)
        k_ref = _f64(k_ref).ravel()
        assert k_ref.size == 64 and flags.size == n_rows
        ch = _f64(cell_h) if cell_h is not None else None
        hp = hc = hv = None
        if hang is not None:
            hp = np.ascontiguousarray(hang[0], dtype=np.int64)
            hc, hv = _i32(hang[1]), _f64(hang[2])
        self._ck(self.lib.gmg_assemble_matrix(
            self.h, which, level, n_rows, cell_dofs.shape[0], cell_dofs.ctypes.data_as(_pi32),
            _pd_of(ch) if ch is not None else None, float(uniform_h), flags.ctypes.data_as(_pu8),
            hp.ctypes.data_as(_pi64) if hp is not None else None, hc.ctypes.data_as(_pi32) if hc is not None else None,
            _pd_of(hv) if hv is not None else None, _pd_of(k_ref)))

    def raw_matrix(self, which, level, n_rows):
        """(rowptr, col, val) of a handed-over / assembled matrix before gmg_setup consumes it."""
        nnz = _i64(0)
        self._ck(self.lib.gmg_raw_matrix_get(self.h, which, level, C.byref(nnz), None, None, None))
        rp = np.zeros(n_rows + 1, dtype=np.int64)
        col = np.zeros(nnz.value, dtype=np.int32)
        val = np.zeros(nnz.value, dtype=np.float64)
        self._ck(self.lib.gmg_raw_matrix_get(self.h, which, level, C.byref(nnz), rp.ctypes.data_as(_pi64),
                                             col.ctypes.data_as(_pi32), _pd_of(val)))
        return rp, col, val

    def set_copy_indices(self, level, g, l):
        g, l = _i32(g), _i32(l)
        self._ck(self.lib.gmg_set_copy_indices(self.h, level, len(g), g.ctypes.data_as(_pi32), l.ctypes.data_as(_pi32)))

    def set_smoother(self, kind, omega=0.5, steps=2):
        self._ck(self.lib.gmg_set_smoother(self.h, kind, omega, steps))

    def set_level_coloring(self, level, color):
        color = _i32(color)
        self._ck(self.lib.gmg_set_level_coloring(self.h, level, len(color), color.ctypes.data_as(_pi32)))

    def set_persistent_smoother(self, on):
        self._ck(self.lib.gmg_set_persistent_smoother(self.h, int(on)))

    def set_graphs(self, on):
        self._ck(self.lib.gmg_set_graphs(self.h, int(on)))

    def set_coarse(self, max_it=1000, tol=1e-10):
        self._ck(self.lib.gmg_set_coarse(self.h, max_it, tol))

    def set_compression(self, on):
        self._ck(self.lib.gmg_set_compression(self.h, int(on)))

    def set_drop_tolerance(self, tol):
        self._ck(self.lib.gmg_set_drop_tolerance(self.h, tol))

    def setup(self):
        self._ck(self.lib.gmg_setup(self.h))

    def set_stream(self, stream_ptr):
        self._ck(self.lib.gmg_set_stream(self.h, C.c_void_p(stream_ptr)))

    def synchronize(self):
        self._ck(self.lib.gmg_synchronize(self.h))

    # ---- solve path (host buffers)
    def pcg_solve(self, b, x0, max_it=500, tol=1e-8):
        b, x = _f64(b), _f64(x0).copy()
        it, r0, r1 = _i(0), _d(0), _d(0)
        rc = self.lib.gmg_pcg_solve(self.h, _pd_of(b), _pd_of(x), max_it, tol, C.byref(it), C.byref(r0), C.byref(r1))
        self._ck(rc)
        return x, it.value, r0.value, r1.value

    def pcg_solve_jacobi(self, b, x0, omega=0.6, max_it=500, tol=1e-8):
        b, x = _f64(b), _f64(x0).copy()
        it, r0, r1 = _i(0), _d(0), _d(0)
        self._ck(self.lib.gmg_pcg_solve_jacobi(self.h, _pd_of(b), _pd_of(x), omega, max_it, tol, C.byref(it),
                                               C.byref(r0), C.byref(r1)))
        return x, it.value, r0.value, r1.value

    def vcycle(self, src):
        src = _f64(src)
        dst = np.zeros_like(src)
        self._ck(self.lib.gmg_vcycle_apply(self.h, _pd_of(src), _pd_of(dst)))
        return dst

    def spmv(self, which, level, x, n_rows):
        x = _f64(x)
        y = np.zeros(n_rows)
        self._ck(self.lib.gmg_spmv(self.h, which, level, _pd_of(x), _pd_of(y)))
        return y

    def cg_solve(self, which, level, b, max_it=1000, tol=1e-10):
        b = _f64(b)
        x = np.zeros_like(b)
        it, r = _i(0), _d(0)
        self._ck(self.lib.gmg_cg_solve(self.h, which, level, _pd_of(b), _pd_of(x), max_it, tol, C.byref(it), C.byref(r)))
        return x, it.value, r.value

    def smooth(self, level, rhs, u, zero_start):
        rhs, u = _f64(rhs), _f64(u).copy()
        self._ck(self.lib.gmg_smooth(self.h, level, _pd_of(rhs), _pd_of(u), int(zero_start)))
        return u

    def matrix_norms(self, which, level=0):
        out = np.zeros(3)
        self._ck(self.lib.gmg_matrix_norms(self.h, which, level, _pd_of(out)))
        return out

    def vector_norms(self, v):
        v = _f64(v)
        out = np.zeros(3)
        self._ck(self.lib.gmg_vector_norms(self.h, len(v), _pd_of(v), _pd_of(out)))
        return out

    def last_coarse_iterations(self):
        buf = np.zeros(4096, dtype=np.int32)
        n = _i(0)
        self._ck(self.lib.gmg_last_coarse_iterations(self.h, buf.ctypes.data_as(_pi32), len(buf), C.byref(n)))
        return buf[:min(n.value, len(buf))].tolist()

    # ---- device-resident variants
    def vec_alloc(self, n):
        p = C.c_void_p()
        self._ck(self.lib.gmg_vec_alloc(self.h, n, C.byref(p)))
        return p

    def vec_free(self, p):
        self._ck(self.lib.gmg_vec_free(self.h, p))

    def vec_upload(self, p, host):
        host = _f64(host)
        self._ck(self.lib.gmg_vec_upload(self.h, p, _pd_of(host), len(host)))

    def vec_download(self, p, n):
        out = np.zeros(n)
        self._ck(self.lib.gmg_vec_download(self.h, _pd_of(out), p, n))
        return out

    def pcg_solve_dev(self, b_dev, x_dev, max_it=500, tol=1e-8):
        it, r0, r1 = _i(0), _d(0), _d(0)
        self._ck(self.lib.gmg_pcg_solve_dev(self.h, b_dev, x_dev, max_it, tol, C.byref(it), C.byref(r0), C.byref(r1)))
        return it.value, r0.value, r1.value

    def vcycle_dev(self, src_dev, dst_dev):
        self._ck(self.lib.gmg_vcycle_apply_dev(self.h, src_dev, dst_dev))

    def spmv_dev(self, which, level, x_dev, y_dev):
        self._ck(self.lib.gmg_spmv_dev(self.h, which, level, x_dev, y_dev))

    def cg_solve_dev(self, which, level, b_dev, x_dev, max_it=1000, tol=1e-10):
        it, r = _i(0), _d(0)
        self._ck(self.lib.gmg_cg_solve_dev(self.h, which, level, b_dev, x_dev, max_it, tol, C.byref(it), C.byref(r)))
        return it.value, r.value

    def matrix_traffic(self, which, level=0):
        out = np.zeros(6)
        self._ck(self.lib.gmg_matrix_traffic(self.h, which, level, _pd_of(out)))
        return dict(nnz=out[0], spmv_bytes=out[1], cg_iter_bytes=out[2], csr_spmv_bytes=out[3], csr_cg_iter_bytes=out[4],
                    compressed=bool(out[5]), format=int(out[5]))

    def error_indicator(self, face_nb, face_kind, hang_children, u, rho, residual_term, gauss2_points, gauss2_weights):
        """eta (float32 per active cell of the last assemble_rhs call) and max eta; rho None = device-resident densities."""
        face_nb, hang_children = _i32(face_nb), _i32(hang_children)
        face_kind = np.ascontiguousarray(face_kind, dtype=np.uint8)
        u, gp, gw = _f64(u), _f64(gauss2_points), _f64(gauss2_weights)
        n_cells = len(face_nb) // 6
        rho_a = _f64(rho).ravel() if rho is not None else None
        eta = np.zeros(n_cells, dtype=np.float32)
        mx = C.c_float(0)
        self._ck(self.lib.gmg_error_indicator(
            self.h, n_cells, face_nb.ctypes.data_as(_pi32), face_kind.ctypes.data_as(_pu8), len(hang_children) // 4,
            hang_children.ctypes.data_as(_pi32), _pd_of(u), len(u), _pd_of(rho_a) if rho_a is not None else None,
            int(residual_term), _pd_of(gp), _pd_of(gw), eta.ctypes.data_as(C.POINTER(C.c_float)), C.byref(mx)))
        return eta, mx.value

    def mark_cells(self, n_cells, fraction=0.6):
        """Refinement flags (uint8 per active cell of the last error_indicator call) and the threshold fraction * max."""
        flags = np.zeros(n_cells, dtype=np.uint8)
        thr = np.zeros(1)
        self._ck(self.lib.gmg_mark_cells(self.h, int(n_cells), float(fraction), flags.ctypes.data_as(_pu8), _pd_of(thr)))
        return flags, float(thr[0])

    def transfer_solution(self, u_old, n_new, copy_old, copy_new, pass_ptr, parent_dofs, constrained):
        """SolutionTransfer::interpolate + set_zero on the device (index tables from the host mesh)."""
        u_old, copy_old, copy_new = _f64(u_old), _i32(copy_old), _i32(copy_new)
        pass_ptr = np.ascontiguousarray(pass_ptr, dtype=np.int64)
        pd = _i32(np.asarray(parent_dofs).ravel())
        con = np.ascontiguousarray(constrained, dtype=np.uint8)
        out = np.zeros(n_new)
        self._ck(self.lib.gmg_transfer_solution(
            self.h, len(u_old), _pd_of(u_old), int(n_new), len(copy_old), copy_old.ctypes.data_as(_pi32),
            copy_new.ctypes.data_as(_pi32), len(pass_ptr) - 1, pass_ptr.ctypes.data_as(_pi64), pd.ctypes.data_as(_pi32),
            con.ctypes.data_as(_pu8), _pd_of(out)))
        return out

    def pair_energies(self, r_c):
        out = np.zeros(2)
        self._ck(self.lib.gmg_pair_energies(self.h, float(r_c), _pd_of(out)))
        return dict(analytic=out[0], short=out[1])

    def energy_norm_error(self, u, r_c, gauss2_points, gauss2_weights):
        u, gp, gw = _f64(u), _f64(gauss2_points), _f64(gauss2_weights)
        out = np.zeros(1)
        self._ck(self.lib.gmg_energy_norm_error(self.h, _pd_of(u), len(u), float(r_c), _pd_of(gp), _pd_of(gw), _pd_of(out)))
        return float(out[0])

    def debug_cg_phases(self, block_plus_1=1):
        out = np.zeros(16)
        self._ck(self.lib.gmg_debug_cg_phases(self.h, int(block_plus_1), _pd_of(out)))
        return out

    def debug_vcycle_profile(self, enable=True):
        out = np.zeros(4)
        self._ck(self.lib.gmg_debug_vcycle_profile(self.h, int(enable), _pd_of(out)))
        return dict(down_ms=out[0], coarse_ms=out[1], up_ms=out[2], vcycles=int(out[3]))

    def coarse_kernel(self, which, level=0):
        k = _i(0)
        self._ck(self.lib.gmg_coarse_kernel(self.h, which, level, C.byref(k)))
        return k.value

    def debug_cg_blocks(self):
        out = np.zeros(768)
        self._ck(self.lib.gmg_debug_cg_blocks(self.h, _pd_of(out)))
        return out.reshape(3, 256)

    def coarse_profile(self, reset=True):
        ms, n, it = _d(0), _i64(0), _i64(0)
        self._ck(self.lib.gmg_coarse_profile(self.h, int(reset), C.byref(ms), C.byref(n), C.byref(it)))
        return dict(ms=ms.value, launches=n.value, iterations=it.value)

    def transfer_bytes(self, reset=True):
        a, b = _i64(0), _i64(0)
        self._ck(self.lib.gmg_transfer_bytes(self.h, int(reset), C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self):
        return int(self.lib.gmg_launch_count(self.h))

    # ---- multi-GPU
    def dist_init(self, rank, world, comm_bytes=512 << 20):
        """Allocate this rank's peer-mapped communication buffer; returns its 64-byte CUDA IPC handle."""
        buf = (C.c_char * 64)()
        self._ck(self.lib.gmg_dist_init(self.h, rank, world, comm_bytes, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def dist_connect(self, all_handles):
        blob = b"".join(all_handles)
        self._ck(self.lib.gmg_dist_connect(self.h, C.cast(C.c_char_p(blob), C.c_void_p)))

    def pingpong(self, iters=1000, mode=0):
        us = _d(0)
        self._ck(self.lib.gmg_dist_pingpong(self.h, iters, mode, C.byref(us)))
        return us.value

    def set_ownership(self, which, level, owner):
        owner = _i32(owner)
        self._ck(self.lib.gmg_set_ownership(self.h, which, level, len(owner), owner.ctypes.data_as(_pi32)))

    # ---- RHS path
    def set_atoms(self, pos, charge):
        pos, charge = _f64(pos), _f64(charge)
        self._ck(self.lib.gmg_set_atoms(self.h, len(charge), _pd_of(pos), _pd_of(charge)))

    def set_atom_lists(self, ptr, idx):
        ptr = np.ascontiguousarray(ptr, dtype=np.int64)
        idx = _i32(idx)
        self._ck(self.lib.gmg_set_atom_lists(self.h, len(ptr) - 1, ptr.ctypes.data_as(_pi64), idx.ctypes.data_as(_pi32)))

    def bin_atoms(self, cell_lo, cell_h, pos, radius):
        cell_lo, cell_h, pos = _f64(cell_lo), _f64(cell_h), _f64(pos)
        n = len(cell_h)
        ptr = np.zeros(n + 1, dtype=np.int64)
        self._ck(self.lib.gmg_bin_atoms(self.h, n, _pd_of(cell_lo), _pd_of(cell_h), len(pos), _pd_of(pos), radius,
                                        ptr.ctypes.data_as(_pi64), None))
        idx = np.zeros(max(int(ptr[-1]), 1), dtype=np.int32)
        self._ck(self.lib.gmg_bin_atoms(self.h, n, _pd_of(cell_lo), _pd_of(cell_h), len(pos), _pd_of(pos), radius,
                                        ptr.ctypes.data_as(_pi64), idx.ctypes.data_as(_pi32)))
        return ptr, idx[:int(ptr[-1])]

    def charge_density(self, cell_lo, cell_h, list_of_cell, qpoints, r_c, download=True):
        cell_lo, cell_h, qpoints = _f64(cell_lo), _f64(cell_h), _f64(qpoints)
        loc = _i32(list_of_cell)
        n, nq = len(cell_h), len(qpoints)
        rho = np.zeros((n, nq)) if download else None
        self._ck(self.lib.gmg_charge_density(self.h, n, _pd_of(cell_lo), _pd_of(cell_h), loc.ctypes.data_as(_pi32), nq,
                                             _pd_of(qpoints), r_c, _pd_of(rho) if download else None))
        return rho

    def assemble_rhs(self, rho, cell_h, cell_dofs, shape, weights, n_dofs, hang_ptr, hang_col, hang_val, constrained,
                     kref=None, ghat=None):
        cell_h, shape, weights = _f64(cell_h), _f64(shape), _f64(weights)
        cell_dofs = _i32(cell_dofs)
        hang_ptr = np.ascontiguousarray(hang_ptr, dtype=np.int64)
        hang_col, hang_val = _i32(hang_col), _f64(hang_val)
        constrained = np.ascontiguousarray(constrained, dtype=np.uint8)
        rho_p = _pd_of(_f64(rho)) if rho is not None else None
        kref_a = _f64(kref) if kref is not None else None
        ghat_a = _f64(ghat) if ghat is not None else None
        b = np.zeros(n_dofs)
        self._ck(self.lib.gmg_assemble_rhs(
            self.h, len(cell_h), rho_p, _pd_of(cell_h), cell_dofs.ctypes.data_as(_pi32), len(weights), _pd_of(shape),
            _pd_of(weights), _pd_of(kref_a) if kref_a is not None else None,
            _pd_of(ghat_a) if ghat_a is not None else None, n_dofs, hang_ptr.ctypes.data_as(_pi64),
            hang_col.ctypes.data_as(_pi32), _pd_of(hang_val), constrained.ctypes.data_as(_pu8), _pd_of(b)))
        return b

    def point_values(self, cell_dofs, ref_coords, u):
        cell_dofs, ref_coords, u = _i32(cell_dofs), _f64(ref_coords), _f64(u)
        out = np.zeros(len(cell_dofs))
        self._ck(self.lib.gmg_point_values(self.h, len(cell_dofs), cell_dofs.ctypes.data_as(_pi32), _pd_of(ref_coords),
                                           _pd_of(u), len(u), _pd_of(out)))
        return out

    def rhs_step_dev(self, b_dev):
        self._ck(self.lib.gmg_rhs_step_dev(self.h, b_dev))


def hand_over_hierarchy(gmg, system_A, level_A, level_I, prolong, copy_global, copy_level):
    """Give an assembled hierarchy (scipy CSR matrices, index arrays) to the device, as `solve()` of
    the reference wires mg_matrices / mg_interface_matrices / MGTransferPrebuilt (src/step-50.cc:957-989)."""
    nl = len(level_A)
    gmg.set_num_levels(nl)
    gmg.set_matrix(GMG_SYSTEM, 0, system_A)
    for l in range(nl):
        gmg.set_matrix(GMG_LEVEL, l, level_A[l])
        if l >= 1:
            gmg.set_matrix(GMG_EDGE, l, level_I[l])
        if l + 1 < nl:
            gmg.set_matrix(GMG_PROLONG, l, prolong[l])
        gmg.set_copy_indices(l, copy_global[l], copy_level[l])


def connect_ranks(gmg, rank, world, all_gather_bytes, comm_bytes=512 << 20):
    """Map every rank's communication buffer on every other rank.  `all_gather_bytes(b) -> [b_0 .. b_{W-1}]` is the
    caller's collective (torch.distributed / MPI); it also acts as the barrier between init and connect."""
    handle = gmg.dist_init(rank, world, comm_bytes)
    gmg.dist_connect(all_gather_bytes(handle))


def partition_probe(rank, world, csr, owner):
    """Host-only: the local matrix / halo maps rank `rank` would build (see csrc/partition.h)."""
    lib = load_library()
    csr = csr.tocsr()
    rp = np.ascontiguousarray(csr.indptr, dtype=np.int64)
    col, val, owner = _i32(csr.indices), _f64(csr.data), _i32(owner)
    n_owned, n_halo = C.c_int32(), C.c_int32()
    l_rp, l_col, l_val = _pi64(), _pi32(), _pd()
    og, hg, sc, si, sb = _pi32(), _pi32(), _pi32(), _pi32(), _pi32()
    rc = lib.gmg_partition_probe(rank, world, csr.shape[0], rp.ctypes.data_as(_pi64), col.ctypes.data_as(_pi32), _pd_of(val),
                                 owner.ctypes.data_as(_pi32), C.byref(n_owned), C.byref(n_halo), C.byref(l_rp), C.byref(l_col),
                                 C.byref(l_val), C.byref(og), C.byref(hg), C.byref(sc), C.byref(si), C.byref(sb))
    if rc != 0:
        raise GmgError(rc, "gmg_partition_probe failed")
    no, nh = n_owned.value, n_halo.value
    take = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(max(n, 1),)).astype(dt)[:n].copy()
    rowptr = take(l_rp, no + 1, np.int64)
    nnz = int(rowptr[-1])
    out = dict(n_owned=no, n_halo=nh, rowptr=rowptr, col=take(l_col, nnz, np.int32), val=take(l_val, nnz, np.float64),
               owned_global=take(og, no, np.int32), halo_global=take(hg, nh, np.int32), send_dst_base=take(sb, world, np.int32))
    sc_all = take(sc, world + 1, np.int32)
    out["send_count"], out["n_halo_lo"] = sc_all[:world], int(sc_all[world])
    out["send_idx"] = take(si, int(out["send_count"].sum()), np.int32)
    for p in (l_rp, l_col, l_val, og, hg, sc, si, sb):
        lib.gmg_free_host(C.cast(p, C.c_void_p))
    return out
