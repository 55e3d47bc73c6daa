// gmg_b200 C ABI, RHS path: atom->cell binning, Gaussian charge densities, load vector, point values.
// Replaces rhs_assembly_optimization (src/step-50.cc:260-306), compute_charge_densities (:509-575)
// and the load-vector part of assemble_system (:798-828) of the reference.
#include <algorithm>
#include <cmath>
#include <cub/device/device_segmented_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <vector>

#include "context.h"

using namespace gmg;

struct RhsState {
  // cells of the last gmg_charge_density call
  int n_cells = 0, n_q = 0;
  double *cell_lo = nullptr, *cell_h = nullptr, *qpts = nullptr, *rho = nullptr;
  int *list_of_cell = nullptr;
  double r_c = 0.0;
  int *nonempty = nullptr;  // cells with a non-empty atom list (the only ones the density kernels visit)
  int n_nonempty = -1;
  int tensor_n1 = 0;  // > 0: the quadrature points are a tensor grid of tensor_n1^3 points, x fastest (QGauss<3>)
  double tensor_p[3][8] = {};
  // inputs of the last gmg_assemble_rhs call
  int a_cells = 0, a_nq = 0, n_dofs = 0;
  double *a_h = nullptr, *shape = nullptr, *weights = nullptr, *kref = nullptr, *ghat = nullptr, *hang_val = nullptr;
  int *cell_dofs = nullptr, *hang_col = nullptr;
  int64_t *hang_ptr = nullptr;
  uint8_t *constrained = nullptr;
  bool have_kref = false, have_ghat = false;
  bool bin_pending = false;  // first call of the two-call gmg_bin_atoms protocol done
  // multi-GPU: the cells that touch a dof this rank owns (the only ones its part of b depends on)
  int *dist_cells = nullptr;
  int n_dist_cells = -1;
};

namespace {

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

template <class T>
int upload(gmg_context *h, T *&dst, const T *src, int64_t n) {
  dfree(dst);
  GMG_CUDA(h, dalloc(&dst, n));
  if (n > 0)
    if (int rc = gmg::staged_h2d(h, dst, src, sizeof(T) * n)) return rc;
  return GMG_OK;
}

// Pair sums of the energy post-processing (src/step-50.cc:1316-1332): sum_{i<j} q_i q_j / r_ij and
// sum_{i<j} q_i q_j erfc(r_ij / r_c) / r_ij.  One thread per atom i, atoms j > i streamed through shared memory in
// tiles; fp64 throughout; per-block partial sums are added on the host in block order (deterministic).
constexpr int PAIR_BLOCK = 128;
__global__ void __launch_bounds__(PAIR_BLOCK) pair_energy_kernel(int n, const double *__restrict__ pos, const double *__restrict__ q,
                                                                   double inv_rc, double *__restrict__ partial /* [2][gridDim.x] */) {
  __shared__ double sx[PAIR_BLOCK], sy[PAIR_BLOCK], sz[PAIR_BLOCK], sq[PAIR_BLOCK];
  __shared__ double red[2][PAIR_BLOCK / 32];
  const int i = blockIdx.x * PAIR_BLOCK + threadIdx.x;
  const bool on = i < n;
  const double xi = on ? pos[3 * i] : 0.0, yi = on ? pos[3 * i + 1] : 0.0, zi = on ? pos[3 * i + 2] : 0.0, qi = on ? q[i] : 0.0;
  double e_coul = 0.0, e_short = 0.0;
  for (int j0 = blockIdx.x * PAIR_BLOCK; j0 < n; j0 += PAIR_BLOCK) {  // tiles at or after this block's own
    const int j = j0 + threadIdx.x;
    __syncthreads();
    sx[threadIdx.x] = j < n ? pos[3 * j] : 0.0;
    sy[threadIdx.x] = j < n ? pos[3 * j + 1] : 0.0;
    sz[threadIdx.x] = j < n ? pos[3 * j + 2] : 0.0;
    sq[threadIdx.x] = j < n ? q[j] : 0.0;
    __syncthreads();
    const int m = min(PAIR_BLOCK, n - j0);
    for (int t = 0; t < m; ++t) {
      if (j0 + t > i && on) {
        const double dx = xi - sx[t], dy = yi - sy[t], dz = zi - sz[t];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        const double qq = qi * sq[t];
        e_coul += qq / r;
        e_short += qi * (sq[t] * (erfc(r * inv_rc) / r));
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    e_coul += __shfl_xor_sync(0xffffffffu, e_coul, o);
    e_short += __shfl_xor_sync(0xffffffffu, e_short, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = e_coul;
    red[1][threadIdx.x >> 5] = e_short;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < PAIR_BLOCK / 32; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    partial[blockIdx.x] = a;
    partial[gridDim.x + blockIdx.x] = b;
  }
}

// postprocess_error_in_energy_norm (src/step-50.cc:1423-1461): sqrt(sum_cells sum_q |grad u_h - grad u_exact|^2 JxW) with
// QGauss<3>(2) and the exact gradient summed over ALL atoms (O(cells * 8 * atoms), the reference runs it every cycle).
// One thread per (cell, quadrature point), atoms streamed through shared memory in tiles; per-block partial sums are
// added on the host in block order.
constexpr int ENORM_BLOCK = 128;
__global__ void __launch_bounds__(ENORM_BLOCK) energy_norm_kernel(int n_cells, const double *__restrict__ cell_lo,
                                                                    const double *__restrict__ cell_h, const int *__restrict__ cell_dofs,
                                                                    const double *__restrict__ u, int n_atoms,
                                                                    const double *__restrict__ pos, const double *__restrict__ q,
                                                                    double r_c, double g0, double g1, double w0, double w1,
                                                                    double *__restrict__ partial) {
  __shared__ double sx[ENORM_BLOCK], sy[ENORM_BLOCK], sz[ENORM_BLOCK], sq[ENORM_BLOCK];
  __shared__ double red[ENORM_BLOCK / 32];
  const int64_t t = (int64_t)blockIdx.x * ENORM_BLOCK + threadIdx.x;
  const int c = (int)(t >> 3), qp = (int)(t & 7);
  const bool on = c < n_cells;
  const int qx = qp & 1, qy = (qp >> 1) & 1, qz = qp >> 2;
  const double xi[3] = {qx ? g1 : g0, qy ? g1 : g0, qz ? g1 : g0};
  const double wq = (qx ? w1 : w0) * (qy ? w1 : w0) * (qz ? w1 : w0);
  double gh[3] = {0.0, 0.0, 0.0}, x[3] = {0.0, 0.0, 0.0}, h = 1.0;
  if (on) {
    h = cell_h[c];
    for (int v = 0; v < 8; ++v) {
      const double uv = u[cell_dofs[8 * (size_t)c + v]];
      for (int g = 0; g < 3; ++g) {
        double w = 1.0;
        for (int k = 0; k < 3; ++k) {
          const bool hi = (v >> k) & 1;
          w *= (k == g) ? (hi ? 1.0 : -1.0) / h : (hi ? xi[k] : 1.0 - xi[k]);
        }
        gh[g] += uv * w;
      }
    }
    for (int k = 0; k < 3; ++k) x[k] = cell_lo[3 * (size_t)c + k] + xi[k] * h;
  }
  const double inv_constant = 1.0 / (sqrt(M_PI) * r_c), inv_rc = 1.0 / r_c;
  double ga[3] = {0.0, 0.0, 0.0};
  for (int j0 = 0; j0 < n_atoms; j0 += ENORM_BLOCK) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    sx[threadIdx.x] = j < n_atoms ? pos[3 * (size_t)j] : 0.0;
    sy[threadIdx.x] = j < n_atoms ? pos[3 * (size_t)j + 1] : 0.0;
    sz[threadIdx.x] = j < n_atoms ? pos[3 * (size_t)j + 2] : 0.0;
    sq[threadIdx.x] = j < n_atoms ? q[j] : 0.0;
    __syncthreads();
    const int m = min(ENORM_BLOCK, n_atoms - j0);
    for (int k = 0; k < m; ++k) {
      const double dx = x[0] - sx[k], dy = x[1] - sy[k], dz = x[2] - sz[k];
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      const double s = r * inv_rc;
      const double fac = sq[k] * (((2.0 * r * exp(-(s * s)) * inv_constant) - erf(s)) / (r * r)) / r;
      ga[0] += fac * dx;
      ga[1] += fac * dy;
      ga[2] += fac * dz;
    }
  }
  double e = 0.0;
  if (on) {
    for (int g = 0; g < 3; ++g) e += (gh[g] - ga[g]) * (gh[g] - ga[g]);
    e *= wq * h * h * h;
  }
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < ENORM_BLOCK / 32; ++w) a += red[w];
    partial[blockIdx.x] = a;
  }
}

RhsState *state(gmg_context *h) {
  if (!h->rhs) h->rhs = new RhsState();
  return h->rhs;
}

// ---------------------------------------------------------------------------------------- binning
struct HashGrid {
  double lo[3], hi[3];
  double inv_s;
  int n[3];
};

__device__ __forceinline__ int hash_coord(const HashGrid &g, double x, int d) {
  int c = (int)floor((x - g.lo[d]) * g.inv_s);
  return min(max(c, 0), g.n[d] - 1);
}

__global__ void hash_count(int n_atoms, const double *__restrict__ pos, HashGrid g, int *__restrict__ cell_of_atom,
                           int *__restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_atoms) return;
  const int cx = hash_coord(g, pos[3 * i], 0), cy = hash_coord(g, pos[3 * i + 1], 1), cz = hash_coord(g, pos[3 * i + 2], 2);
  const int c = cx + g.n[0] * (cy + g.n[1] * cz);
  cell_of_atom[i] = c;
  atomicAdd(count + c, 1);
}

__global__ void hash_fill(int n_atoms, const int *__restrict__ cell_of_atom, const int *__restrict__ start,
                          int *__restrict__ cursor, int *__restrict__ sorted_atoms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_atoms) return;
  const int c = cell_of_atom[i];
  sorted_atoms[start[c] + atomicAdd(cursor + c, 1)] = i;
}

// The reference's criterion, bit for bit: exists vertex v with sqrt(dx^2+dy^2+dz^2) < radius.  The
// minimum over the 8 vertices is attained at the per-axis nearest vertex (rounding is monotone), so
// one distance per (cell, atom) decides.  No FMA contraction: explicit round-to-nearest intrinsics.
__device__ __forceinline__ bool atom_in_cell_support(double lx, double ly, double lz, double hx, double hy, double hz,
                                                     double X, double Y, double Z, double radius) {
  const double dx = fmin(fabs(__dsub_rn(X, lx)), fabs(__dsub_rn(X, hx)));
  const double dy = fmin(fabs(__dsub_rn(Y, ly)), fabs(__dsub_rn(Y, hy)));
  const double dz = fmin(fabs(__dsub_rn(Z, lz)), fabs(__dsub_rn(Z, hz)));
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  return __dsqrt_rn(d2) < radius;
}

// MODE 0: count hits per cell ; MODE 1: write hits at rowptr[cell] + k (hash order, sorted afterwards)
template <int MODE>
__global__ void __launch_bounds__(128) bin_cells(int n_cells, const double *__restrict__ cell_lo,
                                                 const double *__restrict__ cell_h, const double *__restrict__ pos,
                                                 HashGrid g, const int *__restrict__ hstart,
                                                 const int *__restrict__ sorted_atoms, double radius,
                                                 int64_t *__restrict__ rowptr_or_count, int *__restrict__ out_atoms) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const double lx = cell_lo[3 * c], ly = cell_lo[3 * c + 1], lz = cell_lo[3 * c + 2], hh = cell_h[c];
  const double hx = __dadd_rn(lx, hh), hy = __dadd_rn(ly, hh), hz = __dadd_rn(lz, hh);
  if (hx + radius < g.lo[0] || lx - radius > g.hi[0] || hy + radius < g.lo[1] || ly - radius > g.hi[1] ||
      hz + radius < g.lo[2] || lz - radius > g.hi[2]) {
    if (MODE == 0) rowptr_or_count[c + 1] = 0;
    return;
  }
  const int x0 = hash_coord(g, lx - radius, 0), x1 = hash_coord(g, hx + radius, 0);
  const int y0 = hash_coord(g, ly - radius, 1), y1 = hash_coord(g, hy + radius, 1);
  const int z0 = hash_coord(g, lz - radius, 2), z1 = hash_coord(g, hz + radius, 2);
  int64_t k = (MODE == 1) ? rowptr_or_count[c] : 0;
  int cnt = 0;
  for (int z = z0; z <= z1; ++z)
    for (int y = y0; y <= y1; ++y) {
      const int row = g.n[0] * (y + g.n[1] * z);
      const int a0 = hstart[row + x0], a1 = hstart[row + x1 + 1];  // contiguous run of hash cells along x
      for (int a = a0; a < a1; ++a) {
        const int i = sorted_atoms[a];
        if (atom_in_cell_support(lx, ly, lz, hx, hy, hz, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], radius)) {
          if (MODE == 1) out_atoms[k++] = i;
          ++cnt;
        }
      }
    }
  if (MODE == 0) rowptr_or_count[c + 1] = cnt;
}

// ---------------------------------------------------------------------------------------- density
// one block per cell; thread t owns quadrature point t % nq_pad and every (blockDim/nq_pad)-th atom
__global__ void __launch_bounds__(128) density_kernel(int n_cells, const double *__restrict__ cell_lo,
                                                      const double *__restrict__ cell_h,
                                                      const int *__restrict__ list_of_cell,
                                                      const int64_t *__restrict__ list_ptr,
                                                      const int *__restrict__ list_atoms, int n_atoms,
                                                      const double *__restrict__ pos, const double *__restrict__ charge,
                                                      int n_q, const double *__restrict__ qpts, double C, double inv_rc2,
                                                      double *__restrict__ rho, const int *__restrict__ cell_list) {
  extern __shared__ double sacc[];  // blockDim.x
  const int c = cell_list ? cell_list[blockIdx.x] : blockIdx.x;
  const int t = threadIdx.x;
  const int groups = max((int)blockDim.x / n_q, 1);  // threads per q-point
  const int lo_t = t % n_q, grp = t / n_q;
  const int list = list_of_cell[c];
  int64_t a0 = 0, a1 = n_atoms;
  if (list >= 0) {
    a0 = list_ptr[list];
    a1 = list_ptr[list + 1];
  }
  const double hh = cell_h[c];
  for (int q0 = 0; q0 < n_q; q0 += blockDim.x) {  // more q-points than threads: several passes
    const int q = q0 + lo_t;
    double acc = 0.0;
    if (q < n_q && grp < groups) {
      const double xq = cell_lo[3 * c] + hh * qpts[3 * q];
      const double yq = cell_lo[3 * c + 1] + hh * qpts[3 * q + 1];
      const double zq = cell_lo[3 * c + 2] + hh * qpts[3 * q + 2];
      for (int64_t a = a0 + grp; a < a1; a += groups) {
        const int i = (list >= 0) ? list_atoms[a] : (int)a;
        const double dx = pos[3 * i] - xq, dy = pos[3 * i + 1] - yq, dz = pos[3 * i + 2] - zq;
        const double r = sqrt(dx * dx + dy * dy + dz * dz);  // Point::distance, then r*r (src/step-50.cc:563-564)
        acc += C * exp(-(r * r) * inv_rc2) * charge[i];
      }
    }
    sacc[t] = acc;
    __syncthreads();
    if (grp == 0 && q < n_q) {
      double s = 0.0;
      for (int gi = 0; gi < groups; ++gi) s += sacc[gi * n_q + lo_t];
      rho[(int64_t)c * n_q + q] = s;
    }
    __syncthreads();
  }
}

// Tensor-product quadrature (the reference's QGauss<3>(n)): the Gaussian separates,
//   exp(-|X - x_q|^2 / r_c^2) = ex(qx) * ey(qy) * ez(qz),
// so an atom costs 3 n exponentials instead of n^3 (6 instead of 8 at n = 2, 15 instead of 125 at n = 5) and a
// (atom, q-point) pair three shared-memory reads and three multiplies.  The factors agree with the unseparated
// exponential to a few ulp: far inside the 1e-12 relative L2 the RHS has to meet.  One block per cell; atoms in
// tiles of DENS_TILE: phase A fills the factor tables, phase B accumulates (q-point, atom-group) partial sums.
constexpr int DENS_TILE = 32, DENS_MAXQ1 = 8, DENS_PASSES = 4;
struct TensorRule {
  int n1;
  double p[3][DENS_MAXQ1];
};
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) density_sep_kernel(int n_cells, const double *__restrict__ cell_lo,
                                                            const double *__restrict__ cell_h, const int *__restrict__ list_of_cell,
                                                            const int64_t *__restrict__ list_ptr, const int *__restrict__ list_atoms,
                                                            int n_atoms, const double *__restrict__ pos,
                                                            const double *__restrict__ charge, const __grid_constant__ TensorRule R,
                                                            double C, double inv_rc2, double *__restrict__ rho,
                                                            const int *__restrict__ cell_list) {
  __shared__ double tab[3][DENS_TILE][DENS_MAXQ1];
  __shared__ double sacc[BLOCK];
  const int c = cell_list ? cell_list[blockIdx.x] : blockIdx.x;
  const int t = threadIdx.x;
  const int n1 = R.n1, n_q = n1 * n1 * n1;
  const int list = list_of_cell[c];
  int64_t a0 = 0, a1 = n_atoms;
  if (list >= 0) {
    a0 = list_ptr[list];
    a1 = list_ptr[list + 1];
  }
  if (a0 == a1) {  // vacuum cell
    for (int q = t; q < n_q; q += BLOCK) rho[(int64_t)c * n_q + q] = 0.0;
    return;
  }
  const double hh = cell_h[c];
  const double lo[3] = {cell_lo[3 * c], cell_lo[3 * c + 1], cell_lo[3 * c + 2]};
  const int per_pass = min(n_q, BLOCK);
  const int groups = max(BLOCK / n_q, 1);  // threads per q-point (n_q < BLOCK)
  const int lo_t = t % per_pass, grp = t / per_pass;
  double acc[DENS_PASSES];
  int qx[DENS_PASSES], qy[DENS_PASSES], qz[DENS_PASSES];
#pragma unroll
  for (int ps = 0; ps < DENS_PASSES; ++ps) {
    acc[ps] = 0.0;
    const int q = ps * BLOCK + lo_t;
    qx[ps] = q % n1;
    qy[ps] = (q / n1) % n1;
    qz[ps] = q / (n1 * n1);
  }
  for (int64_t base = a0; base < a1; base += DENS_TILE) {
    const int m = (int)min((int64_t)DENS_TILE, a1 - base);
    __syncthreads();
    for (int e = t; e < m * 3 * n1; e += BLOCK) {
      const int a = e / (3 * n1), rem = e - a * 3 * n1, k = rem / n1, j = rem - k * n1;
      const int i = (list >= 0) ? list_atoms[base + a] : (int)(base + a);
      const double d = pos[3 * i + k] - (lo[k] + hh * R.p[k][j]);
      double v = exp(-(d * d) * inv_rc2);
      if (k == 0) v = C * v * charge[i];
      tab[k][a][j] = v;
    }
    __syncthreads();
    if (grp < groups) {
#pragma unroll
      for (int ps = 0; ps < DENS_PASSES; ++ps)
        if (ps * BLOCK + lo_t < n_q)
          for (int a = grp; a < m; a += groups) acc[ps] += tab[0][a][qx[ps]] * tab[1][a][qy[ps]] * tab[2][a][qz[ps]];
    }
  }
#pragma unroll
  for (int ps = 0; ps < DENS_PASSES; ++ps) {
    if (ps * BLOCK >= n_q) break;
    __syncthreads();
    sacc[t] = acc[ps];
    __syncthreads();
    const int q = ps * BLOCK + lo_t;
    if (grp == 0 && q < n_q) {
      double s = 0.0;
      for (int gi = 0; gi < groups; ++gi) s += sacc[gi * per_pass + lo_t];
      rho[(int64_t)c * n_q + q] = s;
    }
  }
}

// cells whose atom list is not empty (the others are vacuum: their densities are 0)
__global__ void cell_has_atoms(int n_cells, const int *__restrict__ list_of_cell, const int64_t *__restrict__ list_ptr, int n_atoms,
                               unsigned char *__restrict__ flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int list = list_of_cell[c];
  flag[c] = (list >= 0) ? (list_ptr[list + 1] > list_ptr[list]) : (n_atoms > 0);
}

// Small rules (n <= 3 points per axis, the cluster runs use n = 2): one thread per atom, the 3 n factors and the n^3
// accumulators in registers, no shared-memory table and no barrier in the atom loop; one shuffle/shared-memory
// reduction per cell at the end.
template <int N1>
__global__ void __launch_bounds__(64, N1 == 2 ? 14 : 8) density_reg_kernel(int n_cells, const double *__restrict__ cell_lo,
                                                         const double *__restrict__ cell_h, const int *__restrict__ list_of_cell,
                                                         const int64_t *__restrict__ list_ptr, const int *__restrict__ list_atoms,
                                                         int n_atoms, const double *__restrict__ pos,
                                                         const double *__restrict__ charge, const __grid_constant__ TensorRule R,
                                                         double C, double inv_rc2, double *__restrict__ rho,
                                                         const int *__restrict__ cell_list) {
  constexpr int NQ = N1 * N1 * N1;
  __shared__ double part[2][NQ];
  const int c = cell_list ? cell_list[blockIdx.x] : blockIdx.x;
  const int t = threadIdx.x;
  const int list = list_of_cell[c];
  int64_t a0 = 0, a1 = n_atoms;
  if (list >= 0) {
    a0 = list_ptr[list];
    a1 = list_ptr[list + 1];
  }
  if (a0 == a1) {  // vacuum cell
    if (t < NQ) rho[(int64_t)c * NQ + t] = 0.0;
    return;
  }
  const double hh = cell_h[c];
  double q0[3][N1];  // quadrature point coordinates per axis
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < N1; ++j) q0[k][j] = cell_lo[3 * c + k] + hh * R.p[k][j];
  double acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
  for (int64_t a = a0 + t; a < a1; a += 64) {
    const int i = (list >= 0) ? list_atoms[a] : (int)a;
    double e[3][N1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double x = pos[3 * i + k];
#pragma unroll
      for (int j = 0; j < N1; ++j) {
        const double d = x - q0[k][j];
        e[k][j] = exp(-(d * d) * inv_rc2);
      }
    }
    const double w = C * charge[i];
#pragma unroll
    for (int j = 0; j < N1; ++j) e[0][j] *= w;
#pragma unroll
    for (int z = 0; z < N1; ++z)
#pragma unroll
      for (int y = 0; y < N1; ++y) {
        const double yz = e[1][y] * e[2][z];
#pragma unroll
        for (int x = 0; x < N1; ++x) acc[(z * N1 + y) * N1 + x] += e[0][x] * yz;
      }
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((t & 31) == 0) part[t >> 5][q] = v;
  }
  __syncthreads();
  if (t < NQ) rho[(int64_t)c * NQ + t] = part[0][t] + part[1][t];
}

// ------------------------------------------------------------------------------------ load vector
__global__ void __launch_bounds__(128) load_vector_kernel(int n_cells, const double *__restrict__ rho,
                                                          const double *__restrict__ cell_h,
                                                          const int *__restrict__ cell_dofs, int n_q,
                                                          const double *__restrict__ shape,
                                                          const double *__restrict__ weights,
                                                          const double *__restrict__ kref, const double *__restrict__ ghat,
                                                          const int64_t *__restrict__ hang_ptr,
                                                          const int *__restrict__ hang_col,
                                                          const double *__restrict__ hang_val,
                                                          const uint8_t *__restrict__ constrained, double *b,
                                                          const int *__restrict__ cell_list) {
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  if (ci >= n_cells) return;
  const int c = cell_list ? cell_list[ci] : ci;
  const double hh = cell_h[c];
  const double jac = hh * hh * hh;
  double f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool any = false;
  for (int q = 0; q < n_q; ++q) {
    const double rw = rho[(int64_t)c * n_q + q] * (weights[q] * jac);
    if (rw != 0.0) any = true;
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += shape[q * 8 + i] * rw;
  }
  int dofs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dofs[i] = cell_dofs[(int64_t)c * 8 + i];
  if (kref != nullptr && ghat != nullptr) {
    double gl[8];
    bool anyg = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gl[j] = ghat[dofs[j]];
      anyg |= gl[j] != 0.0;
    }
    if (anyg) {
      any = true;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += kref[i * 8 + j] * gl[j];
        f[i] -= hh * s;
      }
    }
  }
  if (!any) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int dof = dofs[i];
    const int64_t p0 = hang_ptr[dof], p1 = hang_ptr[dof + 1];
    if (p1 > p0) {
      for (int64_t p = p0; p < p1; ++p) {
        const int par = hang_col[p];
        if (!constrained[par]) atomicAdd(b + par, hang_val[p] * f[i]);
      }
    } else if (!constrained[dof]) {
      atomicAdd(b + dof, f[i]);
    }
  }
}

__global__ void point_values_kernel(int n, const int *__restrict__ cell_dofs, const double *__restrict__ xi,
                                    const double *__restrict__ u, double *__restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double x = xi[3 * p], y = xi[3 * p + 1], z = xi[3 * p + 2];
  double s = 0.0;
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const double w = ((v & 1) ? x : 1.0 - x) * ((v & 2) ? y : 1.0 - y) * ((v & 4) ? z : 1.0 - z);
    s += w * u[cell_dofs[8 * p + v]];
  }
  out[p] = s;
}

// a cell contributes to this rank's rows of b if one of its dofs is owned here, or is a hanging node with a
// parent owned here (its share is redistributed to the parents)
__global__ void select_owned_cells(int n_cells, const int *__restrict__ cell_dofs, const int *__restrict__ owner, int rank,
                                   const int64_t *__restrict__ hang_ptr, const int *__restrict__ hang_col,
                                   int *__restrict__ list, int *__restrict__ count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  bool mine = false;
  for (int v = 0; v < 8; ++v) {
    const int dof = cell_dofs[(int64_t)c * 8 + v];
    mine |= owner[dof] == rank;
    for (int64_t p = hang_ptr[dof]; p < hang_ptr[dof + 1]; ++p) mine |= owner[hang_col[p]] == rank;
  }
  if (mine) list[atomicAdd(count, 1)] = c;
}

int run_density(gmg_context *h, RhsState *s, const int *cell_list = nullptr, int n_list = 0) {
  if (s->n_cells == 0) return GMG_OK;
  const double C = 4.0 * M_PI / (s->r_c * s->r_c * s->r_c * std::pow(M_PI, 1.5));  // src/step-50.cc:522
  const int block = (s->n_q <= 64) ? 64 : 128;  // threads per cell: blockDim / n_q threads share a q-point
  if (!cell_list && s->n_nonempty >= 0) {  // vacuum cells: zero by memset, no block
    GMG_CUDA(h, cudaMemsetAsync(s->rho, 0, sizeof(double) * (size_t)s->n_cells * s->n_q, h->stream));
    cell_list = s->nonempty;
    n_list = s->n_nonempty;
  }
  const int grid = cell_list ? n_list : s->n_cells;
  if (grid == 0) return GMG_OK;
  if (s->tensor_n1 > 0) {  // tensor-product rule: separable Gaussians
    TensorRule R;
    R.n1 = s->tensor_n1;
    for (int k = 0; k < 3; ++k)
      for (int j = 0; j < DENS_MAXQ1; ++j) R.p[k][j] = j < R.n1 ? s->tensor_p[k][j] : 0.0;
    const double irc2 = 1.0 / (s->r_c * s->r_c);
#define GMG_DENSITY_REG(N1)                                                                                              \
  density_reg_kernel<N1><<<grid, 64, 0, h->stream>>>(s->n_cells, s->cell_lo, s->cell_h, s->list_of_cell, h->list_ptr, \
                                                     h->list_atoms, h->n_atoms, h->atom_pos, h->atom_q, R, C, irc2, s->rho, cell_list)
    if (R.n1 == 1) GMG_DENSITY_REG(1);
    else if (R.n1 == 2) GMG_DENSITY_REG(2);
    else if (R.n1 == 3) GMG_DENSITY_REG(3);
#undef GMG_DENSITY_REG
    else if (s->n_q <= 64)
      density_sep_kernel<64><<<grid, 64, 0, h->stream>>>(s->n_cells, s->cell_lo, s->cell_h, s->list_of_cell, h->list_ptr,
                                                          h->list_atoms, h->n_atoms, h->atom_pos, h->atom_q, R, C,
                                                          1.0 / (s->r_c * s->r_c), s->rho, cell_list);
    else
      density_sep_kernel<128><<<grid, 128, 0, h->stream>>>(s->n_cells, s->cell_lo, s->cell_h, s->list_of_cell, h->list_ptr,
                                                            h->list_atoms, h->n_atoms, h->atom_pos, h->atom_q, R, C,
                                                            1.0 / (s->r_c * s->r_c), s->rho, cell_list);
    GMG_LAUNCH_CHECK(h);
    return GMG_OK;
  }
  density_kernel<<<grid, block, sizeof(double) * block, h->stream>>>(
      s->n_cells, s->cell_lo, s->cell_h, s->list_of_cell, h->list_ptr, h->list_atoms, h->n_atoms, h->atom_pos, h->atom_q,
      s->n_q, s->qpts, C, 1.0 / (s->r_c * s->r_c), s->rho, cell_list);
  GMG_LAUNCH_CHECK(h);
  return GMG_OK;
}

int run_load_vector(gmg_context *h, RhsState *s, const double *rho_dev, double *b_dev, const int *cell_list = nullptr,
                    int n_list = 0) {
  GMG_CUDA(h, cudaMemsetAsync(b_dev, 0, sizeof(double) * s->n_dofs, h->stream));
  const int n = cell_list ? n_list : s->a_cells;
  if (n == 0) return GMG_OK;
  load_vector_kernel<<<cdiv(n, 128), 128, 0, h->stream>>>(
      n, rho_dev, s->a_h, s->cell_dofs, s->a_nq, s->shape, s->weights, s->have_kref ? s->kref : nullptr,
      s->have_ghat ? s->ghat : nullptr, s->hang_ptr, s->hang_col, s->hang_val, s->constrained, b_dev, cell_list);
  GMG_LAUNCH_CHECK(h);
  return GMG_OK;
}

}  // namespace

namespace gmg {
void rhs_invalidate_partition(gmg_context *h) {
  if (h->rhs) h->rhs->n_dist_cells = -1;
}
void rhs_free(gmg_context *h) {
  RhsState *s = h->rhs;
  if (!s) return;
  dfree(s->cell_lo);
  dfree(s->cell_h);
  dfree(s->qpts);
  dfree(s->rho);
  dfree(s->nonempty);
  dfree(s->list_of_cell);
  dfree(s->a_h);
  dfree(s->shape);
  dfree(s->weights);
  dfree(s->kref);
  dfree(s->ghat);
  dfree(s->hang_val);
  dfree(s->cell_dofs);
  dfree(s->hang_col);
  dfree(s->hang_ptr);
  dfree(s->constrained);
  dfree(s->dist_cells);
  delete s;
  h->rhs = nullptr;
}
}  // namespace gmg

// the cell arrays kept on the device by the last gmg_assemble_rhs / gmg_charge_density call (indicator.cu)
int gmg_rhs_resident(gmg_context *h, int *n_cells, const double **cell_h, const int **cell_dofs, const double **weights, int *n_q,
                     const double **rho, int *rho_cells, int *rho_nq) {
  RhsState *s = state(h);
  *n_cells = s->a_cells;
  *cell_h = s->a_h;
  *cell_dofs = s->cell_dofs;
  *weights = s->weights;
  *n_q = s->a_nq;
  *rho = s->rho;
  *rho_cells = s->n_cells;
  *rho_nq = s->n_q;
  return GMG_OK;
}

extern "C" {

int gmg_set_atoms(gmg_handle h, int32_t n_atoms, const double *pos, const double *charge) {
  if (!h || n_atoms < 0 || (n_atoms && (!pos || !charge))) return GMG_EINVAL;
  gmg::enter(h);
  h->n_atoms = n_atoms;
  state(h)->n_nonempty = -1;  // (the list of non-vacuum cells follows the atom lists)
  if (int rc = upload(h, h->atom_pos, pos, 3 * (int64_t)n_atoms)) return rc;
  if (int rc = upload(h, h->atom_q, charge, n_atoms)) return rc;
  GMG_CUDA(h, cudaStreamSynchronize(h->stream));
  return GMG_OK;
}

int gmg_set_atom_lists(gmg_handle h, int32_t n_lists, const int64_t *rowptr, const int32_t *atoms) {
  if (!h || n_lists < 0 || !rowptr) return GMG_EINVAL;
  gmg::enter(h);
  h->n_lists = n_lists;
  state(h)->n_nonempty = -1;
  if (int rc = upload(h, h->list_ptr, rowptr, (int64_t)n_lists + 1)) return rc;
  if (int rc = upload(h, h->list_atoms, atoms, rowptr[n_lists])) return rc;
  GMG_CUDA(h, cudaStreamSynchronize(h->stream));
  return GMG_OK;
}

int gmg_bin_atoms(gmg_handle h, int32_t n_cells, const double *cell_lo, const double *cell_h, int32_t n_atoms,
                  const double *pos, double radius, int64_t *rowptr_out, int32_t *atoms_out) {
  if (!h || n_cells < 0 || n_atoms < 0 || !rowptr_out || radius <= 0.0) return GMG_EINVAL;
  gmg::enter(h);
  if (atoms_out != nullptr && state(h)->bin_pending && h->n_lists == n_cells && h->list_ptr && h->list_atoms) {
    state(h)->bin_pending = false;
    // second call of the two-call protocol: hand out the lists computed by the first
    GMG_CUDA(h, gmg::copy_sync(h, rowptr_out, h->list_ptr, sizeof(int64_t) * (n_cells + 1), cudaMemcpyDeviceToHost));
    if (rowptr_out[n_cells] > 0)
      GMG_CUDA(h, gmg::copy_sync(h, atoms_out, h->list_atoms, sizeof(int) * rowptr_out[n_cells], cudaMemcpyDeviceToHost));
    return GMG_OK;
  }
  // hash grid over the atoms' bounding box, about one atom per hash cell
  HashGrid g;
  double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
  if (n_atoms > 0) {
    for (int d = 0; d < 3; ++d) lo[d] = hi[d] = pos[d];
    for (int i = 0; i < n_atoms; ++i)
      for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(lo[d], pos[3 * i + d]);
        hi[d] = std::max(hi[d], pos[3 * i + d]);
      }
  }
  double vol = 1.0;
  for (int d = 0; d < 3; ++d) vol *= std::max(hi[d] - lo[d], 1e-3);
  double s = std::cbrt(vol / std::max(n_atoms, 1));
  s = std::min(std::max(s, radius / 8.0), radius);
  int64_t total = 1;
  for (int d = 0; d < 3; ++d) {
    g.lo[d] = lo[d];
    g.hi[d] = hi[d];
    g.n[d] = std::max(1, (int)std::floor((hi[d] - lo[d]) / s) + 1);
    total *= g.n[d];
  }
  if (total > (int64_t)1 << 28) return fail(h, GMG_EINVAL, "atom hash grid too large");
  g.inv_s = 1.0 / s;
  const int nh = (int)total;
  double *d_pos = nullptr, *d_lo = nullptr, *d_h = nullptr;
  int *d_cell_of_atom = nullptr, *d_count = nullptr, *d_start = nullptr, *d_sorted = nullptr;
  int64_t *d_rowptr = nullptr;
  int *d_atoms = nullptr, *d_atoms_sorted = nullptr;
  void *d_temp = nullptr;
  int rc = GMG_OK;
  auto cleanup = [&]() {
    dfree(d_pos); dfree(d_lo); dfree(d_h); dfree(d_cell_of_atom); dfree(d_count); dfree(d_start); dfree(d_sorted);
    dfree(d_atoms);
    if (d_temp) cudaFreeAsync(d_temp, tl_stream);
  };
#define TRY(expr)                                                                         \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      cleanup();                                                                          \
      dfree(d_rowptr);                                                                    \
      dfree(d_atoms_sorted);                                                              \
      return fail(h, GMG_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));     \
    }                                                                                     \
  } while (0)
  TRY(dalloc(&d_pos, 3 * (int64_t)n_atoms));
  TRY(dalloc(&d_lo, 3 * (int64_t)n_cells));
  TRY(dalloc(&d_h, n_cells));
  TRY(dalloc(&d_cell_of_atom, n_atoms));
  TRY(dalloc(&d_count, nh + 1));
  TRY(dalloc(&d_start, nh + 1));
  TRY(dalloc(&d_sorted, n_atoms));
  TRY(dalloc(&d_rowptr, (int64_t)n_cells + 1));
  TRY(gmg::copy(h, d_pos, pos, sizeof(double) * 3 * n_atoms, cudaMemcpyHostToDevice));
  TRY(gmg::copy(h, d_lo, cell_lo, sizeof(double) * 3 * n_cells, cudaMemcpyHostToDevice));
  TRY(gmg::copy(h, d_h, cell_h, sizeof(double) * n_cells, cudaMemcpyHostToDevice));
  TRY(cudaMemsetAsync(d_count, 0, sizeof(int) * (nh + 1), h->stream));
  if (n_atoms > 0) {
    hash_count<<<cdiv(n_atoms, 256), 256, 0, h->stream>>>(n_atoms, d_pos, g, d_cell_of_atom, d_count);
    h->launches++;
  }
  // exclusive scan of the (small) hash histogram on the host
  std::vector<int> cnt(nh + 1), start(nh + 1, 0);
  TRY(gmg::copy(h, cnt.data(), d_count, sizeof(int) * (nh + 1), cudaMemcpyDeviceToHost));
  TRY(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < nh; ++i) start[i + 1] = start[i] + cnt[i];
  TRY(gmg::copy(h, d_start, start.data(), sizeof(int) * (nh + 1), cudaMemcpyHostToDevice));
  TRY(cudaMemsetAsync(d_count, 0, sizeof(int) * (nh + 1), h->stream));
  if (n_atoms > 0) {
    hash_fill<<<cdiv(n_atoms, 256), 256, 0, h->stream>>>(n_atoms, d_cell_of_atom, d_start, d_count, d_sorted);
    h->launches++;
  }
  // pass 1: counts
  TRY(cudaMemsetAsync(d_rowptr, 0, sizeof(int64_t) * (n_cells + 1), h->stream));
  if (n_cells > 0) {
    bin_cells<0><<<cdiv(n_cells, 128), 128, 0, h->stream>>>(n_cells, d_lo, d_h, d_pos, g, d_start, d_sorted, radius,
                                                            d_rowptr, nullptr);
    h->launches++;
  }
  std::vector<int64_t> rp(n_cells + 1);
  TRY(gmg::copy(h, rp.data(), d_rowptr, sizeof(int64_t) * (n_cells + 1), cudaMemcpyDeviceToHost));
  TRY(cudaStreamSynchronize(h->stream));
  for (int c = 0; c < n_cells; ++c) rp[c + 1] += rp[c];
  const int64_t n_pairs = rp[n_cells];
  TRY(gmg::copy(h, d_rowptr, rp.data(), sizeof(int64_t) * (n_cells + 1), cudaMemcpyHostToDevice));
  TRY(dalloc(&d_atoms, n_pairs));
  TRY(dalloc(&d_atoms_sorted, n_pairs));
  if (n_cells > 0 && n_pairs > 0) {
    bin_cells<1><<<cdiv(n_cells, 128), 128, 0, h->stream>>>(n_cells, d_lo, d_h, d_pos, g, d_start, d_sorted, radius,
                                                            d_rowptr, d_atoms);
    h->launches++;
    // ascending atom index inside every cell (std::set iteration order of the reference)
    size_t temp_bytes = 0;
    TRY(cub::DeviceSegmentedSort::SortKeys(nullptr, temp_bytes, d_atoms, d_atoms_sorted, n_pairs, n_cells, d_rowptr,
                                           d_rowptr + 1, h->stream));
    TRY(cudaMallocAsync(&d_temp, std::max<size_t>(temp_bytes, 1), tl_stream));
    TRY(cub::DeviceSegmentedSort::SortKeys(d_temp, temp_bytes, d_atoms, d_atoms_sorted, n_pairs, n_cells, d_rowptr,
                                           d_rowptr + 1, h->stream));
    h->launches++;
  }
  TRY(cudaStreamSynchronize(h->stream));
  TRY(cudaGetLastError());
#undef TRY
  cleanup();
  // keep the lists on the device as the current atom lists
  dfree(h->list_ptr);
  dfree(h->list_atoms);
  h->list_ptr = d_rowptr;
  h->list_atoms = d_atoms_sorted;
  h->n_lists = n_cells;
  state(h)->n_nonempty = -1;
  std::copy(rp.begin(), rp.end(), rowptr_out);
  state(h)->bin_pending = (atoms_out == nullptr);
  if (atoms_out && n_pairs > 0)
    GMG_CUDA(h, gmg::copy_sync(h, atoms_out, h->list_atoms, sizeof(int) * n_pairs, cudaMemcpyDeviceToHost));
  return rc;
}

int gmg_charge_density(gmg_handle h, int32_t n_cells, const double *cell_lo, const double *cell_h,
                       const int32_t *list_of_cell, int32_t n_q, const double *qpoints, double r_c, double *rho_out) {
  if (!h || n_cells < 0 || n_q < 1 || !cell_lo || !cell_h || !list_of_cell || !qpoints) return GMG_EINVAL;
  if (!h->atom_pos) return fail(h, GMG_EINVAL, "gmg_set_atoms first");
  gmg::enter(h);
  RhsState *s = state(h);
  s->n_cells = n_cells;
  s->n_q = n_q;
  s->r_c = r_c;
  s->n_dist_cells = -1;
  if (int rc = upload(h, s->cell_lo, cell_lo, 3 * (int64_t)n_cells)) return rc;
  if (int rc = upload(h, s->cell_h, cell_h, n_cells)) return rc;
  if (int rc = upload(h, s->list_of_cell, list_of_cell, n_cells)) return rc;
  if (int rc = upload(h, s->qpts, qpoints, 3 * (int64_t)n_q)) return rc;
  // tensor-product rule (x fastest)?  Then the Gaussians separate (density_sep_kernel)
  s->tensor_n1 = 0;
  {
    int n1 = (int)std::lround(std::cbrt((double)n_q));
    if (n1 >= 1 && n1 <= DENS_MAXQ1 && n1 * n1 * n1 == n_q && n_q <= DENS_PASSES * 128) {
      bool ok = true;
      for (int q = 0; q < n_q && ok; ++q) {
        const int j[3] = {q % n1, (q / n1) % n1, q / (n1 * n1)};
        const int first[3] = {j[0], j[1] * n1, j[2] * n1 * n1};  // the first point with this coordinate index
        for (int k = 0; k < 3; ++k) ok = ok && qpoints[3 * q + k] == qpoints[3 * first[k] + k];
      }
      if (ok) {
        s->tensor_n1 = n1;
        for (int j = 0; j < n1; ++j) {
          s->tensor_p[0][j] = qpoints[3 * j];
          s->tensor_p[1][j] = qpoints[3 * (j * n1) + 1];
          s->tensor_p[2][j] = qpoints[3 * (j * n1 * n1) + 2];
        }
      }
    }
  }
  dfree(s->rho);
  GMG_CUDA(h, dalloc(&s->rho, (int64_t)n_cells * n_q));
  // compact list of the cells that have atoms
  s->n_nonempty = -1;
  dfree(s->nonempty);
  if (n_cells > 0) {
    unsigned char *flag = nullptr;
    int *d_n = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    GMG_CUDA(h, dalloc(&flag, n_cells));
    GMG_CUDA(h, dalloc(&d_n, 1));
    GMG_CUDA(h, dalloc(&s->nonempty, n_cells));
    cell_has_atoms<<<cdiv(n_cells, 256), 256, 0, h->stream>>>(n_cells, s->list_of_cell, h->list_ptr, h->n_atoms, flag);
    h->launches++;
    thrust::counting_iterator<int> iota(0);
    GMG_CUDA(h, cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, flag, s->nonempty, d_n, n_cells, h->stream));
    GMG_CUDA(h, cudaMallocAsync(&tmp, std::max<size_t>(tmp_bytes, 1), h->stream));
    GMG_CUDA(h, cub::DeviceSelect::Flagged(tmp, tmp_bytes, iota, flag, s->nonempty, d_n, n_cells, h->stream));
    int hn = 0;
    GMG_CUDA(h, gmg::copy_sync(h, &hn, d_n, sizeof(int), cudaMemcpyDeviceToHost));
    s->n_nonempty = hn;
    cudaFreeAsync(tmp, h->stream);
    dfree(flag);
    dfree(d_n);
  }
  if (int rc = run_density(h, s)) return rc;
  if (rho_out)
    if (int rc = gmg::staged_d2h(h, rho_out, s->rho, sizeof(double) * (int64_t)n_cells * n_q)) return rc;
  GMG_CUDA(h, cudaStreamSynchronize(h->stream));
  return GMG_OK;
}

int gmg_assemble_rhs(gmg_handle h, int32_t n_cells, const double *rho, const double *cell_h, const int32_t *cell_dofs,
                     int32_t n_q, const double *shape, const double *weights, const double *Kref, const double *ghat,
                     int32_t n_dofs, const int64_t *hang_rowptr, const int32_t *hang_col, const double *hang_val,
                     const uint8_t *constrained, double *b_out) {
  if (!h || n_cells < 0 || !cell_h || !cell_dofs || !shape || !weights || !hang_rowptr || !constrained || !b_out)
    return GMG_EINVAL;
  gmg::enter(h);
  RhsState *s = state(h);
  if (rho == nullptr && (s->rho == nullptr || s->n_cells != n_cells || s->n_q != n_q))
    return fail(h, GMG_EINVAL, "rho == NULL needs a matching gmg_charge_density call first");
  s->a_cells = n_cells;
  s->a_nq = n_q;
  s->n_dofs = n_dofs;
  s->n_dist_cells = -1;
  if (int rc = upload(h, s->a_h, cell_h, n_cells)) return rc;
  if (int rc = upload(h, s->cell_dofs, cell_dofs, 8 * (int64_t)n_cells)) return rc;
  if (int rc = upload(h, s->shape, shape, 8 * (int64_t)n_q)) return rc;
  if (int rc = upload(h, s->weights, weights, n_q)) return rc;
  s->have_kref = Kref != nullptr;
  s->have_ghat = ghat != nullptr;
  if (Kref)
    if (int rc = upload(h, s->kref, Kref, 64)) return rc;
  if (ghat)
    if (int rc = upload(h, s->ghat, ghat, n_dofs)) return rc;
  if (int rc = upload(h, s->hang_ptr, hang_rowptr, (int64_t)n_dofs + 1)) return rc;
  if (int rc = upload(h, s->hang_col, hang_col, hang_rowptr[n_dofs])) return rc;
  if (int rc = upload(h, s->hang_val, hang_val, hang_rowptr[n_dofs])) return rc;
  if (int rc = upload(h, s->constrained, constrained, n_dofs)) return rc;
  const double *rho_dev = s->rho;
  if (rho != nullptr) {
    if (int rc = ensure_stage(h, std::max<int64_t>((int64_t)n_cells * n_q, n_dofs))) return rc;
    GMG_CUDA(h, gmg::copy(h, h->stage_a, rho, sizeof(double) * (int64_t)n_cells * n_q, cudaMemcpyHostToDevice));
    rho_dev = h->stage_a;
  } else if (int rc = ensure_stage(h, n_dofs)) {
    return rc;
  }
  if (int rc = run_load_vector(h, s, rho_dev, h->stage_b)) return rc;
  GMG_CUDA(h, gmg::copy(h, b_out, h->stage_b, sizeof(double) * n_dofs, cudaMemcpyDeviceToHost));
  GMG_CUDA(h, cudaStreamSynchronize(h->stream));
  return GMG_OK;
}

int gmg_rhs_step_dev(gmg_handle h, double *b_dev) {
  if (!h || !h->rhs || !b_dev) return GMG_EINVAL;
  gmg::enter(h);
  RhsState *s = h->rhs;
  if (s->n_cells != s->a_cells || s->n_q != s->a_nq) return fail(h, GMG_EINVAL, "density / load-vector inputs differ");
  if (h->dist.on && (int)h->dist.sys_owner.size() == s->n_dofs) {
    // multi-GPU: a rank's rows of b depend only on the cells touching its dofs; the solve reads only those rows
    if (s->n_dist_cells < 0) {
      int *owner = nullptr, *count = nullptr;
      dfree(s->dist_cells);
      GMG_CUDA(h, dalloc(&owner, s->n_dofs));
      GMG_CUDA(h, dalloc(&count, 1));
      GMG_CUDA(h, dalloc(&s->dist_cells, s->a_cells));
      GMG_CUDA(h, gmg::copy(h, owner, h->dist.sys_owner.data(), sizeof(int) * s->n_dofs, cudaMemcpyHostToDevice));
      GMG_CUDA(h, cudaMemsetAsync(count, 0, sizeof(int), h->stream));
      select_owned_cells<<<cdiv(std::max(s->a_cells, 1), 256), 256, 0, h->stream>>>(
          s->a_cells, s->cell_dofs, owner, h->dist.rank, s->hang_ptr, s->hang_col, s->dist_cells, count);
      h->launches++;
      GMG_CUDA(h, gmg::copy_sync(h, &s->n_dist_cells, count, sizeof(int), cudaMemcpyDeviceToHost));
      dfree(owner);
      dfree(count);
    }
    if (int rc = run_density(h, s, s->dist_cells, s->n_dist_cells)) return rc;
    return run_load_vector(h, s, s->rho, b_dev, s->dist_cells, s->n_dist_cells);
  }
  if (int rc = run_density(h, s)) return rc;
  return run_load_vector(h, s, s->rho, b_dev);
}

int gmg_point_values(gmg_handle h, int32_t n_points, const int32_t *cell_dofs, const double *ref_coords, const double *u,
                     int32_t n_dofs, double *phi_out) {
  if (!h || n_points < 0 || !cell_dofs || !ref_coords || !u || !phi_out) return GMG_EINVAL;
  gmg::enter(h);
  int *d_dofs = nullptr;
  double *d_xi = nullptr, *d_u = nullptr, *d_out = nullptr;
  int rc = GMG_OK;
  if ((rc = upload(h, d_dofs, cell_dofs, 8 * (int64_t)n_points)) == GMG_OK &&
      (rc = upload(h, d_xi, ref_coords, 3 * (int64_t)n_points)) == GMG_OK && (rc = upload(h, d_u, u, n_dofs)) == GMG_OK &&
      dalloc(&d_out, n_points) == cudaSuccess) {
    if (n_points > 0) {
      point_values_kernel<<<cdiv(n_points, 128), 128, 0, h->stream>>>(n_points, d_dofs, d_xi, d_u, d_out);
      h->launches++;
    }
    gmg::copy(h, phi_out, d_out, sizeof(double) * n_points, cudaMemcpyDeviceToHost);
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) rc = fail(h, GMG_ECUDA, "point values failed");
  } else if (rc == GMG_OK) {
    rc = fail(h, GMG_ECUDA, "allocation failed");
  }
  dfree(d_dofs);
  dfree(d_xi);
  dfree(d_u);
  dfree(d_out);
  return rc;
}

int gmg_pair_energies(gmg_handle h, double r_c, double out[2]) {
  if (!h || !out || !(r_c > 0.0)) return GMG_EINVAL;
  if (!h->atom_pos) return fail(h, GMG_EINVAL, "gmg_set_atoms first");
  gmg::enter(h);
  const int n = h->n_atoms;
  out[0] = out[1] = 0.0;
  if (n < 2) return GMG_OK;
  const int grid = cdiv(n, PAIR_BLOCK);
  double *partial = nullptr;
  GMG_CUDA(h, dalloc(&partial, 2 * (int64_t)grid));
  pair_energy_kernel<<<grid, PAIR_BLOCK, 0, h->stream>>>(n, h->atom_pos, h->atom_q, 1.0 / r_c, partial);
  h->launches++;
  std::vector<double> hp(2 * (size_t)grid);
  const cudaError_t e = gmg::copy_sync(h, hp.data(), partial, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost);
  dfree(partial);
  if (e != cudaSuccess) return fail(h, GMG_ECUDA, std::string("pair energies: ") + cudaGetErrorString(e));
  for (int b = 0; b < grid; ++b) {
    out[0] += hp[b];
    out[1] += hp[grid + b];
  }
  return GMG_OK;
}

int gmg_energy_norm_error(gmg_handle h, const double *u, int32_t n_dofs, double r_c, const double gauss2_points[2],
                          const double gauss2_weights[2], double *out) {
  if (!h || !u || !out || !gauss2_points || !gauss2_weights) return GMG_EINVAL;
  gmg::enter(h);
  RhsState *s = h->rhs;
  if (!s || !s->cell_lo || !s->cell_dofs || s->n_cells != s->a_cells || n_dofs != s->n_dofs)
    return fail(h, GMG_EINVAL, "energy norm: the cells of gmg_charge_density / gmg_assemble_rhs are not resident");
  *out = 0.0;
  const int64_t threads = 8 * (int64_t)s->n_cells;
  if (threads == 0) return GMG_OK;
  const int grid = cdiv(threads, ENORM_BLOCK);
  double *d_u = nullptr, *partial = nullptr;
  int rc = upload(h, d_u, u, (int64_t)n_dofs);
  if (rc == GMG_OK && dalloc(&partial, grid) != cudaSuccess) rc = fail(h, GMG_ECUDA, "allocation failed");
  if (rc == GMG_OK) {
    energy_norm_kernel<<<grid, ENORM_BLOCK, 0, h->stream>>>(s->n_cells, s->cell_lo, s->cell_h, s->cell_dofs, d_u, h->n_atoms,
                                                            h->atom_pos, h->atom_q, r_c, gauss2_points[0], gauss2_points[1],
                                                            gauss2_weights[0], gauss2_weights[1], partial);
    h->launches++;
    std::vector<double> hp((size_t)grid);
    const cudaError_t e = gmg::copy_sync(h, hp.data(), partial, sizeof(double) * hp.size(), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(h, GMG_ECUDA, std::string("energy norm: ") + cudaGetErrorString(e));
    double sum = 0.0;
    for (int b = 0; b < grid; ++b) sum += hp[b];
    *out = std::sqrt(sum);
  }
  dfree(d_u);
  dfree(partial);
  return rc;
}

}  // extern "C"
