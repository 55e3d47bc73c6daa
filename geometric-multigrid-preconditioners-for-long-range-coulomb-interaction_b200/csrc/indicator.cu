// gmg_b200 C ABI, refinement indicator: the cell-wise error estimate of estimate_error_and_mark_cells
// (src/step-50.cc:1020-1090): KellyErrorEstimator (h_K * sum over interior faces of int [d_n u_h]^2, face rule
// QGauss<2>(2), Dirichlet faces contribute nothing) plus the residual term h_K^2 int (4 pi rho)^2, Vector<float>
// storage.  The host keeps the mesh and hands over the face topology; every cell gathers its own six faces (the
// coarse side of a hanging face visits its four fine neighbours in the reference's cell order), so no atomics and
// the same operation order as the sequential loop: all arithmetic is spelled with __d*_rn (no FMA contraction) and
// the float32 indicators are bit-identical to the host restatement.
#include <cmath>
#include <vector>

#include "context.h"

using namespace gmg;

// defined in rhs.cu: the cell arrays kept on the device by the last gmg_assemble_rhs / gmg_charge_density call
int gmg_rhs_resident(gmg_context *h, int *n_cells, const double **cell_h, const int **cell_dofs, const double **weights, int *n_q,
                     const double **rho, int *rho_cells, int *rho_nq);

namespace {

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }

// d/dx_a of the Q1 interpolant on a cube of edge h at tangential unit coordinates (s, t) of the two other axes
__device__ __forceinline__ double normal_derivative(const double U[8], double h, int a, double s, double t) {
  const int o0 = (a == 0) ? 1 : 0, o1 = (a == 2) ? 1 : 2;
  double r = 0.0;
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    double w = __ddiv_rn(((v >> a) & 1) ? 1.0 : -1.0, h);
    w = mul(w, ((v >> o0) & 1) ? s : __dsub_rn(1.0, s));
    w = mul(w, ((v >> o1) & 1) ? t : __dsub_rn(1.0, t));
    r = add(r, mul(U[v], w));
  }
  return r;
}

// int over a (sub)face of edge h of the squared jump between the cell with values U (edge h) and its neighbour Un
// (edge hn; `coarse`: the face is the (sub0, sub1) quarter of the neighbour's face)
__device__ __forceinline__ double face_integral(const double U[8], double h, const double Un[8], double hn, int a, bool coarse,
                                                int sub0, int sub1, const double gp2[2], const double gw2[2]) {
  double I = 0.0;
  for (int t1 = 0; t1 < 2; ++t1)
    for (int t0 = 0; t0 < 2; ++t0) {
      const double s = gp2[t0], t = gp2[t1];
      const double own = normal_derivative(U, h, a, s, t);
      const double oth = coarse ? normal_derivative(Un, hn, a, __ddiv_rn(add((double)sub0, s), 2.0), __ddiv_rn(add((double)sub1, t), 2.0))
                                : normal_derivative(Un, hn, a, s, t);
      const double j = __dsub_rn(own, oth);
      I = add(I, mul(mul(j, j), mul(mul(mul(gw2[t0], gw2[t1]), h), h)));
    }
  return I;
}

__device__ __forceinline__ void load_cell(const double *__restrict__ u, const int *__restrict__ dofs, int c, double U[8]) {
#pragma unroll
  for (int v = 0; v < 8; ++v) U[v] = u[dofs[8 * (int64_t)c + v]];
}

__global__ void __launch_bounds__(128) indicator_kernel(int n_cells, const double *__restrict__ cell_h, const int *__restrict__ cell_dofs,
                                                        const int *__restrict__ face_nb, const unsigned char *__restrict__ face_kind,
                                                        const int *__restrict__ hang_children, const double *__restrict__ u,
                                                        const double *__restrict__ rho, int n_q, const double *__restrict__ weights,
                                                        double gp0, double gp1, double gw0, double gw1, float *__restrict__ eta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const double gp2[2] = {gp0, gp1}, gw2[2] = {gw0, gw1};
  const double h = cell_h[c];
  const double diam = sqrt(mul(mul(3.0, h), h));
  double U[8], Un[8];
  load_cell(u, cell_dofs, c, U);
  float err = 0.0f;
  for (int face = 0; face < 6; ++face) {
    const int a = face >> 1;
    const int nb = face_nb[6 * (int64_t)c + face];
    const int kind = face_kind[6 * (int64_t)c + face];
    double I = 0.0;
    if (nb >= 0) {
      if ((kind & 3) == 0) {  // same-level neighbour
        load_cell(u, cell_dofs, nb, Un);
        I = face_integral(U, h, Un, h, a, false, 0, 0, gp2, gw2);
      } else if ((kind & 3) == 1) {  // this cell is on the fine side of a hanging face
        load_cell(u, cell_dofs, nb, Un);
        I = face_integral(U, h, Un, mul(2.0, h), a, true, (kind >> 2) & 1, (kind >> 3) & 1, gp2, gw2);
      } else {  // coarse side: the four subface integrals as their fine cells compute them, in the reference's order
        for (int k = 0; k < 4; ++k) {
          const int ch = hang_children[4 * (int64_t)nb + k];
          const int ck = face_kind[6 * (int64_t)ch + (face ^ 1)];
          const double hc = cell_h[ch];
          load_cell(u, cell_dofs, ch, Un);
          I = add(I, face_integral(Un, hc, U, mul(2.0, hc), a, true, (ck >> 2) & 1, (ck >> 3) & 1, gp2, gw2));
        }
      }
    }
    err = (float)add((double)err, mul(I, diam));
  }
  const float kelly = (float)sqrt((double)err);
  double resid = 0.0;
  if (rho) {
    const double *r = rho + (int64_t)c * n_q;
    for (int q = 0; q < n_q; ++q) {
      const double t = mul(mul(4.0, M_PI), r[q]);
      resid = add(resid, mul(mul(t, t), mul(mul(mul(weights[q], h), h), h)));
    }
  }
  eta[c] = (float)sqrt(add(mul((double)kelly, (double)kelly), mul(mul(diam, diam), resid)));
}

__global__ void __launch_bounds__(256) float_max_kernel(int n, const float *__restrict__ v, float *out /* zero-initialised */) {
  float m = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, v[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  // indicators are >= 0: their float order is the order of their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));
}

// src/step-50.cc:1084-1090: a cell is flagged when its (float) indicator, widened to double, reaches the threshold
__global__ void __launch_bounds__(256) mark_kernel(int n, const float *__restrict__ eta, double threshold, bool any, unsigned char *flags) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) flags[c] = (any && (double)eta[c] >= threshold) ? 1 : 0;
}

// ---- solution transfer (src/step-50.cc:1110-1119)
__global__ void __launch_bounds__(256) xfer_init(int n, int *stamp, int *owner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    stamp[i] = 0x7fffffff;  // unknown: later than any pass
    owner[i] = 0x7fffffff;
  }
}
__global__ void __launch_bounds__(256) xfer_copy(int n, const int *__restrict__ src, const int *__restrict__ dst,
                                                 const double *__restrict__ u_old, int n_old, int n_new, double *x, int *stamp,
                                                 int *bad_index) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    if (src[k] < 0 || src[k] >= n_old || dst[k] < 0 || dst[k] >= n_new) {
      *bad_index = 1;
      return;
    }
    x[dst[k]] = u_old[src[k]];
    stamp[dst[k]] = -1;
  }
}
// the hand-over is untrusted: an index outside [-1, n_new) is reported, never dereferenced
__global__ void __launch_bounds__(256) xfer_check(int64_t n, const int *__restrict__ pd, int n_new, int *bad_index) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && (pd[i] < -1 || pd[i] >= n_new)) *bad_index = 1;
}
// pass l, step 1: a refined cell whose corners are known claims its unknown points (lowest cell index wins)
__global__ void __launch_bounds__(256) xfer_claim(int64_t p0, int64_t p1, const int *__restrict__ pd, int pass,
                                                  const int *__restrict__ stamp, int *owner) {
  const int64_t p = p0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= p1) return;
  const int *q = pd + 27 * p;
  for (int v = 0; v < 8; ++v) {
    const int c = q[2 * (v & 1) + 6 * ((v >> 1) & 1) + 18 * ((v >> 2) & 1)];
    if (c < 0 || stamp[c] >= pass) return;
  }
  for (int t = 0; t < 27; ++t) {
    const int t0 = t % 3, t1 = (t / 3) % 3, t2 = t / 9;
    if (t0 != 1 && t1 != 1 && t2 != 1) continue;
    const int dof = q[t];
    if (dof >= 0 && stamp[dof] >= pass) atomicMin(owner + dof, (int)(p - p0));
  }
}
// step 2: the owner interpolates: val = sum_v w_v U_v over the corners with w_v != 0, in vertex order, unfused
__global__ void __launch_bounds__(256) xfer_fill(int64_t p0, int64_t p1, const int *__restrict__ pd, int pass,
                                                 const int *__restrict__ owner, double *x, int *stamp) {
  const int64_t p = p0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= p1) return;
  const int *q = pd + 27 * p;
  double U[8];
  for (int v = 0; v < 8; ++v) {
    const int c = q[2 * (v & 1) + 6 * ((v >> 1) & 1) + 18 * ((v >> 2) & 1)];
    if (c < 0 || stamp[c] >= pass) return;
    U[v] = x[c];
  }
  for (int t = 0; t < 27; ++t) {
    const int tt[3] = {t % 3, (t / 3) % 3, t / 9};
    if (tt[0] != 1 && tt[1] != 1 && tt[2] != 1) continue;
    const int dof = q[t];
    if (dof < 0 || owner[dof] != (int)(p - p0) || stamp[dof] < pass) continue;
    double val = 0.0;
    for (int v = 0; v < 8; ++v) {
      double w = 1.0;
      for (int k = 0; k < 3; ++k) w = __dmul_rn(w, ((v >> k) & 1) ? tt[k] / 2.0 : 1.0 - tt[k] / 2.0);
      if (w != 0.0) val = __dadd_rn(val, __dmul_rn(w, U[v]));
    }
    x[dof] = val;
  }
}
// step 3 (after every owner has read the corner stamps): the claimed points are known from the next pass on
__global__ void __launch_bounds__(256) xfer_stamp(int n, int pass, int *owner, int *stamp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && owner[i] != 0x7fffffff) {
    stamp[i] = pass;
    owner[i] = 0x7fffffff;
  }
}
__global__ void __launch_bounds__(256) xfer_finish(int n, const unsigned char *__restrict__ constrained, const int *__restrict__ stamp,
                                                   double *x, int *missing) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (stamp[i] == 0x7fffffff) atomicAdd(missing, 1);
  if (constrained[i]) x[i] = 0.0;
}

}  // namespace

extern "C" {

int gmg_transfer_solution(gmg_handle h, int32_t n_old, const double *u_old, int32_t n_new, int32_t n_copy, const int32_t *copy_old,
                          const int32_t *copy_new, int32_t n_pass, const int64_t *pass_ptr, const int32_t *parent_dofs,
                          const uint8_t *constrained, double *u_new_out) {
  if (!h || n_old < 0 || n_new < 0 || n_copy < 0 || n_pass < 0 || !u_old || !u_new_out || !constrained || !pass_ptr ||
      (n_copy > 0 && (!copy_old || !copy_new)) || (pass_ptr[n_pass] > 0 && !parent_dofs))
    return GMG_EINVAL;
  gmg::enter(h);
  const int64_t n_par = pass_ptr[n_pass];
  double *d_old = nullptr, *d_x = nullptr;
  int *d_src = nullptr, *d_dst = nullptr, *d_pd = nullptr, *d_stamp = nullptr, *d_owner = nullptr, *d_missing = nullptr, *d_bad = nullptr;
  unsigned char *d_con = nullptr;
  auto cleanup = [&]() {
    dfree(d_old); dfree(d_x); dfree(d_src); dfree(d_dst); dfree(d_pd); dfree(d_stamp); dfree(d_owner); dfree(d_missing); dfree(d_con); dfree(d_bad);
  };
  cudaError_t e = cudaSuccess;
  if ((e = dalloc(&d_old, n_old)) != cudaSuccess || (e = dalloc(&d_x, n_new)) != cudaSuccess || (e = dalloc(&d_src, n_copy)) != cudaSuccess ||
      (e = dalloc(&d_dst, n_copy)) != cudaSuccess || (e = dalloc(&d_pd, 27 * n_par)) != cudaSuccess ||
      (e = dalloc(&d_stamp, n_new)) != cudaSuccess || (e = dalloc(&d_owner, n_new)) != cudaSuccess ||
      (e = dalloc(&d_missing, 1)) != cudaSuccess || (e = dalloc(&d_con, n_new)) != cudaSuccess ||
      (e = dalloc(&d_bad, 1)) != cudaSuccess) {
    cleanup();
    return gmg::fail(h, GMG_ECUDA, std::string("gmg_transfer_solution: ") + cudaGetErrorString(e));
  }
  int rc = GMG_OK;
  if ((rc = staged_h2d(h, d_old, u_old, sizeof(double) * (size_t)n_old)) || (rc = staged_h2d(h, d_src, copy_old, sizeof(int) * (size_t)n_copy)) ||
      (rc = staged_h2d(h, d_dst, copy_new, sizeof(int) * (size_t)n_copy)) ||
      (rc = staged_h2d(h, d_pd, parent_dofs, sizeof(int) * 27 * (size_t)n_par)) || (rc = staged_h2d(h, d_con, constrained, (size_t)n_new))) {
    cleanup();
    return rc;
  }
  cudaMemsetAsync(d_x, 0, sizeof(double) * (size_t)std::max(n_new, 1), h->stream);
  cudaMemsetAsync(d_missing, 0, sizeof(int), h->stream);
  cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream);
  if (n_par > 0) {
    xfer_check<<<cdiv(27 * n_par, 256), 256, 0, h->stream>>>(27 * n_par, d_pd, n_new, d_bad);
    h->launches++;
    int bad = 0;
    e = gmg::copy_sync(h, &bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess || bad) {
      cleanup();
      return gmg::fail(h, e != cudaSuccess ? GMG_ECUDA : GMG_EINVAL, "gmg_transfer_solution: a parent_dofs entry is outside [-1, n_new)");
    }
  }
  if (n_new > 0) {
    xfer_init<<<cdiv(n_new, 256), 256, 0, h->stream>>>(n_new, d_stamp, d_owner);
    h->launches++;
  }
  if (n_copy > 0) {
    xfer_copy<<<cdiv(n_copy, 256), 256, 0, h->stream>>>(n_copy, d_src, d_dst, d_old, n_old, n_new, d_x, d_stamp, d_bad);
    h->launches++;
  }
  for (int l = 0; l < n_pass; ++l) {
    const int64_t p0 = pass_ptr[l], p1 = pass_ptr[l + 1];
    if (p1 <= p0) continue;
    const int grid = (int)cdiv(p1 - p0, (int64_t)256);
    xfer_claim<<<grid, 256, 0, h->stream>>>(p0, p1, d_pd, l, d_stamp, d_owner);
    xfer_fill<<<grid, 256, 0, h->stream>>>(p0, p1, d_pd, l, d_owner, d_x, d_stamp);
    xfer_stamp<<<cdiv(n_new, 256), 256, 0, h->stream>>>(n_new, l, d_owner, d_stamp);
    h->launches += 3;
  }
  if (n_new > 0) {
    xfer_finish<<<cdiv(n_new, 256), 256, 0, h->stream>>>(n_new, d_con, d_stamp, d_x, d_missing);
    h->launches++;
  }
  int missing = 0, bad_copy = 0;
  if ((rc = staged_d2h(h, u_new_out, d_x, sizeof(double) * (size_t)n_new)) == GMG_OK) {
    gmg::copy(h, &bad_copy, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
    e = gmg::copy_sync(h, &missing, d_missing, sizeof(int), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = gmg::fail(h, GMG_ECUDA, std::string("gmg_transfer_solution: ") + cudaGetErrorString(e));
  }
  cleanup();
  if (rc == GMG_OK && bad_copy) return gmg::fail(h, GMG_EINVAL, "gmg_transfer_solution: a copy index is out of range");
  if (rc == GMG_OK && missing > 0) return gmg::fail(h, GMG_EINVAL, "solution transfer left a dof without a value");
  return rc;
}

int gmg_error_indicator(gmg_handle h, int32_t n_cells, const int32_t *face_nb, const uint8_t *face_kind, int32_t n_hang,
                        const int32_t *hang_children, const double *u, int32_t n_dofs, const double *rho, int residual_term,
                        const double gauss2_points[2], const double gauss2_weights[2], float *eta_out, float *max_out) {
  if (!h || n_cells < 0 || !face_nb || !face_kind || !u || !eta_out || !gauss2_points || !gauss2_weights ||
      (n_hang > 0 && !hang_children))
    return GMG_EINVAL;
  gmg::enter(h);
  int r_cells = 0, n_q = 0, rho_cells = 0, rho_nq = 0;
  const double *cell_h = nullptr, *weights = nullptr, *rho_dev = nullptr;
  const int *cell_dofs = nullptr;
  if (int rc = gmg_rhs_resident(h, &r_cells, &cell_h, &cell_dofs, &weights, &n_q, &rho_dev, &rho_cells, &rho_nq)) return rc;
  if (r_cells != n_cells || !cell_h || !cell_dofs)
    return gmg::fail(h, GMG_EINVAL, "gmg_error_indicator: call gmg_assemble_rhs on the same active cells first");
  int *d_nb = nullptr, *d_hang = nullptr;
  unsigned char *d_kind = nullptr;
  double *d_u = nullptr, *d_rho = nullptr;
  float *d_eta = nullptr, *d_max = nullptr;
  auto cleanup = [&]() {
    dfree(d_nb);
    dfree(d_hang);
    dfree(d_kind);
    dfree(d_u);
    dfree(d_rho);
    dfree(d_eta);
    dfree(d_max);
  };
  const int64_t nc6 = 6 * (int64_t)n_cells;
  const double *rho_use = nullptr;
  int rc = GMG_OK;
  cudaError_t e = cudaSuccess;
  if ((e = dalloc(&d_nb, nc6)) != cudaSuccess || (e = dalloc(&d_kind, nc6)) != cudaSuccess ||
      (e = dalloc(&d_hang, 4 * (int64_t)n_hang)) != cudaSuccess || (e = dalloc(&d_u, n_dofs)) != cudaSuccess ||
      (e = dalloc(&d_eta, n_cells)) != cudaSuccess || (e = dalloc(&d_max, 1)) != cudaSuccess) {
    cleanup();
    return gmg::fail(h, GMG_ECUDA, std::string("gmg_error_indicator: ") + cudaGetErrorString(e));
  }
  if ((rc = staged_h2d(h, d_nb, face_nb, sizeof(int) * nc6)) || (rc = staged_h2d(h, d_kind, face_kind, nc6)) ||
      (n_hang > 0 && (rc = staged_h2d(h, d_hang, hang_children, sizeof(int) * 4 * (int64_t)n_hang))) ||
      (rc = staged_h2d(h, d_u, u, sizeof(double) * n_dofs))) {
    cleanup();
    return rc;
  }
  if (residual_term) {
    if (rho) {  // densities handed over by the caller (n_cells x n_q, the quadrature of gmg_assemble_rhs)
      if (dalloc(&d_rho, (int64_t)n_cells * n_q) != cudaSuccess || (rc = staged_h2d(h, d_rho, rho, sizeof(double) * (int64_t)n_cells * n_q))) {
        cleanup();
        return rc ? rc : gmg::fail(h, GMG_ECUDA, "gmg_error_indicator: allocation failed");
      }
      rho_use = d_rho;
    } else if (rho_dev && rho_cells == n_cells && rho_nq == n_q) {
      rho_use = rho_dev;  // still on the device from gmg_charge_density
    } else {
      cleanup();
      return gmg::fail(h, GMG_EINVAL, "gmg_error_indicator: no charge densities for the residual term");
    }
  }
  cudaMemsetAsync(d_max, 0, sizeof(float), h->stream);
  if (n_cells > 0) {
    indicator_kernel<<<cdiv(n_cells, 128), 128, 0, h->stream>>>(n_cells, cell_h, cell_dofs, d_nb, d_kind, d_hang, d_u, rho_use, n_q,
                                                                weights, gauss2_points[0], gauss2_points[1], gauss2_weights[0],
                                                                gauss2_weights[1], d_eta);
    h->launches++;
    float_max_kernel<<<std::min(cdiv(n_cells, 256), 1024), 256, 0, h->stream>>>(n_cells, d_eta, d_max);
    h->launches++;
  }
  float mx = 0.0f;
  if ((rc = staged_d2h(h, eta_out, d_eta, sizeof(float) * (size_t)n_cells)) == GMG_OK) {
    e = gmg::copy_sync(h, &mx, d_max, sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = gmg::fail(h, GMG_ECUDA, std::string("gmg_error_indicator: ") + cudaGetErrorString(e));
  }
  if (max_out) *max_out = mx;
  if (rc == GMG_OK) {  // the indicators stay on the device for gmg_mark_cells
    dfree(h->ind_eta);
    h->ind_eta = d_eta;
    h->ind_n = n_cells;
    h->ind_max = mx;
    d_eta = nullptr;
  }
  cleanup();
  return rc;
}

int gmg_mark_cells(gmg_handle h, int32_t n_cells, double fraction, uint8_t *flags_out, double *threshold_out) {
  if (!h || n_cells < 0 || !flags_out) return GMG_EINVAL;
  gmg::enter(h);
  if (!h->ind_eta || h->ind_n != n_cells) return gmg::fail(h, GMG_EINVAL, "gmg_mark_cells: call gmg_error_indicator on the same cells first");
  const double threshold = fraction * (double)h->ind_max;  // (src/step-50.cc:1084: 0.6 * estimated_error_per_cell.linfty_norm())
  if (threshold_out) *threshold_out = threshold;
  unsigned char *d_flags = nullptr;
  if (cudaError_t e = dalloc(&d_flags, n_cells); e != cudaSuccess) return gmg::fail(h, GMG_ECUDA, cudaGetErrorString(e));
  if (n_cells > 0) {
    mark_kernel<<<cdiv(n_cells, 256), 256, 0, h->stream>>>(n_cells, h->ind_eta, threshold, h->ind_max > 0.0f, d_flags);
    h->launches++;
  }
  int rc = staged_d2h(h, flags_out, d_flags, (size_t)n_cells);
  dfree(d_flags);
  return rc;
}

}  // extern "C"
