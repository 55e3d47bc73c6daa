/* oracle.c -- plain C (OpenMP) restatement of the reference's hot path, for the CPU baseline and for
 * parity checks at sizes where the numpy oracle is slow.  TEST INFRASTRUCTURE: only tests/, smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.  Same algorithms as
 * oracle/solver.py and oracle/rhs.py (which are pinned against the reference's golden stdout):
 *
 *   orc_bin_atoms    rhs_assembly_optimization          src/step-50.cc:260-306
 *   orc_density      compute_charge_densities           src/step-50.cc:509-575
 *   orc_load_vector  assemble_system, load vector part  src/step-50.cc:813-828
 *   orc_pcg_gmg      solve(): SolverCG + PreconditionMG + Multigrid V-cycle + MGTransferPrebuilt +
 *                    Ifpack point relaxation + coarse SolverCG   src/step-50.cc:938-1017
 *
 * Parallelism mirrors the reference's MPI decomposition: rows are split into `n_blocks` contiguous
 * blocks (one per thread); SpMV / vector updates / dot products are block-parallel, SSOR is
 * processor-block SSOR (off-block couplings see zeros within a sweep), exactly as Ifpack's point
 * relaxation with overlap 0 behaves on n_blocks ranks.  n_blocks = 1 is the lexicographic 1-rank case.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int n_rows, n_cols;
  const int64_t *rowptr;
  const int32_t *col;
  const double *val;
} csr_t;

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------ BLAS-1/2 */
static void spmv(const csr_t *A, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < A->n_rows; ++r) {
    double s = 0.0;
    for (int64_t k = A->rowptr[r]; k < A->rowptr[r + 1]; ++k) s += A->val[k] * x[A->col[k]];
    y[r] = s;
  }
}
/* y += alpha * A^T x  (restrict_and_add, Tvmult); serial scatter: the transfer operators are small */
static void spmv_t_add(const csr_t *A, double alpha, const double *x, double *y) {
  for (int r = 0; r < A->n_rows; ++r) {
    const double xr = alpha * x[r];
    if (xr == 0.0) continue;
    for (int64_t k = A->rowptr[r]; k < A->rowptr[r + 1]; ++k) y[A->col[k]] += A->val[k] * xr;
  }
}
static double dot(int n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* ------------------------------------------------------------------------------------ SolverCG */
typedef void (*precond_fn)(void *ctx, const double *src, double *dst);

/* deal.II SolverCG::solve; returns 0 converged, 1 failure.  work: 3 n doubles. */
static int cg(const csr_t *A, const double *b, double *x, int x_is_zero, double tol, int max_it, precond_fn M, void *ctx,
              double *work, int *iters, double *res0, double *res_last) {
  const int n = A->n_rows;
  double *g = work, *d = work + n, *h = work + 2 * (size_t)n;
  if (!x_is_zero) {
    spmv(A, x, g);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) g[i] -= b[i];
  } else {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) g[i] = -b[i];
  }
  double res = sqrt(dot(n, g, g));
  *res0 = res;
  *res_last = res;
  *iters = 0;
  if (res <= tol) return 0;
  double gh;
  if (M) {
    M(ctx, g, h);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) d[i] = -h[i];
    gh = dot(n, g, h);
  } else {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) d[i] = -g[i];
    gh = res * res;
  }
  int it = 0;
  for (;;) {
    ++it;
    spmv(A, d, h);
    const double alpha = gh / dot(n, d, h);
    double gg = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : gg)
    for (int i = 0; i < n; ++i) {
      x[i] += alpha * d[i];
      g[i] += alpha * h[i];
      gg += g[i] * g[i];
    }
    res = sqrt(gg);
    *iters = it;
    *res_last = res;
    if (res <= tol) return 0;
    if (it >= max_it) return 1;
    double beta = gh;
    if (M) {
      M(ctx, g, h);
      gh = dot(n, g, h);
      beta = gh / beta;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) d[i] = beta * d[i] - h[i];
    } else {
      gh = res * res;
      beta = gh / beta;
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) d[i] = beta * d[i] - g[i];
    }
  }
}

/* ------------------------------------------------------------------------------------ multigrid */
typedef struct {
  int n_levels;
  const csr_t *A, *I, *P; /* P[l]: level l -> l+1 */
  const int32_t *const *copy_g, *const *copy_l;
  const int32_t *n_copy;
  int n_sys;
  int smoother; /* 0 Jacobi, 1 SSOR (processor-block with n_blocks blocks) */
  double omega;
  int steps, n_blocks;
  double coarse_tol;
  int coarse_max;
  double **defect, **sol, **t, **tmp, **diag;
  double *cg_work;
  int32_t *coarse_its;
  int n_coarse, coarse_cap, coarse_failed;
} mg_t;

/* y = omega D^-1 r */
static void jacobi_apply(const mg_t *m, int l, const double *r, double *y) {
  const int n = m->A[l].n_rows;
  const double *dg = m->diag[l];
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) y[i] = m->omega * r[i] / dg[i];
}

/* Ifpack symmetric Gauss-Seidel, one sweep, zero start, per contiguous row block */
static void ssor_apply(const mg_t *m, int l, const double *r, double *y) {
  const csr_t *A = &m->A[l];
  const int n = A->n_rows, nb = m->n_blocks < 1 ? 1 : m->n_blocks;
  const double *dg = m->diag[l];
  const double w = m->omega;
#pragma omp parallel for schedule(static, 1)
  for (int b = 0; b < nb; ++b) {
    const int r0 = (int)((int64_t)n * b / nb), r1 = (int)((int64_t)n * (b + 1) / nb);
    for (int i = r0; i < r1; ++i) y[i] = 0.0;
    for (int i = r0; i < r1; ++i) {
      double s = 0.0;
      for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; ++k) {
        const int j = A->col[k];
        if (j >= r0 && j < r1) s += A->val[k] * y[j];
      }
      y[i] += w * (r[i] - s) / dg[i];
    }
    for (int i = r1 - 1; i >= r0; --i) {
      double s = 0.0;
      for (int64_t k = A->rowptr[i]; k < A->rowptr[i + 1]; ++k) {
        const int j = A->col[k];
        if (j >= r0 && j < r1) s += A->val[k] * y[j];
      }
      y[i] += w * (r[i] - s) / dg[i];
    }
  }
}

static void smoother_apply(const mg_t *m, int l, const double *r, double *y) {
  if (m->smoother == 0) jacobi_apply(m, l, r, y);
  else ssor_apply(m, l, r, y);
}

/* MGSmootherPrecondition::smooth: u += P^-1 (rhs - A u), `steps` times */
static void smooth(mg_t *m, int l, double *u, const double *rhs, int zero_start) {
  const csr_t *A = &m->A[l];
  const int n = A->n_rows;
  double *r = m->tmp[l], *dlt = m->t[l];
  for (int s = 0; s < m->steps; ++s) {
    if (zero_start && s == 0) {
      smoother_apply(m, l, rhs, u);
    } else {
      spmv(A, u, r);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) r[i] = rhs[i] - r[i];
      smoother_apply(m, l, r, dlt);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) u[i] += dlt[i];
    }
  }
}

static void level_v_step(mg_t *m, int l) {
  if (l == 0) {
    const int n = m->A[0].n_rows;
    int its;
    double r0, r1;
    memset(m->sol[0], 0, sizeof(double) * n);
    const int fail = cg(&m->A[0], m->defect[0], m->sol[0], 1, m->coarse_tol, m->coarse_max, NULL, NULL, m->cg_work, &its,
                        &r0, &r1);
    if (fail) m->coarse_failed = 1;
    if (m->n_coarse < m->coarse_cap) m->coarse_its[m->n_coarse] = its;
    m->n_coarse++;
    return;
  }
  const int n = m->A[l].n_rows;
  double *u = m->sol[l], *t = m->t[l];
  smooth(m, l, u, m->defect[l], 1);
  /* t = defect - A u - I u */
  spmv(&m->A[l], u, t);
  {
    const csr_t *I = &m->I[l];
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
      double s = 0.0;
      for (int64_t k = I->rowptr[r]; k < I->rowptr[r + 1]; ++k) s += I->val[k] * u[I->col[k]];
      t[r] = m->defect[l][r] - t[r] - s;
    }
  }
  /* save t: smooth() reuses m->t as scratch, so restrict now */
  spmv_t_add(&m->P[l - 1], 1.0, t, m->defect[l - 1]);
  level_v_step(m, l - 1);
  /* u += P sol[l-1] */
  {
    const csr_t *P = &m->P[l - 1];
    const double *c = m->sol[l - 1];
#pragma omp parallel for schedule(static)
    for (int r = 0; r < n; ++r) {
      double s = 0.0;
      for (int64_t k = P->rowptr[r]; k < P->rowptr[r + 1]; ++k) s += P->val[k] * c[P->col[k]];
      u[r] += s;
    }
  }
  /* defect -= I^T u */
  spmv_t_add(&m->I[l], -1.0, u, m->defect[l]);
  smooth(m, l, u, m->defect[l], 0);
}

/* PreconditionMG::vmult */
static void mg_vmult(void *ctx, const double *src, double *dst) {
  mg_t *m = (mg_t *)ctx;
  for (int l = 0; l < m->n_levels; ++l) {
    memset(m->defect[l], 0, sizeof(double) * m->A[l].n_rows);
    for (int i = 0; i < m->n_copy[l]; ++i) m->defect[l][m->copy_l[l][i]] = src[m->copy_g[l][i]];
  }
  level_v_step(m, m->n_levels - 1);
  memset(dst, 0, sizeof(double) * m->n_sys);
  for (int l = 0; l < m->n_levels; ++l)
    for (int i = 0; i < m->n_copy[l]; ++i) dst[m->copy_g[l][i]] = m->sol[l][m->copy_l[l][i]];
}

static void extract_diag(const csr_t *A, double *dg) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < A->n_rows; ++r) {
    double d = 0.0;
    for (int64_t k = A->rowptr[r]; k < A->rowptr[r + 1]; ++k)
      if (A->col[k] == r) d += A->val[k];
    dg[r] = d;
  }
}

/* Flat-array entry point.  Matrices are passed as arrays of pointers per level.
 * Returns 0 ok, 1 outer no-convergence, 2 coarse no-convergence. */
int orc_pcg_gmg(int n_levels, int n_sys, const int64_t *sys_rowptr, const int32_t *sys_col, const double *sys_val,
                const int32_t *level_n, const int64_t *const *A_rowptr, const int32_t *const *A_col,
                const double *const *A_val, const int64_t *const *I_rowptr, const int32_t *const *I_col,
                const double *const *I_val, const int64_t *const *P_rowptr, const int32_t *const *P_col,
                const double *const *P_val, const int32_t *n_copy, const int32_t *const *copy_g,
                const int32_t *const *copy_l, int smoother, double omega, int steps, int n_blocks, const double *b,
                double *x_inout, double tol, int max_it, double coarse_tol, int coarse_max, int *iters, double *res0,
                double *res_last, int32_t *coarse_its, int coarse_cap, int *n_coarse) {
  mg_t m;
  memset(&m, 0, sizeof m);
  csr_t *A = (csr_t *)calloc(n_levels, sizeof(csr_t)), *I = (csr_t *)calloc(n_levels, sizeof(csr_t)),
        *P = (csr_t *)calloc(n_levels, sizeof(csr_t));
  m.defect = (double **)calloc(n_levels, sizeof(double *));
  m.sol = (double **)calloc(n_levels, sizeof(double *));
  m.t = (double **)calloc(n_levels, sizeof(double *));
  m.tmp = (double **)calloc(n_levels, sizeof(double *));
  m.diag = (double **)calloc(n_levels, sizeof(double *));
  for (int l = 0; l < n_levels; ++l) {
    const int n = level_n[l];
    A[l] = (csr_t){n, n, A_rowptr[l], A_col[l], A_val[l]};
    if (l >= 1) I[l] = (csr_t){n, n, I_rowptr[l], I_col[l], I_val[l]};
    if (l + 1 < n_levels) P[l] = (csr_t){level_n[l + 1], n, P_rowptr[l], P_col[l], P_val[l]};
    m.defect[l] = (double *)calloc(n > 0 ? n : 1, sizeof(double));
    m.sol[l] = (double *)calloc(n > 0 ? n : 1, sizeof(double));
    m.t[l] = (double *)calloc(n > 0 ? n : 1, sizeof(double));
    m.tmp[l] = (double *)calloc(n > 0 ? n : 1, sizeof(double));
    m.diag[l] = (double *)calloc(n > 0 ? n : 1, sizeof(double));
    extract_diag(&A[l], m.diag[l]);
  }
  m.n_levels = n_levels;
  m.A = A;
  m.I = I;
  m.P = P;
  m.copy_g = copy_g;
  m.copy_l = copy_l;
  m.n_copy = n_copy;
  m.n_sys = n_sys;
  m.smoother = smoother;
  m.omega = omega;
  m.steps = steps;
  m.n_blocks = n_blocks;
  m.coarse_tol = coarse_tol;
  m.coarse_max = coarse_max;
  m.cg_work = (double *)calloc(3 * (size_t)(level_n[0] > 0 ? level_n[0] : 1), sizeof(double));
  m.coarse_its = coarse_its;
  m.coarse_cap = coarse_cap;
  csr_t S = {n_sys, n_sys, sys_rowptr, sys_col, sys_val};
  double *work = (double *)calloc(3 * (size_t)n_sys, sizeof(double));
  int x_zero = 1;
  for (int i = 0; i < n_sys; ++i)
    if (x_inout[i] != 0.0) { x_zero = 0; break; }
  int rc = cg(&S, b, x_inout, x_zero, tol, max_it, mg_vmult, &m, work, iters, res0, res_last);
  if (m.coarse_failed) rc = 2;
  *n_coarse = m.n_coarse;
  for (int l = 0; l < n_levels; ++l) {
    free(m.defect[l]); free(m.sol[l]); free(m.t[l]); free(m.tmp[l]); free(m.diag[l]);
  }
  free(m.defect); free(m.sol); free(m.t); free(m.tmp); free(m.diag); free(m.cg_work); free(work);
  free(A); free(I); free(P);
  return rc;
}

/* plain CG from x = 0 on one matrix (the coarse solver alone) */
int orc_cg(int n, const int64_t *rowptr, const int32_t *col, const double *val, const double *b, double *x, double tol,
           int max_it, int *iters, double *res_last) {
  csr_t A = {n, n, rowptr, col, val};
  double *work = (double *)calloc(3 * (size_t)n, sizeof(double));
  double r0;
  memset(x, 0, sizeof(double) * n);
  const int rc = cg(&A, b, x, 1, tol, max_it, NULL, NULL, work, iters, &r0, res_last);
  free(work);
  return rc;
}

void orc_spmv(int n_rows, const int64_t *rowptr, const int32_t *col, const double *val, const double *x, double *y) {
  csr_t A = {n_rows, 0, rowptr, col, val};
  spmv(&A, x, y);
}

/* ------------------------------------------------------------------------------------ RHS path */
/* cells of a structured base lattice (reps^3, x fastest): list atom i iff a vertex is within radius.
 * Pass atoms_out == NULL to count (rowptr_out filled); then call again with the buffer. */
int64_t orc_bin_atoms(int reps, double lo, double H, int n_atoms, const double *pos, double radius, int64_t *rowptr_out,
                      int32_t *atoms_out) {
  const int64_t nc = (int64_t)reps * reps * reps;
  const int w = (int)ceil(radius / H) + 1;
  int32_t *count = (int32_t *)calloc(nc, sizeof(int32_t));
  /* the literal criterion on candidate cells around each atom; per cell the atoms are appended in
   * ascending atom order because the outer loop is over atoms */
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) {
      if (!atoms_out) break;
      memset(count, 0, sizeof(int32_t) * nc);
    }
    for (int i = 0; i < n_atoms; ++i) {
      const double *X = pos + 3 * (size_t)i;
      int c0[3];
      for (int d = 0; d < 3; ++d) c0[d] = (int)floor((X[d] - lo) / H);
      for (int k = c0[2] - w; k <= c0[2] + w; ++k) {
        if (k < 0 || k >= reps) continue;
        for (int j = c0[1] - w; j <= c0[1] + w; ++j) {
          if (j < 0 || j >= reps) continue;
          for (int ii = c0[0] - w; ii <= c0[0] + w; ++ii) {
            if (ii < 0 || ii >= reps) continue;
            int hit = 0;
            for (int v = 0; v < 8 && !hit; ++v) {
              const double vx = lo + (ii + (v & 1)) * H, vy = lo + (j + ((v >> 1) & 1)) * H,
                           vz = lo + (k + ((v >> 2) & 1)) * H;
              const double dx = X[0] - vx, dy = X[1] - vy, dz = X[2] - vz;
              if (sqrt(dx * dx + dy * dy + dz * dz) < radius) hit = 1;
            }
            if (!hit) continue;
            const int64_t c = ii + (int64_t)reps * (j + (int64_t)reps * k);
            if (pass == 1) atoms_out[rowptr_out[c] + count[c]] = i;
            count[c]++;
          }
        }
      }
    }
    if (pass == 0) {
      rowptr_out[0] = 0;
      for (int64_t c = 0; c < nc; ++c) rowptr_out[c + 1] = rowptr_out[c] + count[c];
    }
  }
  free(count);
  return rowptr_out[nc];
}

void orc_density(int n_cells, const double *cell_lo, const double *cell_h, const int32_t *list_of_cell,
                 const int64_t *list_ptr, const int32_t *list_atoms, int n_atoms, const double *pos, const double *q,
                 int n_q, const double *qpts, double r_c, double *rho) {
  const double C = 4.0 * M_PI / (pow(r_c, 3) * pow(M_PI, 1.5));
  const double inv = 1.0 / (r_c * r_c);
#pragma omp parallel for schedule(dynamic, 256)
  for (int c = 0; c < n_cells; ++c) {
    const int list = list_of_cell[c];
    const int64_t a0 = list >= 0 ? list_ptr[list] : 0, a1 = list >= 0 ? list_ptr[list + 1] : n_atoms;
    for (int qq = 0; qq < n_q; ++qq) {
      const double x = cell_lo[3 * (size_t)c] + cell_h[c] * qpts[3 * qq], y = cell_lo[3 * (size_t)c + 1] + cell_h[c] * qpts[3 * qq + 1],
                   z = cell_lo[3 * (size_t)c + 2] + cell_h[c] * qpts[3 * qq + 2];
      double s = 0.0;
      for (int64_t a = a0; a < a1; ++a) {
        const int i = list >= 0 ? list_atoms[a] : (int)a;
        const double dx = pos[3 * (size_t)i] - x, dy = pos[3 * (size_t)i + 1] - y, dz = pos[3 * (size_t)i + 2] - z;
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        s += C * exp(-(r * r) * inv) * q[i];
      }
      rho[(size_t)c * n_q + qq] = s;
    }
  }
}

void orc_load_vector(int n_cells, const double *rho, const double *cell_h, const int32_t *cell_dofs, int n_q,
                     const double *shape, const double *weights, const double *Kref, const double *ghat, int n_dofs,
                     const int64_t *hang_ptr, const int32_t *hang_col, const double *hang_val, const uint8_t *constrained,
                     double *b) {
  memset(b, 0, sizeof(double) * n_dofs);
  for (int c = 0; c < n_cells; ++c) {
    const double h = cell_h[c], jac = h * h * h;
    double f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int qq = 0; qq < n_q; ++qq) {
      const double rw = rho[(size_t)c * n_q + qq] * (weights[qq] * jac);
      for (int i = 0; i < 8; ++i) f[i] += shape[qq * 8 + i] * rw;
    }
    const int32_t *dofs = cell_dofs + 8 * (size_t)c;
    if (Kref && ghat)
      for (int i = 0; i < 8; ++i) {
        double s = 0.0;
        for (int j = 0; j < 8; ++j) s += Kref[i * 8 + j] * ghat[dofs[j]];
        f[i] -= h * s;
      }
    for (int i = 0; i < 8; ++i) {
      const int dof = dofs[i];
      if (hang_ptr[dof + 1] > hang_ptr[dof]) {
        for (int64_t p = hang_ptr[dof]; p < hang_ptr[dof + 1]; ++p)
          if (!constrained[hang_col[p]]) b[hang_col[p]] += hang_val[p] * f[i];
      } else if (!constrained[dof]) {
        b[dof] += f[i];
      }
    }
  }
}

/* postprocess_electrostatic_energy, pair sums (src/step-50.cc:1315-1345): out[0] = sum_{i<j} q_i q_j / r,
 * out[1] = sum_{i<j} q_i q_j erfc(r / r_c) / r.  Row sums in parallel, rows added in index order. */
void orc_pair_energies(int n, const double *pos, const double *q, double r_c, double *out) {
  double *ra = (double *)malloc(sizeof(double) * (size_t)n), *rs = (double *)malloc(sizeof(double) * (size_t)n);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    double a = 0.0, s = 0.0;
    for (int j = i + 1; j < n; ++j) {
      const double dx = pos[3 * (size_t)j] - pos[3 * (size_t)i], dy = pos[3 * (size_t)j + 1] - pos[3 * (size_t)i + 1],
                   dz = pos[3 * (size_t)j + 2] - pos[3 * (size_t)i + 2];
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      const double c = q[i] * q[j] / r;
      a += c;
      s += c * erfc(r / r_c);
    }
    ra[i] = a;
    rs[i] = s;
  }
  double a = 0.0, s = 0.0;
  for (int i = 0; i < n; ++i) {
    a += ra[i];
    s += rs[i];
  }
  out[0] = a;
  out[1] = s;
  free(ra);
  free(rs);
}
