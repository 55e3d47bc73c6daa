// sm_100a kernels of the MG-PCG path: sliced-ELL SpMV with fused epilogues, the persistent
// cooperative coarse-grid CG, smoothers, fused PCG vector updates, transfers.
//
// Matrix layout (SELL-32-1, "pair-interleaved"): rows are cut into slices of 32 (one warp, one row
// per lane); a slice of width w (even) stores its w/2 entry pairs so that lane L reads the 16-byte
// pair p at  val2[slice_ptr/2 + p*32 + L]  -- every warp-wide load is one fully coalesced 512 B
// (values) / 256 B (columns) transaction, 128-bit per lane.  Padding entries have value 0.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace gmg {


template <bool NC>
__device__ __forceinline__ double ldx(const double *x, int c) {
  if (NC) return __ldg(x + c);
  return x[c];
}

// sum_j a_rj x_j for the row (slice, lane); j ascending = CSR order, one FMA chain.
template <bool NC>
__device__ __forceinline__ double sell_row_dot(const SellView &A, int slice, int lane, const double *__restrict__ x) {
  const int64_t b = A.slice_ptr[slice];
  const int npairs = (int)((A.slice_ptr[slice + 1] - b) >> 6);
  const double2 *v2 = reinterpret_cast<const double2 *>(A.val) + (b >> 1) + lane;
  const int2 *c2 = reinterpret_cast<const int2 *>(A.col) + (b >> 1) + lane;
  double acc = 0.0;
  int p = 0;
  for (; p + 4 <= npairs; p += 4) {
    double2 v[4];
    int2 c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = ld_stream_d2(v2 + (p + u) * 32);
      c[u] = ld_stream_i2(c2 + (p + u) * 32);
    }
    double xv[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      xv[2 * u] = ldx<NC>(x, c[u].x);
      xv[2 * u + 1] = ldx<NC>(x, c[u].y);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc = fma(v[u].x, xv[2 * u], acc);
      acc = fma(v[u].y, xv[2 * u + 1], acc);
    }
  }
  for (; p < npairs; ++p) {
    const double2 v = ld_stream_d2(v2 + p * 32);
    const int2 c = ld_stream_i2(c2 + p * 32);
    acc = fma(v.x, ldx<NC>(x, c.x), acc);
    acc = fma(v.y, ldx<NC>(x, c.y), acc);
  }
  return acc;
}

// ---- compressed format ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// sum_j a_rj x_j for row (slice, lane) of a CSELL matrix; `dict` points at the value dictionary (shared or global).
// Same entry order and the same FMA chain as sell_row_dot: bit-identical results.
template <bool NC>
__device__ __forceinline__ double csell_row_dot(const CsellView &A, int slice, int lane, const double *__restrict__ x,
                                                const double *__restrict__ dict) {
  const int64_t b = A.slice_ptr[slice];
  const int nchunks = (int)((A.slice_ptr[slice + 1] - b) >> 7);
  const uint4 *e4 = reinterpret_cast<const uint4 *>(A.ent) + (b >> 2) + lane;
  const int row = slice * 32 + lane;
  if (row >= A.n_rows) return 0.0;  // lanes past the last row of the last slice: their offsets are relative to nothing
  double acc = 0.0;
  // Latency-bound on the entry stream (ncu: about half of all stall samples sat on the first use of an entry word):
  // entries are consumed in groups of 4 chunks (16 entries) whose loads are all issued first, and the caller
  // prefetches the next slice of this warp into L2 (csell_prefetch_slice) while this one is processed.
  for (int p0 = 0; p0 < nchunks; p0 += 4) {
    uint4 e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      e[u] = (p0 + u < nchunks) ? ld_stream_u4(e4 + (p0 + u) * 32) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (p0 + u < nchunks) {
        const uint32_t w[4] = {e[u].x, e[u].y, e[u].z, e[u].w};
        double xv[4], av[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          xv[t] = ldx<NC>(x, row + (int)(short)(w[t] & 0xffffu));
          av[t] = dict[w[t] >> 16];
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) acc = fma(av[t], xv[t], acc);
      }
    }
  }
  return acc;
}

// one 128-byte line per lane: the whole entry block of a slice (<= 32 lines = 4 KB) goes to L2 with one instruction
__device__ __forceinline__ void csell_prefetch_slice(const CsellView &A, int slice, int lane) {
  if (slice >= A.n_slices) return;
  const int64_t b = A.slice_ptr[slice];
  const int64_t bytes = (A.slice_ptr[slice + 1] - b) * 4;
  const char *p = reinterpret_cast<const char *>(A.ent) + b * 4 + (int64_t)lane * 128;
  if ((int64_t)lane * 128 < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// open-addressing set of fp64 bit patterns (distinct matrix values); table size is a power of two
__global__ void value_set_insert(int64_t n, const double *__restrict__ val, unsigned long long *table, int mask,
                                 int *count, int limit) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = (unsigned long long)__double_as_longlong(val[i]) + 1ull;  // 0 = empty slot
  unsigned int hsh = (unsigned int)((key * 0x9E3779B97F4A7C15ull) >> 40) & mask;
  for (int probe = 0; probe <= mask; ++probe) {
    const unsigned long long cur = table[hsh];
    if (cur == key) return;
    if (cur == 0ull) {
      if (*(volatile int *)count > limit) return;
      const unsigned long long old = atomicCAS(table + hsh, 0ull, key);
      if (old == 0ull) {
        atomicAdd(count, 1);
        return;
      }
      if (old == key) return;
    }
    hsh = (hsh + 1) & mask;
  }
}

// SELL (pair layout) -> CSELL; slot_code maps hash-table slots to dictionary codes.  flags[0] |= 1 on failure.
__global__ void sell_to_csell(SellView A, const int64_t *__restrict__ cslice_ptr, const unsigned long long *__restrict__ table,
                              const unsigned short *__restrict__ slot_code, int mask, unsigned short zero_code,
                              uint32_t *__restrict__ ent, int *flags) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int slice = r >> 5, lane = r & 31;
  if (slice >= A.n_slices) return;
  const int64_t b = A.slice_ptr[slice];
  const int w = (int)((A.slice_ptr[slice + 1] - b) >> 5);
  const int64_t cb = cslice_ptr[slice];
  const int cw = (int)((cslice_ptr[slice + 1] - cb) >> 5);
  for (int j = 0; j < cw; ++j) {
    uint32_t word = ((uint32_t)zero_code << 16);
    if (j < w && r < A.n_rows) {
      const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
      const int delta = A.col[at] - r;
      if (delta < -32768 || delta > 32767) atomicOr(flags, 1);
      const unsigned long long key = (unsigned long long)__double_as_longlong(A.val[at]) + 1ull;
      unsigned int hsh = (unsigned int)((key * 0x9E3779B97F4A7C15ull) >> 40) & mask;
      int probe = 0;
      while (table[hsh] != key && probe <= mask) {
        hsh = (hsh + 1) & mask;
        ++probe;
      }
      if (probe > mask) atomicOr(flags, 1);
      word = ((uint32_t)slot_code[hsh] << 16) | (uint32_t)(unsigned short)(short)delta;
    }
    ent[cb + (int64_t)(j >> 2) * 128 + lane * 4 + (j & 3)] = word;
  }
}

// ------------------------------------------------------------------------------------------------
// CSR -> SELL conversion (device side; the host hands over plain CSR)
// ------------------------------------------------------------------------------------------------
// `rows` (optional): local row r is CSR row rows[r] (sub-matrices of one colour / wavefront)
__global__ void csr_slice_widths(int n_rows, int n_slices, const int *__restrict__ rows, const int64_t *__restrict__ rowptr,
                                 const double *__restrict__ val, double drop_tol, int *__restrict__ width,
                                 unsigned long long *__restrict__ total_nnz) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  if (r < n_rows) {
    const int src = rows ? rows[r] : r;
    if (drop_tol < 0.0) {
      cnt = (int)(rowptr[src + 1] - rowptr[src]);
    } else {
      for (int64_t k = rowptr[src]; k < rowptr[src + 1]; ++k) cnt += (fabs(val[k]) > drop_tol) ? 1 : 0;
    }
  }
  int m = cnt, sum = cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if ((threadIdx.x & 31) == 0 && (r >> 5) < n_slices) {
    width[r >> 5] = (m + 1) & ~1;
    if (sum) atomicAdd(total_nnz, (unsigned long long)sum);
  }
}

__global__ void csr_to_sell(int n_rows, int n_cols, const int *__restrict__ rows, const int64_t *__restrict__ rowptr,
                            const int *__restrict__ col, const double *__restrict__ val, double drop_tol,
                            const int64_t *__restrict__ slice_ptr, double *__restrict__ sval, int *__restrict__ scol) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int slice = r >> 5, lane = r & 31;
  if (slice >= (n_rows + 31) / 32) return;
  const int64_t b = slice_ptr[slice];
  const int w = (int)((slice_ptr[slice + 1] - b) >> 5);
  int j = 0;
  const int pad_col = rows ? 0 : ((r < n_cols) ? r : 0);
  if (r < n_rows) {
    const int src = rows ? rows[r] : r;
    for (int64_t k = rowptr[src]; k < rowptr[src + 1]; ++k) {
      const double v = val[k];
      if (drop_tol >= 0.0 && !(fabs(v) > drop_tol)) continue;
      const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
      sval[at] = v;
      scol[at] = col[k];
      ++j;
    }
  }
  for (; j < w; ++j) {
    const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
    sval[at] = 0.0;
    scol[at] = pad_col;
  }
}

// CSR transpose on the device (R = P^T): 64-bit keys (column << 32 | row) of all entries, radix-sorted with their
// values; the transposed row pointer from the per-column counts.  Entries of a transposed row come out in ascending
// original-row order: the same (deterministic) order as a sequential transpose.
__global__ void csr_transpose_keys(int n_rows, const int64_t *__restrict__ rowptr, const int *__restrict__ col,
                                   unsigned long long *__restrict__ key, unsigned long long *__restrict__ count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
    key[k] = ((unsigned long long)(unsigned int)col[k] << 32) | (unsigned int)r;
    atomicAdd(count + col[k], 1ull);
  }
}
__global__ void csr_transpose_cols(int64_t nnz, const unsigned long long *__restrict__ key, int *__restrict__ col) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k < nnz) col[k] = (int)(key[k] & 0xffffffffull);
}

// val(i, j) += v for the entries of a (small) CSR matrix whose pattern is contained in the SELL matrix's:
// A + I of a level without a host round trip.  *flag != 0 if an entry has no slot.
__global__ void sell_add_csr(SellView A, double *__restrict__ aval, int n_rows, const int64_t *__restrict__ rowptr,
                             const int *__restrict__ col, const double *__restrict__ val, int *flag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int64_t k0 = rowptr[r], k1 = rowptr[r + 1];
  if (k0 == k1) return;
  const int slice = r >> 5, lane = r & 31;
  const int64_t b = A.slice_ptr[slice];
  const int w = (int)((A.slice_ptr[slice + 1] - b) >> 5);
  for (int64_t k = k0; k < k1; ++k) {
    const int c = col[k];
    bool found = false;
    for (int j = 0; j < w && !found; ++j) {
      const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
      if (A.col[at] == c) {
        aval[at] += val[k];
        found = true;
      }
    }
    if (!found) atomicOr(flag, 1);
  }
}

__global__ void sell_extract_diag_inv(SellView A, double *__restrict__ dinv) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n_rows) return;
  const int slice = r >> 5, lane = r & 31;
  const int64_t b = A.slice_ptr[slice];
  const int w = (int)((A.slice_ptr[slice + 1] - b) >> 5);
  double d = 0.0;
  for (int j = 0; j < w; ++j) {
    const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
    if (A.col[at] == r) d += A.val[at];
  }
  dinv[r] = 1.0 / d;
}

// matrix norms: column abs-sums via atomics into colsum, per-block max row abs-sum and sum of squares
__global__ void __launch_bounds__(256) sell_norm_partials(SellView A, double *colsum, double *rowmax_partial,
                                                          double *frob_partial) {
  __shared__ double red[32];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double rs = 0.0, fs = 0.0;
  if (r < A.n_rows) {
    const int slice = r >> 5, lane = r & 31;
    const int64_t b = A.slice_ptr[slice];
    const int w = (int)((A.slice_ptr[slice + 1] - b) >> 5);
    for (int j = 0; j < w; ++j) {
      const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
      const double a = fabs(A.val[at]);
      rs += a;
      fs += a * a;
      if (a != 0.0) atomicAdd(colsum + A.col[at], a);
    }
  }
  const double m = block_max(rs, red);
  if (threadIdx.x == 0) rowmax_partial[blockIdx.x] = m;
  const double f = block_sum(fs, red);
  if (threadIdx.x == 0) frob_partial[blockIdx.x] = f;
}

__global__ void __launch_bounds__(256) vec_max_partials(int n, const double *__restrict__ v, double *pmax) {
  __shared__ double red[32];
  double m = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmax(m, fabs(v[i]));
  const double sm = block_max(m, red);
  if (threadIdx.x == 0) pmax[blockIdx.x] = sm;
}

// ------------------------------------------------------------------------------------------------
// SpMV with fused epilogues
// ------------------------------------------------------------------------------------------------
enum { EPI_ASSIGN = 0,   // y = A x
       EPI_ADD = 1,      // y += A x
       EPI_SUB = 2,      // y -= A x
       EPI_RESID = 3,    // y = b - A x
       EPI_NRESID = 4,   // y = A x - b            (SolverCG: g = A x - b)
       EPI_JACOBI = 5 }; // y = x + omega * dinv * (b - A x)   (one damped Jacobi step, x != y)

enum { DOT_NONE = 0, DOT_XY = 1 /* sum x_r y_r */, DOT_YY = 2 /* sum y_r^2 */ };

template <int EPI, int DOT>
__global__ void __launch_bounds__(256) sell_spmv(SellView A, const double *__restrict__ x, double *__restrict__ y,
                                                 const double *__restrict__ b, const double *__restrict__ dinv,
                                                 double omega, double *partials, unsigned int *counter, double *out) {
  __shared__ double red[32];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int slice = r >> 5, lane = r & 31;
  double contrib = 0.0;
  if (slice < A.n_slices) {
    const double ax = sell_row_dot<true>(A, slice, lane, x);
    if (r < A.n_rows) {
      double yv;
      if (EPI == EPI_ASSIGN) yv = ax;
      else if (EPI == EPI_ADD) yv = y[r] + ax;
      else if (EPI == EPI_SUB) yv = y[r] - ax;
      else if (EPI == EPI_RESID) yv = b[r] - ax;
      else if (EPI == EPI_NRESID) yv = ax - b[r];
      else yv = x[r] + omega * dinv[r] * (b[r] - ax);
      y[r] = yv;
      if (DOT == DOT_XY) contrib = x[r] * yv;
      if (DOT == DOT_YY) contrib = yv * yv;
    }
  }
  if (DOT != DOT_NONE) {
    const double s = block_sum(contrib, red);
    grid_sum_finalize(s, partials, counter, out, red);
  }
}

// rows of one colour, in place Gauss-Seidel update: u_i += omega (rhs_i - sum_j a_ij u_j) / a_ii
// (A_c holds the rows of this colour; rows[k] is the global row of local row k; dinv global).
// A colour of a patch level is a few thousand rows: the launch is latency-bound, not bandwidth-bound.  Eight lanes
// share a row: every lane loads its part of the row's entries and operands at once (two memory round trips per row
// instead of two per four entry pairs), then the FMA chain runs through the eight lanes in entry order, the
// partial sum handed on by shuffle: same order, same bits as one lane per row.
// (k: local row of the colour or out of range; sub: this lane's position among the eight lanes of the row; all
// 32 lanes of a warp must call it)
// PDL: the kernel was launched with programmatic stream serialisation (the previous colour may still be running):
// the matrix entries, row indices and diagonals (never written during a solve) are fetched first, then
// griddepcontrol.wait orders everything that touches u / rhs after the previous kernel; u is read past L1.
template <bool PDL>
__device__ __forceinline__ void color_relax_row8(const SellView &Ac, const int *__restrict__ rows, double *u,
                                                 const double *__restrict__ rhs, const double *__restrict__ dinv, double omega,
                                                 int k, int sub) {
  const bool on = k < Ac.n_rows;
  const int slice = k >> 5, rlane = k & 31;
  int64_t b = 0;
  int npairs = 0;
  int i = 0;
  double dv = 0.0;
  if (on) {
    b = Ac.slice_ptr[slice];
    npairs = (int)((Ac.slice_ptr[slice + 1] - b) >> 6);
    if (sub == 0) {
      i = rows[k];
      dv = dinv[i];
    }
  }
  const double2 *v2 = reinterpret_cast<const double2 *>(Ac.val) + (b >> 1) + rlane;
  const int2 *c2 = reinterpret_cast<const int2 *>(Ac.col) + (b >> 1) + rlane;
  const int pmax = __reduce_max_sync(0xffffffffu, npairs);
  double acc = 0.0;
  bool waited = !PDL;
  for (int base = 0; base < pmax; base += 32) {  // 32 pairs per round, 4 per lane
    const int n_here = min(max(npairs - base, 0), 32);
    const int chunk = (n_here + 7) >> 3;
    const int p0 = base + sub * chunk;
    const int cnt = max(min(chunk, n_here - sub * chunk), 0);
    double2 v[4];
    int2 c[4];
    double x0[4], x1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < cnt) {
        v[q] = ld_stream_d2(v2 + (p0 + q) * 32);
        c[q] = ld_stream_i2(c2 + (p0 + q) * 32);
      }
    if (!waited) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      waited = true;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < cnt) {
        x0[q] = PDL ? __ldcg(u + c[q].x) : u[c[q].x];
        x1[q] = PDL ? __ldcg(u + c[q].y) : u[c[q].y];
      }
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      if (sub == cc) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < cnt) {
            acc = fma(v[q].x, x0[q], acc);
            acc = fma(v[q].y, x1[q], acc);
          }
      }
      acc = __shfl_sync(0xffffffffu, acc, (threadIdx.x & 24) | cc);
    }
  }
  if (!waited) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (on && sub == 0) {
    const double ui = PDL ? __ldcg(u + i) : u[i];
    u[i] = ui + omega * (rhs[i] - acc) * dv;
  }
}

__global__ void __launch_bounds__(256) sell_color_relax(SellView Ac, const int *__restrict__ rows, double *u,
                                                        const double *__restrict__ rhs, const double *__restrict__ dinv,
                                                        double omega) {
  asm volatile("griddepcontrol.launch_dependents;");  // (a following programmatic launch may start its prologue)
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  color_relax_row8<false>(Ac, rows, u, rhs, dinv, omega, t >> 3, t & 7);
}

// the same launched with programmatic stream serialisation after another relaxation kernel: lets the next colour
// start launching at once and fetches this colour's matrix entries while the previous colour is still running
__global__ void __launch_bounds__(256) sell_color_relax_pdl(SellView Ac, const int *__restrict__ rows, double *u,
                                                            const double *__restrict__ rhs, const double *__restrict__ dinv,
                                                            double omega) {
  asm volatile("griddepcontrol.launch_dependents;");
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  color_relax_row8<true>(Ac, rows, u, rhs, dinv, omega, t >> 3, t & 7);
}

// One smooth() call of the multicolour / level-scheduled SSOR as ONE cooperative launch: colours (or wavefronts)
// are separated by grid.sync() instead of kernel boundaries -- on the patch levels a colour is a few microseconds
// of work, so launch gaps dominated (ncu: 896 launches = 19 % of a step before this kernel).
struct ColorView {
  SellView A;
  const int *rows;
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) ssor_persistent(const ColorView *__restrict__ fwd, int n_fwd,
                                                         const ColorView *__restrict__ bwd, int n_bwd, int bwd_reversed,
                                                         int n, double *u, const double *__restrict__ rhs,
                                                         const double *__restrict__ dinv, double omega, int steps,
                                                         int zero_start) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  if (zero_start) {
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n; i += gridDim.x * BLOCK) u[i] = 0.0;
    grid.sync();
  }
  for (int s = 0; s < steps; ++s)
    for (int pass = 0; pass < 2; ++pass) {
      const ColorView *set = pass == 0 ? fwd : bwd;
      const int nc = pass == 0 ? n_fwd : n_bwd;
      for (int k = 0; k < nc; ++k) {
        const ColorView &C = set[(pass == 1 && bwd_reversed) ? nc - 1 - k : k];
        // eight lanes per row (color_relax_row8); whole warps iterate together
        const int rows_per_sweep = (gridDim.x * BLOCK) >> 3;
        for (int r0 = 0; r0 < C.A.n_rows; r0 += rows_per_sweep)
          color_relax_row8<false>(C.A, C.rows, u, rhs, dinv, omega, r0 + ((blockIdx.x * BLOCK + threadIdx.x) >> 3), threadIdx.x & 7);
        grid.sync();
      }
    }
}

// The same smooth() call for SMALL levels (a few thousand rows per colour) as ONE thread-block cluster: the colours
// are separated by the hardware cluster barrier (~0.3 us) instead of kernel boundaries (~3 us in a CUDA graph) or
// grid.sync() (~2 us).  The barrier's acquire at cluster scope invalidates the L1s, so the plain loads of u see
// the other blocks' updates.  Launched with cluster dimension = grid dimension (<= 8 blocks).
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) ssor_cluster(const ColorView *__restrict__ fwd, int n_fwd,
                                                      const ColorView *__restrict__ bwd, int n_bwd, int bwd_reversed, int n,
                                                      double *u, const double *__restrict__ rhs,
                                                      const double *__restrict__ dinv, double omega, int steps, int zero_start) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int nthreads = (int)cluster.num_blocks() * BLOCK;
  const int tid = (int)cluster.block_rank() * BLOCK + threadIdx.x;
  if (zero_start) {
    for (int i = tid; i < n; i += nthreads) u[i] = 0.0;
    cluster.sync();
  }
  for (int s = 0; s < steps; ++s)
    for (int pass = 0; pass < 2; ++pass) {
      const ColorView *set = pass == 0 ? fwd : bwd;
      const int nc = pass == 0 ? n_fwd : n_bwd;
      for (int k = 0; k < nc; ++k) {
        const ColorView &C = set[(pass == 1 && bwd_reversed) ? nc - 1 - k : k];
        for (int r0 = 0; r0 < C.A.n_rows; r0 += nthreads >> 3)
          color_relax_row8<false>(C.A, C.rows, u, rhs, dinv, omega, r0 + (tid >> 3), threadIdx.x & 7);
        cluster.sync();
      }
    }
}

// ------------------------------------------------------------------------------------------------
// vector kernels
// ------------------------------------------------------------------------------------------------

__global__ void vec_scale_dinv(int n, double omega, const double *__restrict__ dinv, const double *__restrict__ r,
                               double *__restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = omega * dinv[i] * r[i];
}

// y = a x + b y
__global__ void vec_axpby(int n, double a, const double *__restrict__ x, double b, double *__restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i] + b * y[i];
}

// Chebyshev: inc = c1 inc + c2 dinv r ; u += inc
__global__ void cheb_update(int n, double c1, double c2, const double *__restrict__ dinv, const double *__restrict__ r,
                            double *__restrict__ inc, double *__restrict__ u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = c1 * inc[i] + c2 * dinv[i] * r[i];
    inc[i] = v;
    u[i] += v;
  }
}

__global__ void vec_gather(int n, const int *__restrict__ dst_idx, const int *__restrict__ src_idx,
                           const double *__restrict__ src, double *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[dst_idx[i]] = src[src_idx[i]];
}

// d = -h ; gh[slot] = g.h
__global__ void __launch_bounds__(256) pcg_init_direction(int n, const double *__restrict__ g, const double *__restrict__ h,
                                                          double *__restrict__ d, PcgScalars *s, int slot,
                                                          double *partials, unsigned int *counter) {
  __shared__ double red[32];
  double c = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double hv = h[i];
    d[i] = -hv;
    c += g[i] * hv;
  }
  const double bs = block_sum(c, red);
  grid_sum_finalize(bs, partials, counter, &s->gh[slot], red);
}

// alpha = gh[slot]/dh ; x += alpha d ; g += alpha h ; res2 = g.g      (Vector::add + add_and_dot)
__global__ void __launch_bounds__(256) pcg_update(int n, double *__restrict__ x, double *__restrict__ g,
                                                  const double *__restrict__ d, const double *__restrict__ h,
                                                  PcgScalars *s, int slot, double *partials, unsigned int *counter) {
  __shared__ double red[32];
  const double alpha = s->gh[slot] / s->dh;
  double c = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] += alpha * d[i];
    const double gv = g[i] + alpha * h[i];
    g[i] = gv;
    c += gv * gv;
  }
  const double bs = block_sum(c, red);
  grid_sum_finalize(bs, partials, counter, &s->res2, red);
}

// out = a.b
__global__ void __launch_bounds__(256) vec_dot(int n, const double *__restrict__ a, const double *__restrict__ b,
                                               double *out, double *partials, unsigned int *counter) {
  __shared__ double red[32];
  double c = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += a[i] * b[i];
  const double bs = block_sum(c, red);
  grid_sum_finalize(bs, partials, counter, out, red);
}

// beta = gh[slot_new] / gh[slot_old] ; d = beta d - h
__global__ void pcg_new_direction(int n, double *__restrict__ d, const double *__restrict__ h, const PcgScalars *s,
                                  int slot_new) {
  const double beta = s->gh[slot_new] / s->gh[slot_new ^ 1];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = beta * d[i] - h[i];
}

// l1 / l2^2 / linf partials
__global__ void __launch_bounds__(256) vec_norm_partials(int64_t n, const double *__restrict__ v, double *p1, double *p2,
                                                         double *pinf) {
  __shared__ double red[32];
  double a = 0.0, b = 0.0, m = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = fabs(v[i]);
    a += x;
    b += x * x;
    m = fmax(m, x);
  }
  const double sa = block_sum(a, red);
  if (threadIdx.x == 0) p1[blockIdx.x] = sa;
  const double sb = block_sum(b, red);
  if (threadIdx.x == 0) p2[blockIdx.x] = sb;
  const double sm = block_max(m, red);
  if (threadIdx.x == 0) pinf[blockIdx.x] = sm;
}

// ------------------------------------------------------------------------------------------------
// persistent cooperative CG (MGCoarseGridIterativeSolver<SolverCG, PreconditionIdentity>):
// the whole solve is ONE launch; all scalars stay on the device; dot products are reduced through
// per-block partials + grid.sync(), summed in a fixed order by every block.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ double grid_total(const double *partials, int nblocks, double *bc) {
  // every block sums the same partials in the same order
  if (threadIdx.x < 32) {
    const double s = warp_sum_partials(partials, nblocks);
    if (threadIdx.x == 0) *bc = s;
  }
  __syncthreads();
  const double r = *bc;
  __syncthreads();
  return r;
}

constexpr int CSELL_SMEM_DICT = 2048;  // dictionary entries staged in shared memory (16 KB)

// Row kernels of the persistent CG kernels, per matrix format.  init() stages what the format keeps in shared memory;
// operator() is sum_j a_rj x_j of row (slice, lane); valid() tells whether that row belongs to the slice pass at all
// (the row-pattern format handles its rare rows in remainder()).
template <class MAT>
struct RowSmem {
  double unused[1];
};
template <>
struct RowSmem<CsellView> {
  double dict[CSELL_SMEM_DICT];
};
template <class MAT>
struct RowDot;
template <>
struct RowDot<SellView> {
  __device__ __forceinline__ void init(const SellView &, RowSmem<SellView> &) {}
  __device__ __forceinline__ void prefetch(const SellView &, int, int) const {}
  __device__ __forceinline__ double operator()(const SellView &A, int s, int lane, const double *x) {
    return sell_row_dot<false>(A, s, lane, x);
  }
  __device__ __forceinline__ bool valid() const { return true; }
  template <class F>
  __device__ __forceinline__ void remainder(const SellView &, int, int, int, int, const double *, F &&) const {}
};
template <>
struct RowDot<CsellView> {
  const double *dict;
  __device__ __forceinline__ void init(const CsellView &A, RowSmem<CsellView> &sm) {
    if (A.dict_n <= CSELL_SMEM_DICT) {
      for (int i = threadIdx.x; i < A.dict_n; i += blockDim.x) sm.dict[i] = A.dict[i];
      __syncthreads();
      dict = sm.dict;
    } else {
      dict = A.dict;
    }
  }
  __device__ __forceinline__ void prefetch(const CsellView &A, int s, int lane) const { csell_prefetch_slice(A, s, lane); }
  __device__ __forceinline__ double operator()(const CsellView &A, int s, int lane, const double *x) {
    return csell_row_dot<false>(A, s, lane, x, dict);
  }
  __device__ __forceinline__ bool valid() const { return true; }
  template <class F>
  __device__ __forceinline__ void remainder(const CsellView &, int, int, int, int, const double *, F &&) const {}
};

template <int BLOCK, class MAT>
__global__ void __launch_bounds__(BLOCK, 2) cg_persistent(MAT A, const double *__restrict__ b, double *x, double *g,
                                                       double *d, double *h, double *partials /* 3 * gridDim.x */,
                                                       int max_it, double tol, CgResult *result) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[32];
  __shared__ double bc;
  __shared__ RowSmem<MAT> row_smem;
  RowDot<MAT> row_dot;
  row_dot.init(A, row_smem);
  const int nb = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int WPB = BLOCK / 32;
  // contiguous chunk of slices per block, warps stride through it
  const int s_begin = (int)(((int64_t)A.n_slices * blockIdx.x) / nb);
  const int s_end = (int)(((int64_t)A.n_slices * (blockIdx.x + 1)) / nb);
  double *pa = partials, *pb = partials + nb, *pc = partials + 2 * nb;

  // x = 0, g = -b, d = -g ; res0 = ||g||
  double acc = 0.0;
  for (int s = s_begin + warp; s < s_end; s += WPB) {
    const int r = s * 32 + lane;
    if (r < A.n_rows) {
      const double bv = b[r];
      x[r] = 0.0;
      g[r] = -bv;
      d[r] = bv;
      acc += bv * bv;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) pc[blockIdx.x] = acc;
  grid.sync();
  double res2 = grid_total(pc, nb, &bc);
  double res = sqrt(res2);
  const double res0 = res;
  int it = 0, status = 0;
  if (res > tol) {
    double gh = res * res;
    while (true) {
      ++it;
      // h = A d ; dh = d.h
      acc = 0.0;
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        if (s + WPB < s_end) row_dot.prefetch(A, s + WPB, lane);
        const double ad = row_dot(A, s, lane, d);
        const int r = s * 32 + lane;
        if (r < A.n_rows && row_dot.valid()) {
          h[r] = ad;
          acc += d[r] * ad;
        }
      }
      row_dot.remainder(A, blockIdx.x, nb, warp, WPB, d, [&](int r, double ad) {
        h[r] = ad;
        acc += d[r] * ad;
      });
      acc = block_sum(acc, red);
      if (threadIdx.x == 0) pa[blockIdx.x] = acc;
      grid.sync();
      const double alpha = gh / grid_total(pa, nb, &bc);
      // x += alpha d ; g += alpha h ; res2 = g.g
      acc = 0.0;
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        const int r = s * 32 + lane;
        if (r < A.n_rows) {
          x[r] += alpha * d[r];
          const double gv = g[r] + alpha * h[r];
          g[r] = gv;
          acc += gv * gv;
        }
      }
      acc = block_sum(acc, red);
      if (threadIdx.x == 0) pb[blockIdx.x] = acc;
      grid.sync();
      res2 = grid_total(pb, nb, &bc);
      res = sqrt(res2);
      if (res <= tol) break;
      if (it >= max_it) { status = 1; break; }
      const double beta = res2 / gh;
      gh = res2;
      // d = beta d - g
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        const int r = s * 32 + lane;
        if (r < A.n_rows) d[r] = beta * d[r] - g[r];
      }
      grid.sync();
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    result->iterations = it;
    result->status = status;
    result->res0 = res0;
    result->res = res;
  }
}

}  // namespace gmg

#include "pattern.cuh"
#include "pattern_win.cuh"
#include "pattern_win2.cuh"
