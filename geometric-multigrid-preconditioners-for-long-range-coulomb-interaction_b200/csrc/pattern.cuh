// Row-pattern dictionary format ("RPD"): a lossless representation of FE matrices on (mostly) uniform meshes.
//
// A Q1 stiffness matrix on a uniform lattice has a handful of distinct ROWS once a row is written as the sequence
// of (column - row, value) pairs: the interior stencil, the boundary-adjacent variants, identity rows.  The matrix is
// stored as one 32-bit pattern id per row plus the table of distinct patterns.  An SpMV then streams 4 bytes per ROW
// instead of 12 bytes (CSR) or 4 bytes (CSELL) per ENTRY: the level-0 matrix of the 64k-atom case (47 M entries)
// shrinks from 565 MB to 7 MB, the coarse-grid CG's whole working set (4 vectors) fits the 126 MB L2, and the kernel
// is bound by L1/L2 bandwidth and the grid barriers instead of HBM.
//
// Exactness: a pattern keeps the nonzero entries of its rows in CSR order, and the row dot product is the same
// single FMA chain as the plain formats.  Entries that are exactly +-0 are left out: fma(0, x, acc) == acc for every
// finite x (acc starts at +0 and a sum that cancels exactly rounds to +0, so the sign of a zero acc never differs),
// hence results are bit-identical to SELL / CSELL for finite vectors.  Every row is verified against its pattern
// entry by entry after the build (hash collisions cannot go unnoticed); on any mismatch, or when a matrix has too
// many distinct rows to pay off, the caller falls back to CSELL / SELL.
#pragma once
#include "common.cuh"

namespace gmg {

constexpr int PAT_MAX_ENT = 3584;   // entries of the pattern table: always staged in shared memory (42 KB)
constexpr int PAT_MAX_PAT = 1023;   // frequent patterns (+ the empty one)

__device__ __forceinline__ uint64_t pat_mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33;
  return h;
}

// nonzero entries of row r of a SELL matrix, in stored (CSR) order
template <class F>
__device__ __forceinline__ void sell_row_foreach_nonzero(const SellView &A, int r, F &&f) {
  const int slice = r >> 5, lane = r & 31;
  const int64_t b = A.slice_ptr[slice];
  const int w = (int)((A.slice_ptr[slice + 1] - b) >> 5);
  for (int j = 0; j < w; ++j) {
    const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
    const double v = A.val[at];
    if (v != 0.0) f(A.col[at] - r, v);
  }
}

// hash of every row's (offset, value) sequence; 0 is never produced (0 = empty slot of the table)
__global__ void pat_row_hash(SellView A, uint64_t *__restrict__ hash) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n_rows) return;
  uint64_t hsh = 0x243F6A8885A308D3ull;
  sell_row_foreach_nonzero(A, r, [&](int off, double v) {
    hsh = pat_mix(hsh, (uint64_t)(uint32_t)off);
    hsh = pat_mix(hsh, (uint64_t)__double_as_longlong(v));
  });
  hash[r] = hsh ? hsh : 1ull;
}

// Insert the row hashes into an open-addressing table; per slot: smallest row with that hash (the pattern's
// representative) and the number of rows.  Warps aggregate equal keys so the interior pattern (~all rows) does not
// serialise on one address.  *count > limit aborts (too many distinct rows).
__global__ void pat_table_insert(int n_rows, const uint64_t *__restrict__ hash, unsigned long long *keys, int *rep,
                                 int *cnt, int mask, int *count, int limit) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = r < n_rows;
  const unsigned long long key = active ? hash[r] : 0ull;
  const unsigned int peers = __match_any_sync(0xffffffffu, key);
  if (!active) return;
  const int leader = __ffs(peers) - 1;  // lowest lane = smallest row of the group
  if ((int)(threadIdx.x & 31) != leader) return;
  unsigned int slot = (unsigned int)(key >> 20) & mask;
  for (int probe = 0; probe <= mask; ++probe) {
    unsigned long long cur = keys[slot];
    if (cur == 0ull) {
      if (*(volatile int *)count > limit) return;
      cur = atomicCAS(keys + slot, 0ull, key);
      if (cur == 0ull) {
        atomicAdd(count, 1);
        cur = key;
      }
    }
    if (cur == key) {
      atomicMin(rep + slot, r);
      atomicAdd(cnt + slot, __popc(peers));
      return;
    }
    slot = (slot + 1) & mask;
  }
}

// number of nonzero entries of each pattern's representative row
__global__ void pat_widths(SellView A, int n_pat, const int *__restrict__ rep, int *__restrict__ width) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pat) return;
  int w = 0;
  sell_row_foreach_nonzero(A, rep[p], [&](int, double) { ++w; });
  width[p] = w;
}

__global__ void pat_fill(SellView A, int n_pat, const int *__restrict__ rep, const int *__restrict__ ptr,
                         int *__restrict__ off, double *__restrict__ val) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pat) return;
  int k = ptr[p];
  sell_row_foreach_nonzero(A, rep[p], [&](int o, double v) {
    off[k] = o;
    val[k] = v;
    ++k;
  });
}

// pattern id of every row + entry-by-entry verification against the table (flag |= 1 on any difference).
// Rows whose pattern is not in the table get the empty pattern and irregular[r] = 1: they go to the remainder.
__global__ void pat_assign_verify(SellView A, const uint64_t *__restrict__ hash, const unsigned long long *__restrict__ keys,
                                  const int *__restrict__ slot_pid, int mask, const int *__restrict__ ptr,
                                  const int *__restrict__ off, const double *__restrict__ val, int empty_pid,
                                  uint32_t *__restrict__ pat, unsigned char *__restrict__ irregular, int *flag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n_slices * 32) return;
  if (r >= A.n_rows) {
    pat[r] = (uint32_t)empty_pid;
    return;
  }
  const unsigned long long key = hash[r];
  unsigned int slot = (unsigned int)(key >> 20) & mask;
  int probe = 0;
  while (keys[slot] != key && probe <= mask) {
    slot = (slot + 1) & mask;
    ++probe;
  }
  if (probe > mask) {
    atomicOr(flag, 1);
    pat[r] = (uint32_t)empty_pid;
    irregular[r] = 1;
    return;
  }
  const int pid = slot_pid[slot];
  pat[r] = (uint32_t)pid;
  irregular[r] = pid == empty_pid ? 1 : 0;
  if (pid == empty_pid) return;
  int k = ptr[pid];
  const int k1 = ptr[pid + 1];
  bool ok = true;
  sell_row_foreach_nonzero(A, r, [&](int o, double v) {
    if (k >= k1 || off[k] != o || __double_as_longlong(val[k]) != __double_as_longlong(v)) ok = false;
    ++k;
  });
  if (!ok || k != k1) atomicOr(flag, 1);
}

// remainder rows as their own SELL matrix (zeros dropped), cut out of the source SELL matrix: slice widths ...
__global__ void sell_sub_widths(SellView A, int n_sub, const int *__restrict__ rows, int *__restrict__ width) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  int w = 0;
  if (k < n_sub) sell_row_foreach_nonzero(A, rows[k], [&](int, double) { ++w; });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
  if ((threadIdx.x & 31) == 0 && (k >> 5) < (n_sub + 31) / 32) width[k >> 5] = (w + 1) & ~1;
}

// ... and entries (pair-interleaved layout of kernels.cuh; padding: value 0, column = the row itself)
__global__ void sell_sub_fill(SellView A, int n_sub, const int *__restrict__ rows, const int64_t *__restrict__ slice_ptr,
                              double *__restrict__ sval, int *__restrict__ scol) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int slice = k >> 5, lane = k & 31;
  if (slice >= (n_sub + 31) / 32) return;
  const int64_t b = slice_ptr[slice];
  const int w = (int)((slice_ptr[slice + 1] - b) >> 5);
  int j = 0;
  int pad_col = 0;
  if (k < n_sub) {
    const int r = rows[k];
    pad_col = r < A.n_cols ? r : 0;
    sell_row_foreach_nonzero(A, r, [&](int o, double v) {
      const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
      sval[at] = v;
      scol[at] = r + o;
      ++j;
    });
  }
  for (; j < w; ++j) {
    const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
    sval[at] = 0.0;
    scol[at] = pad_col;
  }
}

// the same remainder rows as CSR (zeros dropped): nonzeros per row, then (after an exclusive scan) the entries
__global__ void sell_sub_count(SellView A, int n_sub, const int *__restrict__ rows, int *__restrict__ cnt) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_sub) return;
  int w = 0;
  sell_row_foreach_nonzero(A, rows[k], [&](int, double) { ++w; });
  cnt[k] = w;
}
__global__ void sell_sub_fill_csr(SellView A, int n_sub, const int *__restrict__ rows, const int *__restrict__ ptr,
                                  int *__restrict__ col, double *__restrict__ val) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_sub) return;
  const int r = rows[k];
  int e = ptr[k];
  sell_row_foreach_nonzero(A, r, [&](int o, double v) {
    col[e] = r + o;
    val[e] = v;
    ++e;
  });
}

// The pattern table in shared memory.
struct PatTable {
  const int *ptr;
  const int *off;
  const double *val;
};

struct PatSmem {
  double val[PAT_MAX_ENT];
  int off[PAT_MAX_ENT];
  int ptr[PAT_MAX_PAT + 2];
};

__device__ __forceinline__ PatTable pat_stage(const PatView &A, PatSmem &sm) {
  for (int i = threadIdx.x; i <= A.n_pat; i += blockDim.x) sm.ptr[i] = A.ptr[i];
  for (int i = threadIdx.x; i < A.n_ent; i += blockDim.x) {
    sm.off[i] = A.off[i];
    sm.val[i] = A.val[i];
  }
  __syncthreads();
  return PatTable{sm.ptr, sm.off, sm.val};
}

// sum_j a_rj x_j for the row whose pattern id is pid; xr = x + row.  Same entry order and the same FMA chain as
// sell_row_dot.  When the whole warp shares one pattern (the interior of a uniform level) the table reads are
// shared-memory broadcasts with uniform loop bounds; otherwise every lane walks its own pattern (lanes with the
// same pattern still broadcast).  The x loads of neighbouring rows are coalesced either way.
template <bool NC>
__device__ __forceinline__ double pat_row_dot(const PatTable &T, uint32_t pid, const double *__restrict__ xr) {
  const uint32_t pid0 = __shfl_sync(0xffffffffu, pid, 0);
  double acc = 0.0;
  if (__all_sync(0xffffffffu, pid == pid0)) {
    int k = T.ptr[pid0];
    const int k1 = T.ptr[pid0 + 1];
    for (; k + 4 <= k1; k += 4) {
      double xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = ldx<NC>(xr, T.off[k + u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fma(T.val[k + u], xv[u], acc);
    }
    for (; k < k1; ++k) acc = fma(T.val[k], ldx<NC>(xr, T.off[k]), acc);
  } else {
    const int k = T.ptr[pid];
    const int len = T.ptr[pid + 1] - k;
    const int lmax = __reduce_max_sync(0xffffffffu, len);
    for (int i = 0; i < lmax; i += 4) {
      double xv[4], av[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool on = i + u < len;
        xv[u] = on ? ldx<NC>(xr, T.off[k + i + u]) : 0.0;
        av[u] = on ? T.val[k + i + u] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u < len) acc = fma(av[u], xv[u], acc);
    }
  }
  return acc;
}

// SpMV with the fused epilogues of sell_spmv (kernels.cuh) on a row-pattern matrix.  Blocks take contiguous chunks
// of slices (the table is staged once per block), then their share of the remainder slices.
template <int EPI>
__device__ __forceinline__ double pat_epilogue(int r, double ax, const double *__restrict__ x, double *__restrict__ y,
                                               const double *__restrict__ b, const double *__restrict__ dinv, double omega) {
  double yv;
  if (EPI == EPI_ASSIGN) yv = ax;
  else if (EPI == EPI_ADD) yv = y[r] + ax;
  else if (EPI == EPI_SUB) yv = y[r] - ax;
  else if (EPI == EPI_RESID) yv = b[r] - ax;
  else if (EPI == EPI_NRESID) yv = ax - b[r];
  else yv = x[r] + omega * dinv[r] * (b[r] - ax);
  y[r] = yv;
  return yv;
}

template <int EPI, int DOT>
__global__ void __launch_bounds__(512) pat_spmv(PatView A, const double *__restrict__ x, double *__restrict__ y,
                                                const double *__restrict__ b, const double *__restrict__ dinv, double omega,
                                                double *partials, unsigned int *counter, double *out) {
  __shared__ double red[32];
  __shared__ PatSmem sm;
  const PatTable T = pat_stage(A, sm);
  const int nb = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int WPB = 512 / 32;
  const uint32_t empty = (uint32_t)(A.n_pat - 1);
  double contrib = 0.0;
  {
    const int s_begin = (int)(((int64_t)A.n_slices * blockIdx.x) / nb);
    const int s_end = (int)(((int64_t)A.n_slices * (blockIdx.x + 1)) / nb);
    for (int s = s_begin + warp; s < s_end; s += WPB) {
      const int r = s * 32 + lane;
      const uint32_t pid = __ldg(A.pat + r) & PAT_ID_MASK;
      const double ax = pat_row_dot<true>(T, pid, x + r);
      if (pid != empty) {
        const double yv = pat_epilogue<EPI>(r, ax, x, y, b, dinv, omega);
        if (DOT == DOT_XY) contrib += x[r] * yv;
        if (DOT == DOT_YY) contrib += yv * yv;
      }
    }
  }
  {
    const int s_begin = (int)(((int64_t)A.rem.n_slices * blockIdx.x) / nb);
    const int s_end = (int)(((int64_t)A.rem.n_slices * (blockIdx.x + 1)) / nb);
    for (int s = s_begin + warp; s < s_end; s += WPB) {
      const double ax = sell_row_dot<true>(A.rem, s, lane, x);
      const int k = s * 32 + lane;
      if (k < A.rem.n_rows) {
        const int r = A.rem_rows[k];
        const double yv = pat_epilogue<EPI>(r, ax, x, y, b, dinv, omega);
        if (DOT == DOT_XY) contrib += x[r] * yv;
        if (DOT == DOT_YY) contrib += yv * yv;
      }
    }
  }
  if (DOT != DOT_NONE) {
    const double s = block_sum(contrib, red);
    grid_sum_finalize(s, partials, counter, out, red);
  }
}

// Row kernel of the persistent CG kernels (kernels.cuh: cg_persistent, dist.cuh: cg_persistent_dist) for the
// row-pattern format with L1 gathers: the pattern table in shared memory, the rare rows through the remainder SELL
// matrix (each block takes a contiguous share of its slices).
template <>
struct RowSmem<PatView> {
  PatSmem p;
};
template <>
struct RowDot<PatView> {
  PatTable T;
  uint32_t empty;
  bool ok;
  __device__ __forceinline__ void init(const PatView &A, RowSmem<PatView> &sm) {
    T = pat_stage(A, sm.p);
    empty = (uint32_t)(A.n_pat - 1);
    ok = false;
  }
  __device__ __forceinline__ void prefetch(const PatView &, int, int) const {}
  __device__ __forceinline__ double operator()(const PatView &A, int s, int lane, const double *x) {
    const int r = s * 32 + lane;
    const uint32_t pid = __ldg(A.pat + r) & PAT_ID_MASK;
    ok = pid != empty;
    return pat_row_dot<false>(T, pid, x + r);
  }
  __device__ __forceinline__ bool valid() const { return ok; }
  template <class F>
  __device__ __forceinline__ void remainder(const PatView &A, int block, int nb, int warp, int wpb, const double *x, F &&f) const {
    const int lane = threadIdx.x & 31;
    const int q_begin = (int)(((int64_t)A.rem.n_slices * block) / nb);
    const int q_end = (int)(((int64_t)A.rem.n_slices * (block + 1)) / nb);
    for (int q = q_begin + warp; q < q_end; q += wpb) {
      const double ad = sell_row_dot<false>(A.rem, q, lane, x);
      const int k = q * 32 + lane;
      if (k < A.rem.n_rows) f(A.rem_rows[k], ad);
    }
  }
};

}  // namespace gmg
