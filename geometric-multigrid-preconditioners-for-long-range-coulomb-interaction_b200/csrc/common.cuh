// Shared device helpers for the sm_100a kernels: warp/block reductions, deterministic grid-wide
// reductions ("last block finalises, fixed order"), streaming loads.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>

namespace gmg {

constexpr int WARP = 32;
constexpr int SLICE = 32;  // sliced-ELL slice height = one warp

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp: wait until the first n (<= 32 * U) tagged 16-byte slots, `stride_u64` 64-bit words apart, carry `tag` in both
// words; their payloads summed lane-strided first, then by the shuffle tree: the same bits for whoever adds them.
// SYS: the slots are written by peer GPUs (system-scope loads).  All pending slots of a lane are read before any of
// them is looked at (a load whose issue depends on the previous slot's tag would serialise the L2 round trips); the
// loaded words themselves are the storage of the arrived values.  false: timed out (`limit` clocks).
template <int U, bool SYS>
__device__ __forceinline__ bool ll_collect_slots(const uint64_t *slots, size_t stride_u64, int n, uint32_t tag, long long limit,
                                                 double &sum) {
  const int lane = threadIdx.x & 31;
  uint64_t w0[U], w1[U];
  uint32_t pending = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    w0[u] = w1[u] = 0;
    if (lane + 32 * u < n) pending |= 1u << u;
  }
  const long long t0 = clock64();
  bool ok = true;
  while (__any_sync(0xffffffffu, pending != 0u)) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (pending & (1u << u)) {
        if (SYS)
          asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[u]), "=l"(w1[u]) : "l"(slots + stride_u64 * (lane + 32 * u)) : "memory");
        else
          asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[u]), "=l"(w1[u]) : "l"(slots + stride_u64 * (lane + 32 * u)) : "memory");
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if ((uint32_t)(w0[u] >> 32) == tag && (uint32_t)(w1[u] >> 32) == tag) pending &= ~(1u << u);
    if (clock64() - t0 > limit) {
      ok = false;
      break;
    }
  }
  ok = __all_sync(0xffffffffu, ok);
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (lane + 32 * u < n) s += __longlong_as_double((long long)((w1[u] << 32) | (w0[u] & 0xffffffffull)));
  sum = warp_sum(s);
  return ok;
}

// Sum over the block; result valid in thread 0.  `red` needs blockDim.x/32 doubles.
__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` against a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = (lane < nw) ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}
__device__ __forceinline__ double block_max(double v, double *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = (lane < nw) ? red[lane] : 0.0;
    r = warp_max(r);
  }
  return r;
}

// Sum `n` partials in a fixed order with one warp (lane-strided serial sums, then a shuffle tree):
// every block that calls this on the same data obtains the same bits.
// Loads are ld.global.cg (L2, never a stale L1 line) and independent, so they overlap: with volatile loads the
// ~10 serial L2 round trips of this loop cost ~6 us per reduction inside the persistent CG.
__device__ __forceinline__ double warp_sum_partials(const double *p, int n) {
  const int lane = threadIdx.x & 31;
  double v[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int i = lane + 32 * u;
    v[u] = (i < n) ? __ldcg(p + i) : 0.0;
  }
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < 16; ++u) s += v[u];  // same order as the serial loop
  for (int i = lane + 512; i < n; i += 32) s += __ldcg(p + i);

  return warp_sum(s);
}

// Grid-wide deterministic reduction for ordinary (non-cooperative) launches: every block stores its
// partial, the last block to arrive sums all partials in fixed order and writes *out.
// `counter` must be 0 before the launch and is left at 0.
__device__ __forceinline__ void grid_sum_finalize(double block_value /*thread 0*/, double *partials,
                                                  unsigned int *counter, double *out, double *red) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_value;
    __threadfence();
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += ((volatile double *)partials)[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) *out = s;
  }
}

// streaming (read-once) loads: keep the matrix stream out of L1 so the gathered vector stays there
__device__ __forceinline__ double2 ld_stream_d2(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ld_stream_i2(const int2 *p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

struct SellView {
  int n_rows, n_cols, n_slices;
  const int64_t *slice_ptr;  // n_slices + 1 element offsets (multiples of 64)
  const double *val;
  const int *col;
};

// Compressed sliced ELL ("CSELL"): lossless 4-byte entries = 16-bit dictionary code of the fp64 value
// + signed 16-bit column offset from the row.  FE matrices on (mostly) uniform levels have a handful of
// distinct values and a bounded bandwidth, so 12 bytes per entry shrink to 4.  Same slice layout as SELL
// but 4 entries per 128-bit lane load: lane L reads chunk p at ent4[slice_ptr/4 + 32 p + L].
struct CsellView {
  int n_rows, n_cols, n_slices;
  const int64_t *slice_ptr;   // n_slices + 1 entry offsets (multiples of 128)
  const uint32_t *ent;        // (code << 16) | (uint16)(col - row)
  const double *dict;
  int dict_n;
};

// Row-pattern dictionary format (pattern.cuh): one pattern id per row + the table of distinct rows.
constexpr uint32_t PAT_ID_MASK = 0x3fffffffu;   // pattern id bits of PatView::pat[r]
constexpr uint32_t PAT_GENERAL = 0x40000000u;   // window kernel: row must take the general path (a dominant column is out of range)
constexpr uint32_t PAT_ZEROED = 0x80000000u;    // window kernel: entry r of the zeroed operand copy is 0 (see pattern_win.cuh)
struct PatView {
  int n_rows, n_cols, n_slices;
  const uint32_t *pat;   // pattern id per row (padded to a multiple of 32 rows with the empty pattern)
  const int *ptr;        // n_pat + 1 offsets into off / val
  const int *off;        // column - row
  const double *val;
  int n_pat, n_ent;       // n_pat includes the empty pattern (id n_pat - 1): row handled by the remainder / padding
  SellView rem;           // rows whose pattern is not in the table, as a SELL matrix (zeros dropped) ...
  const int *rem_rows;    // ... and their row indices, ascending
  const int *rem_ptr;     // the same rows as CSR (rem.n_rows + 1 offsets; window kernel: 4 lanes per row)
  const int *rem_ccol;
  const double *rem_cval;
  // the same rows once more for the 4-lanes-per-row walk of the window kernel, in the order the lanes load them: groups
  // of 8 rows, per group [8 entries][32 lanes] (lane = 4 * row-in-group + quarter; lane's i-th entry, column -1 = none):
  // one coalesced line per load instruction instead of 32 sectors.  rem4_long[group] != 0: a row of the group has more
  // than 32 entries (walked through the CSR copy)
  const int *rem4_col;
  const double *rem4_val;
  const unsigned char *rem4_long;
};

// Dominant pattern + window plan of the TMA-staged persistent CG (pattern_win.cuh)
constexpr int WIN_BLOCK = 1024;      // threads of the window kernel: 31 consumer warps + 1 producer warp
constexpr int WIN_SPW = 2;           // slices per consumer warp and tile
constexpr int WIN_TILE_SLICES = (WIN_BLOCK / 32 - 1) * WIN_SPW;
constexpr int WIN_TILE_ROWS = WIN_TILE_SLICES * 32;
constexpr int DOM_MAX = 32;
constexpr int WIN_MAX_SEG = 4;

struct DomPat {
  int len;                    // entries of the dominant pattern (0: none)
  int nseg;                   // window segments per tile
  int win_elems;              // doubles per window stage (even)
  int diag_wbyte;             // window byte offset of column == row
  int wbyte[DOM_MAX];         // window byte offset of entry k for the first row of a tile (+ 8 * local row); 0 past len
  double val[DOM_MAX];
  int seg_lo[WIN_MAX_SEG];    // first column of the segment relative to the tile's first row (even)
  int seg_len[WIN_MAX_SEG];   // doubles (even)
  int seg_base[WIN_MAX_SEG];  // position in the window (even)
};

struct PcgScalars {
  double gh[2];   // g.h, double-buffered (beta = gh[new] / gh[old])
  double dh;      // d.(A d)
  double res2;    // g.g
  double tmp;
};

struct CgResult {
  int iterations;
  int status;  // 0 converged, 1 max iterations reached
  double res0, res;
};

}  // namespace gmg
