// Persistent coarse-grid CG on a row-pattern matrix with the SpMV operand staged through shared-memory WINDOWS
// filled by TMA bulk copies (cp.async.bulk + mbarrier), one 1024-thread block per SM.
//
// Why: with the matrix reduced to 4 bytes per row (pattern.cuh) the SpMV is bound by L1 wavefronts: a 27-point row
// needs 27 gathers per lane, each a misaligned 256-byte warp load = 3 lines = ~6 LSU cycles.  The rows of a tile of
// WIN_TILE_ROWS (1984) consecutive rows with the DOMINANT pattern (offsets o_0..o_L) read the ranges
// [r0 + o_k, r0 + o_k + 1984); their union is a few contiguous segments (3 for a first-touch-numbered Q1 lattice: one
// per lattice plane).  Warp 31 is the producer: one lane issues one bulk copy per segment into a two-stage
// shared-memory window (no LSU work, no registers; full / empty mbarriers, no block barrier inside the SpMV);
// warps 0..30 consume: the 27 gathers become conflict-free LDS.64 [R + UR + imm] at constant window offsets.  The
// dominant pattern itself (values, window byte offsets) is a kernel parameter: fully unrolled, its values and
// offsets are constant-bank operands of the DFMA / address arithmetic, no table loads at all, and each constant is
// loaded once for the two rows a lane owns in a tile.  Which path a row takes (dominant loop / single diagonal entry /
// general table walk / remainder) is a 16-bit row code, staged in shared memory for the block's rows (or read from
// global memory when a block owns more than ~35 k rows).
//
// Rows whose pattern is a sub-sequence of the dominant one with the same values (rows next to an eliminated
// Dirichlet boundary: the couplings to boundary columns are stored zeros) run the SAME unmasked loop: the operand
// vector d holds exact zeros at the skipped columns, and fma(a, 0, acc) == acc, so the chain is bit-identical to the
// one that skips those entries (see pattern.cuh).  The set Z of zeroed columns is derived at build time
// (pat_mark_columns); it is only used when no row other than j itself needs the value of a column j of Z (true for
// Dirichlet-eliminated matrices; otherwise only the exact dominant rows use the windows).  The true values of the
// Z entries live in a side vector dt that only the Z rows themselves touch (their own diagonal term, the x update,
// the direction update): no second operand copy, no extra L2 footprint.
// The entry order of every row is unchanged: results are bit-identical to the other formats.
#pragma once
#include "pattern.cuh"

namespace gmg {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin (a lost transaction must not hang the GPU): returns false on time-out
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy global writes (other SMs' d updates, ordered by the grid barrier) -> async-proxy reads (TMA)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// one tile's window: every segment clamped to [0, n_even) (the clamped parts are never used by a stored entry)
__device__ __forceinline__ void win_issue(const DomPat &D, const double *src, int n_even, int r0, double *stage, uint64_t *bar) {
  uint32_t bytes = 0;
#pragma unroll
  for (int s = 0; s < WIN_MAX_SEG; ++s)
    if (s < D.nseg) {
      const int g0 = r0 + D.seg_lo[s], g1 = g0 + D.seg_len[s];
      const int c0 = max(g0, 0), c1 = min(g1, n_even);
      if (c1 > c0) bytes += (uint32_t)(c1 - c0) * 8u;
    }
  mbar_expect_tx(bar, bytes);
#pragma unroll
  for (int s = 0; s < WIN_MAX_SEG; ++s)
    if (s < D.nseg) {
      const int g0 = r0 + D.seg_lo[s], g1 = g0 + D.seg_len[s];
      const int c0 = max(g0, 0), c1 = min(g1, n_even);
      if (c1 > c0) tma_bulk_g2s(stage + D.seg_base[s] + (c0 - g0), src + c0, (uint32_t)(c1 - c0) * 8u, bar);
    }
}

// per-lane walk of the pattern table (lanes with the same pattern broadcast); pid == empty: length 0.
// zrow: this row is a member of Z, its own entry of the operand vector is `self` (xr[0] holds the zero).
template <bool NC>
__device__ __forceinline__ double pat_row_dot_lanes(const PatTable &T, uint32_t pid, const double *__restrict__ xr, bool zrow,
                                                    double self) {
  const int k = T.ptr[pid];
  const int len = T.ptr[pid + 1] - k;
  const int lmax = __reduce_max_sync(0xffffffffu, len);
  double acc = 0.0;
  for (int i = 0; i < lmax; i += 4) {
    double xv[4], av[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool on = i + u < len;
      const int o = on ? T.off[k + i + u] : 0;
      xv[u] = on ? ldx<NC>(xr, o) : 0.0;
      if (zrow && o == 0) xv[u] = self;
      av[u] = on ? T.val[k + i + u] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i + u < len) acc = fma(av[u], xv[u], acc);
  }
  return acc;
}

// dynamic shared memory: [4 mbarriers | 2 window stages | pattern table | 16-bit row codes of the block's rows]
struct WinLayout {
  int win_off, table_off, code_off, total;
};
__host__ __device__ inline WinLayout win_layout(int win_elems, int rows_per_block) {
  WinLayout L;
  L.win_off = 32;
  L.table_off = L.win_off + 2 * win_elems * 8;
  L.code_off = L.table_off + (int)sizeof(PatSmem);
  L.total = (L.code_off + 2 * rows_per_block + 15) & ~15;
  return L;
}
// row code: pattern id | flags
constexpr uint32_t RC_ID = 0x0fffu, RC_DIAG = 0x1000u, RC_DOM = 0x2000u, RC_Z = 0x4000u, RC_EMPTY = 0x8000u;

// Build time: which columns are needed with their true value (1) / must read as zero on the dominant path (2)?
// A dominant-path row with a dominant column outside the matrix is sent to the general path.  Table rows off the
// dominant path need all their columns except their own diagonal (a Z row reads that from dt).
__global__ void pat_mark_columns(PatView A, uint32_t *__restrict__ pat, const uint32_t *__restrict__ dom_mask, int dom_len,
                                 int *__restrict__ colflag) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n_rows) return;
  const uint32_t pid = pat[r] & PAT_ID_MASK;
  if (pid == (uint32_t)(A.n_pat - 1)) return;  // remainder row: pat_mark_remainder
  const uint32_t m = dom_mask[pid];
  bool dom = m != 0u;
  if (dom) {
    for (int k = 0; k < dom_len; ++k) {
      const int j = r + A.off[k];
      if (j < 0 || j >= A.n_cols) dom = false;
    }
    if (!dom) pat[r] = pid | PAT_GENERAL;
  }
  if (dom) {
    for (int k = 0; k < dom_len; ++k) atomicOr(colflag + r + A.off[k], ((m >> k) & 1u) ? 1 : 2);
  } else {
    for (int k = A.ptr[pid]; k < A.ptr[pid + 1]; ++k)
      if (A.off[k] != 0) atomicOr(colflag + r + A.off[k], 1);
  }
}
__global__ void pat_mark_remainder(SellView R, int *__restrict__ colflag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= R.n_rows) return;
  const int slice = k >> 5, lane = k & 31;
  const int64_t b = R.slice_ptr[slice];
  const int w = (int)((R.slice_ptr[slice + 1] - b) >> 5);
  for (int j = 0; j < w; ++j) {
    const int64_t at = b + (int64_t)(j >> 1) * 64 + lane * 2 + (j & 1);
    if (R.val[at] != 0.0) atomicOr(colflag + R.col[at], 1);
  }
}
// colflag 2 -> member of Z; 3 -> conflict (reported, nothing is zeroed then)
__global__ void pat_apply_zero_set(int n, const int *__restrict__ colflag, uint32_t *__restrict__ pat, int *conflict, int apply) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int f = colflag[r];
  if (f == 3) atomicOr(conflict, 1);
  if (apply && f == 2) pat[r] |= PAT_ZEROED;
}

// the 16-bit row codes of the window kernel (pattern id + path flags) for every row: used from global memory when a
// block's share of them does not fit next to the windows in shared memory (more than ~35 k rows per block)
__global__ void pat_row_codes(PatView A, const uint32_t *__restrict__ dom_mask, unsigned short *__restrict__ code) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n_slices * 32) return;
  const uint32_t p = A.pat[r];
  const uint32_t id = p & PAT_ID_MASK;
  uint32_t c = id;
  if (id == (uint32_t)(A.n_pat - 1)) c |= RC_EMPTY;
  else if (!(p & PAT_GENERAL) && dom_mask[id] != 0u) c |= RC_DOM;
  else if (A.ptr[id + 1] - A.ptr[id] == 1 && A.off[A.ptr[id]] == 0) c |= RC_DIAG;
  if (p & PAT_ZEROED) c |= RC_Z;
  code[r] = (unsigned short)c;
}

// optional phase timing (block 0, thread 0; globaltimer ns): [spmv, barrier 1, update, barrier 2, direction, barrier 3, iterations]
__device__ unsigned long long g_cg_phase_ns[16];
__device__ unsigned long long g_cg_block_ns[3][256];  // per block: SpMV, update, direction phase (prof != 0)
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GMG_PHASE(i)                                   \
  if (prof && blockIdx.x == prof - 1 && threadIdx.x == 0) {   \
    const unsigned long long now = gtime();            \
    g_cg_phase_ns[i] += now - t_prev;                  \
    t_prev = now;                                      \
  }

// Remainder rows (CSR, zeros dropped) with 4 lanes per row: every lane loads its quarter of the row's entries and
// operands at once (two memory round trips per row instead of two per four entries), then the FMA chain runs through
// the four lanes in entry order (the partial sum is handed on by shuffle): same order, same bits as the SELL path.
__device__ __forceinline__ double rem_row_dot4_at(const PatView &A, int p0, int len, int c /* lane & 3 */, const double *__restrict__ x);
__device__ __forceinline__ double rem_row_dot4(const PatView &A, int k /* remainder row or -1 */, int c /* lane & 3 */,
                                               const double *__restrict__ x) {
  int p0 = 0, len = 0;
  if (k >= 0) {
    p0 = __ldg(A.rem_ptr + k);
    len = __ldg(A.rem_ptr + k + 1) - p0;
  }
  return rem_row_dot4_at(A, p0, len, c, x);
}
// (the row's entries are rem_ccol / rem_cval [p0, p0 + len); len == 0: no row)
__device__ __forceinline__ double rem_row_dot4_at(const PatView &A, int p0, int len, int c, const double *__restrict__ x) {
  const int lmax = __reduce_max_sync(0xffffffffu, len);
  double acc = 0.0;
  for (int base = 0; base < lmax; base += 32) {  // 32 entries per round, 8 per lane
    const int n_here = min(max(len - base, 0), 32);
    const int chunk = (n_here + 3) >> 2;
    const int e0 = p0 + base + c * chunk;
    const int cnt = max(min(chunk, n_here - c * chunk), 0);
    double v[8], xv[8];
    int col[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      col[i] = i < cnt ? __ldg(A.rem_ccol + e0 + i) : -1;
      v[i] = i < cnt ? __ldg(A.rem_cval + e0 + i) : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) xv[i] = col[i] >= 0 ? x[col[i]] : 0.0;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      if (c == cc) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < cnt) acc = fma(v[i], xv[i], acc);
      }
      acc = __shfl_sync(0xffffffffu, acc, (threadIdx.x & 28) | cc);
    }
  }
  return acc;
}

constexpr int REM4_NONE = (int)0x80000000;  // "no entry" in rem4_col (rank-local blocks have negative halo columns)
// The lane-ordered copy of the remainder rows (PatView::rem4_*): thread (group, lane) writes the up to 8 entries lane
// `lane` of a remainder warp walks for row 8 * group + lane / 4 -- the quarter `lane & 3` of the row's first 32 entries,
// split exactly as rem_row_dot4_at splits them.
__global__ void __launch_bounds__(256) rem4_build(int n_rem, const int *__restrict__ rem_ptr, const int *__restrict__ ccol,
                                                  const double *__restrict__ cval, int *__restrict__ col4,
                                                  double *__restrict__ val4, unsigned char *__restrict__ is_long) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int group = (int)(t >> 5), lane = (int)(t & 31);
  if (group * 8 >= n_rem) return;
  const int k = group * 8 + (lane >> 2), c = lane & 3;
  int p0 = 0, len = 0;
  if (k < n_rem) {
    p0 = rem_ptr[k];
    len = rem_ptr[k + 1] - p0;
  }
  if (len > 32 && c == 0) is_long[group] = 1;
  const int n_here = min(len, 32);
  const int chunk = (n_here + 3) >> 2;
  const int e0 = p0 + c * chunk;
  const int cnt = max(min(chunk, n_here - c * chunk), 0);
  for (int i = 0; i < 8; ++i) {
    col4[(int64_t)group * 256 + i * 32 + lane] = i < cnt ? ccol[e0 + i] : REM4_NONE;
    val4[(int64_t)group * 256 + i * 32 + lane] = i < cnt ? cval[e0 + i] : 0.0;
  }
}
// the walk of rem_row_dot4_at on that copy (rows of at most 32 entries): same entries per lane, same chain
__device__ __forceinline__ double rem_row_dot4_lanes(const PatView &A, int group, const double *__restrict__ x) {
  const int lane = threadIdx.x & 31, c = lane & 3;
  const int *cp = A.rem4_col + (int64_t)group * 256 + lane;
  const double *vp = A.rem4_val + (int64_t)group * 256 + lane;
  double v[8], xv[8];
  int col[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    col[i] = __ldg(cp + i * 32);
    v[i] = __ldg(vp + i * 32);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) xv[i] = col[i] != REM4_NONE ? x[col[i]] : 0.0;
  double acc = 0.0;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    if (c == cc) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col[i] != REM4_NONE) acc = fma(v[i], xv[i], acc);
    }
    acc = __shfl_sync(0xffffffffu, acc, (threadIdx.x & 28) | cc);
  }
  return acc;
}

// 1024 threads: warps 0..30 compute, warp 31 is the TMA producer.  A tile is 31 * SPW slices; consumer warp w takes
// the slices tile + w, tile + w + 31, ... so that each table constant (window offset, value) is loaded once for SPW
// rows.  Two window stages with full / empty mbarriers; no block-wide barrier inside the SpMV.
template <int SPW>
__global__ void __launch_bounds__(WIN_BLOCK, 1)
    cg_persistent_win(PatView A, const __grid_constant__ DomPat D, const uint32_t *__restrict__ dom_mask, const double *__restrict__ b,
                      double *x, double *g, double *d, double *dt, double *h, double *partials /* 3 * gridDim.x */, int max_it,
                      double tol, CgResult *result, int rows_per_block, const unsigned short *__restrict__ gcode, int prof) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red[32];
  __shared__ double bc;
  constexpr int BLOCK = WIN_BLOCK;
  constexpr int WPB = BLOCK / 32;   // warps (phases 2, 3 and reductions use all of them)
  constexpr int CW = WPB - 1;       // consumer warps of the SpMV
  constexpr int TILE = CW * SPW;    // slices per tile
  static_assert(TILE == WIN_TILE_SLICES, "window plan and kernel disagree on the tile size");
  unsigned long long t_prev = 0;
  const WinLayout lay = win_layout(D.win_elems, rows_per_block);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);  // [2]
  uint64_t *empty_bar = full + 2;                       // [2]
  double *win = reinterpret_cast<double *>(smem + lay.win_off);
  PatSmem &sm = *reinterpret_cast<PatSmem *>(smem + lay.table_off);
  unsigned short *scode = reinterpret_cast<unsigned short *>(smem + lay.code_off);

  const int nb = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s_begin = (int)(((int64_t)A.n_slices * blockIdx.x) / nb);
  const int s_end = (int)(((int64_t)A.n_slices * (blockIdx.x + 1)) / nb);
  const int k_begin = (int)(((int64_t)A.rem.n_rows * blockIdx.x) / nb);
  const int k_end = (int)(((int64_t)A.rem.n_rows * (blockIdx.x + 1)) / nb);
  const int n_tiles = (s_end - s_begin + TILE - 1) / TILE;
  const int n_even = (A.n_rows + 1) & ~1;
  const uint32_t empty_id = (uint32_t)(A.n_pat - 1);

  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(full + 1, 1);
    mbar_init(empty_bar, CW);
    mbar_init(empty_bar + 1, CW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * D.win_elems; i += BLOCK) win[i] = 0.0;  // (lanes off the dominant path read, but never use, window data)
  const PatTable T = pat_stage(A, sm);  // (ends with __syncthreads)
  // row codes of this block's rows (they never change): pattern id + which path the row takes.  RC_DIAG: the row is
  // its diagonal entry only (eliminated Dirichlet rows): a . d[r], no table walk.  In shared memory when they fit
  // (rows_per_block > 0), else read from the precomputed global array.
  const unsigned short *code = gcode + (size_t)s_begin * 32;
  if (rows_per_block > 0) {
    for (int i = threadIdx.x; i < (s_end - s_begin) * 32; i += BLOCK) scode[i] = gcode[(size_t)s_begin * 32 + i];
    code = scode;
  }
  __syncthreads();
  double *pa = partials, *pb = partials + nb, *pc = partials + 2 * nb;
  uint32_t tc = 0;  // tiles produced / consumed so far: stage = tc & 1, use of that stage = tc >> 1
  bool pipeline_ok = true;

  // x = 0 ; g = -b ; d = b (0 on Z, true value in dt) ; res0 = |b|
  double acc = 0.0;
  for (int s = s_begin + warp; s < s_end; s += WPB) {
    const int r = s * 32 + lane;
    if (r < A.n_rows) {
      const double bv = b[r];
      const bool z = code[(s - s_begin) * 32 + lane] & RC_Z;
      x[r] = 0.0;
      g[r] = -bv;
      d[r] = z ? 0.0 : bv;
      if (z) dt[r] = bv;
      acc += bv * bv;
    }
  }
  fence_proxy_async();
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) pc[blockIdx.x] = acc;
  grid.sync();
  double res2 = grid_total(pc, nb, &bc);
  double res = sqrt(res2);
  const double res0 = res;
  int it = 0, status = 0;
  double alpha = 0.0;
  if (res > tol) {
    double gh = res * res;
    while (true) {
      ++it;
      if (prof && blockIdx.x == prof - 1 && threadIdx.x == 0) {
        t_prev = gtime();
        g_cg_phase_ns[6] += 1;
      }
      // ---- h = A d ; dh = d.h ------------------------------------------------------------------
      acc = 0.0;
      unsigned long long tb = 0;
      if (prof && threadIdx.x == 32) tb = gtime();
      if (warp == CW) {
        // producer: fill stage (tc & 1) with tile t as soon as its previous user has released it
        if (lane == 0) {
          fence_proxy_async();
          for (int t = 0; t < n_tiles; ++t, ++tc) {
            const uint32_t st = tc & 1u, use = tc >> 1;
            if (!mbar_wait(empty_bar + st, (use + 1u) & 1u)) pipeline_ok = false;
            win_issue(D, d, n_even, (s_begin + t * TILE) * 32, win + st * D.win_elems, full + st);
          }
        }
        tc = __shfl_sync(0xffffffffu, tc, 0);
      } else {
        for (int t = 0; t < n_tiles; ++t, ++tc) {
          const int tile0 = s_begin + t * TILE;
          const uint32_t st = tc & 1u, use = tc >> 1;
          uint32_t rc[SPW];
          bool any_dom = false, any_general = false;
#pragma unroll
          for (int j = 0; j < SPW; ++j) {
            const int s = tile0 + j * CW + warp;
            rc[j] = (s < s_end) ? code[(s - s_begin) * 32 + lane] : RC_EMPTY;
            any_dom = any_dom || (rc[j] & RC_DOM);
            any_general = any_general || !(rc[j] & (RC_DOM | RC_EMPTY | RC_DIAG));
          }
          any_dom = __any_sync(0xffffffffu, any_dom);
          any_general = __any_sync(0xffffffffu, any_general);
          GMG_PHASE(8)
          if (!mbar_wait(full + st, use & 1u)) pipeline_ok = false;
          GMG_PHASE(9)
          // byte address of this lane's first row in the window; row j of the warp is j * CW slices further
          const char *w = reinterpret_cast<const char *>(win + st * D.win_elems + warp * 32 + lane);
          double ad[SPW];
#pragma unroll
          for (int j = 0; j < SPW; ++j) ad[j] = 0.0;
          if (any_dom) {
#pragma unroll
            for (int k0 = 0; k0 < DOM_MAX; k0 += 4) {
              if (k0 + 4 <= D.len) {
                double xv[4][SPW];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                  for (int j = 0; j < SPW; ++j)
                    xv[u][j] = *reinterpret_cast<const double *>(w + D.wbyte[k0 + u] + j * (CW * 32 * 8));
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                  for (int j = 0; j < SPW; ++j) ad[j] = fma(D.val[k0 + u], xv[u][j], ad[j]);
              } else if (k0 < D.len) {
#pragma unroll
                for (int u = 0; u < 3; ++u)
                  if (k0 + u < D.len) {
#pragma unroll
                    for (int j = 0; j < SPW; ++j)
                      ad[j] = fma(D.val[k0 + u], *reinterpret_cast<const double *>(w + D.wbyte[k0 + u] + j * (CW * 32 * 8)), ad[j]);
                  }
              }
            }
          }
          double dr[SPW];
#pragma unroll
          for (int j = 0; j < SPW; ++j) dr[j] = *reinterpret_cast<const double *>(w + D.diag_wbyte + j * (CW * 32 * 8));
          // this warp is done with the stage: hand it back to the producer
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty_bar + st)) : "memory");
          GMG_PHASE(10)
#pragma unroll
          for (int j = 0; j < SPW; ++j) {
            const int s = tile0 + j * CW + warp;
            const int r = s * 32 + lane;
            const bool general = !(rc[j] & (RC_DOM | RC_EMPTY | RC_DIAG));
            const bool z = !(rc[j] & RC_DOM) && (rc[j] & RC_Z);
            if (z) dr[j] = dt[r];
            if (rc[j] & RC_DIAG) ad[j] = fma(T.val[T.ptr[rc[j] & RC_ID]], dr[j], 0.0);
            if (any_general) {
              const double ag = pat_row_dot_lanes<false>(T, general ? (rc[j] & RC_ID) : empty_id, d + r, z, dr[j]);
              if (general) ad[j] = ag;
            }
            if (!(rc[j] & RC_EMPTY)) {
              h[r] = ad[j];
              acc += dr[j] * ad[j];
            }
          }
          GMG_PHASE(11)
        }
        // rows whose pattern is not in the table: 8 per warp and round (after the tiles: the window fills no longer load the L2)
        for (int k0 = k_begin + warp * 8; k0 < k_end; k0 += CW * 8) {
          const int k = k0 + (lane >> 2);
          const bool on = k < k_end;
          const double aq = rem_row_dot4(A, on ? k : -1, lane & 3, d);
          if (on && (lane & 3) == 0) {
            const int r = A.rem_rows[k];
            h[r] = aq;
            acc += d[r] * aq;
          }
        }
        GMG_PHASE(13)
      }
      acc = block_sum(acc, red);
      if (threadIdx.x == 0) pa[blockIdx.x] = acc;
      if (prof && threadIdx.x == 32) g_cg_block_ns[0][blockIdx.x & 255] += gtime() - tb;
      GMG_PHASE(0)
      grid.sync();
      alpha = gh / grid_total(pa, nb, &bc);
      GMG_PHASE(1)
      // ---- g += alpha h ; res2 = g.g  (x += alpha d is done together with the direction update) ----
      acc = 0.0;
      if (prof && threadIdx.x == 32) tb = gtime();
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        const int r = s * 32 + lane;
        if (r < A.n_rows) {
          const double gv = g[r] + alpha * h[r];
          g[r] = gv;
          acc += gv * gv;
        }
      }
      acc = block_sum(acc, red);
      if (threadIdx.x == 0) pb[blockIdx.x] = acc;
      if (prof && threadIdx.x == 32) g_cg_block_ns[1][blockIdx.x & 255] += gtime() - tb;
      GMG_PHASE(2)
      grid.sync();
      res2 = grid_total(pb, nb, &bc);
      GMG_PHASE(3)
      res = sqrt(res2);
      if (res <= tol) break;
      if (it >= max_it) { status = 1; break; }
      const double beta = res2 / gh;
      gh = res2;
      // ---- x += alpha d ; d = beta d - g -------------------------------------------------------
      if (prof && threadIdx.x == 32) tb = gtime();
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        const int r = s * 32 + lane;
        if (r < A.n_rows) {
          const bool z = code[(s - s_begin) * 32 + lane] & RC_Z;
          const double dv = z ? dt[r] : d[r];
          x[r] += alpha * dv;
          const double dn = beta * dv - g[r];
          if (z) dt[r] = dn;  // (d[r] stays 0)
          else d[r] = dn;
        }
      }
      fence_proxy_async();
      __syncthreads();
      if (prof && threadIdx.x == 32) g_cg_block_ns[2][blockIdx.x & 255] += gtime() - tb;
      GMG_PHASE(4)
      grid.sync();
      GMG_PHASE(5)
    }
    // the x update of the last iteration
    for (int s = s_begin + warp; s < s_end; s += WPB) {
      const int r = s * 32 + lane;
      if (r < A.n_rows) {
        const bool z = code[(s - s_begin) * 32 + lane] & RC_Z;
        x[r] += alpha * (z ? dt[r] : d[r]);
      }
    }
  }
  // a timed-out window transaction (never observed) is reported as status 2 instead of a wrong answer
  const int bad = __syncthreads_or(pipeline_ok ? 0 : 1);
  if (threadIdx.x == 0) pc[blockIdx.x] = bad ? 1.0 : 0.0;
  grid.sync();
  const double n_bad = grid_total(pc, nb, &bc);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    result->iterations = it;
    result->status = n_bad != 0.0 ? 2 : status;
    result->res0 = res0;
    result->res = res;
  }
}

}  // namespace gmg
