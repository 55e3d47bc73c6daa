// Second generation of the TMA-window coarse-grid CG (pattern_win.cuh explains the windows, the dominant pattern and
// the zeroed-operand set Z; all of that is unchanged).  What the per-phase timers of the first kernel showed at 64k
// atoms (48 us per inner iteration: SpMV 23, g update 3.8, x/d update 6.8, barriers 14) is what changed here:
//
//  * NO cooperative-groups grid.sync(): the three grid-wide synchronisations of an iteration are tagged 16-byte words
//    ("LL" words, as the multi-GPU kernel uses across NVLink): every block publishes {partial sum, tag} with one store,
//    one warp per block polls the slots of all blocks and adds them in a fixed order.  A reduction + barrier is one L2
//    round trip after the last block arrives instead of barrier + atomic + spin + barrier + a second pass over the
//    partials; the barrier before the x/d update needs no fence at all (it orders nothing but the scalar).
//  * Rows off the window path (remainder rows: rare patterns) are walked by DEDICATED warps while the tile warps
//    run the dominant loop: their dependent L2 round trips (row pointer -> entries -> operands) no longer sit on the
//    critical path after the tiles, and the tile loop has no divergent general-row epilogue any more (build_pat moves
//    every row that is neither dominant-compatible nor a single diagonal entry to the remainder for this kernel).
//  * h = A d of a block's own window-path rows never leaves the SM: it is written to shared memory in the SpMV and
//    read back in the g update (28 MB less L2 traffic per iteration at 64k atoms); only the remainder rows, which
//    are computed by whichever block got them, go through global memory.
//
// The arithmetic is unchanged (same FMA chains, same summation order of the partials): iteration counts and residuals
// are bit-identical to the first kernel and to the other formats.
#pragma once
#include "pattern_win.cuh"

namespace gmg {

constexpr int WIN2_TILE_SLICES = 48;                     // slices per tile: 24 tile warps x 2 slices (1024 threads) or 12 x 4 (512)
constexpr int WIN2_TILE_ROWS = WIN2_TILE_SLICES * 32;
constexpr int WIN2_MAX_BLOCKS = 256;                     // slots one lane polls: 8
constexpr int WIN2_CHANNELS = 4;
constexpr int WIN2_SLOT_U64_MAX = 128;                   // largest distance of two blocks' slots (in 64-bit words: 1 KB)
// XPAIR (below) is compiled for the dominant pattern of a Q1 lattice with an odd number of vertices per line: 9 runs of
// 3 consecutive columns, runs 0, 2, 4, 6, 8 start at an odd window position, the diagonal is the middle entry of run 4
constexpr int XP_RUNS = 9, XP_DIAG = 4;
constexpr uint32_t XP_ODD = 0x155u;

__device__ __forceinline__ void llg_store(uint64_t *p /*16-byte aligned pair*/, double v, uint32_t tag) {
  const uint64_t bits = (uint64_t)__double_as_longlong(v);
  const uint64_t w0 = ((uint64_t)tag << 32) | (bits & 0xffffffffull);
  const uint64_t w1 = ((uint64_t)tag << 32) | (bits >> 32);
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}

// Sum over all blocks + barrier, every thread contributes `v`.  FENCE: the barrier also orders the blocks' global
// writes before it against the reads after it (release fence before the publish, acquire fence after the poll).
// Two block barriers per call: the staging arrays alternate with the parity of the tag (`red`: [2][32], `bc`: [2]).
// The two halves of ll_grid_sum, for callers that have work to do while the other blocks arrive.
template <bool FENCE>
__device__ __forceinline__ void ll_publish(uint64_t *slots, int slot_u64, double v, uint32_t tag, double *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int par = tag & 1u;
  v = warp_sum(v);
  if (lane == 0) red[par * 32 + w] = v;
  __syncthreads();
  // (warp 0 has passed a block barrier after every thread's writes: its release fence covers them)
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double r = (lane < nw) ? red[par * 32 + lane] : 0.0;
    r = warp_sum(r);
    if (lane == 0) {
      if (FENCE) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      llg_store(slots + (size_t)slot_u64 * blockIdx.x, r, tag);
    }
  }
}
template <bool FENCE>
__device__ __forceinline__ double ll_await(const uint64_t *slots, int slot_u64, int nb, uint32_t tag, double *bc, int *bad) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int par = tag & 1u;
  if (w == 0) {
    double s;
    const bool ok = ll_collect_slots<WIN2_MAX_BLOCKS / 32, false>(slots, (size_t)slot_u64, nb, tag, 4000000000LL, s);
    if (lane == 0) {
      // (the block barrier below hands the acquire fence on to every thread, as cooperative groups' grid.sync() does)
      if (FENCE) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      bc[par] = s;
      if (!ok) *bad = 1;
    }
  }
  __syncthreads();
  return bc[par];
}
template <bool FENCE>
__device__ __forceinline__ double ll_grid_sum(uint64_t *slots, int slot_u64, int nb, double v, uint32_t tag, double *red, double *bc,
                                              int *bad) {
  ll_publish<FENCE>(slots, slot_u64, v, tag, red);
  return ll_await<FENCE>(slots, slot_u64, nb, tag, bc, bad);
}

// dynamic shared memory: [4 mbarriers | 2 window stages | diagonal value per pattern id | row index, first entry and
// length of the remainder rows the block walks | h of the block's rows | row codes]
struct Win2Layout {
  int win_off, diag_off, rem_off, h_off, code_off, total;
  int h_rows, code_rows;  // 0: h / the codes stay in global memory
  int rem_cap;
};
__host__ __device__ inline Win2Layout win2_layout(int win_elems, int n_pat, int rows_per_block, bool h_smem, bool code_smem,
                                                  int rem_per_block) {
  Win2Layout L;
  L.win_off = 32;
  L.diag_off = L.win_off + 2 * win_elems * 8;
  L.rem_off = L.diag_off + ((n_pat + 1) & ~1) * 8;
  L.rem_cap = (rem_per_block + 3) & ~3;
  L.h_off = L.rem_off + L.rem_cap * 12;
  L.h_rows = h_smem ? rows_per_block : 0;
  L.code_off = L.h_off + L.h_rows * 8;
  L.code_rows = code_smem ? rows_per_block : 0;
  L.total = (L.code_off + 2 * L.code_rows + 15) & ~15;
  return L;
}

// BLOCK threads: TW tile warps (SPW slices each per tile, TW * SPW == WIN2_TILE_SLICES), the last warp produces, the
// warps in between walk the remainder rows.  GV > 0: g of the first GV slices a thread owns in the vector phases lives
// in REGISTERS for the whole solve (g is only ever touched by the thread that owns the row): the g update then reads
// shared memory and registers only, the direction update moves 32 instead of 40 bytes per row through L2 (these
// phases run at the L2 slice throughput of the chip, ~43 B / clock / SM, so bytes are what counts).  With 512
// threads a thread has 128 registers: 24 slices = 48 registers of g for blocks of up to 12288 rows.
//
// XPAIR: a lane of a tile warp owns TWO CONSECUTIVE rows (a warp walks chunks of 64 consecutive rows) instead of one row
// of each of SPW slices.  The dominant pattern of a first-touch-numbered Q1 lattice is 9 runs of 3 consecutive columns;
// the two rows of a lane need 4 consecutive operands per run instead of 2 x 3, so the dominant loop reads 8 instead of
// 12 shared-memory wavefronts per run and 64 rows:
//   run starts at an even column: 2 x LDS.128 (lane stride 16 bytes: conflict-free);
//   odd column: LDS.128 for the two middle operands + 2 x LDS.64 for the outer ones, where lanes 0-7 of every group of
//   16 read the left operand while lanes 8-15 read the right one (and the other way round in the second load): the 16
//   lanes of a phase then touch 16 different 8-byte slots of a 128-byte line (with every lane reading the same
//   operand, lanes L and L + 8 would collide), and a select puts the two values in place.
// Fewer wavefronts alone bought nothing (B200, 64k atoms: 12.5 -> 11.5 us for the loop): with 8 tile warps the loop is
// bound by the latency of "loads, wait, dependent DFMAs" per run.  What made it 6.8 us: the run parities are
// compile-time constants (XP_ODD: the loop is one basic block) and run r + 1 of a chunk is loaded right after run r of
// that chunk has been consumed, into the same registers, while the other chunks' DFMAs issue.
// Every row still sums its entries in entry order from 0.0 (bit-identical rows); the partial sums of d.h are
// grouped differently (as they are between the other kernels).  xpair_supported() (host) checks the run structure.
template <int BLOCK, int TW, int SPW, int GV, int DVB, bool XSPLIT, bool XPAIR = false>
__global__ void __launch_bounds__(BLOCK, 1)
    cg_persistent_win2(PatView A, const __grid_constant__ DomPat D, const uint32_t *__restrict__ dom_mask, const double *__restrict__ b,
                       double *x, double *g, double *d, double *dt, double *h,
                       uint64_t *ll /* WIN2_CHANNELS channels of WIN2_MAX_BLOCKS slots, slot_u64 words apart */, int slot_u64,
                       uint32_t tag_base, int max_it, double tol, CgResult *result, int rows_per_block, int h_smem, int code_smem,
                       const unsigned short *__restrict__ gcode, int prof) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red[64];
  __shared__ double bc[2];
  __shared__ int bad;
  constexpr int WPB = BLOCK / 32;          // warps (the vector phases and reductions use all of them)
  constexpr int PW = WPB - 1;              // producer warp
  constexpr int RW = PW - TW;              // remainder warps TW .. PW-1
  constexpr int TILE = TW * SPW;           // slices per tile
  static_assert(TILE == WIN2_TILE_SLICES, "window plan and kernel disagree on the tile size");
  static_assert(RW >= 1, "no warps left for the remainder rows");
  static_assert(GV % DVB == 0, "register-resident slices are processed DVB at a time");
  constexpr int GR = GV > 0 ? GV : 1;
  constexpr int VB = 4;                    // slices per warp and round of the rows whose g lives in global memory
  unsigned long long t_prev = 0;
  const int nb = gridDim.x;
  const int n_groups = (A.rem.n_rows + 7) >> 3;  // remainder rows are walked in groups of 8 (PatView::rem4_*)
  const int rem_per_block = 8 * (n_groups / nb + 1);
  const Win2Layout lay = win2_layout(D.win_elems, A.n_pat, rows_per_block, h_smem != 0, code_smem != 0, rem_per_block);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem);  // [2]
  uint64_t *empty_bar = full + 2;                       // [2]
  double *win = reinterpret_cast<double *>(smem + lay.win_off);
  double *diagval = reinterpret_cast<double *>(smem + lay.diag_off);
  double *hs = reinterpret_cast<double *>(smem + lay.h_off);
  unsigned short *scode = reinterpret_cast<unsigned short *>(smem + lay.code_off);
  int *rrow = reinterpret_cast<int *>(smem + lay.rem_off), *rp0 = rrow + lay.rem_cap, *rlen = rp0 + lay.rem_cap;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s_begin = (int)(((int64_t)A.n_slices * blockIdx.x) / nb);
  const int s_end = (int)(((int64_t)A.n_slices * (blockIdx.x + 1)) / nb);
  const int g_begin = (int)(((int64_t)n_groups * blockIdx.x) / nb);
  const int g_end = (int)(((int64_t)n_groups * (blockIdx.x + 1)) / nb);
  const int k_begin = 8 * g_begin, k_end = min(8 * g_end, A.rem.n_rows);
  const int n_tiles = (s_end - s_begin + TILE - 1) / TILE;
  const int n_even = (A.n_rows + 1) & ~1;
  const int s_glob = s_begin + GV * WPB;   // first slice whose g lives in global memory
  const uint32_t empty_id = (uint32_t)(A.n_pat - 1);
  const PatTable T{A.ptr, A.off, A.val};  // (global memory: only rows flagged PAT_GENERAL walk it, see below)
  const size_t ll_ch = (size_t)slot_u64 * WIN2_MAX_BLOCKS;
  uint64_t *llA = ll, *llB = ll + ll_ch, *llC = ll + 2 * ll_ch, *llS = ll + 3 * ll_ch;
  uint32_t tag = tag_base;

  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(full + 1, 1);
    mbar_init(empty_bar, TW);
    mbar_init(empty_bar + 1, TW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    bad = 0;
  }
  for (int i = threadIdx.x; i < 2 * D.win_elems; i += BLOCK) win[i] = 0.0;  // (lanes off the dominant path read, but never use, window data)
  // the value of the single entry of a diagonal-only pattern (eliminated Dirichlet rows), by pattern id
  for (int p = threadIdx.x; p < A.n_pat; p += BLOCK) {
    const int k = A.ptr[p];
    diagval[p] = (A.ptr[p + 1] - k == 1) ? A.val[k] : 0.0;
  }
  // the remainder rows this block walks (its share of all of them): row, first entry, length
  for (int i = threadIdx.x; i < k_end - k_begin; i += BLOCK) {
    const int p0 = A.rem_ptr[k_begin + i];
    rrow[i] = A.rem_rows[k_begin + i];
    rp0[i] = p0;
    rlen[i] = A.rem_ptr[k_begin + i + 1] - p0;
  }
  // row codes of this block's rows (they never change): pattern id + which path the row takes
  const unsigned short *code = gcode + (size_t)s_begin * 32;
  if (code_smem) {
    for (int i = threadIdx.x; i < (s_end - s_begin) * 32; i += BLOCK) scode[i] = gcode[(size_t)s_begin * 32 + i];
    code = scode;
  }
  // h of the block's own rows: shared memory, or the block's part of the global vector
  double *hloc = h_smem ? hs : h + (size_t)s_begin * 32;
  __syncthreads();
  uint32_t tc = 0;  // tiles produced / consumed so far: stage = tc & 1, use of that stage = tc >> 1
  bool pipeline_ok = true;

  // x = 0 ; g = -b ; d = b (0 on Z, true value in dt) ; res0 = |b|
  double greg[GR];
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < GV; ++u) {
    const int s = s_begin + warp + u * WPB, r = s * 32 + lane;
    greg[u] = 0.0;
    if (s < s_end && r < A.n_rows) {
      const double bv = b[r];
      const bool z = code[(s - s_begin) * 32 + lane] & RC_Z;
      x[r] = 0.0;
      greg[u] = -bv;
      d[r] = z ? 0.0 : bv;
      if (z) dt[r] = bv;
      acc += bv * bv;
    }
  }
  for (int s = s_glob + warp; s < s_end; s += WPB) {
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
  double res2 = ll_grid_sum<true>(llC, slot_u64, nb, acc, ++tag, red, bc, &bad);
  double res = sqrt(res2);
  const double res0 = res;
  int it = 0, status = 0;
  double alpha = 0.0;
  if (res > tol && !bad) {
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
      if (warp == PW) {
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
      } else if (warp >= TW) {
        // remainder rows (this block's share of ALL of them, not the ones it owns): 8 per warp and round, 4 lanes per
        // row; h goes to global memory, the owner reads it there
        const unsigned long long tr = prof ? gtime() : 0ull;
        for (int gq = g_begin + (warp - TW); gq < g_end; gq += RW) {
          const int k = (gq - g_begin) * 8 + (lane >> 2);
          const bool on = k < k_end - k_begin;
          const int r = on ? rrow[k] : 0;
          const double dv = (on && (lane & 3) == 0) ? d[r] : 0.0;  // (in flight together with the row's entries)
          const double aq = A.rem4_long[gq] ? rem_row_dot4_at(A, on ? rp0[k] : 0, on ? rlen[k] : 0, lane & 3, d)
                                            : rem_row_dot4_lanes(A, gq, d);
          if (on && (lane & 3) == 0) {
            h[r] = aq;
            acc += dv * aq;
          }
        }
        if (prof && blockIdx.x == prof - 1 && threadIdx.x == TW * 32) g_cg_phase_ns[13] += gtime() - tr;  // (first remainder warp)
      } else if constexpr (XPAIR) {
        constexpr int PC = SPW / 2;   // chunks of 64 consecutive rows per warp and tile
        static_assert(SPW % 2 == 0, "XPAIR walks pairs of slices");
        const bool hi8 = (lane & 8) != 0;
        for (int t = 0; t < n_tiles; ++t, ++tc) {
          const int tile0 = s_begin + t * TILE;
          const uint32_t st = tc & 1u, use = tc >> 1;
          // chunk j of this warp: slices tile0 + 2 (j TW + warp) and the next one; the lane's rows are 2 lane, 2 lane + 1 of it
          uint32_t rc[PC][2];
          bool any_dom = false, any_general = false;
#pragma unroll
          for (int j = 0; j < PC; ++j) {
            const int s = tile0 + 2 * (j * TW + warp) + (lane >> 4);
            const uint32_t two = (s < s_end) ? *reinterpret_cast<const uint32_t *>(code + (s - s_begin) * 32 + 2 * (lane & 15))
                                             : (RC_EMPTY | (RC_EMPTY << 16));
            rc[j][0] = two & 0xffffu;
            rc[j][1] = two >> 16;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              any_dom = any_dom || (rc[j][m] & RC_DOM);
              any_general = any_general || !(rc[j][m] & (RC_DOM | RC_EMPTY | RC_DIAG));
            }
          }
          any_dom = __any_sync(0xffffffffu, any_dom);
          any_general = __any_sync(0xffffffffu, any_general);
          double dtv[PC][2];
#pragma unroll
          for (int j = 0; j < PC; ++j)
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              dtv[j][m] = 0.0;
              if (!(rc[j][m] & RC_DOM) && (rc[j][m] & RC_Z)) dtv[j][m] = dt[(tile0 + 2 * (j * TW + warp)) * 32 + 2 * lane + m];
            }
          GMG_PHASE(8)
          if (!mbar_wait(full + st, use & 1u)) pipeline_ok = false;
          GMG_PHASE(9)
          // byte address of the lane's first row of chunk 0 in the window; chunk j is j * TW * 64 rows further
          const char *w = reinterpret_cast<const char *>(win + st * D.win_elems + warp * 64 + 2 * lane);
          const char *wp = w + (hi8 ? 24 : 0), *wq = w + (hi8 ? 0 : 24);
          double ad[PC][2], dr[PC][2];
#pragma unroll
          for (int j = 0; j < PC; ++j) ad[j][0] = ad[j][1] = dr[j][0] = dr[j][1] = 0.0;
          if (any_dom) {
            // run r of chunk j is loaded right after run r - 1 of that chunk has been consumed: its registers are free
            // then, and the loads have the arithmetic of the other chunks to arrive (the parity of every run is a
            // compile-time constant: no branches, the whole loop is one basic block)
            double v[PC][4];
            auto load_run = [&](int r, int j) {
              const int wb = D.wbyte[3 * r];
              if ((XP_ODD >> r) & 1u) {
                const double2 mid = *reinterpret_cast<const double2 *>(w + wb + 8 + j * (TW * 64 * 8));
                const double p = *reinterpret_cast<const double *>(wp + wb + j * (TW * 64 * 8));
                const double q = *reinterpret_cast<const double *>(wq + wb + j * (TW * 64 * 8));
                v[j][0] = hi8 ? q : p;
                v[j][1] = mid.x;
                v[j][2] = mid.y;
                v[j][3] = hi8 ? p : q;
              } else {
                const double2 a2 = *reinterpret_cast<const double2 *>(w + wb + j * (TW * 64 * 8));
                const double2 b2 = *reinterpret_cast<const double2 *>(w + wb + 16 + j * (TW * 64 * 8));
                v[j][0] = a2.x;
                v[j][1] = a2.y;
                v[j][2] = b2.x;
                v[j][3] = b2.y;
              }
            };
#pragma unroll
            for (int j = 0; j < PC; ++j) load_run(0, j);
#pragma unroll
            for (int r = 0; r < XP_RUNS; ++r) {
#pragma unroll
              for (int j = 0; j < PC; ++j) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                  ad[j][m] = fma(D.val[3 * r], v[j][m], ad[j][m]);
                  ad[j][m] = fma(D.val[3 * r + 1], v[j][m + 1], ad[j][m]);
                  ad[j][m] = fma(D.val[3 * r + 2], v[j][m + 2], ad[j][m]);
                }
                if (r == XP_DIAG) {  // the run around the diagonal: d of the lane's own rows
                  dr[j][0] = v[j][1];
                  dr[j][1] = v[j][2];
                }
                if (r + 1 < XP_RUNS) load_run(r + 1, j);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < PC; ++j) {
              dr[j][0] = *reinterpret_cast<const double *>(w + D.diag_wbyte + j * (TW * 64 * 8));
              dr[j][1] = *reinterpret_cast<const double *>(w + D.diag_wbyte + 8 + j * (TW * 64 * 8));
            }
          }
          // this warp is done with the stage: hand it back to the producer
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty_bar + st)) : "memory");
          GMG_PHASE(10)
#pragma unroll
          for (int j = 0; j < PC; ++j) {
            const int li = (tile0 + 2 * (j * TW + warp) - s_begin) * 32 + 2 * lane;  // local index of the lane's first row
            const int r0 = (tile0 + 2 * (j * TW + warp)) * 32 + 2 * lane;
            // (measured: branches on the row codes, as here, beat a branch-free version of this epilogue that fetches
            // diagval for every row and selects -- 32.5 against 33.3 us per inner iteration at 64k atoms)
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              const uint32_t c = rc[j][m];
              const bool general = !(c & (RC_DOM | RC_EMPTY | RC_DIAG));
              const bool z = !(c & RC_DOM) && (c & RC_Z);
              if (z) dr[j][m] = dtv[j][m];
              if (c & RC_DIAG) ad[j][m] = fma(diagval[c & RC_ID], dr[j][m], 0.0);
              if (any_general) {  // (a dominant-compatible row with a dominant column outside the matrix: a handful at most)
                const double ag = pat_row_dot_lanes<false>(T, general ? (c & RC_ID) : empty_id, d + r0 + m, z, dr[j][m]);
                if (general) ad[j][m] = ag;
              }
              if (!(c & RC_EMPTY)) acc += dr[j][m] * ad[j][m];
            }
            // (h of a remainder row is written by the warp that walks it -- to global memory, which is where hloc points
            // when h does not fit in shared memory: never store over it)
            const bool e0 = rc[j][0] & RC_EMPTY, e1 = rc[j][1] & RC_EMPTY;
            if (!e0 && !e1) *reinterpret_cast<double2 *>(hloc + li) = make_double2(ad[j][0], ad[j][1]);
            else {
              if (!e0) hloc[li] = ad[j][0];
              if (!e1) hloc[li + 1] = ad[j][1];
            }
          }
          GMG_PHASE(11)
        }
      } else {
        for (int t = 0; t < n_tiles; ++t, ++tc) {
          const int tile0 = s_begin + t * TILE;
          const uint32_t st = tc & 1u, use = tc >> 1;
          uint32_t rc[SPW];
          bool any_dom = false, any_general = false;
#pragma unroll
          for (int j = 0; j < SPW; ++j) {
            const int s = tile0 + j * TW + warp;
            rc[j] = (s < s_end) ? code[(s - s_begin) * 32 + lane] : RC_EMPTY;
            any_dom = any_dom || (rc[j] & RC_DOM);
            any_general = any_general || !(rc[j] & (RC_DOM | RC_EMPTY | RC_DIAG));
          }
          any_dom = __any_sync(0xffffffffu, any_dom);
          any_general = __any_sync(0xffffffffu, any_general);
          // operand of the Z rows off the dominant path (eliminated rows: their own entry lives in dt): the load is in
          // flight while the warp waits for the window and runs the dominant loop
          double dtv[SPW];
#pragma unroll
          for (int j = 0; j < SPW; ++j) {
            dtv[j] = 0.0;
            if (!(rc[j] & RC_DOM) && (rc[j] & RC_Z)) dtv[j] = dt[(tile0 + j * TW + warp) * 32 + lane];
          }
          GMG_PHASE(8)
          if (!mbar_wait(full + st, use & 1u)) pipeline_ok = false;
          GMG_PHASE(9)
          // byte address of this lane's first row in the window; row j of the warp is j * TW slices further
          const char *w = reinterpret_cast<const char *>(win + st * D.win_elems + warp * 32 + lane);
          double ad[SPW];
#pragma unroll
          for (int j = 0; j < SPW; ++j) ad[j] = 0.0;
          if (any_dom) {
            constexpr int KU = SPW <= 2 ? 4 : 2;  // entries per round: KU * SPW window loads in flight per lane
#pragma unroll
            for (int k0 = 0; k0 < DOM_MAX; k0 += KU) {
              if (k0 + KU <= D.len) {
                double xv[KU][SPW];
#pragma unroll
                for (int u = 0; u < KU; ++u)
#pragma unroll
                  for (int j = 0; j < SPW; ++j)
                    xv[u][j] = *reinterpret_cast<const double *>(w + D.wbyte[k0 + u] + j * (TW * 32 * 8));
#pragma unroll
                for (int u = 0; u < KU; ++u)
#pragma unroll
                  for (int j = 0; j < SPW; ++j) ad[j] = fma(D.val[k0 + u], xv[u][j], ad[j]);
              } else if (k0 < D.len) {
#pragma unroll
                for (int u = 0; u < KU - 1; ++u)
                  if (k0 + u < D.len) {
#pragma unroll
                    for (int j = 0; j < SPW; ++j)
                      ad[j] = fma(D.val[k0 + u], *reinterpret_cast<const double *>(w + D.wbyte[k0 + u] + j * (TW * 32 * 8)), ad[j]);
                  }
              }
            }
          }
          double dr[SPW];
#pragma unroll
          for (int j = 0; j < SPW; ++j) dr[j] = *reinterpret_cast<const double *>(w + D.diag_wbyte + j * (TW * 32 * 8));
          // this warp is done with the stage: hand it back to the producer
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty_bar + st)) : "memory");
          GMG_PHASE(10)
#pragma unroll
          for (int j = 0; j < SPW; ++j) {
            const int s = tile0 + j * TW + warp;
            const int r = s * 32 + lane;
            const bool general = !(rc[j] & (RC_DOM | RC_EMPTY | RC_DIAG));
            const bool z = !(rc[j] & RC_DOM) && (rc[j] & RC_Z);
            if (z) dr[j] = dtv[j];
            if (rc[j] & RC_DIAG) ad[j] = fma(diagval[rc[j] & RC_ID], dr[j], 0.0);
            if (any_general) {  // (a dominant-compatible row with a dominant column outside the matrix: a handful at most)
              const double ag = pat_row_dot_lanes<false>(T, general ? (rc[j] & RC_ID) : empty_id, d + r, z, dr[j]);
              if (general) ad[j] = ag;
            }
            if (!(rc[j] & RC_EMPTY)) {
              hloc[(s - s_begin) * 32 + lane] = ad[j];
              acc += dr[j] * ad[j];
            }
          }
          GMG_PHASE(11)
        }
      }
      if (prof && threadIdx.x == 32) g_cg_block_ns[0][blockIdx.x & 255] += gtime() - tb;
      GMG_PHASE(0)
      alpha = gh / ll_grid_sum<true>(llA, slot_u64, nb, acc, ++tag, red, bc, &bad);
      GMG_PHASE(1)
      if (bad) { status = 3; break; }
      // ---- g += alpha h ; res2 = g.g  (x += alpha d is done together with the direction update) ----
      acc = 0.0;
      if (prof && threadIdx.x == 32) tb = gtime();
#pragma unroll
      for (int u0 = 0; u0 < GV; u0 += DVB) {  // rows whose g lives in registers; h of a remainder row: global memory
        double hv[DVB];
#pragma unroll
        for (int u = 0; u < DVB; ++u) {
          const int s = s_begin + warp + (u0 + u) * WPB, r = s * 32 + lane;
          hv[u] = 0.0;
          if (s < s_end && r < A.n_rows) {
            const int li = (s - s_begin) * 32 + lane;
            hv[u] = (code[li] & RC_EMPTY) ? h[r] : hloc[li];
          }
        }
#pragma unroll
        for (int u = 0; u < DVB; ++u) {
          const double gn = greg[u0 + u] + alpha * hv[u];  // (rows outside the block: 0 + alpha * 0)
          greg[u0 + u] = gn;
          acc += gn * gn;
        }
      }
      for (int s0 = s_glob + warp; s0 < s_end; s0 += VB * WPB) {  // VB slices per round: all loads in flight at once
        double gv[VB], hv[VB];
#pragma unroll
        for (int u = 0; u < VB; ++u) {
          const int s = s0 + u * WPB, r = s * 32 + lane;
          gv[u] = hv[u] = 0.0;
          if (s < s_end && r < A.n_rows) {
            const int li = (s - s_begin) * 32 + lane;
            hv[u] = (code[li] & RC_EMPTY) ? h[r] : hloc[li];
            gv[u] = g[r];
          }
        }
#pragma unroll
        for (int u = 0; u < VB; ++u) {
          const int s = s0 + u * WPB, r = s * 32 + lane;
          if (s < s_end && r < A.n_rows) {
            const double gn = gv[u] + alpha * hv[u];
            g[r] = gn;
            acc += gn * gn;
          }
        }
      }
      if (prof && threadIdx.x == 32) g_cg_block_ns[1][blockIdx.x & 255] += gtime() - tb;
      GMG_PHASE(2)
      ++tag;
      ll_publish<false>(llB, slot_u64, acc, tag, red);
      if (XSPLIT) {
        // x += alpha d needs alpha only: it runs while the other blocks' partial sums of g.g arrive (the reduction's latency
        // is hidden behind 24 bytes per row of L2 traffic; d is read once more in the direction update: 40 instead of 32
        // bytes per row in total, same arithmetic)
        for (int s0 = s_begin + warp; s0 < s_end; s0 += DVB * WPB) {
          double dv[DVB], xv[DVB];
#pragma unroll
          for (int u = 0; u < DVB; ++u) {
            const int s = s0 + u * WPB, r = s * 32 + lane;
            dv[u] = xv[u] = 0.0;
            if (s < s_end && r < A.n_rows) {
              dv[u] = (code[(s - s_begin) * 32 + lane] & RC_Z) ? dt[r] : d[r];
              xv[u] = x[r];
            }
          }
#pragma unroll
          for (int u = 0; u < DVB; ++u) {
            const int s = s0 + u * WPB, r = s * 32 + lane;
            if (s < s_end && r < A.n_rows) x[r] = xv[u] + alpha * dv[u];
          }
        }
      }
      res2 = ll_await<false>(llB, slot_u64, nb, tag, bc, &bad);
      GMG_PHASE(3)
      if (bad) { status = 3; break; }
      res = sqrt(res2);
      if (res <= tol) break;
      if (it >= max_it) { status = 1; break; }
      const double beta = res2 / gh;
      gh = res2;
      // ---- (x += alpha d ;) d = beta d - g -----------------------------------------------------
      if (prof && threadIdx.x == 32) tb = gtime();
#pragma unroll
      for (int u0 = 0; u0 < GV; u0 += DVB) {
        double dv[DVB], xv[DVB];
        bool z[DVB];
#pragma unroll
        for (int u = 0; u < DVB; ++u) {
          const int s = s_begin + warp + (u0 + u) * WPB, r = s * 32 + lane;
          dv[u] = xv[u] = 0.0;
          z[u] = false;
          if (s < s_end && r < A.n_rows) {
            z[u] = code[(s - s_begin) * 32 + lane] & RC_Z;
            dv[u] = z[u] ? dt[r] : d[r];
            if (!XSPLIT) xv[u] = x[r];
          }
        }
#pragma unroll
        for (int u = 0; u < DVB; ++u) {
          const int s = s_begin + warp + (u0 + u) * WPB, r = s * 32 + lane;
          if (s < s_end && r < A.n_rows) {
            if (!XSPLIT) x[r] = xv[u] + alpha * dv[u];
            const double dn = beta * dv[u] - greg[u0 + u];
            if (z[u]) dt[r] = dn;  // (d[r] stays 0)
            else d[r] = dn;
          }
        }
      }
      for (int s0 = s_glob + warp; s0 < s_end; s0 += VB * WPB) {
        double dv[VB], xv[VB], gv[VB];
        bool z[VB];
#pragma unroll
        for (int u = 0; u < VB; ++u) {
          const int s = s0 + u * WPB, r = s * 32 + lane;
          dv[u] = xv[u] = gv[u] = 0.0;
          z[u] = false;
          if (s < s_end && r < A.n_rows) {
            z[u] = code[(s - s_begin) * 32 + lane] & RC_Z;
            dv[u] = z[u] ? dt[r] : d[r];
            if (!XSPLIT) xv[u] = x[r];
            gv[u] = g[r];
          }
        }
#pragma unroll
        for (int u = 0; u < VB; ++u) {
          const int s = s0 + u * WPB, r = s * 32 + lane;
          if (s < s_end && r < A.n_rows) {
            if (!XSPLIT) x[r] = xv[u] + alpha * dv[u];
            const double dn = beta * dv[u] - gv[u];
            if (z[u]) dt[r] = dn;  // (d[r] stays 0)
            else d[r] = dn;
          }
        }
      }
      fence_proxy_async();
      if (prof && threadIdx.x == 32) g_cg_block_ns[2][blockIdx.x & 255] += gtime() - tb;
      GMG_PHASE(4)
      ll_grid_sum<true>(llC, slot_u64, nb, 0.0, ++tag, red, bc, &bad);
      GMG_PHASE(5)
      if (bad) { status = 3; break; }
    }
    // the x update of the last iteration (XSPLIT: already done before the convergence test)
    if (status != 3 && !XSPLIT)
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        const int r = s * 32 + lane;
        if (r < A.n_rows) {
          const bool z = code[(s - s_begin) * 32 + lane] & RC_Z;
          x[r] += alpha * (z ? dt[r] : d[r]);
        }
      }
  }
  // a timed-out window transaction or barrier (never observed) is reported as status 2 / 3 instead of a wrong answer
  const int lost = __syncthreads_or(pipeline_ok ? 0 : 1);
  const double n_bad = (status == 3 || bad) ? 1.0
                                            : ll_grid_sum<true>(llS, slot_u64, nb, (lost && threadIdx.x == 0) ? 1.0 : 0.0, ++tag, red, bc, &bad);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    result->iterations = it;
    result->status = (status == 3 || bad) ? 3 : n_bad != 0.0 ? 2 : status;
    result->res0 = res0;
    result->res = res;
  }
}

}  // namespace gmg
