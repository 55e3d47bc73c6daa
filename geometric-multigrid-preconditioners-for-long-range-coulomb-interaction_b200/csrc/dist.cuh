// Multi-GPU primitives over NVLink peer memory (CUDA IPC), one process per GPU.
//
// Every rank owns one "symmetric" communication buffer (same layout on all ranks) that all peers map.
// All communication is  P2P store of the payload -> __threadfence_system -> st.release.sys of a
// sequence-numbered flag in the destination's buffer ; the consumer spins with ld.acquire.sys until
// flag >= seq.  There is no separate exchange step and no NCCL call on the data path: the halo rows of
// a vector are written straight into the neighbour's extended vector, all-reduces are W x W scalar
// pushes summed in rank order (bit-identical on every rank).  Waits are bounded (clock64) and report
// an error instead of hanging.
#pragma once
#include "common.cuh"

namespace gmg {

constexpr int DIST_MAX_RANKS = 8;
constexpr int DIST_NCHAN = 32;
constexpr size_t DIST_HEADER_BYTES = 8192;
// header layout: uint64 flag[DIST_NCHAN][8] @0 ; double red[2 areas][2][8][4] @2048 ; double gsum[4] @3072
// (area 0: stream-ordered dist_allreduce, area 1: the persistent CG kernel -- never share slots)
constexpr size_t DIST_OFF_RED = 2048, DIST_OFF_GSUM = 3072;
enum { CH_RED = 0, CH_HALO_SYS_X = 1, CH_HALO_SYS_D = 2, CH_GATHER_G = 3, CH_GATHER_C = 4, CH_CG_RED = 5, CH_CG_HALO = 6,
       CH_GATHER_X = 7, CH_BARRIER = 8, CH_REV = 9 };

struct DistPeers {
  int rank, world;
  char *peer[DIST_MAX_RANKS];  // peer[rank] = own buffer
};

__device__ __forceinline__ void flag_store(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// polling uses relaxed loads: an acquire load at system scope is a load + a system-wide fence, and hundreds of
// spinning warps issuing fences starve the very store they wait for; one fence follows the successful poll
__device__ __forceinline__ uint64_t flag_load(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t *dist_flag(char *base, int channel, int src_rank) {
  return reinterpret_cast<uint64_t *>(base) + channel * 8 + src_rank;
}
__device__ __forceinline__ double *dist_red(char *base, int parity, int src_rank, int area = 0) {
  return reinterpret_cast<double *>(base + DIST_OFF_RED + 512 * area) + (parity * 8 + src_rank) * 4;
}
// bounded spin
__device__ __forceinline__ bool wait_flag(const uint64_t *p, uint64_t seq) {
  const long long t0 = clock64();
  while (flag_load(p) < seq) {
    if (clock64() - t0 > 40000000000LL) return false;  // ~20 s
    __nanosleep(40);
  }
  return true;
}
// warp-collective: lanes q < world with bit q of `mask` wait for flag[channel][q] >= seq in the local buffer;
// lane 0 then re-reads the flags and issues the acquire fence for the block (follow with __syncthreads/__syncwarp)
__device__ __forceinline__ bool warp_wait_flags(char *mine, int channel, uint64_t seq, uint32_t mask, int world) {
  const int lane = threadIdx.x & 31;
  bool ok = true;
  if (lane < world && (mask >> lane & 1u)) ok = wait_flag(dist_flag(mine, channel, lane), seq);
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    for (int q = 0; q < world; ++q)
      if (mask >> q & 1u) (void)flag_load(dist_flag(mine, channel, q));
    __threadfence_system();
  }
  __syncwarp();
  return ok;
}

// payload[t] -> peer's region ; then flags to every rank in dst_mask (even when there is no payload)
__global__ void __launch_bounds__(256) dist_push(DistPeers P, int n, const int *__restrict__ src_idx,
                                                 const unsigned char *__restrict__ dst_peer,
                                                 const int *__restrict__ dst_idx, size_t region_off,
                                                 const double *__restrict__ src, int channel, uint64_t seq,
                                                 uint32_t dst_mask, unsigned int *counter) {
  __shared__ bool last;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) reinterpret_cast<double *>(P.peer[dst_peer[t]] + region_off)[dst_idx[t]] = src[src_idx[t]];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // release of this block's stores (cumulative over the barrier), gpu scope
    const unsigned int c = atomicInc(counter, gridDim.x - 1);
    last = (c == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < P.world && (dst_mask >> threadIdx.x & 1u)) {
    __threadfence_system();
    flag_store(dist_flag(P.peer[threadIdx.x], channel, P.rank), seq);
  }
}

__global__ void dist_wait(DistPeers P, int channel, uint64_t seq, uint32_t src_mask, int *error) {
  if (!warp_wait_flags(P.peer[P.rank], channel, seq, src_mask, P.world) && threadIdx.x == 0) *error = 1;
}

// in-place sum over ranks of k <= 4 consecutive doubles at v (rank order: identical bits everywhere)
__global__ void dist_allreduce(DistPeers P, double *v, int k, uint64_t seq, int *error) {
  const int q = threadIdx.x;
  const int par = (int)(seq & 1);
  if (q < P.world) {
    double *slot = dist_red(P.peer[q], par, P.rank);
    for (int i = 0; i < k; ++i) slot[i] = v[i];
    __threadfence_system();
    flag_store(dist_flag(P.peer[q], CH_RED, P.rank), seq);
  }
  if (!warp_wait_flags(P.peer[P.rank], CH_RED, seq, (1u << P.world) - 1u, P.world) && q == 0) *error = 1;
  if (q == 0)
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int r = 0; r < P.world; ++r) s += dist_red(P.peer[P.rank], par, r)[i];
      v[i] = s;
    }
}

// reverse halo exchange, receiving side: what peer p accumulated for my owned entry send_src[k] sits in my staging area
// at [p][the position send_dst[k] of that entry in p's extended vector]
__global__ void dist_rev_add(int n, const int *__restrict__ send_src, const unsigned char *__restrict__ send_peer,
                             const int *__restrict__ send_dst, const double *__restrict__ staging, int stride,
                             double *__restrict__ owned) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) atomicAdd(owned + send_src[k], staging[(size_t)send_peer[k] * stride + send_dst[k]]);
}

// scatter/gather helpers on local data
__global__ void vec_gather_from(int n, const int *__restrict__ dst_idx, const double *__restrict__ src, double *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[dst_idx[i]] = src[i];  // dst[dst_idx[i]] = src[i] (src contiguous)
}

__global__ void vec_take(int n, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// ---- "LL" words (as in NCCL's low-latency protocol): 8-byte stores carry 4 bytes of payload and a 4-byte tag,
// so the payload needs no fence and no separate flag: a double travels as two tagged words in one 16-byte store.
constexpr size_t DIST_OFF_LLRED = 4096;  // uint64 llred[4][8][2] in the header: [launch parity * 2 + reduction parity][rank]
__device__ __forceinline__ void ll_store(uint64_t *p /*16-byte aligned pair*/, double v, uint32_t tag) {
  const uint64_t bits = (uint64_t)__double_as_longlong(v);
  const uint64_t w0 = ((uint64_t)tag << 32) | (bits & 0xffffffffull);
  const uint64_t w1 = ((uint64_t)tag << 32) | (bits >> 32);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
// spin until both words carry `tag`; false on timeout
__device__ __forceinline__ bool ll_load(const uint64_t *p, uint32_t tag, double &v) {
  const long long t0 = clock64();
  uint64_t w0, w1;
  while (true) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
    if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag) break;
    if (clock64() - t0 > 40000000000LL) return false;
  }
  v = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
  return true;
}

// latency probe: rank 0 sends tag i to rank 1, which echoes it; cycles for `iters` round trips -> out[0].
// mode 0: tagged LL store only; mode 1: LL store followed by __threadfence_system()
__global__ void dist_pingpong(DistPeers P, size_t region, int iters, int mode, uint32_t tag_base, long long *out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint64_t *mine = reinterpret_cast<uint64_t *>(P.peer[P.rank] + region);
  uint64_t *other = reinterpret_cast<uint64_t *>(P.peer[P.rank ^ 1] + region);
  double v;
  const long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    const uint32_t tag = tag_base + i;
    if (P.rank == 0) {
      ll_store(other, (double)i, tag);
      if (mode == 1) __threadfence_system();
      if (!ll_load(mine, tag, v)) break;
    } else {
      if (!ll_load(mine, tag, v)) break;
      ll_store(other, v, tag);
      if (mode == 1) __threadfence_system();
    }
  }
  out[0] = clock64() - t0;
}

constexpr int DIST_CG_MAXB = 160;  // blocks of the distributed CG kernel a rank may run (one per SM; slots per rank and group)
constexpr size_t DIST_CG_SLOT_BYTES = (size_t)6 * DIST_MAX_RANKS * DIST_CG_MAXB * 16;

struct DistCgArgs {
  DistPeers P;
  int n_owned, n_halo, n_halo_lo;
  int n_send;
  const int *send_src;          // sorted by source row
  const unsigned char *send_peer;
  const int *send_hpos;         // position in the destination's halo
  const int *send_block_ptr;    // gridDim.x + 1: entries whose source row belongs to each block's slice range
  size_t region_d;     // symmetric offset of the extended direction vector [lower halo | owned | upper halo]
  size_t region_ll;    // symmetric offset of the LL receive area: 16 bytes per halo entry
  int prof;            // gmg_debug_cg_phases: block (1-based) whose thread 0 times the phases of an iteration
  size_t region_slots; // symmetric offset of the reduction / barrier slots: [6 groups][8 ranks][DIST_CG_MAXB] x 16 bytes
  uint32_t tag_base;   // launch id * 2^20: tags used inside the kernel are tag_base + counter (never 0)
};

// ------------------------------------------------------------------------------------------------
// distributed persistent CG: the single-GPU cg_persistent plus, inside the same kernel, the halo rows of d pushed
// straight into the neighbours (by the block that just updated them) and the all-reduces of d.h and g.g, all as tagged
// LL words: no fences on the NVLink path, no flags, no extra exchange kernels, and NO grid.sync():
//   * a reduction is ONE all-to-all step: every block stores {its partial sum, tag} into its slot on EVERY rank (W
//     16-byte stores, W - 1 of them over NVLink); warp r of every block polls rank r's slots in local memory and adds
//     them in a fixed order, the W totals are added in rank order: every block of every rank obtains the same bits one
//     NVLink latency after the last block arrived (before: grid.sync + block 0 sums the partials + NVLink + poll);
//   * the barrier after the direction update is rank-local (tagged slots of the rank's own blocks); the halo entries
//     synchronise themselves through their tags.
// The launch stays cooperative (co-residency of all blocks is what the spinning needs).
// ------------------------------------------------------------------------------------------------
template <int BLOCK, class MAT>
__global__ void __launch_bounds__(BLOCK, 1) cg_persistent_dist(MAT A, const double *__restrict__ b, double *x, double *g,
                                                            double *h, double *partials, int max_it, double tol,
                                                            CgResult *result, DistCgArgs D, int *error) {
  __shared__ double red[64];
  __shared__ double s_tot[2][DIST_MAX_RANKS];
  __shared__ RowSmem<MAT> row_smem;
  RowDot<MAT> row_dot;
  row_dot.init(A, row_smem);
  const int nb = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int WPB = BLOCK / 32;
  static_assert(WPB >= DIST_MAX_RANKS, "one polling warp per rank");
  const int s_begin = (int)(((int64_t)A.n_slices * blockIdx.x) / nb);
  const int s_end = (int)(((int64_t)A.n_slices * (blockIdx.x + 1)) / nb);
  char *mine = D.P.peer[D.P.rank];
  double *d_ext = reinterpret_cast<double *>(mine + D.region_d);
  double *d = d_ext + D.n_halo_lo;  // first owned entry; lower-rank halo entries sit at negative indices
  const uint64_t *ll_in = reinterpret_cast<const uint64_t *>(mine + D.region_ll);
  uint32_t nred = 0, nhalo = 0, nbar = 0;
  const int prof = D.prof;
  unsigned long long t_prev = 0;
  // A wait that times out (~20 s) sets *error; every other wait of the kernel is bounded too, so all blocks leave.
  bool aborted = false;
  volatile int *abort_flag = error;
  auto slot_of = [&](char *base, int group, int rank, int block) {
    return reinterpret_cast<uint64_t *>(base + D.region_slots) + 2 * (((size_t)group * DIST_MAX_RANKS + rank) * DIST_CG_MAXB + block);
  };

  // sum over all blocks of all ranks; `fence`: the step also orders the rank's global writes before it against the
  // reads after it (h of the remainder rows is written by one block and read by another)
  auto all_sum = [&](double v, bool fence) -> double {
    ++nred;
    const uint32_t tag = D.tag_base + nred;
    // Slot group = launch parity x reduction parity.  A block can be one reduction ahead of the slowest block (never
    // two: reduction n + 1 completes only when every block has published it, i.e. has finished collecting n), and a rank
    // one launch ahead of a slow peer block still polling the last reduction of the previous launch.
    const int par = (int)(nred & 1u);
    const int group = (int)(((D.tag_base >> 20) & 1u) * 2u) + par;
    v = warp_sum(v);
    if (lane == 0) red[par * 32 + warp] = v;
    __syncthreads();
    if (warp == 0) {
      double r = lane < WPB ? red[par * 32 + lane] : 0.0;
      r = warp_sum(r);
      if (fence && lane == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      __syncwarp();
      if (lane < D.P.world) ll_store(slot_of(D.P.peer[lane], group, D.P.rank, blockIdx.x), r, tag);
    }
    if (warp < D.P.world) {
      double s;
      const bool ok = ll_collect_slots<DIST_CG_MAXB / 32, true>(slot_of(mine, group, warp, 0), 2, nb, tag, 40000000000LL, s);
      if (lane == 0) {
        if (fence && warp == D.P.rank) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        s_tot[par][warp] = s;
        if (!ok) *abort_flag = 1;
      }
    }
    if (__syncthreads_or(*abort_flag != 0)) aborted = true;  // (block-uniform)
    double tot = 0.0;
    for (int r = 0; r < D.P.world; ++r) tot += s_tot[par][r];  // rank order
    return aborted ? 0.0 : tot;
  };
  // barrier over the blocks of this rank, with release / acquire of their global writes (the owned rows and the unpacked
  // halo entries of d)
  auto rank_barrier = [&]() {
    ++nbar;
    const uint32_t tag = D.tag_base + (1u << 18) + nbar;
    const int group = 4 + (int)(nbar & 1u);
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        ll_store(slot_of(mine, group, D.P.rank, blockIdx.x), 0.0, tag);
      }
      double s;
      const bool ok = ll_collect_slots<DIST_CG_MAXB / 32, true>(slot_of(mine, group, D.P.rank, 0), 2, nb, tag, 40000000000LL, s);
      if (lane == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        if (!ok) *abort_flag = 1;
      }
    }
    if (__syncthreads_or(*abort_flag != 0)) aborted = true;
  };
  // after this block's rows of d are written: push its boundary rows, then help unpacking the incoming halo
  auto exchange_halo = [&]() {
    ++nhalo;
    const uint32_t tag = D.tag_base + (1u << 19) + nhalo;
    __syncthreads();
    for (int t = D.send_block_ptr[blockIdx.x] + threadIdx.x; t < D.send_block_ptr[blockIdx.x + 1]; t += BLOCK)
      ll_store(reinterpret_cast<uint64_t *>(D.P.peer[D.send_peer[t]] + D.region_ll) + 2 * (size_t)D.send_hpos[t],
               d[D.send_src[t]], tag);
    bool ok = true;
    for (int k = blockIdx.x * BLOCK + threadIdx.x; k < D.n_halo; k += nb * BLOCK) {
      double v;
      if (ll_load(ll_in + 2 * (size_t)k, tag, v)) d_ext[k < D.n_halo_lo ? k : k + D.n_owned] = v;
      else ok = false;
    }
    if (!ok) *abort_flag = 1;
  };

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
  exchange_halo();
  rank_barrier();  // d (owned rows and unpacked halo) visible to every block
  double res2 = all_sum(acc, false);
  double res = sqrt(res2);
  const double res0 = res;
  int it = 0, status = 0;
  if (aborted) status = 2;
  if (res > tol && !aborted) {
    double gh = res * res;
    while (true) {
      ++it;
      if (prof && blockIdx.x == prof - 1 && threadIdx.x == 0) {
        t_prev = gtime();
        g_cg_phase_ns[6] += 1;
      }
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
      GMG_PHASE(0)
      const double alpha = gh / all_sum(acc, true);
      GMG_PHASE(1)
      if (aborted) { status = 2; break; }
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
      GMG_PHASE(2)
      res2 = all_sum(acc, false);
      GMG_PHASE(3)
      res = sqrt(res2);
      if (aborted) { status = 2; break; }
      if (res <= tol) break;
      if (it >= max_it) { status = 1; break; }
      const double beta = res2 / gh;
      gh = res2;
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        const int r = s * 32 + lane;
        if (r < A.n_rows) d[r] = beta * d[r] - g[r];
      }
      GMG_PHASE(4)
      exchange_halo();
      GMG_PHASE(7)
      rank_barrier();  // d (owned rows and unpacked halo) visible to every block
      GMG_PHASE(5)
      if (aborted) { status = 2; break; }
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
