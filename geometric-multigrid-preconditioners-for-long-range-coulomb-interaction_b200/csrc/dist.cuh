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
       CH_GATHER_X = 7, CH_BARRIER = 8 };

struct DistPeers {
  int rank, world;
  char *peer[DIST_MAX_RANKS];  // peer[rank] = own buffer
};

__device__ __forceinline__ void flag_store(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t flag_load(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
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
    __nanosleep(20);
  }
  return true;
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
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
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
  const int q = threadIdx.x;
  if (q < P.world && (src_mask >> q & 1u))
    if (!wait_flag(dist_flag(P.peer[P.rank], channel, q), seq)) *error = 1;
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
    if (!wait_flag(dist_flag(P.peer[P.rank], CH_RED, q), seq)) *error = 1;
  }
  __syncwarp();
  if (q == 0)
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int r = 0; r < P.world; ++r) s += dist_red(P.peer[P.rank], par, r)[i];
      v[i] = s;
    }
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

struct DistCgArgs {
  DistPeers P;
  int n_send;
  const int *send_src;
  const unsigned char *send_peer;
  const int *send_dst;
  size_t region_d;     // symmetric offset of the extended direction vector d = [owned | halo]
  uint32_t dst_mask, src_mask;
  uint64_t seq_base;   // (launch id << 32): sequence numbers used inside the kernel are seq_base + counter
};

// ------------------------------------------------------------------------------------------------
// distributed persistent CG: the single-GPU cg_persistent plus, inside the same cooperative kernel,
// the halo push of d to the neighbours' extended vectors and the cross-GPU all-reduces of d.h and g.g.
// ------------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) cg_persistent_dist(SellView A, const double *__restrict__ b, double *x, double *g,
                                                            double *h, double *partials, int max_it, double tol,
                                                            CgResult *result, DistCgArgs D, int *error) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[32];
  __shared__ int s_abort;
  const int nb = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int WPB = BLOCK / 32;
  const int s_begin = (int)(((int64_t)A.n_slices * blockIdx.x) / nb);
  const int s_end = (int)(((int64_t)A.n_slices * (blockIdx.x + 1)) / nb);
  char *mine = D.P.peer[D.P.rank];
  double *d = reinterpret_cast<double *>(mine + D.region_d);
  volatile double *gsum = reinterpret_cast<volatile double *>(mine + DIST_OFF_GSUM);
  volatile int *abort_flag = error;
  uint64_t nred = 0, nhalo = 0;

  // sum over the grid and over the ranks; 2 grid syncs
  auto all_sum = [&](double block_value) -> double {
    if (threadIdx.x == 0) partials[blockIdx.x] = block_value;
    grid.sync();
    ++nred;
    const int slot = (int)(nred % 3);
    if (blockIdx.x == 0 && warp == 0) {
      const double s = warp_sum_partials(partials, nb);
      const uint64_t seq = D.seq_base + nred;
      const int par = (int)(seq & 1);
      bool ok = true;
      if (lane < D.P.world) {
        dist_red(D.P.peer[lane], par, D.P.rank, 1)[0] = s;
        __threadfence_system();
        flag_store(dist_flag(D.P.peer[lane], CH_CG_RED, D.P.rank), seq);
        ok = wait_flag(dist_flag(mine, CH_CG_RED, lane), seq);
      }
      if (!ok) *abort_flag = 1;
      __syncwarp();
      if (lane == 0) {
        double tot = 0.0;
        for (int r = 0; r < D.P.world; ++r) tot += dist_red(mine, par, r, 1)[0];
        gsum[slot] = tot;
        __threadfence();
      }
    }
    grid.sync();
    return gsum[slot];
  };
  // push my boundary values of d into the neighbours' halos and wait for theirs; 2 grid syncs
  auto halo = [&]() {
    ++nhalo;
    const uint64_t seq = D.seq_base + nhalo;
    for (int t = blockIdx.x * BLOCK + threadIdx.x; t < D.n_send; t += nb * BLOCK)
      reinterpret_cast<double *>(D.P.peer[D.send_peer[t]] + D.region_d)[D.send_dst[t]] = d[D.send_src[t]];
    __threadfence_system();
    grid.sync();
    if (blockIdx.x == 0 && warp == 0 && lane < D.P.world) {
      if (D.dst_mask >> lane & 1u) flag_store(dist_flag(D.P.peer[lane], CH_CG_HALO, D.P.rank), seq);
      if (D.src_mask >> lane & 1u)
        if (!wait_flag(dist_flag(mine, CH_CG_HALO, lane), seq)) *abort_flag = 1;
    }
    grid.sync();
  };
  auto aborted = [&]() -> bool {
    if (threadIdx.x == 0) s_abort = *abort_flag;
    __syncthreads();
    const int a = s_abort;
    __syncthreads();
    return a != 0;
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
  acc = block_sum(acc, red);
  double res2 = all_sum(acc);
  double res = sqrt(res2);
  const double res0 = res;
  int it = 0, status = 0;
  if (res > tol && !aborted()) {
    double gh = res * res;
    halo();
    while (true) {
      if (aborted()) { status = 2; break; }
      ++it;
      acc = 0.0;
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        const double ad = sell_row_dot<false>(A, s, lane, d);
        const int r = s * 32 + lane;
        if (r < A.n_rows) {
          h[r] = ad;
          acc += d[r] * ad;
        }
      }
      acc = block_sum(acc, red);
      const double alpha = gh / all_sum(acc);
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
      res2 = all_sum(acc);
      res = sqrt(res2);
      if (res <= tol) break;
      if (it >= max_it) { status = 1; break; }
      const double beta = res2 / gh;
      gh = res2;
      for (int s = s_begin + warp; s < s_end; s += WPB) {
        const int r = s * 32 + lane;
        if (r < A.n_rows) d[r] = beta * d[r] - g[r];
      }
      grid.sync();  // all of d written before anyone pushes / reads it
      halo();
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
