// Host-side state behind a gmg_handle (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <emmintrin.h>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gmg_b200.h"
#include "common.cuh"
#include "partition.h"

namespace gmg {

struct HostCsr {
  int n_rows = 0, n_cols = 0;
  std::vector<int64_t> rowptr;
  std::vector<int> col;
  std::vector<double> val;
  bool empty() const { return rowptr.empty(); }
  int64_t nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
};

struct DevCsr {
  bool in_arena = false;  // arrays live in an Arena of the context (not freed individually)
  int n_rows = 0, n_cols = 0;
  int64_t nnz = 0;
  int64_t *rowptr = nullptr;
  int *col = nullptr;
  double *val = nullptr;
};

struct Sell {
  SellView v{};
  int64_t *slice_ptr = nullptr;
  double *val = nullptr;
  int *col = nullptr;
  int64_t stored_nnz = 0;  // entries kept from the CSR (explicit zeros included unless dropped)
  int64_t padded = 0;      // elements in the sliced-ELL arrays
  bool valid = false;
  bool shares_structure = false;  // slice_ptr / col belong to another Sell (A + I shares A's)
  std::vector<int64_t> h_slice_ptr;
  // optional lossless compressed copy (CsellView): 4 bytes per entry
  bool compressed = false;
  CsellView cv{};
  int64_t *cslice_ptr = nullptr;
  uint32_t *ent = nullptr;
  double *dict = nullptr;
  int64_t cpadded = 0;
  // optional row-pattern dictionary copy (PatView, pattern.cuh): 4 bytes per row
  bool patterned = false;
  PatView pv{};
  uint32_t *pat = nullptr;
  int *pat_ptr = nullptr, *pat_off = nullptr;
  double *pat_val = nullptr;
  int *rem_rows = nullptr;
  int64_t *rem_slice_ptr = nullptr;
  double *rem_val = nullptr;
  int *rem_col = nullptr;
  int64_t rem_padded = 0;
  int *rem_ptr = nullptr, *rem_ccol = nullptr;  // the remainder rows as CSR
  double *rem_cval = nullptr;
  int *rem4_col = nullptr;                      // ... and in the lane order of the window kernel (PatView::rem4_*)
  double *rem4_val = nullptr;
  unsigned char *rem4_long = nullptr;
  // dominant pattern + TMA window plan for the persistent CG (pattern_win.cuh); dom.len == 0: not available
  DomPat dom{};
  DomPat dom2{};         // the same pattern planned for the tile size of pattern_win2.cuh
  bool win2_ok = false;  // every table pattern is dominant-compatible or a single diagonal entry
  uint32_t *dom_mask = nullptr;
  unsigned short *row_code = nullptr;  // 16-bit row codes of the window kernel (pat_row_codes)
};

struct ColorSet {
  Sell A;          // rows of this colour
  int *rows = nullptr;
  int n = 0;
  std::vector<int> h_rows;
};

struct Level {
  int n = 0;
  DevCsr rawA;                 // freed after setup
  HostCsr hA, hI, hP;          // host copies of the small operators (levels >= 1; P: this level -> next)
  Sell A, AI, IT, P, R;        // A_l ; A_l + I_l ; I_l^T ; P_l (l -> l+1) ; R_l = P_l^T
  double *dinv = nullptr;
  double *defect = nullptr, *sol = nullptr, *t = nullptr, *tmp = nullptr;
  int n_copy = 0;
  int *copy_g = nullptr, *copy_l = nullptr;
  std::vector<ColorSet> colors;      // multicolour SSOR
  std::vector<ColorSet> wave_fwd, wave_bwd;  // level-scheduled lexicographic SSOR
  void *d_fwd = nullptr, *d_bwd = nullptr;   // device arrays of ColorView for the persistent SSOR kernel
  int n_fwd = 0, n_bwd = 0, ssor_grid = 0;
  int cluster_blocks = 0;            // > 0: small level, smooth() as one thread-block cluster of this many blocks
  double lambda_max = 0.0;           // Chebyshev
  bool edge_free = false;             // the interface matrix of the level has no entry (uniform level)
  std::vector<int32_t> user_color;   // optional colouring handed over by the host (gmg_set_level_coloring)
};

// one distributed (row-partitioned) matrix: local SELL over [owned | halo] columns + its halo send lists
struct DistMat {
  Sell A;
  int n_owned = 0, n_halo = 0, n_halo_lo = 0, n_send = 0;
  int *send_src = nullptr, *send_dst = nullptr, *send_hpos = nullptr;
  unsigned char *send_peer = nullptr;
  uint32_t dst_mask = 0, src_mask = 0;
  std::vector<int> h_send_src;  // host copy (sorted by source row)
  // reverse halo exchange (gmg_matrix_norms: column sums of halo columns go back to their owners): every halo entry
  // of my extended vector is pushed to its owner's staging area [sender rank][position in the sender's extended vector]
  int n_rev = 0, rev_stride = 0;
  int *rev_src = nullptr, *rev_dst = nullptr;
  unsigned char *rev_peer = nullptr;
  size_t rev_region = 0;
};

// "push my owned entries of a list to every rank" (all-gather over peer memory)
struct GatherPlan {
  int n_total = 0, n_send = 0;  // n_send = owned entries x world
  int *send_src = nullptr, *send_dst = nullptr;
  unsigned char *send_peer = nullptr;
  size_t region = 0;
};

struct DistData {
  bool on = false;
  int rank = 0, world = 1;
  char *buf = nullptr;
  char *peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t bytes = 0, bump = 0;
  uint64_t seq[32] = {0};
  uint64_t launch_id = 0;
  int *d_error = nullptr;
  std::vector<int32_t> sys_owner, l0_owner;
  HostCsr hS, hA0;
  std::vector<std::vector<int32_t>> h_copy_g, h_copy_l;  // copy indices as handed over (global numbering)
  DistMat S, A0;
  size_t reg_cg_d = 0, reg_cg_ll = 0, reg_cg_slots = 0, reg_pcg_d = 0, reg_pcg_x = 0;
  int n_sys_owned = 0, n_l0_owned = 0, n_sys = 0, n_l0 = 0;
  int *sys_owned_global = nullptr;      // device: global index of each owned system dof
  // copy_to_mg / copy_from_mg
  int n_copy0 = 0;
  int *copy0_sys = nullptr, *copy0_l0 = nullptr;  // owned pairs on level 0 (local indices)
  GatherPlan gather_g;                            // defect entries of the replicated levels
  std::vector<int> gather_g_offset;               // per level >= 1: offset into the gathered list
  std::vector<int> n_from;                        // per level >= 1: owned pairs for copy_from_mg
  std::vector<int *> from_sys, from_lvl;
  // level 0 <-> level 1 transfer
  Sell R0, P0F;
  GatherPlan gather_c;
  GatherPlan gather_x;  // solution all-gather (global positions)
  double *l0_defect = nullptr, *l0_sol = nullptr, *cg_g = nullptr, *cg_h = nullptr;
  double *g = nullptr, *hh = nullptr;  // outer PCG owned vectors
  double *cg_partials = nullptr;
  int *cg_send_block_ptr = nullptr;
  int cg_grid = 0;
};

}  // namespace gmg

namespace gmg {
// memcpy with non-temporal stores (the destination -- a pinned staging buffer or the caller's array -- is not read
// again by this core: no read-for-ownership traffic, no cache pollution)
inline void stream_copy(char *dst, const char *src, size_t n) {
#if defined(__SSE2__)
  while (n > 0 && (reinterpret_cast<uintptr_t>(dst) & 15u)) {
    *dst++ = *src++;
    --n;
  }
  size_t blocks = n / 64;
  for (; blocks > 0; --blocks, src += 64, dst += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 48));
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst), a);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 48), d);
  }
  _mm_sfence();
  n &= 63;
#endif
  if (n) std::memcpy(dst, src, n);
}

// A few persistent host threads that copy one chunk in parallel (the staging ring of staged_h2d / staged_d2h):
// spawning threads per 32 MB chunk cost ~0.25 ms per chunk.
class CopyPool {
 public:
  explicit CopyPool(int n) : stop_(false), job_(0), pending_(0) {
    for (int t = 0; t < n; ++t) workers_.emplace_back([this, t, n]() { run(t, n); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++job_;
    }
    cv_.notify_all();
    for (auto &w : workers_) w.join();
  }
  // dst[0, n) = src[0, n), split evenly over the workers; returns when done
  void copy(char *dst, const char *src, size_t n) {
    std::unique_lock<std::mutex> lk(m_);
    dst_ = dst;
    src_ = src;
    n_ = n;
    pending_ = (int)workers_.size();
    ++job_;
    cv_.notify_all();
    done_.wait(lk, [this]() { return pending_ == 0; });
  }
  int size() const { return (int)workers_.size(); }

 private:
  void run(int t, int nt) {
    unsigned long seen = 0;
    for (;;) {
      char *d;
      const char *s;
      size_t n;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&]() { return job_ != seen; });
        seen = job_;
        if (stop_) return;
        d = dst_;
        s = src_;
        n = n_;
      }
      const size_t part = (n + nt - 1) / nt, a = std::min(n, part * t), b = std::min(n, part * (t + 1));
      if (b > a) stream_copy(d + a, s + a, b - a);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  bool stop_;
  unsigned long job_;
  int pending_;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t n_ = 0;
};

// Grow-only device arena with bump allocation: transient set-up buffers (raw CSR uploads, the temporaries of the
// row-pattern build) stay out of the stream-ordered pool, whose free blocks they would otherwise fragment: a
// hierarchy that is handed over again then finds its (large) blocks free instead of growing the pool.
struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0, wanted = 0;
  std::vector<char *> overflow;  // blocks taken while the arena was too small (released at the next reset)
};
}  // namespace gmg

struct gmg_context {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  int64_t launches = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;  // host<->device traffic of the host-pointer entry points

  int n_levels = 0;
  std::vector<gmg::Level> levels;
  gmg::Arena scratch;          // temporaries of build_pat
  gmg::Arena upload[2];        // raw CSR of the system matrix / the level-0 matrix
  gmg::DevCsr rawS;
  gmg::Sell S;  // system matrix
  int n_sys = 0;
  double *s_dinv = nullptr;
  double *g = nullptr, *d = nullptr, *hh = nullptr;      // outer PCG work vectors
  double *cg_g = nullptr, *cg_d = nullptr, *cg_h = nullptr, *cg_dz = nullptr;  // coarse CG work vectors
  int cg_n = 0;
  double *stage_a = nullptr, *stage_b = nullptr;  // device staging for host-pointer entry points
  int64_t stage_n = 0;

  int smoother = GMG_SMOOTHER_JACOBI;
  double omega = 0.5;
  int steps = 2;
  int coarse_max_it = 1000;
  double coarse_tol = 1e-10;
  double drop_tol = -1.0;
  int compress = 2;      // 0: plain SELL; 1: CSELL entries; 2: row-pattern dictionary (falls back to 1, then 0)
  int cg_grid_c = 0;     // cooperative grid of the compressed-format CG kernel
  int cg_grid_p = 0;     // cooperative grid of the row-pattern CG kernel
  bool win_global_codes = false;  // force the large-level variant of the window kernel (tests: GMG_WIN_GLOBAL_CODES=1)
  int cg_win = 2;        // TMA-window variants of the row-pattern CG: 2 = pattern_win2.cuh (default), 1 = pattern_win.cuh,
                         // 0 = L1 gathers (GMG_CG_WIN)
  uint64_t *cg_ll = nullptr;   // tagged words of the in-kernel reductions / barriers of pattern_win2.cuh
  uint32_t cg_ll_tag = 0;
  int cg_ll_stride = 32;       // distance of two blocks' slots in 64-bit words (256 B: the slots spread over the L2 slices)
  bool vc_prof = false;  // gmg_debug_vcycle_profile: events around down sweep / coarse solve / up sweep
  std::vector<cudaEvent_t> vc_ev;
  int vc_ev_used = 0;
  int cg_prof = 0;       // gmg_debug_cg_phases: per-phase timing inside the window kernel
  int cg_win2_smem = 0;
  const void *cg_win2_fn = nullptr;  // the kernel cg_win2_smem was last set on
  int cg_win2_variant = -1;    // GMG_WIN2_VARIANT: -1 = pick (XPAIR where the dominant pattern allows it), 0..3 = force (context.cu)
  int cg_win_smem = 0;   // dynamic shared memory the window kernel is currently configured for
  bool is_setup = false;
  char *pin_small = nullptr;  // 256 bytes of pinned host memory: small D2H results without a staging round trip each
  float *ind_eta = nullptr;  // indicators of the last gmg_error_indicator call (gmg_mark_cells)
  int ind_n = 0;
  float ind_max = 0.0f;

  // reductions
  double *partials = nullptr;     // generic partial buffer
  int partials_cap = 0;
  unsigned int *counter = nullptr;
  gmg::PcgScalars *scalars = nullptr;
  double *cg_partials = nullptr;  // 3 * cg_grid
  int cg_grid = 0;
  gmg::CgResult *cg_results = nullptr;  // ring on device
  int cg_ring = 4096;
  int cg_cursor = 0, cg_solve_begin = 0;
  std::vector<int> last_coarse_its;

  // profiling of the persistent coarse CG
  std::vector<cudaEvent_t> ev_begin, ev_end;
  std::vector<int> ev_result_slot;
  int ev_used = 0;
  double prof_ms = 0.0;
  int64_t prof_launches = 0, prof_iters = 0;

  // RHS path state (rhs.cu)
  double *atom_pos = nullptr, *atom_q = nullptr;
  int n_atoms = 0;
  int64_t *list_ptr = nullptr;
  int *list_atoms = nullptr;
  int n_lists = 0;
  struct RhsState *rhs = nullptr;
  gmg::DistData dist;
  // pinned staging ring for host->device uploads of large arrays (csrc/context.cu: staged_h2d)
  char *pin[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t pin_free[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t pin_bytes = 0;
  int stage_threads = 6;
  std::unique_ptr<gmg::CopyPool> copy_pool;  // created with the ring
  // CUDA graphs of the fine-level parts of the V-cycle (down sweep / up sweep), keyed by (src, dst)
  struct VcGraph {
    const double *src = nullptr;
    double *dst = nullptr;
    cudaGraphExec_t down = nullptr, up = nullptr;
    int64_t n_down = 0, n_up = 0;
  };
  std::vector<VcGraph> vc_graphs;
  bool use_graphs = true;
  bool pdl = true;               // colour sweeps as programmatic dependent launches (GMG_PDL=0 disables)
  bool cluster_ssor = false;     // small levels: one cluster launch per smooth() (GMG_CLUSTER_SSOR=1); measured slower than
                                 // graph-replayed launches (level 2 of the 64k case: +2.8 ms per step): the colours' dependent
                                 // L2 round trips serialise on 8 SMs
  bool persistent_ssor = false;  // measured slower than graph-replayed per-colour launches (60.4 vs 54.0 ms/step)
  int ssor_blocks_per_sm = 0;
};

namespace gmg {
// Stream-ordered allocation from the device's memory pool (release threshold = never): cudaMalloc / cudaFree
// cost milliseconds each and a hierarchy hand-over does hundreds of them.
extern thread_local cudaStream_t tl_stream;
inline void enter(gmg_context *h) {
  cudaSetDevice(h->device);
  tl_stream = h->stream;
}
template <class T>
inline cudaError_t dalloc(T **p, int64_t n) {
  return cudaMallocAsync((void **)p, (size_t)(n > 0 ? n : 1) * sizeof(T), tl_stream);
}
template <class T>
inline void dfree(T *&p) {
  if (p) cudaFreeAsync(p, tl_stream);
  p = nullptr;
}
int fail(gmg_context *h, int code, const std::string &msg);
int ensure_stage(gmg_context *h, int64_t n);
int staged_h2d(gmg_context *h, void *dst, const void *src, size_t bytes);  // large copies through the pinned ring
int staged_d2h(gmg_context *h, void *dst, const void *src, size_t bytes);
void rhs_free(gmg_context *h);
void rhs_invalidate_partition(gmg_context *h);
inline cudaError_t copy(gmg_context *h, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
  if (kind == cudaMemcpyHostToDevice) h->h2d_bytes += (int64_t)bytes;
  if (kind == cudaMemcpyDeviceToHost) h->d2h_bytes += (int64_t)bytes;
  return cudaMemcpyAsync(dst, src, bytes, kind, h->stream);
}
inline cudaError_t copy_sync(gmg_context *h, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t e = copy(h, dst, src, bytes, kind);
  return e != cudaSuccess ? e : cudaStreamSynchronize(h->stream);
}
}  // namespace gmg

// GMG_TRACE=1: wall-clock of the set-up phases on stderr
struct TraceScope {
  const char *name;
  std::chrono::steady_clock::time_point t0;
  bool on;
  explicit TraceScope(const char *n) : name(n), t0(std::chrono::steady_clock::now()), on(std::getenv("GMG_TRACE") != nullptr) {}
  ~TraceScope() {
    if (on) {
      cudaDeviceSynchronize();
      std::fprintf(stderr, "[gmg trace] %-28s %9.3f ms\n", name,
                   1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
  }
};

#define GMG_CUDA(h, expr)                                                                          \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return gmg::fail(h, GMG_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
  } while (0)

#define GMG_LAUNCH_CHECK(h)                                                                        \
  do {                                                                                             \
    (h)->launches++;                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return gmg::fail(h, GMG_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e__));    \
  } while (0)
