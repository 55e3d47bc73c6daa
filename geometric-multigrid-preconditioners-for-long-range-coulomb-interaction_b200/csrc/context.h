// Host-side state behind a gmg_handle (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/gmg_b200.h"
#include "common.cuh"

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
};

struct ColorSet {
  Sell A;          // rows of this colour
  int *rows = nullptr;
  int n = 0;
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
  double lambda_max = 0.0;           // Chebyshev
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
  gmg::DevCsr rawS;
  gmg::Sell S;  // system matrix
  int n_sys = 0;
  double *s_dinv = nullptr;
  double *g = nullptr, *d = nullptr, *hh = nullptr;      // outer PCG work vectors
  double *cg_g = nullptr, *cg_d = nullptr, *cg_h = nullptr;  // coarse CG work vectors
  int cg_n = 0;
  double *stage_a = nullptr, *stage_b = nullptr;  // device staging for host-pointer entry points
  int64_t stage_n = 0;

  int smoother = GMG_SMOOTHER_JACOBI;
  double omega = 0.5;
  int steps = 2;
  int coarse_max_it = 1000;
  double coarse_tol = 1e-10;
  double drop_tol = -1.0;
  bool is_setup = false;

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
void rhs_free(gmg_context *h);
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
