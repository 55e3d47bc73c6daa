// Device-side assembly of the Q1 Laplace matrices (gmg_assemble_matrix, include/gmg_b200.h): the host
// hands over the cell -> dof map, the cell sizes, one flag byte per row and the hanging-node constraint lines (a few
// hundred MB less than the assembled CSR: 12 bytes per stored entry stay at home), the device builds the CSR the host
// would have built -- same pattern (explicit zeros included), same value bits (assemble_row.h: sums in the order of the
// sequential cell loop, unfused arithmetic).  Included by context.cu.
//
//   1. asm_slot_count / asm_slot_emit: every (cell, vertex) slot emits its incidence entries (the dof itself; the free
//      parents of a hanging dof); offsets by an exclusive scan so the entries come out in slot order;
//   2. stable radix sort by row (CUB) -> per-row incidence lists in (cell, vertex, parent) order = the encounter order
//      of the reference's cell loop;
//   3. asm_rows_count: one thread per row collects its sorted column set in local memory (64 columns; the few rows next
//      to hanging nodes that need more are redone with 320); exclusive scan -> row pointer;
//   4. asm_rows_fill: the same column set again + the values, written to the CSR arrays.
#include "assemble_row.h"

// the cell arrays kept on the device by the last gmg_assemble_rhs call (rhs.cu)
int gmg_rhs_resident(gmg_context *h, int *n_cells, const double **cell_h, const int **cell_dofs, const double **weights, int *n_q,
                     const double **rho, int *rho_cells, int *rho_nq);

namespace gmg {

// ncu (profiles/r02_ncu_summary.md): system matrix at 64k atoms -- count<64> 2.35 ms / 872 M warp instructions, fill<64>
// 5.93 ms / 1794 M, against 0.35 ms / 215 M and 1.44 ms / 544 M for the hanging-node-free level-0 matrix of the same size:
// three quarters of the instructions are issued by the few warps that hold a row next to a hanging node (a coarse vertex
// that is a parent of many hanging dofs walks 50-100 incidence entries, each with 8 dofs and their constraint lines).
// Measured and dropped in round 2: smaller / shared-memory column arrays (32 / 96 / 320 tiers; [32][128] in shared memory),
// the column set cached between the two passes, batched flag loads -- none moved the two kernels by more than 5 %.
constexpr int ASM_SMALL = 64;
constexpr int ASM_LARGE = 320;

__global__ void asm_slot_count(int64_t n_slots, int n_rows, const int32_t *cell_dofs, const uint8_t *flags,
                               const int64_t *hang_ptr, const int32_t *hang_col, unsigned long long *slot_cnt,
                               unsigned long long *row_cnt, int *err) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const int d = cell_dofs[s];
  if (d < 0 || d >= n_rows) {
    *err = 1;
    if (slot_cnt) slot_cnt[s] = 0;
    return;
  }
  if (hang_ptr && (flags[d] & ASM_HANGING)) {
    const int64_t len = hang_ptr[d + 1] - hang_ptr[d];
    bool ok = len >= 0 && len <= ASM_MAX_PARENTS;
    for (int64_t e = hang_ptr[d]; ok && e < hang_ptr[d + 1]; ++e) ok = hang_col[e] >= 0 && hang_col[e] < n_rows;
    if (!ok) {
      *err = 2;
      if (slot_cnt) slot_cnt[s] = 0;
      return;
    }
  }
  const int n = asm_slot_entries(cell_dofs, flags, hang_ptr, hang_col, s,
                                 [&](int row, uint64_t) { atomicAdd(&row_cnt[row], 1ull); });
  if (slot_cnt) slot_cnt[s] = (unsigned long long)n;
}

__global__ void asm_slot_emit(int64_t n_slots, int n_rows, const int32_t *cell_dofs, const uint8_t *flags,
                              const int64_t *hang_ptr, const int32_t *hang_col, const unsigned long long *slot_off,
                              unsigned *key, unsigned long long *ent) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_slots) return;
  const int d = cell_dofs[s];
  if (d < 0 || d >= n_rows) return;
  int64_t at = slot_off ? (int64_t)slot_off[s] : s;
  asm_slot_entries(cell_dofs, flags, hang_ptr, hang_col, s, [&](int row, uint64_t e) {
    key[at] = (unsigned)row;
    ent[at] = e;
    ++at;
  });
}

__global__ void asm_cell_hanging(int64_t n_cells, const int32_t *cell_dofs, const uint8_t *flags, uint8_t *cell_hang) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  int any = 0;
  for (int a = 0; a < 8; ++a) any |= flags[cell_dofs[8 * c + a]] & ASM_HANGING;
  cell_hang[c] = (uint8_t)(any != 0);
}

// Which row a thread takes.  With hanging nodes the rows are visited in the order of DECREASING incidence-list length
// (stable: the regular rows, all with 8 incident cells, keep their index order and their coalesced accesses): a free
// vertex that is a parent of many hanging dofs walks 50-100 incidence entries, and with one such row among 31 regular
// ones three quarters of the warp instructions of the row kernels were issued with one active lane.  Measured (B200, 64k
// atoms, system matrix): row widths 2.29 -> 2.17 ms, columns + values 5.67 -> 5.06 ms -- the long rows now share warps and
// start first, but what remains is the sequential walk of the longest rows themselves (one thread each, every add in the
// order of the cell loop); shortening it needs a warp per long row with the adds into a column kept in lane order.
__global__ void asm_row_keys(int n_rows, const unsigned long long *row_cnt, unsigned *key, unsigned *idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const unsigned long long c = row_cnt[i];
  key[i] = 255u - (unsigned)(c < 255ull ? c : 255ull);
  idx[i] = (unsigned)i;
}

template <int MAXC, bool LARGE>
__global__ void __launch_bounds__(128) asm_rows_count(AsmView A, const unsigned *__restrict__ order, uint8_t *large,
                                                      unsigned long long *cnt, int *err) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= A.n_rows) return;
  const int i = order ? (int)order[tid] : tid;
  if (LARGE && !large[i]) return;
  int cols[MAXC];
  const int n = asm_row_pattern(A, i, cols, MAXC);
  if (n < 0) {
    if (LARGE) *err = 3;
    else large[i] = 1;
    cnt[i] = 0;
    return;
  }
  if (!LARGE) large[i] = 0;
  cnt[i] = (unsigned long long)n;
}

template <int MAXC, bool LARGE>
__global__ void __launch_bounds__(128) asm_rows_fill(AsmView A, const unsigned *__restrict__ order, const uint8_t *large,
                                                     const int64_t *rowptr, int *col, double *val) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= A.n_rows) return;
  const int i = order ? (int)order[tid] : tid;
  if (LARGE != (large[i] != 0)) return;
  int cols[MAXC];
  double vals[MAXC];
  const int n = asm_row_pattern(A, i, cols, MAXC);
  if (n < 0) return;  // (cannot happen: the count pass has sized the row)
  asm_row_values(A, i, cols, n, vals);
  const int64_t at = rowptr[i];
  for (int k = 0; k < n; ++k) {
    col[at + k] = cols[k];
    val[at + k] = vals[k];
  }
}

// n_cells cells; cell_h null: uniform_h.  The result lands in `out` (arrays in `arena`, like an uploaded CSR).
static int assemble_matrix_device(gmg_context *h, int n_rows, int64_t n_cells, const int32_t *cell_dofs,
                                  const double *cell_h, double uniform_h, const uint8_t *row_flags,
                                  const int64_t *hang_rowptr, const int32_t *hang_col, const double *hang_val,
                                  const double *k_ref, DevCsr &out, Arena &arena) {
  TraceScope tr("gmg_assemble_matrix");
  free_csr(out);
  arena_reset(arena);
  const int64_t n_slots = 8 * n_cells;
  const int64_t n_hang = hang_rowptr ? hang_rowptr[n_rows] : 0;
  const bool hanging = n_hang > 0;
  AsmView A{};
  A.n_rows = n_rows;
  A.n_cells = n_cells;
  A.uniform_h = uniform_h;
  std::memcpy(A.kref, k_ref, sizeof(A.kref));
  // ---- inputs
  std::unique_ptr<TraceScope> ts(new TraceScope("  asm: inputs H2D"));
  int32_t *d_dofs = nullptr, *d_hcol = nullptr;
  double *d_h = nullptr, *d_hval = nullptr;
  uint8_t *d_flags = nullptr, *d_cell_hang = nullptr, *d_large = nullptr;
  int64_t *d_hptr = nullptr;
  int *d_err = nullptr;
  GMG_CUDA(h, arena_alloc(arena, &d_flags, n_rows));
  GMG_CUDA(h, arena_alloc(arena, &d_err, 1));
  GMG_CUDA(h, cudaMemsetAsync(d_err, 0, sizeof(int), h->stream));
  if (cell_dofs) {
    GMG_CUDA(h, arena_alloc(arena, &d_dofs, n_slots));
    if (int rc = staged_h2d(h, d_dofs, cell_dofs, sizeof(int32_t) * n_slots)) return rc;
    if (cell_h) {
      GMG_CUDA(h, arena_alloc(arena, &d_h, n_cells));
      if (int rc = staged_h2d(h, d_h, cell_h, sizeof(double) * n_cells)) return rc;
    }
  } else {
    // the cells of the last gmg_assemble_rhs call are still on the device (the same active cells, the same order)
    int r_cells = 0, n_q = 0, rho_cells = 0, rho_nq = 0;
    const double *r_h = nullptr, *weights = nullptr, *rho_dev = nullptr;
    const int *r_dofs = nullptr;
    if (int rc = gmg_rhs_resident(h, &r_cells, &r_h, &r_dofs, &weights, &n_q, &rho_dev, &rho_cells, &rho_nq)) return rc;
    if (r_cells != n_cells || !r_dofs || !r_h)
      return fail(h, GMG_EINVAL, "gmg_assemble_matrix: cell_dofs == NULL needs the cells of a gmg_assemble_rhs call on the same mesh");
    d_dofs = const_cast<int32_t *>(r_dofs);
    d_h = const_cast<double *>(r_h);
  }
  if (int rc = staged_h2d(h, d_flags, row_flags, sizeof(uint8_t) * n_rows)) return rc;
  if (hanging) {
    GMG_CUDA(h, arena_alloc(arena, &d_hptr, (int64_t)n_rows + 1));
    GMG_CUDA(h, arena_alloc(arena, &d_hcol, n_hang));
    GMG_CUDA(h, arena_alloc(arena, &d_hval, n_hang));
    GMG_CUDA(h, arena_alloc(arena, &d_cell_hang, n_cells));
    if (int rc = staged_h2d(h, d_hptr, hang_rowptr, sizeof(int64_t) * ((size_t)n_rows + 1))) return rc;
    if (int rc = staged_h2d(h, d_hcol, hang_col, sizeof(int32_t) * n_hang)) return rc;
    if (int rc = staged_h2d(h, d_hval, hang_val, sizeof(double) * n_hang)) return rc;
  }
  A.cell_dofs = d_dofs;
  A.cell_h = d_h;
  A.flags = d_flags;
  A.hang_ptr = d_hptr;
  A.hang_col = d_hcol;
  A.hang_val = d_hval;
  // ---- incidence lists
  ts.reset();
  ts.reset(new TraceScope("  asm: incidence lists"));
  unsigned long long *slot_cnt = nullptr, *slot_off = nullptr, *row_cnt = nullptr, *ent = nullptr, *ent_sorted = nullptr;
  unsigned *key = nullptr, *key_sorted = nullptr;
  int64_t *inc_ptr = nullptr;
  GMG_CUDA(h, arena_alloc(arena, &row_cnt, (int64_t)n_rows + 1));
  GMG_CUDA(h, arena_alloc(arena, &inc_ptr, (int64_t)n_rows + 1));
  GMG_CUDA(h, cudaMemsetAsync(row_cnt, 0, sizeof(unsigned long long) * ((size_t)n_rows + 1), h->stream));
  if (hanging) {
    GMG_CUDA(h, arena_alloc(arena, &slot_cnt, n_slots + 1));
    GMG_CUDA(h, arena_alloc(arena, &slot_off, n_slots + 1));
    GMG_CUDA(h, cudaMemsetAsync(slot_cnt + n_slots, 0, sizeof(unsigned long long), h->stream));
  }
  if (n_slots > 0) {
    asm_slot_count<<<cdiv(n_slots, 256), 256, 0, h->stream>>>(n_slots, n_rows, d_dofs, d_flags, d_hptr, d_hcol, slot_cnt,
                                                              row_cnt, d_err);
    GMG_LAUNCH_CHECK(h);
  }
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  int64_t n_ent = n_slots;
  if (hanging) {
    GMG_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, slot_cnt, slot_off, n_slots + 1, h->stream));
    GMG_CUDA(h, arena_alloc(arena, (char **)&tmp, (int64_t)tmp_bytes));
    GMG_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, slot_cnt, slot_off, n_slots + 1, h->stream));
    unsigned long long total = 0;
    GMG_CUDA(h, copy_sync(h, &total, slot_off + n_slots, sizeof(total), cudaMemcpyDeviceToHost));
    n_ent = (int64_t)total;
  }
  int err = 0;
  GMG_CUDA(h, copy_sync(h, &err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err == 1) return fail(h, GMG_EINVAL, "gmg_assemble_matrix: a cell dof is outside [0, n_rows)");
  if (err == 2) return fail(h, GMG_EINVAL, "gmg_assemble_matrix: a constraint line has more than 7 entries or a bad column");
  GMG_CUDA(h, arena_alloc(arena, &key, n_ent));
  GMG_CUDA(h, arena_alloc(arena, &key_sorted, n_ent));
  GMG_CUDA(h, arena_alloc(arena, &ent, n_ent));
  GMG_CUDA(h, arena_alloc(arena, &ent_sorted, n_ent));
  if (n_slots > 0) {
    asm_slot_emit<<<cdiv(n_slots, 256), 256, 0, h->stream>>>(n_slots, n_rows, d_dofs, d_flags, d_hptr, d_hcol, slot_off, key,
                                                             ent);
    GMG_LAUNCH_CHECK(h);
  }
  int end_bit = 1;
  while (end_bit < 32 && ((unsigned)std::max(n_rows, 1) >> end_bit) != 0u) ++end_bit;
  size_t sort_bytes = 0, scan_bytes = 0, order_bytes = 0;
  unsigned *okey = nullptr, *okey_sorted = nullptr, *oidx = nullptr, *order = nullptr;
  if (hanging && n_rows > 0 && !std::getenv("GMG_ASM_UNSORTED")) {
    GMG_CUDA(h, arena_alloc(arena, &okey, n_rows));
    GMG_CUDA(h, arena_alloc(arena, &okey_sorted, n_rows));
    GMG_CUDA(h, arena_alloc(arena, &oidx, n_rows));
    GMG_CUDA(h, arena_alloc(arena, &order, n_rows));
    GMG_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, order_bytes, okey, okey_sorted, oidx, order, n_rows, 0, 8, h->stream));
  }
  GMG_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, key, key_sorted, ent, ent_sorted, n_ent, 0, end_bit,
                                              h->stream));
  GMG_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, row_cnt, (unsigned long long *)inc_ptr, n_rows + 1,
                                            h->stream));
  tmp_bytes = std::max(std::max(sort_bytes, scan_bytes), order_bytes);
  GMG_CUDA(h, arena_alloc(arena, (char **)&tmp, (int64_t)tmp_bytes));
  GMG_CUDA(h, cub::DeviceRadixSort::SortPairs(tmp, sort_bytes, key, key_sorted, ent, ent_sorted, n_ent, 0, end_bit,
                                              h->stream));
  GMG_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, row_cnt, (unsigned long long *)inc_ptr, n_rows + 1, h->stream));
  if (order) {  // (row_cnt still holds the lengths of the incidence lists; it is reused as the row widths below)
    asm_row_keys<<<cdiv(n_rows, 256), 256, 0, h->stream>>>(n_rows, row_cnt, okey, oidx);
    GMG_LAUNCH_CHECK(h);
    GMG_CUDA(h, cub::DeviceRadixSort::SortPairs(tmp, order_bytes, okey, okey_sorted, oidx, order, n_rows, 0, 8, h->stream));
  }
  A.inc_ptr = inc_ptr;
  A.inc = (const uint64_t *)ent_sorted;
  if (hanging && n_cells > 0) {
    asm_cell_hanging<<<cdiv(n_cells, 256), 256, 0, h->stream>>>(n_cells, d_dofs, d_flags, d_cell_hang);
    GMG_LAUNCH_CHECK(h);
    A.cell_hang = d_cell_hang;
  }
  // ---- row pointer
  ts.reset();
  ts.reset(new TraceScope("  asm: row widths"));
  out.n_rows = out.n_cols = n_rows;
  out.in_arena = true;
  GMG_CUDA(h, arena_alloc(arena, &out.rowptr, (int64_t)n_rows + 1));
  GMG_CUDA(h, arena_alloc(arena, &d_large, n_rows));
  GMG_CUDA(h, cudaMemsetAsync(row_cnt + n_rows, 0, sizeof(unsigned long long), h->stream));  // (reused as the row widths)
  if (n_rows > 0) {
    asm_rows_count<ASM_SMALL, false><<<cdiv(n_rows, 128), 128, 0, h->stream>>>(A, order, d_large, row_cnt, d_err);
    GMG_LAUNCH_CHECK(h);
    asm_rows_count<ASM_LARGE, true><<<cdiv(n_rows, 128), 128, 0, h->stream>>>(A, order, d_large, row_cnt, d_err);  // flagged rows
    GMG_LAUNCH_CHECK(h);
  }
  GMG_CUDA(h, cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, row_cnt, (unsigned long long *)out.rowptr, n_rows + 1,
                                            h->stream));
  GMG_CUDA(h, copy(h, &out.nnz, out.rowptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost));
  GMG_CUDA(h, copy_sync(h, &err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err == 3) {
    out = DevCsr{};  // (arena memory: nothing to free)
    return fail(h, GMG_EINVAL, "gmg_assemble_matrix: a row has more than 320 columns");
  }
  // ---- columns and values
  ts.reset();
  ts.reset(new TraceScope("  asm: columns + values"));
  GMG_CUDA(h, arena_alloc(arena, &out.col, out.nnz));
  GMG_CUDA(h, arena_alloc(arena, &out.val, out.nnz));
  if (n_rows > 0) {
    asm_rows_fill<ASM_SMALL, false><<<cdiv(n_rows, 128), 128, 0, h->stream>>>(A, order, d_large, out.rowptr, out.col, out.val);
    GMG_LAUNCH_CHECK(h);
    asm_rows_fill<ASM_LARGE, true><<<cdiv(n_rows, 128), 128, 0, h->stream>>>(A, order, d_large, out.rowptr, out.col, out.val);
    GMG_LAUNCH_CHECK(h);
  }
  // no synchronisation here: the borrowed host buffers were consumed by the staged copies above, and the fill kernels
  // overlap with the caller's next hand-over call (the host-side staging of the next matrix); gmg_setup orders after them
  ts.reset();
  return GMG_OK;
}

}  // namespace gmg
