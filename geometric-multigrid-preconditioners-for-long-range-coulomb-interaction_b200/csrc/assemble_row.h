// Row-gather assembly of the Q1 Laplace matrices (system matrix with hanging-node / Dirichlet constraints condensed,
// level matrices with boundary and refinement-edge dofs eliminated): the arithmetic of one matrix row, shared by the
// CUDA kernels (assemble.inl) and by a sequential host emulation (host/capi_host.cc: ms_assemble_emulate) that the CPU
// tests compare with the host assembly entry by entry.
//
// Replaces, row by row, what the reference does cell by cell (src/step-50.cc:771-795 with
// constraints.distribute_local_to_global and make_sparsity_pattern(dof, dsp, constraints, keep_constrained = true),
// :699-701; level matrices :855-889):
//   * pattern of row i = the dofs of every cell containing i (explicit zeros kept) + the columns its condensed
//     contributions land in;
//   * constrained row (Dirichlet, hanging, level boundary / refinement edge): diagonal = sum_c |K^c_aa|, rest zero;
//   * free row: sum over the cells c and local rows a it receives a share w_a of (itself: w = 1; hanging dofs it is a
//     parent of: their weight), of w_a * w_b * K^c_ab into the column dof b resolves to (itself; its free parents;
//     nothing when eliminated), added in the order (cell, a, parent of a, b, parent of b) -- the order of the
//     sequential cell loop, so the sums carry the same bits.
// K^c = h_c * K_ref (K_ref: unit-cube cell matrix, handed over by the host so its bits are the host's).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GMG_HD __host__ __device__ __forceinline__
#else
#define GMG_HD inline
#endif

namespace gmg {

constexpr int ASM_ELIMINATED = 1;  // Dirichlet dof / level boundary or refinement-edge dof: row and column eliminated
constexpr int ASM_HANGING = 2;     // hanging dof: row and column condensed into the free parents
constexpr int ASM_MAX_PARENTS = 7;

// incidence entry of a row: (slot << 3) | t with slot = 8 * cell + local index a; t = 0: the row's dof IS vertex a of
// the cell, t >= 1: vertex a is a hanging dof and the row is its parent number t - 1 (position in its constraint line)
GMG_HD uint64_t asm_entry(int64_t slot, int t) { return ((uint64_t)slot << 3) | (uint64_t)t; }

struct AsmView {
  int n_rows;
  int64_t n_cells;
  const int32_t *cell_dofs;  // [n_cells][8]
  const double *cell_h;      // [n_cells], or null: uniform_h
  double uniform_h;
  const uint8_t *flags;      // [n_rows] ASM_* bits
  const uint8_t *cell_hang;  // [n_cells] 1: the cell has a hanging vertex (null: look at the vertices)
  const int64_t *hang_ptr;   // constraint lines as CSR over the rows (null: no hanging dofs)
  const int32_t *hang_col;
  const double *hang_val;
  const int64_t *inc_ptr;    // [n_rows + 1] incidence lists, entries in (cell, a, t) order
  const uint64_t *inc;
  double kref[64];
};

GMG_HD double asm_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);  // never contracted into an FMA with the running sum
#else
  return a * b;
#endif
}
GMG_HD double asm_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}

// sorted insert without duplicates; false: the row does not fit into max_n columns
GMG_HD bool asm_insert(int *cols, int &n, int max_n, int c) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int m = (lo + hi) >> 1;
    if (cols[m] < c) lo = m + 1;
    else hi = m;
  }
  if (lo < n && cols[lo] == c) return true;
  if (n == max_n) return false;
  for (int k = n; k > lo; --k) cols[k] = cols[k - 1];
  cols[lo] = c;
  ++n;
  return true;
}

GMG_HD int asm_find(const int *cols, int n, int c) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int m = (lo + hi) >> 1;
    if (cols[m] < c) lo = m + 1;
    else hi = m;
  }
  return lo;
}

// columns of row i, ascending, into cols[max_n]; returns their number or -1 when they do not fit
GMG_HD int asm_row_pattern(const AsmView &A, int i, int *cols, int max_n) {
  int n = 0;
  const bool free_row = (A.flags[i] & (ASM_ELIMINATED | ASM_HANGING)) == 0;
  for (int64_t e = A.inc_ptr[i]; e < A.inc_ptr[i + 1]; ++e) {
    const uint64_t w = A.inc[e];
    const int t = (int)(w & 7u);
    const int64_t cell = (int64_t)(w >> 6);
    const int32_t *dofs = A.cell_dofs + 8 * cell;
    if (t == 0)
      for (int b = 0; b < 8; ++b)
        if (!asm_insert(cols, n, max_n, dofs[b])) return -1;
    if (free_row && A.hang_ptr && (t != 0 || !A.cell_hang || A.cell_hang[cell]))
      for (int b = 0; b < 8; ++b) {
        const int d = dofs[b];
        const int f = A.flags[d];
        if (f & ASM_HANGING) {
          for (int64_t e2 = A.hang_ptr[d]; e2 < A.hang_ptr[d + 1]; ++e2) {
            const int p = A.hang_col[e2];
            if ((A.flags[p] & (ASM_ELIMINATED | ASM_HANGING)) == 0)
              if (!asm_insert(cols, n, max_n, p)) return -1;
          }
        } else if (t != 0 && !(f & ASM_ELIMINATED)) {
          if (!asm_insert(cols, n, max_n, d)) return -1;
        }
      }
  }
  return n;
}

// values of row i for the columns found by asm_row_pattern
GMG_HD void asm_row_values(const AsmView &A, int i, const int *cols, int n, double *vals) {
  for (int k = 0; k < n; ++k) vals[k] = 0.0;
  const bool free_row = (A.flags[i] & (ASM_ELIMINATED | ASM_HANGING)) == 0;
  double diag = 0.0;
  for (int64_t e = A.inc_ptr[i]; e < A.inc_ptr[i + 1]; ++e) {
    const uint64_t w = A.inc[e];
    const int t = (int)(w & 7u);
    const int a = (int)((w >> 3) & 7u);
    const int64_t cell = (int64_t)(w >> 6);
    const int32_t *dofs = A.cell_dofs + 8 * cell;
    const double hc = A.cell_h ? A.cell_h[cell] : A.uniform_h;
    if (!free_row) {
      const double k = asm_mul(hc, A.kref[a * 8 + a]);
      diag = asm_add(diag, k < 0.0 ? -k : k);
      continue;
    }
    const double wa = t == 0 ? 1.0 : A.hang_val[A.hang_ptr[dofs[a]] + (t - 1)];
    for (int b = 0; b < 8; ++b) {
      const int d = dofs[b];
      const int f = A.flags[d];
      const double kab = asm_mul(hc, A.kref[a * 8 + b]);
      if (f & ASM_HANGING) {
        for (int64_t e2 = A.hang_ptr[d]; e2 < A.hang_ptr[d + 1]; ++e2) {
          const int p = A.hang_col[e2];
          if ((A.flags[p] & (ASM_ELIMINATED | ASM_HANGING)) == 0) {
            const int k = asm_find(cols, n, p);
            vals[k] = asm_add(vals[k], asm_mul(asm_mul(wa, A.hang_val[e2]), kab));
          }
        }
      } else if (!(f & ASM_ELIMINATED)) {
        const int k = asm_find(cols, n, d);
        vals[k] = asm_add(vals[k], asm_mul(wa, kab));  // (w_a * 1) * K_ab
      }
    }
  }
  if (!free_row) vals[asm_find(cols, n, i)] = diag;
}

// incidence entries generated by vertex `a` of `cell` (slot = 8 * cell + a), in order: the dof itself, then the free
// parents of a hanging dof.  emit(row, entry) is called for each; returns their number.
template <class Emit>
GMG_HD int asm_slot_entries(const int32_t *cell_dofs, const uint8_t *flags, const int64_t *hang_ptr, const int32_t *hang_col,
                            int64_t slot, Emit emit) {
  const int d = cell_dofs[slot];
  int n = 1;
  emit(d, asm_entry(slot, 0));
  if (hang_ptr && (flags[d] & ASM_HANGING)) {
    int t = 1;
    for (int64_t e = hang_ptr[d]; e < hang_ptr[d + 1]; ++e, ++t) {
      const int p = hang_col[e];
      if ((flags[p] & (ASM_ELIMINATED | ASM_HANGING)) == 0) {
        emit(p, asm_entry(slot, t));
        ++n;
      }
    }
  }
  return n;
}

}  // namespace gmg
