// Host-side row-block partitioning of a global CSR matrix for the multi-GPU path (no CUDA calls here,
// so the N > 1 host logic is testable on the CPU).  Mirrors what Epetra's Import/Export objects set up
// for the reference's distributed matrices: owned rows, ghost (halo) columns, send lists.
//
// Halo entries are ordered by (owner rank, global index); the extended vector of rank r is laid out as
//     [ halo entries owned by lower ranks | owned entries (ascending global order) | halo entries of higher ranks ]
// and local column indices are relative to the first OWNED entry (so lower-rank halo columns are negative).
// With slab-like partitions every row then reaches its neighbours at small column offsets, which keeps the
// 16-bit offsets of the compressed matrix format valid.  What rank r sends to rank q is exactly q's halo
// segment for owner r, in q's halo order.
#pragma once
#include <cstdint>
#include <vector>

namespace gmg {

struct LocalMatrix {
  int n_owned = 0, n_halo = 0, n_halo_lo = 0;  // n_halo_lo: halo entries owned by lower ranks
  std::vector<int64_t> rowptr;      // n_owned + 1
  std::vector<int32_t> col;         // local column indices in [-n_halo_lo, n_owned + n_halo - n_halo_lo)
  std::vector<double> val;
  std::vector<int32_t> owned_global;  // global row/col index of each owned entry
  std::vector<int32_t> halo_global;   // global index of each halo entry
  std::vector<int32_t> halo_owner;
};

struct ExchangePlan {
  int rank = 0, world = 1;
  // per peer q: local (owned) indices to send, and where they land in q's extended vector
  std::vector<std::vector<int32_t>> send_idx;   // [world][..]
  std::vector<int32_t> send_dst_base;           // [world]: index in q's extended vector where my segment starts
  std::vector<int32_t> send_hpos_base;          // [world]: position of my segment in q's halo list
  std::vector<int32_t> n_halo_lo_of;            // [world]
  std::vector<int32_t> recv_count;              // [world]: entries received from each peer
  std::vector<int32_t> n_owned_of, n_halo_of;   // [world]
};

// col_owner may equal row_owner (square matrices).  Every rank calls this with the same global data.
void partition_matrix(int rank, int world, int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col,
                      const double *val, const int32_t *row_owner, const int32_t *col_owner, LocalMatrix &out,
                      ExchangePlan &plan);

// rows owned by `rank`, all columns kept global (operators whose input vector is replicated)
void extract_owned_rows(int rank, int32_t n_rows, const int64_t *rowptr, const int32_t *col, const double *val,
                        const int32_t *row_owner, std::vector<int64_t> &l_rowptr, std::vector<int32_t> &l_col,
                        std::vector<double> &l_val);

}  // namespace gmg
