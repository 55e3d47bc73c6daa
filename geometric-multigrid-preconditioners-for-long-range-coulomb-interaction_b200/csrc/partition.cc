// see partition.h
#include "partition.h"

#include <algorithm>
#include <stdexcept>

namespace gmg {

void partition_matrix(int rank, int world, int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col,
                      const double *val, const int32_t *row_owner, const int32_t *col_owner, LocalMatrix &out,
                      ExchangePlan &plan) {
  if (world > 32) throw std::invalid_argument("partition_matrix: at most 32 ranks");
  // which ranks reference each column from one of their own rows
  std::vector<uint32_t> wanted(n_cols, 0u);
  for (int32_t i = 0; i < n_rows; ++i) {
    const uint32_t bit = 1u << row_owner[i];
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) wanted[col[k]] |= bit;
  }
  plan.rank = rank;
  plan.world = world;
  plan.n_owned_of.assign(world, 0);
  plan.n_halo_of.assign(world, 0);
  for (int32_t j = 0; j < n_cols; ++j) plan.n_owned_of[col_owner[j]]++;
  // halo(q, o): columns owned by o wanted by q != o, ascending global index; sizes for every pair
  std::vector<std::vector<int32_t>> halo_cnt(world, std::vector<int32_t>(world, 0));
  for (int32_t j = 0; j < n_cols; ++j) {
    const int o = col_owner[j];
    uint32_t w = wanted[j] & ~(1u << o);
    while (w) {
      const int q = __builtin_ctz(w);
      w &= w - 1;
      halo_cnt[q][o]++;
    }
  }
  for (int q = 0; q < world; ++q)
    for (int o = 0; o < world; ++o) plan.n_halo_of[q] += halo_cnt[q][o];
  // my owned and halo lists; global -> local map
  std::vector<int32_t> g2l(n_cols, -1);
  out.owned_global.clear();
  for (int32_t j = 0; j < n_cols; ++j)
    if (col_owner[j] == rank) {
      g2l[j] = (int32_t)out.owned_global.size();
      out.owned_global.push_back(j);
    }
  out.n_owned = (int)out.owned_global.size();
  std::vector<std::vector<int32_t>> my_halo(world);
  for (int32_t j = 0; j < n_cols; ++j)
    if (col_owner[j] != rank && (wanted[j] >> rank & 1u)) my_halo[col_owner[j]].push_back(j);
  out.halo_global.clear();
  out.halo_owner.clear();
  plan.recv_count.assign(world, 0);
  for (int o = 0; o < world; ++o) {
    plan.recv_count[o] = (int32_t)my_halo[o].size();
    for (int32_t j : my_halo[o]) {
      g2l[j] = out.n_owned + (int32_t)out.halo_global.size();
      out.halo_global.push_back(j);
      out.halo_owner.push_back(o);
    }
  }
  out.n_halo = (int)out.halo_global.size();
  out.n_halo_lo = 0;
  for (int o = 0; o < rank; ++o) out.n_halo_lo += plan.recv_count[o];
  // local column of a halo entry: relative to the first owned entry
  for (int k = 0; k < out.n_halo; ++k)
    g2l[out.halo_global[k]] = (k < out.n_halo_lo) ? k - out.n_halo_lo : out.n_owned + (k - out.n_halo_lo);
  plan.n_halo_lo_of.assign(world, 0);
  for (int q = 0; q < world; ++q)
    for (int o = 0; o < q; ++o) plan.n_halo_lo_of[q] += halo_cnt[q][o];
  // what I send to q: q's halo segment for owner == rank, in ascending global order
  plan.send_idx.assign(world, {});
  plan.send_dst_base.assign(world, 0);
  plan.send_hpos_base.assign(world, 0);
  for (int q = 0; q < world; ++q) {
    if (q == rank) continue;
    int32_t off = 0;
    for (int o = 0; o < rank; ++o) off += halo_cnt[q][o];
    plan.send_hpos_base[q] = off;
    plan.send_dst_base[q] = off + (rank > q ? plan.n_owned_of[q] : 0);
  }
  for (int32_t j = 0; j < n_cols; ++j) {
    if (col_owner[j] != rank) continue;
    uint32_t w = wanted[j] & ~(1u << rank);
    while (w) {
      const int q = __builtin_ctz(w);
      w &= w - 1;
      plan.send_idx[q].push_back(g2l[j]);
    }
  }
  // local rows (requires the square case row i <-> col i when used for SpMV on owned vectors)
  out.rowptr.assign(1, 0);
  out.col.clear();
  out.val.clear();
  for (int32_t i = 0; i < n_rows; ++i) {
    if (row_owner[i] != rank) continue;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      out.col.push_back(g2l[col[k]]);
      out.val.push_back(val[k]);
    }
    out.rowptr.push_back((int64_t)out.col.size());
  }
}

void extract_owned_rows(int rank, int32_t n_rows, const int64_t *rowptr, const int32_t *col, const double *val,
                        const int32_t *row_owner, std::vector<int64_t> &l_rowptr, std::vector<int32_t> &l_col,
                        std::vector<double> &l_val) {
  l_rowptr.assign(1, 0);
  l_col.clear();
  l_val.clear();
  for (int32_t i = 0; i < n_rows; ++i) {
    if (row_owner[i] != rank) continue;
    l_col.insert(l_col.end(), col + rowptr[i], col + rowptr[i + 1]);
    l_val.insert(l_val.end(), val + rowptr[i], val + rowptr[i + 1]);
    l_rowptr.push_back((int64_t)l_col.size());
  }
}

}  // namespace gmg
