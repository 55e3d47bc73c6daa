// extern "C" shim over ministep for the Python test harness and the CPU arm of bench.py: forest, numbering,
// sparsity, host assembly, indicator, solution transfer.  Pure host code: this file and ministep.cc also build into
// libministep_b200.so, which links no CUDA library (the CPU baseline must not map the product's kernels).
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <sstream>

#include "capi_ministep.h"
#include "../csrc/assemble_row.h"

using namespace ministep;

namespace {
struct Bundle {
  std::unique_ptr<Forest> forest;
  std::unique_ptr<DoFs> dofs;
  Csr system;
  LevelOperators ops;
  std::vector<std::vector<float>> eta;
  std::vector<std::vector<char>> flags;
  double threshold = 0.0;
  IndicatorTopology topo;
  std::vector<int32_t> scratch_i32;
  std::vector<double> scratch_f64;
  std::string err;
};
}  // namespace
namespace ministep_capi {
thread_local std::string g_err;
}
using ministep_capi::g_err;

extern "C" {

const char *ms_last_error() { return g_err.c_str(); }

void *ms_create(int reps, double lo, double hi) {
  Bundle *b = new Bundle();
  b->forest.reset(new Forest(reps, lo, hi));
  return b;
}
void ms_destroy(void *p) { delete (Bundle *)p; }

int ms_refine_global(void *p, int times) {
  try {
    ((Bundle *)p)->forest->refine_global(times);
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

// flags: concatenated per level (n_cells(l) bytes each)
int ms_refine(void *p, const uint8_t *flags) {
  Bundle *b = (Bundle *)p;
  try {
    std::vector<std::vector<char>> fl(b->forest->n_levels());
    size_t off = 0;
    for (int l = 0; l < b->forest->n_levels(); ++l) {
      fl[l].assign(flags + off, flags + off + b->forest->n_cells(l));
      off += b->forest->n_cells(l);
    }
    b->forest->refine(fl);
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

int ms_build(void *p, int step16_coefficient, int with_matrices) {
  Bundle *b = (Bundle *)p;
  try {
    b->dofs.reset(new DoFs(*b->forest));
    if (with_matrices) {
      Coefficient coef;
      if (step16_coefficient) coef = [](double x, double y, double z) { return (x * x + y * y + z * z < 0.25) ? 5.0 : 1.0; };
      b->system = assemble_system_matrix(*b->forest, *b->dofs, coef);
      b->ops = assemble_level_operators(*b->forest, *b->dofs, coef);
    }
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

int ms_n_levels(void *p) { return ((Bundle *)p)->forest->n_levels(); }
int64_t ms_n_cells(void *p, int l) { return ((Bundle *)p)->forest->n_cells(l); }

// dtype: 0 = int32, 1 = int64, 2 = float64, 3 = uint8, 4 = float32
}  // extern "C"

namespace ministep_capi {
int get_array(const Forest *fp, const DoFs *dp, const Csr *system, const LevelOperators *ops,
                     const std::vector<std::vector<float>> *eta, const std::vector<std::vector<char>> *flags,
                     const char *name, int l, const void **ptr, int64_t *count, int *dtype) {
  const std::string n(name);
  const Forest &f = *fp;
  auto ret = [&](const void *q, int64_t c, int t) { *ptr = q; *count = c; *dtype = t; return 0; };
  auto csr = [&](const Csr &m, const std::string &part) {
    if (part == "rowptr") return ret(m.rowptr.data(), (int64_t)m.rowptr.size(), 1);
    if (part == "col") return ret(m.col.data(), (int64_t)m.col.size(), 0);
    return ret(m.val.data(), (int64_t)m.val.size(), 2);
  };
  try {
    if (n == "ijk") return ret(f.L.at(l).ijk.data(), 3 * (int64_t)f.n_cells(l), 0);
    if (n == "parent") return ret(f.L.at(l).parent.data(), f.n_cells(l), 0);
    if (n == "child0") return ret(f.L.at(l).child0.data(), f.n_cells(l), 0);
    if (!dp) { g_err = "DoFs not built"; return -1; }
    const DoFs &d = *dp;
    if (n == "dof_xyz") return ret(d.xyz.data(), 3 * (int64_t)d.n, 0);
    if (n == "active_cells") return ret(d.active_cells.at(l).data(), (int64_t)d.active_cells[l].size(), 0);
    if (n == "cell_dofs") return ret(d.cell_dofs.at(l).data(), 8 * (int64_t)d.cell_dofs[l].size(), 0);
    if (n == "boundary") return ret(d.boundary.data(), d.n, 3);
    if (n == "hanging") return ret(d.hanging.data(), d.n, 3);
    if (n == "dirichlet") return ret(d.dirichlet.data(), d.n, 3);
    if (n == "constrained") return ret(d.constrained.data(), d.n, 3);
    if (n == "hang_rowptr") return csr(d.hang, "rowptr");
    if (n == "hang_col") return csr(d.hang, "col");
    if (n == "hang_val") return csr(d.hang, "val");
    if (n == "level_n") return ret(d.level_n.data(), (int64_t)d.level_n.size(), 0);
    if (n == "level_cell_dofs") return ret(d.level_cell_dofs.at(l).data(), 8 * (int64_t)d.level_cell_dofs[l].size(), 0);
    if (n == "level_xyz") return ret(d.level_xyz.at(l).data(), 3 * (int64_t)d.level_xyz[l].size(), 0);
    if (n == "level_edge") return ret(d.level_edge.at(l).data(), (int64_t)d.level_edge[l].size(), 3);
    if (n == "level_boundary") return ret(d.level_boundary.at(l).data(), (int64_t)d.level_boundary[l].size(), 3);
    if (n == "copy_global") return ret(d.copy_global.at(l).data(), (int64_t)d.copy_global[l].size(), 0);
    if (n == "copy_level") return ret(d.copy_level.at(l).data(), (int64_t)d.copy_level[l].size(), 0);
    if (n.rfind("sys_", 0) == 0) return csr(*system, n.substr(4));
    if (n.rfind("A_", 0) == 0) return csr(ops->A.at(l), n.substr(2));
    if (n.rfind("I_", 0) == 0) return csr(ops->I.at(l), n.substr(2));
    if (n.rfind("P_", 0) == 0) return csr(ops->P.at(l), n.substr(2));
    if (n == "eta" && eta) return ret(eta->at(l).data(), (int64_t)(*eta)[l].size(), 4);
    if (n == "flags" && flags) return ret(flags->at(l).data(), (int64_t)(*flags)[l].size(), 3);
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  g_err = "unknown array " + n;
  return -1;
}
}  // namespace ministep_capi
using ministep_capi::get_array;

extern "C" {

int ms_get(void *p, const char *name, int l, const void **ptr, int64_t *count, int *dtype) {
  Bundle *b = (Bundle *)p;
  const std::string nm(name);
  if (nm.rfind("topo_", 0) == 0) {  // face topology for gmg_error_indicator
    try {
      if (!b->dofs) { g_err = "DoFs not built"; return -1; }
      b->topo = indicator_topology(*b->forest, *b->dofs);
    } catch (std::exception &e) { g_err = e.what(); return -1; }
    if (nm == "topo_face_nb") { *ptr = b->topo.face_nb.data(); *count = (int64_t)b->topo.face_nb.size(); *dtype = 0; return 0; }
    if (nm == "topo_face_kind") { *ptr = b->topo.face_kind.data(); *count = (int64_t)b->topo.face_kind.size(); *dtype = 3; return 0; }
    if (nm == "topo_hang_children") { *ptr = b->topo.hang_children.data(); *count = (int64_t)b->topo.hang_children.size(); *dtype = 0; return 0; }
    g_err = "unknown array " + nm;
    return -1;
  }
  return get_array(b->forest.get(), b->dofs.get(), &b->system, &b->ops, &b->eta, &b->flags, name, l, ptr, count, dtype);
}

// unit-cube Q1 Laplace cell matrix (row-major 8 x 8): the k_ref argument of gmg_assemble_matrix
int ms_unit_stiffness(double *out64) {
  double K[NV][NV];
  unit_stiffness(K);
  std::memcpy(out64, K, sizeof(K));
  return 0;
}

// Sequential emulation of the device-side matrix assembly (csrc/assemble.inl runs the same row routines of
// csrc/assemble_row.h, one thread per row): builds the matrix and compares it entry by entry -- row pointer, columns and
// the BITS of the values -- with the host assembly.  which: 0 = system matrix, 1 = level matrix.  Returns the number of
// differing words in *n_diff and the longest row in *max_row.
int ms_assemble_emulate(void *p, int which, int level, int64_t *n_diff, int *max_row) {
  Bundle *b = (Bundle *)p;
  try {
    if (!b->dofs) { g_err = "DoFs not built"; return -1; }
    const DoFs &d = *b->dofs;
    const AssemblyInputs in = which == 0 ? assembly_inputs_system(*b->forest, d) : assembly_inputs_level(*b->forest, d, level);
    const Csr &ref = which == 0 ? b->system : b->ops.A.at(level);
    const bool hang = which == 0 && in.hanging;
    // incidence entries in slot order, then a stable sort by row (the device: radix sort)
    std::vector<std::pair<int, uint64_t>> ent;
    for (int64_t s = 0; s < 8 * in.n_cells; ++s)
      gmg::asm_slot_entries(in.cell_dofs.data(), in.flags.data(), hang ? d.hang.rowptr.data() : nullptr,
                            hang ? d.hang.col.data() : nullptr, s,
                            [&](int row, uint64_t e) { ent.push_back({row, e}); });
    std::stable_sort(ent.begin(), ent.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
    std::vector<int64_t> inc_ptr(in.n_rows + 1, 0);
    std::vector<uint64_t> inc(ent.size());
    for (size_t k = 0; k < ent.size(); ++k) {
      inc_ptr[ent[k].first + 1]++;
      inc[k] = ent[k].second;
    }
    for (int i = 0; i < in.n_rows; ++i) inc_ptr[i + 1] += inc_ptr[i];
    gmg::AsmView A{};
    A.n_rows = in.n_rows;
    A.n_cells = in.n_cells;
    A.cell_dofs = in.cell_dofs.data();
    A.cell_h = in.cell_h.empty() ? nullptr : in.cell_h.data();
    A.uniform_h = in.uniform_h;
    A.flags = in.flags.data();
    A.hang_ptr = hang ? d.hang.rowptr.data() : nullptr;
    A.hang_col = hang ? d.hang.col.data() : nullptr;
    A.hang_val = hang ? d.hang.val.data() : nullptr;
    A.inc_ptr = inc_ptr.data();
    A.inc = inc.data();
    double K[NV][NV];
    unit_stiffness(K);
    std::memcpy(A.kref, K, sizeof(K));
    constexpr int MAXC = 1024;
    int64_t diff = 0;
    int longest = 0;
    if (ref.n_rows != in.n_rows) { g_err = "row count differs"; return -1; }
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : diff) reduction(max : longest)
    for (int i = 0; i < in.n_rows; ++i) {
      int cols[MAXC];
      double vals[MAXC];
      const int n = gmg::asm_row_pattern(A, i, cols, MAXC);
      const int64_t r0 = ref.rowptr[i], r1 = ref.rowptr[i + 1];
      if (n < 0 || n != (int)(r1 - r0)) {
        diff += 1 + (r1 - r0);
        continue;
      }
      longest = std::max(longest, n);
      gmg::asm_row_values(A, i, cols, n, vals);
      for (int k = 0; k < n; ++k) {
        if (cols[k] != ref.col[r0 + k]) ++diff;
        if (std::memcmp(&vals[k], &ref.val[r0 + k], sizeof(double)) != 0) ++diff;
      }
    }
    *n_diff = diff;
    *max_row = longest;
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

int ms_error_indicator(void *p, const double *u, int64_t n_rho, const double *rho, int nq, int residual_term,
                       double *threshold) {
  Bundle *b = (Bundle *)p;
  try {
    std::vector<double> uu(u, u + b->dofs->n), rr(rho, rho + n_rho);
    b->eta = error_indicator(*b->forest, *b->dofs, uu, rr, nq, residual_term != 0);
    b->threshold = mark_cells(*b->forest, *b->dofs, b->eta, b->flags);
    *threshold = b->threshold;
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

int ms_transfer(void *p_new, int old_res, void *p_old, const double *u_old, double *x_out) {
  Bundle *nb = (Bundle *)p_new, *ob = (Bundle *)p_old;
  try {
    std::vector<double> uo(u_old, u_old + ob->dofs->n);
    std::vector<double> x = transfer_solution(old_res, *ob->dofs, uo, *nb->forest, *nb->dofs);
    std::memcpy(x_out, x.data(), sizeof(double) * x.size());
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

// index tables of the device solution transfer (gmg_transfer_solution); arrays owned by the new bundle until the next call
int ms_transfer_tables(void *p_new, int old_res, void *p_old, int64_t *n_copy, const int32_t **copy_old, const int32_t **copy_new,
                       int64_t *n_pass, const int64_t **pass_ptr, const int32_t **parent_dofs) {
  Bundle *nb = (Bundle *)p_new, *ob = (Bundle *)p_old;
  try {
    static thread_local TransferTables T;
    T = transfer_tables(old_res, *ob->dofs, *nb->forest, *nb->dofs);
    *n_copy = (int64_t)T.copy_old.size();
    *copy_old = T.copy_old.data();
    *copy_new = T.copy_new.data();
    *n_pass = (int64_t)T.pass_ptr.size() - 1;
    *pass_ptr = T.pass_ptr.data();
    *parent_dofs = T.parent_dofs.data();
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

int ms_distribute(void *p, const double *g, double *x_inout) {
  Bundle *b = (Bundle *)p;
  std::vector<double> gg(g, g + b->dofs->n), x(x_inout, x_inout + b->dofs->n);
  distribute(*b->dofs, gg, x);
  std::memcpy(x_inout, x.data(), sizeof(double) * x.size());
  return 0;
}

int ms_locate(void *p, const double *X, int *level, int *cell, double *xi) {
  Bundle *b = (Bundle *)p;
  locate(*b->forest, *b->dofs, X, *level, *cell, xi);
  return 0;
}

// Gauss rule on [0,1] (host copy of QGauss<1>(n))
int ms_gauss(int n, double *pts, double *wts) {
  std::vector<double> p, w;
  gauss_unit(n, p, w);
  std::memcpy(pts, p.data(), sizeof(double) * n);
  std::memcpy(wts, w.data(), sizeof(double) * n);
  return 0;
}

}  // extern "C"
