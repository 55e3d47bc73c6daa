// extern "C" shim over the host library for the Python test harness (tests/ only): drive ministep
// without a GPU (numbering / sparsity / assembly parity against the oracle) and run the full
// LaplaceProblem from a parameter string (needs the GPU).
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <sstream>

#include "step_50.h"

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
  std::vector<int32_t> scratch_i32;
  std::vector<double> scratch_f64;
  std::string err;
};
thread_local std::string g_err;
}  // namespace

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
int ms_get(void *p, const char *name, int l, const void **ptr, int64_t *count, int *dtype) {
  Bundle *b = (Bundle *)p;
  const std::string n(name);
  const Forest &f = *b->forest;
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
    if (!b->dofs) { g_err = "ms_build first"; return -1; }
    const DoFs &d = *b->dofs;
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
    if (n.rfind("sys_", 0) == 0) return csr(b->system, n.substr(4));
    if (n.rfind("A_", 0) == 0) return csr(b->ops.A.at(l), n.substr(2));
    if (n.rfind("I_", 0) == 0) return csr(b->ops.I.at(l), n.substr(2));
    if (n.rfind("P_", 0) == 0) return csr(b->ops.P.at(l), n.substr(2));
    if (n == "eta") return ret(b->eta.at(l).data(), (int64_t)b->eta[l].size(), 4);
    if (n == "flags") return ret(b->flags.at(l).data(), (int64_t)b->flags[l].size(), 3);
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  g_err = "unknown array " + n;
  return -1;
}

int ms_error_indicator(void *p, const double *u, int64_t n_rho, const double *rho, int nq, double *threshold) {
  Bundle *b = (Bundle *)p;
  try {
    std::vector<double> uu(u, u + b->dofs->n), rr(rho, rho + n_rho);
    b->eta = error_indicator(*b->forest, *b->dofs, uu, rr, nq);
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

// parse a parameter string with the reference's ParameterHandler grammar; returns 0 / -1 (message in ms_last_error)
int step50_check_prm(const char *text, char **echo) {
  try {
    ParameterHandler prm;
    ParameterReader param(prm);
    param.declare_parameters();
    prm.parse_input_from_string(text);
    std::ostringstream os;
    prm.enter_subsection("Geometry");
    os << prm.get_double("Domain limit left") << " " << prm.get_double("Domain limit right") << " "
       << prm.get_double("Mesh size") << " " << prm.get_integer("Vacuum repetitions") << " ";
    prm.leave_subsection();
    prm.enter_subsection("Misc");
    os << prm.get_integer("Number of Adaptive Refinement") << " " << prm.get_bool("Flag for RHS evaluation optimization") << " ";
    prm.leave_subsection();
    prm.enter_subsection("Solver input data");
    os << prm.get("Preconditioner") << " " << prm.get("Smoother");
    prm.leave_subsection();
    if (echo) *echo = strdup(os.str().c_str());
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

// run LaplaceProblem from a parameter string; stdout text and a JSON array of cycle records are returned
// in malloc'ed strings (free with step50_free).  Needs a B200.
int step50_run_string(const char *prm_text, char **stdout_text, char **records_json) {
  try {
    std::ostringstream out;
    std::vector<Step50::CycleRecord> recs;
    step50_run_from_string(prm_text, out, &recs);
    std::ostringstream js;
    js << std::setprecision(17) << "[";
    for (size_t i = 0; i < recs.size(); ++i) {
      const auto &r = recs[i];
      js << (i ? "," : "") << "{\"n_active_cells\":" << r.n_active_cells << ",\"n_dofs\":" << r.n_dofs << ",\"n_dofs_level\":[";
      for (size_t k = 0; k < r.n_dofs_level.size(); ++k) js << (k ? "," : "") << r.n_dofs_level[k];
      js << "],\"rhs_l1\":" << r.rhs_l1 << ",\"rhs_l2\":" << r.rhs_l2 << ",\"rhs_linf\":" << r.rhs_linf
         << ",\"mat_l1\":" << r.mat_l1 << ",\"mat_linf\":" << r.mat_linf << ",\"mat_frob\":" << r.mat_frob
         << ",\"start\":" << r.start << ",\"its\":" << r.its << ",\"conv\":" << r.conv << ",\"sol_l1\":" << r.sol_l1
         << ",\"sol_l2\":" << r.sol_l2 << ",\"sol_linf\":" << r.sol_linf << ",\"threshold\":" << r.threshold
         << ",\"n_flagged\":" << r.n_flagged << ",\"solve_seconds\":" << r.solve_seconds
         << ",\"rhs_seconds\":" << r.rhs_seconds << ",\"coarse_its\":[";
      for (size_t k = 0; k < r.coarse_its.size(); ++k) js << (k ? "," : "") << r.coarse_its[k];
      js << "]";
      if (r.have_energy)
        js << ",\"energy\":{\"analytic\":" << r.e_analytic << ",\"short\":" << r.e_short << ",\"fe\":" << r.e_fe
           << ",\"self\":" << r.e_self << ",\"total\":" << r.e_total << "},\"energy_norm_error\":" << r.energy_norm_error;
      js << "}";
    }
    js << "]";
    if (stdout_text) *stdout_text = strdup(out.str().c_str());
    if (records_json) *records_json = strdup(js.str().c_str());
  } catch (std::exception &e) { g_err = e.what(); return -1; }
  return 0;
}

void step50_free(char *p) { free(p); }

}  // extern "C"
