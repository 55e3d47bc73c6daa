// ministep: the minimal host substrate the B200 path needs in place of deal.II / p4est, which are
// absent here: an octree forest on a structured base lattice, Q1 DoF
// numbering, hanging-node / Dirichlet constraints, sparsity patterns and CSR assembly of the system,
// level, interface and prolongation matrices, the error indicator and the solution transfer.
// It restates the deal.II 9.0 semantics the reference relies on (src/step-50.cc:646-933, 1020-1121);
// every routine names the reference lines it stands in for.  3D only (the LAMMPS path of the
// reference is 3D only, src/step-50.cc:195, 252-256).
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <unordered_map>
#include <vector>

namespace ministep {

constexpr int DIM = 3;
constexpr int NV = 8;  // vertices = Q1 dofs per cell; vertex v has offset bit d of v along axis d

using Int3 = std::array<int, 3>;
using Dofs8 = std::array<int, NV>;

inline int vo(int v, int d) { return (v >> d) & 1; }

struct Csr {
  int n_rows = 0, n_cols = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<double> val;
  int64_t nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
};

// ---- mesh (parallel::distributed::Triangulation with limit_level_difference_at_vertices and
//      construct_multigrid_hierarchy, src/step-50.cc:120-122)
class Forest {
 public:
  Forest(int reps, double lo, double hi);
  int reps;
  double lo, hi, H;
  struct LevelCells {
    std::vector<Int3> ijk;
    std::vector<int> parent, child0;
    std::unordered_map<int64_t, int> index;  // levels >= 1
  };
  std::vector<LevelCells> L;

  int n_levels() const { return (int)L.size(); }
  int64_t cells_per_axis(int l) const { return (int64_t)reps << l; }
  double h(int l) const { return H / (double)(1 << l); }
  int n_cells(int l) const { return (int)L[l].ijk.size(); }
  bool active(int l, int c) const { return L[l].child0[c] < 0; }
  int64_t n_active_cells() const;
  int lookup(int l, int i, int j, int k) const;  // -1: none / outside
  bool inside(int l, int i, int j, int k) const;
  int resolution() const { return n_levels() - 1; }
  int64_t points_per_axis() const { return ((int64_t)reps << resolution()) + 1; }
  // refine flagged active cells after closing the flags under 2:1 balance over faces, edges and
  // corners (p4est_balance(P8EST_CONNECT_FULL)); children appended in parent order
  void refine(std::vector<std::vector<char>> flags);
  void refine_global(int times);
};

// ---- DoFs and constraints (setup_system, src/step-50.cc:646-732)
struct DoFs {
  explicit DoFs(const Forest &f);
  const Forest &f;
  int res;
  int n = 0;
  std::vector<std::vector<int>> active_cells;   // per level, ascending
  std::vector<std::vector<int>> active_pos;     // per level: cell -> position in active_cells or -1
  std::vector<std::vector<Dofs8>> cell_dofs;    // per level, per active cell: global dofs
  std::vector<Int3> xyz;                        // integer vertex coordinates (resolution res) of each dof
  std::unordered_map<int64_t, int> key2dof;
  std::vector<char> boundary, hanging, dirichlet, constrained;
  Csr hang;                                     // rows: dofs; entries: (parent dof, weight) of hanging dofs
  // level data (distribute_mg_dofs, MGConstrainedDoFs)
  std::vector<int> level_n;
  std::vector<std::vector<Dofs8>> level_cell_dofs;  // per level, all cells
  std::vector<std::vector<Int3>> level_xyz;
  std::vector<std::vector<char>> level_boundary, level_edge;
  std::vector<std::vector<int32_t>> copy_global, copy_level;

  int64_t key(const Int3 &p) const;
  int lookup(const Int3 &p) const;
  std::array<double, 3> coords(int dof) const;
};

// Q1 Laplace cell matrix of the unit cube by 2x2x2 Gauss quadrature; a cube of edge h has h * Kref
void unit_stiffness(double Kref[NV][NV]);
// per-quadrature-point contributions (coefficient != 1): K = h * sum_q c_q G[q]
void unit_stiffness_q(double G[8][NV][NV], double pts[8][3]);
void gauss_unit(int n, std::vector<double> &pts, std::vector<double> &wts);

using Coefficient = std::function<double(double, double, double)>;  // nullptr-like (empty) = 1

// system_matrix with constraints condensed + make_sparsity_pattern(dof, dsp, constraints, true)
// (assemble_system matrix part, src/step-50.cc:771-795)
Csr assemble_system_matrix(const Forest &f, const DoFs &d, const Coefficient &coef);
struct LevelOperators {
  std::vector<Csr> A, I, P;  // mg_matrices, mg_interface_matrices, prolongation l -> l+1
};
// assemble_multigrid (src/step-50.cc:835-933) + MGTransferPrebuilt::build_matrices (:957-958)
// first_level > 0: A and I of the levels below it are left empty (assembled elsewhere: gmg_assemble_matrix)
LevelOperators assemble_level_operators(const Forest &f, const DoFs &d, const Coefficient &coef, int first_level = 0);

// Inputs of the device-side matrix assembly (gmg_assemble_matrix, include/gmg_b200.h): the cell -> dof map in the
// order of the sequential cell loop, the cell sizes and one flag byte per row (bit 0: eliminated row / column =
// Dirichlet dof, level boundary or refinement-edge dof; bit 1: hanging dof, constraint line in DoFs::hang).
struct AssemblyInputs {
  int n_rows = 0;
  int64_t n_cells = 0;
  std::vector<int32_t> cell_dofs;  // [n_cells][8]
  std::vector<double> cell_h;      // [n_cells]; empty: uniform_h
  double uniform_h = 0.0;
  std::vector<uint8_t> flags;      // [n_rows]
  bool hanging = false;            // any hanging dof (the constraint lines are DoFs::hang)
};
AssemblyInputs assembly_inputs_system(const Forest &f, const DoFs &d);          // active cells, level by level
AssemblyInputs assembly_inputs_level(const Forest &f, const DoFs &d, int level);  // all cells of the level

// b_i -= sum_j K_ij ghat_j over unconstrained rows is done on the device; the host only provides
// ghat = T g (hanging nodes with Dirichlet parents get their interpolated value)
std::vector<double> resolve_inhomogeneity(const DoFs &d, const std::vector<double> &g);
// constraints.distribute(solution) (src/step-50.cc:1016)
void distribute(const DoFs &d, const std::vector<double> &g, std::vector<double> &x);

// face topology of the active cells for the device indicator (gmg_error_indicator, include/gmg_b200.h): active cells
// numbered level by level in active order (the order of the RHS cell arrays)
struct IndicatorTopology {
  std::vector<int32_t> face_nb;        // [n_active][6]
  std::vector<uint8_t> face_kind;      // [n_active][6]
  std::vector<int32_t> hang_children;  // [n_hang][4]
};
IndicatorTopology indicator_topology(const Forest &f, const DoFs &d);
// Kelly(cell_diameter) + h_K^2 * int (4 pi rho)^2, Vector<float> storage (src/step-50.cc:1020-1090)
std::vector<std::vector<float>> error_indicator(const Forest &f, const DoFs &d, const std::vector<double> &u,
                                                const std::vector<double> &rho /*active cells x nq^3*/, int nq,
                                                bool residual_term = true);
// threshold = 0.6 max eta; flags where eta >= threshold (GridRefinement::refine, :1084-1089)
double mark_cells(const Forest &f, const DoFs &d, const std::vector<std::vector<float>> &eta,
                  std::vector<std::vector<char>> &flags);
// SolutionTransfer::interpolate + constraints.set_zero (src/step-50.cc:1118-1119)
std::vector<double> transfer_solution(int old_res, const DoFs &old_dofs, const std::vector<double> &u_old,
                                      const Forest &f, const DoFs &d);
// The index side of the same transfer for the device (gmg_transfer_solution): which old dof lands on which new dof, and
// per level (coarse to fine) the 27 new dofs (t0 + 3 t1 + 9 t2, -1: no such dof) of every refined cell, in cell order.
struct TransferTables {
  std::vector<int32_t> copy_old, copy_new;
  std::vector<int64_t> pass_ptr;      // [levels]: first refined cell of each pass in parent_dofs
  std::vector<int32_t> parent_dofs;   // [refined cells][27]
};
TransferTables transfer_tables(int old_res, const DoFs &old_dofs, const Forest &f, const DoFs &d);
// find_active_cell_around_point + unit-cell coordinates (src/step-50.cc:1353-1356)
void locate(const Forest &f, const DoFs &d, const double X[3], int &level, int &cell, double xi[3]);

}  // namespace ministep
