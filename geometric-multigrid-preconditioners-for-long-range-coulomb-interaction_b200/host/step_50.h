// Host-side mirror of the reference's application class for the B200 path.
//
// Same public surface as `Step50::LaplaceProblem<dim>` / `ParameterReader` in the reference's
// include/step_50.h:111-214 (20-argument constructor in the order of src/main.cc:72-76, `run()`,
// protected setup_system / assemble_system / assemble_multigrid / solve / refine_grid / ... that tests
// reach by subclassing) and the same stdout.  deal.II, p4est and Trilinos are replaced by `ministep`
// (host mesh / DoFs / assembly) and by the CUDA library behind include/gmg_b200.h (everything
// floating-point on the hot path).  dim == 3 only: the atom path of the reference is 3D only.
#ifndef STEP_50_B200_H
#define STEP_50_B200_H

#include <chrono>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/gmg_b200.h"
#include "ministep.h"
#include "rendezvous.h"
#include "parameter_handler.h"

namespace Step50 {

// SolverControl::NoConvergence
class NoConvergence : public std::runtime_error {
 public:
  explicit NoConvergence(const std::string &m) : std::runtime_error(m) {}
};
class ExcMessage : public std::runtime_error {
 public:
  explicit ExcMessage(const std::string &m) : std::runtime_error(m) {}
};

// dealii::TimerOutput(wall_times) stand-in: named sections, summary table
class TimerOutput {
 public:
  class Scope {
   public:
    Scope(TimerOutput &t, const std::string &name) : timer(t), section(name), t0(std::chrono::steady_clock::now()) {}
    ~Scope() {
      auto &s = timer.sections[section];
      s.first += 1;
      s.second += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

   private:
    TimerOutput &timer;
    std::string section;
    std::chrono::steady_clock::time_point t0;
  };
  void reset() {
    sections.clear();
    start = std::chrono::steady_clock::now();
  }
  void print_summary(std::ostream &out) const;
  double total() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count(); }
  std::map<std::string, std::pair<int, double>> sections;
  std::chrono::steady_clock::time_point start = std::chrono::steady_clock::now();
};

// flattened active cells (lower corner, edge, dofs, atom-list id), cached per mesh
struct ActiveCells {
  std::vector<double> lo, h;
  std::vector<int32_t> dofs, list;
};

// what one refinement cycle printed / computed (kept for the test shim)
struct CycleRecord {
  long n_active_cells = 0, n_dofs = 0;
  std::vector<long> n_dofs_level;
  double rhs_l1 = 0, rhs_l2 = 0, rhs_linf = 0, mat_l1 = 0, mat_linf = 0, mat_frob = 0;
  double start = 0, conv = 0, sol_l1 = 0, sol_l2 = 0, sol_linf = 0, threshold = 0;
  int its = 0;
  long n_flagged = 0;
  std::vector<int> coarse_its;
  bool have_energy = false;
  double e_analytic = 0, e_short = 0, e_fe = 0, e_self = 0, e_total = 0, energy_norm_error = 0;
  double solve_seconds = 0, rhs_seconds = 0;
};

template <int dim>
class LaplaceProblem {
 public:
  LaplaceProblem(const unsigned int degree, ParameterHandler &param, const std::string &Problemtype,
                 const std::string &PreconditionerType, const std::string &LammpsInputFile,
                 const std::string &Boundary_conditions, const double &domain_size_left,
                 const double &domain_size_right, const double &mesh_size_h, const unsigned int &repetitions_for_vacuum,
                 const unsigned int &number_of_global_refinement,
                 const unsigned int &number_of_adaptive_refinement_cycles, const double &r_c,
                 const double &nonzero_density_radius_parameter, const bool &flag_rhs_assembly,
                 const bool &flag_analytical_solution, const bool &flag_rhs_field, const bool &flag_atoms_support,
                 const bool &flag_output_time, const unsigned int &quadrature_degree_rhs);
  ~LaplaceProblem();
  void run();

  // B200-path extras (not in the reference): where stdout goes, what the cycles produced
  void set_output(std::ostream &os) { pcout = &os; }
  const std::vector<CycleRecord> &records() const { return cycle_records; }
  gmg_handle device() const { return gmg; }

 protected:
  void setup_system(const unsigned int &cycle);
  void assemble_system();
  void assemble_multigrid();
  void solve();
  void estimate_error_and_mark_cells();
  void refine_grid(const unsigned int &cycle);
  void read_lammps_input_file(const std::string &filename);
  void output_results(const unsigned int cycle) const;
  void rhs_assembly_optimization();
  void compute_charge_densities();
  void compute_moments();
  void postprocess_electrostatic_energy();
  void postprocess_error_in_energy_norm();
  double long_ranged_potential(const double p[3], const double atom[3], const double &charge) const;
  double short_ranged_potential(const double p[3], const double atom[3], const double &charge) const;
  void make_mesh();
  // run() = begin_run(); for every cycle { cycle_until_solve; solve; cycle_after_solve }; end_run()
  void begin_run();
  void cycle_until_solve(const unsigned int cycle);
  void cycle_after_solve(const unsigned int cycle);
  void end_run();
  void assemble_rhs_on_device();
  void hand_over_hierarchy();
  void boundary_values();

  std::ostream *pcout;
  TimerOutput computing_timer;
  ParameterHandler &prm;
  const unsigned int degree;
  const unsigned int number_of_global_refinement, number_of_adaptive_refinement_cycles;
  const double domain_size_left, domain_size_right, mesh_size_h;
  const unsigned int repetitions_for_vacuum;
  const std::string Problemtype, PreconditionerType, LammpsInputFilename, Boundary_conditions;
  bool lammpsinput = false;
  const bool flag_analytical_solution, flag_rhs_field, flag_atoms_support, flag_rhs_assembly, flag_output_time;
  unsigned int number_of_atoms = 0;
  std::vector<double> atom_positions;  // [n][3]
  std::vector<unsigned int> atom_types;
  std::vector<double> charges;
  const double r_c, nonzero_density_radius_parameter;
  const unsigned int quadrature_degree_rhs;
  double dipole_moment[3] = {0, 0, 0};

  // smoother selection: a source-level toggle in the reference (src/step-50.cc:969-973), prm keys here
  int smoother_kind = GMG_SMOOTHER_LEX_SSOR;
  double smoother_omega = 0.5;
  int smoothing_steps = 2;
  int gpu_device = 0;
  bool densities_on_host = false;  // compute_charge_densities also downloads the densities (no consumer on the host)
  unsigned int energy_atom_limit = 300;
  unsigned int energy_norm_atom_limit = 0;  // 0: always (reference source)
  bool indicator_with_residual = true;  // false: Kelly part only (the build behind the cluster logs)
  bool zero_initial_guess = false;      // true: every cycle's solve starts from 0 (the build behind the step-16 goldens)
  // Matrix assembly = Device: the system matrix and the level-0 matrix are assembled on the GPU from the cell -> dof
  // maps (gmg_assemble_matrix: the same CSR, bit for bit) instead of on the host; patch levels stay on the host
  bool device_assembly = false;
  // Coarse levels below the base mesh (0 = the reference's hierarchy): level base_level() is the base lattice
  int coarse_levels_below_base = 0;
  int base_level() const { return Problemtype == "Step16" ? 0 : coarse_levels_below_base; }
  std::vector<double> rhs_ghat;          // inhomogeneities resolved through the hanging-node lines, per mesh
  std::vector<uint8_t> rhs_constrained;  // constraints.is_constrained(i) as bytes, per mesh
  bool rhs_inhom = false;
  const void *rhs_cells_on_device = nullptr;  // the cell arrays gmg_assemble_rhs last uploaded (system assembly reuses them)
  std::vector<int> hanging_list;  // hanging dofs, ascending (constraints.distribute), per mesh
  int hanging_list_n = -1;
  std::vector<uint8_t> asm_flags_system, asm_flags_level0;  // row flags of gmg_assemble_matrix, built once per mesh
  bool assemble_on_device() const { return device_assembly && Problemtype != "Step16" && PreconditionerType == "GMG"; }

  std::unique_ptr<ministep::Forest> triangulation;
  std::unique_ptr<ministep::DoFs> mg_dof_handler;
  std::shared_ptr<ActiveCells> active_cells_cache;
  ministep::Csr system_matrix;
  ministep::LevelOperators mg_ops;  // mg_matrices, mg_interface_matrices, transfer matrices
  std::vector<double> solution, system_rhs, boundary_g, distributed_solution;
  std::vector<double> density_values;      // active cells x nq^3 (density_values_for_each_cell)
  std::vector<int64_t> charges_list_ptr;   // charges_list_for_each_cell of the base cells (CSR);
  std::vector<int32_t> charges_list_atoms; // children inherit through their level-0 ancestor
  std::vector<std::vector<float>> error_per_cell;
  std::vector<std::vector<char>> refine_flags;
  gmg_handle gmg = nullptr;
  // multi-GPU run(): one process per GPU (the reference: one MPI rank per subdomain, src/step-50.cc:116-122); every
  // rank keeps the whole mesh and the patch levels, the system matrix and level 0 are row-partitioned on the devices
  std::unique_ptr<Rendezvous> ranks;
  bool connect_in_run = false;  // run() joins the ranks the launcher started; the bench hooks connect through the caller
  std::ostream null_out{nullptr};
  void connect_ranks();
  std::chrono::steady_clock::time_point run_start;
  std::vector<CycleRecord> cycle_records;
  CycleRecord *rec = nullptr;
  void gmg_check(int rc, const char *what);
};

}  // namespace Step50

class ParameterReader {
 public:
  explicit ParameterReader(ParameterHandler &);
  void read_parameters(const std::string &);
  void declare_parameters();

 private:
  ParameterHandler &prm;
};

// the reference's main(): parse the .prm, build LaplaceProblem<2|3>, run (src/main.cc:6-121)
int step50_main(int argc, char **argv, std::ostream &out);
// same from a parameter string (tests/gaussian-charges.cc builds its parameters this way)
void step50_run_from_string(const std::string &prm_text, std::ostream &out,
                            std::vector<Step50::CycleRecord> *records = nullptr);

#endif
