// Host side of the B200 path: the reference's LaplaceProblem life cycle (src/step-50.cc:104-1573)
// with deal.II/Trilinos replaced by ministep (mesh, DoFs, assembly: host) and gmg_b200 (RHS path,
// V-cycle, PCG: CUDA).  Method by method it cites the reference lines it stands in for.
#include "step_50.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <thread>

#include <signal.h>
#include <sys/wait.h>

using namespace ministep;

// ------------------------------------------------------------------------------- ParameterReader
ParameterReader::ParameterReader(ParameterHandler &paramhandler) : prm(paramhandler) {}

void ParameterReader::declare_parameters() {
  // the reference's 19 entries, names / defaults / patterns unchanged (src/step-50.cc:15-95)
  prm.enter_subsection("Geometry");
  // new: k > 0 builds the base lattice as k global refinements of a lattice 2^k times coarser, so the
  // multigrid hierarchy continues below the base mesh and the coarse-grid CG runs on (reps / 2^k + 1)^3 dofs instead of
  // (reps + 1)^3.  0 = the reference's hierarchy (coarse solve on the base lattice).
  prm.declare_entry("Coarse levels below the base mesh", "0", Patterns::Integer(),
                    "Number of geometric coarsenings of the base lattice added below it as multigrid levels");
  prm.declare_entry("Number of global refinement", "2", Patterns::Integer(),
                    "The uniform global mesh refinement on the Domain in the power of 4");
  prm.declare_entry("Domain limit left", "-1", Patterns::Double(), "Left limit of domain");
  prm.declare_entry("Domain limit right", "1", Patterns::Double(), "Right limit of domain");
  prm.declare_entry("Mesh size", "0.25", Patterns::Double(), "Mesh size for initial domain");
  prm.declare_entry("Vacuum repetitions", "1", Patterns::Integer(),
                    "Number of repetitions for vacuum on each side in terms of 2 * Mesh size");
  prm.leave_subsection();
  prm.enter_subsection("Problem Selection");
  prm.declare_entry("Problem", "Step16", Patterns::Selection("Step16 | GaussianCharges"),
                    "Problem definition for RHS Function");
  prm.declare_entry("Dimension", "2", Patterns::Integer(), "Problem space dimension");
  prm.declare_entry("Boundary conditions selection", "Inhomogeneous",
                    Patterns::Selection("Homogeneous | Inhomogeneous | Exact"),
                    "Selection between Homogeneous, Inhomogeneous or Exact dirichlet boundary condtions");
  prm.leave_subsection();
  prm.enter_subsection("Misc");
  prm.declare_entry("Number of Adaptive Refinement", "2", Patterns::Integer(),
                    "Number of Adaptive refinement cycles to be done");
  prm.declare_entry("smoothing length", "0.5", Patterns::Double(),
                    "The smoothing length parameter for each Gaussian atom");
  prm.declare_entry("Nonzero Density radius parameter around each charge", "3", Patterns::Double(),
                    "Set the parameter to localize the density around each charge where it is nonzero");
  prm.declare_entry("Output and calculation of Analytical solution", "false", Patterns::Bool(),
                    "Set flag for whether to calculate and output the analytical solution");
  prm.declare_entry("Output of RHS field", "false", Patterns::Bool(), "Set flag for whether to output the RHS field");
  prm.declare_entry("Output of support of each atom", "false", Patterns::Bool(),
                    "Set flag for whether to output the support of each atom");
  prm.declare_entry("Flag for RHS evaluation optimization", "false", Patterns::Bool(),
                    "Set flag for whether to evaluate the RHS field with local optimization");
  prm.declare_entry("Quadrature points for RHS function", "1", Patterns::Integer(),
                    "Number of quadrature points for RHS function (total points = degree + these points)");
  prm.declare_entry("Output time summary table", "true", Patterns::Bool(),
                    "Set flag for whether to output the time summary");
  // new, B200 path only (defaults keep the reference's behaviour)
  prm.declare_entry("Refinement indicator", "KellyAndResidual", Patterns::Selection("KellyAndResidual | Kelly"),
                    "KellyAndResidual = the shipped source (src/step-50.cc:1040-1081); Kelly = the older build that "
                    "produced the cluster logs (marks with the Kelly estimator alone)");
  prm.declare_entry("Initial guess", "Transferred", Patterns::Selection("Transferred | Zero"),
                    "Transferred = the shipped source: the solve of a refined mesh starts from the interpolated solution "
                    "(src/step-50.cc:1095-1121); Zero = the older build behind the step-16 / tests_2D / tests_3D goldens");
  prm.declare_entry("Energy postprocessing atom limit", "300", Patterns::Integer(),
                    "postprocess_electrostatic_energy runs only below this atom count (reference: 300)");
  prm.declare_entry("Energy norm error atom limit", "0", Patterns::Integer(),
                    "postprocess_error_in_energy_norm runs only below this atom count; 0 = always, as the shipped "
                    "reference source does (src/step-50.cc:1555; the builds behind the cluster logs did not have it)");
  prm.leave_subsection();
  prm.declare_entry("Polynomial degree", "1", Patterns::Integer(), "Polynomial degree of finite elements");
  prm.enter_subsection("Solver input data");
  prm.declare_entry("Preconditioner", "GMG", Patterns::Selection("GMG | Jacobi"),
                    "Preconditioner type to be applied to the system matrix");
  // new: the smoother is a source-level toggle in the reference (src/step-50.cc:969-973)
  prm.declare_entry("Smoother", "SSOR", Patterns::Selection("SSOR | MulticolourSSOR | Jacobi | Chebyshev"),
                    "Multigrid smoother: SSOR = the reference's lexicographic SSOR (level-scheduled on the GPU)");
  prm.declare_entry("Smoother relaxation", "0.5", Patterns::Double(), "Damping factor of the smoother");
  prm.declare_entry("Smoothing steps", "2", Patterns::Integer(), "Pre- and post-smoothing steps on every level");
  prm.declare_entry("GPU device", "0", Patterns::Integer(), "CUDA device ordinal");
  prm.declare_entry("Keep densities on the host", "false", Patterns::Bool(),
                    "compute_charge_densities also copies the densities back to the host (they are consumed on the device)");
  prm.declare_entry("Matrix assembly", "Host", Patterns::Selection("Host | Device"),
                    "Device: system and level-0 matrices are assembled on the GPU from the cell-dof maps "
                    "(bit-identical to the host assembly) instead of being handed over assembled");
  prm.leave_subsection();
  prm.enter_subsection("Lammps data");
  prm.declare_entry("Lammps input file", "atom_8.data", Patterns::Anything(),
                    "Lammps input file with atoms, charges and positions");
  prm.leave_subsection();
}

void ParameterReader::read_parameters(const std::string &parameter_file) { prm.parse_input(parameter_file); }

namespace Step50 {

void TimerOutput::print_summary(std::ostream &out) const {
  const double tot = total();
  char buf[256];
  out << "\n\n+---------------------------------------------+------------+------------+\n";
  std::snprintf(buf, sizeof buf, "| Total wallclock time elapsed since start    | %9.3es |            |\n", tot);
  out << buf;
  out << "|                                             |            |            |\n";
  out << "| Section                         | no. calls |  wall time | % of total |\n";
  out << "+---------------------------------+-----------+------------+------------+\n";
  for (auto &s : sections) {
    std::string name = s.first.substr(0, 32);
    std::snprintf(buf, sizeof buf, "| %-32s| %9d | %9.3es | %9.2e%% |\n", name.c_str(), s.second.first, s.second.second,
                  100.0 * s.second.second / std::max(tot, 1e-300));
    out << buf;
  }
  out << "+---------------------------------+-----------+------------+------------+\n\n";
}

// =============================================================================== construction
template <int dim>
LaplaceProblem<dim>::LaplaceProblem(
    const unsigned int degree_, ParameterHandler &param, const std::string &Problemtype_,
    const std::string &PreconditionerType_, const std::string &LammpsInputFile, const std::string &Boundary_conditions_,
    const double &domain_size_left_, const double &domain_size_right_, const double &mesh_size_h_,
    const unsigned int &repetitions_for_vacuum_, const unsigned int &number_of_global_refinement_,
    const unsigned int &number_of_adaptive_refinement_cycles_, const double &r_c_,
    const double &nonzero_density_radius_parameter_, const bool &flag_rhs_assembly_,
    const bool &flag_analytical_solution_, const bool &flag_rhs_field_, const bool &flag_atoms_support_,
    const bool &flag_output_time_, const unsigned int &quadrature_degree_rhs_)
    : pcout(&std::cout),
      prm(param),
      degree(degree_),
      number_of_global_refinement(number_of_global_refinement_),
      number_of_adaptive_refinement_cycles(number_of_adaptive_refinement_cycles_),
      domain_size_left(domain_size_left_),
      domain_size_right(domain_size_right_),
      mesh_size_h(mesh_size_h_),
      repetitions_for_vacuum(repetitions_for_vacuum_),
      Problemtype(Problemtype_),
      PreconditionerType(PreconditionerType_),
      LammpsInputFilename(LammpsInputFile),
      Boundary_conditions(Boundary_conditions_),
      flag_analytical_solution(flag_analytical_solution_),
      flag_rhs_field(flag_rhs_field_),
      flag_atoms_support(flag_atoms_support_),
      flag_rhs_assembly(flag_rhs_assembly_),
      flag_output_time(flag_output_time_),
      r_c(r_c_),
      nonzero_density_radius_parameter(nonzero_density_radius_parameter_),
      quadrature_degree_rhs(quadrature_degree_rhs_) {
  if (degree != 1) throw ExcMessage("The B200 path implements Q1 elements only (Polynomial degree = 1).");
  // optional B200 keys (present when ParameterReader::declare_parameters of this library declared them)
  try {
    prm.enter_subsection("Solver input data");
    const std::string sm = prm.get("Smoother");
    smoother_kind = sm == "SSOR" ? GMG_SMOOTHER_LEX_SSOR
                    : sm == "MulticolourSSOR" ? GMG_SMOOTHER_MC_SSOR
                    : sm == "Jacobi" ? GMG_SMOOTHER_JACOBI
                                     : GMG_SMOOTHER_CHEBYSHEV;
    smoother_omega = prm.get_double("Smoother relaxation");
    smoothing_steps = (int)prm.get_integer("Smoothing steps");
    gpu_device = (int)prm.get_integer("GPU device");
    densities_on_host = prm.get_bool("Keep densities on the host");
    device_assembly = prm.get("Matrix assembly") == "Device";
    prm.leave_subsection();
    prm.enter_subsection("Geometry");
    coarse_levels_below_base = (int)prm.get_integer("Coarse levels below the base mesh");
    prm.leave_subsection();
    prm.enter_subsection("Misc");
    energy_atom_limit = (unsigned int)prm.get_integer("Energy postprocessing atom limit");
    energy_norm_atom_limit = (unsigned int)prm.get_integer("Energy norm error atom limit");
    indicator_with_residual = prm.get("Refinement indicator") == "KellyAndResidual";
    zero_initial_guess = prm.get("Initial guess") == "Zero";
    prm.leave_subsection();
  } catch (const ExcParameter &) {
    prm.leave_subsection();
  }
}

template <int dim>
LaplaceProblem<dim>::~LaplaceProblem() {
  if (gmg) gmg_destroy(gmg);
}

template <int dim>
void LaplaceProblem<dim>::gmg_check(int rc, const char *what) {
  if (rc == GMG_OK) return;
  const std::string msg = std::string(what) + ": " + gmg_last_error(gmg);
  if (rc == GMG_ENOCONVERGENCE) throw NoConvergence(msg);
  throw ExcMessage(msg);
}

// =============================================================================== input
// src/step-50.cc:181-258: whitespace-token walk; token #2 = atom count, atoms from token #35
template <int dim>
void LaplaceProblem<dim>::read_lammps_input_file(const std::string &filename) {
  TimerOutput::Scope t(computing_timer, "Read LAMMPS input file");
  std::ifstream file(filename);
  if (dim != 3) {
    lammpsinput = false;
    *pcout << "\nReading of Lammps input file implemented for 3D only\n" << std::endl;
    return;
  }
  if (!file.is_open()) {
    lammpsinput = false;
    *pcout << "Unable to open the file." << std::endl;
    return;
  }
  lammpsinput = true;
  unsigned int count = 0;
  std::string input;
  while (!file.eof()) {
    if (count == 2) {
      file >> number_of_atoms;
      *pcout << "Number of atoms: " << number_of_atoms << std::endl;
      atom_types.resize(number_of_atoms);
      charges.resize(number_of_atoms);
      atom_positions.resize(3 * (size_t)number_of_atoms);
    } else if (count == 35) {
      double a, b;
      for (unsigned int i = 0; i < number_of_atoms; ++i) {
        file >> a >> b >> atom_types[i] >> charges[i] >> atom_positions[3 * i] >> atom_positions[3 * i + 1] >>
            atom_positions[3 * i + 2];
      }
    } else {
      file >> input;
    }
    count++;
  }
}

// src/step-50.cc:1490-1527
template <int dim>
void LaplaceProblem<dim>::make_mesh() {
  if (Problemtype == "Step16") {
    triangulation.reset(new Forest(1, domain_size_left, domain_size_right));
    triangulation->refine_global((int)number_of_global_refinement);
  } else {
    const double a = 2 * mesh_size_h;
    const double N = (domain_size_right - domain_size_left) / a;
    const double M = repetitions_for_vacuum;
    const double repetitions_in_each_direction = 2 * (N + 2 * M);
    const unsigned int reps = (unsigned int)repetitions_in_each_direction;
    const int k = coarse_levels_below_base;
    if (k < 0 || k > 8 || (k > 0 && (reps % (1u << k)) != 0))
      throw ExcMessage("Coarse levels below the base mesh: the base lattice (" + std::to_string(reps) +
                       " cells per direction) is not divisible by 2^" + std::to_string(k));
    // (hi - lo) / (reps / 2^k) / 2^k == (hi - lo) / reps bit for bit: level k IS the reference's base lattice
    triangulation.reset(new Forest((int)(reps >> k), domain_size_left - (M * a), domain_size_right + (M * a)));
    if (k > 0) triangulation->refine_global(k);
  }
}

// =============================================================================== RHS path (device)
namespace {
struct HostTrace {
  const char *name;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  bool on = std::getenv("GMG_TRACE") != nullptr;
  explicit HostTrace(const char *n) : name(n) {}
  ~HostTrace() {
    if (on)
      std::fprintf(stderr, "[step50 trace] %-34s %9.3f ms\n", name,
                   1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
};
}  // namespace
namespace {
ActiveCells flatten(const Forest &f, const DoFs &d, bool with_lists, int base_level) {
  ActiveCells a;
  for (int l = 0; l < f.n_levels(); ++l)
    for (size_t p = 0; p < d.active_cells[l].size(); ++p) {
      const int c = d.active_cells[l][p];
      const Int3 &ijk = f.L[l].ijk[c];
      const double h = f.h(l);
      for (int k = 0; k < 3; ++k) a.lo.push_back(f.lo + ijk[k] * h);
      a.h.push_back(h);
      for (int v = 0; v < NV; ++v) a.dofs.push_back(d.cell_dofs[l][p][v]);
      int anc = c;
      for (int k = l; k > base_level; --k) anc = f.L[k].parent[anc];  // children inherit the parent's list (:441-449)
      a.list.push_back(with_lists ? anc : -1);
    }
  return a;
}
}  // namespace

// src/step-50.cc:260-306: on the base mesh (cycle 0), cell lists atom i iff a vertex is within the cutoff
template <int dim>
void LaplaceProblem<dim>::rhs_assembly_optimization() {
  TimerOutput::Scope t(computing_timer, "RHS assembly optimization");
  const Forest &f = *triangulation;
  const int bl = base_level();  // the cells of the reference's base lattice (level 0 unless coarser levels were added)
  const int n = f.n_cells(bl);
  const double hb = f.h(bl);
  std::vector<double> lo(3 * (size_t)n), h(n, hb);
  for (int c = 0; c < n; ++c)
    for (int k = 0; k < 3; ++k) lo[3 * (size_t)c + k] = f.lo + f.L[bl].ijk[c][k] * hb;
  charges_list_ptr.assign(n + 1, 0);
  gmg_check(gmg_bin_atoms(gmg, n, lo.data(), h.data(), (int)number_of_atoms, atom_positions.data(),
                          nonzero_density_radius_parameter * r_c, charges_list_ptr.data(), nullptr),
            "gmg_bin_atoms");
  charges_list_atoms.assign(std::max<int64_t>(charges_list_ptr[n], 1), 0);
  gmg_check(gmg_bin_atoms(gmg, n, lo.data(), h.data(), (int)number_of_atoms, atom_positions.data(),
                          nonzero_density_radius_parameter * r_c, charges_list_ptr.data(), charges_list_atoms.data()),
            "gmg_bin_atoms");
  charges_list_atoms.resize(charges_list_ptr[n]);
}

// src/step-50.cc:509-575
template <int dim>
void LaplaceProblem<dim>::compute_charge_densities() {
  TimerOutput::Scope t(computing_timer, "Compute charge densities");
  HostTrace tr("compute_charge_densities");
  const auto t0 = std::chrono::steady_clock::now();
  const int nq = (int)(degree + quadrature_degree_rhs);
  std::vector<double> gp, gw;
  gauss_unit(nq, gp, gw);
  std::vector<double> qpts;
  for (int z = 0; z < nq; ++z)
    for (int y = 0; y < nq; ++y)
      for (int x = 0; x < nq; ++x) {
        qpts.push_back(gp[x]);
        qpts.push_back(gp[y]);
        qpts.push_back(gp[z]);
      }
  const ActiveCells &a = *active_cells_cache;
  if (tr.on)
    std::fprintf(stderr, "[step50 trace]   flatten %.3f ms\n",
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  const int nc = (int)a.h.size();
  // density_values_for_each_cell of the reference feeds its host-side assembly (src/step-50.cc:813-816); here both
  // consumers (load vector, residual term of the indicator) read the device-resident copy: nothing comes back
  // (118 MB per call at 64k atoms) unless `Keep densities on the host` asks for it
  double *rho_out = nullptr;
  if (densities_on_host) {
    density_values.resize((size_t)nc * nq * nq * nq);  // (every entry is written by the device call)
    rho_out = density_values.data();
  } else {
    density_values.clear();
  }
  gmg_check(gmg_charge_density(gmg, nc, a.lo.data(), a.h.data(), a.list.data(), nq * nq * nq, qpts.data(), r_c, rho_out),
            "gmg_charge_density");
  if (rec) rec->rhs_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// src/step-50.cc:578-644: only the dipole survives (the quadrupole is overwritten by 0 at :624)
template <int dim>
void LaplaceProblem<dim>::compute_moments() {
  TimerOutput::Scope t(computing_timer, "Compute dipole moments");
  for (int k = 0; k < 3; ++k) dipole_moment[k] = 0.0;
  for (unsigned int i = 0; i < number_of_atoms; ++i)
    for (int k = 0; k < 3; ++k) dipole_moment[k] += charges[i] * atom_positions[3 * i + k];
}

// Dirichlet values g (src/step-50.cc:681-694; include/step_50.h:338-353, 378-385)
template <int dim>
void LaplaceProblem<dim>::boundary_values() {
  const DoFs &d = *mg_dof_handler;
  boundary_g.assign(d.n, 0.0);
  if (Problemtype != "GaussianCharges" || Boundary_conditions == "Homogeneous" || !lammpsinput) return;
  const double inv_constant = 1.0 / (std::sqrt(M_PI) * r_c);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < d.n; ++i) {
    if (!d.dirichlet[i]) continue;
    const auto x = d.coords(i);
    if (Boundary_conditions == "Exact") {
      double v = 0.0;
      for (unsigned int k = 0; k < number_of_atoms; ++k) {
        const double dx = atom_positions[3 * k] - x[0], dy = atom_positions[3 * k + 1] - x[1],
                     dz = atom_positions[3 * k + 2] - x[2];
        const double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        v += (r < 1e-10) ? charges[k] * 2.0 * inv_constant : charges[k] * (std::erf(r / r_c) / r);
      }
      boundary_g[i] = v;
    } else {  // Inhomogeneous: p0.x / |x|^3 about the origin
      const double r = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
      boundary_g[i] = (dipole_moment[0] * x[0] + dipole_moment[1] * x[1] + dipole_moment[2] * x[2]) / std::pow(r, 3);
    }
  }
}

// src/step-50.cc:646-732
template <int dim>
void LaplaceProblem<dim>::setup_system(const unsigned int &cycle) {
  TimerOutput::Scope t(computing_timer, "Setup system");
  mg_dof_handler.reset(new DoFs(*triangulation));
  asm_flags_system.clear();
  asm_flags_level0.clear();
  rhs_constrained.clear();
  hanging_list.clear();
  hanging_list_n = -1;
  rhs_cells_on_device = nullptr;
  active_cells_cache.reset(new ActiveCells(flatten(*triangulation, *mg_dof_handler, flag_rhs_assembly, base_level())));
  solution.assign(mg_dof_handler->n, 0.0);
  system_rhs.assign(mg_dof_handler->n, 0.0);
  if ((cycle == 0) && flag_rhs_assembly && lammpsinput) rhs_assembly_optimization();
  if (lammpsinput) {
    compute_charge_densities();
    compute_moments();
  }
  boundary_values();
}

// src/step-50.cc:735-833: matrix on the host, load vector + constraints on the device
template <int dim>
void LaplaceProblem<dim>::assemble_system() {
  TimerOutput::Scope t(computing_timer, "Assemble system");
  const Forest &f = *triangulation;
  const DoFs &d = *mg_dof_handler;
  Coefficient coef;
  if (Problemtype == "Step16")
    coef = [](double x, double y, double z) { return (x * x + y * y + z * z < 0.5 * 0.5) ? 5.0 : 1.0; };
  if (assemble_on_device()) system_matrix = Csr{};  // built on the device at the hand-over (gmg_assemble_matrix)
  else system_matrix = assemble_system_matrix(f, d, coef);
  assemble_rhs_on_device();
}

// load vector + constraints on the device (the RHS part of assemble_system, src/step-50.cc:798-828)
template <int dim>
void LaplaceProblem<dim>::assemble_rhs_on_device() {
  HostTrace tr("assemble_rhs_on_device");
  const DoFs &d = *mg_dof_handler;
  const auto t0 = std::chrono::steady_clock::now();
  const int nq = (int)(degree + quadrature_degree_rhs), nq3 = nq * nq * nq;
  std::vector<double> gp, gw;
  gauss_unit(nq, gp, gw);
  std::vector<double> shape((size_t)nq3 * NV), weights(nq3);
  for (int z = 0, q = 0; z < nq; ++z)
    for (int y = 0; y < nq; ++y)
      for (int x = 0; x < nq; ++x, ++q) {
        const double p[3] = {gp[x], gp[y], gp[z]};
        weights[q] = gw[x] * gw[y] * gw[z];
        for (int v = 0; v < NV; ++v) {
          double s = 1.0;
          for (int k = 0; k < 3; ++k) s *= vo(v, k) ? p[k] : 1.0 - p[k];
          shape[(size_t)q * NV + v] = s;
        }
      }
  const ActiveCells &a = *active_cells_cache;
  const int nc = (int)a.h.size();
  const double *rho = nullptr;  // densities already on the device from compute_charge_densities
  std::vector<double> rho_host;
  if (!lammpsinput) {
    // rhs_func->value_list (src/step-50.cc:799-803)
    rho_host.resize((size_t)nc * nq3);
    for (int c = 0; c < nc; ++c)
      for (int z = 0, q = 0; z < nq; ++z)
        for (int y = 0; y < nq; ++y)
          for (int x = 0; x < nq; ++x, ++q) {
            if (Problemtype == "Step16") {
              rho_host[(size_t)c * nq3 + q] = 10.0;
            } else {
              const double px = a.lo[3 * c] + a.h[c] * gp[x], py = a.lo[3 * c + 1] + a.h[c] * gp[y],
                           pz = a.lo[3 * c + 2] + a.h[c] * gp[z];
              const double cv = (px * px + py * py + pz * pz) / (r_c * r_c);
              rho_host[(size_t)c * nq3 + q] = (8.0 * std::exp(-4.0 * cv) - std::exp(-cv)) / (std::pow(r_c, 3) * std::pow(M_PI, 1.5));
            }
          }
    rho = rho_host.data();
    density_values = rho_host;
  }
  double Kref[NV][NV];
  unit_stiffness(Kref);
  // derived per mesh (setup_system clears them): resolved inhomogeneities and the constraint flags as bytes
  if ((int)rhs_constrained.size() != d.n) {
    rhs_ghat = resolve_inhomogeneity(d, boundary_g);
    rhs_inhom = false;
    for (double v : rhs_ghat) rhs_inhom |= v != 0.0;
    rhs_constrained.assign(d.constrained.begin(), d.constrained.end());
  }
  const std::vector<double> &ghat = rhs_ghat;
  const bool inhom = rhs_inhom;
  const std::vector<uint8_t> &constrained = rhs_constrained;
  gmg_check(gmg_assemble_rhs(gmg, nc, rho, a.h.data(), a.dofs.data(), nq3, shape.data(), weights.data(),
                             inhom ? &Kref[0][0] : nullptr, inhom ? ghat.data() : nullptr, d.n, d.hang.rowptr.data(),
                             d.hang.col.data(), d.hang.val.data(), constrained.data(), system_rhs.data()),
            "gmg_assemble_rhs");
  rhs_cells_on_device = (const void *)a.dofs.data();
  if (rec) rec->rhs_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// src/step-50.cc:835-933 (+ MGTransferPrebuilt::build_matrices, :957-958)
template <int dim>
void LaplaceProblem<dim>::assemble_multigrid() {
  TimerOutput::Scope t(computing_timer, "Assemble Multigrid");
  Coefficient coef;
  if (Problemtype == "Step16")
    coef = [](double x, double y, double z) { return (x * x + y * y + z * z < 0.5 * 0.5) ? 5.0 : 1.0; };
  // Matrix assembly = Device: level 0 (all but a few percent of the level entries) is built on the device
  mg_ops = assemble_level_operators(*triangulation, *mg_dof_handler, coef, assemble_on_device() ? 1 : 0);
}


template <int dim>
void LaplaceProblem<dim>::hand_over_hierarchy() {
  HostTrace tr_all("hand_over_hierarchy total");
  const DoFs &d = *mg_dof_handler;
  const bool mg = PreconditionerType == "GMG";
  const int nl = mg ? triangulation->n_levels() : 1;
  gmg_check(gmg_set_num_levels(gmg, nl), "gmg_set_num_levels");
  // multi-GPU: rows follow the z-slab of the base lattice their vertex lies in (the reference: p4est
  // subdomains); a system dof and its level-0 twin share coordinates, hence the owner.  Patch levels are replicated.
  int dist_rank = 0, dist_world = 1;
  gmg_dist_rank(gmg, &dist_rank, &dist_world);
  if (dist_world > 1) {
    if (!mg) throw ExcMessage("the multi-GPU path implements the GMG preconditioner");
    if (base_level() > 0) throw ExcMessage("the multi-GPU path partitions level 0 = the base lattice (Coarse levels below the base mesh = 0)");
    const int res = triangulation->resolution();
    const int planes = triangulation->reps + 1;
    auto slab = [&](int z_fine) { return std::min(dist_world - 1, (int)(((long)(z_fine >> res)) * dist_world / planes)); };
    std::vector<int32_t> own_sys(d.n), own_l0(d.level_n[0]);
    for (int i = 0; i < d.n; ++i) own_sys[i] = slab(d.xyz[i][2]);
    for (int i = 0; i < d.level_n[0]; ++i) own_l0[i] = slab(d.level_xyz[0][i][2]);
    gmg_check(gmg_set_ownership(gmg, GMG_SYSTEM, 0, d.n, own_sys.data()), "gmg_set_ownership");
    gmg_check(gmg_set_ownership(gmg, GMG_LEVEL, 0, d.level_n[0], own_l0.data()), "gmg_set_ownership");
  }
  auto set = [&](int which, int l, const Csr &m) {
    gmg_check(gmg_set_matrix(gmg, which, l, m.n_rows, m.n_cols, m.rowptr.data(), m.col.data(), m.val.data()),
              "gmg_set_matrix");
  };
  bool on_device = assemble_on_device();
  if (on_device && dist_world > 1) {
    // row-partitioned matrices are handed over assembled: build what the earlier stages left to the device
    if (system_matrix.rowptr.empty()) system_matrix = assemble_system_matrix(*triangulation, d, Coefficient());
    if (mg_ops.A[0].rowptr.empty()) mg_ops = assemble_level_operators(*triangulation, d, Coefficient(), 0);
    on_device = false;
  }
  double Kref[NV][NV];
  unit_stiffness(Kref);
  if (on_device) {
    HostTrace tr("  gmg_assemble_matrix (system)");
    const ActiveCells &a = *active_cells_cache;
    std::vector<uint8_t> &flags = asm_flags_system;  // per mesh (setup_system clears them)
    if ((int)flags.size() != d.n) {
      flags.resize(d.n);
      for (int i = 0; i < d.n; ++i) flags[i] = (uint8_t)(d.hanging[i] ? 2 : (d.dirichlet[i] ? 1 : 0));
    }
    // the active cells (dofs, edge lengths) are still on the device from assemble_rhs_on_device() of this cycle
    const bool resident = rhs_cells_on_device == (const void *)a.dofs.data();
    gmg_check(gmg_assemble_matrix(gmg, GMG_SYSTEM, 0, d.n, (int64_t)a.h.size(), resident ? nullptr : a.dofs.data(),
                                  resident ? nullptr : a.h.data(), 0.0, flags.data(), d.hang.rowptr.data(), d.hang.col.data(),
                                  d.hang.val.data(), &Kref[0][0]),
              "gmg_assemble_matrix");
  } else {
    set(GMG_SYSTEM, 0, system_matrix);
  }
  if (mg) {
    for (int l = 0; l < nl; ++l) {
      if (l == 0 && on_device) {
        HostTrace tr("  gmg_assemble_matrix (level 0)");
        const int n0 = d.level_n[0];
        std::vector<uint8_t> &flags = asm_flags_level0;
        if ((int)flags.size() != n0) {
          flags.resize(n0);
          for (int i = 0; i < n0; ++i) flags[i] = (uint8_t)((d.level_edge[0][i] || d.level_boundary[0][i]) ? 1 : 0);
        }
        static_assert(sizeof(Dofs8) == 8 * sizeof(int32_t), "cell dofs are handed over as int32[8]");
        gmg_check(gmg_assemble_matrix(gmg, GMG_LEVEL, 0, n0, (int64_t)triangulation->n_cells(0),
                                      d.level_cell_dofs[0].data()->data(), nullptr, triangulation->h(0), flags.data(),
                                      nullptr, nullptr, nullptr, &Kref[0][0]),
                  "gmg_assemble_matrix");
      } else {
        set(GMG_LEVEL, l, mg_ops.A[l]);
      }
      if (l >= 1) set(GMG_EDGE, l, mg_ops.I[l]);
      if (l + 1 < nl) set(GMG_PROLONG, l, mg_ops.P[l]);
      gmg_check(gmg_set_copy_indices(gmg, l, (int)d.copy_global[l].size(), d.copy_global[l].data(),
                                     d.copy_level[l].data()),
                "gmg_set_copy_indices");
      if (l >= 1) {
        // vertex-parity colouring of the level: neighbours in a Q1 stencil differ in at least one coordinate by one
        const int sh = triangulation->resolution() - l;
        std::vector<int32_t> color(d.level_n[l]);
        for (int i = 0; i < d.level_n[l]; ++i) {
          const Int3 &q = d.level_xyz[l][i];
          color[i] = ((q[0] >> sh) & 1) | (((q[1] >> sh) & 1) << 1) | (((q[2] >> sh) & 1) << 2);
        }
        gmg_check(gmg_set_level_coloring(gmg, l, d.level_n[l], color.data()), "gmg_set_level_coloring");
      }
    }
  } else {
    set(GMG_LEVEL, 0, system_matrix);  // unused by the Jacobi-preconditioned solve
    std::vector<int32_t> id(d.n);
    for (int i = 0; i < d.n; ++i) id[i] = i;
    gmg_check(gmg_set_copy_indices(gmg, 0, d.n, id.data(), id.data()), "gmg_set_copy_indices");
  }
  gmg_check(gmg_set_smoother(gmg, smoother_kind, smoother_omega, smoothing_steps), "gmg_set_smoother");
  gmg_check(gmg_set_coarse(gmg, 1000, 1e-10), "gmg_set_coarse");
  HostTrace tr_setup("  gmg_setup (host view)");
  gmg_check(gmg_setup(gmg), "gmg_setup");
}

// src/step-50.cc:938-1017
template <int dim>
void LaplaceProblem<dim>::solve() {
  TimerOutput::Scope t(computing_timer, "Solve");
  const auto t0 = std::chrono::steady_clock::now();
  std::ostream &out = *pcout;
  const bool trace = std::getenv("GMG_TRACE") != nullptr;
  hand_over_hierarchy();
  if (trace)
    std::fprintf(stderr, "[step50 trace] hand_over_hierarchy %.3f ms\n",
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  double bn[3], mn[3], sn[3];
  gmg_check(gmg_vector_norms(gmg, (int64_t)system_rhs.size(), system_rhs.data(), bn), "gmg_vector_norms");
  gmg_check(gmg_matrix_norms(gmg, GMG_SYSTEM, 0, mn), "gmg_matrix_norms");
  out << "   L1 rhs norm " << std::setprecision(10) << std::scientific << bn[0] << std::endl;
  out << "   L2 rhs norm " << std::setprecision(10) << std::scientific << bn[1] << std::endl;
  out << "   LInfinity rhs norm " << std::setprecision(10) << std::scientific << bn[2] << std::endl;
  out << "   L1 Matrix norm " << std::setprecision(10) << std::scientific << mn[0] << std::endl;
  out << "   LInfinity Matrix norm " << std::setprecision(10) << std::scientific << mn[1] << std::endl;
  out << "   Frobenius Matrix norm " << std::setprecision(10) << std::scientific << mn[2] << std::endl;
  const double tol = 1e-8 * bn[1];  // SolverControl(500, 1e-8 * system_rhs.l2_norm())
  int its = 0;
  double res0 = 0.0, res = 0.0;
  int rc;
  if (PreconditionerType == "GMG")
    rc = gmg_pcg_solve(gmg, system_rhs.data(), solution.data(), 500, tol, &its, &res0, &res);
  else
    rc = gmg_pcg_solve_jacobi(gmg, system_rhs.data(), solution.data(), 0.6, 500, tol, &its, &res0, &res);
  gmg_check(rc, "solver.solve");
  if (trace)
    std::fprintf(stderr, "[step50 trace] through pcg %.3f ms\n",
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
 gmg_check(gmg_vector_norms(gmg, (int64_t)solution.size(), solution.data(), sn), "gmg_vector_norms");
  if (trace)
    std::fprintf(stderr, "[step50 trace] through solution norms %.3f ms\n",
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  out << "   Starting value " << std::fixed << res0 << std::endl;
  out << "   CG converged in " << its << " iterations." << std::endl;
  out << "   Convergence value " << std::scientific << res << std::endl;
  out << "   L1 solution norm " << std::setprecision(10) << std::scientific << sn[0] << std::endl;
  out << "   L2 solution norm " << std::setprecision(10) << std::scientific << sn[1] << std::endl;
  out << "   LInfinity solution norm " << std::setprecision(10) << std::scientific << sn[2] << std::endl;
  if (rec) {
    rec->rhs_l1 = bn[0]; rec->rhs_l2 = bn[1]; rec->rhs_linf = bn[2];
    rec->mat_l1 = mn[0]; rec->mat_linf = mn[1]; rec->mat_frob = mn[2];
    rec->start = res0; rec->its = its; rec->conv = res;
    rec->sol_l1 = sn[0]; rec->sol_l2 = sn[1]; rec->sol_linf = sn[2];
    int32_t buf[4096];
    int n = 0;
    gmg_last_coarse_iterations(gmg, buf, 4096, &n);
    rec->coarse_its.assign(buf, buf + std::min(n, 4096));
    rec->solve_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  if (trace)
    std::fprintf(stderr, "[step50 trace] through records %.3f ms\n",
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  {
    // constraints.distribute(solution) = ministep::distribute on a copy, fused: one parallel pass writes the copy with
    // the Dirichlet values in place, the hanging dofs follow in ascending order (sequential: same sums, same bits)
    const DoFs &d = *mg_dof_handler;
    if (hanging_list_n != d.n) {  // (setup_system resets it for every new mesh)
      hanging_list.clear();
      for (int i = 0; i < d.n; ++i)
        if (d.hanging[i]) hanging_list.push_back(i);
      hanging_list_n = d.n;
    }
    distributed_solution.resize(solution.size());
    const double *src = solution.data(), *gv = boundary_g.data();
    double *dst = distributed_solution.data();
    const char *dir = d.dirichlet.data();
#pragma omp parallel for schedule(static)
    for (int i = 0; i < d.n; ++i) dst[i] = dir[i] ? gv[i] : src[i];
    for (int i : hanging_list) {
      double sum = 0.0;
      for (int64_t e = d.hang.rowptr[i]; e < d.hang.rowptr[i + 1]; ++e) sum += d.hang.val[e] * dst[d.hang.col[e]];
      dst[i] = sum;
    }
  }
  if (trace)
    std::fprintf(stderr, "[step50 trace] through distribute %.3f ms\n",
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
}

// src/step-50.cc:1020-1090
template <int dim>
void LaplaceProblem<dim>::estimate_error_and_mark_cells() {
  TimerOutput::Scope t(computing_timer, "Estimate error and mark cells");
  const int nq = (int)(degree + quadrature_degree_rhs);
  // face jumps + residual term on the device (gmg_error_indicator); the host keeps the mesh and provides the face
  // topology.  The float32 indicators are bit-identical to ministep's error_indicator (tests/test_gpu_indicator.py).
  {
    const Forest &f = *triangulation;
    const DoFs &d = *mg_dof_handler;
    const IndicatorTopology topo = indicator_topology(f, d);
    const int nc = (int)(topo.face_nb.size() / 6);
    std::vector<double> gp2, gw2;
    gauss_unit(2, gp2, gw2);
    std::vector<float> eta((size_t)nc);
    float eta_max = 0.0f;
    const bool with_rho = indicator_with_residual && (lammpsinput || !density_values.empty());
    // GaussianCharges with atoms: the densities of compute_charge_densities are still on the device
    const double *rho = (with_rho && !lammpsinput) ? density_values.data() : nullptr;
    (void)nq;
    gmg_check(gmg_error_indicator(gmg, nc, topo.face_nb.data(), topo.face_kind.data(), (int)(topo.hang_children.size() / 4),
                                  topo.hang_children.data(), distributed_solution.data(), (int)distributed_solution.size(), rho,
                                  with_rho ? 1 : 0, gp2.data(), gw2.data(), eta.data(), &eta_max),
              "gmg_error_indicator");
    error_per_cell.assign(f.n_levels(), std::vector<float>());
    size_t off = 0;
    for (int l = 0; l < f.n_levels(); ++l) {
      error_per_cell[l].assign(eta.begin() + off, eta.begin() + off + d.active_cells[l].size());
      off += d.active_cells[l].size();
    }
  }
  // marking on the device too (gmg_mark_cells: float indicator widened to double >= 0.6 * max, src/step-50.cc:1084-1090);
  // the host only scatters the per-active-cell flags into the forest's per-level flag arrays
  double threshold = 0.0;
  {
    const Forest &f = *triangulation;
    const DoFs &d = *mg_dof_handler;
    size_t nc = 0;
    for (int l = 0; l < f.n_levels(); ++l) nc += d.active_cells[l].size();
    std::vector<uint8_t> marked(nc);
    gmg_check(gmg_mark_cells(gmg, (int)nc, 0.6, marked.data(), &threshold), "gmg_mark_cells");
    refine_flags.assign(f.n_levels(), {});
    size_t off = 0;
    for (int l = 0; l < f.n_levels(); ++l) {
      refine_flags[l].assign(f.n_cells(l), 0);
      for (size_t p = 0; p < d.active_cells[l].size(); ++p) refine_flags[l][d.active_cells[l][p]] = (char)marked[off + p];
      off += d.active_cells[l].size();
    }
  }
  *pcout << "Threshold value for refinement:\t" << threshold << std::endl;
  if (rec) {
    rec->threshold = threshold;
    for (auto &fl : refine_flags)
      for (char c : fl) rec->n_flagged += c;
  }
}

// src/step-50.cc:1095-1121
template <int dim>
void LaplaceProblem<dim>::refine_grid(const unsigned int &cycle) {
  TimerOutput::Scope t(computing_timer, "Refine, solution transfer and sending atoms list to child cells");
  std::unique_ptr<DoFs> old_dofs = std::move(mg_dof_handler);
  const int old_res = triangulation->resolution();
  const std::vector<double> previous_solution = distributed_solution;
  triangulation->refine(refine_flags);
  setup_system(cycle);
  if (zero_initial_guess) {  // `Initial guess = Zero`
    solution.assign(mg_dof_handler->n, 0.0);
    return;
  }
  // SolutionTransfer::interpolate + set_zero on the device (gmg_transfer_solution): the host provides the index tables
  // (which old dof is which new dof, the 27 dofs of every refined cell), the values never pass through host arithmetic
  const TransferTables T = transfer_tables(old_res, *old_dofs, *triangulation, *mg_dof_handler);
  const DoFs &d = *mg_dof_handler;
  std::vector<uint8_t> constrained(d.constrained.begin(), d.constrained.end());
  solution.assign(d.n, 0.0);
  gmg_check(gmg_transfer_solution(gmg, old_dofs->n, previous_solution.data(), d.n, (int)T.copy_old.size(), T.copy_old.data(),
                                  T.copy_new.data(), (int)T.pass_ptr.size() - 1, T.pass_ptr.data(), T.parent_dofs.data(),
                                  constrained.data(), solution.data()),
            "gmg_transfer_solution");
}

// VTU / PVTU / VisIt output (src/step-50.cc:1149-1308) is visualisation, out of scope of this path
template <int dim>
void LaplaceProblem<dim>::output_results(const unsigned int) const {}

template <int dim>
double LaplaceProblem<dim>::long_ranged_potential(const double p[3], const double a[3], const double &charge) const {
  const double r = std::sqrt((p[0] - a[0]) * (p[0] - a[0]) + (p[1] - a[1]) * (p[1] - a[1]) + (p[2] - a[2]) * (p[2] - a[2]));
  return charge * (std::erf(r / r_c) / r);
}
template <int dim>
double LaplaceProblem<dim>::short_ranged_potential(const double p[3], const double a[3], const double &charge) const {
  const double r = std::sqrt((p[0] - a[0]) * (p[0] - a[0]) + (p[1] - a[1]) * (p[1] - a[1]) + (p[2] - a[2]) * (p[2] - a[2]));
  return charge * (std::erfc(r / r_c) / r);
}

// src/step-50.cc:1310-1420
template <int dim>
void LaplaceProblem<dim>::postprocess_electrostatic_energy() {
  TimerOutput::Scope t(computing_timer, "Postprocess electrostatic energy");
  const long n = number_of_atoms;
  // pair sums on the device (O(N^2): 2e9 pairs at 64k atoms)
  double pair[2] = {0.0, 0.0};
  gmg_check(gmg_pair_energies(gmg, r_c, pair), "gmg_pair_energies");
  const double analytical_energy = pair[0], short_ranged = pair[1];
  // FE part: phi_h(X_i) evaluated on the device in the active cell around each atom
  std::vector<int32_t> cd(8 * (size_t)n);
  std::vector<double> xi(3 * (size_t)n), phi(n);
  for (long i = 0; i < n; ++i) {
    int lev, cell;
    locate(*triangulation, *mg_dof_handler, &atom_positions[3 * i], lev, cell, &xi[3 * i]);
    const int pos = mg_dof_handler->active_pos[lev][cell];
    for (int v = 0; v < NV; ++v) cd[8 * i + v] = mg_dof_handler->cell_dofs[lev][pos][v];
  }
  gmg_check(gmg_point_values(gmg, (int)n, cd.data(), xi.data(), distributed_solution.data(),
                             (int)distributed_solution.size(), phi.data()),
            "gmg_point_values");
  double fe = 0.0, self_energy = 0.0;
  for (long i = 0; i < n; ++i) {
    fe += 0.5 * charges[i] * phi[i];
    self_energy += charges[i] * charges[i] / (std::sqrt(M_PI) * r_c);
  }
  const double total = short_ranged + fe - self_energy;
  std::ostream &out = *pcout;
  out << "\nTotal analytical electrostatic energy :   " << analytical_energy << std::endl;
  out << "Short-ranged energy contribution :  " << short_ranged << std::endl;
  out << "FE solution long-ranged energy contribution :    " << fe << std::endl;
  out << "Self energy contribution : " << self_energy << std::endl;
  out << "Total electrostatic energy with split in short- and long-ranged : " << total << std::endl;
  out << "Absolute Error between both energies :\t" << std::abs(std::abs(analytical_energy) - std::abs(total)) << "\n"
      << std::endl;
  out << "Relative Error in total electrostatic energy :\t"
      << std::abs((std::abs(analytical_energy) - std::abs(total)) / analytical_energy) << std::endl;
  if (rec) {
    rec->have_energy = true;
    rec->e_analytic = analytical_energy; rec->e_short = short_ranged; rec->e_fe = fe;
    rec->e_self = self_energy; rec->e_total = total;
  }
}

// src/step-50.cc:1423-1461: O(cells * 8 * atoms), on the device (gmg_energy_norm_error) over the cells already resident
// for the RHS.  Without atoms the reference's exact_solution is absent (Step16): nothing is printed there either.
template <int dim>
void LaplaceProblem<dim>::postprocess_error_in_energy_norm() {
  TimerOutput::Scope t(computing_timer, "Postprocess FE error");
  if (!lammpsinput) return;
  std::vector<double> gp, gw;
  gauss_unit(2, gp, gw);
  double err = 0.0;
  gmg_check(gmg_energy_norm_error(gmg, distributed_solution.data(), (int)distributed_solution.size(), r_c, gp.data(),
                                  gw.data(), &err),
            "gmg_energy_norm_error");
  *pcout << "Error in FE solution in energy norm:  " << err << std::endl;
  if (rec) rec->energy_norm_error = err;
}

// =============================================================================== run (src/step-50.cc:1463-1573)
// One process per GPU: exchange the CUDA IPC handles of the ranks' communication buffers (host all-gather = control
// plane), map the peers.  From here on gmg_set_ownership / gmg_setup partition the system matrix and level 0.
template <int dim>
void LaplaceProblem<dim>::connect_ranks() {
  if (!ranks || ranks->world == 1) return;
  const char *cb = std::getenv("GMG_COMM_BYTES");
  const int64_t comm_bytes = cb ? std::atoll(cb) : (int64_t)512 << 20;
  std::vector<char> mine(64), all(64 * (size_t)ranks->world);
  gmg_check(gmg_dist_init(gmg, ranks->rank, ranks->world, comm_bytes, mine.data()), "gmg_dist_init");
  ranks->all_gather(mine.data(), all.data(), 64);
  gmg_check(gmg_dist_connect(gmg, all.data()), "gmg_dist_connect");
  ranks->barrier();
}

template <int dim>
void LaplaceProblem<dim>::begin_run() {
  if (connect_in_run) {
    ranks.reset(new Rendezvous());
    if (ranks->rank != 0) pcout = &null_out;  // ConditionalOStream pcout(std::cout, this_mpi_process == 0)
  }
  std::ostream &out = *pcout;
  out << "Problem type is:   " << Problemtype << std::endl;
  out << "Preconditioner :    " << PreconditionerType << std::endl;
  if (flag_rhs_assembly)
    out << "Rhs assembly optimization ENABLED" << std::endl;
  else
    out << "Without rhs assembly optimization" << std::endl;
  // the reference: "Running with Trilinos on N MPI rank(s)..." (src/step-50.cc:1476-1482)
  out << "Running with B200 (sm_100a CUDA) on " << (ranks ? ranks->world : 1) << " GPU(s)..." << std::endl;
  if (dim != 3) throw ExcMessage("Only dim = 3 is implemented on the B200 path.");
  if (gmg_create(gpu_device + (ranks ? ranks->local_rank : 0), &gmg) != GMG_OK)
    throw ExcMessage("gmg_create failed: no B200 (sm_100) CUDA device; this path has no CPU fallback.");
  connect_ranks();
  computing_timer.reset();
  run_start = std::chrono::steady_clock::now();
  out << "Dimension:\t" << dim << std::endl;
  read_lammps_input_file(LammpsInputFilename);
  if (lammpsinput)
    gmg_check(gmg_set_atoms(gmg, (int)number_of_atoms, atom_positions.data(), charges.data()), "gmg_set_atoms");
  cycle_records.clear();
  cycle_records.reserve(number_of_adaptive_refinement_cycles);
}

template <int dim>
void LaplaceProblem<dim>::cycle_until_solve(const unsigned int cycle) {
  std::ostream &out = *pcout;
  cycle_records.emplace_back();
  rec = &cycle_records.back();
  out << "Cycle " << cycle << ':' << std::endl;
  if (cycle == 0)
    make_mesh();
  else
    refine_grid(cycle);
  rec->n_active_cells = (long)triangulation->n_active_cells();
  out << "   Number of active cells:       " << rec->n_active_cells << std::endl;
  if (cycle == 0) setup_system(cycle);
  rec->n_dofs = mg_dof_handler->n;
  out << "   Number of degrees of freedom: " << mg_dof_handler->n << " (by level: ";
  for (int level = 0; level < triangulation->n_levels(); ++level) {
    rec->n_dofs_level.push_back(mg_dof_handler->level_n[level]);
    out << mg_dof_handler->level_n[level] << (level == triangulation->n_levels() - 1 ? ")" : ", ");
  }
  out << std::endl;
  assemble_system();
  if (PreconditionerType == "GMG") assemble_multigrid();
}

template <int dim>
void LaplaceProblem<dim>::cycle_after_solve(const unsigned int cycle) {
  estimate_error_and_mark_cells();
  output_results(cycle);
  // src/step-50.cc:1553-1555: the energies only below 300 atoms, the energy-norm error every cycle
  if (lammpsinput && number_of_atoms < energy_atom_limit) postprocess_electrostatic_energy();
  if (lammpsinput && (energy_norm_atom_limit == 0 || number_of_atoms < energy_norm_atom_limit))
    postprocess_error_in_energy_norm();
}

template <int dim>
void LaplaceProblem<dim>::end_run() {
  std::ostream &out = *pcout;
  rec = nullptr;
  if (ranks) ranks->barrier();  // no rank unmaps its buffers while a peer may still store into them
  if (flag_output_time) computing_timer.print_summary(out);
  if (flag_output_time)
    out << "   \nTotal Elapsed wall time for solution: "
        << std::chrono::duration<double>(std::chrono::steady_clock::now() - run_start).count() << " seconds.\n"
        << std::endl;
}

template <int dim>
void LaplaceProblem<dim>::run() {
  connect_in_run = true;
  begin_run();
  for (unsigned int cycle = 0; cycle < number_of_adaptive_refinement_cycles; ++cycle) {
    cycle_until_solve(cycle);
    solve();
    cycle_after_solve(cycle);
  }
  end_run();
}

template class LaplaceProblem<2>;
template class LaplaceProblem<3>;

}  // namespace Step50

// =============================================================================== main (src/main.cc:6-121)
namespace {
template <class F>
void run_with(ParameterHandler &prm, std::ostream &out, F &&after) {
  using namespace Step50;
  prm.enter_subsection("Geometry");
  unsigned int number_of_global_refinement = prm.get_integer("Number of global refinement");
  double domain_size_left = prm.get_double("Domain limit left");
  double domain_size_right = prm.get_double("Domain limit right");
  double mesh_size_h = prm.get_double("Mesh size");
  unsigned int repetitions_for_vacuum = prm.get_integer("Vacuum repetitions");
  prm.leave_subsection();
  prm.enter_subsection("Misc");
  unsigned int number_of_adaptive_refinement_cycles = prm.get_integer("Number of Adaptive Refinement");
  double r_c = prm.get_double("smoothing length");
  double nonzero_density_radius_parameter = prm.get_double("Nonzero Density radius parameter around each charge");
  bool flag_analytical_solution = prm.get_bool("Output and calculation of Analytical solution");
  bool flag_rhs_field = prm.get_bool("Output of RHS field");
  bool flag_atoms_support = prm.get_bool("Output of support of each atom");
  bool flag_rhs_assembly = prm.get_bool("Flag for RHS evaluation optimization");
  const unsigned int quadrature_degree_rhs = prm.get_integer("Quadrature points for RHS function");
  const bool flag_output_time = prm.get_bool("Output time summary table");
  prm.leave_subsection();
  const unsigned int Degree = prm.get_integer("Polynomial degree");
  prm.enter_subsection("Solver input data");
  std::string PreconditionerType = prm.get("Preconditioner");
  prm.leave_subsection();
  prm.enter_subsection("Problem Selection");
  std::string Problemtype = prm.get("Problem");
  const unsigned int d = prm.get_integer("Dimension");
  const std::string Boundary_conditions = prm.get("Boundary conditions selection");
  prm.leave_subsection();
  prm.enter_subsection("Lammps data");
  std::string LammpsInputFile = prm.get("Lammps input file");
  prm.leave_subsection();
  if (d == 3) {
    LaplaceProblem<3> laplace_problem(Degree, prm, Problemtype, PreconditionerType, LammpsInputFile, Boundary_conditions,
                                      domain_size_left, domain_size_right, mesh_size_h, repetitions_for_vacuum,
                                      number_of_global_refinement, number_of_adaptive_refinement_cycles, r_c,
                                      nonzero_density_radius_parameter, flag_rhs_assembly, flag_analytical_solution,
                                      flag_rhs_field, flag_atoms_support, flag_output_time, quadrature_degree_rhs);
    laplace_problem.set_output(out);
    laplace_problem.run();
    after(laplace_problem.records());
  } else if (d == 2) {
    LaplaceProblem<2> laplace_problem(Degree, prm, Problemtype, PreconditionerType, LammpsInputFile, Boundary_conditions,
                                      domain_size_left, domain_size_right, mesh_size_h, repetitions_for_vacuum,
                                      number_of_global_refinement, number_of_adaptive_refinement_cycles, r_c,
                                      nonzero_density_radius_parameter, flag_rhs_assembly, flag_analytical_solution,
                                      flag_rhs_field, flag_atoms_support, flag_output_time, quadrature_degree_rhs);
    laplace_problem.set_output(out);
    laplace_problem.run();
  } else {
    throw ExcMessage("Only 2 and 3 dimensions are supported.");
  }
}
}  // namespace

void step50_run_from_string(const std::string &prm_text, std::ostream &out, std::vector<Step50::CycleRecord> *records) {
  ParameterHandler prm;
  ParameterReader param(prm);
  param.declare_parameters();
  prm.parse_input_from_string(prm_text.c_str());
  run_with(prm, out, [&](const std::vector<Step50::CycleRecord> &r) {
    if (records) *records = r;
  });
}

// `main -np N file.prm`: what `mpirun -np N main file.prm` is for the reference, without an MPI installation: fork one
// process per GPU (before anything touches CUDA), give each its RANK / LOCAL_RANK / WORLD_SIZE and a free rendezvous port.
// Returns the rank of this process; `children` holds the pids rank 0 has to wait for.
static int fork_ranks(int n, std::vector<pid_t> &children) {
  int port = 0;
  {
    const int fd = ::socket(AF_INET, SOCK_STREAM, 0);
    sockaddr_in sa{};
    sa.sin_family = AF_INET;
    sa.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
    socklen_t len = sizeof(sa);
    if (fd >= 0 && ::bind(fd, (sockaddr *)&sa, sizeof(sa)) == 0 && ::getsockname(fd, (sockaddr *)&sa, &len) == 0)
      port = ntohs(sa.sin_port);
    if (fd >= 0) ::close(fd);
  }
  if (port == 0) throw Step50::ExcMessage("main -np: no free TCP port for the rendezvous");
  ::setenv("MASTER_ADDR", "127.0.0.1", 1);
  ::setenv("GMG_RENDEZVOUS_PORT", std::to_string(port).c_str(), 1);
  ::setenv("WORLD_SIZE", std::to_string(n).c_str(), 1);
  if (!std::getenv("OMP_NUM_THREADS")) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    ::setenv("OMP_NUM_THREADS", std::to_string(std::max(1u, hw / (unsigned)n)).c_str(), 1);
  }
  int rank = 0;
  for (int r = 1; r < n; ++r) {
    const pid_t pid = ::fork();
    if (pid < 0) throw Step50::ExcMessage("main -np: fork failed");
    if (pid == 0) {
      rank = r;
      children.clear();
      break;
    }
    children.push_back(pid);
  }
  ::setenv("RANK", std::to_string(rank).c_str(), 1);
  ::setenv("LOCAL_RANK", std::to_string(rank).c_str(), 1);
  return rank;
}

int step50_main(int argc, char **argv, std::ostream &out) {
  std::vector<pid_t> children;
  int rank = 0;
  try {
    if (argc <= 1) throw Step50::ExcMessage("Invalid inputs. \nCall this program as <./main para_filename.prm>");
    int file_arg = 1;
    if (std::string(argv[1]) == "-np") {
      if (argc <= 3 || std::atoi(argv[2]) < 1)
        throw Step50::ExcMessage("Invalid inputs. \nCall this program as <./main -np N para_filename.prm>");
      file_arg = 3;
      if (std::atoi(argv[2]) > 1) rank = fork_ranks(std::atoi(argv[2]), children);
    }
    ParameterHandler prm;
    ParameterReader param(prm);
    param.declare_parameters();
    param.read_parameters(argv[file_arg]);
    run_with(prm, out, [](const std::vector<Step50::CycleRecord> &) {});
    int failed = 0;
    for (pid_t pid : children) {
      int status = 0;
      if (::waitpid(pid, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) ++failed;
    }
    if (failed) throw Step50::ExcMessage(std::to_string(failed) + " rank(s) of main -np failed");
    if (rank != 0) {  // a forked rank: leave without running the parent's exit handlers twice
      std::cout.flush();
      ::_exit(0);
    }
  } catch (std::exception &exc) {
    std::cerr << std::endl
              << std::endl
              << "----------------------------------------------------" << std::endl;
    std::cerr << "Exception on processing: " << std::endl
              << exc.what() << std::endl
              << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    if (rank != 0) ::_exit(1);
    for (pid_t pid : children) ::kill(pid, SIGTERM);
    throw;
  }
  return 0;
}
