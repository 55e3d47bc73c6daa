/* gmg_b200.h -- C ABI of the B200-native MG-PCG + Gaussian-charge RHS path.
 *
 * The reference (vinayak-gholap1993/Geometric-Multigrid-preconditioners-for-long-range-Coulomb-
 * interaction) has no plugin/FFI layer: its hot path is `LaplaceProblem<dim>::solve()`
 * (src/step-50.cc:938-1017) wiring deal.II / Trilinos objects together, fed by
 * `rhs_assembly_optimization()` (:260-306), `compute_charge_densities()` (:509-575) and the
 * load-vector part of `assemble_system()` (:798-828).  The entry points below are exactly the
 * operations that path performs on data the host hands over (assembled CSR matrices, transfer
 * matrices, copy indices, atoms, cell lists); each cites the reference construct it replaces.
 *
 * Conventions: every function returns 0 on success or a negative GMG_E* code; the message is
 * available from gmg_last_error().  All pointers are HOST memory borrowed for the duration of the
 * call unless the name ends in `_dev` / the function says "device".  One caller thread per handle.
 * Indices are 32-bit (deal.II `types::global_dof_index` is `unsigned int` in the reference build),
 * row pointers 64-bit, values fp64.  There is no CPU fallback: without a CUDA device gmg_create
 * fails with GMG_ENODEVICE.
 */
#ifndef GMG_B200_H
#define GMG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmg_context *gmg_handle;

enum {
  GMG_OK = 0,
  GMG_EINVAL = -1,        /* bad argument / call order */
  GMG_ENODEVICE = -2,     /* no CUDA device, or CUDA runtime error at creation */
  GMG_ECUDA = -3,         /* CUDA runtime error */
  GMG_ENOCONVERGENCE = -4,/* SolverControl::NoConvergence (outer > max_it, or coarse > its limit) */
  GMG_ENCCL = -5
};

/* which matrix */
enum {
  GMG_SYSTEM = 0,  /* system_matrix              (include/step_50.h:156)                         */
  GMG_LEVEL = 1,   /* mg_matrices[level]         (include/step_50.h:168, src/step-50.cc:888-889) */
  GMG_EDGE = 2,    /* mg_interface_matrices[l]   (include/step_50.h:169, src/step-50.cc:896-925) */
  GMG_PROLONG = 3  /* MGTransferPrebuilt prolongation level -> level+1 (src/step-50.cc:957-958)  */
};

/* smoother kinds (src/step-50.cc:969-973 selects the smoother at source level) */
enum {
  GMG_SMOOTHER_JACOBI = 0,      /* LA::MPI::PreconditionJacobi(omega), `steps` steps               */
  GMG_SMOOTHER_CHEBYSHEV = 1,   /* Chebyshev-accelerated Jacobi of degree `steps`                  */
  GMG_SMOOTHER_MC_SSOR = 2,     /* multicolour SSOR(omega): stands in for Ifpack's sequential SSOR */
  GMG_SMOOTHER_LEX_SSOR = 3     /* Ifpack's lexicographic SSOR(omega) itself, level-scheduled      */
};

/* ---- lifecycle ------------------------------------------------------------------------------ */
int gmg_create(int device, gmg_handle *out);
int gmg_destroy(gmg_handle h);
const char *gmg_last_error(gmg_handle h);
/* Launch all kernels of this handle on an existing CUDA stream (a `cudaStream_t` passed as void*),
 * e.g. torch's current stream, so that the caller's CUDA events bracket the work. */
int gmg_set_stream(gmg_handle h, void *cuda_stream);
int gmg_synchronize(gmg_handle h);
/* library build info: returns the sm arch the kernels were compiled for (100) */
int gmg_compiled_arch(void);

/* ---- hierarchy hand-over (replaces the deal.II objects `solve()` wires together) ------------- */
int gmg_set_num_levels(gmg_handle h, int n_levels);
/* CSR, sorted or unsorted columns, explicit zeros allowed (Epetra keeps the pattern's zeros). */
int gmg_set_matrix(gmg_handle h, int which, int level, int32_t n_rows, int32_t n_cols,
                   const int64_t *rowptr, const int32_t *col, const double *val);
/* Device-side assembly of the Q1 Laplace matrix instead of handing it over assembled: replaces the
 * matrix part of assemble_system (src/step-50.cc:771-795: cell matrices condensed through
 * constraints.distribute_local_to_global into the pattern of make_sparsity_pattern(dof, dsp, constraints, true),
 * :699-701) for which = GMG_SYSTEM, and the level-0 loop of assemble_multigrid (:855-889: boundary / refinement-edge
 * dofs eliminated) for which = GMG_LEVEL, level = 0.  Patch levels (>= 1) are small and stay with gmg_set_matrix.
 *   cell_dofs[n_cells][8]  dofs of the cells in the order of the reference's cell loop (vertex v: bit d = offset along d);
 *                          which = GMG_SYSTEM: cell_dofs = cell_h = NULL takes the cell arrays the last gmg_assemble_rhs
 *                          call left on the device (the same active cells in the same order: n_cells must agree)
 *   cell_h[n_cells]        edge length per cell, or NULL: every cell has uniform_h (cell matrix = h * k_ref)
 *   row_flags[n_rows]      bit 0: eliminated row and column (Dirichlet dof, level boundary / refinement-edge dof);
 *                          bit 1: hanging dof, constraint line in the CSR (hang_rowptr[n_rows + 1], hang_col, hang_val;
 *                          at most 7 parents per line, no chains); hang_* may be NULL when nothing hangs
 *   k_ref[64]              cell matrix of the unit cube, row-major (the host's bits)
 * The result is the CSR the sequential assembly produces: same pattern (explicit zeros kept), same value bits (every
 * entry is summed in the order of the cell loop, unfused).  Afterwards gmg_setup as usual. */
int gmg_assemble_matrix(gmg_handle h, int which, int level, int32_t n_rows, int64_t n_cells,
                        const int32_t *cell_dofs, const double *cell_h, double uniform_h,
                        const uint8_t *row_flags, const int64_t *hang_rowptr, const int32_t *hang_col,
                        const double *hang_val, const double *k_ref);
/* Raw CSR of a matrix handed over (gmg_set_matrix) or assembled (gmg_assemble_matrix) but not yet consumed by
 * gmg_setup: *nnz always; rowptr[n_rows + 1] / col[nnz] / val[nnz] when non-NULL (call twice: sizes, then arrays).
 * which = GMG_SYSTEM or GMG_LEVEL (level 0).  For tests and debugging. */
int gmg_raw_matrix_get(gmg_handle h, int which, int level, int64_t *nnz, int64_t *rowptr, int32_t *col, double *val);
/* MGLevelGlobalTransfer::copy_indices[level]: (global dof, level dof) pairs. */
int gmg_set_copy_indices(gmg_handle h, int level, int32_t n, const int32_t *global_idx,
                         const int32_t *level_idx);
/* MGSmootherPrecondition<...>::initialize(mg_matrices, AdditionalData(omega)); set_steps(steps). */
int gmg_set_smoother(gmg_handle h, int kind, double omega, int steps);
/* Optional colouring of a level's rows for the multicolour SSOR (e.g. the 8 vertex-parity colours of a Q1
 * mesh level); validated against the matrix graph, the library falls back to its greedy colouring. */
int gmg_set_level_coloring(gmg_handle h, int level, int32_t n, const int32_t *color);
/* Run a whole smoothing call of the (multicolour / level-scheduled) SSOR as one cooperative kernel with grid-wide
 * barriers between colours instead of one launch per colour.  Default off: on a B200 the graph-replayed per-colour
 * launches are faster (54.0 vs 60.4 ms per 64k-atom step); kept as an option. */
int gmg_set_persistent_smoother(gmg_handle h, int on);
/* Replay the fine-level parts of the V-cycle as CUDA graphs (default on). */
int gmg_set_graphs(gmg_handle h, int on);
/* SolverControl coarse_solver_control(max_it, abs_tol) + SolverCG + PreconditionIdentity. */
int gmg_set_coarse(gmg_handle h, int max_it, double abs_tol);
/* Drop stored entries with |a_ij| <= drop_tol when building the device format (default: keep all). */
int gmg_set_drop_tolerance(gmg_handle h, double drop_tol);
/* Lossless device formats of the coarse-level matrix (results are bit-identical in all three):
 *   mode 0  plain sliced ELL, 12 bytes per entry;
 *   mode 1  CSELL: 16-bit value-dictionary code + 16-bit column offset, 4 bytes per entry (falls back to 0 when a
 *           matrix has more than 60000 distinct values or a half bandwidth >= 32768);
 *   mode 2  (default) row-pattern dictionary: one 32-bit pattern id per ROW + the table of distinct
 *           (column offset, value) rows -- an FE matrix on a uniform level has a few dozen; falls back to mode 1
 *           when the rows are not repetitive enough (more than n/8 or 20000 distinct rows). */
int gmg_set_compression(gmg_handle h, int mode);
/* Build device formats (sliced ELL), transposes, colourings, eigenvalue bounds. */
int gmg_setup(gmg_handle h);

/* ---- the solve path --------------------------------------------------------------------------- */
/* solver.solve(system_matrix, solution, system_rhs, PreconditionMG) -- SolverCG recurrences,
 * stop when ||g||_2 <= abs_tol, fail (GMG_ENOCONVERGENCE) at max_it.  x_inout: initial guess in,
 * solution out.  res0 = "Starting value", res_final = "Convergence value". */
int gmg_pcg_solve(gmg_handle h, const double *b, double *x_inout, int max_it, double abs_tol,
                  int *iters, double *res0, double *res_final);
/* PreconditionerType == "Jacobi": PCG with omega * D^-1 (src/step-50.cc:996-1006). */
int gmg_pcg_solve_jacobi(gmg_handle h, const double *b, double *x_inout, double omega, int max_it,
                         double abs_tol, int *iters, double *res0, double *res_final);
/* PreconditionMG::vmult: dst = M^-1 src (one V-cycle incl. copy_to_mg / copy_from_mg). */
int gmg_vcycle_apply(gmg_handle h, const double *src, double *dst);
/* matrix.vmult */
int gmg_spmv(gmg_handle h, int which, int level, const double *x, double *y);
/* MGCoarseGridIterativeSolver / plain SolverCG with identity preconditioner from x = 0. */
int gmg_cg_solve(gmg_handle h, int which, int level, const double *b, double *x, int max_it,
                 double abs_tol, int *iters, double *res_final);
/* one application of the level smoother: u <- smooth(u, rhs) (`steps` steps), zero_start != 0
 * ignores u on input. */
int gmg_smooth(gmg_handle h, int level, const double *rhs, double *u_inout, int zero_start);
/* system_matrix.l1_norm / linfty_norm / frobenius_norm (src/step-50.cc:950-952) -> out[3] */
int gmg_matrix_norms(gmg_handle h, int which, int level, double out[3]);
/* vector l1 / l2 / linfty norms (src/step-50.cc:946-948, 1012-1014) -> out[3] */
int gmg_vector_norms(gmg_handle h, int64_t n, const double *v, double out[3]);
/* inner coarse-CG iteration counts of the V-cycles of the last gmg_pcg_solve (at most cap). */
int gmg_last_coarse_iterations(gmg_handle h, int32_t *out, int cap, int *n_out);

/* ---- device-resident variants (bench `value` leg: inputs already in HBM) ---------------------- */
int gmg_vec_alloc(gmg_handle h, int64_t n, double **dev_out);
int gmg_vec_free(gmg_handle h, double *dev);
int gmg_vec_upload(gmg_handle h, double *dev, const double *host, int64_t n);
int gmg_vec_download(gmg_handle h, double *host, const double *dev, int64_t n);
int gmg_vec_copy_dev(gmg_handle h, double *dst_dev, const double *src_dev, int64_t n);
/* bytes moved host->device / device->host by the host-pointer entry points since the last reset */
int gmg_transfer_bytes(gmg_handle h, int reset, int64_t *h2d, int64_t *d2h);
int gmg_pcg_solve_dev(gmg_handle h, const double *b_dev, double *x_dev, int max_it, double abs_tol,
                      int *iters, double *res0, double *res_final);
int gmg_vcycle_apply_dev(gmg_handle h, const double *src_dev, double *dst_dev);
int gmg_spmv_dev(gmg_handle h, int which, int level, const double *x_dev, double *y_dev);
int gmg_cg_solve_dev(gmg_handle h, int which, int level, const double *b_dev, double *x_dev,
                     int max_it, double abs_tol, int *iters, double *res_final);
/* bytes one SpMV / one coarse-CG iteration with this matrix moves algorithmically:
 * out[0] = stored nnz, out[1] = SpMV bytes and out[2] = CG-iteration bytes of the format actually stored,
 * out[3], out[4] = the same for plain CSR (12 B per entry), out[5] = format in use (0, 1, 2 as in gmg_set_compression) */
int gmg_matrix_traffic(gmg_handle h, int which, int level, double out[6]);
/* Developer probe: per-phase time of the windowed coarse-CG kernel as seen by thread 0 of block (block_plus_1 - 1)
 * (globaltimer ns, summed over iterations): out[0] SpMV tail + block reduction, [1] barrier + grid reduction,
 * [2] x/g update, [3] barrier + reduction, [4] new direction, [5] barrier, [6] iterations; inside the SpMV, per tile:
 * [8] issue of the next window + pattern ids, [9] wait for the window, [10] dominant loop, [11] other rows + stores,
 * [12] block barrier, [13] remainder rows.  Returns the counters since the last call, resets them and selects the
 * block (0 switches timing off). */
int gmg_debug_cg_phases(gmg_handle h, int block_plus_1, double out_ns[16]);
/* Developer probe: device time (ms, CUDA events) of the three parts of the V-cycles since the last call: out[0] down
 * sweep (copy_to_mg, pre-smoothing, residuals, restrictions), out[1] coarse solve, out[2] up sweep (prolongations,
 * post-smoothing, copy_from_mg), out[3] number of V-cycles; resets the counters and switches the timing on/off. */
int gmg_debug_vcycle_profile(gmg_handle h, int enable, double out_ms[4]);
/* ... and per block (256 slots each): time in the SpMV, update and direction phases while timing was on. */
int gmg_debug_cg_blocks(gmg_handle h, double out_ns[768]);
/* Which persistent CG kernel gmg_cg_solve / the coarse solve runs on this matrix: 0 plain SELL, 1 CSELL,
 * 2 row patterns with L1 gathers (also the multi-GPU kernel), 3 row patterns with TMA-filled shared-memory windows,
 * 4 the same with the row codes read from global memory (more than ~35 k rows per SM); 5 / 6 the second-generation
 * window kernel (tagged-word grid reductions, g in registers) with h = A d of a block's rows in shared / global memory,
 * 7 / 8 the same with two consecutive rows per lane in the dominant loop (the default on Q1 lattices). */
int gmg_coarse_kernel(gmg_handle h, int which, int level, int *kernel);
/* accumulated device time (ms, CUDA events on the handle's stream) and launch count of the
 * persistent coarse-CG kernel since the last reset; inner iterations summed in *iters. */
int gmg_coarse_profile(gmg_handle h, int reset, double *ms, int64_t *launches, int64_t *iters);
/* number of kernel launches issued by this handle since creation */
int64_t gmg_launch_count(gmg_handle h);

/* ---- multi-GPU (one process per GPU) -------------------------------------------------
 * The reference distributes rows over MPI ranks (p4est subdomains; Epetra Import/Export halos,
 * MPI_Allreduce dots).  Here every rank maps every peer's communication buffer (CUDA IPC over NVLink):
 *   gmg_dist_init   allocate this rank's buffer, return its 64-byte IPC handle
 *   (caller all-gathers the handles, e.g. torch.distributed / MPI)
 *   gmg_dist_connect  map the peers
 *   gmg_set_ownership owner rank of every row of the system matrix / of level 0 (patch levels >= 1 are
 *                   replicated); then gmg_set_matrix / gmg_set_copy_indices / gmg_setup as usual with the
 *                   GLOBAL matrices on every rank; gmg_pcg_solve(_dev) takes and returns global vectors. */
int gmg_dist_init(gmg_handle h, int rank, int world, int64_t comm_bytes, void *ipc_handle_out /*64 B*/);
int gmg_dist_connect(gmg_handle h, const void *all_handles /*world x 64 B*/);
int gmg_set_ownership(gmg_handle h, int which, int level, int32_t n, const int32_t *owner);
int gmg_dist_rank(gmg_handle h, int *rank, int *world);
/* host-only probe of the row partitioner (CPU tests of the N > 1 logic); outputs malloc'ed, gmg_free_host */
int gmg_partition_probe(int rank, int world, int32_t n_rows, const int64_t *rowptr, const int32_t *col,
                        const double *val, const int32_t *owner, int32_t *n_owned, int32_t *n_halo,
                        int64_t **l_rowptr, int32_t **l_col, double **l_val, int32_t **owned_global,
                        int32_t **halo_global, int32_t **send_count, int32_t **send_idx, int32_t **send_dst_base);
void gmg_free_host(void *p);
/* developer probe: NVLink round-trip latency (us) of tagged 16-byte words between ranks 0 and 1 */
int gmg_dist_pingpong(gmg_handle h, int iters, int mode, double *us_per_round_trip);

/* ---- the RHS path ----------------------------------------------------------------------------- */
/* rhs_assembly_optimization (src/step-50.cc:260-306): cell c lists atom i iff some vertex v of the
 * axis-aligned cube [lo_c, lo_c + h_c]^3 has ||X_i - v||_2 < radius (strict).  Output CSR with
 * ascending atom indices per cell.  Call with atoms_out == NULL to obtain rowptr_out (n_cells+1)
 * and the total count, then again with a buffer of rowptr_out[n_cells] entries. */
int gmg_bin_atoms(gmg_handle h, int32_t n_cells, const double *cell_lo /*[n_cells][3]*/,
                  const double *cell_h /*[n_cells]*/, int32_t n_atoms, const double *pos /*[n][3]*/,
                  double radius, int64_t *rowptr_out, int32_t *atoms_out);
/* Keep per-list atom indices on the device for gmg_charge_density; children inherit the parent's
 * list (src/step-50.cc:441-449) by pointing at the same list id. */
int gmg_set_atom_lists(gmg_handle h, int32_t n_lists, const int64_t *rowptr, const int32_t *atoms);
int gmg_set_atoms(gmg_handle h, int32_t n_atoms, const double *pos /*[n][3]*/, const double *charge);
/* compute_charge_densities (src/step-50.cc:509-575): rho[c][q] = sum_k C exp(-|X_k - x_q|^2/r_c^2) q_k,
 * x_q = lo_c + h_c * qpoint[q]; list_of_cell[c] < 0 sums over all atoms (flag off). */
int gmg_charge_density(gmg_handle h, int32_t n_cells, const double *cell_lo, const double *cell_h,
                       const int32_t *list_of_cell, int32_t n_q, const double *qpoints /*[n_q][3] unit cell*/,
                       double r_c, double *rho_out /*[n_cells][n_q]*/);
/* load vector + constraints (src/step-50.cc:813-828): for every cell, cell_rhs(i) = sum_q shape[q][i]
 * rho[c][q] w[q] h_c^3 - h_c * sum_j Kref[i][j] ghat[dof_j]; distributed into b with hanging-node
 * weights (constraint CSR over dofs; rows of unconstrained dofs empty); dofs flagged in
 * `constrained` end at 0.  ghat may be NULL (homogeneous). */
int gmg_assemble_rhs(gmg_handle h, int32_t n_cells, const double *rho, const double *cell_h,
                     const int32_t *cell_dofs /*[n_cells][8]*/, int32_t n_q, const double *shape /*[n_q][8]*/,
                     const double *weights /*[n_q]*/, const double *Kref /*[8][8] or NULL*/,
                     const double *ghat /*[n_dofs] or NULL*/, int32_t n_dofs,
                     const int64_t *hang_rowptr, const int32_t *hang_col, const double *hang_val,
                     const uint8_t *constrained, double *b_out);
/* phi_h(X_i) for the energy (src/step-50.cc:1353-1366): trilinear evaluation in given cells. */
int gmg_point_values(gmg_handle h, int32_t n_points, const int32_t *cell_dofs /*[n][8]*/,
                     const double *ref_coords /*[n][3]*/, const double *u, int32_t n_dofs,
                     double *phi_out);
/* Error indicator of estimate_error_and_mark_cells (src/step-50.cc:1020-1090): eta_K = float(sqrt(h_K * sum over the
 * interior faces of K of int [d_n u_h]^2 + h_K^2 * int_K (4 pi rho)^2)), h_K = cell diameter, face rule QGauss<2>(2),
 * Vector<float> accumulation as KellyErrorEstimator does.  The active cells are those of the last gmg_assemble_rhs
 * call (their edges, dofs and quadrature weights are still on the device).  Face topology from the host's mesh
 * (face = 2 * axis + side):
 *   face_nb[c][face] < 0     no contribution (Dirichlet boundary face)
 *   face_kind & 3 == 0       face_nb = active neighbour of the same level
 *   face_kind & 3 == 1       c is the fine side of a hanging face: face_nb = the coarse neighbour; bits 2 and 3 of
 *                            face_kind = position of the subface in the coarse face along the two tangential axes
 *   face_kind & 3 == 2       c is the coarse side: face_nb = row of hang_children, the four fine neighbours in the
 *                            order the reference visits them (ascending active index)
 * u = solution after constraints.distribute.  rho: n_cells x n_q densities, or NULL to use those of the last
 * gmg_charge_density call (device-resident); ignored unless residual_term.  eta_out[n_cells] (float32, bit-identical
 * to the sequential restatement), *max_out = max eta (the refinement threshold is 0.6 * max, :1084). */
int gmg_error_indicator(gmg_handle h, int32_t n_cells, const int32_t *face_nb /*[n_cells][6]*/,
                        const uint8_t *face_kind /*[n_cells][6]*/, int32_t n_hang, const int32_t *hang_children /*[n_hang][4]*/,
                        const double *u, int32_t n_dofs, const double *rho, int residual_term,
                        const double gauss2_points[2], const double gauss2_weights[2], float *eta_out, float *max_out);
/* Marking (src/step-50.cc:1084-1090): flags_out[c] = 1 where the indicator of the last gmg_error_indicator call reaches
 * fraction * max (the reference: 0.6 * linfty_norm, refine_and_coarsen with a single threshold and no coarsening); the
 * float32 indicators are widened to double for the comparison, as Vector<float> entries are.  *threshold_out = the
 * printed "Threshold value for refinement". */
int gmg_mark_cells(gmg_handle h, int32_t n_cells, double fraction, uint8_t *flags_out, double *threshold_out);
/* SolutionTransfer::interpolate + constraints.set_zero (src/step-50.cc:1110-1119) after a refinement-only step: the
 * new vector takes the old values at the dofs both meshes share (copy_old[k] -> copy_new[k]); pass by pass (coarse to
 * fine) every refined cell whose 8 corners are known gives its 19 edge / face / centre points (parent_dofs: 27 new dofs
 * per refined cell, index t0 + 3 t1 + 9 t2, -1: none) their trilinear values, the first cell in the hand-over order
 * winning a shared point, sums over the corners in vertex order with unfused arithmetic -- the bits of the sequential
 * host loop; constrained dofs end at 0.  GMG_EINVAL when a dof is left without a value. */
int gmg_transfer_solution(gmg_handle h, int32_t n_old, const double *u_old, int32_t n_new, int32_t n_copy,
                          const int32_t *copy_old, const int32_t *copy_new, int32_t n_pass, const int64_t *pass_ptr,
                          const int32_t *parent_dofs /*[pass_ptr[n_pass]][27]*/, const uint8_t *constrained /*[n_new]*/,
                          double *u_new_out);
/* The O(N^2) pair sums of postprocess_electrostatic_energy (src/step-50.cc:1316-1332) over the atoms of gmg_set_atoms:
 * out[0] = sum_{i<j} q_i q_j / r_ij, out[1] = sum_{i<j} q_i q_j erfc(r_ij / r_c) / r_ij. */
int gmg_pair_energies(gmg_handle h, double r_c, double out[2]);
/* postprocess_error_in_energy_norm (src/step-50.cc:1423-1461): sqrt(int |grad u_h - grad u_exact|^2) with QGauss<3>(2),
 * grad u_exact summed over all atoms of gmg_set_atoms (:1437-1452 via exact_solution->gradient_list).  Cells = those of the
 * last gmg_charge_density / gmg_assemble_rhs calls (resident); u = solution after constraints.distribute. */
int gmg_energy_norm_error(gmg_handle h, const double *u, int32_t n_dofs, double r_c, const double gauss2_points[2],
                          const double gauss2_weights[2], double *out);
/* fused device-resident RHS step used by the bench `value` leg: densities + load vector with the
 * inputs of the last gmg_charge_density / gmg_assemble_rhs calls kept on the device. */
int gmg_rhs_step_dev(gmg_handle h, double *b_dev);

#ifdef __cplusplus
}
#endif
#endif /* GMG_B200_H */
