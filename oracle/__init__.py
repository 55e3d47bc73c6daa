"""CPU oracle for the MG-PCG + Gaussian-charge RHS hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of what the
reference (`/root/reference/src/step-50.cc`, `include/step_50.h`) computes on the
hot path, including the deal.II 9.0 / Trilinos (Epetra, Ifpack) semantics that the
reference delegates to libraries which are absent from `/root/reference`
(deal.II pinned only as ">= 9.0.0", CMakeLists.txt:29; Trilinos unpinned,
CMakeLists.txt:43-53).  It may be imported only by `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` -- never by the product path under
`geometric-multigrid-preconditioners-for-long-range-coulomb-interaction_b200/`.

Parity status: PINNED against the reference's own golden stdout files
(`tests/gaussian-charges.mpirun=1.output`, `tests_3D/*.output`, `tests_2D/*.output`,
`tests/test_with_optimal_parameters.mpirun=1.output`, the cluster logs under
`Cluster runs output and postprocessing/`).  The numbers are transcribed in
`tests/golden/reference_goldens.json` with their file:line.  Quantities that the
reference never prints (DoF numbering, sparsity entries, per-cell atom lists,
individual vector entries) are pinned only indirectly through those norms and
iteration counts; see DESIGN.md "Oracle".

The real reference cannot be compiled here (needs deal.II+Trilinos+p4est+MPI,
none installed, no network), so there is no `oracle/_ref`.
"""
