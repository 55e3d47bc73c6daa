"""Device-side matrix assembly (gmg_assemble_matrix), CPU part: the row routines of csrc/assemble_row.h -- the code the
CUDA kernels run, one thread per row -- executed sequentially on the host (ms_assemble_emulate) reproduce the host
assembly of the system matrix (hanging nodes + Dirichlet constraints condensed, src/step-50.cc:771-795) and of every
level matrix (boundary / refinement-edge dofs eliminated, :855-889): same row pointer, same columns, same value BITS."""
import numpy as np
import pytest

import hostlib


def random_forest(seed, reps, cycles, frac):
    rng = np.random.default_rng(seed)
    M = hostlib.Ministep(reps, -1.0, 2.0)
    for _ in range(cycles):
        flags = []
        for l in range(M.n_levels):
            child0 = M.get("child0", l)
            flags.append(((rng.random(len(child0)) < frac) & (child0 < 0)).astype(np.uint8))
        M.refine(flags)
    M.build()
    return M


@pytest.mark.parametrize("seed,reps,cycles,frac", [(1, 4, 3, 0.15), (2, 6, 2, 0.3), (3, 3, 4, 0.1), (4, 8, 2, 0.05),
                                                   (5, 5, 0, 0.0)])
def test_row_gather_assembly_is_bit_identical_to_the_cell_loop(seed, reps, cycles, frac):
    M = random_forest(seed, reps, cycles, frac)
    assert (M.get("hanging").sum() > 0) == (cycles > 0)
    diff, longest = M.assemble_emulate(0)
    assert diff == 0
    assert longest >= 27
    for l in range(M.n_levels):
        diff, longest = M.assemble_emulate(1, l)
        assert diff == 0 and longest == 27


def test_assembly_inputs_helper_matches_the_host_structures():
    """hostapi.assembly_inputs (what the GPU test hands to gmg_assemble_matrix) is consistent with the mesh."""
    M = random_forest(7, 4, 2, 0.2)
    n, dofs, h, uh, flags, hang = hostlib.assembly_inputs(M, 0)
    assert n == len(M.get("hanging")) and dofs.shape[1] == 8 and len(h) == len(dofs)
    assert dofs.min() == 0 and dofs.max() == n - 1
    assert set(np.unique(flags)) <= {0, 1, 2} and hang is not None and len(hang[0]) == n + 1
    # every hanging dof has 2 or 4 parents with weights 1/2 or 1/4, none of them hanging
    lens = np.diff(hang[0])
    assert set(np.unique(lens[flags == 2])) <= {2, 4} and (lens[flags != 2] == 0).all()
    assert (flags[hang[1]] != 2).all()
    n0, dofs0, h0, uh0, flags0, hang0 = hostlib.assembly_inputs(M, 1, 0)
    assert h0 is None and hang0 is None and uh0 == 3.0 / 4 and len(flags0) == n0 == 5 ** 3
    assert flags0.sum() == 5 ** 3 - 3 ** 3  # the boundary dofs of the base lattice
    K = hostlib.unit_stiffness()
    assert np.allclose(K, K.T) and abs(K.sum()) < 1e-14 and abs(K[0, 0] - 1.0 / 3.0) < 1e-15


def test_matrix_assembly_prm_key_is_declared_with_a_selection_pattern():
    from conftest import make_prm
    ok = "subsection Solver input data\n set Matrix assembly = Device\nend\n"
    hostlib.check_prm(make_prm(extra=ok))
    with pytest.raises(hostlib.HostError):
        hostlib.check_prm(make_prm(extra=ok.replace("Device", "Elsewhere")))
