"""CPU guard for the structure the coarse-grid CG's window kernel is compiled for (csrc/pattern_win2.cuh, XPAIR): the
dominant row of the level-0 matrix of a Q1 lattice -- as the host substrate numbers it (first touch, explicit zeros kept)
-- is 9 runs of 3 consecutive columns with the diagonal in the middle of run 4, and with an odd number of vertices per
line the runs start at odd, even, odd, ... offsets (XP_ODD = 0x155: the window plan keeps the parity of the offsets).
If the numbering or the sparsity ever changes, `xpair_supported()` falls back to the one-row-per-lane loop silently; this
test says so loudly."""
from collections import Counter

import pytest

import hostlib


@pytest.mark.parametrize("reps", [12, 20])
def test_dominant_row_of_a_lattice_level0_matrix(reps):
    M = hostlib.Ministep(reps, -1.5, 2.5)
    M.build()
    rp, col, val = M.csr("A", 0)
    n = len(rp) - 1
    nx = reps + 1
    assert n == nx ** 3 and nx % 2 == 1
    patterns = Counter()
    for r in range(n):
        patterns[tuple(int(c) - r for c in col[rp[r]:rp[r + 1]])] += 1
    offsets, count = patterns.most_common(1)[0]
    assert count >= (nx - 4) ** 3  # at least every vertex two layers away from the boundary (its columns are eliminated)
    assert len(offsets) == 27
    runs = [offsets[3 * k:3 * k + 3] for k in range(9)]
    for k, (a, b, c) in enumerate(runs):
        assert b == a + 1 and c == a + 2, "not a run of three consecutive columns"
        assert (a % 2 == 1) == bool((0x155 >> k) & 1), "parity of the run start"
    assert runs[4] == (-1, 0, 1)
    assert [r[1] for r in runs] == [dz * nx * nx + dy * nx for dz in (-1, 0, 1) for dy in (-1, 0, 1)]
    # the pattern dictionary keys on offsets AND value bits: the most frequent such pattern has these offsets too
    keyed = Counter()
    for r in range(n):
        if rp[r + 1] - rp[r] == 27:
            keyed[(tuple(int(c) - r for c in col[rp[r]:rp[r + 1]]), val[rp[r]:rp[r + 1]].tobytes())] += 1
    (dom_offsets, _), dom_count = keyed.most_common(1)[0]
    assert dom_offsets == offsets and dom_count > n // 4
