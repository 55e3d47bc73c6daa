"""Oracle DoF handling for Q1 elements: numbering, hanging-node and Dirichlet constraints,
refinement-edge sets, copy indices.

Restates deal.II semantics used by the reference at src/step-50.cc:646-732
(`distribute_dofs`, `distribute_mg_dofs`, `make_hanging_node_constraints`,
`interpolate_boundary_values`, `MGConstrainedDoFs::initialize`,
`make_zero_boundary_constraints`) and by `MGTransferPrebuilt::build_matrices`
(src/step-50.cc:957-958; copy indices).  On one rank deal.II numbers DoFs
first-touch while walking active cells in (level, index) order and, per level, all
cells of that level in index order; Q1 DoF i of a cell sits on vertex i
(SURVEY.md Appendix A2).

TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import itertools

import numpy as np


def first_touch(keys):
    """Number the distinct values of `keys` (cells x vertices) in order of first appearance."""
    flat = keys.ravel()
    uniq, first = np.unique(flat, return_index=True)
    order = np.argsort(first, kind="stable")
    ids = np.empty(len(uniq), dtype=np.int64)
    ids[order] = np.arange(len(uniq), dtype=np.int64)
    pos = np.searchsorted(uniq, flat)
    return ids[pos].reshape(keys.shape), uniq[order]


class DoFs:
    """Active ("global") and per-level Q1 DoF numbering on a Forest."""

    def __init__(self, forest):
        f = self.forest = forest
        self.res = f.resolution()
        nl = f.n_levels
        # ---- active numbering (distribute_dofs)
        self.active_cells = [f.active(l) for l in range(nl)]
        keys = [f.vertex_key(f.vertex_coords(l, self.active_cells[l])) for l in range(nl)]
        allkeys = np.concatenate(keys, axis=0)
        dofs, self.key = first_touch(allkeys)
        self.n = len(self.key)
        split = np.cumsum([len(k) for k in keys])[:-1]
        self.cell_dofs = np.split(dofs, split, axis=0)  # per level: (n_active_l, 2^dim)
        self.xyz = f.key_to_xyz(self.key)  # integer vertex coordinates of each dof
        # ---- level numbering (distribute_mg_dofs)
        self.level_cell_dofs, self.level_key, self.level_n, self.level_xyz = [], [], [], []
        for l in range(nl):
            k = f.vertex_key(f.vertex_coords(l))
            d, key = first_touch(k)
            self.level_cell_dofs.append(d)
            self.level_key.append(key)
            self.level_n.append(len(key))
            self.level_xyz.append(f.key_to_xyz(key))
        self._boundary()
        self._hanging()
        self._refinement_edges()
        self._copy_indices()

    # ---------------------------------------------------------------- boundary
    def _on_boundary(self, xyz):
        n = self.forest.points_per_axis()
        return np.any((xyz == 0) | (xyz == n - 1), axis=-1)

    def _boundary(self):
        self.boundary = self._on_boundary(self.xyz)
        self.level_boundary = [self._on_boundary(x) for x in self.level_xyz]

    def lookup_dof(self, key):
        order = getattr(self, "_key_order", None)
        if order is None:
            order = self._key_order = np.argsort(self.key, kind="stable")
            self._key_sorted = self.key[order]
        pos = np.minimum(np.searchsorted(self._key_sorted, key), self.n - 1)
        found = self._key_sorted[pos] == key
        return np.where(found, order[pos], -1)

    # ----------------------------------------------------------- hanging nodes
    def _hanging(self):
        """make_hanging_node_constraints: every vertex in the interior of a face (or edge)
        of an active cell whose neighbour across that face is refined is constrained to the
        Q1 interpolant of the face corners (edge midpoint = 1/2,1/2; face centre = 4 x 1/4)."""
        f = self.forest
        dim = f.dim
        rows, cols, vals = [], [], []
        tang_pts = [t for t in itertools.product((0, 1, 2), repeat=dim - 1) if any(c == 1 for c in t)]
        for l in range(f.n_levels):
            act = self.active_cells[l]
            if len(act) == 0:
                continue
            half = 1 << (self.res - l - 1) if self.res > l else 0
            for a in range(dim):
                others = [d for d in range(dim) if d != a]
                for side in (0, 1):
                    nb = f.ijk[l][act].copy()
                    nb[:, a] += 1 if side else -1
                    nidx = f.lookup(l, nb)
                    refined = (nidx >= 0)
                    refined[refined] = f.child0[l][nidx[refined]] >= 0
                    cells = act[refined]
                    if len(cells) == 0:
                        continue
                    assert half > 0
                    base = f.ijk[l][cells] << (self.res - l)  # lower corner, fine coords
                    base = base.copy()
                    base[:, a] += (2 * half) if side else 0
                    for t in tang_pts:
                        p = base.copy()
                        for d, td in zip(others, t):
                            p[:, d] += td * half
                        hdof = self.lookup_dof(f.vertex_key(p))
                        assert (hdof >= 0).all()
                        halves = [d for d, td in zip(others, t) if td == 1]
                        w = 1.0 / (1 << len(halves))
                        for sg in itertools.product((-1, 1), repeat=len(halves)):
                            q = p.copy()
                            for d, s in zip(halves, sg):
                                q[:, d] += s * half
                            pdof = self.lookup_dof(f.vertex_key(q))
                            assert (pdof >= 0).all()
                            rows.append(hdof)
                            cols.append(pdof)
                            vals.append(np.full(len(hdof), w))
        if rows:
            rows = np.concatenate(rows)
            cols = np.concatenate(cols)
            vals = np.concatenate(vals)
            pair = rows * self.n + cols
            _, keep = np.unique(pair, return_index=True)
            rows, cols, vals = rows[keep], cols[keep], vals[keep]
        else:
            rows = cols = np.zeros(0, dtype=np.int64)
            vals = np.zeros(0)
        self.hang_rows, self.hang_cols, self.hang_vals = rows, cols, vals
        self.hanging = np.zeros(self.n, dtype=bool)
        self.hanging[rows] = True
        # with 2:1 balance across corners a hanging node never depends on another hanging node
        assert not self.hanging[cols].any(), "chained hanging-node constraints"
        # interpolate_boundary_values skips DoFs that are already constrained (src/step-50.cc:692-694)
        self.dirichlet = self.boundary & ~self.hanging
        self.constrained = self.hanging | self.dirichlet

    # -------------------------------------------------------- refinement edges
    def _refinement_edges(self):
        """MGConstrainedDoFs refinement-edge indices: level-l DoFs on faces of level-l cells whose
        neighbour across the face is coarser (no level-l cell there, but inside the domain)."""
        f = self.forest
        dim = f.dim
        self.level_edge = []
        for l in range(f.n_levels):
            edge = np.zeros(self.level_n[l], dtype=bool)
            if l > 0:
                n = f.cells_per_axis(l)
                for a in range(dim):
                    for side in (0, 1):
                        nb = f.ijk[l].copy()
                        nb[:, a] += 1 if side else -1
                        inside = (nb[:, a] >= 0) & (nb[:, a] < n)
                        coarser = inside & (f.lookup(l, nb) < 0)
                        vs = [v for v in range(1 << dim) if f.VO[v, a] == side]
                        edge[self.level_cell_dofs[l][coarser][:, vs].ravel()] = True
            self.level_edge.append(edge)

    # ------------------------------------------------------------ copy indices
    def _copy_indices(self):
        """MGLevelGlobalTransfer copy_indices[l]: (global, level) pairs over the DoFs of active
        level-l cells that are not on the refinement edge of level l."""
        self.copy_global, self.copy_level = [], []
        for l in range(self.forest.n_levels):
            g = self.cell_dofs[l].ravel()
            lv = self.level_cell_dofs[l][self.active_cells[l]].ravel()
            keep = ~self.level_edge[l][lv]
            g, lv = g[keep], lv[keep]
            _, first = np.unique(lv, return_index=True)
            self.copy_global.append(g[first])
            self.copy_level.append(lv[first])

    # ----------------------------------------------------------------- helpers
    def real_coords(self):
        return self.forest.real_coords(self.xyz)

    def level_real_coords(self, l):
        return self.forest.real_coords(self.level_xyz[l])
