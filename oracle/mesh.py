"""Oracle mesh: a forest of quad/octrees on a structured base lattice.

Restates what the reference obtains from deal.II's
`parallel::distributed::Triangulation` with `limit_level_difference_at_vertices`
and `construct_multigrid_hierarchy` (src/step-50.cc:120-122):

* base mesh `subdivided_hyper_rectangle(reps, lo, hi, colorize=false)`
  (src/step-50.cc:1504-1526) or `hyper_cube` + `refine_global`
  (src/step-50.cc:1496-1497): level-0 cells lexicographic, x fastest;
* isotropic refinement, children of a parent stored consecutively in deal.II
  child order (child c has offset bit d of c along axis d), new children
  appended on level l+1 in parent-index order (no coarsening ever happens,
  `GridRefinement::refine` only sets refine flags, src/step-50.cc:1089);
* p4est 2:1 balance across faces, edges and corners (deal.II calls
  `p4est_balance(..., P{4,8}EST_CONNECT_FULL)`).

TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import itertools

import numpy as np


def vertex_offsets(dim):
    """deal.II GeometryInfo<dim>: vertex/child v has offset bit d of v along axis d."""
    return np.array([[(v >> d) & 1 for d in range(dim)] for v in range(1 << dim)], dtype=np.int64)


class Forest:
    def __init__(self, reps, lo, hi, dim=3):
        self.dim = dim
        self.reps = int(reps)
        self.lo = float(lo)
        self.hi = float(hi)
        self.H = (self.hi - self.lo) / self.reps  # level-0 cell edge
        self.VO = vertex_offsets(dim)
        ax = np.arange(self.reps, dtype=np.int64)
        grids = np.meshgrid(*([ax] * dim), indexing="ij")
        # lexicographic, x fastest: index = i + reps*(j + reps*k)
        ijk = np.stack([g.ravel(order="F") for g in grids], axis=1)
        self.ijk = [ijk]
        self.parent = [np.full(len(ijk), -1, dtype=np.int64)]
        self.child0 = [np.full(len(ijk), -1, dtype=np.int64)]
        self._sorted = [None]

    # ------------------------------------------------------------------ basics
    @property
    def n_levels(self):
        return len(self.ijk)

    def cells_per_axis(self, l):
        return self.reps << l

    def h(self, l):
        return self.H / (1 << l)

    def n_cells(self, l):
        return len(self.ijk[l])

    def active_mask(self, l):
        return self.child0[l] < 0

    def active(self, l):
        return np.nonzero(self.child0[l] < 0)[0]

    def n_active_cells(self):
        return int(sum(self.active_mask(l).sum() for l in range(self.n_levels)))

    def _key(self, l, ijk):
        n = self.cells_per_axis(l)
        key = ijk[..., 0].astype(np.int64).copy()
        mult = n
        for d in range(1, self.dim):
            key += ijk[..., d] * mult
            mult *= n
        return key

    def _rebuild_index(self, l):
        if l == 0:
            return
        key = self._key(l, self.ijk[l])
        order = np.argsort(key, kind="stable")
        self._sorted[l] = (key[order], order)

    def lookup(self, l, ijk):
        """Index of the level-l cell at integer position ijk, -1 if there is none."""
        ijk = np.asarray(ijk, dtype=np.int64)
        n = self.cells_per_axis(l)
        inside = np.all((ijk >= 0) & (ijk < n), axis=-1)
        key = self._key(l, np.where(inside[..., None], ijk, 0))
        if l == 0:
            idx = key
        else:
            sk, si = self._sorted[l]
            if len(sk) == 0:
                return np.full(key.shape, -1, dtype=np.int64)
            pos = np.minimum(np.searchsorted(sk, key), len(sk) - 1)
            idx = np.where(sk[pos] == key, si[pos], -1)
        return np.where(inside, idx, -1)

    def neighbor_offsets(self):
        return [np.array(d, dtype=np.int64) for d in itertools.product((-1, 0, 1), repeat=self.dim) if any(d)]

    # -------------------------------------------------------------- refinement
    def balance_flags(self, flags):
        """Close a set of refine flags under 2:1 balance across faces, edges and corners."""
        flags = [np.asarray(f, dtype=bool).copy() for f in flags]
        while len(flags) < self.n_levels:
            flags.append(np.zeros(self.n_cells(len(flags)), dtype=bool))
        for l in range(self.n_levels - 1, 0, -1):
            idx = np.nonzero(flags[l])[0]
            if len(idx) == 0:
                continue
            n = self.cells_per_axis(l)
            for d in self.neighbor_offsets():
                nb = self.ijk[l][idx] + d
                inside = np.all((nb >= 0) & (nb < n), axis=1)
                missing = inside & (self.lookup(l, nb) < 0)
                if not missing.any():
                    continue
                pidx = self.lookup(l - 1, nb[missing] >> 1)
                assert (pidx >= 0).all(), "mesh was not 2:1 balanced"
                assert (self.child0[l - 1][pidx] < 0).all()
                flags[l - 1][pidx] = True
        return flags

    def refine(self, flags):
        """Refine flagged active cells (after balancing).  Returns, per level, the parent indices refined."""
        flags = self.balance_flags(flags)
        refined = []
        nl = self.n_levels
        for l in range(nl):
            fl = np.zeros(self.n_cells(l), dtype=bool)  # children created in this pass are never flagged
            fl[:len(flags[l])] = flags[l]
            idx = np.nonzero(fl & self.active_mask(l))[0]
            refined.append(idx)
            if len(idx) == 0:
                continue
            if l + 1 == self.n_levels:
                self.ijk.append(np.zeros((0, self.dim), dtype=np.int64))
                self.parent.append(np.zeros(0, dtype=np.int64))
                self.child0.append(np.zeros(0, dtype=np.int64))
                self._sorted.append(None)
            nch = 1 << self.dim
            start = self.n_cells(l + 1)
            self.child0[l][idx] = start + nch * np.arange(len(idx), dtype=np.int64)
            cijk = (2 * self.ijk[l][idx])[:, None, :] + self.VO[None, :, :]
            self.ijk[l + 1] = np.concatenate([self.ijk[l + 1], cijk.reshape(-1, self.dim)])
            self.parent[l + 1] = np.concatenate([self.parent[l + 1], np.repeat(idx, nch)])
            self.child0[l + 1] = np.concatenate([self.child0[l + 1], np.full(nch * len(idx), -1, dtype=np.int64)])
            self._rebuild_index(l + 1)
        return refined

    def refine_global(self, times=1):
        for _ in range(times):
            self.refine([self.active_mask(l) for l in range(self.n_levels)])

    # ------------------------------------------------------------- geometry
    def resolution(self):
        """Vertex coordinates are integers on the grid of the finest level."""
        return self.n_levels - 1

    def vertex_coords(self, l, idx=None, res=None):
        """Integer coordinates (at resolution `res`) of the 2^dim vertices of level-l cells: (n, 2^dim, dim)."""
        res = self.resolution() if res is None else res
        ijk = self.ijk[l] if idx is None else self.ijk[l][idx]
        return (ijk[:, None, :] + self.VO[None, :, :]) << (res - l)

    def points_per_axis(self, res=None):
        res = self.resolution() if res is None else res
        return (self.reps << res) + 1

    def vertex_key(self, xyz, res=None):
        n = self.points_per_axis(res)
        key = xyz[..., 0].astype(np.int64).copy()
        mult = n
        for d in range(1, self.dim):
            key += xyz[..., d] * mult
            mult *= n
        return key

    def key_to_xyz(self, key, res=None):
        n = self.points_per_axis(res)
        out = np.empty(key.shape + (self.dim,), dtype=np.int64)
        k = key.copy()
        for d in range(self.dim):
            out[..., d] = k % n
            k //= n
        return out

    def real_coords(self, xyz, res=None):
        res = self.resolution() if res is None else res
        return self.lo + xyz * (self.H / (1 << res))
