"""Oracle RHS path: atom->cell binning, list inheritance, Gaussian charge densities.

Restates src/step-50.cc:260-306 (`rhs_assembly_optimization`), :377-491 (children inherit the
parent's list on refinement) and :509-575 (`compute_charge_densities`).
TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import math

import numpy as np

from . import fe


def bin_atoms_base(forest, pos, radius, chunk=256, base_level=0):
    """Cell k of level 0 lists atom i iff some vertex v of the cell has ||X_i - v||_2 < radius
    (strict, src/step-50.cc:278-283).  Returns CSR (ptr, idx) with ascending atom indices per cell
    (std::set iteration order).  base_level > 0 (the product's `Coarse levels below the base mesh`): the base lattice
    is level `base_level` of the forest (all of its cells exist); the lists are returned in that level's cell order."""
    assert forest.dim == 3
    reps, H, lo = forest.reps << base_level, forest.h(base_level), forest.lo
    nv = reps + 1
    w = int(math.ceil(radius / H)) + 1
    cells, atoms = [], []
    off = np.arange(-w, w + 2)
    for s in range(0, len(pos), chunk):
        X = pos[s:s + chunk]
        c0 = np.floor((X - lo) / H).astype(np.int64)  # (m, 3)
        vi = c0[:, None, :] + off[None, :, None]  # vertex indices per axis (m, nw, 3)
        valid = (vi >= 0) & (vi < nv)
        d = np.where(valid, lo + vi * H - X[:, None, :], np.inf)
        # distance as Point::distance: sqrt(sum of squares)
        d2 = d[:, :, None, None, 0] ** 2 + d[:, None, :, None, 1] ** 2 + d[:, None, None, :, 2] ** 2
        near = np.sqrt(d2) < radius  # (m, nw, nw, nw) over vertices
        cm = np.zeros(tuple(np.array(near.shape) - [0, 1, 1, 1]), dtype=bool)
        for v in range(8):
            a, b, c = v & 1, (v >> 1) & 1, (v >> 2) & 1
            cm |= near[:, a:a + cm.shape[1], b:b + cm.shape[2], c:c + cm.shape[3]]
        m, ii, jj, kk = np.nonzero(cm)
        ci = c0[m, 0] + off[ii]
        cj = c0[m, 1] + off[jj]
        ck = c0[m, 2] + off[kk]
        ok = (ci >= 0) & (ci < reps) & (cj >= 0) & (cj < reps) & (ck >= 0) & (ck < reps)
        cells.append((ci + reps * (cj + reps * ck))[ok])
        atoms.append((m + s)[ok])
    cells = np.concatenate(cells) if cells else np.zeros(0, dtype=np.int64)
    atoms = np.concatenate(atoms) if atoms else np.zeros(0, dtype=np.int64)
    order = np.lexsort((atoms, cells))
    cells, atoms = cells[order], atoms[order]
    ptr = np.zeros(reps ** 3 + 1, dtype=np.int64)
    np.add.at(ptr, cells + 1, 1)
    ptr, atoms = np.cumsum(ptr), atoms.astype(np.int64)
    if base_level == 0:
        return ptr, atoms
    # lexicographic lattice index -> cell order of the forest's level `base_level`
    ijk = forest.ijk[base_level].astype(np.int64)
    assert len(ijk) == reps ** 3
    lex = ijk[:, 0] + reps * (ijk[:, 1] + reps * ijk[:, 2])
    cnt = ptr[lex + 1] - ptr[lex]
    out_ptr = np.concatenate([[0], np.cumsum(cnt)])
    src = np.repeat(ptr[lex], cnt) + (np.arange(out_ptr[-1]) - np.repeat(out_ptr[:-1], cnt))
    return out_ptr, atoms[src]


def bin_atoms_bruteforce(forest, pos, radius):
    """The reference's literal triple loop over (cells, atoms, vertices); for small cases only."""
    verts = forest.real_coords(forest.vertex_coords(0, res=0), res=0)  # (nc, 8, 3)
    ptr, idx = [0], []
    for c in range(len(verts)):
        d = np.sqrt(((verts[c][None, :, :] - pos[:, None, :]) ** 2).sum(2))
        hit = np.nonzero((d < radius).any(1))[0]
        idx.append(hit)
        ptr.append(ptr[-1] + len(hit))
    return np.array(ptr, dtype=np.int64), np.concatenate(idx).astype(np.int64)


def inherit_lists(forest, lists0, base_level=0):
    """Per-level CSR lists: every cell of level l+1 copies its parent's list (src/step-50.cc:441-449).  Levels below
    `base_level` (never active) get empty lists."""
    out = [(np.zeros(forest.n_cells(l) + 1, dtype=np.int64), np.zeros(0, dtype=np.int64)) for l in range(base_level)]
    out.append(lists0)
    for l in range(base_level + 1, forest.n_levels):
        pptr, pidx = out[l - 1]
        par = forest.parent[l]
        cnt = pptr[par + 1] - pptr[par]
        ptr = np.concatenate([[0], np.cumsum(cnt)])
        src = np.repeat(pptr[par], cnt) + (np.arange(ptr[-1]) - np.repeat(ptr[:-1], cnt))
        out.append((ptr, pidx[src]))
    return out


def density_constant(r_c):
    return 4.0 * math.pi / (r_c ** 3 * math.pi ** 1.5)  # src/step-50.cc:522


def charge_densities(forest, dofs, pos, charges, r_c, nq, lists=None, pair_chunk=2_000_000):
    """rho_q = sum_{k in list(cell)} C exp(-|X_k - x_q|^2 / r_c^2) q_k on active cells (src/step-50.cc:535-571).
    lists=None sums over all atoms (flag_rhs_assembly == false)."""
    dim = forest.dim
    pts, _ = fe.tensor_rule(nq, dim)
    C = density_constant(r_c)
    inv = 1.0 / (r_c * r_c)
    out = []
    for l in range(forest.n_levels):
        act = dofs.active_cells[l]
        h = forest.h(l)
        org = forest.lo + forest.ijk[l][act] * h
        dens = np.zeros((len(act), len(pts)))
        if lists is None:
            for c0 in range(0, len(act), 4096):
                xq = org[c0:c0 + 4096, None, :] + h * pts[None, :, :]
                for k in range(len(pos)):
                    r = np.sqrt(((xq - pos[k]) ** 2).sum(-1))
                    dens[c0:c0 + 4096] += C * np.exp(-(r * r) * inv) * charges[k]
        else:
            ptr, idx = lists[l]
            cnt = ptr[act + 1] - ptr[act]
            cell_of_pair = np.repeat(np.arange(len(act)), cnt)
            start = np.concatenate([[0], np.cumsum(cnt)])
            src = np.repeat(ptr[act], cnt) + (np.arange(start[-1]) - np.repeat(start[:-1], cnt))
            atom_of_pair = idx[src]
            for s in range(0, len(cell_of_pair), pair_chunk):
                cp = cell_of_pair[s:s + pair_chunk]
                ap = atom_of_pair[s:s + pair_chunk]
                xq = org[cp][:, None, :] + h * pts[None, :, :]
                r = np.sqrt(((xq - pos[ap][:, None, :]) ** 2).sum(-1))
                contrib = C * np.exp(-(r * r) * inv) * charges[ap][:, None]
                np.add.at(dens, cp, contrib)
        out.append(dens)
    return out
