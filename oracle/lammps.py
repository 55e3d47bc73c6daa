"""Oracle LAMMPS 'full' atom-file reader (src/step-50.cc:181-258): a whitespace-token walk;
token #2 is the atom count, tokens 3..34 are skipped, from token #35 on each atom is
`id mol type q x y z`.  TEST INFRASTRUCTURE -- see oracle/__init__.py."""
import numpy as np


def read(path):
    with open(path) as f:
        tok = f.read().split()
    n = int(tok[2])
    body = tok[35:35 + 7 * n]
    if len(body) < 7 * n:
        raise ValueError("truncated atom file")
    a = np.array(body, dtype=object).reshape(n, 7)
    types = a[:, 2].astype(np.int64)
    charges = a[:, 3].astype(np.float64)
    pos = a[:, 4:7].astype(np.float64)
    return pos, charges, types


def nacl_lattice(n):
    """The reference's atom/atom_n{n}_{8 n^3}.data lattices."""
    basis = np.array([(0, 0, 0), (.5, 0, 0), (.5, .5, 0), (0, .5, 0), (.5, 0, .5), (0, 0, .5), (0, .5, .5), (.5, .5, .5)])
    g = np.arange(n)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)  # x slowest, z fastest
    pos = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3).astype(np.float64)
    par = np.rint(2 * pos.sum(1)).astype(np.int64)
    q = np.where(par % 2 == 0, 1.0, -1.0)
    return pos, q


def write(path, pos, q):
    n = len(q)
    hi = float(np.ceil(pos.max() + 0.5)) if n else 1.0
    with open(path, "w") as f:
        f.write("LAMMPS Description\n\n")
        f.write(f"     {n}  atoms\n     0  bonds\n     0  angles\n     0  dihedrals\n     0  impropers\n\n")
        f.write("     2  atom types\n\n")
        for ax in "xyz":
            f.write(f"  0.0 {hi:.1f} {ax}lo {ax}hi\n")
        f.write("\nMasses\n\n      1\t\t22.989\n      2 \t35.453\n\nAtoms # full\n\n")
        for i in range(n):
            t = 1 if q[i] > 0 else 2
            x, y, z = (repr(float(v)) for v in pos[i])  # shortest text that reads back to the same double
            f.write(f"{i + 1} {i + 1} {t} {q[i]:.1f} {x} {y} {z}\n")
