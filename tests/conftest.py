import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def make_prm(problem="GaussianCharges", dim=3, nref=0, left=0.0, right=1.0, cycles=1, bc="Homogeneous",
             atom="atom_n1_8.data", h=0.25, vacuum=10, cutoff=3.5, nq=1, flag="true", precond="GMG", extra=""):
    """A .prm text in the reference's grammar (src/step-50.cc:13-96)."""
    if atom and not os.path.isabs(atom) and os.path.exists(os.path.join(GOLDEN, atom)):
        atom = os.path.join(GOLDEN, atom)
    return f"""
subsection Geometry
  set Number of global refinement = {nref}
  set Domain limit left = {left}
  set Domain limit right = {right}
  set Mesh size = {h}
  set Vacuum repetitions = {vacuum}
end
subsection Misc
  set Number of Adaptive Refinement = {cycles}
  set smoothing length = 0.5
  set Nonzero Density radius parameter around each charge = {cutoff}
  set Flag for RHS evaluation optimization = {flag}
  set Quadrature points for RHS function = {nq}
  set Output time summary table = false
end
set Polynomial degree = 1
subsection Solver input data
  set Preconditioner = {precond}
end
subsection Problem Selection
  set Problem = {problem}
  set Dimension = {dim}
  set Boundary conditions selection = {bc}
end
subsection Lammps data
  set Lammps input file = {atom}
end
{extra}
"""


@pytest.fixture(scope="session")
def goldens():
    import json
    with open(os.path.join(GOLDEN, "reference_goldens.json")) as f:
        return json.load(f)
