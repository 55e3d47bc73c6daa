"""Oracle `.prm` reader: the 19 entries the reference declares (src/step-50.cc:13-96) with their
defaults and patterns, in deal.II ParameterHandler syntax (`subsection X` / `end` / `set K = V` /
`#` comments).  TEST INFRASTRUCTURE -- see oracle/__init__.py."""

DECLARED = {
    ("Geometry", "Number of global refinement"): ("2", "int"),
    ("Geometry", "Domain limit left"): ("-1", "double"),
    ("Geometry", "Domain limit right"): ("1", "double"),
    ("Geometry", "Mesh size"): ("0.25", "double"),
    ("Geometry", "Vacuum repetitions"): ("1", "int"),
    # not a key of the reference: the product's optional deeper hierarchy, restated here so that the
    # option has an oracle too; 0 = the reference's behaviour
    ("Geometry", "Coarse levels below the base mesh"): ("0", "int"),
    ("Problem Selection", "Problem"): ("Step16", ("Step16", "GaussianCharges")),
    ("Problem Selection", "Dimension"): ("2", "int"),
    ("Problem Selection", "Boundary conditions selection"): ("Inhomogeneous", ("Homogeneous", "Inhomogeneous", "Exact")),
    ("Misc", "Number of Adaptive Refinement"): ("2", "int"),
    ("Misc", "smoothing length"): ("0.5", "double"),
    ("Misc", "Nonzero Density radius parameter around each charge"): ("3", "double"),
    ("Misc", "Output and calculation of Analytical solution"): ("false", "bool"),
    ("Misc", "Output of RHS field"): ("false", "bool"),
    ("Misc", "Output of support of each atom"): ("false", "bool"),
    ("Misc", "Flag for RHS evaluation optimization"): ("false", "bool"),
    ("Misc", "Quadrature points for RHS function"): ("1", "int"),
    ("Misc", "Output time summary table"): ("true", "bool"),
    ("", "Polynomial degree"): ("1", "int"),
    ("Solver input data", "Preconditioner"): ("GMG", ("GMG", "Jacobi")),
    ("Lammps data", "Lammps input file"): ("atom_8.data", "any"),
}


def _convert(value, pattern, key):
    v = value.strip()
    if pattern == "int":
        return int(v)
    if pattern == "double":
        return float(v)
    if pattern == "bool":
        if v.lower() in ("true", "yes", "on"):
            return True
        if v.lower() in ("false", "no", "off"):
            return False
        raise ValueError(f"entry <{key}> does not match pattern Bool: {v}")
    if pattern == "any":
        return v
    if v not in pattern:
        raise ValueError(f"entry <{key}> = <{v}> does not match Selection {pattern}")
    return v


def parse_string(text):
    values = {k: _convert(d, p, k[1]) for k, (d, p) in DECLARED.items()}
    stack = []
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        low = line.lower()
        if low.startswith("subsection"):
            stack.append(" ".join(line.split()[1:]))
        elif low == "end":
            if not stack:
                raise ValueError("unbalanced 'end'")
            stack.pop()
        elif low.startswith("set"):
            body = line[3:].strip()
            name, _, val = body.partition("=")
            name = " ".join(name.split())
            key = (stack[-1] if stack else "", name)
            if key not in DECLARED:
                raise ValueError(f"No entry with name <{name}> was declared in subsection <{key[0]}>")
            values[key] = _convert(val, DECLARED[key][1], name)
        else:
            raise ValueError(f"Could not parse line: {raw!r}")
    if stack:
        raise ValueError("unclosed subsection")
    return values


def parse_file(path):
    with open(path) as f:
        return parse_string(f.read())
