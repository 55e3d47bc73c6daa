"""Oracle Q1 element on axis-aligned cubes/squares: Gauss quadrature, shape values, stiffness.

Restates what the reference gets from `FE_Q<dim>(1)`, `QGauss<dim>(n)` and `FEValues`
(src/step-50.cc:744-790, 515-518).  Tensor-product points are ordered x fastest as in
deal.II's `QGauss`.  TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import numpy as np

from .mesh import vertex_offsets


def gauss_unit(n):
    """n-point Gauss-Legendre rule mapped to [0, 1]."""
    p, w = np.polynomial.legendre.leggauss(n)
    return (p + 1.0) / 2.0, w / 2.0


def tensor_rule(n, dim):
    p1, w1 = gauss_unit(n)
    pts = np.zeros((n ** dim, dim))
    wts = np.ones(n ** dim)
    for q in range(n ** dim):
        r = q
        for d in range(dim):  # x fastest
            i = r % n
            r //= n
            pts[q, d] = p1[i]
            wts[q] *= w1[i]
    return pts, wts


def shape_values(pts, dim):
    """(nq, 2^dim) values of the Q1 shape functions (vertex ordering of deal.II) at unit-cell points."""
    VO = vertex_offsets(dim)
    out = np.ones((len(pts), 1 << dim))
    for v in range(1 << dim):
        for d in range(dim):
            out[:, v] *= pts[:, d] if VO[v, d] else (1.0 - pts[:, d])
    return out


def shape_grads(pts, dim):
    """(nq, 2^dim, dim) unit-cell gradients."""
    VO = vertex_offsets(dim)
    out = np.ones((len(pts), 1 << dim, dim))
    for v in range(1 << dim):
        for g in range(dim):
            for d in range(dim):
                if d == g:
                    out[:, v, g] *= 1.0 if VO[v, d] else -1.0
                else:
                    out[:, v, g] *= pts[:, d] if VO[v, d] else (1.0 - pts[:, d])
    return out


def stiffness_q(dim, nq=2):
    """Per-quadrature-point unit-cell stiffness contributions G[q, i, j] = grad_i . grad_j * w_q.
    The cell matrix of a cube of edge h with coefficient values c_q is h^(dim-2) * sum_q c_q G[q]."""
    pts, wts = tensor_rule(nq, dim)
    g = shape_grads(pts, dim)
    return np.einsum("qid,qjd,q->qij", g, g, wts), pts


def stiffness(h, dim, coef=None):
    """Q1 Laplace cell matrix on a cube of edge h (src/step-50.cc:782-790), optional coefficient per q-point."""
    G, _ = stiffness_q(dim)
    if coef is None:
        return G.sum(0) * h ** (dim - 2)
    return np.einsum("...q,qij->...ij", coef, G) * h ** (dim - 2)
