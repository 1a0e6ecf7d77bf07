"""CPU oracle of the natural boundary conditions — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Entity-by-entity restatement (plain loops, one facet / one element at a time) of
  mech_boundary_forces       src/mech/elem/distributed.jl:76-152
  mech_solid_body_forces     src/mech/elem/distributed.jl:157-217
  norm2                      src/tools/linalg.jl:52-72
  F[map] += Fd accumulation  src/bc.jl:116-136,175-194
with the facet shape functions written out in the reference's expanded polynomial form
  LIN2 / LIN3  src/shape/lines.jl:10-24,58-75      TRI6  src/shape/solids2d.jl:110-136
(QUAD4 / QUAD8 faces and the cell shapes come from the C oracle, amaru_oracle.c).  Default quadrature of each shape
(get_ip_coords(shape), src/shape/shape.jl:61-64): LIN_IP2, TRI_IP3, QUAD_IP4, HEX_IP8, TET_IP4 (src/shape/quadrature.jl).

Parity pinning: the reference's own known answers for edge / triangular-face / quadrilateral-face / body loads
(test/mech/elem/elastic-hex8.jl:29-59, elastic-quad4.jl) are reproduced through this file in tests/test_loads.py.
Only tests/ may import this.
"""
from __future__ import annotations

import numpy as np

from . import oracle as O

LIN2, LIN3, TRI6 = 101, 102, 103
QUAD4, QUAD8, HEX8, HEX20, TET10 = 1, 2, 3, 4, 5
KEY_X, KEY_Y, KEY_Z, KEY_N = 0, 1, 2, 3


def shape_dim(shape):
    return {LIN2: 1, LIN3: 1, TRI6: 2, QUAD4: 2, QUAD8: 2, HEX8: 3, HEX20: 3, TET10: 3}[shape]


def shape_nn(shape):
    return {LIN2: 2, LIN3: 3, TRI6: 6, QUAD4: 4, QUAD8: 8, HEX8: 8, HEX20: 20, TET10: 10}[shape]


def quadrature(shape):
    if shape in (LIN2, LIN3):          # quadrature.jl:16-18
        g = 0.577350269189625764509149
        return np.array([[-g, 0.0, 0.0, 1.0], [g, 0.0, 0.0, 1.0]])
    if shape == TRI6:                  # quadrature.jl:37-40
        return np.array([[1 / 6, 1 / 6, 0.0, 1 / 6], [2 / 3, 1 / 6, 0.0, 1 / 6], [1 / 6, 2 / 3, 0.0, 1 / 6]])
    return O.quadrature(shape)


def func(shape, R):
    r, s = R[0], R[1]
    if shape == LIN2:
        return np.array([0.5 * (1 - r), 0.5 * (1 + r)])
    if shape == LIN3:
        return np.array([0.5 * (r * r - r), 0.5 * (r * r + r), 1.0 - r * r])
    if shape == TRI6:
        return np.array([1.0 - (r + s) * (3.0 - 2.0 * (r + s)), r * (2.0 * r - 1.0), s * (2.0 * s - 1.0),
                         4.0 * r * (1.0 - (r + s)), 4.0 * r * s, 4.0 * s * (1.0 - (r + s))])
    return O.shape_func(shape, np.asarray(R, dtype=np.float64))


def deriv(shape, R):
    r, s = R[0], R[1]
    if shape == LIN2:
        return np.array([[-0.5], [0.5]])
    if shape == LIN3:
        return np.array([[r - 0.5], [r + 0.5], [-2.0 * r]])
    if shape == TRI6:
        return np.array([[-3.0 + 4.0 * (r + s), -3.0 + 4.0 * (r + s)], [4.0 * r - 1.0, 0.0], [0.0, 4.0 * s - 1.0],
                         [4.0 - 8.0 * r - 4.0 * s, -4.0 * r], [4.0 * s, 4.0 * r], [-4.0 * s, 4.0 - 4.0 * r - 8.0 * s]])
    return O.shape_deriv(shape, np.asarray(R, dtype=np.float64))


def norm2(J):
    """tools/linalg.jl:52-72"""
    r, c = J.shape
    if r == c:
        return float(np.linalg.det(J))
    if r == 1 or c == 1:
        return float(np.sqrt((J * J).sum()))
    if r == 3 and c == 2:
        j1 = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
        j2 = J[0, 0] * J[2, 1] - J[0, 1] * J[2, 0]
        j3 = J[1, 0] * J[2, 1] - J[1, 1] * J[2, 0]
        return float((j1 * j1 + j2 * j2 + j3 * j3) ** 0.5)
    raise ValueError("No rule to calculate norm2")


def ip_coords(shape, coords, nodes, ndim):
    """X = C'N at every integration point of every entity -> (nents*nip, 3)."""
    ips = quadrature(shape)
    X = np.zeros((nodes.shape[0] * ips.shape[0], 3))
    for e in range(nodes.shape[0]):
        C = coords[nodes[e], :ndim]
        for q in range(ips.shape[0]):
            X[e * ips.shape[0] + q, :ndim] = C.T @ func(shape, ips[q, :3])
    return X


def entity_forces(shape, C, ndim, th, key, vals, axi=False):
    """One call of mech_boundary_forces (facet) / mech_solid_body_forces (cell): C (nn, ndim), vals (nip,) -> (nn, ndim)."""
    ips = quadrature(shape)
    facet = shape_dim(shape) < ndim        # faces, 2D edges, and the 3D edges of EdgeBC (qx qy qz; th = 1.0, distributed.jl:95)
    F = np.zeros((C.shape[0], ndim))
    for q in range(ips.shape[0]):
        R, w = ips[q, :3], ips[q, 3]
        N = func(shape, R)
        D = deriv(shape, R)
        J = C.T @ D
        Q = np.zeros(ndim)
        if key == KEY_N:
            assert facet
            n = np.array([J[1, 0], -J[0, 0]]) if ndim == 2 else np.cross(J[:, 0], J[:, 1])
            Q = vals[q] * n / np.sqrt((n * n).sum())
        else:
            Q[key] = vals[q]
        if axi:
            th = 2 * np.pi * float(C[:, 0] @ N)      # ctx.stressmodel==:axisymmetric && (th = 2*pi*X[1]), distributed.jl:121,193
        coef = (norm2(J) if facet else float(np.linalg.det(J))) * w * th
        F += coef * np.outer(N, Q)
    return F


def apply(shape, coords, nodes, eqid, ndim, th, key, vals, F, axi=False):
    """F[map] += Fd for every entity in order (bc.jl:131-134,189-192). vals: scalar or (nents*nip,)."""
    nip = quadrature(shape).shape[0]
    vals = np.broadcast_to(np.asarray(vals, dtype=np.float64), (nodes.shape[0] * nip,)) if np.ndim(vals) == 0 else vals
    for e in range(nodes.shape[0]):
        Fd = entity_forces(shape, coords[nodes[e], :ndim], ndim, th, key, vals[e * nip:(e + 1) * nip], axi)
        emap = eqid[nodes[e]].reshape(-1)
        F[emap] += Fd.reshape(-1)      # node-major map, no repeated node inside one entity
    return F
