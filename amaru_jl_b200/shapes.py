"""Cell shapes used by the mechanical hot path (host side, numpy).

Mirrors the part of the reference's ``CellShape`` registry that the path needs
(``src/shape/shape.jl:19-37``): natural coordinates, shape functions N(R), their
derivatives dN/dR, facet index tables and the *default* quadrature of each shape
(``shape.quadrature[0]``).

The functions are written in tensor-product / barycentric form from the node
natural coordinates instead of the reference's expanded per-node polynomials
(``src/shape/solids2d.jl:295-412``, ``src/shape/solids3d.jl:21-191,472-691``,
``src/shape/lines.jl:10-84``); they evaluate the same polynomials.

Shape ids are the ones of ``include/amaru_b200.h`` (``AMARU_SHAPE_*``).
"""
from __future__ import annotations

import numpy as np

# --- quadrature tables (literal constants are data: src/shape/quadrature.jl) ---------------
_G2 = 0.577350269189626            # quadrature.jl:62-66,167-175 (QUAD_IP4, HEX_IP8)
_G2L = 0.577350269189625764509149  # quadrature.jl:16-18 (LIN_IP2)
_TA, _TB = 0.5854101966249685, 0.1381966011250105  # quadrature.jl:110-114 (TET_IP4)


def _lin_ip2():
    return np.array([[-_G2L, 0, 0, 1.0], [_G2L, 0, 0, 1.0]])


def _quad_ip4():
    # r fastest, then s (quadrature.jl:62-66)
    return np.array([[r * _G2, s * _G2, 0.0, 1.0] for s in (-1, 1) for r in (-1, 1)])


def _hex_ip8():
    # r fastest, then s, then t (quadrature.jl:167-175)
    return np.array([[r * _G2, s * _G2, t * _G2, 1.0]
                     for t in (-1, 1) for s in (-1, 1) for r in (-1, 1)])


def _tet_ip4():
    w = 0.04166666666666667
    return np.array([[_TA, _TB, _TB, w], [_TB, _TA, _TB, w], [_TB, _TB, _TA, w], [_TB, _TB, _TB, w]])


def _tri_ip3():
    # quadrature.jl:37-40
    return np.array([[1 / 6, 1 / 6, 0, 1 / 6], [2 / 3, 1 / 6, 0, 1 / 6], [1 / 6, 2 / 3, 0, 1 / 6]])


class CellShape:
    """One cell shape: name, ndim, natural node coordinates, N, dN/dR, facets, default IPs."""

    def __init__(self, name, sid, ndim, nat, facet_idxs, facet_shape, quad):
        self.name = name
        self.id = sid
        self.ndim = ndim
        self.nat_coords = np.asarray(nat, dtype=np.float64)
        self.npoints = self.nat_coords.shape[0]
        self.facet_idxs = [np.asarray(f, dtype=np.int64) - 1 for f in facet_idxs]  # 0-based
        self.facet_shape = facet_shape
        self.quadrature = quad  # rows (r, s, t, w)

    def __repr__(self):
        return f"CellShape({self.name})"

    # N at one natural point R -> (npoints,)
    def func(self, R):
        return _FUNC[self.name](np.asarray(R, dtype=np.float64), self.nat_coords)[0]

    # dN/dR at one natural point R -> (npoints, ndim)
    def deriv(self, R):
        return _FUNC[self.name](np.asarray(R, dtype=np.float64), self.nat_coords)[1]


# --- shape function families ----------------------------------------------------------------
def _lagrange_linear(R, nat):
    """LIN2 / QUAD4 / HEX8: N_i = prod_d (1 + x_d xi_d)/2."""
    nd = nat.shape[1]
    f = 0.5 * (1.0 + nat * R[:nd])             # (n, nd)
    N = np.prod(f, axis=1)
    D = np.empty_like(f)
    for d in range(nd):
        g = 0.5 * nat[:, d]
        for e in range(nd):
            if e != d:
                g = g * f[:, e]
        D[:, d] = g
    return N, D


def _lin3(R, nat):
    r = R[0]
    N = np.array([0.5 * (r * r - r), 0.5 * (r * r + r), 1.0 - r * r])
    D = np.array([[r - 0.5], [r + 0.5], [-2.0 * r]])
    return N, D


def _serendipity(R, nat):
    """QUAD8 / HEX20 serendipity family from the node natural coordinates.

    corner i : N = prod(1+x xi)/2^nd * (sum(x xi) - (nd-1))
    midside i (xi_m = 0): N = (1 - x_m^2) * prod_{d != m}(1 + x_d xi_d) / 2^(nd-1)
    """
    n, nd = nat.shape
    x = R[:nd]
    N = np.empty(n)
    D = np.empty((n, nd))
    for i in range(n):
        xi = nat[i]
        zero = [d for d in range(nd) if xi[d] == 0.0]
        if not zero:
            f = 1.0 + xi * x
            P = np.prod(f) / 2 ** nd
            S = float(np.dot(xi, x)) - (nd - 1)
            N[i] = P * S
            for d in range(nd):
                Pd = xi[d] * np.prod([f[e] for e in range(nd) if e != d]) / 2 ** nd
                D[i, d] = Pd * S + P * xi[d]
        else:
            m = zero[0]
            others = [d for d in range(nd) if d != m]
            f = {d: 1.0 + xi[d] * x[d] for d in others}
            c = 1.0 / 2 ** (nd - 1)
            q = 1.0 - x[m] * x[m]
            N[i] = c * q * np.prod([f[d] for d in others])
            for d in range(nd):
                if d == m:
                    D[i, d] = c * (-2.0 * x[m]) * np.prod([f[e] for e in others])
                else:
                    D[i, d] = c * q * xi[d] * np.prod([f[e] for e in others if e != d])
    return N, D


def _simplex_quadratic(R, nat):
    """TRI6 / TET10 in barycentric form: corners L(2L-1), mid-edges 4 L_a L_b."""
    n, nd = nat.shape
    x = R[:nd]
    L = np.concatenate(([1.0 - x.sum()], x))          # L0 = u, L1 = r, ...
    dL = np.vstack((-np.ones(nd), np.eye(nd)))         # (nd+1, nd)
    N = np.empty(n)
    D = np.empty((n, nd))
    for i in range(n):
        Li = np.concatenate(([1.0 - nat[i].sum()], nat[i]))
        nz = np.nonzero(Li > 0.25)[0]
        if len(nz) == 1:                               # corner
            a = nz[0]
            N[i] = L[a] * (2.0 * L[a] - 1.0)
            D[i] = (4.0 * L[a] - 1.0) * dL[a]
        else:                                          # mid-edge between corners a, b
            a, b = nz
            N[i] = 4.0 * L[a] * L[b]
            D[i] = 4.0 * (L[a] * dL[b] + L[b] * dL[a])
    return N, D


_FUNC = {
    "LIN2": _lagrange_linear, "QUAD4": _lagrange_linear, "HEX8": _lagrange_linear,
    "LIN3": _lin3, "QUAD8": _serendipity, "HEX20": _serendipity,
    "TRI6": _simplex_quadratic, "TET10": _simplex_quadratic,
}

# --- registry ----------------------------------------------------------------------------------
LIN2 = CellShape("LIN2", 101, 1, [[-1.0], [1.0]], [], None, _lin_ip2())            # lines.jl:48
LIN3 = CellShape("LIN3", 102, 1, [[-1.0], [1.0], [0.0]], [], None, _lin_ip2())     # lines.jl:108
QUAD4 = CellShape("QUAD4", 1, 2, [[-1, -1], [1, -1], [1, 1], [-1, 1]],
                  [[1, 2], [2, 3], [3, 4], [4, 1]], LIN2, _quad_ip4())             # solids2d.jl:287-327
QUAD8 = CellShape("QUAD8", 2, 2,
                  [[-1, -1], [1, -1], [1, 1], [-1, 1], [0, -1], [1, 0], [0, 1], [-1, 0]],
                  [[1, 2, 5], [2, 3, 6], [3, 4, 7], [4, 1, 8]], LIN3, _quad_ip4())  # solids2d.jl:359-427
TRI6 = CellShape("TRI6", 103, 2, [[0, 0], [1, 0], [0, 1], [.5, 0], [.5, .5], [0, .5]],
                 [[1, 2, 4], [2, 3, 5], [3, 1, 6]], LIN3, _tri_ip3())             # solids2d.jl:99-151
HEX8 = CellShape("HEX8", 3, 3,
                 [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                  [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]],
                 [[1, 5, 8, 4], [2, 3, 7, 6], [1, 2, 6, 5], [3, 4, 8, 7], [1, 4, 3, 2], [5, 6, 7, 8]],
                 QUAD4, _hex_ip8())                                                # solids3d.jl:459-519
HEX20 = CellShape("HEX20", 4, 3,
                  [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                   [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1],
                   [0, -1, -1], [1, 0, -1], [0, 1, -1], [-1, 0, -1],
                   [0, -1, 1], [1, 0, 1], [0, 1, 1], [-1, 0, 1],
                   [-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]],
                  [[1, 5, 8, 4, 17, 16, 20, 12], [2, 3, 7, 6, 10, 19, 14, 18],
                   [1, 2, 6, 5, 9, 18, 13, 17], [3, 4, 8, 7, 11, 20, 15, 19],
                   [1, 4, 3, 2, 12, 11, 10, 9], [5, 6, 7, 8, 13, 14, 15, 16]],
                  QUAD8, _hex_ip8())                                               # solids3d.jl:557-706
TET10 = CellShape("TET10", 5, 3,
                  [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [.5, 0, 0], [.5, .5, 0],
                   [0, .5, 0], [0, 0, .5], [.5, 0, .5], [0, .5, .5]],
                  [[1, 4, 3, 8, 10, 7], [1, 2, 4, 5, 9, 8], [1, 3, 2, 7, 6, 5], [2, 3, 4, 6, 10, 9]],
                  TRI6, _tet_ip4())                                                # solids3d.jl:108-206

# VTK cell types (src/shape/shape.jl:66-96)
for _s, _v in ((LIN2, 3), (LIN3, 21), (QUAD4, 9), (QUAD8, 23), (TRI6, 22), (HEX8, 12), (HEX20, 25), (TET10, 24)):
    _s.vtk_type = _v

SHAPES = {s.name: s for s in (LIN2, LIN3, QUAD4, QUAD8, TRI6, HEX8, HEX20, TET10)}
SOLID_SHAPES_BY_ID = {s.id: s for s in (QUAD4, QUAD8, HEX8, HEX20, TET10)}


def deriv_table(shape: CellShape) -> np.ndarray:
    """dN/dR at the default integration points -> (nip, npoints, ndim)."""
    return np.stack([shape.deriv(q[:3]) for q in shape.quadrature])


def func_table(shape: CellShape) -> np.ndarray:
    """N at the default integration points -> (nip, npoints)."""
    return np.stack([shape.func(q[:3]) for q in shape.quadrature])
