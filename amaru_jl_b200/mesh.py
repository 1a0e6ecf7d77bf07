"""Structured block mesher (host side, numpy) producing flat SoA arrays.

Mirrors ``Block([x0 y0 z0; x1 y1 z1], nx=, ny=, nz=, cellshape=, tag=)`` and ``Mesh(blocks...)`` of the
reference (``src/mesh/block.jl:3-30,126-152``, ``src/mesh/structured.jl:37-554``,
``src/mesh/mesh.jl:335-391``): same node creation order (k outer, j, i inner; serendipity grid points
skipped), same cell order and local node order, coordinates rounded to 8 digits (``src/node.jl:57-61``),
boundary points of several blocks merged through their rounded coordinates (``structured.jl:481-487``).
Only straight two-corner boxes with uniform spacing (``rx=ry=rz=1``) are generated here; nothing in the
hot path depends on how the mesh was made — the C ABI takes ``coords/conn`` from any host.

Output is flat: ``mesh.coords (nnodes,3) float64``, ``mesh.conn (nelem, nn) int32`` (0-based),
``mesh.shape``, ``mesh.tags``.  No per-node / per-cell Python objects are created.
"""
from __future__ import annotations

import numpy as np

from . import shapes as S

_QUADRATIC = ("QUAD8", "HEX20", "TET10")


class Block:
    """A two-corner box to be split into nx*ny(*nz) cells (block.jl:126-152)."""

    def __init__(self, coords, nx=1, ny=1, nz=1, cellshape=None, tag=""):
        c = np.asarray(coords, dtype=np.float64)
        if c.shape[0] != 2:
            raise ValueError("Block: only two-corner boxes are supported")
        if c.shape[1] == 2:
            c = np.hstack((c, np.zeros((2, 1))))
        self.ndim = 3 if np.abs(c[:, 2]).sum() != 0 else 2
        if cellshape is None:
            cellshape = S.HEX8 if self.ndim == 3 else S.QUAD4
        if isinstance(cellshape, str):
            cellshape = S.SHAPES[cellshape]
        if cellshape.ndim != self.ndim:
            raise ValueError(f"Block: invalid cell type {cellshape.name} for dimension {self.ndim}")
        self.c0, self.c1 = c[0], c[1]
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz if self.ndim == 3 else 0)
        self.cellshape = cellshape
        self.tag = tag


def _axis(x0, x1, m):
    """Points of one direction: block shape function interpolation r=-1+2i/m (structured.jl:392,475)."""
    r = -1.0 + 2.0 * ((1.0 / m) * np.arange(m + 1))
    return 0.5 * (1.0 - r) * x0 + 0.5 * (1.0 + r) * x1


def _split_block(bl: Block):
    """-> coords (n,3), boundary flags (n,), conn (nelem, nn) local 0-based, in reference creation order."""
    sh = bl.cellshape
    quad = sh.name in _QUADRATIC
    f = 2 if quad else 1
    mx, my = f * bl.nx, f * bl.ny
    mz = f * bl.nz if bl.ndim == 3 else 0
    xs = _axis(bl.c0[0], bl.c1[0], mx)
    ys = _axis(bl.c0[1], bl.c1[1], my)
    zs = _axis(bl.c0[2], bl.c1[2], mz) if bl.ndim == 3 else np.array([0.0])
    K, J, I = np.meshgrid(np.arange(mz + 1), np.arange(my + 1), np.arange(mx + 1), indexing="ij")
    keep = np.ones(I.shape, dtype=bool)
    if sh.name == "QUAD8":
        keep = ~((I % 2 == 1) & (J % 2 == 1))                       # structured.jl:186
    elif sh.name == "HEX20":
        odd = (I % 2) + (J % 2) + (K % 2)
        keep = odd < 2                                              # structured.jl:466-470
    ids = np.full(I.shape, -1, dtype=np.int64)
    ids[keep] = np.arange(int(keep.sum()))                          # creation order k, j, i
    coords = np.stack((xs[I[keep]], ys[J[keep]], zs[K[keep]]), axis=1)
    coords = np.round(coords, 8) + 0.0                              # node.jl:58-60
    onb = (I == 0) | (I == mx) | (J == 0) | (J == my)
    if bl.ndim == 3:
        onb |= (K == 0) | (K == mz)
    boundary = onb[keep]

    # cells: k, j, i loops (stride f)
    if bl.ndim == 2:
        jj, ii = np.meshgrid(np.arange(0, my, f), np.arange(0, mx, f), indexing="ij")
        ii, jj = ii.ravel(), jj.ravel()
        g = lambda di, dj: ids[0, jj + dj, ii + di]
        if sh.name == "QUAD4":
            conn = np.stack((g(0, 0), g(1, 0), g(1, 1), g(0, 1)), axis=1)
        elif sh.name == "QUAD8":                                    # structured.jl:211-222
            conn = np.stack((g(0, 0), g(2, 0), g(2, 2), g(0, 2),
                             g(1, 0), g(2, 1), g(1, 2), g(0, 1)), axis=1)
        else:
            raise ValueError(f"block: cannot discretize using shape {sh.name}")
        return coords, boundary, conn
    kk, jj, ii = np.meshgrid(np.arange(0, mz, f), np.arange(0, my, f), np.arange(0, mx, f), indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    g = lambda di, dj, dk: ids[kk + dk, jj + dj, ii + di]
    if sh.name == "HEX8":                                           # structured.jl:416-426
        conn = np.stack((g(0, 0, 0), g(1, 0, 0), g(1, 1, 0), g(0, 1, 0),
                         g(0, 0, 1), g(1, 0, 1), g(1, 1, 1), g(0, 1, 1)), axis=1)
        return coords, boundary, conn
    # 27-point stencil p1..p27 (structured.jl:499-533)
    off = [(0, 0, 0), (2, 0, 0), (2, 2, 0), (0, 2, 0), (0, 0, 2), (2, 0, 2), (2, 2, 2), (0, 2, 2),
           (1, 0, 0), (2, 1, 0), (1, 2, 0), (0, 1, 0), (1, 0, 2), (2, 1, 2), (1, 2, 2), (0, 1, 2),
           (0, 0, 1), (2, 0, 1), (2, 2, 1), (0, 2, 1),
           (0, 1, 1), (2, 1, 1), (1, 0, 1), (1, 2, 1), (1, 1, 0), (1, 1, 2), (1, 1, 1)]
    if sh.name == "HEX20":
        conn = np.stack([g(*o) for o in off[:20]], axis=1)
        return coords, boundary, conn
    if sh.name == "TET10":
        p = [None] + [g(*o) for o in off]                            # 1-based like the reference
        tets = [(2, 4, 1, 8, 25, 12, 9, 27, 20, 21), (2, 1, 5, 8, 9, 17, 23, 27, 21, 16),
                (2, 5, 6, 8, 23, 13, 18, 27, 16, 26), (2, 6, 7, 8, 18, 14, 22, 27, 26, 15),
                (2, 3, 4, 8, 10, 11, 25, 27, 24, 20), (2, 7, 3, 8, 22, 19, 10, 27, 15, 24)]  # structured.jl:537-542
        per = [np.stack([p[q] for q in t], axis=1) for t in tets]    # 6 x (ncell, 10)
        conn = np.stack(per, axis=1).reshape(-1, 10)                 # six tets of a cell are consecutive
        return coords, boundary, conn
    raise ValueError(f"block: cannot discretize using shape {sh.name}")


class Mesh:
    """Flat mesh: coords (nnodes,3), conn (nelem,nn) int32, one cell shape, per-element tag index."""

    def __init__(self, *blocks, quiet=True, native=True):
        """``native``: single blocks are generated by the C++ mesher behind the ABI (``amaru_mesh_block``, csrc/mesher.cpp:
        same arrays, bit for bit, as the numpy path below, which stays for multi-block meshes and as the cross-check)."""
        bl = []
        for b in blocks:
            bl.extend(b if isinstance(b, (list, tuple)) else [b])
        if not bl:
            raise ValueError("Mesh: no blocks")
        shape = bl[0].cellshape
        if any(b.cellshape is not shape for b in bl):
            raise ValueError("Mesh: all blocks must use the same cell shape in this build")
        self.shape = shape
        self.ndim = max(b.ndim for b in bl)
        coords_all, conn_all, tag_all = [], [], []
        self.tags = []
        pointdict = {}
        n = 0
        for b in bl:
            if len(bl) == 1 and native:
                from . import lib as L
                c, conn = L.mesh_block(b.cellshape.id, b.c0, b.c1, b.nx, b.ny, b.nz)
                onb = None
            else:
                c, onb, conn = _split_block(b)
            if len(bl) == 1:
                gid = np.arange(c.shape[0], dtype=np.int64)
                coords_all.append(c)
                n = c.shape[0]
            else:                                                    # merge boundary points (structured.jl:481-487)
                gid = np.empty(c.shape[0], dtype=np.int64)
                new_rows = []
                for i in range(c.shape[0]):
                    if onb[i]:
                        key = (c[i, 0], c[i, 1], c[i, 2])
                        j = pointdict.get(key)
                        if j is None:
                            j = n
                            pointdict[key] = j
                            new_rows.append(i)
                            n += 1
                        gid[i] = j
                    else:
                        gid[i] = n
                        new_rows.append(i)
                        n += 1
                coords_all.append(c[new_rows])
            conn_all.append(gid[conn])
            if b.tag not in self.tags:
                self.tags.append(b.tag)
            tag_all.append(np.full(conn.shape[0], self.tags.index(b.tag), dtype=np.int32))
        self.coords = np.ascontiguousarray(np.vstack(coords_all))
        self.conn = np.ascontiguousarray(np.vstack(conn_all).astype(np.int32))
        self.elem_tag = np.concatenate(tag_all)
        self._facets = None

    @property
    def nnodes(self):
        return self.coords.shape[0]

    @property
    def nelems(self):
        return self.conn.shape[0]

    def outer_facets(self, native=True):
        """Boundary facets (faces in 3D, edges in 2D): those seen once (mesh.jl:69-85).

        -> (facet_nodes (nf, nfn) int32 in the owner's facet_idxs order, owner element (nf,))"""
        if self._facets is None and native:
            from . import lib as L
            self._facets = L.outer_facets(self.shape.id, self.conn)     # C++ behind the ABI (csrc/mesher.cpp)
        if self._facets is None:
            sh = self.shape
            fl, ow = [], []
            for fi in sh.facet_idxs:
                fl.append(self.conn[:, fi])
                ow.append(np.arange(self.nelems))
            nfl = len(sh.facet_idxs)
            F = np.stack(fl, axis=1).reshape(-1, fl[0].shape[1])      # element-major, local face order
            O = np.stack(ow, axis=1).reshape(-1)
            key = np.sort(F, axis=1)
            _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
            once = cnt[inv.reshape(-1)] == 1
            self._facets = (np.ascontiguousarray(F[once]), O[once])
        return self._facets
