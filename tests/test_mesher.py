"""Host-side model generation behind the ABI: the C++ structured mesher and dof numbering against the numpy
restatement of src/mesh/structured.jl / src/bc.jl:198-233 and the reference's own node counts (test/mesh/structured.jl:24-64)."""
import numpy as np
import pytest

from amaru_jl_b200 import lib as L
from amaru_jl_b200.mesh import Block, Mesh, _split_block


@pytest.mark.parametrize("shape,n", [("QUAD4", (5, 3, 0)), ("QUAD8", (4, 7, 0)), ("QUAD8", (1, 1, 0)), ("HEX8", (3, 4, 5)),
                                     ("HEX20", (4, 3, 5)), ("HEX20", (1, 1, 1)), ("TET10", (3, 2, 4)), ("HEX20", (17, 9, 11))])
def test_native_mesher_is_bit_identical_to_the_numpy_restatement(shape, n):
    if shape.startswith("QUAD"):
        b = Block([[0.1, 0.2], [2.3, 1.7]], nx=n[0], ny=n[1], cellshape=shape)
    else:
        b = Block([[0.1, 0.2, -0.3], [2.3, 1.7, 0.9]], nx=n[0], ny=n[1], nz=n[2], cellshape=shape)
    c, _, conn = _split_block(b)
    c2, conn2 = L.mesh_block(b.cellshape.id, b.c0, b.c1, b.nx, b.ny, b.nz)
    assert conn2.dtype == np.int32 and np.array_equal(conn, conn2)
    assert np.array_equal(c, c2)                                     # includes the 8-digit rounding (node.jl:57-61)
    m1, m2 = Mesh(b, native=True), Mesh(b, native=False)
    assert np.array_equal(m1.coords, m2.coords) and np.array_equal(m1.conn, m2.conn)


# reference test/mesh/structured.jl:24-64
@pytest.mark.parametrize("shape,nnodes", [("QUAD8", 341), ("HEX8", 1331), ("HEX20", 4961), ("TET10", 9261)])
def test_native_mesher_node_counts(shape, nnodes):
    sid = {"QUAD8": 2, "HEX8": 3, "HEX20": 4, "TET10": 5}[shape]
    c, conn = L.mesh_block(sid, (0, 0, 0), (1, 1, 0 if shape == "QUAD8" else 1), 10, 10, 0 if shape == "QUAD8" else 10)
    assert c.shape[0] == nnodes and conn.min() == 0 and conn.max() == nnodes - 1
    assert np.unique(conn).size == nnodes
    assert np.unique(np.round(c, 8), axis=0).shape[0] == nnodes      # no duplicated points


def test_native_mesher_refuses_bad_input():
    with pytest.raises(L.AmaruError):
        L.mesh_block(4, (0, 0, 0), (1, 1, 1), 0, 1, 1)
    with pytest.raises(L.AmaruError):
        L.mesh_block(103, (0, 0, 0), (1, 1, 1), 1, 1, 1)


def test_configure_dofs_unknowns_first_stable():
    rng = np.random.default_rng(0)
    presc = rng.random((1000, 3)) < 0.2
    eq, nu = L.configure_dofs(presc)
    flat = presc.reshape(-1)
    order = np.concatenate((np.nonzero(~flat)[0], np.nonzero(flat)[0]))      # bc.jl:220-224
    ref = np.empty(flat.size, dtype=np.int64)
    ref[order] = np.arange(flat.size)
    assert nu == int((~flat).sum()) and np.array_equal(eq.reshape(-1), ref)
    eq2, nu2 = L.configure_dofs(np.zeros((5, 2), dtype=bool))
    assert nu2 == 10 and np.array_equal(eq2.reshape(-1), np.arange(10))


@pytest.mark.parametrize("shape,n", [("QUAD4", 7), ("QUAD8", 5), ("HEX8", 6), ("HEX20", 5), ("TET10", 4), ("HEX20", 1)])
def test_native_outer_facets_match_the_numpy_restatement(shape, n):
    """get_outer_facets (mesh.jl:69-85): same facets, same order (cell, local facet), same node order, same owners."""
    if shape.startswith("QUAD"):
        b = Block([[0, 0], [1, 1]], nx=n, ny=n + 1, cellshape=shape)
    else:
        b = Block([[0, 0, 0], [1, 1, 1]], nx=n, ny=n + 1, nz=n + 2, cellshape=shape)
    fn, ow = Mesh(b).outer_facets(native=False)
    fn2, ow2 = L.outer_facets(Mesh(b).shape.id, Mesh(b).conn)
    assert np.array_equal(fn, fn2) and np.array_equal(ow, ow2)
    if not shape.startswith("QUAD"):                                   # closed surface: every boundary edge is shared twice
        nx, ny, nz = n, n + 1, n + 2
        per_cell_face = 2 if shape == "TET10" else 1
        assert fn2.shape[0] == per_cell_face * 2 * (nx * ny + ny * nz + nx * nz)
