"""Patch plan of the matrix-free operator (csrc/patches.cpp), checked on the host: amaru_patch_plan_check rebuilds the plan
amaru_create would use and verifies that every element sits in exactly one slot, the 8 elements of a group share no node,
the patch-local ids of both DMMA fragment orders name the element's nodes, first-touch / ghost flags are right and every
patch lists every earlier patch it shares a node with (the accumulation order at a node is then fixed)."""
import numpy as np
import pytest

from amaru_jl_b200 import Block, Mesh
from amaru_jl_b200 import lib as L
from amaru_jl_b200.shapes import SHAPES


def plan(shape, nx, ny, nz=0, jitter=0.0, nowned=None):
    if nz:
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=nx, ny=ny, nz=nz, cellshape=shape, tag="s"))
    else:
        mesh = Mesh(Block([[0, 0], [2, 1]], nx=nx, ny=ny, cellshape=shape, tag="s"))
    X = np.zeros((mesh.coords.shape[0], 3))
    X[:, :mesh.coords.shape[1]] = mesh.coords
    if jitter:
        X[:, :mesh.coords.shape[1]] += jitter / nx * np.random.default_rng(1).uniform(-1, 1, size=mesh.coords.shape)
    X = np.ascontiguousarray(X)
    conn = np.ascontiguousarray(mesh.conn, dtype=np.int32)
    st = np.zeros(6)
    rc = L.load().amaru_patch_plan_check(SHAPES[shape].id, X.shape[0], nowned or X.shape[0], X.ctypes.data, conn.shape[0],
                                         conn.ctypes.data, st.ctypes.data)
    return rc, dict(zip(("patches", "colours", "fill", "slots", "maxgroups", "deps"), st)), conn.shape[0]


@pytest.mark.parametrize("shape,dims,jitter", [
    ("HEX20", (8, 8, 8), 0.0), ("HEX20", (10, 9, 7), 0.0), ("HEX20", (6, 6, 6), 0.15), ("HEX8", (9, 7, 5), 0.1),
    ("TET10", (8, 8, 8), 0.0), ("TET10", (5, 4, 3), 0.1), ("QUAD8", (20, 10), 0.0), ("QUAD4", (17, 9), 0.1),
    ("HEX20", (1, 1, 1), 0.0), ("QUAD8", (1, 1), 0.0)])
def test_plan_invariants(shape, dims, jitter):
    rc, st, nelem = plan(shape, *dims, jitter=jitter)
    assert rc == 0, f"rule {rc} violated"
    assert st["slots"] % 8 == 0 and st["slots"] >= nelem and abs(st["fill"] - nelem / st["slots"]) < 1e-12
    assert 1 <= st["colours"] <= 64 and st["patches"] >= 1


def test_structured_bricks_fill_every_slot():
    """A block whose edge counts are multiples of the brick (4x4x4 hexahedra, 8x8 quadrilaterals) has no empty slot: 8 patch
    colours, 8 groups of 8 elements per hexahedral patch."""
    rc, st, nelem = plan("HEX20", 12, 8, 8)
    assert rc == 0 and st["fill"] == 1.0 and st["colours"] == 8 and st["maxgroups"] == 8 and st["patches"] == nelem // 64
    rc, st, nelem = plan("QUAD8", 32, 16)
    assert rc == 0 and st["fill"] == 1.0 and st["colours"] == 4 and st["patches"] == 8


def test_ghost_rows_are_flagged():
    rc, st, _ = plan("HEX8", 6, 6, 6, nowned=200)   # nodes >= 200 belong to a neighbour rank
    assert rc == 0
