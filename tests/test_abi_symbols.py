"""CPU-side checks of the boundary: libamaru_b200.so loads, exports every symbol include/amaru_b200.h declares, the
ctypes table covers them all, and without a GPU the library refuses to create a model (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from amaru_jl_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "amaru_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amaru_[a-z_A-Z0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/amaru_b200.h but not exported"
        assert n in L.SYMBOLS, f"{n} has no ctypes signature in amaru_jl_b200/lib.py"
    assert set(L.SYMBOLS) == set(names)
    assert b"sm_100a" in lib.amaru_version()


def test_no_cpu_fallback_without_device():
    if L.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from amaru_jl_b200 import Block, FEModel, LinearElastic, MechAnalysis, MechContext, MechSolid, Mesh, NodeBC, addstage, solve
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=1, ny=1, nz=1, cellshape="HEX8", tag="s"))
    model = FEModel(mesh, [("s", MechSolid, LinearElastic, dict(E=1.0, nu=0.3))], MechContext())
    ana = MechAnalysis(model)
    addstage(ana, [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==1", NodeBC(fz=1))])
    with pytest.raises(L.AmaruStatus) as e:
        solve(ana)
    assert e.value.code == L.ERR_NO_DEVICE


def test_unsupported_features_are_refused_on_the_host():
    from amaru_jl_b200 import AmaruError, Block, FEModel, LinearElastic, MechContext, MechSolid, Mesh
    mesh = Mesh(Block([[0, 0], [1, 1]], nx=1, ny=1, cellshape="QUAD8", tag="s"))
    mesh3 = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=1, ny=1, nz=1, cellshape="HEX8", tag="s"))
    with pytest.raises(AmaruError):                      # axisymmetric / plane models are 2D
        FEModel(mesh3, [("s", MechSolid, LinearElastic, dict(E=1.0, nu=0.3))], MechContext(stressmodel="axisymmetric"))
    from amaru_jl_b200 import VonMises
    with pytest.raises(AmaruError):                      # plane stress: LinearElastic only (linear-elastic.jl:99-108)
        FEModel(mesh, [("s", MechSolid, VonMises, dict(E=1.0, nu=0.3, fy=1.0))], MechContext(stressmodel="planestress"))
    with pytest.raises(AmaruError):
        FEModel(mesh, [("s", object, LinearElastic, dict(E=1.0, nu=0.3))], MechContext())
    with pytest.raises(AmaruError):
        LinearElastic(E=-1.0)
