"""The BASELINE.json configurations as GPU parity cases (through the public solve() driver and the C ABI).

config 1 is in test_gpu_parity.py (QUAD8 20x10 cantilever, full size).  Here: config 2 at FULL size with size-independent
checks (the CPU oracle's direct solve cannot factorise 634 k dofs in test time), config 3 and config 4 at sizes the oracle
finishes in seconds, compared increment by increment with the oracle's restatement of mech_stage_solver!.
"""
import numpy as np
import pytest

from amaru_jl_b200 import (Block, BodyC, DruckerPrager, FEModel, LinearElastic, MechAnalysis, MechContext, MechSolid, Mesh,
                           NodeBC, SurfaceBC, VonMises, addstage, solve)
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    d = np.abs(b).max()
    return np.abs(a - b).max() / (d if d > 0 else 1.0)


def oracle_run(model, stages, **kw):
    om, Uacc = None, np.zeros((model.nnodes, model.ndim))
    last = None
    for bcs, nincs in stages:
        eqid, nu, setup = model.configure_dofs(bcs)
        Uex, Fex = model.get_bc_vals(eqid, setup)
        if om is None:
            om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
        om.eqid, om.nu = np.ascontiguousarray(eqid), nu
        last = O.mech_stage_solver(om, Uex, Fex, nincs=nincs, **kw)
        Uacc += last["U"][eqid]
    return last, Uacc, om


def test_config2_hex8_200k_full_size():
    """config 2: HEX8 100x50x40 (200 000 elements, 211 191 nodes, 633 573 dofs) linear elastic, z=0 clamped, tz=-10 on top,
    single static solve."""
    mesh = Mesh(Block([[0, 0, 0], [2, 1, 0.8]], nx=100, ny=50, nz=40, cellshape="HEX8", tag="solids"))
    model = FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=100.0, nu=0.2))], MechContext())
    assert model.nelems == 200000 and model.nnodes == 211191 and model.ndofs == 633573
    ana = MechAnalysis(model)
    addstage(ana, [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==0.8", SurfaceBC(tz=-10.0))], nincs=1)
    status = solve(ana, cg_rtol=1e-10)
    assert status.success and len(ana.stats) == 1                       # linear: one Newton iteration
    X = model.coords
    # global equilibrium: the reactions balance the applied load (-10 * 2 * 1)
    assert abs(model.F[:, 2].sum()) < 1e-6 * 20.0
    base = np.abs(X[:, 2]) < 1e-9
    assert abs(model.F[base, 2].sum() - 20.0) < 1e-6 * 20.0
    assert abs(model.F[base, 0].sum()) < 1e-6 and abs(model.F[base, 1].sum()) < 1e-6
    # symmetry of the solution about the planes x = 1 and y = 0.5 (structured mesh, symmetric load)
    key = {tuple(np.round(p, 6)): i for i, p in enumerate(X)}
    idx = np.arange(0, model.nnodes, 97)
    mx = np.array([key[(round(2 - X[i, 0], 6), round(X[i, 1], 6), round(X[i, 2], 6))] for i in idx])
    assert np.abs(model.U[idx, 2] - model.U[mx, 2]).max() < 1e-7 * np.abs(model.U[:, 2]).max()
    assert np.abs(model.U[idx, 0] + model.U[mx, 0]).max() < 1e-7 * np.abs(model.U[:, 2]).max()
    # the mean vertical strain is close to the 1-D estimate sigma/E' (confined column), sanity of magnitude
    top = np.abs(X[:, 2] - 0.8) < 1e-9
    assert -0.12 < model.U[top, 2].mean() < -0.04
    st = ana.stats[0]
    assert st["cg_relres"] <= 1e-10 and st["cg_iters"] > 10


def test_config3_footing_reduced_vs_oracle():
    """config 3 at 8^3 HEX20: von Mises footing, prescribed uz on the central patch, 10 equal increments, maxits=5."""
    def mk():
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=8, ny=8, nz=8, cellshape="HEX20", tag="solids"))
        return FEModel(mesh, [("solids", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0.0))], MechContext())
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==1 and x>=0.375 and x<=0.625 and y>=0.375 and y<=0.625", NodeBC(uz=-0.004))]
    model = mk()
    ana = MechAnalysis(model)
    addstage(ana, bcs, nincs=10)
    status = solve(ana, cg_rtol=1e-12, maxits=5)
    ref = mk()
    r, Uacc, om = oracle_run(ref, [(bcs, 10)], maxits=5)
    assert status.success == r["success"]
    assert len(ana.stats) == r["its"]
    assert rel(model.U, Uacc) < 1e-7
    assert rel(model.state["sigma"], om.sig) < 1e-6
    assert (om.epa > 0).sum() > 0 and np.array_equal(model.state["epa"] > 0, om.epa > 0)


def test_config4_tet10_dp_slope_reduced_vs_oracle():
    """config 4 at 4^3 x 6 TET10: Drucker-Prager block under gravity body force (BodyC wz), autoinc off."""
    def mk():
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=4, ny=4, nz=4, cellshape="TET10", tag="solids"))
        return FEModel(mesh, [("solids", MechSolid, DruckerPrager, dict(E=100.0, nu=0.25, alpha=0.05, kappa=0.1))], MechContext())
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("x==0", NodeBC(ux=0)), ("y==0 || y==1", NodeBC(uy=0)), ("z>=0", BodyC(wz=-0.3))]
    model = mk()
    ana = MechAnalysis(model)
    addstage(ana, bcs, nincs=4)
    status = solve(ana, cg_rtol=1e-12, tol=1e-3)
    ref = mk()
    r, Uacc, om = oracle_run(ref, [(bcs, 4)], tol=1e-3)
    assert status.success == r["success"]
    assert len(ana.stats) == r["its"]
    assert rel(model.U, Uacc) < 1e-7
    assert rel(model.state["sigma"], om.sig) < 1e-6
