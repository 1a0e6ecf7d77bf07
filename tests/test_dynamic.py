"""Next tier: consistent mass + Newmark driver.  CPU checks pin the oracle's restatement; the GPU test
compares solve_dynamic (C ABI) with the oracle on the reference's own dynamic scenario (test/dynamic/dyn-solid.jl)."""
import numpy as np
import pytest

from amaru_jl_b200 import Block, FEModel, LinearElastic, MechContext, MechSolid, Mesh, NodeBC
from amaru_jl_b200.model import DynamicAnalysis, addstage
from oracle import oracle as O
from oracle import oracle_dyn as OD


def beam(ny=10):
    mesh = Mesh(Block([[0, 0, 0], [0.2, 2.0, 0.2]], nx=1, ny=ny, nz=1, cellshape="HEX8", tag="solids"))
    return FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=30e6, nu=0.2, rho=24.0))], MechContext())


BCS = [("y==0 && z==0", NodeBC(ux=0, uy=0, uz=0)), ("y==2 && z==0", NodeBC(uz=0)), ("y==1 && z==0.2", NodeBC(fz=-10))]


@pytest.mark.parametrize("shape", ["QUAD8", "HEX8", "HEX20", "TET10"])
def test_oracle_mass_matrix_total_mass(shape):
    """Consistent mass (mech-solid.jl:169-205): every direction block sums to the total mass rho*V*th, M is symmetric."""
    if shape == "QUAD8":
        mesh = Mesh(Block([[0, 0], [2, 1]], nx=3, ny=2, cellshape=shape, tag="s"))
        vol, th = 2.0, 0.5
    else:
        mesh = Mesh(Block([[0, 0, 0], [1, 2, 1.5]], nx=2, ny=2, nz=2, cellshape=shape, tag="s"))
        vol, th = 3.0, 1.0
    model = FEModel(mesh, [("s", MechSolid, LinearElastic, dict(E=1.0, nu=0.3, rho=7.0))], MechContext(), thickness=th)
    assert np.all(model.elem_rho == 7.0)
    eqid, nu, _ = model.configure_dofs([("x==0", NodeBC(ux=0))])
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    st, M = om.mount_M(model.elem_rho, filter_eps=False)
    assert st == 0
    for d in range(model.ndim):
        e = np.zeros(eqid.size)
        e[eqid[:, d]] = 1.0
        assert abs(e @ (M @ e) - 7.0 * vol * th) < 1e-10
    assert abs(M - M.T).max() < 1e-14


def test_oracle_newmark_static_limit():
    """A suddenly applied load held over one very long average-acceleration step: Kp -> K and the load vector becomes
    Fex + M*A0 = 2*Fex (A0 = M^-1 Fex from the initial-acceleration solve, dyn-solver.jl:289,378), i.e. exactly twice the
    static displacement — the classic dynamic amplification of a step load."""
    model = beam(6)
    eqid, nu, setup = model.configure_dofs(BCS)
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    r = OD.dyn_stage_solver(om, lambda t: (Uex.copy(), Fex.copy()), model.elem_rho, tspan=1e4, nincs=1, tol=1e-6)
    om2 = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    rs = O.mech_stage_solver(om2, Uex, Fex, nincs=1)
    assert r["success"] and np.abs(r["U"] - 2.0 * rs["U"]).max() < 1e-6 * np.abs(rs["U"]).max()


@pytest.mark.gpu
def test_solve_dynamic_matches_oracle():
    """reference test/dynamic/dyn-solid.jl scenario (10 steps instead of 1): Rayleigh alpha=4.2038, beta=174.2803e-6."""
    from amaru_jl_b200.dyn_solver import solve_dynamic
    model = beam()
    ana = DynamicAnalysis(model)
    addstage(ana, BCS, tspan=0.1, nincs=10, nouts=1)
    status = solve_dynamic(ana, alpha=4.2038, beta=174.2803e-6, tol=1e-6, cg_rtol=1e-13)
    ref = beam()
    eqid, nu, setup = ref.configure_dofs(BCS)
    om = O.OracleModel(ref.flatten(), eqid, eqid.size, nu)
    r = OD.dyn_stage_solver(om, lambda t: ref.get_bc_vals(eqid, setup, t), ref.elem_rho, tspan=0.1, nincs=10, alpha=4.2038,
                            beta=174.2803e-6, tol=1e-6)
    assert status.success and r["success"] and len(ana.stats) == r["its"]

    def rel(a, b):
        return np.abs(a - b).max() / np.abs(b).max()
    assert rel(model.U, r["U"][eqid]) < 1e-7
    assert rel(model.V, r["V"][eqid]) < 1e-7
    assert rel(model.A, r["A"][eqid]) < 1e-6
    assert rel(model.F, r["F"][eqid]) < 1e-6
    assert np.abs(model.U).max() > 0
