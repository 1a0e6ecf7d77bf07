"""Natural boundary conditions: SurfaceBC tx/ty/tz/tn and BodyC wx/wy/wz.

CPU: the entity-by-entity oracle (oracle/oracle_loads.py, a restatement of src/mech/elem/distributed.jl:76-217) is pinned
to the reference's known answers (test/mech/elem/elastic-hex8.jl:29-59: triangular face load `tx=3z`, body load) and to
closed-form resultants, and the host's vectorised get_bc_vals is checked against it.
GPU: the device load sets (amaru_loadset_*, csrc/loads.cu) against the oracle on the same facets, through the C ABI.
"""
import numpy as np
import pytest

from amaru_jl_b200 import (Block, BodyC, FEModel, LinearElastic, MechAnalysis, MechContext, MechSolid, Mesh, NodeBC,
                           SurfaceBC, addstage, solve)
from amaru_jl_b200.expr import evaluate
from oracle import oracle as O
from oracle import oracle_loads as OL

MATS = [("solids", MechSolid, LinearElastic, dict(E=100.0, nu=0.2))]


def make(shape, n=3, jitter=0.0, thickness=1.0):
    if shape in ("QUAD4", "QUAD8"):
        mesh = Mesh(Block([[0, 0], [2, 1]], nx=n, ny=n, cellshape=shape, tag="solids"))
        ctx = MechContext(stressmodel="planestrain")
    else:
        mesh = Mesh(Block([[0, 0, 0], [2, 1, 1.5]], nx=n, ny=n, nz=n, cellshape=shape, tag="solids"))
        ctx = MechContext()
    if jitter:
        X = mesh.coords                                          # smooth warp: curved facets, distorted cells
        x0, y0 = X[:, 0].copy(), X[:, 1].copy()
        X[:, 0] += jitter * np.sin(2.0 * y0 + 0.3)
        X[:, 1] += jitter * np.cos(1.5 * x0)
        if mesh.ndim == 3:
            X[:, 2] += jitter * np.sin(x0 + y0)
    return FEModel(mesh, MATS, ctx, thickness=thickness)


def oracle_F(model, eqid, setup, t=0.0):
    """get_bc_vals restricted to the distributed loads, one entity at a time through the oracle."""
    nd = model.ndim
    axi = model.ctx.stressmodel == "axisymmetric"
    F = np.zeros(eqid.size)
    for bc, target in setup:
        for key, val in bc.conds.items():
            if isinstance(bc, SurfaceBC) and key in ("tx", "ty", "tz", "tn"):
                fn, _ = target
                sid = model.shape.facet_shape.id
                X = OL.ip_coords(sid, model.coords, fn, nd)
                vals = np.broadcast_to(evaluate(val, x=X[:, 0], y=X[:, 1], z=X[:, 2], t=t), (X.shape[0],))
                OL.apply(sid, model.coords, fn, eqid, nd, model.thickness, ("tx", "ty", "tz", "tn").index(key), vals, F, axi=axi)
            elif isinstance(bc, BodyC) and key in ("wx", "wy", "wz"):
                nodes = model.conn[target]
                X = OL.ip_coords(model.shape.id, model.coords, nodes, nd)
                vals = np.broadcast_to(evaluate(val, x=X[:, 0], y=X[:, 1], z=X[:, 2]), (X.shape[0],))
                OL.apply(model.shape.id, model.coords, nodes, eqid, nd, model.thickness, ("wx", "wy", "wz").index(key), vals, F, axi=axi)
    return F


CASES = [
    ("QUAD4", [("y>=0.9", SurfaceBC(ty=-10.0)), ("x>=1.9", SurfaceBC(tn="3*y+1"))]),
    ("QUAD8", [("y>=0.9", SurfaceBC(ty="-0.1*x", tx=2.0)), ("x>=1.9", SurfaceBC(tn=-4.0)), ("x>=0", BodyC(wy=-2.5, wx="x*y"))]),
    ("HEX8", [("z>=1.4", SurfaceBC(tz=-10.0)), ("x>=1.9", SurfaceBC(tx="3*z", tn="1+y")), ("x>=0", BodyC(wz=-1.0))]),
    ("HEX20", [("z>=1.4", SurfaceBC(tz="-10*x*y", ty=1.0)), ("y<=0.1", SurfaceBC(tn=7.0)), ("z<=0.6", BodyC(wz="-x", wy=3.0))]),
    ("TET10", [("z>=1.4", SurfaceBC(tz=-10.0, tn="x+2*y")), ("x>=0", BodyC(wz=-0.3, wx="z"))]),
]


# ---------------------------------------------------------------------------------------------- CPU: pin the oracle
# reference test/mech/elem/elastic-hex8.jl:29-59 — triangular face load and body load known answers (atol 1e-5)
@pytest.mark.parametrize("extra,uz", [
    (("x==1", SurfaceBC(tx="3*z")), [0, 0, 0, 0, 1.51044, -2.4501, 1.4499, -2.31023]),
    (("x>=0", BodyC(wz=-1)), [0, 0, 0, 0, -0.5, -0.5, -0.5, -0.5]),
])
def test_oracle_loads_reference_known_answers(extra, uz):
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=1, ny=1, nz=1, cellshape="HEX8", tag="solid"))
    base = [("x==0 && y==0 && z==0", NodeBC(ux=0, uy=0)), ("x==1 && y==0 && z==0", NodeBC(uy=0)),
            ("x==0 && y==1 && z==0", NodeBC(ux=0)), ("z==0", NodeBC(uz=0))]
    model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=1.0, nu=0.3))], MechContext())
    eqid, nu, setup = model.configure_dofs(base + [extra])
    Uex, _ = model.get_bc_vals(eqid, setup)
    Fex = oracle_F(model, eqid, setup)                                  # loads from the oracle, not from the host path
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    res = O.mech_stage_solver(om, Uex, Fex, nincs=1)
    assert res["success"]
    assert np.abs(res["U"][eqid][:, 2] - np.array(uz)).max() < 1e-5


def test_oracle_loads_resultants():
    m = make("HEX20", 2)
    eqid, nu, setup = m.configure_dofs([("z==1.5", SurfaceBC(tz=-10.0)), ("x==2", SurfaceBC(tn=4.0)), ("x>=0", BodyC(wy=-2.0))])
    F = oracle_F(m, eqid, setup)[eqid]                                    # (nnodes, 3)
    assert abs(F[:, 2].sum() - (-10.0 * 2 * 1)) < 1e-12                   # traction x area
    assert abs(F[:, 0].sum() - (4.0 * 1 * 1.5)) < 1e-12                   # outward normal of x==2 is +x
    assert abs(F[:, 1].sum() - (-2.0 * 2 * 1 * 1.5)) < 1e-12              # body force x volume
    m2 = make("QUAD8", 2, thickness=0.25)                                 # thickness multiplies tractions and body forces
    eqid, nu, setup = m2.configure_dofs([("y==1", SurfaceBC(ty=-8.0)), ("x>=0", BodyC(wx=1.0))])
    F2 = oracle_F(m2, eqid, setup)[eqid]
    assert abs(F2[:, 1].sum() - (-8.0 * 2 * 0.25)) < 1e-12 and abs(F2[:, 0].sum() - 2 * 1 * 0.25) < 1e-12


@pytest.mark.parametrize("shape,bcs", CASES)
def test_host_get_bc_vals_matches_oracle(shape, bcs):
    m = make(shape, 3, jitter=0.03)
    eqid, nu, setup = m.configure_dofs(bcs)
    _, F = m.get_bc_vals(eqid, setup, t=0.0)
    Fo = oracle_F(m, eqid, setup)
    assert np.abs(Fo).max() > 0
    assert np.abs(F - Fo).max() <= 1e-13 * np.abs(Fo).max()


@pytest.mark.parametrize("shape", ["QUAD4", "QUAD8"])
def test_axisymmetric_loads(shape):
    """stressmodel = :axisymmetric: th = 2*pi*X[1] at every integration point of a facet / cell (distributed.jl:121,193).  The
    host integration equals the oracle's, and the resultants are pressure x lateral area / weight of the revolved body."""
    mesh = Mesh(Block([[1, 0], [3, 1]], nx=4, ny=3, cellshape=shape, tag="solids"))
    m = FEModel(mesh, MATS, MechContext(stressmodel="axisymmetric"))
    bcs = [("x==3", SurfaceBC(tx=-2.0)), ("y==1", SurfaceBC(ty="-0.5*x")), ("x>=0", BodyC(wy=-1.5))]
    eqid, nu, setup = m.configure_dofs(bcs)
    _, F = m.get_bc_vals(eqid, setup)
    Fo = oracle_F(m, eqid, setup)
    assert np.abs(F - Fo).max() <= 1e-13 * np.abs(Fo).max()
    m1 = FEModel(mesh, MATS, MechContext(stressmodel="axisymmetric"))
    eqid, nu, setup = m1.configure_dofs([("x==3", SurfaceBC(tx=-2.0))])
    _, F1 = m1.get_bc_vals(eqid, setup)
    assert abs(F1[eqid[:, 0]].sum() - (-2.0 * 2 * np.pi * 3 * 1)) < 1e-12        # pressure x (2 pi r h)
    eqid, nu, setup = m1.configure_dofs([("x>=0", BodyC(wy=-1.5))])
    _, F2 = m1.get_bc_vals(eqid, setup)
    assert abs(F2[eqid[:, 1]].sum() - (-1.5 * np.pi * (9 - 1) * 1)) < 1e-11       # weight of the hollow cylinder


def test_unsuitable_keys_are_refused():
    from amaru_jl_b200 import AmaruError
    m = make("QUAD8", 2)
    eqid, nu, setup = m.configure_dofs([("y==1", SurfaceBC(tz=1.0))])
    with pytest.raises(AmaruError):
        m.get_bc_vals(eqid, setup)
    eqid, nu, setup = m.configure_dofs([("x>=0", BodyC(wz=1.0))])
    with pytest.raises(AmaruError):
        m.get_bc_vals(eqid, setup)


# ---------------------------------------------------------------------------------------------- GPU: device load sets
@pytest.mark.gpu
@pytest.mark.parametrize("shape,bcs", CASES)
def test_device_loads_match_oracle(shape, bcs):
    from amaru_jl_b200 import lib as L
    m = make(shape, 3, jitter=0.03, thickness=1.0 if shape not in ("QUAD4", "QUAD8") else 0.7)
    eqid, nu, setup = m.configure_dofs(bcs + [("x==0", NodeBC(ux=0, uy=0))])
    dm = L.DeviceModel(m.flatten(), eqid, eqid.size, nu)
    try:
        l0 = dm.launches
        U, F = m.get_bc_vals(eqid, setup, t=0.0, device=dm)
        assert dm.launches > l0                                           # the CUDA path ran
        Fo = oracle_F(m, eqid, setup)
        assert np.abs(F - Fo).max() <= 1e-13 * np.abs(Fo).max()
        F2 = m.get_bc_vals(eqid, setup, t=0.0, device=dm)[1]              # cached load sets, bitwise repeatable
        assert np.array_equal(F, F2)
        # integration-point coordinates returned for the host-side expression evaluation
        for (ls, X), (bc, target) in zip(dm._loadsets.values(), [s for s in setup if not isinstance(s[0], NodeBC)]):
            if X is None:
                continue
            nodes = target[0] if isinstance(bc, SurfaceBC) else m.conn[target]
            sid = m.shape.facet_shape.id if isinstance(bc, SurfaceBC) else m.shape.id
            assert np.abs(X - OL.ip_coords(sid, m.coords, nodes, m.ndim)).max() < 1e-14
    finally:
        dm.close()


@pytest.mark.gpu
def test_device_loads_refuse_unsuitable_keys():
    from amaru_jl_b200 import lib as L
    m = make("QUAD8", 2)
    eqid, nu, setup = m.configure_dofs([("x==0", NodeBC(ux=0, uy=0))])
    dm = L.DeviceModel(m.flatten(), eqid, eqid.size, nu)
    try:
        fn, _ = m.mesh.outer_facets()
        ls = dm.loadset(m.shape.facet_shape.id, fn)
        F = np.zeros(eqid.size)
        with pytest.raises(L.AmaruStatus) as e:
            ls.apply("tz", 1.0, F)                                        # distributed.jl:88
        assert e.value.code == L.ERR_ARG and "2D" in e.value.message
        cells = dm.loadset(m.shape.id, m.conn)
        with pytest.raises(L.AmaruStatus):
            cells.apply(3, 1.0, F)                                        # tn on cells: distributed.jl:163
        with pytest.raises(L.AmaruStatus):
            dm.loadset(3, m.conn)                                         # HEX8 entities in a 2D analysis
    finally:
        dm.close()


@pytest.mark.gpu
def test_solve_with_device_loads_known_answers():
    """reference test/mech/elem/elastic-hex8.jl:29-59 through solve(): the loads now come from the device load sets."""
    for extra, uz in [(("x==1", SurfaceBC(tx="3*z")), [0, 0, 0, 0, 1.51044, -2.4501, 1.4499, -2.31023]),
                      (("x>=0", BodyC(wz=-1)), [0, 0, 0, 0, -0.5, -0.5, -0.5, -0.5])]:
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=1, ny=1, nz=1, cellshape="HEX8", tag="solid"))
        base = [("x==0 && y==0 && z==0", NodeBC(ux=0, uy=0)), ("x==1 && y==0 && z==0", NodeBC(uy=0)),
                ("x==0 && y==1 && z==0", NodeBC(ux=0)), ("z==0", NodeBC(uz=0))]
        model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=1.0, nu=0.3))], MechContext())
        ana = MechAnalysis(model)
        addstage(ana, base + [extra], nincs=1)
        assert solve(ana, cg_rtol=1e-12).success
        assert np.abs(model.U[:, 2] - np.array(uz)).max() < 1e-5
