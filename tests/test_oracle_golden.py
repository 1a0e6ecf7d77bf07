"""Pins the CPU oracle (oracle/) and the host logic (mesher, dof numbering, BC vectors) to every known answer the
reference's own tests hold for the mechanical path.  CPU only."""
import numpy as np
import pytest

from amaru_jl_b200 import shapes as S
from amaru_jl_b200.mesh import Block, Mesh
from amaru_jl_b200.model import (BodyC, DruckerPrager, FEModel, LinearElastic, MechContext, MechSolid, NodeBC,
                                 SurfaceBC, VonMises)
from oracle import oracle as O


def run(model, bcs, om=None, **kw):
    eqid, nu, setup = model.configure_dofs(bcs)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    om = om or O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    om.eqid, om.nu = np.ascontiguousarray(eqid), nu
    return O.mech_stage_solver(om, Uex, Fex, **kw), eqid, om


# reference test/mesh/structured.jl:24-64
@pytest.mark.parametrize("shape,nnodes", [("QUAD8", 341), ("HEX8", 1331), ("HEX20", 4961), ("TET10", 9261)])
def test_structured_node_counts(shape, nnodes):
    if shape == "QUAD8":
        m = Mesh(Block([[0, 0], [1, 1]], nx=10, ny=10, cellshape=shape))
    else:
        m = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=10, ny=10, nz=10, cellshape=shape))
    assert m.nnodes == nnodes
    assert m.conn.min() == 0 and m.conn.max() == nnodes - 1


# reference test/mesh/shape/shape_deriv.jl:6-59
@pytest.mark.parametrize("sh", [S.QUAD4, S.QUAD8, S.HEX8, S.HEX20, S.TET10])
def test_shape_functions(sh):
    for j in range(sh.npoints):                        # N_i(node_j) = delta_ij
        R = np.zeros(3)
        R[:sh.ndim] = sh.nat_coords[j]
        e = np.zeros(sh.npoints)
        e[j] = 1.0
        assert np.abs(O.shape_func(sh.id, R) - e).max() < 1e-10
        assert np.abs(sh.func(R) - e).max() < 1e-10
    for q in O.quadrature(sh.id):                      # partition of unity + FD derivative at the default IPs
        R = q[:3]
        assert abs(O.shape_func(sh.id, R).sum() - 1.0) < 1e-10
        D = O.shape_deriv(sh.id, R)
        h = 1e-6
        for d in range(sh.ndim):
            Rp, Rm = R.copy(), R.copy()
            Rp[d] += h
            Rm[d] -= h
            fd = (O.shape_func(sh.id, Rp) - O.shape_func(sh.id, Rm)) / (2 * h)
            assert np.abs(fd - D[:, d]).max() < 1e-6
        assert np.abs(sh.deriv(R) - D).max() < 1e-14   # product host tables == oracle
    assert np.allclose(sh.quadrature, O.quadrature(sh.id), rtol=0, atol=0)


# reference test/tools/tensors.jl:5-20
def test_tensor_invariants():
    s = np.array([10., 20., 30., 4., 5., 6.])
    assert s[:3].sum() == 60
    d = O.dev(s)
    assert np.abs(O.dev(d) - d).max() < 1e-10
    assert abs(O.J2(s) - O.J2(d)) < 1e-10
    assert abs(d[:3].sum()) < 1e-10


# reference test/mech/elem/elastic-quad4.jl:10-46
def test_elastic_quad4():
    mesh = Mesh(Block([[0, 0], [1, 1]], nx=1, ny=1, cellshape="QUAD4", tag="solid"))
    model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=1.0, nu=0.25))],
                    MechContext(stressmodel="planestrain"))
    bcs = [("x==0.", SurfaceBC(ux=0.)), ("y==0.", SurfaceBC(uy=0)), ("y==1.", SurfaceBC(ty=-1.))]
    res, eqid, _ = run(model, bcs, nincs=1)
    dis = np.array([[0.0, 0.0], [0.3125, 0.0], [0.0, -0.9375], [0.3125, -0.9375]])
    assert res["success"]
    assert np.abs(res["U"][eqid] - dis).max() < 1e-5


def test_elastic_quad4_planestress():
    """Plane stress (calcDe branch linear-elastic.jl:99-108) on the plate of elastic-quad4.jl: the state is uniaxial
    (σyy = -1, σxx = σzz = 0), so uy = -1/E and ux = ν/E exactly; the in-plane block of the plane-stress matrix equals the
    plane-strain matrix of E* = E(1+2ν)/(1+ν)², ν* = ν/(1+ν) — the identity the CUDA path relies on."""
    mesh = Mesh(Block([[0, 0], [1, 1]], nx=2, ny=2, cellshape="QUAD8", tag="solid"))
    model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=1.0, nu=0.25))], MechContext(stressmodel="planestress"))
    bcs = [("x==0.", SurfaceBC(ux=0.)), ("y==0.", SurfaceBC(uy=0)), ("y==1.", SurfaceBC(ty=-1.))]
    res, eqid, om = run(model, bcs, nincs=1)
    assert res["success"]
    U = res["U"][eqid]
    assert np.abs(U[:, 0] - 0.25 * model.coords[:, 0]).max() < 1e-12
    assert np.abs(U[:, 1] + 1.0 * model.coords[:, 1]).max() < 1e-12
    assert np.abs(om.sig - np.array([0, -1.0, 0, 0, 0, 0])).max() < 1e-12
    E, nu = 3.7, 0.31
    Es, nus = E * (1 + 2 * nu) / (1 + nu) ** 2, nu / (1 + nu)
    ip = [0, 1, 5]
    assert np.abs(O.calcDe_planestress(E, nu)[np.ix_(ip, ip)] - O.calcDe(Es, nus)[np.ix_(ip, ip)]).max() < 1e-14 * E
    assert np.abs(O.calcDe_planestress(E, nu)[2]).max() == 0.0


# reference test/mech/elem/axisymmetric.jl:4-61 (the axisymmetric half; its 3D half needs `revolve`, outside the path):
# solid cylinder r, y in [0, 1], E = 100, nu = 0.2, ty = -10 on top -> uniaxial state, (ux, uy)(1, 1) = (nu*10/E, -10/E)
@pytest.mark.parametrize("shape", ["QUAD4", "QUAD8"])
def test_axisymmetric_reference_case(shape):
    mesh = Mesh(Block([[0, 0], [1, 1]], nx=4, ny=4, cellshape=shape, tag="solids"))
    model = FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=100.0, nu=0.2))], MechContext(stressmodel="axisymmetric"))
    bcs = [("x==0", SurfaceBC(ux=0)), ("y==0", SurfaceBC(uy=0)), ("y==1", SurfaceBC(ty=-10))]
    res, eqid, om = run(model, bcs, nincs=1)
    assert res["success"]
    U = res["U"][eqid]
    node = model.select_nodes("x==1 && y==1")[0]
    assert np.abs(U[node] - np.array([0.02, -0.1])).max() < 1e-10      # the reference holds its 3D twin to 1e-3
    assert np.abs(om.sig - np.array([0, -10.0, 0, 0, 0, 0])).max() < 1e-9


def test_axisymmetric_lame_cylinder():
    """stressmodel = :axisymmetric (hoop row of B, th = 2*pi*r: mech-solid.jl:94-108,143; distributed.jl:121) against the Lamé
    solution of a thick-walled cylinder (a = 1, b = 2) under internal pressure with plane-strain ends:
    u_r = p a² (1+ν) / (E (b²-a²)) · ((1-2ν) r + b²/r); σ_θθ(a) = p (a²+b²)/(b²-a²)."""
    E, nu, p, a, b = 1000.0, 0.3, 1.0, 1.0, 2.0
    mesh = Mesh(Block([[a, 0], [b, 0.2]], nx=24, ny=1, cellshape="QUAD8", tag="solid"))
    model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=E, nu=nu))], MechContext(stressmodel="axisymmetric"))
    bcs = [("y==0", NodeBC(uy=0)), ("y==0.2", NodeBC(uy=0)), (f"x=={a}", SurfaceBC(tx=p))]
    res, eqid, om = run(model, bcs, nincs=1)
    assert res["success"]
    U = res["U"][eqid]
    r = model.coords[:, 0]
    ur = p * a * a * (1 + nu) / (E * (b * b - a * a)) * ((1 - 2 * nu) * r + b * b / r)
    assert np.abs(U[:, 0] - ur).max() < 2e-6 * np.abs(ur).max() and np.abs(U[:, 1]).max() < 1e-12
    # hoop stress at the integration points (Mandel component 3 = θθ): Lamé σ_θθ = p a²/(b²-a²) (1 + b²/r²)
    ipr = model.ip_coords()[:, 0]
    sth = p * a * a / (b * b - a * a) * (1 + b * b / ipr ** 2)
    assert np.abs(om.sig[:, 2] - sth).max() < 2e-4 * sth.max()


# reference test/mech/elem/elastic-hex8.jl:10-75 (nodal, triangular face and volume load cases)
@pytest.mark.parametrize("extra,uz", [
    (("z==1", NodeBC(fz=1)), [0, 0, 0, 0, 4.0, 4.0, 4.0, 4.0]),
    (("x==1", SurfaceBC(tx="3*z")), [0, 0, 0, 0, 1.51044, -2.4501, 1.4499, -2.31023]),
    (("x>=0", BodyC(wz=-1)), [0, 0, 0, 0, -0.5, -0.5, -0.5, -0.5]),
])
def test_elastic_hex8(extra, uz):
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=1, ny=1, nz=1, cellshape="HEX8", tag="solid"))
    base = [("x==0 && y==0 && z==0", NodeBC(ux=0, uy=0)), ("x==1 && y==0 && z==0", NodeBC(uy=0)),
            ("x==0 && y==1 && z==0", NodeBC(ux=0)), ("z==0", NodeBC(uz=0))]
    model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=1.0, nu=0.3))], MechContext())
    res, eqid, _ = run(model, base + [extra], nincs=1)
    assert res["success"]
    assert np.abs(res["U"][eqid][:, 2] - np.array(uz)).max() < 1e-5


# reference test/mech/elem/elastic-elems.jl:6-68
@pytest.mark.parametrize("shape", ["QUAD8", "TET10", "HEX8", "HEX20"])
def test_elastic_elems(shape):
    mats = [("solids", MechSolid, LinearElastic, dict(E=100.0, nu=0.2))]
    if shape == "QUAD8":
        mesh = Mesh(Block([[0, 0], [1, 1]], nx=2, ny=2, cellshape=shape, tag="solids"))
        model = FEModel(mesh, mats, MechContext())
        res, eqid, _ = run(model, [("y==0", SurfaceBC(ux=0, uy=0)), ("y==1", SurfaceBC(ty=-10.))])
        top = model.select_nodes("y==1")[0]
        assert np.abs(res["U"][eqid][top] - np.array([-0.012, -0.095])).max() < 4e-2
    else:
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=2, ny=2, nz=2, cellshape=shape, tag="solids"))
        model = FEModel(mesh, mats, MechContext())
        res, eqid, _ = run(model, [("z==0", SurfaceBC(ux=0, uy=0, uz=0)), ("x==0 || x==1", SurfaceBC(ux=0)),
                                   ("z==1", SurfaceBC(tz=-10.))])
        top = model.select_nodes("z==1")[0]
        assert np.abs(res["U"][eqid][top][1:] - np.array([-0.012, -0.095])).max() < 1e-2
    assert res["success"]


# reference test/mech/mat/vm-3d.jl:12-37 : plastic moment of the cantilever, fz ~ -30 +- 0.7
def test_vm_3d_cantilever():
    th = 0.05
    mesh = Mesh(Block([[0, 0, -0.05], [0.05, 1.0, 0.05]], nx=1, ny=50, nz=2, cellshape="HEX20"))
    model = FEModel(mesh, [("bulks", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0))], MechContext())
    bcs = [("y==0", NodeBC(uy=0)), ("y==0 && z==0", NodeBC(uz=0)), (f"x=={th/2} && y==0 && z==0", NodeBC(ux=0)),
           (f"x=={th/2} && y==1 && z==0", NodeBC(uz=-0.08))]
    res, eqid, _ = run(model, bcs, nincs=20, nouts=1, autoinc=True)
    n = model.select_nodes(f"x=={th/2} && y==1 && z==0")[0]
    assert res["success"]
    assert abs(res["F"][eqid[n, 2]] - (-30.0)) < 0.7


# reference test/mech/mat/dp.jl:5-37 : two stages (load, unload), IP state carried over; `.success`
def test_drucker_prager_two_stages():
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 0.5]], nx=2, ny=2, nz=2, tag="solids"))
    model = FEModel(mesh, [("solids", MechSolid, DruckerPrager, dict(E=100., nu=0.25, alpha=0.05, kappa=0.1))],
                    MechContext())
    bcs = [("z==0.0", NodeBC(ux=0, uy=0, uz=0)), ("z==0.5", NodeBC(uz=-0.033)),
           ("x==0 || x==1.0", NodeBC(ux=0, uy=0)), ("y==0 || y==1.0", NodeBC(ux=0, uy=0))]
    r1, _, om = run(model, bcs, nincs=10, tol=1e-2, autoinc=True)
    bcs[1] = ("z==0.5", NodeBC(uz=+0.008))
    r2, _, om = run(model, bcs, om=om, nincs=10, tol=1e-2, autoinc=True)
    assert r1["success"] and r2["success"]
    assert om.epa.max() > 0                      # the load stage yields
