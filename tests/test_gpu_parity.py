"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same inputs.  Needs a B200: -m gpu.

Tolerances are BASELINE.json's: CSR pattern and dof numbering exact; K, f_int and IP stresses within 1e-12 relative
(max-norm); displacements from the PCG within 1e-8 relative of the direct solve.
"""
import numpy as np
import pytest

from amaru_jl_b200 import (Block, BodyC, DruckerPrager, FEModel, LinearElastic, MechAnalysis, MechContext, MechSolid,
                           Mesh, NodeBC, SurfaceBC, VonMises, addstage, solve)
from amaru_jl_b200 import lib as L
from oracle import oracle as O

pytestmark = pytest.mark.gpu

MATS = {
    "le": (LinearElastic, dict(E=100.0, nu=0.2)),
    "vm": (VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=1e6)),
    "vm0": (VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0.0)),
    "dp": (DruckerPrager, dict(E=100.0, nu=0.25, alpha=0.05, kappa=0.1, H=0.0)),
}


def make_model(shape, n, mat="le", jitter=0.0, seed=0, mixed=False):
    if shape in ("QUAD4", "QUAD8"):
        mesh = Mesh(Block([[0, 0], [2, 1]], nx=2 * n, ny=n, cellshape=shape, tag="solids"))
    else:
        mesh = Mesh(Block([[0, 0, 0], [1, 1.5, 2]], nx=n, ny=n, nz=n + 1, cellshape=shape, tag="solids"))
    if jitter:
        rng = np.random.default_rng(seed)
        c = mesh.coords
        lo, hi = c.min(0), c.max(0)
        interior = np.all((c > lo + 1e-9) | (hi - lo == 0), axis=1) & np.all((c < hi - 1e-9) | (hi - lo == 0), axis=1)
        h = (hi - lo) / (2 * n + 2)
        c[interior] += rng.uniform(-jitter, jitter, (interior.sum(), 3)) * h
        if mesh.ndim == 2:
            c[:, 2] = 0.0
        mesh.coords[...] = np.round(c, 8)
    mty, par = MATS[mat]
    binds = [("solids", MechSolid, mty, par)]
    if mixed:   # second material on the upper half: two material ids in one batch
        binds.append(("y>=0.4" if mesh.ndim == 2 else "z>=0.9", MechSolid, LinearElastic, dict(E=50.0, nu=0.3)))
    thickness = 0.7 if mesh.ndim == 2 else 1.0
    return FEModel(mesh, binds, MechContext(), thickness=thickness)


def clamp_bcs(model):
    if model.ndim == 2:
        return [("x==0", NodeBC(ux=0, uy=0)), ("x==2", NodeBC(uy=-0.01))]
    return [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==2", NodeBC(uz=-0.02, ux=0.003))]


def pair(model, bcs):
    eqid, nu, setup = model.configure_dofs(bcs)
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
    return om, dm, eqid, nu, setup


def rel(a, b):
    d = np.abs(b).max()
    return np.abs(a - b).max() / (d if d > 0 else 1.0)


def check_K(om, dm, tol=1e-12):
    dm.assemble_K()
    rp, ci, val = dm.get_csr()
    st, K = om.mount_K(filter_eps=False)
    assert st == 0
    K = K.tocsr()
    K.sort_indices()
    srp, sci = om.symbolic_csr()
    assert np.array_equal(rp, srp) and np.array_equal(ci, sci), "symbolic CSR pattern differs"
    assert np.array_equal(K.indptr, rp) and np.array_equal(K.indices, ci)
    assert rel(val, K.data) < tol
    # the reference's stored pattern (|v| < eps dropped, mech-solver.jl:90) is a subset whose complement is ~0
    st, Kf = om.mount_K(filter_eps=True)
    Kf = Kf.tocsr()
    assert Kf.nnz <= K.nnz
    return K


SHAPES = [("QUAD4", 3), ("QUAD8", 3), ("HEX8", 3), ("HEX20", 2), ("TET10", 2)]


@pytest.mark.parametrize("shape,n", SHAPES)
def test_K_pattern_and_values_elastic(shape, n):
    model = make_model(shape, n, "le", jitter=0.2, mixed=True)
    om, dm, *_ = pair(model, clamp_bcs(model))
    check_K(om, dm)
    dm.close()


@pytest.mark.parametrize("shape,n", SHAPES)
@pytest.mark.parametrize("mat", ["le", "vm", "vm0", "dp"])
def test_update_state_and_tangent(shape, n, mat):
    """update_state! parity (f_int, sigma, eps, epa, dlam), then mount_K on the updated (trial) state."""
    model = make_model(shape, n, mat, jitter=0.15, seed=1)
    om, dm, eqid, nu, _ = pair(model, clamp_bcs(model))
    rng = np.random.default_rng(2)
    E = MATS[mat][1]["E"]
    scale = {"le": 1e-3, "vm": 4e-3, "vm0": 4e-3, "dp": 5e-3}[mat]
    for step in range(3):                                   # later steps start from a plastic, non-zero state
        dU = rng.uniform(-1, 1, eqid.size) * scale * (step + 1)
        dFo, st = om.update_state(dU)
        assert st == 0
        dF = dm.update_state(dU)
        s = dm.get_state()
        if mat != "le":
            assert (om.dlam > 0).sum() > 0, "test must exercise the plastic branch"
        assert rel(dF, dFo) < 1e-12
        assert rel(s["sigma"], om.sig) < 1e-12
        assert rel(s["eps"], om.eps) < 1e-12
        assert rel(s["epa"], om.epa) < 1e-12
        assert rel(s["dlam"], om.dlam) < 1e-11
        assert np.array_equal(s["dlam"] > 0, om.dlam > 0)
        check_K(om, dm, tol=1e-12)
        assert rel(dm.internal_forces(), om.internal_forces()) < 1e-12
    dm.close()


def test_dp_apex_branch():
    """Hydrostatic tension drives Drucker-Prager to the apex return (drucker-prager.jl:135-139) on the second step
    (the cone/apex switch uses the previous Δγ, :130)."""
    model = make_model("HEX8", 2, "dp")
    om, dm, eqid, nu, _ = pair(model, clamp_bcs(model))
    X = model.coords
    U = np.zeros(eqid.size)
    for k, (amp, sh) in enumerate(((0.02, 0.01), (0.05, 0.0005))):
        D = amp * X[:, :3].copy()                           # dilation + a little shear (keeps J2tr well conditioned)
        D[:, 0] += sh * X[:, 1]
        U[eqid.reshape(-1)] = D.reshape(-1)
        dFo, st = om.update_state(U)
        dF = dm.update_state(U)
        s = dm.get_state()
        assert st == 0 and rel(dF, dFo) < 1e-12 and rel(s["sigma"], om.sig) < 1e-12 and rel(s["dlam"], om.dlam) < 1e-12
    j2 = np.array([O.J2(x) for x in om.sig])
    assert (j2 < 1e-20).any(), "apex return not reached"
    check_K(om, dm, tol=1e-12)
    dm.close()


@pytest.mark.parametrize("shape,n", [("QUAD8", 4), ("HEX8", 4), ("HEX20", 3), ("TET10", 3)])
@pytest.mark.parametrize("precond", ["jacobi", "block-jacobi"])
def test_solve_matches_direct(shape, n, precond):
    model = make_model(shape, n, "le", jitter=0.1)
    bcs = clamp_bcs(model) + [("z==2" if model.ndim == 3 else "y==1", SurfaceBC(**({"tz": -3.0} if model.ndim == 3 else {"ty": -3.0})))]
    om, dm, eqid, nu, setup = pair(model, bcs)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    st, K = om.mount_K()
    dm.assemble_K()
    U, F = Uex.copy(), Fex.copy()
    iters, rr = dm.solve(U, F, cg_rtol=1e-12, precond=L.PRECOND[precond])
    Uo, Fo = Uex.copy(), Fex.copy()
    ok, _ = O.solve_system(K, Uo, Fo, nu)
    assert ok and iters > 0 and rr <= 1e-12
    assert rel(U, Uo) < 1e-8                                 # displacements incl. untouched prescribed part
    assert np.array_equal(U[nu:], Uex[nu:])                  # solve_system! leaves U2 alone (solver.jl:74)
    assert np.array_equal(F[:nu], Fex[:nu])                  # ... and F1 (solver.jl:75)
    assert rel(F[nu:], Fo[nu:]) < 1e-8                       # reactions
    dm.close()


def test_state_roundtrip_backup_restore():
    model = make_model("HEX20", 2, "vm")
    om, dm, eqid, *_ = pair(model, clamp_bcs(model))
    rng = np.random.default_rng(3)
    n = dm.nip_total
    sig, eps, epa, dl = rng.normal(size=(n, 6)), rng.normal(size=(n, 6)), rng.uniform(size=n), rng.uniform(size=n)
    dm.set_state(sig, eps, epa, dl)
    s = dm.get_state()
    assert np.array_equal(s["sigma"], sig) and np.array_equal(s["eps"], eps)
    assert np.array_equal(s["epa"], epa) and np.array_equal(s["dlam"], dl)
    dm.state_backup()
    dm.update_state(rng.uniform(-1, 1, eqid.size) * 1e-3)
    assert not np.array_equal(dm.get_state()["sigma"], sig)
    dm.state_restore()
    assert np.array_equal(dm.get_state()["sigma"], sig)
    dm.close()


def test_assembly_is_deterministic():
    model = make_model("HEX20", 3, "vm0", jitter=0.1)
    om, dm, eqid, *_ = pair(model, clamp_bcs(model))
    dm.update_state(np.random.default_rng(5).uniform(-1, 1, eqid.size) * 5e-3)
    dm.assemble_K()
    a = dm.get_csr()[2].copy()
    f1 = dm.internal_forces().copy()
    for _ in range(3):
        dm.assemble_K()
        assert np.array_equal(dm.get_csr()[2], a)            # bitwise
        assert np.array_equal(dm.internal_forces(), f1)
    dm.close()


def test_failure_statuses():
    # negative Jacobian (mech-solid.jl:150): swap two nodes of one element
    model = make_model("HEX8", 2, "le")
    model.conn[0, [0, 1]] = model.conn[0, [1, 0]]
    eqid, nu, _ = model.configure_dofs(clamp_bcs(model))
    dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
    with pytest.raises(L.AmaruStatus) as e:
        dm.assemble_K()
    assert e.value.code == L.FAIL_NEG_JACOBIAN
    dm.close()
    # no essential BCs: K11 singular -> the CG cannot converge / blows up, reported as a ReturnStatus failure
    model = make_model("HEX8", 2, "le")
    eqid, nu, setup = model.configure_dofs([("z==2", NodeBC(fz=1.0))])
    dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
    dm.assemble_K()
    U, F = model.get_bc_vals(eqid, setup)
    with pytest.raises(L.AmaruStatus) as e:
        dm.solve(U, F, cg_rtol=1e-12, cg_maxit=300)
    assert e.value.code in (L.FAIL_CG_NOCONV, L.FAIL_SINGULAR)
    dm.close()
    # unsupported shape id / material kind are refused (no CPU fallback)
    flat = model.flatten()
    flat["mat_kind"] = np.array([99], dtype=np.int32)
    with pytest.raises(L.AmaruStatus) as e:
        L.DeviceModel(flat, eqid, eqid.size, nu)
    assert e.value.code == L.ERR_UNSUPPORTED


# ------------------------------------------------------------------------------------------ whole solve! through the driver
def drive_both(model_fn, bcs_list, **kw):
    """Run the product solve() and the oracle's mech_stage_solver on the same stages; return both end states."""
    model = model_fn()
    ana = MechAnalysis(model)
    for bcs, nincs in bcs_list:
        addstage(ana, bcs, nincs=nincs)
    status = solve(ana, cg_rtol=1e-12, **kw)
    ref = model_fn()
    om = None
    Uacc = np.zeros((ref.nnodes, ref.ndim))
    for bcs, nincs in bcs_list:
        eqid, nu, setup = ref.configure_dofs(bcs)
        Uex, Fex = ref.get_bc_vals(eqid, setup)
        if om is None:
            om = O.OracleModel(ref.flatten(), eqid, eqid.size, nu)
        om.eqid, om.nu = np.ascontiguousarray(eqid), nu
        r = O.mech_stage_solver(om, Uex, Fex, nincs=nincs, **kw)
        Uacc += r["U"][eqid]
    return status, model, ana, r, Uacc, om


def test_solve_known_answers_quad4_hex8():
    # reference test/mech/elem/elastic-quad4.jl and elastic-hex8.jl (nodal load case), through the GPU path
    mesh = Mesh(Block([[0, 0], [1, 1]], nx=1, ny=1, cellshape="QUAD4", tag="solid"))
    model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=1.0, nu=0.25))], MechContext(stressmodel="planestrain"))
    ana = MechAnalysis(model)
    addstage(ana, [("x==0.", SurfaceBC(ux=0.)), ("y==0.", SurfaceBC(uy=0)), ("y==1.", SurfaceBC(ty=-1.))], nincs=1)
    assert solve(ana).success
    assert np.abs(model.U - np.array([[0, 0], [0.3125, 0], [0, -0.9375], [0.3125, -0.9375]])).max() < 1e-5
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=1, ny=1, nz=1, cellshape="HEX8", tag="solid"))
    base = [("x==0 && y==0 && z==0", NodeBC(ux=0, uy=0)), ("x==1 && y==0 && z==0", NodeBC(uy=0)),
            ("x==0 && y==1 && z==0", NodeBC(ux=0)), ("z==0", NodeBC(uz=0))]
    for extra, uz in [(("z==1", NodeBC(fz=1)), [0, 0, 0, 0, 4.0, 4.0, 4.0, 4.0]),
                      (("x==1", SurfaceBC(tx="3*z")), [0, 0, 0, 0, 1.51044, -2.4501, 1.4499, -2.31023]),
                      (("x>=0", BodyC(wz=-1)), [0, 0, 0, 0, -0.5, -0.5, -0.5, -0.5])]:
        model = FEModel(mesh, [("solid", MechSolid, LinearElastic, dict(E=1.0, nu=0.3))], MechContext())
        ana = MechAnalysis(model)
        addstage(ana, base + [extra], nouts=1)
        assert solve(ana).success
        assert np.abs(model.U[:, 2] - np.array(uz)).max() < 1e-5


def test_solve_config1_quad8_cantilever():
    """BASELINE config 1: QUAD8 20x10 plane-strain cantilever, E=200e6 nu=0.2, one load step."""
    def mk():
        mesh = Mesh(Block([[0, 0], [3, 0.4]], nx=20, ny=10, cellshape="QUAD8", tag="solids"))
        return FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=200e6, nu=0.2))], MechContext(stressmodel="planestrain"))
    bcs = [("x==0", NodeBC(ux=0, uy=0)), ("y==0.4", SurfaceBC(ty="-0.1*x"))]
    status, model, ana, r, Uacc, om = drive_both(mk, [(bcs, 1)])
    assert status.success and r["success"]
    assert model.nnodes == 661
    assert rel(model.U, Uacc) < 1e-8
    assert rel(model.state["sigma"], om.sig) < 1e-7


def test_solve_vm_cantilever_fixed_increments():
    """reference test/mech/mat/vm-3d.jl scenario with fixed increments (autoinc trajectories amplify solver noise):
    same increments, Newton iterations and end state as the oracle driver with the direct solver."""
    th = 0.05

    def mk():
        mesh = Mesh(Block([[0, 0, -0.05], [0.05, 1.0, 0.05]], nx=1, ny=30, nz=2, cellshape="HEX20"))
        return FEModel(mesh, [("bulks", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0))], MechContext())
    bcs = [("y==0", NodeBC(uy=0)), ("y==0 && z==0", NodeBC(uz=0)), (f"x=={th/2} && y==0 && z==0", NodeBC(ux=0)),
           (f"x=={th/2} && y==1 && z==0", NodeBC(uz=-0.02))]
    status, model, ana, r, Uacc, om = drive_both(mk, [(bcs, 10)], maxits=5, tol=1e-2, rtol=1e-2)
    assert status.success == r["success"]
    assert len(ana.stats) == r["its"]
    assert rel(model.U, Uacc) < 1e-8
    assert rel(model.state["sigma"], om.sig) < 1e-7
    assert (model.state["epa"] > 0).sum() == (om.epa > 0).sum() > 0


@pytest.mark.parametrize("scheme", ["ME", "BE", "Ralston"])
def test_solve_predictor_corrector_schemes(scheme):
    """scheme = :ME / :BE / :Ralston (mech-solver.jl:279-288, corrector :341-350: K = a1*K + a2*K2 through
    amaru_tangent_save / amaru_tangent_blend) on the vm-3d cantilever: same iteration history and end state as the oracle."""
    th = 0.05

    def mk():
        mesh = Mesh(Block([[0, 0, -0.05], [0.05, 1.0, 0.05]], nx=1, ny=12, nz=2, cellshape="HEX20"))
        return FEModel(mesh, [("bulks", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=1e5))], MechContext())
    bcs = [("y==0", NodeBC(uy=0)), ("y==0 && z==0", NodeBC(uz=0)), (f"x=={th/2} && y==0 && z==0", NodeBC(ux=0)),
           (f"x=={th/2} && y==1 && z==0", NodeBC(uz=-0.02))]
    status, model, ana, r, Uacc, om = drive_both(mk, [(bcs, 8)], maxits=5, tol=1e-2, rtol=1e-2, scheme=scheme)
    assert status.success == r["success"]
    assert len(ana.stats) == r["its"]
    assert rel(model.U, Uacc) < 1e-8
    assert rel(model.state["sigma"], om.sig) < 1e-7
    assert (om.epa > 0).sum() > 0 and (model.state["epa"] > 0).sum() == (om.epa > 0).sum()


def test_solve_dp_two_stages_autoinc():
    """reference test/mech/mat/dp.jl: load then unload, IP state carried across stages, autoinc; `.success`."""
    def mk():
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 0.5]], nx=2, ny=2, nz=2, tag="solids"))
        return FEModel(mesh, [("solids", MechSolid, DruckerPrager, dict(E=100., nu=0.25, alpha=0.05, kappa=0.1))], MechContext())
    b1 = [("z==0.0", NodeBC(ux=0, uy=0, uz=0)), ("z==0.5", NodeBC(uz=-0.033)),
          ("x==0 || x==1.0", NodeBC(ux=0, uy=0)), ("y==0 || y==1.0", NodeBC(ux=0, uy=0))]
    b2 = list(b1)
    b2[1] = ("z==0.5", NodeBC(uz=+0.008))
    status, model, ana, r, Uacc, om = drive_both(mk, [(b1, 10), (b2, 10)], tol=1e-2, autoinc=True)
    assert status.success and r["success"]
    assert rel(model.U, Uacc) < 1e-6
    assert rel(model.state["epa"], om.epa) < 1e-6


# ------------------------------------------------------------------------------------------ next tier: mass matrix, Kp, matvec
@pytest.mark.parametrize("shape,n", [("QUAD8", 3), ("HEX8", 3), ("HEX20", 2), ("TET10", 2)])
def test_mass_matrix_and_system_matrix(shape, n):
    """mount_M (dyn-solver.jl:72-103, elem_mass mech-solid.jl:169-205), Kp = a*K + b*M (:376-377) and its products."""
    model = make_model(shape, n, "le", jitter=0.1)
    om, dm, eqid, nu, setup = pair(model, clamp_bcs(model))
    rho = np.linspace(1.0, 3.0, model.nelems)
    st, M = om.mount_M(rho, filter_eps=False)
    st, K = om.mount_K(filter_eps=False)
    assert st == 0
    dm.assemble_K()
    dm.assemble_M(rho)
    rng = np.random.default_rng(7)
    x = rng.normal(size=eqid.size)
    a, b = 1.0 + 2 * 174.28e-6 / 1e-3, 4 / 1e-3 ** 2 + 2 * 4.2038 / 1e-3     # Newmark Kp coefficients (dyn-solid.jl:30)
    assert rel(dm.matvec(0.0, 1.0, x), M @ x) < 1e-12
    assert rel(dm.matvec(1.0, 0.0, x), K @ x) < 1e-12
    assert rel(dm.matvec(a, b, x), a * (K @ x) + b * (M @ x)) < 1e-12
    dm.set_system_matrix(a, b)
    rp, ci, val = dm.get_csr()
    Kp = (a * K + b * M).tocsr()
    Kp.sort_indices()
    assert np.array_equal(rp, Kp.indptr) and np.array_equal(ci, Kp.indices) and rel(val, Kp.data) < 1e-12
    Uex, Fex = model.get_bc_vals(eqid, setup)
    Fex = Fex + rng.normal(size=eqid.size) * (np.arange(eqid.size) < nu)
    U, F = Uex.copy(), Fex.copy()
    dm.solve(U, F, cg_rtol=1e-12)
    Uo, Fo = Uex.copy(), Fex.copy()
    ok, _ = O.solve_system(Kp.tocsc(), Uo, Fo, nu)
    assert ok and rel(U, Uo) < 1e-8 and rel(F[nu:], Fo[nu:]) < 1e-8
    dm.set_system_matrix(1.0, 0.0)
    assert rel(dm.get_csr()[2], K.tocsr().data) < 1e-12
    dm.close()


# ------------------------------------------------------------------------------------------ full-size properties (config 3 mesh)
def test_full_size_properties_hex20_1M():
    """At BASELINE config-3 size (HEX20 100^3, 12.27 M dofs, 2.1 G non-zeros) the oracle cannot run; check properties
    that do not depend on size: rigid-body translations are in the null space of the unconstrained K, K is symmetric
    (x'Ky = y'Kx), the product is linear, a uniform strain field gives the exact uniform stress at every IP and zero
    internal force at every interior node, and assembly is bitwise reproducible."""
    import sys
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    from bench import footing_model
    model, bcs = footing_model(100)
    eqid, nu, setup = model.configure_dofs(bcs)
    ndofs = eqid.size
    dm = L.DeviceModel(model.flatten(), eqid, ndofs, nu)
    assert dm.nnz == 2118711609 and model.nnodes == 4090601       # symbolic nnz / node count of SURVEY §8
    dm.assemble_K()
    E, nuu = 210e6, 0.3
    kscale = E * 0.01                                             # ~ diagonal magnitude E*h
    t = np.zeros(ndofs)
    t[eqid[:, 2]] = 1.0                                           # rigid translation in z
    assert np.abs(dm.matvec(1.0, 0.0, t)).max() < 1e-9 * kscale
    rng = np.random.default_rng(11)
    x, y = rng.normal(size=ndofs), rng.normal(size=ndofs)
    Kx, Ky = dm.matvec(1.0, 0.0, x), dm.matvec(1.0, 0.0, y)
    assert abs(x @ Ky - y @ Kx) < 1e-10 * abs(x @ Kx)
    assert rel(dm.matvec(1.0, 0.0, 2.0 * x - 3.0 * y), 2.0 * Kx - 3.0 * Ky) < 1e-12
    # patch test: u = (a x, b y, c z) -> uniform strain (elastic: strains well below yield)
    a, b, c = 1e-5, -2e-5, 1.5e-5
    U = np.zeros(ndofs)
    for d, g in enumerate((a, b, c)):
        U[eqid[:, d]] = g * model.coords[:, d]
    dF = dm.update_state(U)
    st = dm.get_state()
    lam, mu = E * nuu / ((1 + nuu) * (1 - 2 * nuu)), E / (2 * (1 + nuu))
    tr = a + b + c
    sig = np.array([lam * tr + 2 * mu * a, lam * tr + 2 * mu * b, lam * tr + 2 * mu * c, 0, 0, 0])
    assert np.abs(st["sigma"] - sig).max() < 1e-9 * np.abs(sig).max()
    X = model.coords
    interior = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    assert np.abs(dF[eqid[interior]]).max() < 1e-9 * np.abs(dF).max()
    v1 = dm.matvec(1.0, 0.0, x)
    dm.assemble_K()
    assert np.array_equal(dm.matvec(1.0, 0.0, x), Kx) and np.array_equal(v1, Kx)   # bitwise reproducible
    dm.close()


@pytest.mark.parametrize("shape,n", [("QUAD8", 6), ("HEX8", 5), ("HEX20", 4), ("TET10", 3)])
@pytest.mark.parametrize("precond", ["jacobi", "block-jacobi"])
def test_symmetric_storage_spmv_matches_full_storage(shape, n, precond, monkeypatch):
    """AMARU_SPMV_SYM=1: the CG loop multiplies with the upper blocks only (k_spmv_sym); default is the full-storage kernel.
    Same solution within the PCG tolerance, same iteration count within rounding, on the elastic and on a plastic tangent."""
    model = make_model(shape, n, "vm0" if shape != "QUAD8" else "vm", jitter=0.15, seed=3)
    bcs = clamp_bcs(model)
    eqid, nu, setup = model.configure_dofs(bcs)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    out = {}
    monkeypatch.setenv("AMARU_OPERATOR", "csr")   # the assembled-matrix CG operators (the default is the matrix-free one)
    for sym in ("1", "0"):
        monkeypatch.setenv("AMARU_SPMV_SYM", sym)
        dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
        try:
            assert ("sym" in dm.spmv_kernel) == (sym == "1")
            res = []
            for it in range(2):
                dm.assemble_K()
                U, F = Uex.copy(), Fex.copy()
                iters, rr = dm.solve(U, F, cg_rtol=1e-12, precond=L.PRECOND[precond])
                dm.update_state(3.0 * U)                                  # second pass: tangent on a plastic trial state
                res.append((U, F, iters))
            out[sym] = res
        finally:
            dm.close()
    for (U1, F1, i1), (U0, F0, i0) in zip(out["1"], out["0"]):
        assert rel(U1, U0) < 1e-9 and rel(F1[nu:], F0[nu:]) < 1e-9
        assert abs(i1 - i0) <= max(2, i0 // 50)


def test_edge_cases_single_element_all_prescribed_empty_sets():
    """Smallest / degenerate inputs: one element; every dof prescribed (nu = 0: solve_system! only forms the reactions
    K22*U2, solver.jl:32); a boundary condition whose filter selects nothing (empty load set)."""
    # (HEX8: a single HEX20 with its default 2x2x2 quadrature has spurious zero-energy modes, K11 would be singular)
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=1, ny=1, nz=1, cellshape="HEX8", tag="solids"))
    model = FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=100.0, nu=0.2))], MechContext())
    # one element, ordinary solve
    om, dm, eqid, nu, setup = pair(model, [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==1", SurfaceBC(tz=-1.0)),
                                          ("x>5", SurfaceBC(tx=3.0)), ("x>5", BodyC(wz=1.0))])      # last two select nothing
    U, F = model.get_bc_vals(eqid, setup, device=dm)
    assert abs(F.sum() + 1.0) < 1e-12
    K = check_K(om, dm)
    Uo, Fo = U.copy(), F.copy()
    ok, _ = O.solve_system(K.tocsc(), Uo, Fo, nu)
    dm.solve(U, F, cg_rtol=1e-13)
    assert ok and rel(U, Uo) < 1e-9 and rel(F[nu:], Fo[nu:]) < 1e-9
    dm.close()
    # nu = 0: all dofs prescribed
    eqid, nu, setup = model.configure_dofs([("x>=0", NodeBC(ux="0.01*x", uy=0, uz="-0.02*z"))])
    assert nu == 0
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
    K = check_K(om, dm)
    U, F = model.get_bc_vals(eqid, setup)
    Uo, Fo = U.copy(), F.copy()
    iters, rr = dm.solve(U, F, cg_rtol=1e-12)
    assert np.array_equal(U, Uo) and rel(F, K @ Uo) < 1e-12
    dF = dm.update_state(U)
    dFo, st = om.update_state(Uo)
    assert st == 0 and rel(dF, dFo) < 1e-12
    dm.close()


@pytest.mark.parametrize("shape,n", [("QUAD4", 5), ("QUAD8", 4)])
@pytest.mark.parametrize("op", ["csr", "ebe"])
def test_plane_stress_linear_elastic(shape, n, op):
    """stressmodel = :planestress (LinearElastic's own calcDe branch, linear-elastic.jl:99-108; AMARU_STRESS_PLANESTRESS).  The
    device runs on the equivalent plane-strain constants and clears σzz; the oracle uses the reference's plane-stress matrix
    literally: K bit pattern and values (1e-12), solve (1e-8), state update incl. σzz = 0 (1e-12), operator products."""
    mesh = Mesh(Block([[0, 0], [2, 1]], nx=2 * n, ny=n, cellshape=shape, tag="solids"))
    rng = np.random.default_rng(3)
    c = mesh.coords
    interior = (c[:, 0] > 1e-9) & (c[:, 0] < 2 - 1e-9) & (c[:, 1] > 1e-9) & (c[:, 1] < 1 - 1e-9)
    c[interior, :2] += rng.uniform(-0.15, 0.15, (interior.sum(), 2)) / (2 * n + 2)
    mesh.coords[...] = np.round(c, 8)
    model = FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=100.0, nu=0.3)),
                           ("y>=0.4", MechSolid, LinearElastic, dict(E=50.0, nu=0.2))], MechContext(stressmodel="planestress"),
                    thickness=0.7)
    om, dm, eqid, nu, setup = pair(model, clamp_bcs(model))
    dm.set_operator(op)
    K = check_K(om, dm)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    U, F = Uex.copy(), Fex.copy()
    dm.solve(U, F, cg_rtol=1e-12)
    Uo, Fo = Uex.copy(), Fex.copy()
    ok, _ = O.solve_system(K, Uo, Fo, nu)
    assert ok and rel(U, Uo) < 1e-8 and rel(F[nu:], Fo[nu:]) < 1e-8
    x = rng.uniform(-1, 1, eqid.size)
    y, _ = dm.operator_apply(x, masked=False)
    assert rel(y, K @ x) < 1e-12
    dF = dm.update_state(Uo)
    dFo, st = om.update_state(Uo)
    assert st == 0 and rel(dF, dFo) < 1e-12
    s = dm.get_state()
    assert rel(s["sigma"], om.sig) < 1e-12 and np.abs(s["sigma"][:, 2]).max() == 0.0
    assert rel(s["eps"], om.eps) < 1e-12
    dm.close()


@pytest.mark.parametrize("shape,n", [("QUAD4", 5), ("QUAD8", 4)])
@pytest.mark.parametrize("mat", ["le", "vm", "dp"])
def test_axisymmetric(shape, n, mat):
    """stressmodel = :axisymmetric (AMARU_STRESS_AXISYMMETRIC): hoop row N_a/r of B and th = 2*pi*r in mount_K, update_state!,
    elem_internal_forces and mount_M (mech-solid.jl:94-108,143,180,222,260) against the oracle (pinned to the Lamé cylinder in
    tests/test_oracle_golden.py): pattern bit-exact, K / f_int / IP state 1e-12, u 1e-8 over two chained plastic steps; the
    PCG of such handles runs on the block-CSR matrix (the matrix-free operator is refused)."""
    mesh = Mesh(Block([[1.0, 0], [3.0, 1.0]], nx=2 * n, ny=n, cellshape=shape, tag="solids"))
    rng = np.random.default_rng(7)
    c = mesh.coords
    interior = (c[:, 0] > 1 + 1e-9) & (c[:, 0] < 3 - 1e-9) & (c[:, 1] > 1e-9) & (c[:, 1] < 1 - 1e-9)
    c[interior, :2] += rng.uniform(-0.15, 0.15, (interior.sum(), 2)) / (2 * n + 2)
    mesh.coords[...] = np.round(c, 8)
    mty, par = MATS[mat]
    model = FEModel(mesh, [("solids", MechSolid, mty, par)], MechContext(stressmodel="axisymmetric"))
    bcs = [("y==0", NodeBC(uy=0)), ("x==1", NodeBC(ux=0.002 if mat != "dp" else 0.004)), ("y==1", SurfaceBC(ty="-0.1*x")),
           ("x>=0", BodyC(wy=-0.05))]
    om, dm, eqid, nu, setup = pair(model, bcs)
    assert "spmv" in dm.spmv_kernel
    with pytest.raises(L.AmaruStatus):
        dm.set_operator("ebe")
    Uex, Fex = model.get_bc_vals(eqid, setup, device=dm)              # axisymmetric loads: host path (th = 2*pi*r)
    import oracle.oracle_loads as OL
    fn, _ = [t for b, t in setup if isinstance(b, SurfaceBC)][0]
    Fo = np.zeros(eqid.size)
    ipx = OL.ip_coords(model.shape.facet_shape.id, model.coords, fn, 2)
    OL.apply(model.shape.facet_shape.id, model.coords, fn, eqid, 2, 1.0, 1, -0.1 * ipx[:, 0], Fo, axi=True)
    OL.apply(model.shape.id, model.coords, model.conn, eqid, 2, 1.0, 1, -0.05, Fo, axi=True)
    assert rel(Fex, Fo) < 1e-12
    for it in range(2):
        K = check_K(om, dm)
        U, F = Uex.copy(), Fex.copy()
        dm.solve(U, F, cg_rtol=1e-12)
        Uo, Fo2 = Uex.copy(), Fex.copy()
        ok, _ = O.solve_system(K, Uo, Fo2, nu)
        assert ok and rel(U, Uo) < 1e-8 and rel(F[nu:], Fo2[nu:]) < 1e-8
        dF = dm.update_state((1.0 + it) * Uo)
        dFo, st = om.update_state((1.0 + it) * Uo)
        assert st == 0 and rel(dF, dFo) < 1e-12
        s = dm.get_state()
        assert rel(s["sigma"], om.sig) < 1e-12 and rel(s["eps"], om.eps) < 1e-12 and rel(s["epa"], om.epa) < 1e-11
        assert rel(dm.internal_forces(), om.internal_forces()) < 1e-12
    if mat != "le":
        assert (om.dlam > 0).sum() > 0
    rho = rng.uniform(1.0, 3.0, model.nelems)
    dm.assemble_M(rho)
    _, M = om.mount_M(rho, filter_eps=False)
    x = rng.uniform(-1, 1, eqid.size)
    assert rel(dm.matvec(0.0, 1.0, x), M.tocsr() @ x) < 1e-12
    dm.close()
