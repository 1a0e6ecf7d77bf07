"""Multi-GPU handles of ONE process: ``amaru_create(..., ngpus > 1)`` / ``solve(ana, ngpus=N)``.

Needs >= 2 B200s on the box (skipped otherwise).  The same calls as the single-GPU tests, through the same C ABI with the
same global-length host vectors; compared with the single-GPU handle (entry point by entry point) and with the CPU oracle
(through the public ``solve`` driver on reduced BASELINE configs 3 and 4).
"""
import os

import numpy as np
import pytest

from amaru_jl_b200 import (Block, BodyC, DruckerPrager, FEModel, LinearElastic, MechAnalysis, MechContext, MechSolid, Mesh,
                           NodeBC, SurfaceBC, VonMises, addstage, solve)
from amaru_jl_b200 import lib as L
from amaru_jl_b200.dyn_solver import solve_dynamic
from amaru_jl_b200.model import DynamicAnalysis
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def ngpus_or_skip(want=2):
    n = L.device_count()
    if n < want:
        pytest.skip(f"needs at least {want} GPUs")
    return n


def rel(a, b):
    d = np.abs(b).max()
    return np.abs(a - b).max() / (d if d > 0 else 1.0)


def block_model(shape, n, mat):
    if shape == "QUAD8":
        mesh = Mesh(Block([[0, 0], [2, 1]], nx=4 * n, ny=2 * n, cellshape=shape, tag="s"))
        ctx = MechContext(stressmodel="planestrain")
    else:
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 2]], nx=n, ny=n, nz=2 * n, cellshape=shape, tag="s"))
        ctx = MechContext()
    mats = {"vm": (VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=1e6)),
            "dp": (DruckerPrager, dict(E=100.0, nu=0.25, alpha=0.05, kappa=0.1, H=0.0)),
            "le": (LinearElastic, dict(E=100.0, nu=0.2))}
    mty, par = mats[mat]
    model = FEModel(mesh, [("s", MechSolid, mty, par)], ctx)
    if model.ndim == 2:
        bcs = [("x==0", NodeBC(ux=0, uy=0)), ("x==2", NodeBC(uy=-0.05))]
    elif mat == "vm":
        bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==2 and x>=0.3 and x<=0.7", NodeBC(uz=-0.006))]
    else:
        bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==2", NodeBC(uz=-0.02, ux=0.004))]
    return model, bcs


@pytest.mark.parametrize("shape,n,mat,partitioner", [("HEX20", 6, "vm", "rcb"), ("HEX20", 6, "vm", "metis"),
                                                      ("TET10", 5, "dp", "metis"), ("HEX8", 8, "le", "rcb"),
                                                      ("QUAD8", 6, "vm", "rcb")])
def test_group_handle_matches_single_gpu(shape, n, mat, partitioner):
    """assemble_K -> solve -> update_state (twice: the second tangent is on the plastic trial state), IP state in / out,
    operator and matrix products, internal forces: ngpus = 2 (and 4 when the box has them) against one GPU."""
    have = ngpus_or_skip()
    model, bcs = block_model(shape, n, mat)
    eqid, nu, setup = model.configure_dofs(bcs)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    ndofs = int(eqid.size)
    flat = model.flatten()
    ref = L.DeviceModel(flat, eqid, ndofs, nu)
    for ng in [g for g in (2, 4, 8) if g <= have]:
        dm = L.DeviceModel(flat, eqid, ndofs, nu, ngpus=ng, partitioner=partitioner)
        assert dm.ngpus == ng
        ref.set_state(*[np.zeros_like(v) for v in ref.get_state().values()])
        ref.state_backup()
        rng = np.random.default_rng(5)
        for it in range(2):
            dm.assemble_K()
            ref.assemble_K()
            U, F = Uex.copy(), Fex.copy()
            U0, F0 = Uex.copy(), Fex.copy()
            iters, rr = dm.solve(U, F, cg_rtol=1e-12)
            it0, rr0 = ref.solve(U0, F0, cg_rtol=1e-12)
            assert abs(iters - it0) <= max(3, it0 // 20)
            assert rel(U, U0) < 1e-8 and rel(F[nu:], F0[nu:]) < 1e-7
            x = rng.normal(size=ndofs)
            x[nu:] = 0.0
            y, pq = dm.operator_apply(x, masked=True)
            y0, pq0 = ref.operator_apply(x, masked=True)
            assert rel(y, y0) < 1e-11 and abs(pq - pq0) < 1e-11 * abs(pq0)
            assert rel(dm.matvec(1.0, 0.0, x), ref.matvec(1.0, 0.0, x)) < 1e-12
            if it:
                dm.state_restore()
                ref.state_restore()
            dF = dm.update_state(U0)
            dF0 = ref.update_state(U0)
            assert rel(dF, dF0) < 1e-12
            s, s0 = dm.get_state(), ref.get_state()
            for k in ("sigma", "eps", "epa", "dlam"):
                assert rel(s[k], s0[k]) < 1e-12, k
            assert rel(dm.internal_forces(), ref.internal_forces()) < 1e-12
        if mat != "le":
            assert (s0["dlam"] > 0).sum() > 0
        # set_state / backup / restore round trip through the partition (owner's copy wins on the way out)
        st = {k: rng.normal(size=v.shape) for k, v in s0.items()}
        dm.set_state(**st)
        back = dm.get_state()
        for k in st:
            assert np.array_equal(back[k], st[k]), k
        dm.close()
    ref.close()


def test_group_mass_and_system_matrix():
    """mount_M, Kp = a*K + b*M and the products / solves of the Newmark algebra on a 2-GPU handle."""
    ngpus_or_skip()
    model, bcs = block_model("HEX8", 6, "le")
    eqid, nu, setup = model.configure_dofs(bcs)
    flat, ndofs = model.flatten(), int(eqid.size)
    rho = np.linspace(1.0, 3.0, model.nelems)
    ref = L.DeviceModel(flat, eqid, ndofs, nu)
    dm = L.DeviceModel(flat, eqid, ndofs, nu, ngpus=2)
    for d in (ref, dm):
        d.assemble_K()
        d.assemble_M(rho)
    x = np.random.default_rng(1).normal(size=ndofs)
    a, b = 1.0 + 2 * 174.28e-6 / 1e-3, 4 / 1e-3 ** 2 + 2 * 4.2038 / 1e-3
    for aa, bb in ((0.0, 1.0), (1.0, 0.0), (a, b)):
        assert rel(dm.matvec(aa, bb, x), ref.matvec(aa, bb, x)) < 1e-12
    for d in (ref, dm):
        d.set_system_matrix(a, b)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    U, F, U0, F0 = Uex.copy(), Fex.copy(), Uex.copy(), Fex.copy()
    dm.solve(U, F, cg_rtol=1e-12)
    ref.solve(U0, F0, cg_rtol=1e-12)
    assert rel(U, U0) < 1e-8 and rel(F[nu:], F0[nu:]) < 1e-7
    dm.close()
    ref.close()


def oracle_run(model, stages, **kw):
    om, Uacc, last = None, np.zeros((model.nnodes, model.ndim)), None
    for bcs, nincs in stages:
        eqid, nu, setup = model.configure_dofs(bcs)
        Uex, Fex = model.get_bc_vals(eqid, setup)
        if om is None:
            om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
        om.eqid, om.nu = np.ascontiguousarray(eqid), nu
        last = O.mech_stage_solver(om, Uex, Fex, nincs=nincs, **kw)
        Uacc += last["U"][eqid]
    return last, Uacc, om


@pytest.mark.parametrize("ng,partitioner", [(2, "rcb"), (4, "metis")])
def test_solve_ngpus_config3_reduced_vs_oracle(ng, partitioner):
    """solve(ana, ngpus=N) on config 3 at 8^3 HEX20 (von Mises footing, 10 increments) against the oracle driver."""
    ngpus_or_skip(ng)

    def mk():
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=8, ny=8, nz=8, cellshape="HEX20", tag="solids"))
        return FEModel(mesh, [("solids", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0.0))], MechContext())
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==1 and x>=0.375 and x<=0.625 and y>=0.375 and y<=0.625", NodeBC(uz=-0.004))]
    model = mk()
    ana = MechAnalysis(model)
    addstage(ana, bcs, nincs=10)
    status = solve(ana, cg_rtol=1e-12, maxits=5, ngpus=ng, partitioner=partitioner)
    r, Uacc, om = oracle_run(mk(), [(bcs, 10)], maxits=5)
    assert status.success == r["success"]
    assert len(ana.stats) == r["its"]
    assert rel(model.U, Uacc) < 1e-7
    assert rel(model.state["sigma"], om.sig) < 1e-6
    assert (om.epa > 0).sum() > 0 and np.array_equal(model.state["epa"] > 0, om.epa > 0)


@pytest.mark.parametrize("ng,partitioner", [(2, "metis"), (4, "rcb")])
def test_solve_ngpus_config4_reduced_vs_oracle(ng, partitioner):
    """solve(ana, ngpus=N) on config 4 at 4^3 x 6 TET10 (Drucker-Prager, gravity BodyC integrated by the device load sets of
    the multi-GPU handle) against the oracle driver."""
    ngpus_or_skip(ng)

    def mk():
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=4, ny=4, nz=4, cellshape="TET10", tag="solids"))
        return FEModel(mesh, [("solids", MechSolid, DruckerPrager, dict(E=100.0, nu=0.25, alpha=0.05, kappa=0.1))], MechContext())
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("x==0", NodeBC(ux=0)), ("y==0 || y==1", NodeBC(uy=0)), ("z>=0", BodyC(wz=-0.3))]
    model = mk()
    ana = MechAnalysis(model)
    addstage(ana, bcs, nincs=4)
    status = solve(ana, cg_rtol=1e-12, tol=1e-3, ngpus=ng, partitioner=partitioner)
    r, Uacc, om = oracle_run(mk(), [(bcs, 4)], tol=1e-3)
    assert status.success == r["success"]
    assert len(ana.stats) == r["its"]
    assert rel(model.U, Uacc) < 1e-7
    assert rel(model.state["sigma"], om.sig) < 1e-6


def test_solve_dynamic_ngpus_matches_single():
    """config 5's driver (Newmark, consistent mass) on a 2-GPU handle against the same run on one GPU."""
    ngpus_or_skip()

    def run(ng):
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 2]], nx=6, ny=6, nz=12, cellshape="HEX8", tag="solids"))
        model = FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=30e6, nu=0.2, rho=24.0))], MechContext())
        ana = DynamicAnalysis(model)
        addstage(ana, [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==2", SurfaceBC(tz="-100*t"))], tspan=0.01, nincs=5)
        st = solve_dynamic(ana, alpha=4.2038, beta=174.28e-6, cg_rtol=1e-12, ngpus=ng)
        assert st.success
        return model
    m1, m2 = run(1), run(2)
    assert rel(m2.U, m1.U) < 1e-8 and rel(m2.V, m1.V) < 1e-7 and rel(m2.A, m1.A) < 1e-6


def test_group_output_side_loads_and_recovery(tmp_path):
    """Device load sets and nodal patch recovery of a multi-GPU handle (integrated on its first GPU from the global arrays
    and the owners' IP state) against the single-GPU handle, through solve() with VTU output."""
    ngpus_or_skip()
    fields = {}
    for ng in (1, 2):
        mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=5, ny=5, nz=5, cellshape="HEX20", tag="solids"))
        # (H = 2e7: with H = 1e6 the three increments do not converge under the reference's rules — the CPU oracle agrees)
        model = FEModel(mesh, [("solids", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=2e7))], MechContext())
        ana = MechAnalysis(model, outdir=str(tmp_path / f"out{ng}"))
        addstage(ana, [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==1", SurfaceBC(tz="-3e5*x")), ("z>=0", BodyC(wz=-20.0))],
                 nincs=3, nouts=1)
        assert solve(ana, cg_rtol=1e-12, ngpus=ng).success
        fields[ng] = (model.U.copy(), {k: v.copy() for k, v in model.node_data.items()})
        assert os.path.exists(tmp_path / f"out{ng}" / f"{ana.outkey}-1.vtu")
    assert rel(fields[2][0], fields[1][0]) < 1e-8
    assert set(fields[2][1]) == set(fields[1][1])
    for k, v in fields[1][1].items():
        assert rel(fields[2][1][k], v) < 1e-7, k


def test_group_comm_failure_is_reported_not_hung():
    """a rank that leaves the collective sequence must surface as AMARU_ERR_COMM, not as a device hang.  One GPU
    skips a scalar all-reduce; the peers' bounded waits give up after AMARU_P2P_TIMEOUT_MS and raise the abort flag."""
    ngpus_or_skip()
    model, bcs = block_model("HEX8", 4, "le")
    eqid, nu, _ = model.configure_dofs(bcs)
    os.environ["AMARU_P2P_TIMEOUT_MS"] = "300"
    try:
        dm = L.DeviceModel(model.flatten(), eqid, int(eqid.size), nu, ngpus=2)
    finally:
        os.environ.pop("AMARU_P2P_TIMEOUT_MS")
    dm.comm_selftest(-1)                                  # everybody takes part: fine
    with pytest.raises(L.AmaruStatus) as ei:
        dm.comm_selftest(1)
    assert ei.value.code == L.ERR_COMM
    dm.close()


def test_group_rejects_bad_requests():
    have = ngpus_or_skip()
    model, bcs = block_model("HEX8", 2, "le")
    eqid, nu, _ = model.configure_dofs(bcs)
    with pytest.raises(L.AmaruStatus) as ei:
        L.DeviceModel(model.flatten(), eqid, int(eqid.size), nu, ngpus=have + 1)
    assert ei.value.code == L.ERR_ARG
    with pytest.raises(L.AmaruStatus):
        L.DeviceModel(model.flatten(), eqid, int(eqid.size), nu, ngpus=2, devices=[0, 0])
