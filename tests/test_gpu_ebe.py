"""The matrix-free (element-by-element) CG operator (csrc/ebe.cu) against the assembled tangent.  Needs a B200: -m gpu.

The operator must be THE SAME linear map as the matrix mount_K builds (reference src/mech/mech-solver.jl:78-110 with
elem_stiffness, src/mech/elem/mech-solid.jl:124-166): products are compared with the CPU oracle's K (and a*K + b*M) at
1e-12 relative, the fused x.Ax with the explicit dot, and the PCG run on either operator must return the same
displacements (1e-8) in the same number of iterations (+-1: the two products differ in the last bits)."""
import numpy as np
import pytest

from amaru_jl_b200 import lib as L
from oracle import oracle as O
from test_gpu_parity import MATS, SHAPES, clamp_bcs, make_model, pair, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["patch", "colour"])
def operator_form(request, monkeypatch):
    """Every test runs on both forms of the matrix-free operator: k_ebe_patch (x / y of a patch in shared memory; forced here
    even where the small test meshes fill the patch slots poorly) and the colour-ordered k_ebe_mma."""
    if request.param == "patch":
        monkeypatch.setenv("AMARU_EBE_PATCH_MINFILL", "0")
        monkeypatch.setenv("AMARU_EBE_PATCH_MINPATCH", "0")
    else:
        monkeypatch.setenv("AMARU_EBE_PATCH", "0")
    return request.param


def plastic_state(model, om, dm, eqid, mat, seed=3):
    rng = np.random.default_rng(seed)
    scale = {"le": 1e-3, "vm": 4e-3, "vm0": 4e-3, "dp": 5e-3}[mat]
    for step in range(2):
        dU = rng.uniform(-1, 1, eqid.size) * scale * (step + 1)
        _, st = om.update_state(dU)
        assert st == 0
        dm.update_state(dU)
    if mat != "le":
        assert (om.dlam > 0).sum() > 0


@pytest.mark.parametrize("shape,n", SHAPES)
@pytest.mark.parametrize("mat", ["le", "vm", "vm0", "dp"])
def test_ebe_operator_equals_assembled_tangent(shape, n, mat):
    model = make_model(shape, n, mat, jitter=0.15, seed=1, mixed=(mat == "vm"))
    om, dm, eqid, nu, _ = pair(model, clamp_bcs(model))
    plastic_state(model, om, dm, eqid, mat)
    dm.assemble_K()
    st, K = om.mount_K(filter_eps=False)
    assert st == 0
    K = K.tocsr()
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, eqid.size)
    for op in ("ebe", "csr"):
        dm.set_operator(op)
        y, _ = dm.operator_apply(x, masked=False)
        assert rel(y, K @ x) < 1e-12, op
        xm = x.copy()
        xm[nu:] = 0.0                                            # p of the CG loop vanishes on the prescribed dofs
        ym, pq = dm.operator_apply(xm, masked=True)
        ref = K @ xm
        ref[nu:] = 0.0
        assert rel(ym, ref) < 1e-12, op
        assert abs(pq - xm @ ref) <= 1e-12 * abs(xm @ ref), op
    dm.close()


@pytest.mark.parametrize("shape,n", [("HEX8", 3), ("HEX20", 2), ("QUAD8", 3), ("TET10", 2)])
def test_ebe_operator_with_mass_term(shape, n):
    """Newmark system matrix a*K + b*M (dyn-solver.jl:376-377) through the matrix-free operator."""
    model = make_model(shape, n, "le", jitter=0.1, seed=4)
    om, dm, eqid, nu, _ = pair(model, clamp_bcs(model))
    rho = np.random.default_rng(6).uniform(1.0, 3.0, model.nelems)
    dm.assemble_K()
    dm.assemble_M(rho)
    a, b = 1.0 + 2 * 174.28e-6 / 1e-3, 4 / 1e-3 ** 2 + 2 * 4.2038 / 1e-3
    dm.set_system_matrix(a, b)
    _, K = om.mount_K(filter_eps=False)
    _, M = om.mount_M(rho, filter_eps=False)
    A = (a * K + b * M).tocsr()
    x = np.random.default_rng(7).uniform(-1, 1, eqid.size)
    for op in ("ebe", "csr"):
        dm.set_operator(op)
        y, _ = dm.operator_apply(x, masked=False)
        assert rel(y, A @ x) < 1e-12, op
    dm.close()


@pytest.mark.parametrize("shape,n,mat", [("HEX20", 4, "vm0"), ("TET10", 3, "dp"), ("HEX8", 5, "le"), ("QUAD8", 6, "vm")])
@pytest.mark.parametrize("precond", ["jacobi", "block-jacobi"])
def test_solve_is_operator_independent(shape, n, mat, precond):
    model = make_model(shape, n, mat, jitter=0.1, seed=2)
    bcs = clamp_bcs(model)
    om, dm, eqid, nu, setup = pair(model, bcs)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    plastic_state(model, om, dm, eqid, mat)
    dm.assemble_K()
    out = {}
    for op in ("ebe", "csr"):
        dm.set_operator(op)
        U, F = Uex.copy(), Fex.copy()
        it, rr = dm.solve(U, F, cg_rtol=1e-12, precond=L.PRECOND[precond])
        out[op] = (U, F, it)
    assert rel(out["ebe"][0], out["csr"][0]) < 1e-8
    assert rel(out["ebe"][1][nu:], out["csr"][1][nu:]) < 1e-8
    assert abs(out["ebe"][2] - out["csr"][2]) <= max(2, out["csr"][2] // 50)
    # and against the direct solve of the oracle
    st, K = om.mount_K()
    U, F = Uex.copy(), Fex.copy()
    ok, _ = O.solve_system(K, U, F, nu)
    assert ok and rel(out["ebe"][0], U) < 1e-8
    dm.close()


def test_operator_form_is_the_requested_one(operator_form):
    model = make_model("HEX20", 2, "le")
    om, dm, eqid, nu, _ = pair(model, clamp_bcs(model))
    dm.set_operator("ebe")                               # (handles of tiny models start with the CSR operator)
    assert dm.spmv_kernel.startswith("k_ebe_patch" if operator_form == "patch" else "k_ebe_mma")
    dm.close()


@pytest.mark.parametrize("shape,n", [("HEX20", 12), ("HEX20", 9), ("TET10", 8), ("HEX8", 13), ("QUAD8", 40)])
def test_many_patches_match_csr_and_repeat_bitwise(shape, n):
    """Enough patches (27 ... 100) that several warps and CTAs work concurrently and wait on each other's rows: the product
    equals the block-CSR product of the assembled tangent (plastic state) and repeats bit for bit."""
    model = make_model(shape, n, "vm0" if shape != "QUAD8" else "vm", jitter=0.1, seed=5)
    bcs = clamp_bcs(model)
    eqid, nu, setup = model.configure_dofs(bcs)
    dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
    rng = np.random.default_rng(8)
    dm.state_backup()
    dm.update_state(rng.uniform(-1, 1, eqid.size) * 6e-3)
    dm.assemble_K()
    x = rng.uniform(-1, 1, eqid.size)
    xm = x.copy()
    xm[nu:] = 0.0
    dm.set_operator("csr")
    yc, _ = dm.operator_apply(x, masked=False)
    ycm, pqc = dm.operator_apply(xm, masked=True)
    dm.set_operator("ebe")
    y0, _ = dm.operator_apply(x, masked=False)
    ym0, pq0 = dm.operator_apply(xm, masked=True)
    assert rel(y0, yc) < 1e-12 and rel(ym0, ycm) < 1e-12 and abs(pq0 - pqc) <= 1e-12 * abs(pqc)
    for _ in range(5):
        y, _ = dm.operator_apply(x, masked=False)
        ym, pq = dm.operator_apply(xm, masked=True)
        assert np.array_equal(y, y0) and np.array_equal(ym, ym0) and pq == pq0
    dm.close()


def test_ebe_is_deterministic():
    model = make_model("HEX20", 3, "vm", jitter=0.1)
    om, dm, eqid, nu, _ = pair(model, clamp_bcs(model))
    plastic_state(model, om, dm, eqid, "vm")
    dm.assemble_K()
    dm.set_operator("ebe")
    x = np.random.default_rng(1).uniform(-1, 1, eqid.size)
    x[nu:] = 0.0
    y0, p0 = dm.operator_apply(x, masked=True)
    for _ in range(3):
        y, p = dm.operator_apply(x, masked=True)
        assert np.array_equal(y, y0) and p == p0
    dm.close()
