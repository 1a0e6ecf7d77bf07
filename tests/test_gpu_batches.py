"""Several element batches in one handle (north_star: element batches sorted by cell type).

1. The same HEX20 / QUAD8 model passed as ONE batch and as TWO batches of the same shape must give the same matrix, state
   update, IP-state order at the ABI, solve and nodal recovery.
2. A model mixing cell shapes (HEX8 + TET10 bodies sharing their interface corner nodes; QUAD4 + QUAD8 in 2D) against the
   CPU oracle, which is composed per batch (K = K_a + K_b on the shared dof numbering).
"""
import numpy as np
import pytest
import scipy.sparse as sp

from amaru_jl_b200 import Block, FEModel, LinearElastic, MechContext, MechSolid, Mesh, NodeBC, VonMises
from amaru_jl_b200 import lib as L
from amaru_jl_b200.output import boundary_nodes
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["ebe-patch", "ebe-colour", "ebe-mixed", "csr"])
def cg_operator(request, monkeypatch):
    """Multi-batch handles with every CG operator: the patch form of the matrix-free operator (forced on these small meshes),
    its colour-ordered form, and the block-CSR SpMV (what handles this small start with)."""
    monkeypatch.setenv("AMARU_OPERATOR", "csr" if request.param == "csr" else "ebe")
    if request.param == "ebe-patch":
        monkeypatch.setenv("AMARU_EBE_PATCH_MINFILL", "0")
        monkeypatch.setenv("AMARU_EBE_PATCH_MINPATCH", "0")
    elif request.param == "ebe-colour":
        monkeypatch.setenv("AMARU_EBE_PATCH", "0")
    elif request.param == "ebe-mixed":               # batches with >= 2 patches take the patch form, the others the colour form
        monkeypatch.setenv("AMARU_EBE_PATCH_MINFILL", "0")   # (HEX8 + TET10 model: one batch of each form in one handle)
        monkeypatch.setenv("AMARU_EBE_PATCH_MINPATCH", "2")
    return request.param


def rel(a, b):
    d = np.abs(b).max()
    return np.abs(a - b).max() / (d if d > 0 else 1.0)


@pytest.mark.parametrize("shape", ["HEX20", "QUAD8"])
def test_two_batches_of_one_shape_equal_one_batch(shape):
    if shape == "HEX20":
        mesh = Mesh(Block([[0, 0, 0], [1, 1.5, 2]], nx=3, ny=3, nz=4, cellshape=shape, tag="s"))
        ctx = MechContext()
        bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==2", NodeBC(uz=-0.02))]
    else:
        mesh = Mesh(Block([[0, 0], [2, 1]], nx=6, ny=4, cellshape=shape, tag="s"))
        ctx = MechContext(stressmodel="planestrain")
        bcs = [("x==0", NodeBC(ux=0, uy=0)), ("x==2", NodeBC(uy=-0.01))]
    model = FEModel(mesh, [("s", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=1e6))], ctx)
    eqid, nu, setup = model.configure_dofs(bcs)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    flat1 = model.flatten()
    k = model.nelems // 3
    flat2 = dict(flat1)
    flat2["batch_shape"] = np.array([model.shape.id, model.shape.id], dtype=np.int32)
    flat2["batch_nelem"] = np.array([k, model.nelems - k], dtype=np.int64)
    out = []
    for flat in (flat1, flat2):
        dm = L.DeviceModel(flat, eqid, eqid.size, nu)
        try:
            dm.assemble_K()
            rp, ci, val = dm.get_csr()
            U, F = Uex.copy(), Fex.copy()
            dm.solve(U, F, cg_rtol=1e-13)
            dF = dm.update_state(3.0 * U)
            st = dm.get_state()
            dm.assemble_K()
            val2 = dm.get_csr()[2]
            dm.recovery_create(boundary_nodes(model))
            V = np.array(dm.recover_nodal(model.nnodes))
            names = dm.recovery_fields()
            fin = dm.internal_forces()
            out.append(dict(rp=rp, ci=ci, val=val, U=U, F=F, dF=dF, st=st, val2=val2, V=V, names=names, fin=fin))
        finally:
            dm.close()
    a, b = out
    assert np.array_equal(a["rp"], b["rp"]) and np.array_equal(a["ci"], b["ci"]) and a["names"] == b["names"]
    assert rel(b["val"], a["val"]) < 1e-13 and rel(b["val2"], a["val2"]) < 1e-12
    assert rel(b["U"], a["U"]) < 1e-9 and rel(b["F"], a["F"]) < 1e-9
    assert rel(b["dF"], a["dF"]) < 1e-11 and rel(b["fin"], a["fin"]) < 1e-11
    for key in ("sigma", "eps", "epa", "dlam"):                       # IP order at the ABI: element-major, batch after batch
        assert rel(b["st"][key], a["st"][key]) < 1e-10, key
    assert (a["st"]["dlam"] > 0).sum() > 0
    assert rel(b["V"], a["V"]) < 1e-9


def merged(meshes):
    """Concatenate meshes, merging coincident nodes (8-digit keys like the reference's point dictionary)."""
    key = {}
    coords, conns = [], []
    for m in meshes:
        gid = np.empty(m.nnodes, dtype=np.int32)
        for i, p in enumerate(np.round(m.coords, 8)):
            t = tuple(p)
            if t not in key:
                key[t] = len(coords)
                coords.append(m.coords[i])
            gid[i] = key[t]
        conns.append(gid[m.conn])
    return np.array(coords), conns


@pytest.mark.parametrize("dim", [3, 2])
def test_mixed_cell_shapes_vs_oracle(dim):
    if dim == 3:
        ma = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=3, ny=3, nz=3, cellshape="HEX8"))
        mb = Mesh(Block([[0, 0, 1], [1, 1, 2]], nx=3, ny=3, nz=3, cellshape="TET10"))
        stress, th = 0, 1.0
        fixed = lambda X: np.abs(X[:, 2]) < 1e-9
        top = lambda X: np.abs(X[:, 2] - 2) < 1e-9
    else:
        ma = Mesh(Block([[0, 0], [1, 1]], nx=4, ny=4, cellshape="QUAD4"))
        mb = Mesh(Block([[1, 0], [2, 1]], nx=4, ny=4, cellshape="QUAD8"))
        stress, th = 1, 0.5
        fixed = lambda X: np.abs(X[:, 0]) < 1e-9
        top = lambda X: np.abs(X[:, 0] - 2) < 1e-9
    coords, (ca, cb) = merged([ma, mb])
    nn, nd = coords.shape[0], dim
    assert nn < ma.nnodes + mb.nnodes                                   # the interface corner nodes are shared
    presc = np.zeros((nn, nd), dtype=bool)
    presc[fixed(coords)] = True
    presc[top(coords), nd - 1] = True
    eqid, nu = L.configure_dofs(presc)
    ndofs = eqid.size
    mat_kind = np.array([2, 1], dtype=np.int32)
    mat_par = np.array([[210e6, 0.3, 240e3, 1e6, 0, 0, 0, 0], [100e6, 0.2, 0, 0, 0, 0, 0, 0]])
    common = dict(ndim=nd, stressmodel=stress, thickness=th, coords=coords, mat_kind=mat_kind, mat_params=mat_par)
    fa = dict(common, batch_shape=np.array([ma.shape.id], dtype=np.int32), batch_nelem=np.array([ca.shape[0]], dtype=np.int64),
              conn=ca, elem_mat=np.zeros(ca.shape[0], dtype=np.int32))
    fb = dict(common, batch_shape=np.array([mb.shape.id], dtype=np.int32), batch_nelem=np.array([cb.shape[0]], dtype=np.int64),
              conn=cb, elem_mat=np.ones(cb.shape[0], dtype=np.int32))
    fab = dict(common, batch_shape=np.array([ma.shape.id, mb.shape.id], dtype=np.int32),
               batch_nelem=np.array([ca.shape[0], cb.shape[0]], dtype=np.int64),
               conn=np.concatenate((ca.reshape(-1), cb.reshape(-1))).astype(np.int32),
               elem_mat=np.concatenate((np.zeros(ca.shape[0]), np.ones(cb.shape[0]))).astype(np.int32))
    oa, ob = O.OracleModel(fa, eqid, ndofs, nu), O.OracleModel(fb, eqid, ndofs, nu)
    dm = L.DeviceModel(fab, eqid, ndofs, nu)
    try:
        Uex = np.zeros(ndofs)
        Uex[eqid[top(coords), nd - 1]] = -0.004
        for it in range(2):
            dm.assemble_K()
            rp, ci, val = dm.get_csr()
            Ka, Kb = oa.mount_K(filter_eps=False)[1], ob.mount_K(filter_eps=False)[1]
            pa, pb = oa.symbolic_csr(), ob.symbolic_csr()
            P = (sp.csr_matrix((np.ones(pa[1].size), pa[1], pa[0]), shape=(ndofs, ndofs)) +
                 sp.csr_matrix((np.ones(pb[1].size), pb[1], pb[0]), shape=(ndofs, ndofs))).tocsr()
            P.sort_indices()
            assert np.array_equal(rp, P.indptr) and np.array_equal(ci, P.indices)             # union pattern, bit-exact
            K = (Ka + Kb).tocsr()
            Kd = sp.csr_matrix((val, ci, rp), shape=(ndofs, ndofs))
            assert abs(Kd - K).max() < 1e-12 * abs(K).max()
            U, F = Uex.copy(), np.zeros(ndofs)
            dm.solve(U, F, cg_rtol=1e-13)
            Uo, Fo = Uex.copy(), np.zeros(ndofs)
            ok, _ = O.solve_system(K.tocsc(), Uo, Fo, nu)
            assert ok and rel(U, Uo) < 1e-8 and rel(F[nu:], Fo[nu:]) < 1e-8
            dF = dm.update_state(Uo)
            da, sa = oa.update_state(Uo)
            db, sb = ob.update_state(Uo)
            assert sa == 0 and sb == 0 and rel(dF, da + db) < 1e-12
            st = dm.get_state()
            assert rel(st["sigma"], np.vstack((oa.sig, ob.sig))) < 1e-12 and rel(st["eps"], np.vstack((oa.eps, ob.eps))) < 1e-12
            assert np.array_equal(st["dlam"] > 0, np.concatenate((oa.dlam, ob.dlam)) > 0)
        assert (oa.dlam > 0).sum() > 0                                   # the von Mises body yields on the second pass
    finally:
        dm.close()
