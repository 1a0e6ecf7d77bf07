"""Committed golden vectors (tests/golden/, made by tests/golden/make_golden.py).

CPU: the oracle still reproduces every fixture (guards the oracle against drift) and the known-answer table carries the
reference's file:line for every entry.  GPU: the CUDA path, through the C ABI, against the same fixtures — pattern and dof
numbering bit-exact; K, f_int, IP state within 1e-12 relative; PCG displacements within 1e-8 relative."""
import glob
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from test_gpu_parity import clamp_bcs, make_model, rel  # noqa: E402

from amaru_jl_b200 import lib as L  # noqa: E402
from oracle import oracle as O  # noqa: E402

FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "hotpath_*.npz")))


def load(fn):
    shape, mat = os.path.basename(fn)[len("hotpath_"):-4].split("_")
    n = 3 if shape in ("QUAD8", "HEX8") else 2
    g = np.load(fn)
    model = make_model(shape, n, mat, jitter=0.2, seed=7)
    assert np.array_equal(model.coords, g["coords"]) and np.array_equal(model.conn, g["conn"])     # mesher + jitter frozen
    eqid, nu, setup = model.configure_dofs(clamp_bcs(model))
    assert np.array_equal(eqid, g["eqid"]) and nu == int(g["nu"])                                 # dof numbering bit-exact
    return model, eqid, nu, setup, g


def test_known_answer_table_cites_the_reference():
    known = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))
    assert len(known) >= 7 and len(FIXTURES) >= 7
    for k, v in known.items():
        assert v["ref"].startswith("test/") and ".jl:" in v["ref"], k
    assert known["elastic-quad4"]["dis"][3] == [0.3125, -0.9375] and known["vm-3d"]["fz"] == -30.0


@pytest.mark.parametrize("fn", FIXTURES, ids=[os.path.basename(f)[8:-4] for f in FIXTURES])
def test_oracle_reproduces_fixture(fn):
    model, eqid, nu, setup, g = load(fn)
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    st, K = om.mount_K(filter_eps=False)
    K = K.tocsr()
    K.sort_indices()
    assert np.array_equal(K.indptr, g["rowptr"]) and np.array_equal(K.indices, g["colind"])
    assert rel(K.data, g["K_elastic"]) < 1e-14
    dF, st = om.update_state(g["dU"])
    assert st == 0 and rel(dF, g["f_int"]) < 1e-13 and rel(om.sig, g["sigma"]) < 1e-13
    assert int((om.dlam > 0).sum()) == int(g["plastic_ips"])
    st, K2 = om.mount_K(filter_eps=False)
    K2 = K2.tocsr()
    K2.sort_indices()
    assert rel(K2.data, g["K_trial"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("fn", FIXTURES, ids=[os.path.basename(f)[8:-4] for f in FIXTURES])
def test_cuda_path_reproduces_fixture(fn):
    model, eqid, nu, setup, g = load(fn)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
    try:
        dm.assemble_K()
        rp, ci, val = dm.get_csr()
        assert np.array_equal(rp, g["rowptr"]) and np.array_equal(ci, g["colind"])               # pattern bit-exact
        assert rel(val, g["K_elastic"]) < 1e-12
        U, F = Uex.copy(), Fex.copy()
        dm.solve(U, F, cg_rtol=1e-13)
        assert rel(U, g["U"]) < 1e-8 and rel(F[nu:], g["reactions"][nu:]) < 1e-8
        dF = dm.update_state(g["dU"])
        s = dm.get_state()
        assert rel(dF, g["f_int"]) < 1e-12 and rel(s["sigma"], g["sigma"]) < 1e-12 and rel(s["eps"], g["eps"]) < 1e-12
        assert np.abs(s["epa"] - g["epa"]).max() <= 1e-12 * max(np.abs(g["epa"]).max(), 1e-300)
        assert int((s["dlam"] > 0).sum()) == int(g["plastic_ips"])
        dm.assemble_K()
        assert rel(dm.get_csr()[2], g["K_trial"]) < 1e-12
    finally:
        dm.close()
