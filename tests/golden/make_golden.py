#!/usr/bin/env python
"""Generates the committed fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference (Amaru.jl) is Julia-only and Julia is not installed in this image, so fixtures cannot be produced by
running the reference here.  Two kinds of fixture are kept instead:

  reference_known_answers.json   every numeric known answer the reference's own tests hold for the mechanical path,
                                 copied as DATA with its file:line (these pin the oracle, tests/test_oracle_golden.py)
  hotpath_<shape>_<mat>.npz      input/output vectors of the hot path produced by the CPU oracle (oracle/) on small
                                 seeded meshes: CSR pattern, K values on the elastic state and on a plastic trial state,
                                 f_int and the IP state after update_state!.  They freeze the oracle (a later edit of the
                                 oracle that changes any value fails tests/test_golden_fixtures.py) and give the CUDA
                                 parity tests a second, committed, target.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

KNOWN = {
    "elastic-quad4": {"ref": "test/mech/elem/elastic-quad4.jl:32-46", "atol": 1e-5,
                      "dis": [[0.0, 0.0], [0.3125, 0.0], [0.0, -0.9375], [0.3125, -0.9375]]},
    "elastic-hex8-nodal": {"ref": "test/mech/elem/elastic-hex8.jl:55-62", "atol": 1e-5, "uz": [0, 0, 0, 0, 4.0, 4.0, 4.0, 4.0]},
    "elastic-hex8-triface": {"ref": "test/mech/elem/elastic-hex8.jl:64-69", "atol": 1e-5,
                             "uz": [0, 0, 0, 0, 1.51044, -2.4501, 1.4499, -2.31023]},
    "elastic-hex8-body": {"ref": "test/mech/elem/elastic-hex8.jl:71-75", "atol": 1e-5, "uz": [0, 0, 0, 0, -0.5, -0.5, -0.5, -0.5]},
    "elastic-elems": {"ref": "test/mech/elem/elastic-elems.jl:4,31,67", "top_node_uy_uz": [-0.012, -0.095], "atol": 1e-2},
    "vm-3d": {"ref": "test/mech/mat/vm-3d.jl:37", "fz": -30.0, "atol": 0.7},
    "structured-node-counts": {"ref": "test/mesh/structured.jl:24-64", "QUAD8": 341, "HEX8": 1331, "HEX20": 4961, "TET10": 9261},
}

CASES = [("QUAD8", 3, "le"), ("QUAD8", 3, "vm"), ("HEX8", 3, "dp"), ("HEX20", 2, "vm0"), ("HEX20", 2, "le"), ("TET10", 2, "dp"),
         ("TET10", 2, "vm")]


def case_arrays(shape, n, mat):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_parity import clamp_bcs, make_model
    from oracle import oracle as O
    model = make_model(shape, n, mat, jitter=0.2, seed=7)
    eqid, nu, setup = model.configure_dofs(clamp_bcs(model))
    Uex, Fex = model.get_bc_vals(eqid, setup)
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    st, K = om.mount_K(filter_eps=False)
    assert st == 0
    K = K.tocsr()
    K.sort_indices()
    U, F = Uex.copy(), Fex.copy()
    ok, _ = O.solve_system(K, U, F, nu)
    assert ok
    scale = 3.0 if mat != "le" else 1.0                       # push the plastic materials past yield
    dU = scale * U
    dF, st = om.update_state(dU)
    assert st == 0
    st, K2 = om.mount_K(filter_eps=False)                     # tangent on the trial state
    K2 = K2.tocsr()
    K2.sort_indices()
    return dict(coords=model.coords, conn=model.conn, eqid=eqid, nu=np.int64(nu), rowptr=K.indptr.astype(np.int64),
                colind=K.indices.astype(np.int32), K_elastic=K.data, U=U, reactions=F, dU=dU, f_int=dF, sigma=om.sig, eps=om.eps,
                epa=om.epa, dlam=om.dlam, K_trial=K2.data, plastic_ips=np.int64((om.dlam > 0).sum()))


def main():
    json.dump(KNOWN, open(os.path.join(HERE, "reference_known_answers.json"), "w"), indent=1)
    for shape, n, mat in CASES:
        a = case_arrays(shape, n, mat)
        fn = os.path.join(HERE, f"hotpath_{shape}_{mat}.npz")
        np.savez_compressed(fn, **a)
        print(fn, os.path.getsize(fn) // 1024, "KiB  plastic IPs:", int(a["plastic_ips"]))


if __name__ == "__main__":
    main()
