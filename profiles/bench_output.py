"""Timing of the next-tier kernels at BASELINE config-3 size (HEX20 100^3): device load integration (K11) and nodal patch
recovery (K12), with the host-side numpy / oracle paths on a bounded sample beside them.  Prints one JSON line.

  python profiles/bench_output.py [--size 100]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from amaru_jl_b200 import BodyC, NodeBC, SurfaceBC  # noqa: E402
from amaru_jl_b200 import lib as L  # noqa: E402
from amaru_jl_b200.output import boundary_nodes, update_output_data  # noqa: E402
from bench import footing_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=100)
ap.add_argument("--cpu-size", type=int, default=8)
args = ap.parse_args()


def timed(f, reps=1):
    t0 = time.perf_counter()
    for _ in range(reps):
        out = f()
    return (time.perf_counter() - t0) / reps, out


out = {"workload": f"HEX20 {args.size}^3 von Mises footing"}
model, bcs = footing_model(args.size)
bcs = bcs + [("z==1", SurfaceBC(tz="-10*x")), ("z>=0", BodyC(wz=-1.0))]
eqid, nu, setup = model.configure_dofs(bcs)
dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
t_first, (U, F) = timed(lambda: model.get_bc_vals(eqid, setup, device=dm))          # builds the load sets
t_dev, _ = timed(lambda: model.get_bc_vals(eqid, setup, device=dm), reps=3)          # per-step cost (cached sets)
out["loads"] = {"facets": int(setup[2][1][0].shape[0]), "cells": int(model.nelems), "device_first_call_s": t_first,
                "device_per_call_s": t_dev, "resultant_z": float(F[eqid[:, 2]].sum())}
# recovery
dm.assemble_K()
Us, Fs = 0.1 * U, 0.1 * F
dm.solve(Us, Fs, cg_rtol=1e-6)
dm.update_state(Us)
t_ab, ab = timed(lambda: boundary_nodes(model))
t_create, _ = timed(lambda: dm.recovery_create(ab))
t_rec, V = timed(lambda: dm.recover_nodal(model.nnodes), reps=3)
out["recovery"] = {"nodes": int(model.nnodes), "fields": dm.recovery_fields(), "boundary_nodes_s": t_ab, "setup_host_s": t_create,
                   "recover_nodal_incl_d2h_s": t_rec, "d2h_bytes": int(V.size * 8)}
dm.close()

# host-side comparison on a bounded sample: numpy load integration and the oracle's literal nodal_patch_recovery
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from oracle import oracle_recovery as OR  # noqa: E402
small, sb = footing_model(args.cpu_size)
sb = sb + [("z==1", SurfaceBC(tz="-10*x")), ("z>=0", BodyC(wz=-1.0))]
eq2, nu2, setup2 = small.configure_dofs(sb)
t_np, _ = timed(lambda: small.get_bc_vals(eq2, setup2))
rng = np.random.default_rng(0)
nip = small.nip_total
sig, eps, epa = rng.standard_normal((nip, 6)), rng.standard_normal((nip, 6)), np.abs(rng.standard_normal(nip))
kinds = np.array([small.materials[i].kind for i in small.elem_mat])
t_or, _ = timed(lambda: OR.nodal_patch_recovery(3, False, small.coords, small.conn, small.shape.id, small.ip_coords(), kinds, sig,
                                                 eps, epa, boundary_nodes(small).astype(bool)))
out["cpu_sample"] = {"elements": int(small.nelems), "numpy_loads_s": t_np, "oracle_recovery_s": t_or,
                     "oracle_recovery_nodes_per_s": small.nnodes / t_or}
out["recovery"]["nodes_per_s_device"] = model.nnodes / t_rec
print(json.dumps(out))
