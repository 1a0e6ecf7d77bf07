import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from amaru_jl_b200 import lib as L
from bench import footing_model
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
model, bcs = footing_model(n)
eqid, nu, setup = model.configure_dofs(bcs)
Uex, Fex = model.get_bc_vals(eqid, setup)
dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
dm.state_backup(); dm.assemble_K()
dU = 0.1 * Uex
dm.update_state(dU); dm.assemble_K()
for op in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("csr", "ebe")):
    dm.set_operator(op)
    print(op, dm.spmv_kernel, "bytes", dm.spmv_bytes)
    for kind, nm in ((0, "operator+dot"), (3, "cg_update"), (4, "cg_pupdate"), (1, "assemble_K"), (2, "update")):
        print(f"   {nm:14s} {dm.time_kernel(kind, reps=10):9.4f} ms")
    dm.set_device_vectors(dU, 0.1 * Fex)
    for r in range(2):
        print("   newton", dm.newton_iteration_device(1e-10, 200000, 1))
