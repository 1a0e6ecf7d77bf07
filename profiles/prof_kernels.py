"""Short driver for ncu captures: builds the HEX20 von Mises footing model at --size, then launches each hot kernel a
few times (assemble_K, update_state, SpMV+dot, fused CG update, p-update).  Usage (under gpurun):

  ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 2 -c 2 -o gpurun_out/spmv \
      python profiles/prof_kernels.py --size 100
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from amaru_jl_b200 import lib as L  # noqa: E402
from bench import footing_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=100)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--precond", default="block-jacobi")
args = ap.parse_args()

model, bcs = footing_model(args.size)
eqid, nu, setup = model.configure_dofs(bcs)
Uex, Fex = model.get_bc_vals(eqid, setup)
dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
pc = L.PRECOND[args.precond]
dm.state_backup()
dm.assemble_K()
dU = 0.1 * Uex
dm.update_state(dU)          # plastic trial state
dm.assemble_K()
try:                         # a few CG iterations so that p, q, r hold real data (the timed kernels run on them)
    dm.solve(dU.copy(), 0.1 * Fex, cg_rtol=1e-30, cg_maxit=5, precond=pc)
except L.AmaruStatus:
    pass
print("spmv kernel:", dm.spmv_kernel, " bytes/launch:", dm.spmv_bytes)
names = {1: "assemble_K (all colours)", 2: "update (internal forces mode)", 0: "spmv+dot", 3: "cg_update", 4: "cg_pupdate"}
for kind in (1, 2, 0, 3, 4):
    ms = dm.time_kernel(kind, reps=args.reps, precond=pc)
    print(f"{names[kind]:32s} {ms:9.4f} ms/launch")
dm.close()
