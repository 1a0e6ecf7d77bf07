"""ncu driver for the matrix-free operator: footing model at --size, plastic trial state, a few operator applications.
  ncu --set full --clock-control none --import-source on -k regex:k_ebe_apply -s 8 -c 1 -o gpurun_out/ebe python profiles/prof_ebe.py"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amaru_jl_b200 import lib as L  # noqa: E402
from bench import footing_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
model, bcs = footing_model(args.size)
eqid, nu, setup = model.configure_dofs(bcs)
Uex, Fex = model.get_bc_vals(eqid, setup)
dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
dm.state_backup()
dm.assemble_K()
dm.update_state(0.1 * Uex)
dm.assemble_K()
dm.set_operator("ebe")
print("operator:", dm.spmv_kernel, "bytes/apply:", dm.spmv_bytes)
print("apply+dot ms:", dm.time_kernel(0, reps=args.reps))
dm.close()
