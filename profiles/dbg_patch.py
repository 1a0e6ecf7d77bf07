"""Debug driver: one small model through the patch-form operator with a watchdog that dumps the Python stack."""
import faulthandler
import os
import sys

faulthandler.dump_traceback_later(280, exit=True)
os.environ.setdefault("AMARU_EBE_PATCH_MINFILL", "0")
os.environ.setdefault("AMARU_EBE_PATCH_MINPATCH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_gpu_parity import clamp_bcs, make_model, rel
from amaru_jl_b200 import lib as L

shape, n = sys.argv[1], int(sys.argv[2])
model = make_model(shape, n, "le", jitter=0.15, seed=1)
eqid, nu, setup = model.configure_dofs(clamp_bcs(model))
print("create", flush=True)
dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
print("kernel", dm.spmv_kernel, flush=True)
dm.assemble_K()
print("assembled", flush=True)
x = np.random.default_rng(5).uniform(-1, 1, eqid.size)
dm.set_operator("csr")
yc, _ = dm.operator_apply(x, masked=False)
dm.set_operator("ebe")
print("apply", flush=True)
y, _ = dm.operator_apply(x, masked=False)
print("rel", rel(y, yc), flush=True)
dm.close()
