import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amaru_jl_b200 import lib as L
from bench import footing_model
model, bcs = footing_model(int(sys.argv[1]) if len(sys.argv) > 1 else 100)
eqid, nu, setup = model.configure_dofs(bcs)
dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
dm.assemble_K()
ms = dm.time_kernel(0, reps=1)
out = np.zeros(32, dtype=np.int64)
dm.lib.amaru_debug_read.argtypes = [C.c_void_p, C.c_void_p]
dm.lib.amaru_debug_read(dm.h, out.ctypes.data_as(C.c_void_p))
n = max(out[16], 1)
print("ms %.3f tiles/CTA %d" % (ms, n))
print("consumer warp0 cycles/tile: other %.0f fullr_wait %.0f gather_issue %.0f cpwait %.0f csync %.0f fullv_wait %.0f contract %.0f" % tuple(out[:7] / n))
print("producer cycles/tile: other %.0f emptyr_wait %.0f emptyv_wait %.0f" % tuple(out[8:11] / n))
