"""One device-timed Newton iteration (assemble_K -> PCG -> state restore -> update_state) of every BASELINE configuration
that fits one GPU, at FULL size, through the same measurement hook bench.py uses.  Prints one JSON line per config.

  python profiles/bench_configs.py            (run under gpurun; ~1 min)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from amaru_jl_b200 import (Block, BodyC, DruckerPrager, FEModel, LinearElastic, MechContext, MechSolid, Mesh, NodeBC,  # noqa: E402
                           SurfaceBC, VonMises)
from amaru_jl_b200 import lib as L  # noqa: E402

PEAK = 6456.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def configs():
    m = Mesh(Block([[0, 0], [3, 0.4]], nx=20, ny=10, cellshape="QUAD8", tag="s"))
    yield ("config 1: QUAD8 20x10 cantilever, plane strain, linear elastic",
           FEModel(m, [("s", MechSolid, LinearElastic, dict(E=200e6, nu=0.2))], MechContext(stressmodel="planestrain")),
           [("x==0", NodeBC(ux=0, uy=0)), ("y==0.4", SurfaceBC(ty="-0.1*x"))], 1.0)
    m = Mesh(Block([[0, 0, 0], [2, 1, 0.8]], nx=100, ny=50, nz=40, cellshape="HEX8", tag="s"))
    yield ("config 2: HEX8 100x50x40 linear elastic block",
           FEModel(m, [("s", MechSolid, LinearElastic, dict(E=100.0, nu=0.2))], MechContext()),
           [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==0.8", NodeBC(fz=-0.01))], 1.0)
    m = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=100, ny=100, nz=100, cellshape="HEX20", tag="s"))
    yield ("config 3: HEX20 100^3 von Mises footing (first of 10 increments)",
           FEModel(m, [("s", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0.0))], MechContext()),
           [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==1 and x>=0.4 and x<=0.6 and y>=0.4 and y<=0.6", NodeBC(uz=-0.01))], 0.1)
    m = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=94, ny=94, nz=94, cellshape="TET10", tag="s"))
    yield ("config 4: TET10 94^3x6 Drucker-Prager block under gravity (on ONE GPU)",
           FEModel(m, [("s", MechSolid, DruckerPrager, dict(E=100.0, nu=0.25, alpha=0.05, kappa=0.1))], MechContext()),
           [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("x==0 || x==1", NodeBC(ux=0)), ("y==0 || y==1", NodeBC(uy=0)),
            ("z>=0", BodyC(wz=-0.01))], 1.0)
    m = Mesh(Block([[0, 0, 0], [2, 1, 1]], nx=200, ny=100, nz=100, cellshape="HEX8", tag="s"))
    yield ("config 5 (static tangent only): HEX8 200x100x100",
           FEModel(m, [("s", MechSolid, LinearElastic, dict(E=30e6, nu=0.2, rho=24.0))], MechContext()),
           [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==1 && x>=0.9 && x<=1.1", NodeBC(fz=-10.0))], 1.0)


for name, model, bcs, frac in configs():
    t0 = time.perf_counter()
    eqid, nu, setup = model.configure_dofs(bcs)
    dm = L.DeviceModel(model.flatten(), eqid, eqid.size, nu)
    Uex, Fex = model.get_bc_vals(eqid, setup, device=dm)
    setup_s = time.perf_counter() - t0
    dm.state_backup()
    dm.set_device_vectors(frac * Uex, frac * Fex)
    infos = [dm.newton_iteration_device(1e-10, 200000, L.PRECOND_BLOCK_JACOBI) for _ in range(3)]
    info = infos[-1]                                  # timed as the solver runs it (CG batches replayed as a CUDA graph)
    dm.set_profiling(True)                            # one more pass with events around every SpMV launch
    dm.newton_iteration_device(1e-10, 200000, L.PRECOND_BLOCK_JACOBI)
    spmv_ms, spmv_n = dm.get_profile()
    avg = spmv_ms / max(spmv_n, 1)
    print(json.dumps({"config": name, "elements": model.nelems, "dofs": int(eqid.size), "nnz": int(dm.nnz),
                      "newton_iteration_ms": info["total_ms"], "elements_per_s": model.nelems / info["total_ms"] * 1e3,
                      "assemble_ms": info["assemble_ms"], "solve_ms": info["solve_ms"], "update_ms": info["update_ms"],
                      "cg_iters": info["cg_iters"], "spmv_kernel": dm.spmv_kernel, "spmv_avg_ms": avg,
                      "spmv_gbs": dm.spmv_bytes / (avg * 1e-3) / 1e9 if spmv_n else None,
                      "spmv_frac_of_peak": dm.spmv_bytes / (avg * 1e-3) / 1e9 / PEAK if spmv_n else None,
                      "setup_s": round(setup_s, 2)}))
    dm.close()
