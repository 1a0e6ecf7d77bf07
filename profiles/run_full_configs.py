"""BASELINE.json configs[1..4] driven END TO END through the public drivers (solve / solve_dynamic) at full size.

    python profiles/run_full_configs.py --config 3            # HEX20 1 M von Mises footing, 10 increments, one B200
    python profiles/run_full_configs.py --config 5            # HEX8 2 M Newmark, 100 time steps
    python profiles/run_full_configs.py --config 4 --ngpus 8  # TET10 5 M Drucker-Prager under gravity on 8 B200 (one handle)
    python profiles/run_full_configs.py --config 2            # HEX8 200 k linear elastic, one static solve

--scale S shrinks every edge count by S (the 2-GPU rehearsal of the 8-GPU run).  One JSON record per run is appended to
--out: per Newton iteration (stage, increment, iteration, residue, CG iterations, CG residual, wall seconds), the status, the
wall time of the whole solve, and size-independent checks of the result (equilibrium of the reactions, finite fields,
plastic integration points).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amaru_jl_b200 import (Block, BodyC, DruckerPrager, FEModel, LinearElastic, MechAnalysis, MechContext, MechSolid, Mesh,  # noqa: E402
                           NodeBC, SurfaceBC, VonMises, addstage, solve)
from amaru_jl_b200.dyn_solver import solve_dynamic  # noqa: E402
from amaru_jl_b200.model import DynamicAnalysis  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True, choices=[2, 3, 4, 5])
ap.add_argument("--ngpus", type=int, default=1)
ap.add_argument("--partitioner", default="rcb")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--nincs", type=int, default=0)
ap.add_argument("--autoinc", action="store_true")
ap.add_argument("--cg-rtol", type=float, default=1e-10)
ap.add_argument("--uz", type=float, default=-0.01, help="config 3: prescribed footing displacement")
ap.add_argument("--out", default="gpurun_out/full_configs_r2.jsonl")
args = ap.parse_args()


def sc(n):
    return max(2, int(round(n / args.scale)))


t_all = time.perf_counter()
kw = dict(cg_rtol=args.cg_rtol, ngpus=args.ngpus, partitioner=args.partitioner)
checks = {}
if args.config == 2:
    mesh = Mesh(Block([[0, 0, 0], [2, 1, 0.8]], nx=sc(100), ny=sc(50), nz=sc(40), cellshape="HEX8", tag="solids"))
    model = FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=100.0, nu=0.2))], MechContext())
    ana = MechAnalysis(model)
    addstage(ana, [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==0.8", SurfaceBC(tz=-10.0))], nincs=args.nincs or 1)
    name = "configs[1]: HEX8 linear-elastic block, single static solve"
    t0 = time.perf_counter()
    status = solve(ana, **kw)
    base = np.abs(model.coords[:, 2]) < 1e-9
    checks["sum_reaction_z_over_load"] = float(model.F[base, 2].sum() / 20.0)
elif args.config == 3:
    n = sc(100)
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=n, ny=n, nz=n, cellshape="HEX20", tag="solids"))
    model = FEModel(mesh, [("solids", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0.0))], MechContext())
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)),
           ("z==1 and x>=0.4 and x<=0.6 and y>=0.4 and y<=0.6", NodeBC(uz=args.uz))]
    ana = MechAnalysis(model)
    addstage(ana, bcs, nincs=args.nincs or 10)
    name = f"configs[2]: HEX20 von Mises footing (uz = {args.uz}), 10 Newton-Raphson load increments" + (" (autoinc)" if args.autoinc else "")
    t0 = time.perf_counter()
    status = solve(ana, autoinc=args.autoinc, **kw)
    base = np.abs(model.coords[:, 2]) < 1e-9
    top = model.select_nodes("z==1 and x>=0.4 and x<=0.6 and y>=0.4 and y<=0.6")
    checks["footing_force"] = float(model.F[top, 2].sum())
    checks["base_reaction_z"] = float(model.F[base, 2].sum())
    checks["balance_rel"] = float(abs(model.F[top, 2].sum() + model.F[base, 2].sum()) / max(abs(model.F[top, 2].sum()), 1e-300))
elif args.config == 4:
    n = sc(94)
    mesh = Mesh(Block([[0, 0, 0], [1, 1, 1]], nx=n, ny=n, nz=n, cellshape="TET10", tag="solids"))
    model = FEModel(mesh, [("solids", MechSolid, DruckerPrager, dict(E=100.0, nu=0.25, alpha=0.05, kappa=0.1))], MechContext())
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("x==0", NodeBC(ux=0)), ("y==0 || y==1", NodeBC(uy=0)), ("z>=0", BodyC(wz=-0.3))]
    ana = MechAnalysis(model)
    addstage(ana, bcs, nincs=args.nincs or 4)
    name = "configs[3]: TET10 Drucker-Prager block under gravity, 4 increments"
    t0 = time.perf_counter()
    status = solve(ana, tol=1e-3, autoinc=args.autoinc, **kw)
    base = np.abs(model.coords[:, 2]) < 1e-9
    checks["base_reaction_z_over_weight"] = float(model.F[base, 2].sum() / 0.3)
else:
    mesh = Mesh(Block([[0, 0, 0], [2, 1, 1]], nx=sc(200), ny=sc(100), nz=sc(100), cellshape="HEX8", tag="solids"))
    model = FEModel(mesh, [("solids", MechSolid, LinearElastic, dict(E=30e6, nu=0.2, rho=24.0))], MechContext())
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)), ("z==1 && x>=0.9 && x<=1.1", NodeBC(fz=-10.0))]
    ana = DynamicAnalysis(model)
    nst = args.nincs or 100
    addstage(ana, bcs, tspan=1e-4 * nst, nincs=nst)
    name = f"configs[4]: HEX8 Newmark (consistent mass, Rayleigh damping), {nst} time steps"
    t0 = time.perf_counter()
    status = solve_dynamic(ana, alpha=4.2038, beta=174.28e-6, tol=1e-4, **kw)
    top = model.select_nodes("z==1 && x>=0.9 && x<=1.1")
    checks["mean_uz_loaded_nodes"] = float(model.U[top, 2].mean())
    checks["mean_vz_loaded_nodes"] = float(model.V[top, 2].mean())
t_solve = time.perf_counter() - t0

st = model.state
its = [dict(stage=s.get("stage"), inc=s.get("inc"), it=s.get("it"), residue=s.get("residue"), cg_iters=s.get("cg_iters"),
            cg_relres=s.get("cg_relres"), wall_s=round(s.get("wall_s", 0.0), 4)) for s in ana.stats]
rec = {
    "workload": name, "elements": int(model.nelems), "nodes": int(model.nnodes), "dofs": int(model.ndofs), "ngpus": args.ngpus,
    "partitioner": args.partitioner if args.ngpus > 1 else None, "scale": args.scale, "autoinc": args.autoinc,
    "success": bool(status.success), "message": status.message, "increments": len(ana.records), "newton_iterations": len(its),
    "cg_iterations_total": int(sum(i["cg_iters"] or 0 for i in its)), "solve_wall_s": round(t_solve, 2),
    "total_wall_s": round(time.perf_counter() - t_all, 2),
    "elements_per_s_per_newton_iteration_wall": float(model.nelems * len(its) / t_solve) if its else None,
    "u_max": float(np.abs(model.U).max()), "finite": bool(np.isfinite(model.U).all() and np.isfinite(st["sigma"]).all()),
    "plastic_ips": int((st["epa"] > 0).sum()) if args.config != 4 else int((st["dlam"] != 0).sum()),
    "checks": checks, "iterations": its,
}
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "a") as f:
    f.write(json.dumps(rec) + "\n")
print(json.dumps({k: v for k, v in rec.items() if k != "iterations"}))
for i in its[:200]:
    print(i)
