#!/usr/bin/env python
"""bench.py — elements/s per Newton iteration (assembly + PCG + state update) of the mechanical hot path.

Workload (BASELINE.json configs[2], the configuration the metric's target is quoted on): HEX20 block of n^3 elements
(n=100 -> 1 M elements, 12.27 M dofs), von Mises E=210e6 nu=0.3 fy=240e3 H=0, bottom clamped, rigid footing
= prescribed uz on a 0.2 x 0.2 patch in the middle of the top face, ten equal load increments.  A "step"
is ONE Newton iteration of the first increment: mount_K on the plastic trial state -> solve_system! (PCG to cg_rtol) ->
state restore -> update_state!.  Every timed step does identical work (each restarts from the converged state,
mech-solver.jl:333).

  value   : device-timed (CUDA events on the library's stream), vectors resident in HBM; whole job = all ranks' elements
  e2e     : the same iteration through the C ABI with HOST buffers (pinned), H2D/D2H copies inside the timed region
  roofline: the dominant kernel (block-CSR SpMV of the PCG), algorithmic bytes / live CUDA-event duration
  cpu_baseline: the CPU oracle (restated reference path: COO assembly -> sparse -> direct LU -> state update) on a
            bounded sample (rank 0, N=1 only)

`--impl reference` times that CPU path alone.  Multi-GPU (torchrun, one rank per GPU): STRONG scaling by default — the same
1 M-element configs[2] problem is partitioned over the N GPUs by recursive coordinate bisection with duplicated halo
elements (the CG iteration count is then the same for every N, so the driver's efficiency measures kernels + exchange and
not the growth of the Jacobi-PCG iteration count with the domain, profiles/README.md); the PCG exchanges halo entries of p
and two small all-reduces per iteration over NCCL.  `--scaling weak` grows the block to (Nx*n, Ny*n, Nz*n) elements instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "elements/sec per Newton iteration (assembly+PCG)"
UNIT = "elements/s"
GRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def footing_model(n, mult=(1, 1, 1), native=True):
    """``native=False``: mesh and dof numbering in numpy — the CPU reference arm must not map libamaru_b200.so."""
    from amaru_jl_b200 import Block, FEModel, MechContext, MechSolid, Mesh, NodeBC, VonMises
    mx, my, mz = mult
    mesh = Mesh(Block([[0, 0, 0], [mx, my, mz]], nx=n * mx, ny=n * my, nz=n * mz, cellshape="HEX20", tag="solids"),
                native=native)
    model = FEModel(mesh, [("solids", MechSolid, VonMises, dict(E=210e6, nu=0.3, fy=240e3, H=0.0))], MechContext())
    cx, cy = mx / 2.0, my / 2.0
    bcs = [("z==0", NodeBC(ux=0, uy=0, uz=0)),
           (f"z=={mz} and x>={cx - 0.1} and x<={cx + 0.1} and y>={cy - 0.1} and y<={cy + 0.1}", NodeBC(uz=-0.01))]
    return model, bcs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_reference_step(n, solver="pcg", steps=1, warmup=0, cg_rtol=1e-10):
    """The restated reference CPU path on a bounded sample of the same footing problem (n^3 HEX20 elements), one Newton
    iteration per step: mount_K (COO -> sparse, mech-solver.jl:78-110) + solve_system! + update_state!, on all host cores.
      solver="direct": lu(K11) like the reference (solver.jl:42-43), scipy SuperLU standing in for UMFPACK, COLAMD column
                       ordering (measured here on the 12^3 sample: COLAMD 14.5 s, MMD_ATA 19.7 s, MMD_AT_PLUS_A 28.4 s);
      solver="pcg":    the same algorithm the GPU path runs (Jacobi-PCG to cg_rtol, oracle orc_pcg_jacobi, OpenMP).
    -> dict(value=elements/s, cores, seconds per step and per phase, cg iterations, achieved GB/s of the PCG)"""
    from oracle import oracle as O
    model, bcs = footing_model(n, native=False)
    eqid, nu, setup = model.configure_dofs(bcs, native=False)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    om = O.OracleModel(model.flatten(), eqid, eqid.size, nu)
    om.state_backup()
    dUex = 0.1 * Uex
    times, ph = [], {}
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        st, K = om.mount_K()
        t1 = time.perf_counter()
        U, F = dUex.copy(), 0.1 * Fex
        if solver == "direct":
            ok, msg = O.solve_system(K, U, F, nu)
            its, gbs = 0, None
        else:
            Kr = K.tocsr()
            K11 = Kr[:nu, :nu]
            rhs = F[:nu] - Kr[:nu, nu:] @ U[nu:]
            tp = time.perf_counter()
            x, its, rr = O.pcg_jacobi(K11, rhs, rtol=cg_rtol)
            tpcg = time.perf_counter() - tp
            U[:nu] = x
            F[nu:] = Kr[nu:, :] @ U
            ok, msg = rr <= 1.5 * cg_rtol, "pcg did not converge"
            gbs = its * (12.0 * K11.nnz + 148.0 * nu) / tpcg / 1e9     # SURVEY.md §8(d) B_cg
            ph["pcg_s"] = tpcg
        t2 = time.perf_counter()
        om.state_restore()
        dF, st2 = om.update_state(U)
        t3 = time.perf_counter()
        if s >= warmup:
            times.append(t3 - t0)
            ph.update(mount_K_s=t1 - t0, solve_s=t2 - t1, update_s=t3 - t2, cg_iters=its, pcg_gbs=gbs, nnz=int(K.nnz))
        assert st == 0 and ok and st2 == 0, (st, msg, st2)
    t = float(np.mean(times))
    return dict(value=model.nelems / t, cores=O.num_threads(), t=t, nelems=model.nelems, ndofs=int(eqid.size), **ph)


def cpu_sample_text(n, r, solver):
    how = ("COO mount_K + scipy SuperLU direct solve (COLAMD) standing in for UMFPACK lu(K11) + update_state!" if solver == "direct"
           else f"COO mount_K + Jacobi-PCG to 1e-10 on CSR ({r['cg_iters']} iterations, {r['pcg_gbs']:.1f} GB/s, OpenMP) + update_state!")
    return (f"HEX20 {n}^3 = {r['nelems']} elements / {r['ndofs']} dofs von Mises footing, one Newton iteration: {how}; "
            f"{r['t']:.1f} s per step (mount_K {r['mount_K_s']:.1f} s, solve {r['solve_s']:.1f} s, update {r['update_s']:.2f} s) "
            f"on {r['cores']} host threads; CPU oracle port (the reference is Julia-only)")


def run_reference(args, rank, world):
    """CPU arm: (ii) the same-algorithm PCG path is the line's value, (i) the reference's own direct
    solve is reported beside it on the largest sample SuperLU finishes in seconds.  Never loads libamaru_b200.so."""
    if rank != 0:
        return
    n = args.ref_size
    warm = min(args.warmup, 1)
    r = cpu_reference_step(n, "pcg", steps=args.steps, warmup=warm, cg_rtol=args.cg_rtol)
    d = cpu_reference_step(args.cpu_direct_size, "direct", steps=1, warmup=0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * r["t"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"configs[2]: HEX20 von Mises footing, bounded sample {n}x{n}x{n} = {r['nelems']} elements "
                                   f"({r['ndofs']} dofs), one Newton iteration per step, CPU Jacobi-PCG path",
                       "sample_n": n, "cg_rtol": args.cg_rtol, "cg_iters_per_step": r["cg_iters"]},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": cpu_sample_text(n, r, "pcg"),
                             "direct": {"value": d["value"], "unit": UNIT, "cores": d["cores"],
                                        "sample": cpu_sample_text(args.cpu_direct_size, d, "direct")}},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "repo_so_mapped": sorted({ln.split()[-1][len(ROOT) + 1:] for ln in open("/proc/self/maps")
                                      if ln.rstrip().endswith(".so") and ROOT in ln})}
    assert not any("libamaru_b200" in x for x in line["repo_so_mapped"]), "the reference arm must not map the product library"
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=100, help="HEX20 elements per side (100 -> 1 M elements; per GPU under --scaling weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-size", type=int, default=24, help="elements per side of the cpu_baseline sample (PCG leg)")
    ap.add_argument("--cpu-direct-size", type=int, default=12, help="elements per side of the direct-solve sample")
    ap.add_argument("--ref-size", type=int, default=40, help="elements per side of the --impl reference sample")
    ap.add_argument("--operator", default=None, choices=["ebe", "csr"], help="CG operator (default: the library's, ebe)")
    ap.add_argument("--cg-rtol", type=float, default=1e-10)
    ap.add_argument("--cg-maxit", type=int, default=200000)
    ap.add_argument("--precond", default="block-jacobi", choices=["jacobi", "block-jacobi"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from amaru_jl_b200 import lib as L
    if not torch.cuda.is_available() or L.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    # stdout carries the JSON line only: native libraries (NCCL prints its version banner on stdout when the first
    # communicator is made) are pointed at stderr for the duration of the run, the line is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    pc = L.PRECOND[args.precond]
    view = None
    n = args.size
    mult = GRID.get(world, (world, 1, 1)) if args.scaling == "weak" else (1, 1, 1)
    t_setup = time.perf_counter()
    model, bcs = footing_model(n, mult)
    eqid, nu, setup = model.configure_dofs(bcs)
    Uex, Fex = model.get_bc_vals(eqid, setup)
    ndofs = int(eqid.size)
    if world == 1:
        dm = L.DeviceModel(model.flatten(), eqid, ndofs, nu, device=local_rank)
        nlocal_elems, nlocal_nodes = model.nelems, model.nnodes
    else:
        from amaru_jl_b200.partition import local_flat, partition_mesh
        view = partition_mesh(model.coords, model.conn, world, rank)
        lf, eql = local_flat(model.flatten(), eqid, view)
        uid = [L.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        dm = L.DeviceModel(lf, eql, ndofs, nu, device=local_rank, view=view, nccl_uid=uid[0])
        p2p_on = False
        if os.environ.get("AMARU_P2P", "1") == "1":   # CG-loop exchanges through peer memory instead of NCCL (all ranks or none)

            def gather(obj):
                out = [None] * world
                dist.all_gather_object(out, obj)
                return out
            p2p_on = dm.p2p_connect(gather)
        nlocal_elems, nlocal_nodes = int(view.elem_gid.size), int(view.node_gid.size)
    t_setup = time.perf_counter() - t_setup
    if args.operator:
        dm.set_operator(args.operator)
    dUex, dFex = 0.1 * Uex, 0.1 * Fex                      # first of ten equal increments
    dm.state_backup()
    dm.set_device_vectors(dUex, dFex)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: the first pass assembles the elastic K; later passes run on the plastic trial state
    for _ in range(max(args.warmup, 3)):
        info = dm.newton_iteration_device(args.cg_rtol, args.cg_maxit, pc)
    # ---- timed region (device-resident)
    sampler = ClockSampler(local_rank)
    sampler.start()
    dm.set_profiling(True)
    l0 = dm.launches
    barrier()
    w0 = time.perf_counter()
    dev_ms, phases = 0.0, []
    for _ in range(args.steps):
        info = dm.newton_iteration_device(args.cg_rtol, args.cg_maxit, pc)
        dev_ms += info["total_ms"]
        phases.append(info)
    barrier()
    wall_s = time.perf_counter() - w0
    gpu_launches = dm.launches - l0
    spmv_ms, spmv_n = dm.get_profile()
    dm.set_profiling(False)

    # ---- e2e: the reference-facing C-ABI calls with pinned HOST buffers, copies inside the timed region
    e2e_s, resid = None, None
    if not args.no_e2e:
        hU = torch.empty(ndofs, dtype=torch.float64).pin_memory().numpy()
        hF = torch.empty(ndofs, dtype=torch.float64).pin_memory().numpy()
        hdF = torch.empty(ndofs, dtype=torch.float64).pin_memory().numpy()
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            hU[:] = dUex
            hF[:] = dFex
            dm.assemble_K()
            dm.solve(hU, hF, args.cg_rtol, args.cg_maxit, pc)
            dm.state_restore()
            dm.update_state(hU, hdF)
            resid = float(np.abs(hF[:nu] - hdF[:nu]).max())   # the Newton residual read back on the host
        barrier()
        e2e_s = time.perf_counter() - e0
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- parity record: scalars of the Newton iteration just computed through the host ABI (solution U, ΔFin, plastic
    # IPs of the trial state), compared with the committed single-GPU values of the same problem (profiles/): the line
    # itself shows that the partitioned answer is the single-GPU answer (contract: u within 1e-8 relative)
    parity = None
    if not args.no_e2e:
        st = dm.get_state()
        if world == 1:
            npl = int((st["dlam"] > 0).sum())
        else:
            own = np.repeat(np.asarray(view.elem_owned, dtype=bool), model.nip)
            tpl = torch.tensor([int((st["dlam"][own] > 0).sum())], dtype=torch.int64, device="cuda")
            dist.all_reduce(tpl)
            npl = int(tpl.item())
        parity = {"u_max": float(np.abs(hU[:nu]).max()), "u_l2": float(np.sqrt(np.dot(hU[:nu], hU[:nu]))),
                  "react_l2": float(np.sqrt(np.dot(hF[nu:], hF[nu:]))), "fint_l2": float(np.sqrt(np.dot(hdF, hdF))),
                  "plastic_ips": npl}
        try:
            ref = json.load(open(os.path.join(ROOT, "profiles", "bench_parity_ref.json"))).get(f"hex20_footing_{n}")
        except Exception:
            ref = None
        if ref is not None and args.scaling == "strong":
            rd = {k: abs(parity[k] - ref[k]) / abs(ref[k]) for k in ("u_max", "u_l2", "react_l2", "fint_l2")}
            parity.update(ref=ref, rel_diff=rd, plastic_ips_equal=parity["plastic_ips"] == ref["plastic_ips"],
                          ok=bool(max(rd["u_max"], rd["u_l2"]) < 1e-8 and max(rd["react_l2"], rd["fint_l2"]) < 1e-7
                                  and parity["plastic_ips"] == ref["plastic_ips"]))

    t_dev = dev_ms / 1e3
    if dist is not None:
        tt = torch.tensor([t_dev, e2e_s or 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, e2e_max = float(tt[0]), float(tt[1])
        e2e_s = e2e_max if e2e_s is not None else None
    nelem_total = model.nelems
    value = nelem_total * args.steps / t_dev

    # ---- roofline of the dominant kernel (SpMV): algorithmic bytes per launch / average live duration (this rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    nd = model.ndim
    nblk = dm.nblocks
    nloc = nlocal_nodes * nd
    spmv_bytes = int(dm.spmv_bytes)   # values + tile records of the stored format + x once + y once
    avg_spmv_ms = spmv_ms / max(spmv_n, 1)
    achieved = spmv_bytes / (avg_spmv_ms * 1e-3) / 1e9 if spmv_n else None
    traffic = None
    is_ebe = dm.spmv_kernel.startswith("k_ebe")
    for tf in ("ebe_traffic_r2.json", "spmv_traffic_r1.json"):
        try:   # DRAM bytes per operator application of the same kernel on the same problem from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", tf)))
            if tj.get("kernel", "k_spmv_stream2<3,true>").split("(")[0].strip() == dm.spmv_kernel and abs(tj["algorithmic_bytes_per_launch"] - spmv_bytes) < 0.02 * spmv_bytes:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
    # FP64 work of one matrix-free application (ebe.cu): per IP nn*nd*nd + 2*nd^3 + 30 FMA, per element nn*nip*nd*nd FMA
    fp64 = None
    if is_ebe and spmv_n:
        nn_, q_ = model.conn.shape[1], model.nip
        fma = nlocal_elems * (q_ * (nn_ * nd * nd + 2 * nd ** 3 + 30) + nn_ * q_ * nd * nd)
        try:
            fpk = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r2.json")))["fp64_tflops_sustained"]
        except Exception:
            fpk = None
        ach = 2.0 * fma / (avg_spmv_ms * 1e-3) / 1e12
        fp64 = {"achieved_tflops": ach, "peak_tflops": fpk, "frac": ach / fpk if fpk else None,
                "peak_source": "profiles/fp64_peak_r2.json (DFMA microbenchmark on this pool's B200, sustained)"}
    kdesc = (("matrix-free tangent operator on the FP64 tensor cores (DMMA) + fused p.Ap, x / y of a 64-element patch in shared "
              "memory, one launch per application" if "patch" in dm.spmv_kernel else
              "matrix-free tangent operator (DMMA) + fused p.Ap, one persistent cooperative launch over the element colours")
             if is_ebe else "TMA-streamed, warp-specialised block-CSR SpMV + p.Ap dot")
    cg_iters = [p["cg_iters"] for p in phases]
    # whole-iteration algorithmic bytes of this rank
    S = 14
    nip = nlocal_elems * model.nip
    # vector kernels per dof: update reads p,q,x,r (32) + M^-1 (16: symmetric 3x3 inverse per node; 8 Jacobi), writes x,r,z (24);
    # p-update reads z,p and writes p (24)
    b_cg = spmv_bytes + nloc * (56 + (16 if args.precond == 'block-jacobi' else 8) + 24)
    b_asm = nlocal_elems * 20 * 4 + 24 * nlocal_nodes + nip * 56 + nlocal_elems * 400 * 4 + 8 * nblk * nd * nd
    b_upd = nlocal_elems * 20 * 4 + 48 * nlocal_nodes + 2 * nip * 8 * S + 8 * nloc
    b_it = b_asm + b_upd + float(np.mean(cg_iters)) * b_cg + 2 * nip * 8 * S
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"configs[2]: HEX20 von Mises footing, {n * mult[0]}x{n * mult[1]}x{n * mult[2]} elements over {world} GPU(s), "
                               "one Newton iteration per step", "elements": nelem_total, "dofs": ndofs, "nnz_rank0": int(dm.nnz),
                   "cg_rtol": args.cg_rtol, "precond": args.precond, "cg_iters_per_step": cg_iters,
                   "operator": "ebe (matrix-free; K is still assembled every iteration for the preconditioner and get_csr)" if is_ebe else "csr",
                   "l2": "inputs (K = %.1f GB per GPU) larger than L2" % (dm.nnz * 8 / 1e9),
                   "parallelism": (f"dd{world} (element partition + halo elements, " +
                                   ("peer-memory halo push + all-reduce kernels)" if p2p_on else "NCCL halo exchange)"))
                   if world > 1 else "dd1",
                   "setup_s": round(t_setup, 2)},
        "phases_ms": {k: float(np.mean([p[k] for p in phases])) for k in ("assemble_ms", "solve_ms", "update_ms")},
        "ms_per_cg_iteration": float(np.mean([p["solve_ms"] / max(p["cg_iters"], 1) for p in phases])),
        "iteration_gbs_rank0": b_it / (t_dev / args.steps) / 1e9,
        "iteration_frac_of_peak": b_it / (t_dev / args.steps) / 1e9 / peak,
        "wall_s_timed_region": wall_s,
        "gpu_launches": int(gpu_launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": f"{dm.spmv_kernel} ({kdesc}), rank 0",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": traffic, "bytes_per_launch": spmv_bytes, "avg_launch_ms": avg_spmv_ms,
                     "launches_timed": int(spmv_n), "fp64": fp64,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
    }
    if parity is not None:
        line["parity"] = parity
    if e2e_s is not None:
        line["e2e"] = {"value": nelem_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 3 * ndofs * 8,
                       "d2h_bytes_per_step": 3 * ndofs * 8, "newton_residual": resid}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_step(args.cpu_size, "pcg", steps=1, warmup=0, cg_rtol=args.cg_rtol)
        d = cpu_reference_step(args.cpu_direct_size, "direct", steps=1, warmup=0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": cpu_sample_text(args.cpu_size, r, "pcg"),
                                "direct": {"value": d["value"], "unit": UNIT, "cores": d["cores"],
                                           "sample": cpu_sample_text(args.cpu_direct_size, d, "direct")}}
    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    dm.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
