"""``solve(ana, ...)`` — drop-in for the reference's ``solve!(ana::MechAnalysis; ...)`` on the B200 path.

The increment / Newton-Raphson driver is the reference's (src/mech/mech-solver.jl:186-492; stage iteration
src/solver.jl:82-198), kept on the host exactly as the Julia glue keeps it; the three hot calls
of every iteration go through the C ABI:

    mount_K(active_elems, ndofs)            -> amaru_assemble_K          (:327)
    solve_system!(K, ΔUi, R, nu)            -> amaru_solve               (:331)
    copyto!.(State, StateBk)                -> amaru_state_restore       (:333)
    update_state!(active_elems, ΔUt, 0.0)   -> amaru_update_state        (:335)
    copyto!.(StateBk, State)                -> amaru_state_backup        (:391)

Solver keyword arguments are ``mech_solver_params`` (mech-solver.jl:155-168) plus the GPU knobs ``cg_rtol``,
``cg_maxit``, ``precond``, ``device`` and ``ngpus`` / ``partitioner`` (one handle over N B200s of the box; the driver
below does not change).  The schemes :FE (default), :ME, :BE and :Ralston are wired (mech-solver.jl:279-288,341-350); there
is no CPU fallback.  ``ana.records`` gets one line of scalars per converged increment; ``keep_fields=True`` additionally
keeps copies of U and F per increment (hundreds of MB each at the benchmark size, so off by default).
"""
from __future__ import annotations

import math
import os
import time

import numpy as np

from . import lib as L
from .model import AmaruError, MechAnalysis, ReturnStatus, failure, success

_EPS = float(np.finfo(np.float64).eps)


# scheme -> (p1, q11, a1, a2)   (mech-solver.jl:279-288)
_SCHEMES = {"FE": (1.0, 1.0, None, None), "ME": (1.0, 1.0, 0.5, 0.5), "BE": (1.0, 1.0, 0.0, 1.0),
            "Ralston": (2 / 3, 2 / 3, 1 / 4, 3 / 4)}


def _round_sig(x, sig):
    if x == 0:
        return 0.0
    return round(x, sig - int(math.floor(math.log10(abs(x)))) - 1)


def solve(ana: MechAnalysis, tol=0.01, rspan=0.01, rtol=0.01, dT0=0.01, dTmin=1e-7, dTmax=0.1, scheme="FE", maxits=5,
          autoinc=False, quiet=True, cg_rtol=1e-10, cg_maxit=200000, precond="block-jacobi", device=0, ngpus=1,
          partitioner="rcb", keep_fields=False) -> ReturnStatus:
    if not tol > 0 or not rtol > 0 or not 0 < rspan < 1 or not 0 < dT0 <= 1 or not 0 < dTmin < 1 or not 0 < dTmax < 1:
        raise AmaruError("solve!: invalid solver parameters")
    if not 1 <= maxits <= 10:
        raise AmaruError("solve!: 1<=maxits<=10 required")
    scheme = str(scheme).lstrip(":")
    if scheme not in _SCHEMES:
        raise AmaruError("solve!: scheme must be one of :FE, :ME, :BE, :Ralston")
    pc = L.PRECOND[precond] if isinstance(precond, str) else int(precond)
    stages = [s for s in ana.stages if s.status != "done"]           # solver.jl:88 (resume at first stage not done)
    if not ana.stages:
        raise AmaruError("stage_iterator!: No stages have been set")
    status = success()
    for stage in stages:
        status = _mech_stage_solver(ana, stage, tol, rtol, dT0, dTmin, dTmax, rspan, maxits, autoinc, quiet, cg_rtol,
                                    cg_maxit, pc, device, scheme, ngpus, partitioner, keep_fields)
        if not status.success:
            stage.status = "failed"
            break
        stage.status = "done"
    return status


def _mech_stage_solver(ana, stage, tol, rtol, dT0, dTmin, dTmax, rspan, maxits, autoinc, quiet, cg_rtol, cg_maxit, pc,
                       device, scheme="FE", ngpus=1, partitioner="rcb", keep_fields=False):
    model = ana.model
    log = ana.log
    log.append(f"Mechanical FE analysis: Stage {stage.id}")
    ftol = tol
    nincs, nouts = stage.nincs, stage.nouts
    saveouts = nouts > 0

    eqid, nu, setup = model.configure_dofs(stage.bcs)                # unknown dofs first (mech-solver.jl:223)
    ndofs = eqid.size
    log.append(f"unknown dofs: {nu}")
    if stage.id == 1:                                                 # :234-242
        model.U[...] = 0.0
        model.F[...] = 0.0

    p1, q11, a1, a2 = _SCHEMES[scheme]
    dm = L.DeviceModel(model.flatten(), eqid, ndofs, nu, device=device, ngpus=ngpus, partitioner=partitioner)
    try:
        st = model.state                                              # IP state persists across stages (:245)
        dm.set_state(st["sigma"], st["eps"], st["epa"], st["dlam"])
        dm.state_backup()                                             # StateBk = copy.(State)  (:246)

        dTbk = 0.0
        dTcheck = 1.0 / nouts if saveouts else 1.0
        Tcheck = dTcheck
        T = 0.0
        dT = 1.0 / nincs
        if autoinc:
            dT = min(dT, dTmax, dTcheck, dT0)
        inc = 0
        F = np.zeros(ndofs)
        U = np.zeros(ndofs)
        R = np.zeros(ndofs)
        dFin = np.zeros(ndofs)
        dUa = np.zeros(ndofs)
        dUi = np.zeros(ndofs)
        Rc = np.zeros(ndofs)
        Uex, Fex = model.get_bc_vals(eqid, setup, device=dm)          # :269
        solstatus = success()
        eqflat = eqid.reshape(-1)

        while T < 1.0 - dTmin:                                        # :292
            inc += 1
            log.append(f"  inc {inc}")
            dUex, dFex = dT * Uex, dT * Fex
            dTcr = min(rspan, 1 - T)
            acr = min(dT / dTcr, 1.0)
            if T < 1 - rspan:
                dFex = dFex + acr * Rc
            R[:] = dFex
            dUa[:] = 0.0
            dUi[:] = dUex
            nits, err, res, res1 = 0, 0.0, 0.0, 0.0
            converged, syserror, sysmsg = False, False, ""
            for it in range(1, maxits + 1):                           # :319
                nits += 1
                if it > 1:
                    dUi[:] = 0.0
                lastres = res
                t0 = time.perf_counter()
                try:
                    dm.assemble_K()                                   # K = mount_K(...)
                    dUitr = np.ascontiguousarray(p1 * dUi)
                    Rtr = np.ascontiguousarray(q11 * R)
                    cg_it, cg_rr = dm.solve(dUitr, Rtr, cg_rtol, cg_maxit, pc)   # solve_system!
                    dm.state_restore()                                # copyto!.(State, StateBk)
                    dUt = dUa + dUitr
                    dm.update_state(dUt, dFin)                        # ΔFin, status = update_state!
                    if scheme == "FE":
                        dUi = dUitr
                    else:                                             # corrector step (:341-350)
                        dm.tangent_save()
                        dm.assemble_K()                               # K2 = mount_K(...) on the predictor's state
                        dm.tangent_blend(a1, a2)                      # K = a1*K + a2*K2
                        dUi = np.ascontiguousarray(dUi)
                        Rc2 = np.ascontiguousarray(1.0 * R)           # solve_system! overwrites R[nu+1:end] with reactions
                        it2, cg_rr = dm.solve(dUi, Rc2, cg_rtol, cg_maxit, pc)
                        cg_it += it2
                        dm.state_restore()
                        dUt = dUa + dUi
                        dm.update_state(dUt, dFin)
                except L.AmaruStatus as e:
                    if e.code < 0:
                        raise
                    syserror, sysmsg = True, e.message
                    break
                dUa += dUi
                R[:] = dFex - dFin
                R[nu:] = 0.0
                res = float(np.max(np.abs(R))) if ndofs else 0.0
                with np.errstate(invalid="ignore", divide="ignore"):
                    err = float(np.max(np.abs(dUi)) / np.max(np.abs(dUa)))
                log.append(f"    it {it}  residue: {res:-10.4e}")
                ana.stats.append(dict(stage=stage.id, inc=inc, it=it, residue=res, cg_iters=cg_it, cg_relres=cg_rr,
                                      wall_s=time.perf_counter() - t0))
                if it == 1:
                    res1 = res
                if res < ftol:
                    converged = True
                    break
                if err < rtol:
                    converged = True
                    break
                if math.isnan(res):
                    break
                if it > 1 and res > lastres:
                    break
            if syserror:
                log.append(sysmsg)
                converged = False
            if converged:                                             # :384-397
                U += dUa
                F += dFin
                Rc[:] = (1.0 - acr) * Rc + R
                dm.state_backup()
                model.U.reshape(-1)[:] += dUa[eqflat]                 # dof.vals[dof.name]    += ΔUa[i]
                model.F.reshape(-1)[:] += dFin[eqflat]                # dof.vals[dof.natname] += ΔFin[i]
                T += dT
                checkpoint = T > Tcheck - dTmin
                if checkpoint:
                    Tcheck += dTcheck
                    model.state = dm.get_state()                      # what update_records! samples at outputs
                    ana.out += 1                                      # update_records!(checkpoint=true), analysis.jl:83-90
                    if ana.outdir is not None:
                        from .output import save, update_output_data
                        os.makedirs(ana.outdir, exist_ok=True)
                        update_output_data(model, dm)                 # nodal recovery on the device
                        save(model, os.path.join(ana.outdir, f"{ana.outkey}-{ana.out}.vtu"))
                rec = dict(stage=stage.id, inc=inc, T=T, residue=res, nits=nits)   # scalars only
                if keep_fields:
                    rec.update(U=model.U.copy(), F=model.F.copy())
                ana.records.append(rec)
                if autoinc:                                           # :426-455
                    if dTbk > 0.0:
                        dT = min(dTbk, Tcheck - T)
                        dTbk = 0.0
                    else:
                        q = 1 + math.tanh(math.log10(ftol / (res1 + _EPS)))
                        q = max(q, 1.1)
                        dTtr = min(q * dT, dTmax, 1 - T)
                        if T + dTtr > Tcheck - dTmin:
                            dTbk = dT
                            dT = Tcheck - T
                        else:
                            dT = dTtr
                            dTbk = 0.0
            else:                                                     # :456-485
                inc -= 1
                dm.state_restore()
                if autoinc:
                    log.append("      increment failed")
                    q = 1 + math.tanh(math.log10(ftol / (res1 + _EPS)))
                    q = min(max(q, 0.2), 0.9)
                    if syserror:
                        q = 0.7
                    dT = _round_sig(q * dT, 3)
                    if dT < dTmin:
                        solstatus = failure("Solver did not converge.")
                        break
                else:
                    solstatus = failure("Solver did not converge. Try `autoinc=true`. ")
                    break
        model.state = dm.get_state()                                  # IP history goes back to the model
        ana.last_stage = dict(U=U, F=F, eqid=eqid, nu=nu, incs=inc, launches=dm.launches)
        return solstatus
    finally:
        dm.close()
