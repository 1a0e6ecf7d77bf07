"""``solve_dynamic(ana, ...)`` — drop-in for the reference's ``solve!(ana::DynamicAnalysis; ...)`` (next tier, config 5).

Average-acceleration Newmark with Rayleigh damping, mirrored from ``dyn_stage_solver!`` (reference
src/mech/dyn-solver.jl:203-535); the hot calls go through the C ABI:

    mount_M(elems, ndofs)                      -> amaru_assemble_M            (:72-103; constant, assembled ONCE here)
    mount_K(elems, ndofs)                      -> amaru_assemble_K            (:373)
    Kp = K + 4/Δt² M + 2/Δt (αM + βK)          -> amaru_set_system_matrix(1 + 2β/Δt, 4/Δt² + 2α/Δt)      (:376-377)
    M*(...) , C*(...)                          -> amaru_matvec(0, 1, ·), amaru_matvec(β, α, ·)           (:378, :399)
    solve_system!(Kp, ΔUi, ΔFp, nu)            -> amaru_solve                 (:381)
    solve_system!(M, A, Fex, nu)               -> amaru_set_system_matrix(0, 1) + amaru_solve            (:289)
    copyto! / update_state!                    -> amaru_state_restore / amaru_update_state / amaru_state_backup

Keyword arguments are ``dyn_solver_params`` (dyn-solver.jl:153-172) without the seismic options, plus the GPU knobs.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import lib as L
from .model import AmaruError, DynamicAnalysis, ReturnStatus, failure, success


def _round_sig(x, sig):
    if x == 0:
        return 0.0
    return round(x, sig - int(math.floor(math.log10(abs(x)))) - 1)


def solve_dynamic(ana: DynamicAnalysis, tol=0.01, dTmin=1e-7, dTmax=0.1, rspan=0.01, scheme="FE", maxits=5, autoinc=False,
                  quiet=True, alpha=0.0, beta=0.0, cg_rtol=1e-10, cg_maxit=200000, precond="block-jacobi", device=0, ngpus=1,
                  partitioner="rcb", keep_fields=False) -> ReturnStatus:
    if str(scheme).lstrip(":") != "FE":
        raise AmaruError("solve!: only scheme=:FE is available on the B200 path")
    pc = L.PRECOND[precond] if isinstance(precond, str) else int(precond)
    if not ana.stages:
        raise AmaruError("stage_iterator!: No stages have been set")
    status = success()
    for stage in [s for s in ana.stages if s.status != "done"]:
        status = _dyn_stage_solver(ana, stage, tol, dTmin, dTmax, maxits, autoinc, alpha, beta, cg_rtol, cg_maxit, pc, device,
                                   ngpus, partitioner, keep_fields)
        if not status.success:
            stage.status = "failed"
            break
        stage.status = "done"
    return status


def _dyn_stage_solver(ana, stage, tol, dTmin, dTmax, maxits, autoinc, alpha, beta, cg_rtol, cg_maxit, pc, device,
                      ngpus=1, partitioner="rcb", keep_fields=False):
    model = ana.model
    log = ana.log
    log.append(f"Dynamic FE analysis: Stage {stage.id}")
    nincs, nouts, tspan = stage.nincs, stage.nouts, stage.tspan
    saveouts = nouts > 0
    eqid, nu, setup = model.configure_dofs(stage.bcs)
    ndofs = eqid.size
    eqflat = eqid.reshape(-1)
    dm = L.DeviceModel(model.flatten(), eqid, ndofs, nu, device=device, ngpus=ngpus, partitioner=partitioner)
    try:
        st = model.state
        dm.set_state(st["sigma"], st["eps"], st["epa"], st["dlam"])
        dm.state_backup()
        dm.assemble_M(model.elem_rho)                                  # M is constant: once (the reference: every iteration)
        A = np.zeros(ndofs)
        V = np.zeros(ndofs)
        if stage.id == 1:                                              # :251-300
            model.U[...] = 0.0
            model.F[...] = 0.0
            model.V[...] = 0.0
            model.A[...] = 0.0
            dm.assemble_K()
            Uex, Fex = model.get_bc_vals(eqid, setup, 0.0, device=dm)
            dm.set_system_matrix(0.0, 1.0)
            Fex0 = np.ascontiguousarray(Fex.copy())
            dm.solve(A, Fex0, cg_rtol, cg_maxit, pc)                   # solve_system!(M, A, Fex, nu)
            dm.set_system_matrix(1.0, 0.0)
            model.V.reshape(-1)[:] = V[eqflat]
            model.A.reshape(-1)[:] = A[eqflat]
            model.F.reshape(-1)[:] = Fex0[eqflat]
        else:
            V[eqflat] = model.V.reshape(-1)
            A[eqflat] = model.A.reshape(-1)

        dTbk = 0.0
        dTcheck = 1.0 / nouts if saveouts else 1.0
        Tcheck = dTcheck
        T = 0.0
        dT = 1.0 / nincs
        if autoinc:
            dT = min(dT, dTmax, dTcheck)
        t = ana.t
        inc = 0
        U = np.zeros(ndofs)
        Fin = np.zeros(ndofs)
        dFin = np.zeros(ndofs)
        dUa = np.zeros(ndofs)
        dUi = np.zeros(ndofs)
        Fina = np.zeros(ndofs)
        TFin = np.zeros(ndofs)
        Aa = np.zeros(ndofs)
        Va = np.zeros(ndofs)
        solstatus = success()

        while T < 1.0 - dTmin:                                         # :325
            dt = tspan * dT
            inc += 1
            log.append(f"  inc {inc}")
            Uex, Fex = model.get_bc_vals(eqid, setup, t + dt, device=dm)   # loads re-integrated on the device every step (:346)
            Fex_Fin = Fex - Fina
            dUa[:] = 0.0
            dUi[:] = Uex
            residue, nfails, residue1, converged, maxfails = 0.0, 0, 0.0, False, 3
            for it in range(1, maxits + 1):                            # :366
                if it > 1:
                    dUi[:] = 0.0
                lastres = residue
                try:
                    dm.assemble_K()
                    dm.set_system_matrix(1.0 + 2.0 * beta / dt, 4.0 / dt ** 2 + 2.0 * alpha / dt)       # Kp
                    dFp = Fex_Fin + dm.matvec(0.0, 1.0, A + 4 * V / dt - 4 * dUa / dt ** 2) \
                        + dm.matvec(beta, alpha, V - 2 * dUa / dt)
                    dUi = np.ascontiguousarray(dUi)
                    dFp = np.ascontiguousarray(dFp)
                    cg_it, cg_rr = dm.solve(dUi, dFp, cg_rtol, cg_maxit, pc)
                    dm.state_restore()
                    dUt = dUa + dUi
                    dm.update_state(dUt, dFin)
                except L.AmaruStatus as e:
                    if e.code < 0:
                        raise
                    log.append(e.message)
                    break
                Fina = Fin + dFin
                Va = -V + 2 * dUt / dt
                Aa = -A + 4 * (dUt - V * dt) / dt ** 2
                TFin = Fina + dm.matvec(beta, alpha, Va) + dm.matvec(0.0, 1.0, Aa)
                residue = float(np.max(np.abs((Fex - TFin)[:nu]))) if nu else 0.0
                dUa += dUi
                Fex_Fin = Fex - Fina
                Fex_Fin[nu:] = 0.0
                log.append(f"    it {it}  residue: {residue:-10.4e}")
                ana.stats.append(dict(stage=stage.id, inc=inc, it=it, residue=residue, cg_iters=cg_it, cg_relres=cg_rr))
                if it == 1:
                    residue1 = residue
                if residue > tol:
                    Fina = Fina - dFin
                if residue < tol:
                    converged = True
                    break
                if math.isnan(residue):
                    break
                if it > 1 and residue > lastres:
                    break
                if residue > 0.9 * lastres:
                    nfails += 1
                if nfails == maxfails:
                    break
            if converged:                                              # :424-446
                Fin = Fina
                U += dUa
                dm.state_backup()
                A, V = Aa, Va
                model.U.reshape(-1)[:] = U[eqflat]
                model.F.reshape(-1)[:] = TFin[eqflat]
                model.V.reshape(-1)[:] = V[eqflat]
                model.A.reshape(-1)[:] = A[eqflat]
                T += dT
                t += dt
                ana.t = t
                checkpoint = T > Tcheck - dTmin
                if checkpoint:
                    Tcheck += dTcheck
                    model.state = dm.get_state()
                    ana.out += 1                                      # update_records!(checkpoint=true), analysis.jl:83-90
                    if ana.outdir is not None:
                        from .output import save, update_output_data
                        os.makedirs(ana.outdir, exist_ok=True)
                        update_output_data(model, dm)                 # nodal recovery on the device
                        save(model, os.path.join(ana.outdir, f"{ana.outkey}-{ana.out}.vtu"))
                rec = dict(stage=stage.id, inc=inc, T=T, t=t)             # scalars only; full fields are opt-in
                if keep_fields:
                    rec.update(U=model.U.copy(), V=model.V.copy(), A=model.A.copy())
                ana.records.append(rec)
                if autoinc:
                    if dTbk > 0.0:
                        dT = min(dTbk, Tcheck - T)
                        dTbk = 0.0
                    else:
                        q = 1.0                                        # nits is never incremented in the reference (:490)
                        dTtr = min(q * dT, 1.0 / nincs, 1 - T)
                        if T + dTtr > Tcheck - dTmin:
                            dTbk = dT
                            dT = Tcheck - T
                        else:
                            dT = dTtr
                            dTbk = 0.0
            else:
                inc -= 1
                dm.state_restore()
                if autoinc:
                    q = 1 + math.tanh(math.log10(tol / residue1)) if residue1 > 0 else 0.2
                    q = min(max(q, 0.2), 0.9)
                    dT = _round_sig(q * dT, 3)
                    if dT < dTmin:
                        solstatus = failure("solver did not converge")
                        break
                else:
                    solstatus = failure("solver did not converge")
                    break
        model.state = dm.get_state()
        return solstatus
    finally:
        dm.close()
