"""CPU oracle of the Newmark dynamic stage solver — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ``dyn_stage_solver!`` (reference src/mech/dyn-solver.jl:203-535, scheme :FE, no seismic input, no autoinc) on top of
the oracle's ``mount_K`` / ``mount_M`` / ``update_state`` and the direct ``solve_system``.  Parity pinning: the reference's
only test of this path (test/dynamic/dyn-solid.jl) just runs it without an assertion, so parity with the reference is
UNPINNED here; the restatement is checked for physical consistency in tests/test_oracle_golden.py (static limit, energy).
"""
from __future__ import annotations

import math

import numpy as np

from . import oracle as O


def dyn_stage_solver(om: O.OracleModel, bc_vals, rho, tspan, nincs=1, alpha=0.0, beta=0.0, tol=0.01, maxits=5, first_stage=True,
                     V0=None, A0=None):
    """``bc_vals(t) -> (Uex, Fex)`` in eq order.  -> dict(success, U, V, A, F, its, history)."""
    ndofs, nu = om.ndofs, om.nu
    st, M = om.mount_M(rho)
    A = np.zeros(ndofs) if A0 is None else A0.copy()
    V = np.zeros(ndofs) if V0 is None else V0.copy()
    F0 = None
    if first_stage:                                                   # initial accelerations (:279-289)
        Uex, Fex = bc_vals(0.0)
        F0 = Fex.copy()
        ok, msg = O.solve_system(M, A, F0, nu)
        assert ok, msg
    om.state_backup()
    T, dT, t = 0.0, 1.0 / nincs, 0.0
    U = np.zeros(ndofs)
    Fin = np.zeros(ndofs)
    Fina = np.zeros(ndofs)
    TFin = np.zeros(ndofs)
    dUa = np.zeros(ndofs)
    its, history, success = 0, [], True
    while T < 1.0 - 1e-7:
        dt = tspan * dT
        Uex, Fex = bc_vals(t + dt)
        Fex_Fin = Fex - Fina
        dUa[:] = 0.0
        dUi = Uex.copy()
        residue, nfails, converged = 0.0, 0, False
        for it in range(1, maxits + 1):
            if it > 1:
                dUi[:] = 0.0
            lastres = residue
            st, K = om.mount_K()
            st, M = om.mount_M(rho)
            C = alpha * M + beta * K
            Kp = K + (4 / dt ** 2) * M + (2 / dt) * C
            dFp = Fex_Fin + M @ (A + 4 * V / dt - 4 * dUa / dt ** 2) + C @ (V - 2 * dUa / dt)
            ok, msg = O.solve_system(Kp.tocsc(), dUi, dFp, nu)
            assert ok, msg
            om.state_restore()
            dUt = dUa + dUi
            dFin, st = om.update_state(dUt)
            Fina = Fin + dFin
            Va = -V + 2 * dUt / dt
            Aa = -A + 4 * (dUt - V * dt) / dt ** 2
            TFin = Fina + C @ Va + M @ Aa
            residue = float(np.max(np.abs((Fex - TFin)[:nu])))
            dUa += dUi
            Fex_Fin = Fex - Fina
            Fex_Fin[nu:] = 0.0
            its += 1
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
            if nfails == 3:
                break
        if not converged:
            om.state_restore()
            success = False
            break
        Fin = Fina
        U += dUa
        om.state_backup()
        A, V = Aa, Va
        T += dT
        t += dt
        history.append((t, U.copy(), V.copy(), A.copy()))
    return dict(success=success, U=U, V=V, A=A, F=TFin, F0=F0, its=its, history=history)
