"""CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of oracle/liboracle.so (plain-C restatement of the reference's element / material / assembly
algorithm, see amaru_oracle.c) plus numpy/scipy restatements of
  mount_K's ``sparse(R,C,V)``          src/mech/mech-solver.jl:102        (scipy coo -> csc, duplicates summed)
  solve_system!                        src/solver.jl:5-79                 (scipy ``splu`` stands in for UMFPACK ``lu``)
  mech_stage_solver!                   src/mech/mech-solver.jl:186-492    (increment / Newton driver, FE scheme)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Parity pinning: see the header of amaru_oracle.c — pinned to the reference's own known answers
(tests/test_oracle_golden.py); below those tolerances parity with the reference is unpinned.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "amaru_oracle.c")):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_J2.restype = C.c_double
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _l(a):
    return a.ctypes.data_as(_lp)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


# ---------------------------------------------------------------------------------------------- small wrappers
def shape_func(shape_id, R):
    nn = lib().orc_shape_nn(shape_id)
    N = np.zeros(nn)
    R3 = np.zeros(3)
    R3[:len(R)] = R
    lib().orc_shape_func(shape_id, _d(R3), _d(N))
    return N


def shape_deriv(shape_id, R):
    nn, nd = lib().orc_shape_nn(shape_id), lib().orc_shape_ndim(shape_id)
    D = np.zeros((nn, nd))
    R3 = np.zeros(3)
    R3[:len(R)] = R
    lib().orc_shape_deriv(shape_id, _d(R3), _d(D))
    return D


def quadrature(shape_id):
    ips = np.zeros((8, 4))
    n = lib().orc_quadrature(shape_id, _d(ips))
    return ips[:n].copy()


def J2(s):
    s = np.ascontiguousarray(s, dtype=np.float64)
    return lib().orc_J2(_d(s))


def dev(s):
    s = np.ascontiguousarray(s, dtype=np.float64)
    d = np.zeros(6)
    lib().orc_dev(_d(s), _d(d))
    return d


def calcDe(E, nu):
    D = np.zeros((6, 6))
    lib().orc_calcDe(C.c_double(E), C.c_double(nu), _d(D))
    return D


def calcDe_planestress(E, nu):
    D = np.zeros((6, 6))
    lib().orc_calcDe_ps(C.c_double(E), C.c_double(nu), _d(D))
    return D


def calcD(kind, params, sig, dlam):
    D = np.zeros((6, 6))
    p = np.zeros(8)
    p[:len(params)] = params
    sig = np.ascontiguousarray(sig, dtype=np.float64)
    st = lib().orc_calcD(kind, _d(p), _d(sig), C.c_double(dlam), _d(D))
    return D, st


def update_ip(kind, params, sig, eps, epa, dlam, deps):
    """-> (status, sig, eps, epa, dlam, dsig)"""
    p = np.zeros(8)
    p[:len(params)] = params
    sig = np.array(sig, dtype=np.float64)
    eps = np.array(eps, dtype=np.float64)
    deps = np.ascontiguousarray(deps, dtype=np.float64)
    a, l = C.c_double(epa), C.c_double(dlam)
    ds = np.zeros(6)
    st = lib().orc_update_ip(kind, _d(p), _d(sig), _d(eps), C.byref(a), C.byref(l), _d(deps), _d(ds))
    return st, sig, eps, a.value, l.value, ds


def elem_stiffness(shape_id, th, Cmat, kind, params, sig, dlam):
    nn, nd = lib().orc_shape_nn(shape_id), lib().orc_shape_ndim(shape_id)
    ne = nn * nd
    K = np.zeros((ne, ne))
    p = np.zeros(8)
    p[:len(params)] = params
    Cm = np.ascontiguousarray(Cmat[:, :nd], dtype=np.float64)
    st = lib().orc_elem_stiffness(shape_id, C.c_double(th), _d(Cm), kind, _d(p),
                                  _d(np.ascontiguousarray(sig)), _d(np.ascontiguousarray(dlam)), _d(K))
    return K, st


# ---------------------------------------------------------------------------------------------- the model
class OracleModel:
    """Flat model description + IP state, same arrays the C ABI of the product takes."""

    def __init__(self, flat: dict, eqid: np.ndarray, ndofs: int, nu: int):
        assert len(flat["batch_shape"]) == 1, "oracle handles one element batch"
        self.shape = int(flat["batch_shape"][0])
        self.th = float(flat["thickness"])
        self.coords = np.ascontiguousarray(flat["coords"], dtype=np.float64)
        self.conn = np.ascontiguousarray(flat["conn"], dtype=np.int32)
        self.nelem = self.conn.shape[0]
        self.elem_mat = np.ascontiguousarray(flat["elem_mat"], dtype=np.int32)
        self.mat_kind = np.ascontiguousarray(flat["mat_kind"], dtype=np.int32).copy()
        if int(flat.get("stressmodel", 0)) == 2:   # :planestress — LinearElastic's own calcDe branch (linear-elastic.jl:99-108)
            assert (self.mat_kind == 1).all(), "planestress: LinearElastic only"
            self.mat_kind[:] = 4
        self.mat_par = np.ascontiguousarray(flat["mat_params"], dtype=np.float64)
        self.axi = int(flat.get("stressmodel", 0)) == 3   # :axisymmetric (hoop row of B, th = 2*pi*r; mech-solid.jl:94-108,143)
        self.eqid = np.ascontiguousarray(eqid, dtype=np.int32)
        self.ndofs, self.nu = int(ndofs), int(nu)
        self.nn = lib().orc_shape_nn(self.shape)
        self.nd = lib().orc_shape_ndim(self.shape)
        self.ne = self.nn * self.nd
        self.nip = quadrature(self.shape).shape[0]
        n = self.nelem * self.nip
        self.sig = np.zeros((n, 6))
        self.eps = np.zeros((n, 6))
        self.epa = np.zeros(n)
        self.dlam = np.zeros(n)
        self._bk = None

    # copy.(State) / copyto! (mech-solver.jl:246,333,391)
    def state_backup(self):
        self._bk = (self.sig.copy(), self.eps.copy(), self.epa.copy(), self.dlam.copy())

    def state_restore(self):
        for dst, src in zip((self.sig, self.eps, self.epa, self.dlam), self._bk):
            dst[...] = src

    def _coo(self, mode, filter_eps, rho=None):
        nt = self.nelem * self.ne * self.ne
        rows = np.empty(nt, dtype=np.int64)
        cols = np.empty(nt, dtype=np.int64)
        vals = np.empty(nt, dtype=np.float64)
        ntrip = C.c_int64(0)
        rho = np.zeros(self.nelem) if rho is None else np.ascontiguousarray(rho, dtype=np.float64)
        lib().orc_set_axisymmetric(int(self.axi))
        st = lib().orc_mount_coo(mode, self.shape, C.c_double(self.th), C.c_int64(self.nelem), _d(self.coords),
                                 _i(self.conn), _i(self.elem_mat), _i(self.mat_kind), _d(self.mat_par), _d(rho),
                                 _i(self.eqid), _d(self.sig), _d(self.dlam), int(filter_eps), _l(rows), _l(cols),
                                 _d(vals), C.byref(ntrip))
        n = ntrip.value
        return st, rows[:n], cols[:n], vals[:n]

    def mount_K(self, filter_eps=True):
        """mount_K (mech-solver.jl:78-110) -> (status, scipy CSC ndofs x ndofs)."""
        st, r, c, v = self._coo(0, filter_eps)
        K = sp.coo_matrix((v, (r, c)), shape=(self.ndofs, self.ndofs)).tocsc()
        K.sum_duplicates()
        K.sort_indices()
        return st, K

    def mount_M(self, rho, filter_eps=True):
        """mount_M (dyn-solver.jl:72-103)."""
        st, r, c, v = self._coo(1, filter_eps, rho)
        M = sp.coo_matrix((v, (r, c)), shape=(self.ndofs, self.ndofs)).tocsc()
        M.sum_duplicates()
        M.sort_indices()
        return st, M

    def symbolic_csr(self):
        """Symbolic pattern (connectivity x eq_id) as CSR rowptr/colind with ascending columns."""
        map_e = self.eqid[self.conn].reshape(self.nelem, self.ne).astype(np.int64)
        r = np.repeat(map_e, self.ne, axis=1).reshape(-1)
        c = np.tile(map_e, (1, self.ne)).reshape(-1)
        P = sp.coo_matrix((np.ones(r.size, dtype=np.int8), (r, c)), shape=(self.ndofs, self.ndofs)).tocsr()
        P.sum_duplicates()
        P.sort_indices()
        return P.indptr.astype(np.int64), P.indices.astype(np.int32)

    def update_state(self, dU):
        """update_state! (mech-solver.jl:124-144) -> (dFin, status)."""
        dU = np.ascontiguousarray(dU, dtype=np.float64)
        dF = np.zeros(self.ndofs)
        lib().orc_set_axisymmetric(int(self.axi))
        st = lib().orc_update_state(self.shape, C.c_double(self.th), C.c_int64(self.nelem), _d(self.coords),
                                    _i(self.conn), _i(self.elem_mat), _i(self.mat_kind), _d(self.mat_par),
                                    _i(self.eqid), C.c_int64(self.ndofs), _d(self.sig), _d(self.eps), _d(self.epa),
                                    _d(self.dlam), _d(dU), _d(dF))
        return dF, st

    def internal_forces(self):
        F = np.zeros(self.ndofs)
        lib().orc_set_axisymmetric(int(self.axi))
        lib().orc_internal_forces(self.shape, C.c_double(self.th), C.c_int64(self.nelem), _d(self.coords),
                                  _i(self.conn), _i(self.eqid), C.c_int64(self.ndofs), _d(self.sig), _d(F))
        return F


def solve_system(K, U, F, nu):
    """solve_system!(K, U, F, nu) (solver.jl:5-79); in place on U[:nu], F[nu:].  -> (ok, message)"""
    K = K.tocsc()
    K11 = K[:nu, :nu]
    K12 = K[:nu, nu:]
    K21 = K[nu:, :nu]
    K22 = K[nu:, nu:]
    F1 = F[:nu]
    U2 = U[nu:]
    F2 = K22 @ U2
    U1 = np.zeros(nu)
    if nu > 0:
        rhs = F1 - K12 @ U2
        try:
            lu = spla.splu(K11.tocsc())
            U1 = lu.solve(rhs)
        except Exception as err:                                      # solver.jl:58-66
            return False, f"solve_system!: {err}"
        F2 = F2 + K21 @ U1
    if nu > 0 and (not np.all(np.isfinite(U1)) or np.max(np.abs(U1)) > 1e8):   # solver.jl:68-71
        return False, "solve_system!: Possible syngular matrix"
    U[:nu] = U1
    F[nu:] = F2
    return True, ""


def pcg_jacobi(K11, b, rtol=1e-10, maxit=100000):
    """Jacobi-PCG on the host cores (same algorithm as the device path), BASELINE.md §3 (ii)."""
    A = K11.tocsr()
    A.sort_indices()
    n = A.shape[0]
    x = np.zeros(n)
    it, rr = C.c_int(0), C.c_double(0)
    rp = A.indptr.astype(np.int64)
    ci = A.indices.astype(np.int32)
    lib().orc_pcg_jacobi(C.c_int64(n), _l(rp), _i(ci), _d(A.data), _d(np.ascontiguousarray(b)), _d(x),
                         C.c_double(rtol), int(maxit), C.byref(it), C.byref(rr))
    return x, it.value, rr.value


def _round_sig(x, sig):
    if x == 0:
        return 0.0
    return round(x, sig - int(math.floor(math.log10(abs(x)))) - 1)


def mech_stage_solver(om: OracleModel, Uex, Fex, nincs=1, tol=0.01, rtol=0.01, rspan=0.01, dT0=0.01,
                      dTmin=1e-7, dTmax=0.1, maxits=5, autoinc=False, nouts=0, solver=None, log=None, scheme="FE"):
    """mech_stage_solver! (mech-solver.jl:186-492), schemes :FE / :ME / :BE / :Ralston (:279-288, predictor-corrector
    :341-350).  Vectors are in eq_id order.

    -> dict(success, message, U, F, incs, its, history=[(T, U, F)])
    ``solver(K, U, F, nu)`` defaults to ``solve_system`` (direct)."""
    solver = solver or solve_system
    p1, q11, a1, a2 = {"FE": (1.0, 1.0, None, None), "ME": (1.0, 1.0, 0.5, 0.5), "BE": (1.0, 1.0, 0.0, 1.0),
                       "Ralston": (2 / 3, 2 / 3, 1 / 4, 3 / 4)}[scheme]       # mech-solver.jl:279-288
    ndofs, nu = om.ndofs, om.nu
    ftol = tol
    om.state_backup()                                                 # StateBk = copy.(State)  :246
    dTcheck = 1.0 / nouts if nouts > 0 else 1.0
    Tcheck = dTcheck
    dTbk = 0.0
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
    ok, message = True, ""
    total_its = 0
    history = []
    while T < 1.0 - dTmin:
        inc += 1
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
        for it in range(1, maxits + 1):
            nits += 1
            if it > 1:
                dUi[:] = 0.0
            lastres = res
            st, K = om.mount_K()
            if st:
                syserror, sysmsg = True, f"mount_K status {st}"
                break
            dUitr = p1 * dUi
            Rtr = q11 * R
            sok, smsg = solver(K, dUitr, Rtr, nu)
            if not sok:
                syserror, sysmsg = True, smsg
                break
            om.state_restore()
            dUt = dUa + dUitr
            dFin, st = om.update_state(dUt)
            if st:
                syserror, sysmsg = True, f"update_state status {st}"
                break
            if scheme == "FE":
                dUi = dUitr
            else:                                                     # corrector step, mech-solver.jl:341-350
                st, K2 = om.mount_K()
                if st:
                    syserror, sysmsg = True, f"mount_K status {st}"
                    break
                K = a1 * K + a2 * K2
                Rc2 = 1.0 * R                                         # solve_system! mutates R[nu+1:end]; R is rebuilt below
                sok, smsg = solver(K, dUi, Rc2, nu)
                if not sok:
                    syserror, sysmsg = True, smsg
                    break
                om.state_restore()
                dUt = dUa + dUi
                dFin, st = om.update_state(dUt)
                if st:
                    syserror, sysmsg = True, f"update_state status {st}"
                    break
            dUa += dUi
            R[:] = dFex - dFin
            R[nu:] = 0.0
            res = np.max(np.abs(R)) if ndofs else 0.0
            with np.errstate(invalid="ignore", divide="ignore"):
                err = np.max(np.abs(dUi)) / np.max(np.abs(dUa))
            if log is not None:
                log.append(f"    it {it}  residue: {res:-10.4e}")
            if it == 1:
                res1 = res
            total_its += 1
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
            converged = False
            if log is not None:
                log.append(sysmsg)
        if converged:
            U += dUa
            F += dFin
            Rc[:] = (1.0 - acr) * Rc + R
            om.state_backup()
            T += dT
            history.append((T, U.copy(), F.copy()))
            checkpoint = T > Tcheck - dTmin
            if checkpoint:
                Tcheck += dTcheck
            if autoinc:
                if dTbk > 0.0:
                    dT = min(dTbk, Tcheck - T)
                    dTbk = 0.0
                else:
                    q = 1 + math.tanh(math.log10(ftol / (res1 + np.finfo(float).eps)))
                    q = max(q, 1.1)
                    dTtr = min(q * dT, dTmax, 1 - T)
                    if T + dTtr > Tcheck - dTmin:
                        dTbk = dT
                        dT = Tcheck - T
                    else:
                        dT = dTtr
                        dTbk = 0.0
        else:
            inc -= 1
            om.state_restore()
            if autoinc:
                q = 1 + math.tanh(math.log10(ftol / (res1 + np.finfo(float).eps)))
                q = min(max(q, 0.2), 0.9)
                if syserror:
                    q = 0.7
                dT = _round_sig(q * dT, 3)
                if dT < dTmin:
                    ok, message = False, "Solver did not converge."
                    break
            else:
                ok, message = False, "Solver did not converge. Try `autoinc=true`. "
                break
    return dict(success=ok, message=message, U=U, F=F, incs=inc, its=total_its, history=history)
