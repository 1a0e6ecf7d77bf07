"""ctypes binding of libamaru_b200.so (include/amaru_b200.h) — the same calls the Julia glue makes with ``ccall``.

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible, every compute entry point
raises.  Nothing here imports torch; device memory is owned by the library.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .model import AmaruError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libamaru_b200.so")

OK = 0
FAIL_MATERIAL, FAIL_NAN, FAIL_SINGULAR, FAIL_NEG_JACOBIAN, FAIL_CG_NOCONV, FAIL_TANGENT = 1, 2, 3, 4, 5, 6
ERR_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_COMM = -1, -2, -3, -4, -5
PRECOND_JACOBI, PRECOND_BLOCK_JACOBI = 0, 1
PARTITION_RCB, PARTITION_METIS = 0, 1
PARTITIONER = {"rcb": PARTITION_RCB, "metis": PARTITION_METIS}
PRECOND = {"jacobi": PRECOND_JACOBI, "block-jacobi": PRECOND_BLOCK_JACOBI, "block_jacobi": PRECOND_BLOCK_JACOBI,
           "bjacobi": PRECOND_BLOCK_JACOBI}

_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p

# every symbol include/amaru_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "amaru_device_count": (C.c_int, []),
    "amaru_version": (C.c_char_p, []),
    "amaru_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int64, _dp, C.c_int, _i32p, _i64p, _i32p, _i32p,
                               C.c_int, _i32p, _dp, _i32p, C.c_int64, C.c_int64, C.c_int, _i32p, C.c_int,
                               C.POINTER(_vp), C.c_char_p, C.c_int]),
    "amaru_ngpus": (C.c_int, [_vp]),
    "amaru_partition_elements_abi": (C.c_int, [C.c_int, C.c_int, C.c_int64, _dp, C.c_int, _i32p, _i64p, _i32p, _i32p,
                                               C.c_char_p, C.c_int]),
    "amaru_comm_selftest": (C.c_int, [_vp, C.c_int, C.c_char_p, C.c_int]),
    "amaru_create_partitioned": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int64, C.c_int64, _dp,
                                           C.c_int, _i32p, _i64p, _i32p, _i32p, C.c_int, _i32p, _dp, _i32p,
                                           C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, _i32p, _i64p, _i32p, _i64p,
                                           _i64p, _vp, C.c_int, C.POINTER(_vp), C.c_char_p, C.c_int]),
    "amaru_nccl_unique_id": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "amaru_p2p_export": (C.c_int, [_vp, _vp, C.c_char_p, C.c_int]),
    "amaru_p2p_connect": (C.c_int, [_vp, _vp, _i64p, C.c_char_p, C.c_int]),
    "amaru_p2p_enable": (C.c_int, [_vp, C.c_int]),
    "amaru_destroy": (C.c_int, [_vp]),
    "amaru_nip_total": (C.c_int64, [_vp]),
    "amaru_nnz": (C.c_int64, [_vp]),
    "amaru_nblocks": (C.c_int64, [_vp]),
    "amaru_ncolors": (C.c_int, [_vp]),
    "amaru_set_state": (C.c_int, [_vp, _dp, _dp, _dp, _dp, C.c_char_p, C.c_int]),
    "amaru_get_state": (C.c_int, [_vp, _dp, _dp, _dp, _dp, C.c_char_p, C.c_int]),
    "amaru_state_backup": (C.c_int, [_vp]),
    "amaru_state_restore": (C.c_int, [_vp]),
    "amaru_assemble_K": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "amaru_tangent_save": (C.c_int, [_vp, C.c_char_p, C.c_int]),
    "amaru_tangent_blend": (C.c_int, [_vp, C.c_double, C.c_double, C.c_char_p, C.c_int]),
    "amaru_get_csr": (C.c_int, [_vp, _i64p, _i32p, _dp, C.c_char_p, C.c_int]),
    "amaru_solve": (C.c_int, [_vp, _dp, _dp, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), _dp, C.c_char_p,
                              C.c_int]),
    "amaru_update_state": (C.c_int, [_vp, _dp, _dp, C.c_char_p, C.c_int]),
    "amaru_internal_forces": (C.c_int, [_vp, _dp, C.c_char_p, C.c_int]),
    "amaru_assemble_M": (C.c_int, [_vp, _dp, C.c_char_p, C.c_int]),
    "amaru_set_system_matrix": (C.c_int, [_vp, C.c_double, C.c_double, C.c_char_p, C.c_int]),
    "amaru_matvec": (C.c_int, [_vp, C.c_double, C.c_double, _dp, _dp, C.c_char_p, C.c_int]),
    "amaru_loadset_create": (C.c_int, [_vp, C.c_int, C.c_int64, _i32p, C.POINTER(_vp), C.c_char_p, C.c_int]),
    "amaru_loadset_nip": (C.c_int64, [_vp]),
    "amaru_loadset_ip_coords": (C.c_int, [_vp, _dp, C.c_char_p, C.c_int]),
    "amaru_loadset_apply": (C.c_int, [_vp, C.c_int, C.c_double, _dp, _dp, C.c_char_p, C.c_int]),
    "amaru_loadset_destroy": (C.c_int, [_vp]),
    "amaru_recovery_create": (C.c_int, [_vp, _u8p, C.c_char_p, C.c_int]),
    "amaru_recovery_nfields": (C.c_int, [_vp]),
    "amaru_recovery_field": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "amaru_recover_nodal": (C.c_int, [_vp, _dp, C.c_char_p, C.c_int]),
    "amaru_write_vtu": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int64, _dp, C.c_int, _i32p, _i64p, _i32p,
                                  C.c_int, C.POINTER(C.c_char_p), _i32p, _i32p, C.POINTER(_vp),
                                  C.c_int, C.POINTER(C.c_char_p), _i32p, _i32p, C.POINTER(_vp), C.c_char_p, C.c_int]),
    "amaru_mesh_block_sizes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _i64p, _i64p, C.POINTER(C.c_int)]),
    "amaru_mesh_block": (C.c_int, [C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, _i32p, C.c_char_p, C.c_int]),
    "amaru_configure_dofs": (C.c_int, [C.c_int64, C.c_int, _u8p, _i32p, _i64p]),
    "amaru_outer_facets": (C.c_int64, [C.c_int, C.c_int64, _i32p, _i32p, _i64p, C.c_int64, C.POINTER(C.c_int)]),
    "amaru_newton_iteration_device": (C.c_int, [_vp, C.c_double, C.c_int, C.c_int, _dp, C.POINTER(C.c_int), _dp,
                                                C.c_char_p, C.c_int]),
    "amaru_set_device_vectors": (C.c_int, [_vp, _dp, _dp, C.c_char_p, C.c_int]),
    "amaru_time_kernel": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, C.c_char_p, C.c_int]),
    "amaru_set_operator": (C.c_int, [_vp, C.c_int]),
    "amaru_patch_plan_check": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _vp, C.c_int64, _vp, _vp]),
    "amaru_operator_apply": (C.c_int, [_vp, _dp, _dp, C.c_int, _dp, C.c_char_p, C.c_int]),
    "amaru_set_profiling": (C.c_int, [_vp, C.c_int]),
    "amaru_get_profile": (C.c_int, [_vp, _dp, _i64p]),
    "amaru_spmv_bytes": (C.c_int64, [_vp]),
    "amaru_spmv_kernel": (C.c_char_p, [_vp]),
    "amaru_launch_count": (C.c_int64, [_vp]),
}

_LIB = None


def load():
    """Load libamaru_b200.so; fails loudly if it was not built (python __graft_entry__.py build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise AmaruError(f"{LIB_PATH} is missing: build it with `make -C amaru_jl_b200/csrc` "
                             "(there is no CPU fallback for the mechanical hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def device_count():
    return load().amaru_device_count()


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId (rank 0 makes it, the host broadcasts it to the other ranks)."""
    buf, msg = C.create_string_buffer(128), C.create_string_buffer(256)
    st = load().amaru_nccl_unique_id(C.cast(buf, _vp), msg, 256)
    if st != OK:
        raise AmaruError(f"amaru_nccl_unique_id [{st}] {msg.value.decode(errors='replace')}")
    return bytes(buf.raw)


class AmaruStatus(AmaruError):
    """A non-zero status from the library: ``code > 0`` = ReturnStatus failure, ``code < 0`` = usage / CUDA error."""

    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code, self.message = code, msg


class DeviceModel:
    """One stage's device handle (``amaru_create`` ... ``amaru_destroy``)."""

    def __init__(self, flat: dict, eqid: np.ndarray, ndofs: int, nu: int, device: int = 0, view=None, nccl_uid=None,
                 ngpus: int = 1, devices=None, partitioner="rcb"):
        """``amaru_create``: one handle over ``ngpus`` B200s of this box (``devices`` = their ordinals, default
        ``device, device+1, ...``; the library partitions the mesh itself), or — with ``view`` (partition.LocalView) and the
        128-byte ``nccl_uid`` — this rank's handle of a model partitioned by the caller, one process per GPU
        (``amaru_create_partitioned``); ``flat``/``eqid`` are then the LOCAL arrays from ``partition.local_flat`` while
        ``ndofs``/``nu`` stay global."""
        self.lib = load()
        self._msg = C.create_string_buffer(512)
        self.ndofs, self.nu = int(ndofs), int(nu)
        coords = np.ascontiguousarray(flat["coords"], dtype=np.float64)
        bshape = np.ascontiguousarray(flat["batch_shape"], dtype=np.int32)
        bnelem = np.ascontiguousarray(flat["batch_nelem"], dtype=np.int64)
        conn = np.ascontiguousarray(flat["conn"], dtype=np.int32).reshape(-1)
        emat = np.ascontiguousarray(flat["elem_mat"], dtype=np.int32)
        mkind = np.ascontiguousarray(flat["mat_kind"], dtype=np.int32)
        mpar = np.ascontiguousarray(flat["mat_params"], dtype=np.float64).reshape(-1)
        eq = np.ascontiguousarray(eqid, dtype=np.int32).reshape(-1)
        h = _vp()
        self.view = view
        if view is None:
            devs = np.ascontiguousarray(devices if devices is not None else np.arange(int(ngpus)) + int(device), dtype=np.int32)
            if devs.size != int(ngpus):
                raise AmaruError("DeviceModel: len(devices) must equal ngpus")
            part = PARTITIONER[partitioner] if isinstance(partitioner, str) else int(partitioner)
            st = self.lib.amaru_create(int(flat["ndim"]), int(flat["stressmodel"]), float(flat["thickness"]),
                                       coords.shape[0], _d(coords), len(bshape), bshape.ctypes.data_as(_i32p),
                                       bnelem.ctypes.data_as(_i64p), conn.ctypes.data_as(_i32p),
                                       emat.ctypes.data_as(_i32p), len(mkind), mkind.ctypes.data_as(_i32p), _d(mpar),
                                       eq.ctypes.data_as(_i32p), self.ndofs, self.nu, int(ngpus),
                                       devs.ctypes.data_as(_i32p), part, C.byref(h), self._msg, len(self._msg))
        else:
            neigh = np.ascontiguousarray(view.neigh, dtype=np.int32)
            sptr = np.ascontiguousarray(view.send_ptr, dtype=np.int64)
            snod = np.ascontiguousarray(view.send_nodes, dtype=np.int32)
            rst = np.ascontiguousarray(view.recv_start, dtype=np.int64)
            rct = np.ascontiguousarray(view.recv_count, dtype=np.int64)
            uid = (C.c_char * 128).from_buffer_copy(bytes(nccl_uid)) if nccl_uid is not None else None
            st = self.lib.amaru_create_partitioned(
                int(flat["ndim"]), int(flat["stressmodel"]), float(flat["thickness"]), coords.shape[0], int(view.nowned),
                _d(coords), len(bshape), bshape.ctypes.data_as(_i32p), bnelem.ctypes.data_as(_i64p),
                conn.ctypes.data_as(_i32p), emat.ctypes.data_as(_i32p), len(mkind), mkind.ctypes.data_as(_i32p), _d(mpar),
                eq.ctypes.data_as(_i32p), self.ndofs, self.nu, int(view.rank), int(view.nranks), len(neigh),
                neigh.ctypes.data_as(_i32p), sptr.ctypes.data_as(_i64p), snod.ctypes.data_as(_i32p),
                rst.ctypes.data_as(_i64p), rct.ctypes.data_as(_i64p), C.cast(uid, _vp) if uid is not None else None,
                int(device), C.byref(h), self._msg, len(self._msg))
        if st != OK:
            raise AmaruStatus(st, self._msg.value.decode(errors="replace"))
        self.h = h
        self.nip_total = self.lib.amaru_nip_total(h)

    # -- helpers
    def _check(self, st):
        if st != OK:
            raise AmaruStatus(st, self._msg.value.decode(errors="replace"))

    def close(self):
        if getattr(self, "h", None):
            for ls, _ in list(self.__dict__.get("_loadsets", {}).values()):   # load sets hold a pointer to the model
                ls.close()
            self.__dict__.pop("_loadsets", None)
            self.lib.amaru_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ngpus(self):
        return self.lib.amaru_ngpus(self.h)

    def comm_selftest(self, skip_rank=-1):
        """Failure injection for the bounded peer-memory waits (multi-GPU handles); raises AmaruStatus(ERR_COMM) on timeout."""
        self._check(self.lib.amaru_comm_selftest(self.h, int(skip_rank), self._msg, len(self._msg)))

    @property
    def nnz(self):
        return self.lib.amaru_nnz(self.h)

    @property
    def nblocks(self):
        return self.lib.amaru_nblocks(self.h)

    @property
    def ncolors(self):
        return self.lib.amaru_ncolors(self.h)

    @property
    def spmv_bytes(self):
        return self.lib.amaru_spmv_bytes(self.h)

    @property
    def spmv_kernel(self):
        return self.lib.amaru_spmv_kernel(self.h).decode()

    @property
    def launches(self):
        return self.lib.amaru_launch_count(self.h)

    # -- state
    def set_state(self, sigma=None, eps=None, epa=None, dlam=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (sigma, eps, epa, dlam)]
        self._check(self.lib.amaru_set_state(self.h, *[_d(a) for a in arrs], self._msg, len(self._msg)))

    def get_state(self):
        n = self.nip_total
        sig, eps, epa, dl = np.empty((n, 6)), np.empty((n, 6)), np.empty(n), np.empty(n)
        self._check(self.lib.amaru_get_state(self.h, _d(sig), _d(eps), _d(epa), _d(dl), self._msg, len(self._msg)))
        return dict(sigma=sig, eps=eps, epa=epa, dlam=dl)

    def state_backup(self):
        self._check(self.lib.amaru_state_backup(self.h))

    def state_restore(self):
        self._check(self.lib.amaru_state_restore(self.h))

    # -- the three hot calls
    def assemble_K(self):
        self._check(self.lib.amaru_assemble_K(self.h, self._msg, len(self._msg)))

    def tangent_save(self):
        self._check(self.lib.amaru_tangent_save(self.h, self._msg, len(self._msg)))

    def tangent_blend(self, a1, a2):
        self._check(self.lib.amaru_tangent_blend(self.h, float(a1), float(a2), self._msg, len(self._msg)))

    def solve(self, U, F, cg_rtol=1e-10, cg_maxit=100000, precond=PRECOND_BLOCK_JACOBI):
        """In place on U[:nu] and F[nu:] like solve_system!; returns (iters, relres)."""
        assert U.dtype == np.float64 and F.dtype == np.float64 and U.flags.c_contiguous and F.flags.c_contiguous
        it, rr = C.c_int(0), C.c_double(0)
        self._check(self.lib.amaru_solve(self.h, _d(U), _d(F), float(cg_rtol), int(cg_maxit), int(precond),
                                         C.byref(it), C.byref(rr), self._msg, len(self._msg)))
        return it.value, rr.value

    def update_state(self, dU, dFin=None):
        dU = np.ascontiguousarray(dU, dtype=np.float64)
        if dFin is None:
            dFin = np.empty(self.ndofs)
        self._check(self.lib.amaru_update_state(self.h, _d(dU), _d(dFin), self._msg, len(self._msg)))
        return dFin

    def internal_forces(self):
        F = np.empty(self.ndofs)
        self._check(self.lib.amaru_internal_forces(self.h, _d(F), self._msg, len(self._msg)))
        return F

    def get_csr(self, values=True):
        nnz = self.nnz
        rowptr = np.empty(self.ndofs + 1, dtype=np.int64)
        col = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz) if values else None
        self._check(self.lib.amaru_get_csr(self.h, rowptr.ctypes.data_as(_i64p), col.ctypes.data_as(_i32p), _d(val),
                                           self._msg, len(self._msg)))
        return rowptr, col, val

    # -- next tier (Newmark)
    def assemble_M(self, rho):
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        self._check(self.lib.amaru_assemble_M(self.h, _d(rho), self._msg, len(self._msg)))

    def set_system_matrix(self, a, b):
        self._check(self.lib.amaru_set_system_matrix(self.h, float(a), float(b), self._msg, len(self._msg)))

    def matvec(self, a, b, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.ndofs)
        self._check(self.lib.amaru_matvec(self.h, float(a), float(b), _d(x), _d(y), self._msg, len(self._msg)))
        return y

    # -- multi-GPU: peer-memory path of the CG-loop exchanges (one box, NVLink)
    def p2p_connect(self, all_gather):
        """``all_gather(obj) -> list of every rank's obj in rank order`` (e.g. torch.distributed.all_gather_object).
        Exchanges the cudaIpc handles and the ghost offsets and, if EVERY rank could map its peers, switches the CG loop's
        halo exchange and scalar all-reduces from NCCL to the peer-memory kernels (csrc/halo.cu).  Returns True when the
        peer-memory path is on (all ranks agree), False when every rank stays on NCCL."""
        v = self.view
        buf = C.create_string_buffer(128)
        st = self.lib.amaru_p2p_export(self.h, C.cast(buf, _vp), self._msg, len(self._msg))
        mine = dict(ok=st == OK, handles=bytes(buf.raw), neigh=[int(q) for q in v.neigh],
                    recv_start=[int(s) for s in v.recv_start])
        everyone = all_gather(mine)
        ok = all(e["ok"] for e in everyone)
        if ok:
            allh = b"".join(e["handles"] for e in everyone)
            from .partition import peer_recv_starts
            peer_start = peer_recv_starts(v.rank, mine["neigh"], everyone)
            hb = C.create_string_buffer(allh, len(allh))
            st = self.lib.amaru_p2p_connect(self.h, C.cast(hb, _vp), peer_start.ctypes.data_as(_i64p), self._msg, len(self._msg))
            ok = st == OK
        ok = all(all_gather(bool(ok)))
        if ok:
            self._check(self.lib.amaru_p2p_enable(self.h, 1))
        return ok

    # -- next tier: natural boundary conditions integrated on the device
    def loadset(self, shape_id, nodes):
        return LoadSet(self, shape_id, nodes)

    # -- next tier: output side (nodal_patch_recovery on the device)
    def recovery_create(self, at_bound):
        ab = np.ascontiguousarray(at_bound, dtype=np.uint8)
        self._check(self.lib.amaru_recovery_create(self.h, ab.ctypes.data_as(_u8p), self._msg, len(self._msg)))

    def recovery_fields(self):
        out = []
        name, code = C.create_string_buffer(32), C.c_int(0)
        for i in range(self.lib.amaru_recovery_nfields(self.h)):
            self.lib.amaru_recovery_field(self.h, i, C.byref(code), name, 32)
            out.append(name.value.decode("utf-8"))
        return out

    def recover_nodal(self, nnodes):
        """-> V (nnodes, nfields) like V_rec of nodal_patch_recovery (a transposed view of the field-major buffer)."""
        nf = self.lib.amaru_recovery_nfields(self.h)
        if nf < 0:
            raise AmaruError("recover_nodal: call recovery_create first")
        V = np.zeros((max(nf, 0), nnodes))
        if nf > 0:
            self._check(self.lib.amaru_recover_nodal(self.h, _d(V), self._msg, len(self._msg)))
        return V.T

    # -- measurement hooks
    def set_device_vectors(self, U, F):
        U = np.ascontiguousarray(U, dtype=np.float64)
        F = np.ascontiguousarray(F, dtype=np.float64)
        self._check(self.lib.amaru_set_device_vectors(self.h, _d(U), _d(F), self._msg, len(self._msg)))

    def newton_iteration_device(self, cg_rtol=1e-10, cg_maxit=100000, precond=PRECOND_BLOCK_JACOBI):
        ms = np.zeros(4)
        it, rr = C.c_int(0), C.c_double(0)
        self._check(self.lib.amaru_newton_iteration_device(self.h, float(cg_rtol), int(cg_maxit), int(precond), _d(ms),
                                                           C.byref(it), C.byref(rr), self._msg, len(self._msg)))
        return dict(assemble_ms=ms[0], solve_ms=ms[1], update_ms=ms[2], total_ms=ms[3], cg_iters=it.value,
                    relres=rr.value)

    def time_kernel(self, kind, reps=20, precond=PRECOND_BLOCK_JACOBI):
        t = C.c_double(0)
        self._check(self.lib.amaru_time_kernel(self.h, int(kind), int(precond), int(reps), C.byref(t), self._msg,
                                               len(self._msg)))
        return t.value

    def set_operator(self, kind):
        """CG operator of this handle: "ebe" (matrix-free, default) or "csr" (SpMV on the assembled block-CSR values)."""
        k = {"csr": 0, "ebe": 1}[kind] if isinstance(kind, str) else int(kind)
        self._check(self.lib.amaru_set_operator(self.h, k))

    def operator_apply(self, x, masked=False):
        """(A x, x.Ax) with the CG operator of this handle (``set_operator``); eq_id ordering."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.ndofs)
        pq = C.c_double(0)
        self._check(self.lib.amaru_operator_apply(self.h, _d(x), _d(y), 1 if masked else 0, C.byref(pq), self._msg,
                                                  len(self._msg)))
        return y, pq.value

    def set_profiling(self, on=True):
        self.lib.amaru_set_profiling(self.h, 1 if on else 0)

    def get_profile(self):
        t, n = C.c_double(0), C.c_int64(0)
        self.lib.amaru_get_profile(self.h, C.byref(t), C.byref(n))
        return t.value, n.value


LOAD_KEYS = {"tx": 0, "ty": 1, "tz": 2, "tn": 3, "wx": 0, "wy": 1, "wz": 2}


class LoadSet:
    """The entities (facets or cells) one SurfaceBC / BodyC selected, resident on the device (``amaru_loadset_*``)."""

    def __init__(self, dm: DeviceModel, shape_id: int, nodes: np.ndarray):
        self.dm = dm
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        self.nents = nodes.shape[0]
        h = _vp()
        st = dm.lib.amaru_loadset_create(dm.h, int(shape_id), self.nents, nodes.ctypes.data_as(_i32p), C.byref(h), dm._msg,
                                         len(dm._msg))
        dm._check(st)
        self.h = h
        self.nip = dm.lib.amaru_loadset_nip(h)

    def ip_coords(self):
        X = np.empty((self.nip, 3))
        self.dm._check(self.dm.lib.amaru_loadset_ip_coords(self.h, _d(X), self.dm._msg, len(self.dm._msg)))
        return X

    def apply(self, key, vals, F):
        """F += nodal forces of the set; ``key`` in tx ty tz tn | wx wy wz (or the AMARU_LOAD_* code); ``vals`` a scalar
        or one value per integration point."""
        assert F.dtype == np.float64 and F.flags.c_contiguous
        k = LOAD_KEYS.get(key, key) if isinstance(key, str) else int(key)
        if isinstance(k, str):
            k = -1                              # the library answers with the reference's "not applicable" message
        if np.ndim(vals) == 0:
            cval, vip = float(vals), None
        else:
            cval, vip = 0.0, np.ascontiguousarray(vals, dtype=np.float64)
            assert vip.size == self.nip
        self.dm._check(self.dm.lib.amaru_loadset_apply(self.h, k, cval, _d(vip), _d(F), self.dm._msg, len(self.dm._msg)))
        return F

    def close(self):
        if getattr(self, "h", None) and getattr(self.dm, "h", None):
            self.dm.lib.amaru_loadset_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_VTU_TYPES = {np.dtype(np.float64): 0, np.dtype(np.int64): 1, np.dtype(np.int32): 2, np.dtype(np.uint64): 3}


def write_vtu(filename, coords, batch_shape, batch_nelem, conn, point_data=(), cell_data=(), desc=""):
    """``amaru_write_vtu``: ASCII .vtu formatted like the reference's save_vtu (src/mesh/io.jl:167-276).
    ``point_data`` / ``cell_data``: iterables of (name, array) with one row per node / cell."""
    lib = load()
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    bshape = np.ascontiguousarray(batch_shape, dtype=np.int32)
    bnelem = np.ascontiguousarray(batch_nelem, dtype=np.int64)
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1)
    keep = [coords, bshape, bnelem, conn]

    def pack(items, nrows):
        items = list(items)
        n = len(items)
        names = (C.c_char_p * max(n, 1))()
        types = np.zeros(max(n, 1), dtype=np.int32)
        ncomp = np.zeros(max(n, 1), dtype=np.int32)
        ptrs = (_vp * max(n, 1))()
        for i, (name, arr) in enumerate(items):
            arr = np.asarray(arr)
            if arr.dtype not in _VTU_TYPES:
                arr = arr.astype(np.float64 if arr.dtype.kind == "f" else np.int64)
            arr = np.ascontiguousarray(arr)
            if arr.shape[0] != nrows:
                raise AmaruError(f"write_vtu: array {name} has {arr.shape[0]} rows, expected {nrows}")
            keep.append(arr)
            names[i] = name.encode("utf-8")
            types[i] = _VTU_TYPES[arr.dtype]
            ncomp[i] = 1 if arr.ndim == 1 else arr.shape[1]
            ptrs[i] = arr.ctypes.data
        keep.extend([types, ncomp])
        return n, names, types.ctypes.data_as(_i32p), ncomp.ctypes.data_as(_i32p), ptrs

    npt, pn, pt, pc, pp = pack(point_data, coords.shape[0])
    ncl, cn, ct, cc, cp = pack(cell_data, int(bnelem.sum()))
    msg = C.create_string_buffer(512)
    st = lib.amaru_write_vtu(str(filename).encode(), desc.encode("utf-8"), coords.shape[0], _d(coords), len(bshape),
                             bshape.ctypes.data_as(_i32p), bnelem.ctypes.data_as(_i64p), conn.ctypes.data_as(_i32p),
                             npt, pn, pt, pc, pp, ncl, cn, ct, cc, cp, msg, 512)
    if st != OK:
        raise AmaruStatus(st, msg.value.decode(errors="replace"))


def mesh_block(shape_id, c0, c1, nx, ny, nz):
    """``amaru_mesh_block``: structured box -> (coords (nnodes,3), conn (nelems,nn) int32), reference creation order."""
    lib = load()
    nn_, ne_, k_ = C.c_int64(0), C.c_int64(0), C.c_int(0)
    if lib.amaru_mesh_block_sizes(int(shape_id), int(nx), int(ny), int(nz), C.byref(nn_), C.byref(ne_), C.byref(k_)) != OK:
        raise AmaruError("block: cannot discretize using this shape / these divisions")
    coords = np.empty((nn_.value, 3))
    conn = np.empty((ne_.value, k_.value), dtype=np.int32)
    box = np.array([c0[0], c0[1], c0[2], c1[0], c1[1], c1[2]], dtype=np.float64)
    msg = C.create_string_buffer(256)
    st = lib.amaru_mesh_block(int(shape_id), _d(box), int(nx), int(ny), int(nz), _d(coords), conn.ctypes.data_as(_i32p), msg, 256)
    if st != OK:
        raise AmaruStatus(st, msg.value.decode(errors="replace"))
    return coords, conn


def partition_elements(flat: dict, nparts: int, partitioner="rcb"):
    """The element partition ``amaru_create(ngpus=nparts)`` uses (host only): part index per element, ABI element order."""
    lib = load()
    coords = np.ascontiguousarray(flat["coords"], dtype=np.float64)
    bshape = np.ascontiguousarray(flat["batch_shape"], dtype=np.int32)
    bnelem = np.ascontiguousarray(flat["batch_nelem"], dtype=np.int64)
    conn = np.ascontiguousarray(flat["conn"], dtype=np.int32).reshape(-1)
    out = np.zeros(int(bnelem.sum()), dtype=np.int32)
    msg = C.create_string_buffer(256)
    part = PARTITIONER[partitioner] if isinstance(partitioner, str) else int(partitioner)
    st = lib.amaru_partition_elements_abi(part, int(nparts), coords.shape[0], _d(coords), len(bshape),
                                          bshape.ctypes.data_as(_i32p), bnelem.ctypes.data_as(_i64p),
                                          conn.ctypes.data_as(_i32p), out.ctypes.data_as(_i32p), msg, 256)
    if st != OK:
        raise AmaruStatus(st, msg.value.decode(errors="replace"))
    return out


def configure_dofs(prescribed):
    """``amaru_configure_dofs``: (nnodes, nd) bool -> (eqid (nnodes, nd) int32, nu)."""
    p = np.ascontiguousarray(prescribed, dtype=np.uint8)
    eq = np.empty(p.shape, dtype=np.int32)
    nu = C.c_int64(0)
    if load().amaru_configure_dofs(p.shape[0], p.shape[1], p.ctypes.data_as(_u8p), eq.ctypes.data_as(_i32p), C.byref(nu)) != OK:
        raise AmaruError("configure_dofs!: bad arguments")
    return eq, int(nu.value)


def outer_facets(shape_id, conn):
    """``amaru_outer_facets``: boundary facets of a one-shape mesh -> (facet_nodes (nf, nfn) int32, owner (nf,) int64)."""
    lib = load()
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    nfn = C.c_int(0)
    n = lib.amaru_outer_facets(int(shape_id), conn.shape[0], conn.ctypes.data_as(_i32p), None, None, 0, C.byref(nfn))
    if n < 0:
        raise AmaruError("get_outer_facets: unsupported cell shape")
    fn = np.empty((n, nfn.value), dtype=np.int32)
    ow = np.empty(n, dtype=np.int64)
    n2 = lib.amaru_outer_facets(int(shape_id), conn.shape[0], conn.ctypes.data_as(_i32p), fn.ctypes.data_as(_i32p),
                                ow.ctypes.data_as(_i64p), n, C.byref(nfn))
    assert n2 == n
    return fn, ow
