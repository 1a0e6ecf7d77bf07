"""Host-side mirror of the reference's model / stage API for the mechanical path (flat arrays, no object graph).

Names and argument meaning follow the reference:
  MechContext            src/mech/mech-solver.jl:15-33
  MechSolid              src/mech/elem/mech-solid.jl:31-46 (rho, gamma properties :13-28)
  LinearElastic          src/mech/mat/linear-elastic.jl:5-20
  VonMises               src/mech/mat/von-mises.jl:4-26
  DruckerPrager          src/mech/mat/drucker-prager.jl:5-34
  FEModel(mesh, matbinds, ctx; thickness)   src/fe-model.jl:66-255
  NodeBC / SurfaceBC / BodyC                src/bc.jl:10-58,62-136,142-194
  configure_dofs         src/bc.jl:198-233   (unknown dofs first, then prescribed)
  get_bc_vals            src/bc.jl:237-249 with src/mech/elem/distributed.jl:76-217
  MechAnalysis, addstage src/mech/mech-solver.jl:43-74, src/analysis.jl:27-31

Where the reference keeps Node/Dof/Ip objects this keeps arrays:
  model.coords (nnodes,3)   model.conn (nelem,nn)   model.elem_mat (nelem,)
  model.U (nnodes,nd) = dof.vals[:ux..]   model.F (nnodes,nd) = dof.vals[:fx..]
  model.state = dict(sigma (nip,6), eps (nip,6), epa (nip,), dlam (nip,)) element-major then IP order.
"""
from __future__ import annotations

import numpy as np

from . import shapes as S
from .expr import evaluate, select
from .mesh import Mesh

ESSENTIAL = ("ux", "uy", "uz")
NATURAL = ("fx", "fy", "fz")

MAT_LINEAR_ELASTIC, MAT_VON_MISES, MAT_DRUCKER_PRAGER = 1, 2, 3
NPARAMS = 8


class AmaruError(Exception):
    pass


# ------------------------------------------------------------------------------------------ context / materials
class MechContext:
    def __init__(self, ndim=0, stressmodel="d3", g=0.0):
        stressmodel = str(stressmodel).lstrip(":")
        if stressmodel not in ("planestress", "planestrain", "axisymmetric", "d3", "none"):
            raise AmaruError(f"MechContext: invalid stressmodel {stressmodel}")
        if g < 0:
            raise AmaruError("MechContext: g>=0 required")
        self.ndim = ndim
        self.stressmodel = stressmodel
        self.g = g
        self.thickness = 1.0


class MechSolid:
    """Element type marker; instances hold the element properties rho / gamma."""

    def __init__(self, rho=0.0, gamma=0.0):
        if rho < 0 or gamma < 0:
            raise AmaruError("MechSolid: rho>=0 and gamma>=0 required")
        self.rho, self.gamma = float(rho), float(gamma)


class Material:
    kind = 0

    def params(self):
        raise NotImplementedError


class LinearElastic(Material):
    kind = MAT_LINEAR_ELASTIC

    def __init__(self, E=None, nu=0.0):
        if E is None or not E > 0.0:
            raise AmaruError("LinearElastic: E>0.0 required")
        if not 0.0 <= nu < 0.5:
            raise AmaruError("LinearElastic: 0.0<=nu<0.5 required")
        self.E, self.nu = float(E), float(nu)

    def params(self):
        return [self.E, self.nu, 0, 0, 0, 0, 0, 0]


class VonMises(Material):
    kind = MAT_VON_MISES

    def __init__(self, E=None, nu=0.0, fy=None, H=0.0, rho=0.0):
        if E is None or not E > 0.0:
            raise AmaruError("VonMises: E>0.0 required")
        if not 0.0 <= nu < 0.5:
            raise AmaruError("VonMises: 0.0<=nu<0.5 required")
        if fy is None or not fy >= 0.0:
            raise AmaruError("VonMises: fy>=0.0 required")
        if not H >= 0.0:
            raise AmaruError("VonMises: H>=0.0 required")
        self.E, self.nu, self.fy, self.H, self.rho = float(E), float(nu), float(fy), float(H), float(rho)

    def params(self):
        return [self.E, self.nu, self.fy, self.H, 0, self.rho, 0, 0]


class DruckerPrager(Material):
    kind = MAT_DRUCKER_PRAGER

    def __init__(self, E=float("nan"), nu=0.0, alpha=0.0, kappa=0.0, H=0.0, rho=0.0):
        if not (E > 0.0 and 0.0 <= nu < 0.5 and alpha >= 0.0 and kappa > 0.0 and H >= 0.0 and rho >= 0.0):
            raise AmaruError("DruckerPrager: invalid parameters")
        self.E, self.nu, self.alpha, self.kappa, self.H, self.rho = map(float, (E, nu, alpha, kappa, H, rho))

    def params(self):
        return [self.E, self.nu, self.alpha, self.kappa, self.H, self.rho, 0, 0]


# ------------------------------------------------------------------------------------------ boundary conditions
class BC:
    def __init__(self, **conds):
        if not conds:
            raise AmaruError(f"{type(self).__name__} must have at least one condition")
        self.conds = conds


class NodeBC(BC):
    pass


class SurfaceBC(BC):
    pass


FaceBC = SurfaceBC


class BodyC(BC):
    pass


ElemBC = BodyC


# ------------------------------------------------------------------------------------------ model
class FEModel:
    """FEModel(mesh, [filter => MechSolid => Material => (params)], ctx; thickness=1.0)  (fe-model.jl:66-255).

    matbinds: list of tuples ``(filter, MechSolid | MechSolid(...), MaterialClass | material, params_dict)``;
    ``filter`` is a block tag, ``"solids"``/``"bulks"``/``"all"`` or a coordinate expression (all nodes of a cell
    must satisfy it, src/mesh/cell.jl:249-265).
    """

    def __init__(self, mesh: Mesh, matbinds, ctx: MechContext | None = None, thickness=1.0):
        ctx = ctx or MechContext()
        if not thickness > 0:
            raise AmaruError("FEModel: thickness>0 required")
        self.mesh = mesh
        self.ctx = ctx
        ctx.ndim = mesh.ndim                                         # fe-model.jl:94
        if ctx.ndim == 3 and ctx.stressmodel not in ("d3", "none"):
            raise AmaruError("FEModel: 3D models need stressmodel d3")
        if ctx.stressmodel == "axisymmetric" and ctx.ndim != 2:
            raise AmaruError("FEModel: axisymmetric models are 2D")
        self.thickness = float(thickness)
        self.ndim = ctx.ndim
        self.shape = mesh.shape
        self.coords = mesh.coords
        self.conn = mesh.conn
        self.nnodes, self.nelems = mesh.nnodes, mesh.nelems
        self.materials: list[Material] = []
        self.elem_mat = np.full(self.nelems, -1, dtype=np.int32)
        self.elem_rho = np.zeros(self.nelems)
        for bind in matbinds:
            flt, ety, mty = bind[0], bind[1], bind[2]
            par = bind[3] if len(bind) > 3 else {}
            if not (ety is MechSolid or isinstance(ety, MechSolid)):
                raise AmaruError("only MechSolid elements are on the B200 hot path (no CPU fallback)")
            par = dict(par)
            # `rho` / `gamma` in the parameter tuple are element properties (mech-solid.jl:13-28); materials that also
            # declare rho (VonMises, DruckerPrager) receive it too, like the reference's shared kwarg list
            eprops = {k: par[k] for k in ("rho", "gamma") if k in par}
            props = ety if isinstance(ety, MechSolid) else MechSolid(**eprops)
            if not isinstance(mty, Material):
                import inspect
                accepted = inspect.signature(mty.__init__).parameters
                par = {k: v for k, v in par.items() if k in accepted}
            mat = mty if isinstance(mty, Material) else mty(**par)
            if not isinstance(mat, (LinearElastic, VonMises, DruckerPrager)):
                raise AmaruError("only LinearElastic, VonMises, DruckerPrager are on the B200 hot path")
            sel = self._select_elems(flt)
            self.materials.append(mat)
            self.elem_mat[sel] = len(self.materials) - 1
            self.elem_rho[sel] = props.rho
        if (self.elem_mat < 0).any():
            raise AmaruError("FEModel: some cells have no element/material binding")
        if ctx.stressmodel == "planestress" and (self.ndim != 2 or not all(isinstance(m, LinearElastic) for m in self.materials)):
            raise AmaruError("stressmodel planestress is on the B200 hot path for 2D LinearElastic models only (no CPU fallback)")
        self.nip = self.shape.quadrature.shape[0]
        self.nip_total = self.nelems * self.nip
        nd = self.ndim
        self.U = np.zeros((self.nnodes, nd))
        self.F = np.zeros((self.nnodes, nd))
        self._state = None                                           # allocated on first use (3.6 GB at 8 M HEX20 elements)
        self.ndofs = self.nnodes * nd
        self._ipcoords = None

    @property
    def state(self):
        """ip.state of every IP, element-major: zero-initialised like the IpState constructors (von-mises.jl:35-42)."""
        if self._state is None:
            self._state = dict(sigma=np.zeros((self.nip_total, 6)), eps=np.zeros((self.nip_total, 6)),
                               epa=np.zeros(self.nip_total), dlam=np.zeros(self.nip_total))
        return self._state

    @state.setter
    def state(self, value):
        self._state = value

    def _select_elems(self, flt):
        if isinstance(flt, str):
            key = flt.lstrip(":")
            if key in self.mesh.tags and key != "":
                return self.mesh.elem_tag == self.mesh.tags.index(key)
            if key in ("solids", "bulks", "all"):
                return np.ones(self.nelems, dtype=bool)
        nodesel = select(flt, self.coords)
        return nodesel[self.conn].all(axis=1)

    # ip.coord = C'N (element.jl:160-164)
    def ip_coords(self):
        if self._ipcoords is None:
            Nt = S.func_table(self.shape)                             # (nip, nn)
            self._ipcoords = np.einsum("qa,ead->eqd", Nt, self.coords[self.conn]).reshape(-1, 3)
        return self._ipcoords

    def select_nodes(self, flt):
        return np.nonzero(select(flt, self.coords))[0]

    # ------------------------------------------------------------------ configure_dofs! (bc.jl:198-233)
    def configure_dofs(self, bcs, native=True):
        """``native=False`` numbers the dofs in numpy instead of through ``amaru_configure_dofs`` (the CPU reference arm of
        bench.py must not load the product library)."""
        nd = self.ndim
        presc = np.zeros((self.nnodes, nd), dtype=bool)
        setup = []
        for flt, bc in bcs:
            if isinstance(bc, NodeBC):
                nodes = self.select_nodes(flt)
                setup.append((bc, nodes))
                for key in bc.conds:
                    if key in ESSENTIAL[:nd]:
                        presc[nodes, ESSENTIAL.index(key)] = True
            elif isinstance(bc, SurfaceBC):
                fn, owner = self.mesh.outer_facets()
                sel = select(flt, self.coords)[fn].all(axis=1)         # all facet nodes satisfy the filter
                setup.append((bc, (fn[sel], owner[sel])))
                for key in bc.conds:
                    if key in ESSENTIAL[:nd]:
                        presc[np.unique(fn[sel]), ESSENTIAL.index(key)] = True
            elif isinstance(bc, BodyC):
                sel = self._select_elems(flt)
                setup.append((bc, np.nonzero(sel)[0]))
                for key in bc.conds:
                    if key in ESSENTIAL[:nd]:
                        presc[np.unique(self.conn[sel]), ESSENTIAL.index(key)] = True
            else:
                raise AmaruError(f"unsupported boundary condition {type(bc).__name__}")
        if native:
            from . import lib as L                                    # node-major, ux,uy,uz per node; stable split
            eq, nu = L.configure_dofs(presc)                          # (bc.jl:220-224) in C++ behind the ABI
        else:
            flat = presc.reshape(-1)
            nu = int((~flat).sum())
            eq = np.empty(flat.size, dtype=np.int32)
            eq[~flat] = np.arange(nu, dtype=np.int32)
            eq[flat] = nu + np.arange(flat.size - nu, dtype=np.int32)
            eq = eq.reshape(presc.shape)
        return eq, nu, setup

    # ------------------------------------------------------------------ get_bc_vals (bc.jl:237-249)
    def get_bc_vals(self, eqid, setup, t=0.0, device=None):
        """``device``: a lib.DeviceModel of this stage -> the distributed loads (SurfaceBC tx/ty/tz/tn, BodyC wx/wy/wz) are
        integrated on the GPU (amaru_loadset_*, csrc/loads.cu); the value expressions are evaluated here, on the
        integration-point coordinates the library returns.  Without it the numpy path below is used (multi-GPU runs)."""
        nd = self.ndim
        ndofs = eqid.size
        U = np.zeros(ndofs)
        F = np.zeros(ndofs)
        X = self.coords
        if self.ctx.stressmodel == "axisymmetric":
            device = None          # th = 2*pi*r per integration point: integrated by the numpy path (loads.cu refuses it)
        for bc, target in setup:
            if isinstance(bc, NodeBC):                                # bc.jl:43-58
                nodes = target
                x, y, z = X[nodes, 0], X[nodes, 1], X[nodes, 2]
                for key, cond in bc.conds.items():
                    if key in ESSENTIAL[:nd]:
                        U[eqid[nodes, ESSENTIAL.index(key)]] = evaluate(cond, x=x, y=y, z=z, t=t)
                    elif key in NATURAL[:nd]:
                        np.add.at(F, eqid[nodes, NATURAL.index(key)],
                                  np.broadcast_to(evaluate(cond, x=x, y=y, z=z, t=t), nodes.shape))
                    # keys that are not dofs of the node are skipped (bc.jl:50)
            elif isinstance(bc, SurfaceBC):                           # bc.jl:116-136
                fn, owner = target
                for key, val in bc.conds.items():
                    if key in ESSENTIAL[:nd]:
                        n = fn.reshape(-1)
                        U[eqid[n, ESSENTIAL.index(key)]] = np.broadcast_to(
                            evaluate(val, x=X[n, 0], y=X[n, 1], z=X[n, 2], t=t), n.shape)
                    elif device is not None:
                        self._device_load(device, ("S", id(bc)), self.shape.facet_shape.id, fn, key, val, t, F)
                    else:
                        Fd = self._boundary_forces(fn, key, val, t)   # (nf, nfn, nd)
                        np.add.at(F, eqid[fn].reshape(-1), Fd.reshape(-1))
            elif isinstance(bc, BodyC):                               # bc.jl:175-194
                elems = target
                for key, val in bc.conds.items():
                    if key in ESSENTIAL[:nd]:
                        n = self.conn[elems].reshape(-1)
                        U[eqid[n, ESSENTIAL.index(key)]] = np.broadcast_to(
                            evaluate(val, x=X[n, 0], y=X[n, 1], z=X[n, 2], t=t), n.shape)
                    elif device is not None:
                        self._device_load(device, ("B", id(bc)), self.shape.id, self.conn[elems], key, val, None, F)
                    else:
                        Fd = self._body_forces(elems, key, val)
                        np.add.at(F, eqid[self.conn[elems]].reshape(-1), Fd.reshape(-1))
        return U, F

    def _device_load(self, dm, ckey, shape_id, nodes, key, val, t, F):
        """One (bc, key) of bc.jl:116-136 / 175-194 through the device load set of that bc (cached on the handle)."""
        cache = dm.__dict__.setdefault("_loadsets", {})
        if ckey not in cache:
            ls = dm.loadset(shape_id, nodes)
            cache[ckey] = (ls, None)
        ls, X = cache[ckey]
        if isinstance(val, (int, float)):
            vals = float(val)
        else:
            if X is None:
                X = ls.ip_coords()
                cache[ckey] = (ls, X)
            kw = dict(x=X[:, 0], y=X[:, 1], z=X[:, 2])
            if t is not None:
                kw["t"] = t
            vals = np.ascontiguousarray(np.broadcast_to(evaluate(val, **kw), (X.shape[0],)), dtype=np.float64)
        ls.apply(key, vals, F)

    # mech_boundary_forces (src/mech/elem/distributed.jl:76-152), vectorised over facets
    def _boundary_forces(self, fn, key, val, t):
        nd = self.ndim
        keys = ("tx", "ty", "tn") if nd == 2 else ("tx", "ty", "tz", "tn")
        if key not in keys:
            raise AmaruError(f"mech_boundary_forces: boundary condition {key} is not applicable as distributed bc. "
                             f"Suitable keys are {keys}")
        fshape = self.shape.facet_shape
        th = self.thickness
        C = self.coords[fn][:, :, :nd]                                # (nf, nfn, nd)
        Fd = np.zeros(C.shape)
        for q in fshape.quadrature:
            N = fshape.func(q[:3])                                    # (nfn,)
            D = fshape.deriv(q[:3])                                   # (nfn, fdim)
            J = np.einsum("fai,aj->fij", C, D)                        # C'*D  (nf, nd, fdim)
            Xq = np.einsum("fai,a->fi", C, N)
            x, y = Xq[:, 0], Xq[:, 1]
            z = Xq[:, 2] if nd == 3 else np.zeros_like(x)
            vip = np.broadcast_to(evaluate(val, t=t, x=x, y=y, z=z), x.shape)
            Q = np.zeros((C.shape[0], nd))
            if nd == 2:
                nrm = np.stack((J[:, 1, 0], -J[:, 0, 0]), axis=1)
                nJ = np.sqrt((J[:, :, 0] ** 2).sum(axis=1))
            else:
                nrm = np.cross(J[:, :, 0], J[:, :, 1])
                nJ = np.sqrt((nrm ** 2).sum(axis=1))                  # norm2 of a 3x2 J (tools/linalg.jl:64-69)
            if key == "tn":
                Q = vip[:, None] * nrm / np.sqrt((nrm ** 2).sum(axis=1))[:, None]
            else:
                Q[:, ("tx", "ty", "tz").index(key)] = vip
            tq = 2 * np.pi * x if self.ctx.stressmodel == "axisymmetric" else th      # distributed.jl:121
            coef = nJ * q[3] * tq
            Fd += coef[:, None, None] * N[None, :, None] * Q[:, None, :]
        return Fd

    # mech_solid_body_forces (distributed.jl:157-217)
    def _body_forces(self, elems, key, val):
        nd = self.ndim
        if key not in ("wx", "wy", "wz") or (key == "wz" and nd == 2):
            raise AmaruError(f"mech_solid_body_forces: condition {key} is not applicable")
        sh = self.shape
        C = self.coords[self.conn[elems]][:, :, :nd]
        Fd = np.zeros(C.shape)
        for q in sh.quadrature:
            N = sh.func(q[:3])
            D = sh.deriv(q[:3])
            J = np.einsum("eai,aj->eij", C, D)
            Xq = np.einsum("eai,a->ei", C, N)
            x, y = Xq[:, 0], Xq[:, 1]
            z = Xq[:, 2] if nd == 3 else np.zeros_like(x)
            vip = np.broadcast_to(evaluate(val, x=x, y=y, z=z), x.shape)
            Q = np.zeros((C.shape[0], nd))
            Q[:, ("wx", "wy", "wz").index(key)] = vip
            tq = 2 * np.pi * x if self.ctx.stressmodel == "axisymmetric" else self.thickness   # distributed.jl:193
            coef = np.linalg.det(J) * q[3] * tq
            Fd += coef[:, None, None] * N[None, :, None] * Q[:, None, :]
        return Fd

    # ------------------------------------------------------------------ flattening for the C ABI
    def flatten(self):
        """Arrays exactly as amaru_create takes them (include/amaru_b200.h)."""
        return dict(
            ndim=self.ndim, stressmodel={"d3": 0, "none": 0, "planestrain": 1, "planestress": 2, "axisymmetric": 3}[self.ctx.stressmodel],
            thickness=self.thickness, coords=np.ascontiguousarray(self.coords, dtype=np.float64),
            batch_shape=np.array([self.shape.id], dtype=np.int32),
            batch_nelem=np.array([self.nelems], dtype=np.int64),
            conn=np.ascontiguousarray(self.conn, dtype=np.int32),
            elem_mat=np.ascontiguousarray(self.elem_mat, dtype=np.int32),
            mat_kind=np.array([m.kind for m in self.materials], dtype=np.int32),
            mat_params=np.array([m.params() for m in self.materials], dtype=np.float64).reshape(-1, NPARAMS),
        )


# ------------------------------------------------------------------------------------------ analysis / stages
class Stage:
    def __init__(self, sid, bcs, nincs=1, nouts=0, tspan=0.0):
        self.id, self.bcs, self.nincs, self.nouts = sid, list(bcs), int(nincs), int(nouts)
        self.tspan = float(tspan)
        self.status = "pending"


class ReturnStatus:
    def __init__(self, success=True, message=""):
        self.success, self.message = success, message

    def __repr__(self):
        return f"ReturnStatus(success={self.success}, message={self.message!r})"


def success(msg=""):
    return ReturnStatus(True, msg)


def failure(msg=""):
    return ReturnStatus(False, msg)


class MechAnalysis:
    """MechAnalysis(model) (mech-solver.jl:43-74): sets ctx.thickness, default stress model."""

    def __init__(self, model: FEModel, outdir=None, outkey="out"):
        """``outdir``: where ``<outkey>-<n>.vtu`` files go at output checkpoints (analysis.jl:76-90).  The reference defaults
        to "."; here the default None writes nothing (the records stay in ``ana.records``)."""
        self.model = model
        self.ctx = model.ctx
        self.stages: list[Stage] = []
        self.outdir, self.outkey = outdir, outkey
        model.ctx.thickness = model.thickness
        if model.ctx.stressmodel == "none":
            model.ctx.stressmodel = "planestrain" if model.ctx.ndim == 2 else "d3"
        self.out = 0                            # sctx.out: number of output files written
        self.log: list[str] = []
        self.records: list[dict] = []          # one entry per converged increment (what loggers would sample)
        self.stats: list[dict] = []            # per Newton iteration: cg iterations, residue, timings


class DynamicAnalysis(MechAnalysis):
    """DynamicAnalysis(model) (src/mech/dyn-solver.jl:35-66): Newmark time integration; stages carry `tspan`."""

    def __init__(self, model: FEModel, outdir=None, outkey="out"):
        super().__init__(model, outdir, outkey)
        self.t = 0.0
        nd = model.ndim
        model.V = np.zeros((model.nnodes, nd))       # dof.vals[:vx..]
        model.A = np.zeros((model.nnodes, nd))       # dof.vals[:ax..]


DynAnalysis = DynamicAnalysis


def addstage(ana: MechAnalysis, bcs, nincs=1, nouts=0, tspan=0.0):
    """addstage!(ana, bcs; nincs, nouts, tspan) (analysis.jl:27-31)."""
    st = Stage(len(ana.stages) + 1, bcs, nincs, nouts, tspan)
    ana.stages.append(st)
    return st
