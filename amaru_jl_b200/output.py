"""Output side of the mechanical path: ``update_output_data!`` and ``save(model, "file.vtu")`` of the reference fed from
the device handle instead of the Node / Ip object graph.

  update_output_data!   src/fe-model.jl:342-431   node_data / elem_data dictionaries
  nodal_patch_recovery  src/fe-model.jl:506-692   -> amaru_recovery_create / amaru_recover_nodal (csrc/recovery.cu)
  save_vtu              src/mesh/io.jl:167-276     -> amaru_write_vtu (csrc/vtu.cpp)

The recovered fields need the integration-point state on the device, so ``update_output_data`` takes the stage's
``lib.DeviceModel`` (the solver calls it at output increments); there is no CPU fallback for the recovery.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import lib as L
from .model import AmaruError, FEModel

_DOF_FIELDS = (("ux", "U", 0), ("fx", "F", 0), ("vx", "V", 0), ("ax", "A", 0),
               ("uy", "U", 1), ("fy", "F", 1), ("vy", "V", 1), ("ay", "A", 1),
               ("uz", "U", 2), ("fz", "F", 2), ("vz", "V", 2), ("az", "A", 2))


def boundary_nodes(model: FEModel) -> np.ndarray:
    """at_bound of nodal_patch_recovery: nodes of model.faces = the outer facets (fe-model.jl:523-529)."""
    fn, _ = model.mesh.outer_facets()
    flag = np.zeros(model.nnodes, dtype=np.uint8)
    flag[np.unique(fn)] = 1
    return flag


def _encode_tag(s: str) -> int:
    """encode_string_to_uint64 (src/tools/encode.jl:3-11)."""
    x = 0
    for i, b in enumerate(s.encode("utf-8")[:8]):
        x |= b << (8 * i)
    return x


def update_output_data(model: FEModel, dm: "L.DeviceModel"):
    """Fills ``model.node_data`` / ``model.elem_data`` (ordered like the reference's OrderedDicts)."""
    if dm is None:
        raise AmaruError("update_output_data!: the nodal recovery runs on the device handle of the stage (no CPU fallback)")
    nd = model.ndim
    node_data, elem_data = OrderedDict(), OrderedDict()
    node_data["node-id"] = np.arange(1, model.nnodes + 1, dtype=np.int64)
    elem_data["elem-id"] = np.arange(1, model.nelems + 1, dtype=np.int64)
    elem_data["cell-type"] = np.full(model.nelems, model.shape.vtk_type, dtype=np.int64)
    for name, src, d in _DOF_FIELDS:                              # keys(dof.vals) node by node, dof by dof (:356-374)
        arr = getattr(model, src, None)
        if arr is None or d >= nd:
            continue
        node_data[name] = np.ascontiguousarray(arr[:, d])
    if not getattr(dm, "_recovery_ready", False):
        dm.recovery_create(boundary_nodes(model))
        dm._recovery_ready = True
    names = dm.recovery_fields()
    V = dm.recover_nodal(model.nnodes)                            # nodal_patch_recovery (:377-381)
    for i, name in enumerate(names):
        node_data[name] = np.ascontiguousarray(V[:, i])
    z = np.zeros(model.nnodes)
    node_data["U"] = np.column_stack([model.U[:, 0], model.U[:, 1], model.U[:, 2] if nd == 3 else z])   # :391-399
    if getattr(model, "V", None) is not None:
        node_data["V"] = np.column_stack([model.V[:, 0], model.V[:, 1], model.V[:, 2] if nd == 3 else z])
    model.node_data, model.elem_data = node_data, elem_data
    return node_data, elem_data


def save(model: FEModel, filename: str, desc: str = ""):
    """save(model, filename) for .vtu (src/mesh/io.jl:167-276, uncompressed ASCII)."""
    if not str(filename).endswith(".vtu"):
        raise AmaruError("save: only the .vtu format is written by the B200 path")
    if not hasattr(model, "node_data"):
        raise AmaruError("save: call update_output_data(model, dm) first")
    elem_data = OrderedDict(model.elem_data)
    tags = list(dict.fromkeys(model.mesh.tags[t] for t in model.mesh.elem_tag))
    if len(tags) > 1 or tags[0] != "":                            # io.jl:169-193
        for t in tags:
            if len(t.encode("utf-8")) > 16:
                raise AmaruError(f"Mesh: tag '{t}' too long. Max length is 16 UTF units.")
        tid = {t: i for i, t in enumerate(tags)}
        per = np.array([tid[model.mesh.tags[t]] for t in model.mesh.elem_tag], dtype=np.int64)
        s1 = np.array([_encode_tag(t.encode("utf-8")[:8].decode("utf-8", "ignore")) for t in tags], dtype=np.uint64)
        s2 = np.array([_encode_tag(t.encode("utf-8")[8:].decode("utf-8", "ignore")) for t in tags], dtype=np.uint64)
        elem_data["tag-s1"] = s1[per]
        elem_data["tag-s2"] = s2[per]
        elem_data["tag"] = per
    flat = model.flatten()
    L.write_vtu(filename, model.coords, flat["batch_shape"], flat["batch_nelem"], flat["conn"],
                point_data=model.node_data.items(), cell_data=elem_data.items(), desc=desc)
