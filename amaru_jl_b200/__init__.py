"""amaru_jl_b200 — B200-native (sm_100a) implementation of Amaru.jl's mechanical Newton-iteration hot path.

Host-side mirror of the reference API (flat arrays instead of object graphs) over the C ABI of
``libamaru_b200.so`` (``include/amaru_b200.h``).  The compute path is CUDA only; there is no CPU fallback.
"""
from .mesh import Block, Mesh
from .model import (AmaruError, BodyC, DruckerPrager, ElemBC, FaceBC, FEModel, LinearElastic, MechAnalysis,
                    MechContext, MechSolid, NodeBC, SurfaceBC, VonMises, addstage)
from .shapes import HEX8, HEX20, QUAD4, QUAD8, TET10
from .output import save, update_output_data
from .solver import solve

__all__ = ["Block", "Mesh", "FEModel", "MechContext", "MechAnalysis", "MechSolid", "LinearElastic", "VonMises",
           "DruckerPrager", "NodeBC", "SurfaceBC", "FaceBC", "BodyC", "ElemBC", "addstage", "solve", "save", "update_output_data", "AmaruError",
           "QUAD4", "QUAD8", "HEX8", "HEX20", "TET10"]
