"""
The pytrees that cross the model boundary, named as in the reference
(states.py:25-50, 132-154, 202-218).  Fields are torch tensors (float64) here.
"""

from typing import Any, NamedTuple

import numpy as np
import torch


def _is_mesh(datastructure):
    """The reference dispatches on isinstance(FDNetwork | FDMesh) (COMPAS classes, absent here); any object with the
    same accessors is accepted: meshes are told apart by their vertex accessors."""
    if hasattr(datastructure, "vertices_coordinates"):
        return True
    if hasattr(datastructure, "nodes_coordinates"):
        return False
    raise ValueError(f"Input datastructure {datastructure} is invalid")


def _t(a, device, shape=None):
    t = torch.as_tensor(np.asarray(a, dtype=np.float64), device=device)
    return t if shape is None else t.reshape(shape)

__all__ = ["DatastructureState", "EquilibriumState", "LoadState", "EquilibriumParametersState"]


class EquilibriumState(NamedTuple):
    xyz: Any         # (nodes, 3)
    residuals: Any   # (nodes, 3)
    lengths: Any     # (edges, 1)
    forces: Any      # (edges, 1)
    loads: Any       # (nodes, 3)
    vectors: Any     # (edges, 3)

    @classmethod
    def from_datastructure(cls, datastructure, structure, dtype=None, device=None):
        """states.py:52-104: the state as STORED on a network / mesh (no solve); vectors = C xyz."""
        mesh = _is_mesh(datastructure)
        xyz = datastructure.vertices_coordinates() if mesh else datastructure.nodes_coordinates()
        residuals = datastructure.vertices_residual() if mesh else datastructure.nodes_residual()
        loads = datastructure.vertices_loads() if mesh else datastructure.nodes_loads()
        xyz_np = np.asarray(xyz, dtype=np.float64)
        ei = structure.edges_indexed
        return cls(xyz=_t(xyz_np, device), residuals=_t(residuals, device),
                   lengths=_t(datastructure.edges_lengths(), device, (-1, 1)),
                   forces=_t(datastructure.edges_forces(), device, (-1, 1)), loads=_t(loads, device),
                   vectors=_t(xyz_np[ei[:, 1]] - xyz_np[ei[:, 0]], device))


class LoadState(NamedTuple):
    nodes: Any           # (nodes, 3)
    edges: Any = 0.0     # (edges, 3) or 0.0
    faces: Any = 0.0     # (faces, 3) or 0.0

    @classmethod
    def from_datastructure(cls, datastructure, dtype=None, device=None):
        """states.py:155-200: all-zero edge / face loads collapse to the scalar 0.0."""
        mesh = _is_mesh(datastructure)
        edges = np.asarray(datastructure.edges_loads(), dtype=np.float64)
        loads_edges = 0.0 if np.allclose(edges, 0.0) else _t(edges, device)
        loads_faces = 0.0
        if mesh:
            faces = np.asarray(datastructure.faces_loads(), dtype=np.float64)
            loads_faces = 0.0 if np.allclose(faces, 0.0) else _t(faces, device)
        nodes = datastructure.vertices_loads() if mesh else datastructure.nodes_loads()
        return cls(nodes=_t(nodes, device), edges=loads_edges, faces=loads_faces)


class EquilibriumParametersState(NamedTuple):
    q: Any            # (edges,)
    xyz_fixed: Any    # (nodes_fixed, 3)
    loads: LoadState

    @classmethod
    def from_datastructure(cls, datastructure, dtype=None, device=None):
        """states.py:219-262."""
        mesh = _is_mesh(datastructure)
        xyz_fixed = datastructure.vertices_fixedcoordinates() if mesh else datastructure.nodes_fixedcoordinates()
        return cls(q=_t(datastructure.edges_forcedensities(), device), xyz_fixed=_t(xyz_fixed, device, (-1, 3)),
                   loads=LoadState.from_datastructure(datastructure, device=device))


class DatastructureState(NamedTuple):
    """states.py: what `datastructure_state` bundles for a goal or constraint read without a solve."""
    eq_state: Any
    structure: Any
    parameters: Any
