"""
The pytrees that cross the model boundary, named as in the reference
(states.py:25-50, 132-154, 202-218).  Fields are torch tensors (float64) here.
"""

from typing import Any, NamedTuple

__all__ = ["EquilibriumState", "LoadState", "EquilibriumParametersState"]


class EquilibriumState(NamedTuple):
    xyz: Any         # (nodes, 3)
    residuals: Any   # (nodes, 3)
    lengths: Any     # (edges, 1)
    forces: Any      # (edges, 1)
    loads: Any       # (nodes, 3)
    vectors: Any     # (edges, 3)


class LoadState(NamedTuple):
    nodes: Any           # (nodes, 3)
    edges: Any = 0.0     # (edges, 3) or 0.0
    faces: Any = 0.0     # (faces, 3) or 0.0


class EquilibriumParametersState(NamedTuple):
    q: Any            # (edges,)
    xyz_fixed: Any    # (nodes_fixed, 3)
    loads: LoadState
