"""
Goals (mirror of `jax_fdm.goals` for the goals that read one element of the equilibrium state).

A goal names an element by KEY (node key, edge key pair), a target and a weight, as in the reference
(goals/goal.py:152-260); `index(structure)` resolves the key against the structure's ordering
(equilibrium/indexing.py:88-150).  `prediction(eq_state, structure)` is the reference's per-goal
prediction on torch tensors (used by tests and by `Error.evaluate_state`); the hot path does not call
it: a `Loss` compiles its goals into a table and evaluates them with one fused kernel
(csrc/goals.cu, `fdm_goals_loss`).
"""
import numpy as np

__all__ = ["Goal", "NodeGoal", "EdgeGoal", "NetworkGoal", "NodePointGoal", "NodeResidualForceGoal",
           "NodeResidualVectorGoal", "NodeResidualDirectionGoal", "NodeResidualPlaneGoal",
           "VertexResidualDirectionGoal", "VertexResidualPlaneGoal", "NodeLineGoal", "NodePlaneGoal", "NodeSegmentGoal", "VertexLineGoal",
           "VertexPlaneGoal", "VertexSegmentGoal", "NodeXCoordinateGoal", "NodeYCoordinateGoal", "NodeZCoordinateGoal",
           "EdgeLengthGoal", "EdgeForceGoal", "EdgeLoadPathGoal", "EdgeDirectionGoal", "EdgeAngleGoal",
           "EdgesLengthEqualGoal", "EdgesForceEqualGoal", "NodesColinearGoal", "VerticesColinearGoal",
           "NetworkLoadPathGoal",
           "VertexGoal", "VertexPointGoal", "VertexResidualForceGoal", "VertexResidualVectorGoal",
           "VertexXCoordinateGoal", "VertexYCoordinateGoal", "VertexZCoordinateGoal", "MeshLoadPathGoal", "GoalState"]

from collections import namedtuple

GoalState = namedtuple("GoalState", "goal weight prediction")     # goals/state.py:8-25


class Goal:
    """goals/goal.py:152-379: key, target, weight; `kind` is the row of the kernel's prediction table."""
    kind = -1
    components = 1          # of the prediction
    target_size = None      # of the target when it is not a value of the prediction's shape (line, plane: 2 x 3)

    def __init__(self, key, target, weight=1.0):
        self.key = key
        self.target = np.asarray(target, dtype=np.float64)
        self.weight = float(weight)
        want = self.target_size or self.components
        if self.target.size != want:
            raise ValueError(f"{type(self).__name__}: target must have {want} component(s)")

    def goal(self, target, prediction):
        """goals/goal.py: the reference value the prediction is compared with (the target itself by default)."""
        return target

    def index(self, structure):
        raise NotImplementedError

    def prediction(self, eq_state, structure):
        raise NotImplementedError

    def __call__(self, eq_state, structure):
        import torch
        p = self.prediction(eq_state, structure)
        t = torch.as_tensor(self.target, dtype=p.dtype, device=p.device)
        t = t.reshape(2, 3) if self.target_size == 6 else t.reshape(p.shape[-1:] if self.components == 3 else ())
        return GoalState(goal=self.goal(t, p), weight=self.weight, prediction=p)


class NodeGoal(Goal):
    """goals/node/node.py: the key is a node key."""

    def index(self, structure):
        return structure.node_index[int(self.key)]


class EdgeGoal(Goal):
    """goals/edge/edge.py: the key is the (u, v) pair of the edge as the structure lists it."""

    def index(self, structure):
        return structure.edge_index[(int(self.key[0]), int(self.key[1]))]


class NetworkGoal(Goal):
    """goals/network/network.py: one goal for the whole structure (key -1)."""

    def __init__(self, target=0.0, weight=1.0, key=-1):
        super().__init__(key, target, weight)

    def index(self, structure):
        return 0


class NodePointGoal(NodeGoal):
    """goals/node/point.py: xyz of the node."""
    kind, components = 0, 3

    def prediction(self, eq_state, structure):
        return eq_state.xyz[..., self.index(structure), :]


class NodeResidualForceGoal(NodeGoal):
    """goals/node/residual.py: magnitude of the residual force at the node."""
    kind = 1

    def prediction(self, eq_state, structure):
        return eq_state.residuals[..., self.index(structure), :].norm(dim=-1)


class NodeResidualVectorGoal(NodeGoal):
    """goals/node/residual.py: residual force vector at the node."""
    kind, components = 2, 3

    def prediction(self, eq_state, structure):
        return eq_state.residuals[..., self.index(structure), :]


class EdgeLengthGoal(EdgeGoal):
    """goals/edge/length.py:16-25."""
    kind = 3

    def prediction(self, eq_state, structure):
        return eq_state.lengths[..., self.index(structure), 0]


class EdgeForceGoal(EdgeGoal):
    """goals/edge/force.py:16-25."""
    kind = 4

    def prediction(self, eq_state, structure):
        return eq_state.forces[..., self.index(structure), 0]


class EdgeLoadPathGoal(EdgeGoal):
    """goals/edge/loadpath.py: |force| * length of the edge."""
    kind = 5

    def prediction(self, eq_state, structure):
        i = self.index(structure)
        return eq_state.forces[..., i, 0].abs() * eq_state.lengths[..., i, 0]


class NetworkLoadPathGoal(NetworkGoal):
    """goals/network/loadpath.py: sum over the edges of |force * length|."""
    kind = 6

    def prediction(self, eq_state, structure):
        return (eq_state.lengths * eq_state.forces).abs().sum(dim=(-2, -1))


class _NodeCoordinateGoal(NodeGoal):
    axis = 0

    def prediction(self, eq_state, structure):
        return eq_state.xyz[..., self.index(structure), self.axis]


class NodeXCoordinateGoal(_NodeCoordinateGoal):
    """goals/node/coordinates.py:16-44."""
    kind, axis = 7, 0


class NodeYCoordinateGoal(_NodeCoordinateGoal):
    """goals/node/coordinates.py:47-75."""
    kind, axis = 8, 1


class NodeZCoordinateGoal(_NodeCoordinateGoal):
    """goals/node/coordinates.py:78-106."""
    kind, axis = 9, 2


# Mesh flavour (goals/vertex/*.py, goals/mesh/loadpath.py): the same predictions, keys are vertex keys.
class VertexGoal(NodeGoal):
    """goals/vertex/vertex.py."""


class VertexPointGoal(VertexGoal, NodePointGoal):
    """goals/vertex/point.py:7."""


class VertexResidualForceGoal(VertexGoal, NodeResidualForceGoal):
    """goals/vertex/residual.py:17."""


class VertexResidualVectorGoal(VertexGoal, NodeResidualVectorGoal):
    """goals/vertex/residual.py:23."""


class VertexXCoordinateGoal(VertexGoal, NodeXCoordinateGoal):
    """goals/vertex/coordinates.py:13."""


class VertexYCoordinateGoal(VertexGoal, NodeYCoordinateGoal):
    """goals/vertex/coordinates.py:19."""


class VertexZCoordinateGoal(VertexGoal, NodeZCoordinateGoal):
    """goals/vertex/coordinates.py:25."""


class MeshLoadPathGoal(NetworkLoadPathGoal):
    """goals/mesh/loadpath.py:7."""


class NodeLineGoal(NodePointGoal):
    """goals/node/line.py: the node against its closest point on the line through target[0], target[1]."""
    kind, target_size = 10, 6

    def goal(self, target, prediction):                 # geometry.py:288-313
        a, b = target[0], target[1]
        ab = b - a
        s = ((prediction - a) * ab).sum(dim=-1, keepdim=True) / (ab * ab).sum()
        return a + s * ab


class NodePlaneGoal(NodePointGoal):
    """goals/node/plane.py: the node against its closest point on the plane (origin target[0], normal target[1])."""
    kind, target_size = 11, 6

    def goal(self, target, prediction):                 # geometry.py:260-286
        origin, normal = target[0], target[1]
        normal = normal / normal.norm() if bool((normal.abs() > 1e-8).any()) else normal * 0.0
        e = (prediction * normal).sum(dim=-1, keepdim=True) - (normal * origin).sum()
        return prediction - normal * (e / (normal * normal).sum())


class NodeSegmentGoal(NodeLineGoal):
    """goals/node/segment.py: the node against the closest point of the segment target[0] - target[1]."""
    kind = 12

    def goal(self, target, prediction):                 # geometry.py:315-357
        import torch
        a, b = target[0], target[1]
        p = NodeLineGoal.goal(self, target, prediction)
        d = ((a - b) ** 2).sum()
        d1 = ((a - p) ** 2).sum(dim=-1, keepdim=True)
        d2 = ((b - p) ** 2).sum(dim=-1, keepdim=True)
        end = torch.where(d1 < d2, a.expand_as(p), b.expand_as(p))
        return torch.where((d1 > d) | (d2 > d), end, p)


class VertexLineGoal(VertexGoal, NodeLineGoal):
    """goals/vertex/line.py."""


class VertexPlaneGoal(VertexGoal, NodePlaneGoal):
    """goals/vertex/plane.py."""


class VertexSegmentGoal(VertexGoal, NodeSegmentGoal):
    """goals/vertex/segment.py."""


def _unit(v):
    """normalize_vector (geometry.py:138-168): the zero vector stays zero."""
    n = v.norm(dim=-1, keepdim=True)
    zero = (v.abs() <= 1e-8).all(dim=-1, keepdim=True)
    return v / n.masked_fill(zero, 1.0) * (~zero)


class NodeResidualDirectionGoal(NodeGoal):
    """goals/node/residual.py:86-145: unit residual against the unit target."""
    kind, components = 13, 3

    def prediction(self, eq_state, structure):
        return _unit(eq_state.residuals[..., self.index(structure), :])

    def goal(self, target, prediction):
        return _unit(target)


class NodeResidualPlaneGoal(NodeGoal):
    """goals/node/residual.py:148-219: unit residual against its projection on the plane with normal `target`."""
    kind, components = 14, 3

    def prediction(self, eq_state, structure):
        return _unit(eq_state.residuals[..., self.index(structure), :])

    def goal(self, target, prediction):
        n = _unit(target)
        return prediction - (prediction * n).sum(dim=-1, keepdim=True) * n


class VertexResidualDirectionGoal(VertexGoal, NodeResidualDirectionGoal):
    """goals/vertex/residual.py:29."""


class VertexResidualPlaneGoal(VertexGoal, NodeResidualPlaneGoal):
    """goals/vertex/residual.py:40."""


class EdgeDirectionGoal(EdgeGoal):
    """goals/edge/direction.py: unit edge vector against the unit target."""
    kind, components = 15, 3

    def prediction(self, eq_state, structure):
        return _unit(eq_state.vectors[..., self.index(structure), :])

    def goal(self, target, prediction):
        return _unit(target)


class EdgeAngleGoal(EdgeGoal):
    """goals/edge/angle.py: angle (radians) between the edge vector and `vector` against the target angle."""
    kind = 16

    def __init__(self, key, target, weight=1.0, *, vector):
        self.vector = np.asarray(vector, dtype=np.float64).reshape(3)
        super().__init__(key, target, weight)
        self.target = np.concatenate([np.asarray(target, dtype=np.float64).reshape(1), self.vector])   # kernel row
        self.target_size = 4

    def prediction(self, eq_state, structure):
        import torch
        u = _unit(eq_state.vectors[..., self.index(structure), :])
        w = _unit(torch.as_tensor(self.vector, dtype=u.dtype, device=u.device))
        return torch.arccos(torch.clamp((u * w).sum(dim=-1), -1.0, 1.0))

    def __call__(self, eq_state, structure):
        import torch
        p = self.prediction(eq_state, structure)
        return GoalState(goal=torch.as_tensor(self.target[0], dtype=p.dtype, device=p.device), weight=self.weight, prediction=p)


class _EdgesAggregateGoal(EdgeGoal):
    """goals/goal.py:165-190: an aggregate goal takes the keys of the elements it reduces over; target 0."""
    is_aggregate = True

    def __init__(self, key, weight=1.0):
        super().__init__(key=key, target=0.0, weight=weight)

    def indices(self, structure):
        ei = structure.edge_index
        return [ei[(int(u), int(v))] for u, v in self.key]

    def index(self, structure):
        return 0


class EdgesLengthEqualGoal(_EdgesAggregateGoal):
    """goals/edge/length.py:49-75: var(lengths) / mean(lengths) over the edge set."""
    kind = 17

    def prediction(self, eq_state, structure):
        import torch
        idx = torch.as_tensor(self.indices(structure), device=eq_state.lengths.device)
        v = eq_state.lengths[..., 0].index_select(-1, idx)
        return v.var(dim=-1, unbiased=False) / v.mean(dim=-1)


class EdgesForceEqualGoal(_EdgesAggregateGoal):
    """goals/edge/force.py:49-75: var(forces) / mean(|forces|) over the edge set."""
    kind = 18

    def prediction(self, eq_state, structure):
        import torch
        idx = torch.as_tensor(self.indices(structure), device=eq_state.forces.device)
        v = eq_state.forces[..., 0].index_select(-1, idx)
        return v.var(dim=-1, unbiased=False) / v.abs().mean(dim=-1)


class NodesColinearGoal(NodeGoal):
    """goals/node/colinear.py:19-60: length-normalised colinearity energy of an ordered node sequence (geometry.py:660-694)."""
    kind = 19
    is_aggregate = True

    def __init__(self, key, weight=1.0):
        super().__init__(key=key, target=0.0, weight=weight)

    def indices(self, structure):
        ni = structure.node_index
        return [ni[int(k)] for k in self.key]

    def index(self, structure):
        return 0

    def prediction(self, eq_state, structure):
        import torch
        idx = torch.as_tensor(self.indices(structure), device=eq_state.xyz.device)
        P = eq_state.xyz.index_select(-2, idx)
        d = P[..., 1:, :] - P[..., :-1, :]
        l2 = (d * d).sum(dim=-1)
        dt = d[..., 1:, :] - d[..., :-1, :]
        lbar = 0.5 * (l2[..., 1:] + l2[..., :-1])
        return ((dt * dt).sum(dim=-1) / lbar).sum(dim=-1) / max(dt.shape[-2], 1)


class VerticesColinearGoal(VertexGoal, NodesColinearGoal):
    """goals/vertex/colinear.py."""
