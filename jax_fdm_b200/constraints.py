"""
Constraints on the equilibrium state for the constrained optimisers -- mirror of `jax_fdm.constraints` for the
constraint kinds that read one element (or every edge) of the state:

    EdgeLengthConstraint (constraints/edge/length.py:12-40)        lengths[e, 0]
    EdgeForceConstraint (constraints/edge/force.py:12-40)          forces[e, 0]
    Node{X,Y,Z}CoordinateConstraint (constraints/node/coordinates.py:16-106)   xyz[i, 0 | 1 | 2]
    NetworkEdgesLengthConstraint (constraints/network/length.py:13-41)         ravel(lengths)
    NetworkEdgesForceConstraint (constraints/network/force.py:13-41)           ravel(forces)

A constraint is `bound_low <= g(state) <= bound_up`.  The reference evaluates g under `jit` and its Jacobian with
`jacfwd` on the DENSE model (optimizers/constrained.py:38-102, fdm.py:266-271: the sparse solve has no forward
rule).  Here the Jacobian comes from the sparse path's own forward mode (`EquilibriumModel.jvp`): one tangent per
optimisation parameter, all of them right-hand sides of the factor of the primal solve, in one batched launch.
"""
import numpy as np

__all__ = ["Constraint", "EdgeLengthConstraint", "EdgeForceConstraint", "NodeXCoordinateConstraint",
           "NodeYCoordinateConstraint", "NodeZCoordinateConstraint", "NetworkEdgesLengthConstraint",
           "NetworkEdgesForceConstraint"]


class Constraint:
    """constraints/constraint.py: a key, a lower and an upper bound (None = unbounded)."""
    field, component = None, 0

    def __init__(self, key, bound_low=None, bound_up=None):
        self.key = key
        self.bound_low = -np.inf if bound_low is None else bound_low
        self.bound_up = np.inf if bound_up is None else bound_up

    def rows(self, structure):
        """Flat indices into `state.<field>.reshape(-1)` that this constraint reads."""
        raise NotImplementedError


class _EdgeConstraint(Constraint):
    def rows(self, structure):
        return np.array([structure.edge_index[tuple(int(k) for k in self.key)]], dtype=np.int64)


class EdgeLengthConstraint(_EdgeConstraint):
    field = "lengths"


class EdgeForceConstraint(_EdgeConstraint):
    field = "forces"


class _NodeCoordinateConstraint(Constraint):
    field = "xyz"

    def rows(self, structure):
        return np.array([3 * structure.node_index[int(self.key)] + self.component], dtype=np.int64)


class NodeXCoordinateConstraint(_NodeCoordinateConstraint):
    component = 0


class NodeYCoordinateConstraint(_NodeCoordinateConstraint):
    component = 1


class NodeZCoordinateConstraint(_NodeCoordinateConstraint):
    component = 2


class _NetworkConstraint(Constraint):
    def __init__(self, bound_low=None, bound_up=None):
        super().__init__(-1, bound_low, bound_up)

    def rows(self, structure):
        return np.arange(structure.num_edges, dtype=np.int64)


class NetworkEdgesLengthConstraint(_NetworkConstraint):
    field = "lengths"


class NetworkEdgesForceConstraint(_NetworkConstraint):
    field = "forces"
