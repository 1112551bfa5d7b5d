"""
oracle/goals.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Goal predictions, error terms and the loss, restated in numpy from the reference one-liners:
  goals/node/point.py           NodePointGoal.prediction          xyz[i, :]
  goals/node/residual.py        NodeResidualForceGoal             |residuals[i]|
                                NodeResidualVectorGoal            residuals[i, :]
  goals/edge/length.py:16-25    EdgeLengthGoal                    lengths[e, 0]
  goals/edge/force.py:16-25     EdgeForceGoal                     forces[e, 0]
  goals/edge/loadpath.py        EdgeLoadPathGoal                  |forces[e, 0]| * lengths[e, 0]
  goals/network/loadpath.py     NetworkLoadPathGoal               sum |lengths * forces|
  losses/errors.py:232-431      SquaredError, AbsoluteError, PredictionError, LogMaxError (.error),
                                Mean*, RootMeanSquaredError (.errors), Error.__call__ (alpha)
  losses/loss.py:60-90          Loss.__call__  (sum of error terms + regularizers)
  losses/regularizers.py:43-70  L2Regularizer  alpha * sum(q^2)
Pinned: tests/test_oracle_goals_golden.py checks every prediction, reference value, error kernel and aggregator
here against values produced by the reference's own goal / error / geometry code (tests/golden/make_golden_goals.py).
`state` is the dict of oracle.equilibrium_state.
"""
import numpy as np

PREDICTIONS = {
    "NodePointGoal": lambda st, i: st["xyz"][i, :],
    "NodeResidualForceGoal": lambda st, i: np.linalg.norm(st["residuals"][i, :]),
    "NodeResidualVectorGoal": lambda st, i: st["residuals"][i, :],
    "EdgeLengthGoal": lambda st, i: st["lengths"][i, 0],
    "EdgeForceGoal": lambda st, i: st["forces"][i, 0],
    "EdgeLoadPathGoal": lambda st, i: np.abs(st["forces"][i, 0]) * st["lengths"][i, 0],
    "NetworkLoadPathGoal": lambda st, i: np.sum(np.abs(st["lengths"] * st["forces"])),
    "NodeXCoordinateGoal": lambda st, i: st["xyz"][i, 0],      # goals/node/coordinates.py:44, 75, 106
    "NodeYCoordinateGoal": lambda st, i: st["xyz"][i, 1],
    "NodeZCoordinateGoal": lambda st, i: st["xyz"][i, 2],
}



def _closest_point_on_line(p, line):
    """geometry.py:288-313."""
    a, b = np.asarray(line, dtype=np.float64).reshape(2, 3)
    ab = b - a
    return a + ab * ((p - a) @ ab) / (ab @ ab)


def _closest_point_on_plane(p, plane):
    """geometry.py:260-286 (normal through normalize_vector, :138-168)."""
    origin, normal = np.asarray(plane, dtype=np.float64).reshape(2, 3)
    normal = np.zeros(3) if np.all(np.abs(normal) <= 1e-8) else normal / np.linalg.norm(normal)
    e = normal @ p - normal @ origin
    return p - normal * (e / np.sum(np.square(normal)))


def _closest_point_on_segment(p, seg):
    """geometry.py:315-357."""
    a, b = np.asarray(seg, dtype=np.float64).reshape(2, 3)
    c = _closest_point_on_line(p, seg)
    d, d1, d2 = np.sum((a - b) ** 2), np.sum((a - c) ** 2), np.sum((b - c) ** 2)
    if d1 > d or d2 > d:
        return a if d1 < d2 else b
    return c


def _normalize_vector(v):
    """geometry.py:138-168."""
    v = np.asarray(v, dtype=np.float64)
    return np.zeros(3) if np.all(np.abs(v) <= 1e-8) else v / np.linalg.norm(v)


# goals whose reference value is computed from the target and the prediction (goals/node/line.py, plane.py,
# segment.py, residual.py:125-145, 197-219)
GOALS = {"NodeLineGoal": _closest_point_on_line, "NodePlaneGoal": _closest_point_on_plane,
         "NodeSegmentGoal": _closest_point_on_segment,
         "NodeResidualDirectionGoal": lambda p, t: _normalize_vector(t),
         "NodeResidualPlaneGoal": lambda p, t: p - (p @ _normalize_vector(t)) * _normalize_vector(t)}
for _n in ("NodeLineGoal", "NodePlaneGoal", "NodeSegmentGoal"):
    PREDICTIONS[_n] = PREDICTIONS["NodePointGoal"]
PREDICTIONS["NodeResidualDirectionGoal"] = lambda st, i: _normalize_vector(st["residuals"][i, :])   # residual.py:98-123
PREDICTIONS["NodeResidualPlaneGoal"] = PREDICTIONS["NodeResidualDirectionGoal"]                      # residual.py:170-195
PREDICTIONS["EdgeDirectionGoal"] = lambda st, i: _normalize_vector(st["vectors"][i, :])             # edge/direction.py:23-47
GOALS["EdgeDirectionGoal"] = lambda p, t: _normalize_vector(t)                                       # edge/direction.py:49-69


def edges_equal(st, field, idx):
    """goals/edge/length.py:49-75, force.py:49-75: var(v) / mean(v) for lengths, var(v) / mean(|v|) for forces."""
    v = st[field][np.asarray(idx), 0]
    return np.var(v) / (np.mean(v) if field == "lengths" else np.mean(np.abs(v)))


def nodes_colinear(st, idx):
    """goals/node/colinear.py:19-60 with geometry.py:660-694 (colinearity_points)."""
    P = st["xyz"][np.asarray(idx), :]
    d = P[1:] - P[:-1]
    l2 = np.sum(d * d, axis=-1)
    dt = d[1:] - d[:-1]
    lbar = 0.5 * (l2[1:] + l2[:-1])
    return np.sum(np.sum(dt ** 2, axis=-1) / lbar) / max(dt.shape[0], 1)


def edge_angle(st, i, vector):
    """goals/edge/angle.py:74-101 with geometry.py:46-99 (radians)."""
    c = _normalize_vector(st["vectors"][i, :]) @ _normalize_vector(vector)
    return np.arccos(np.clip(c, -1.0, 1.0))

ERRORS = {
    "SquaredError": lambda w, p, t: np.sum(w * np.square(p - t)),
    "AbsoluteError": lambda w, p, t: np.sum(w * np.abs(p - t)),
    "PredictionError": lambda w, p, t: np.sum(p * w),
    "LogMaxError": lambda w, p, t: np.sum(w * np.log1p(np.maximum(p - t, 0.0))),
}
BASE = {"MeanSquaredError": "SquaredError", "RootMeanSquaredError": "SquaredError",
        "MeanAbsoluteError": "AbsoluteError", "MeanPredictionError": "PredictionError"}


def error_term(name, goals, alpha, state):
    """goals: list of (goal class name, element index, target, weight)."""
    fn = ERRORS[BASE.get(name, name)]
    total = 0.0
    for g, i, t, w in goals:
        if g in ("EdgesLengthEqualGoal", "EdgesForceEqualGoal"):      # i = the list of edge indices
            total = total + fn(w, edges_equal(state, "lengths" if g == "EdgesLengthEqualGoal" else "forces", i), np.float64(0.0))
            continue
        if g == "NodesColinearGoal":                  # i = the ordered list of node indices
            total = total + fn(w, nodes_colinear(state, i), np.float64(0.0))
            continue
        if g == "EdgeAngleGoal":                      # t = (target angle, reference vector)
            total = total + fn(w, edge_angle(state, i, t[1]), np.float64(t[0]))
            continue
        p = PREDICTIONS[g](state, i)
        t = GOALS[g](p, t) if g in GOALS else np.asarray(t, dtype=np.float64)
        total = total + fn(w, p, t)
    if name.startswith("Mean") or name.startswith("RootMean"):
        total = total / len(goals)
    if name.startswith("Root"):
        total = np.sqrt(total)
    return alpha * total


def loss(terms, state, q=None, l2_alpha=0.0):
    """terms: list of (error class name, goals, alpha)."""
    value = sum(error_term(n, g, a, state) for n, g, a in terms)
    if l2_alpha:
        value = value + l2_alpha * np.sum(np.square(q))
    return value
