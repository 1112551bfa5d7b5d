#!/usr/bin/env python
"""
Golden vectors for goals and error terms, made by EXECUTING THE REFERENCE'S OWN CODE:

    src/jax_fdm/goals/{node,edge,network}/*.py   (the `prediction` / `goal` methods)
    src/jax_fdm/losses/errors.py                 (the `error` kernels and the Mean / Root aggregators)
    src/jax_fdm/geometry/geometry.py             (closest points, angles, normalisation)

loaded unmodified from /root/reference under the numpy stand-in for `jax` of make_golden.py (equinox.Module is a
plain class there, so the goal classes are not instantiated: their methods are called unbound, which is all the
reference's `Goal.__call__` does with them, goals/goal.py:262-330).  The equilibrium state they read comes
from the oracle's forward solve of a small cable-net (itself pinned by make_golden.py).

Run here (needs /root/reference):   python tests/golden/make_golden_goals.py
Output: tests/golden/goals.npz
"""
import importlib
import os
import sys
import types
from collections import namedtuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import make_golden_loads as GL  # noqa: E402,F401  (vmap and the extra jnp functions of the stand-in)

GOAL_MODULES = {
    "NodePointGoal": "jax_fdm.goals.node.point", "NodeLineGoal": "jax_fdm.goals.node.line",
    "NodePlaneGoal": "jax_fdm.goals.node.plane", "NodeSegmentGoal": "jax_fdm.goals.node.segment",
    "NodeResidualForceGoal": "jax_fdm.goals.node.residual", "NodeResidualVectorGoal": "jax_fdm.goals.node.residual",
    "NodeResidualDirectionGoal": "jax_fdm.goals.node.residual", "NodeResidualPlaneGoal": "jax_fdm.goals.node.residual",
    "NodeXCoordinateGoal": "jax_fdm.goals.node.coordinates", "NodeYCoordinateGoal": "jax_fdm.goals.node.coordinates",
    "NodeZCoordinateGoal": "jax_fdm.goals.node.coordinates",
    "EdgeLengthGoal": "jax_fdm.goals.edge.length", "EdgesLengthEqualGoal": "jax_fdm.goals.edge.length",
    "EdgeForceGoal": "jax_fdm.goals.edge.force", "EdgesForceEqualGoal": "jax_fdm.goals.edge.force",
    "EdgeLoadPathGoal": "jax_fdm.goals.edge.loadpath", "EdgeDirectionGoal": "jax_fdm.goals.edge.direction",
    "EdgeAngleGoal": "jax_fdm.goals.edge.angle", "NetworkLoadPathGoal": "jax_fdm.goals.network.loadpath",
    "NodesColinearGoal": "jax_fdm.goals.node.colinear",
}
G.REAL |= set(GOAL_MODULES.values()) | {
    "jax_fdm.goals.goal", "jax_fdm.goals.state", "jax_fdm.goals.node.node", "jax_fdm.goals.edge.edge",
    "jax_fdm.goals.network.network", "jax_fdm.losses.errors", "jax_fdm.geometry", "jax_fdm.geometry.geometry"}

GoalState = namedtuple("GoalState", "goal weight prediction")

_prev_build_fake = G._build_fake


def _build_fake(name):
    m = _prev_build_fake(name)
    if name == "jax.lax":                                   # geometry.py:338-357 picks the closest point with lax.cond
        m.cond = lambda pred, true_fn, false_fn: true_fn() if bool(pred) else false_fn()
    return m


G._build_fake = _build_fake


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import oracle
    from jax_fdm_b200 import datasets

    G.load_reference()
    prob = datasets.cablenet_problem(6, seed=2)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    st, _ = oracle.forward(prob.q, prob.xyz_fixed, prob.loads, so)
    eq_state = types.SimpleNamespace(**{k: G.jarr(v) for k, v in st.items()})
    out = {"nodes": prob.nodes, "edges": prob.edges, "supports": prob.supports, "q": prob.q,
           "xyz_fixed": prob.xyz_fixed, "loads": prob.loads}
    rng = np.random.default_rng(5)
    fixed = np.flatnonzero(prob.supports)
    cases = []          # (goal name, index or index list, target, extra)
    for n in (3, 11, 20):
        cases += [("NodePointGoal", n, rng.standard_normal(3), None),
                  ("NodeLineGoal", n, rng.standard_normal((2, 3)) * 3, None),
                  ("NodePlaneGoal", n, rng.standard_normal((2, 3)), None),
                  ("NodeSegmentGoal", n, st["xyz"][n] + rng.standard_normal((2, 3)), None),
                  ("NodeSegmentGoal", n, st["xyz"][n] + 5.0 + rng.standard_normal((2, 3)), None),
                  ("NodeXCoordinateGoal", n, rng.standard_normal(), None), ("NodeYCoordinateGoal", n, 0.3, None),
                  ("NodeZCoordinateGoal", n, -0.2, None)]
    for n in fixed[:3]:
        cases += [("NodeResidualForceGoal", int(n), 0.4, None), ("NodeResidualVectorGoal", int(n), rng.standard_normal(3), None),
                  ("NodeResidualDirectionGoal", int(n), rng.standard_normal(3), None),
                  ("NodeResidualPlaneGoal", int(n), rng.standard_normal(3), None)]
    for e in (0, 7, 19):
        cases += [("EdgeLengthGoal", e, 1.1, None), ("EdgeForceGoal", e, 0.7, None), ("EdgeLoadPathGoal", e, 0.0, None),
                  ("EdgeDirectionGoal", e, rng.standard_normal(3), None),
                  ("EdgeAngleGoal", e, 0.5, rng.standard_normal(3))]
    cases += [("EdgesLengthEqualGoal", list(range(4, 15)), 0.0, None), ("EdgesForceEqualGoal", list(range(0, 30)), 0.0, None),
              ("NetworkLoadPathGoal", 0, 0.0, None), ("NodesColinearGoal", [8, 9, 10, 11, 12], 0.0, None),
              ("NodesColinearGoal", [1, 9, 17, 24], 0.0, None)]

    errors = importlib.import_module("jax_fdm.losses.errors")
    kinds = ["SquaredError", "AbsoluteError", "PredictionError", "LogMaxError"]
    names, preds, goals_, errs = [], [], [], []
    for name, idx, target, extra in cases:
        cls = getattr(importlib.import_module(GOAL_MODULES[name]), name)
        self = types.SimpleNamespace(vector=G.jarr(extra)) if name == "EdgeAngleGoal" else None
        index = G.jarr(idx, np.int64) if isinstance(idx, list) else idx
        p = np.asarray(cls.prediction(self, eq_state, None, index), dtype=np.float64)
        t = G.jarr(np.asarray(target, dtype=np.float64))
        g = np.asarray(cls.goal(self, t, G.jarr(p)), dtype=np.float64) if "goal" in cls.__dict__ or any(
            "goal" in b.__dict__ and b.__name__ != "Goal" for b in cls.__mro__[1:]) else np.asarray(target, dtype=np.float64)
        names.append(name)
        preds.append(p.ravel())
        goals_.append(np.asarray(g, dtype=np.float64).ravel())
        gs = GoalState(goal=G.jarr(g), weight=1.7, prediction=G.jarr(p))
        errs.append([float(getattr(errors, k).error(gs)) for k in kinds])
    out["case_names"] = np.array(names)
    out["case_index"] = np.array([c[1] if not isinstance(c[1], list) else -1 for c in cases])
    for i, c in enumerate(cases):
        out[f"case{i}_target"] = np.asarray(c[2], dtype=np.float64)
        if isinstance(c[1], list):
            out[f"case{i}_set"] = np.asarray(c[1])
        if c[3] is not None:
            out[f"case{i}_vector"] = np.asarray(c[3], dtype=np.float64)
        out[f"case{i}_ref_prediction"] = preds[i]
        out[f"case{i}_ref_goal"] = goals_[i]
    out["ref_errors_weight_1p7"] = np.asarray(errs)
    # aggregators (errors.py:262-400): per-collection errors [2.0, 3.5, 0.25], five goals, alpha = 0.8
    coll = G.jarr([2.0, 3.5, 0.25])
    agg = {}
    for k in ("SquaredError", "MeanSquaredError", "RootMeanSquaredError", "AbsoluteError", "MeanAbsoluteError",
              "PredictionError", "MeanPredictionError", "LogMaxError"):
        term = getattr(errors, k)(goals=[None] * 5, alpha=0.8)
        agg[k] = float(term.errors(coll)) * term.alpha
    out["agg_names"] = np.array(list(agg))
    out["ref_agg_values"] = np.array(list(agg.values()))
    path = os.path.join(HERE, "goals.npz")
    np.savez_compressed(path, **out)
    print(f"goals.npz: {len(cases)} goal cases, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
