"""
CPU: the oracle's goal predictions / reference values / error kernels / aggregators, and the goal classes of
jax_fdm_b200.goals evaluated goal by goal on torch CPU tensors, against values produced by the reference's OWN
goal, error and geometry code (tests/golden/make_golden_goals.py -> tests/golden/goals.npz).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from oracle import goals as og  # noqa: E402
from jax_fdm_b200 import goals as G, losses as LS  # noqa: E402
from jax_fdm_b200.equilibrium.states import EquilibriumState  # noqa: E402
from jax_fdm_b200.equilibrium.structures import EquilibriumStructureSparse  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "goals.npz")
KINDS = ["SquaredError", "AbsoluteError", "PredictionError", "LogMaxError"]


@pytest.fixture(scope="module")
def setup():
    g = dict(np.load(GOLDEN))
    so = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    st, _ = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], so)
    return g, so, st


def _oracle_case(g, st, i, name):
    idx, target = int(g["case_index"][i]), g[f"case{i}_target"]
    if name in ("EdgesLengthEqualGoal", "EdgesForceEqualGoal"):
        p = og.edges_equal(st, "lengths" if name == "EdgesLengthEqualGoal" else "forces", g[f"case{i}_set"])
        return np.atleast_1d(p), np.atleast_1d(0.0)
    if name == "NodesColinearGoal":
        return np.atleast_1d(og.nodes_colinear(st, g[f"case{i}_set"])), np.atleast_1d(0.0)
    if name == "EdgeAngleGoal":
        return np.atleast_1d(og.edge_angle(st, idx, g[f"case{i}_vector"])), np.atleast_1d(target)
    p = og.PREDICTIONS[name](st, idx)
    t = og.GOALS[name](p, target) if name in og.GOALS else np.asarray(target, dtype=np.float64)
    return np.atleast_1d(p).ravel(), np.atleast_1d(t).ravel()


def test_oracle_predictions_goals_and_errors_vs_reference(setup):
    g, so, st = setup
    for i, name in enumerate(g["case_names"]):
        p, t = _oracle_case(g, st, i, str(name))
        rp, rt = g[f"case{i}_ref_prediction"], g[f"case{i}_ref_goal"]
        assert np.abs(p - rp).max() <= 1e-13 * max(1.0, np.abs(rp).max()), (i, name)
        assert np.abs(t - rt).max() <= 1e-13 * max(1.0, np.abs(rt).max()), (i, name)
        for k, kind in enumerate(KINDS):
            got = og.ERRORS[kind](1.7, p, t)
            ref = g["ref_errors_weight_1p7"][i, k]
            assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (i, name, kind)


def test_oracle_aggregators_vs_reference(setup):
    g, _, _ = setup
    for name, ref in zip(g["agg_names"], g["ref_agg_values"]):
        name = str(name)
        total = 2.0 + 3.5 + 0.25
        if name.startswith("Mean") or name.startswith("RootMean"):
            total /= 5
        if name.startswith("Root"):
            total = np.sqrt(total)
        assert abs(0.8 * total - ref) <= 1e-14 * abs(ref), name
        # the same rule as coded in oracle.goals.error_term and in jax_fdm_b200.losses.Error.errors
        term = getattr(LS, name)([G.EdgeLengthGoal((0, 1), 1.0)] * 5, alpha=0.8)
        assert abs(float(term.errors(torch.tensor(5.75, dtype=torch.float64))) * term.alpha - ref) <= 1e-14 * abs(ref)


def test_torch_goal_classes_vs_reference(setup):
    g, so, st = setup
    s = EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=None)
    t = {k: torch.as_tensor(np.asarray(v, dtype=np.float64)) for k, v in st.items()}
    eqs = EquilibriumState(xyz=t["xyz"], residuals=t["residuals"], lengths=t["lengths"], forces=t["forces"],
                           loads=t["loads"], vectors=t["vectors"])
    nodes, edges = g["nodes"], g["edges"]
    for i, name in enumerate(g["case_names"]):
        name = str(name)
        cls = getattr(G, name)
        idx, target = int(g["case_index"][i]), g[f"case{i}_target"]
        if name in ("EdgesLengthEqualGoal", "EdgesForceEqualGoal"):
            goal = cls([(int(edges[e][0]), int(edges[e][1])) for e in g[f"case{i}_set"]])
        elif name == "NodesColinearGoal":
            goal = cls([int(nodes[n]) for n in g[f"case{i}_set"]])
        elif name == "NetworkLoadPathGoal":
            goal = cls()
        elif name == "EdgeAngleGoal":
            goal = cls((int(edges[idx][0]), int(edges[idx][1])), float(target), vector=g[f"case{i}_vector"])
        elif name.startswith("Edge"):
            goal = cls((int(edges[idx][0]), int(edges[idx][1])), target)
        else:
            goal = cls(int(nodes[idx]), target)
        gs = goal(eqs, s)
        rp, rt = g[f"case{i}_ref_prediction"], g[f"case{i}_ref_goal"]
        assert np.abs(np.atleast_1d(gs.prediction.numpy()).ravel() - rp).max() <= 1e-13 * max(1.0, np.abs(rp).max()), (i, name)
        assert np.abs(np.atleast_1d(gs.goal.numpy()).ravel() - rt).max() <= 1e-13 * max(1.0, np.abs(rt).max()), (i, name)
