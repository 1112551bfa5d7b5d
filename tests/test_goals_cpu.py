"""
CPU: the goal / error classes evaluated goal by goal on torch CPU tensors (Error.evaluate_state, the
reference's errors.py:165-200 route) against the oracle's numpy restatement, and the goal program a Loss
compiles for the fused kernel (term table sorted by error term, scales of the Mean / Root aggregates).
No CUDA calls.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from oracle import goals as og  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402
from jax_fdm_b200 import goals as G, losses as LS  # noqa: E402
from jax_fdm_b200.equilibrium.states import EquilibriumState  # noqa: E402
from jax_fdm_b200.equilibrium.structures import EquilibriumStructureSparse  # noqa: E402


def _state(prob):
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    st, _ = oracle.forward(prob.q, prob.xyz_fixed, prob.loads, so)
    t = {k: torch.as_tensor(np.asarray(v, dtype=np.float64)) for k, v in st.items()}
    return st, EquilibriumState(xyz=t["xyz"], residuals=t["residuals"], lengths=t["lengths"], forces=t["forces"],
                                loads=t["loads"], vectors=t["vectors"])


def test_goalwise_errors_match_the_oracle():
    prob = datasets.cablenet_problem(6, seed=2)
    s = EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=None)
    st, eqs = _state(prob)
    e0, e1 = [(int(u), int(v)) for u, v in prob.edges[:2]]
    n0, n1 = int(prob.nodes[7]), int(prob.nodes[0])
    cases = [
        (LS.SquaredError, "SquaredError", [(G.NodePointGoal, "NodePointGoal", n0, [0.1, 0.2, 0.3], 2.0)]),
        (LS.MeanSquaredError, "MeanSquaredError", [(G.EdgeLengthGoal, "EdgeLengthGoal", e0, 1.2, 1.0), (G.EdgeLengthGoal, "EdgeLengthGoal", e1, 0.7, 3.0)]),
        (LS.RootMeanSquaredError, "RootMeanSquaredError", [(G.EdgeForceGoal, "EdgeForceGoal", e0, 0.5, 1.0), (G.EdgeForceGoal, "EdgeForceGoal", e1, 1.5, 1.0)]),
        (LS.AbsoluteError, "AbsoluteError", [(G.NodeResidualForceGoal, "NodeResidualForceGoal", n1, 0.3, 1.5)]),
        (LS.MeanAbsoluteError, "MeanAbsoluteError", [(G.NodeResidualVectorGoal, "NodeResidualVectorGoal", n1, [0.1, -0.2, 0.0], 1.0)]),
        (LS.PredictionError, "PredictionError", [(G.EdgeLoadPathGoal, "EdgeLoadPathGoal", e1, 0.0, 2.0)]),
        (LS.LogMaxError, "LogMaxError", [(G.EdgeLengthGoal, "EdgeLengthGoal", e0, 0.5, 1.0)]),
        (LS.SquaredError, "SquaredError", [(G.NodeLineGoal, "NodeLineGoal", n0, [[0, 0, 1], [3, 4, 2]], 1.5),
                                           (G.NodePlaneGoal, "NodePlaneGoal", n0, [[0, 0, 0.5], [0.1, 0.2, 2.0]], 1.0),
                                           (G.NodeSegmentGoal, "NodeSegmentGoal", n0, [[0, 0, 1], [0.5, 0.5, 1.2]], 2.0),
                                           (G.NodeSegmentGoal, "NodeSegmentGoal", n0, [[-50, 0, 1], [60, 40, 2]], 1.0)]),
        (LS.SquaredError, "SquaredError", [(G.NodeResidualDirectionGoal, "NodeResidualDirectionGoal", n1, [0.1, -0.3, 1.0], 1.5),
                                           (G.NodeResidualPlaneGoal, "NodeResidualPlaneGoal", n1, [0.2, 1.0, 0.3], 2.0),
                                           (G.EdgeDirectionGoal, "EdgeDirectionGoal", e0, [0.3, 1.0, -0.2], 1.0)]),
    ]
    for ecls, ename, goals in cases:
        term = ecls([g(key, target, weight) for g, _, key, target, weight in goals], alpha=0.7)
        idx = [(gn, s.node_index[key] if isinstance(key, int) else s.edge_index[key], target, weight)
               for _, gn, key, target, weight in goals]
        ref = og.error_term(ename, idx, 0.7, st)
        got = term.evaluate_state(eqs, s).item()
        assert abs(got - ref) <= 1e-13 * max(1.0, abs(ref)), (ename, got, ref)
    keys = [(int(u), int(v)) for u, v in prob.edges[:7]]
    agg = LS.SquaredError([G.EdgesLengthEqualGoal(keys, 2.0), G.EdgesForceEqualGoal(keys)], alpha=1.3)
    ref = og.error_term("SquaredError", [("EdgesLengthEqualGoal", list(range(7)), 0.0, 2.0),
                                         ("EdgesForceEqualGoal", list(range(7)), 0.0, 1.0)], 1.3, st)
    assert abs(agg.evaluate_state(eqs, s).item() - ref) <= 1e-13 * abs(ref)
    net = LS.MeanPredictionError([G.NetworkLoadPathGoal()], alpha=0.01)
    ref = og.error_term("MeanPredictionError", [("NetworkLoadPathGoal", 0, 0.0, 1.0)], 0.01, st)
    assert abs(net.evaluate_state(eqs, s).item() - ref) <= 1e-13 * abs(ref)


def test_goal_program_layout():
    prob = datasets.cablenet_problem(5, seed=1)
    s = EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=None)
    e = [(int(u), int(v)) for u, v in prob.edges[:3]]
    loss = LS.Loss(LS.MeanSquaredError([G.EdgeLengthGoal(e[0], 1.0), G.EdgeLengthGoal(e[2], 2.0, 3.0)], alpha=0.5),
                   LS.RootMeanSquaredError([G.NodePointGoal(int(prob.nodes[4]), [1, 2, 3])]),
                   LS.PredictionError([G.NetworkLoadPathGoal()], alpha=0.1), LS.L2Regularizer(1e-3))
    assert loss.number_of_goals() == 4 and loss.number_of_regularizers() == 1
    prog = loss.program(s, torch.device("cpu"))
    kind, index, err, weight, target, gptr, gfinal, ginner, gouter, agg_ptr, agg_index = [a.numpy() for a in prog.arrays]
    assert agg_ptr.tolist() == [0, 0, 0, 0, 0] and prog.struct.num_aggregates == 0
    assert kind.tolist() == [3, 3, 0, 6] and err.tolist() == [0, 0, 0, 2]
    assert index.tolist() == [0, 2, 4, 0] and weight.tolist() == [1.0, 3.0, 1.0, 1.0]
    assert target.tolist() == [[1, 0, 0, 0, 0, 0], [2, 0, 0, 0, 0, 0], [1, 2, 3, 0, 0, 0], [0, 0, 0, 0, 0, 0]]
    assert gptr.tolist() == [0, 2, 3, 4] and gfinal.tolist() == [0, 1, 0]
    assert ginner.tolist() == [0.5, 1.0, 1.0] and gouter.tolist() == [0.5, 1.0, 0.1]
    assert prog.struct.num_terms == 4 and prog.struct.num_groups == 3 and prog.struct.has_network_loadpath == 1
    with pytest.raises(TypeError):
        LS.SquaredError([object()])
    with pytest.raises(ValueError):
        G.NodePointGoal(0, 1.0)
