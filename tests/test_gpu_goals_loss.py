"""
Goals and losses on the GPU: the fused kernel (fdm_goals_loss) against the oracle's numpy restatement
of the reference's goal predictions / error terms (goals/*, losses/errors.py:232-431, losses/loss.py:
60-90), against the goal-by-goal torch evaluation, and the gradient of Loss(params, model, structure)
through the whole forward path against central differences of the oracle.  Tolerances: 1e-12 relative on
the loss, 1e-6 relative on finite-difference gradients.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import goals as og  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402


def T(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=torch.device("cuda", 0))
    return t.requires_grad_(True) if grad else t


def build_loss(prob, s, rng):
    """A loss with every goal kind and every error kind; returns (Loss, the same terms for the oracle)."""
    from jax_fdm_b200 import goals as G, losses as LS

    nodes, edges = prob.nodes, prob.edges
    free = np.flatnonzero(prob.supports == 0)
    fixed = np.flatnonzero(prob.supports != 0)
    pick_n = rng.choice(free, size=min(12, free.size), replace=False)
    pick_e = rng.choice(len(edges), size=min(15, len(edges)), replace=False)
    ni, ei = s.node_index, s.edge_index

    def ekey(e):
        return (int(edges[e][0]), int(edges[e][1]))

    spec = [
        ("SquaredError", 1.0, [("NodePointGoal", int(nodes[n]), prob.xyz0[n] + 0.1 * rng.standard_normal(3), rng.uniform(0.5, 2)) for n in pick_n]),
        ("MeanSquaredError", 0.7, [("EdgeLengthGoal", ekey(e), rng.uniform(0.5, 1.5), rng.uniform(0.5, 2)) for e in pick_e]),
        ("RootMeanSquaredError", 1.3, [("EdgeForceGoal", ekey(e), rng.uniform(0.5, 3.0), 1.0) for e in pick_e[:7]]),
        ("AbsoluteError", 0.5, [("NodeResidualForceGoal", int(nodes[n]), 0.3, 1.5) for n in fixed[:3]]),
        ("MeanAbsoluteError", 0.9, [("NodeResidualVectorGoal", int(nodes[n]), 0.2 * rng.standard_normal(3), 1.0) for n in fixed[:4]]),
        ("PredictionError", 0.01, [("NetworkLoadPathGoal", -1, 0.0, 1.0)]),
        ("MeanPredictionError", 0.05, [("EdgeLoadPathGoal", ekey(e), 0.0, 2.0) for e in pick_e[:5]]),
        ("LogMaxError", 2.0, [("EdgeLengthGoal", ekey(e), 0.8, 1.0) for e in pick_e[5:12]]),
        ("SquaredError", 1e-4, [("NetworkLoadPathGoal", -1, 50.0, 1.0), ("EdgeLengthGoal", ekey(pick_e[0]), 1.0, 3.0)]),
        ("SquaredError", 0.3, [(f"Node{ax}CoordinateGoal", int(nodes[n]), prob.xyz0[n, c] - 0.2, 1.0 + c)
                               for c, ax in enumerate("XYZ") for n in pick_n[:4]]),
        # reference values that are projections of the prediction (differentiated through, as JAX does)
        ("SquaredError", 0.8, [("NodeLineGoal", int(nodes[pick_n[0]]), [prob.xyz0[pick_n[0]] + [0.3, -0.2, 0.5], prob.xyz0[pick_n[0]] + [1.0, 2.0, 3.0]], 1.2),
                               ("NodePlaneGoal", int(nodes[pick_n[1]]), [prob.xyz0[pick_n[1]] + [0.0, 0.0, 0.4], [0.2, -0.1, 1.0]], 0.7),
                               ("NodeSegmentGoal", int(nodes[pick_n[2]]), [prob.xyz0[pick_n[2]] + [-1.0, 0.1, 0.3], prob.xyz0[pick_n[2]] + [1.0, 0.4, 0.6]], 1.0),
                               ("NodeSegmentGoal", int(nodes[pick_n[3]]), [prob.xyz0[pick_n[3]] + [5.0, 5.0, 1.0], prob.xyz0[pick_n[3]] + [9.0, 6.0, 2.0]], 1.0)]),
        ("SquaredError", 1.1, [("NodeResidualDirectionGoal", int(nodes[n]), [0.1, -0.3, 1.0], 1.0 + 0.5 * j) for j, n in enumerate(fixed[:3])]
                              + [("NodeResidualPlaneGoal", int(nodes[fixed[3 % len(fixed)]]), [0.2, 1.0, 0.3], 2.0)]),
        ("SquaredError", 0.9, [("EdgeDirectionGoal", ekey(e), [0.3, 1.0, -0.2], 1.0) for e in pick_e[:4]]
                              + [("EdgeAngleGoal", ekey(e), (0.4 + 0.1 * j, [0.0, 0.2, 1.0]), 1.5) for j, e in enumerate(pick_e[4:8])]),
        ("SquaredError", 3.0, [("EdgesLengthEqualGoal", [ekey(e) for e in pick_e[:9]], 0.0, 2.0),
                               ("EdgesForceEqualGoal", [ekey(e) for e in range(len(edges))], 0.0, 1.0),
                               ("NodesColinearGoal", [int(nodes[n]) for n in free[:6]], 0.0, 1.5)]),
        ("AbsoluteError", 0.4, [("NodeLineGoal", int(nodes[pick_n[4]]), [prob.xyz0[pick_n[4]] + [0.3, 0.2, -0.4], prob.xyz0[pick_n[4]] + [-1.0, 1.5, 2.0]], 1.0),
                                ("NodePlaneGoal", int(nodes[pick_n[5]]), [prob.xyz0[pick_n[5]] + [0.1, 0.0, 0.3], [1.0, 0.5, 0.25]], 2.0)]),
    ]
    terms, oterms = [], []
    for ename, alpha, goals in spec:
        gl, ol = [], []
        for gname, key, target, weight in goals:
            cls = getattr(G, gname)
            if gname in ("EdgesLengthEqualGoal", "EdgesForceEqualGoal", "NodesColinearGoal"):
                gl.append(cls(key, weight))
                ol.append((gname, [ni[k] for k in key] if gname == "NodesColinearGoal" else [ei[k] for k in key], 0.0, weight))
                continue
            if gname == "EdgeAngleGoal":
                g = cls(key, target[0], weight, vector=target[1])
            else:
                g = cls(target=target, weight=weight) if gname == "NetworkLoadPathGoal" else cls(key, target, weight)
            gl.append(g)
            idx = 0 if gname == "NetworkLoadPathGoal" else (ni[key] if isinstance(key, int) else ei[key])
            ol.append((gname, idx, target if gname == "EdgeAngleGoal" else np.asarray(target, dtype=np.float64), weight))
        terms.append(getattr(LS, ename)(gl, alpha=alpha))
        oterms.append((ename, ol, alpha))
    l2 = 1e-3
    return LS.Loss(*terms, LS.L2Regularizer(l2)), oterms, l2


def oracle_loss(q, xs, loads, so, oterms, l2):
    st, _ = oracle.forward(q, xs, loads, so)
    return og.loss(oterms, st, q=q, l2_alpha=l2)


def test_fused_loss_matches_oracle_and_goalwise_torch():
    import jax_fdm_b200.equilibrium as eq

    prob = datasets.cablenet_problem(8, seed=4)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    rng = np.random.default_rng(0)
    loss, oterms, l2 = build_loss(prob, s, rng)
    model = eq.EquilibriumModelSparse(tmax=1)
    B = 5
    qb = prob.q[None, :] * rng.uniform(0.8, 1.25, size=(B, prob.q.size))
    params = eq.EquilibriumParametersState(T(qb), T(prob.xyz_fixed), eq.LoadState(T(prob.loads), 0.0, 0.0))
    got = loss(params, model, s).cpu().numpy()
    assert got.shape == (B,)
    for b in range(B):
        ref = oracle_loss(qb[b], prob.xyz_fixed, prob.loads, so, oterms, l2)
        assert abs(got[b] - ref) <= 1e-12 * abs(ref), (b, got[b], ref)
    # unbatched, and the goal-by-goal evaluation of errors.py:165-200
    p1 = eq.EquilibriumParametersState(T(prob.q), T(prob.xyz_fixed), eq.LoadState(T(prob.loads), 0.0, 0.0))
    one = loss(p1, model, s)
    ref = oracle_loss(prob.q, prob.xyz_fixed, prob.loads, so, oterms, l2)
    assert one.dim() == 0 and abs(one.item() - ref) <= 1e-12 * abs(ref)
    st = model(p1, s)
    slow = sum(t.evaluate_state(st, s) for t in loss.terms_error) + loss.terms_regularization[0](p1)
    assert abs(slow.item() - ref) <= 1e-12 * abs(ref)


def test_loss_gradients_through_the_model_vs_finite_differences():
    import jax_fdm_b200.equilibrium as eq

    prob = datasets.cablenet_problem(7, seed=11)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    rng = np.random.default_rng(1)
    loss, oterms, l2 = build_loss(prob, s, rng)
    model = eq.EquilibriumModelSparse(tmax=1)
    q, xs, p = T(prob.q, True), T(prob.xyz_fixed, True), T(prob.loads, True)
    val = loss(eq.EquilibriumParametersState(q, xs, eq.LoadState(p, 0.0, 0.0)), model, s)
    val.backward()
    for grad, arr, mk in [(q.grad, prob.q, lambda a: (a, prob.xyz_fixed, prob.loads)),
                          (xs.grad, prob.xyz_fixed, lambda a: (prob.q, a, prob.loads)),
                          (p.grad, prob.loads, lambda a: (prob.q, prob.xyz_fixed, a))]:
        gnp = grad.cpu().numpy()
        scale = np.abs(gnp).max()
        from conftest import fd5
        for k in rng.choice(arr.size, size=8, replace=False):
            fd = fd5(lambda a: oracle_loss(*mk(a), so, oterms, l2), arr, k, h=1e-4)   # kinks (|.|, max) stay far at this step
            assert abs(gnp.flat[k] - fd) <= 1e-8 * scale + 1e-11, (k, gnp.flat[k], fd)


def test_batched_gradients_are_per_design():
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS

    prob = datasets.pillow_problem(10)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    loss = LS.Loss(LS.SquaredError([G.EdgeLengthGoal((int(u), int(v)), 0.9) for u, v in prob.edges]),
                   LS.PredictionError([G.NetworkLoadPathGoal()], alpha=0.01))
    model = eq.EquilibriumModelSparse(tmax=1)
    B = 6
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    q = T(qb, True)
    vals = loss(eq.EquilibriumParametersState(q, T(prob.xyz_fixed), eq.LoadState(T(prob.loads), 0.0, 0.0)), model, s)
    vals.sum().backward()
    for b in (0, B - 1):
        q1 = T(qb[b], True)
        v1 = loss(eq.EquilibriumParametersState(q1, T(prob.xyz_fixed), eq.LoadState(T(prob.loads), 0.0, 0.0)), model, s)
        v1.backward()
        assert abs(v1.item() - vals[b].item()) <= 1e-13 * abs(v1.item())
        assert (q1.grad - q.grad[b]).abs().max().item() <= 1e-12 * q1.grad.abs().max().item()


@pytest.mark.parametrize("kinds", ["nodes", "edges", "residuals", "nodes+edges"])
def test_partial_goal_programs_skip_what_they_do_not_read(kinds):
    """A loss of node goals only never launches the state kernel or its VJP; one of edge goals only gets no xyz
    cotangent array, and so on (fdm_goal_kind_arrays).  Value and all three gradients against the goal-by-goal route
    through the model (autograd over the state), for the fused node and for the library-owned pipeline."""
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS
    from jax_fdm_b200.pipeline import NativeForwardAdjoint

    prob = datasets.cablenet_problem(9, seed=2)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    rng = np.random.default_rng(5)
    free = np.flatnonzero(prob.supports == 0)
    fixed = np.flatnonzero(prob.supports != 0)
    terms = []
    if "nodes" in kinds:
        terms.append(LS.SquaredError([G.NodePointGoal(int(prob.nodes[n]), prob.xyz0[n] + 0.1 * rng.standard_normal(3))
                                      for n in free[::3]], alpha=0.7))
    if "edges" in kinds:
        terms.append(LS.MeanSquaredError([G.EdgeLengthGoal((int(u), int(v)), 0.9) for u, v in prob.edges[::4]]))
        terms.append(LS.SquaredError([G.EdgeForceGoal((int(u), int(v)), 1.5) for u, v in prob.edges[1::5]], alpha=0.2))
    if "residuals" in kinds:
        terms.append(LS.SquaredError([G.NodeResidualVectorGoal(int(prob.nodes[n]), [0.1, 0.0, -0.2]) for n in fixed]))
    loss = LS.Loss(*terms)
    assert loss.program(s, torch.device("cuda", 0)).touches == ("nodes" in kinds, "edges" in kinds, "edges" in kinds,
                                                                 "residuals" in kinds)
    model = eq.EquilibriumModelSparse(tmax=1)
    B = 9
    qb = prob.q[None, :] * rng.uniform(0.8, 1.25, size=(B, prob.q.size))

    def run(fused):
        q, xs, p = T(qb, True), T(prob.xyz_fixed, True), T(prob.loads, True)
        params = eq.EquilibriumParametersState(q, xs, eq.LoadState(p, 0.0, 0.0))
        val = loss(params, model, s) if fused else loss.errors_from_state(model(params, s), s)
        w = T(np.linspace(0.5, 1.5, B))                      # a non-trivial upstream cotangent
        (val * w).sum().backward()
        return val.detach().cpu().numpy(), [t.grad.cpu().numpy() for t in (q, xs, p)]

    v1, g1 = run(True)
    v2, g2 = run(False)
    assert np.allclose(v1, v2, rtol=1e-13, atol=0)
    for a, b in zip(g1, g2):
        assert np.abs(a - b).max() <= 1e-11 * max(np.abs(b).max(), 1e-300)
    # the library-owned pipeline on the same program (unit upstream cotangent)
    fa = NativeForwardAdjoint(s, B, chunks=2, solver="cholesky", with_state=False)
    fa.set_inputs(qb, prob.xyz_fixed, prob.loads)
    fa.set_loss(loss)
    fa.step()
    fa.stream.synchronize()
    q = T(qb, True)
    loss(eq.EquilibriumParametersState(q, T(prob.xyz_fixed), eq.LoadState(T(prob.loads), 0.0, 0.0)), model, s).sum().backward()
    assert np.allclose(fa.array("loss").cpu().numpy(), v1, rtol=1e-13, atol=0)
    assert np.abs(fa.array("q_bar").cpu().numpy() - q.grad.cpu().numpy()).max() <= 1e-11 * np.abs(q.grad.cpu().numpy()).max()
    fa.close()
