"""
The scipy seam on the GPU path (reference: optimization/optimizers/optimizer.py:283-415, 453-515 and the
README's constrained_fdm example): L-BFGS-B over bounded force densities with a goal loss.  Every
iteration is one forward + adjoint pass on the device.  Checks: the optimiser's own first-order
optimality, a >= 100x loss reduction, bounds respected, and agreement of the optimum with the same scipy
run driven by the ORACLE's loss and central-difference gradients (small problem).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import goals as og  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402


def test_lbfgsb_over_q_reaches_the_oracle_optimum():
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS
    from jax_fdm_b200.optimization import LBFGSB, optimize_q
    from scipy.optimize import minimize

    prob = datasets.arch_problem(10)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    keys = [(int(u), int(v)) for u, v in prob.edges]
    target = 0.55
    loss = LS.Loss(LS.SquaredError([G.EdgeLengthGoal(k, target) for k in keys]), LS.L2Regularizer(1e-6))
    model = eq.EquilibriumModelSparse(tmax=1)
    q_opt, res = optimize_q(model, s, loss, prob.q, prob.xyz_fixed, prob.loads, optimizer=LBFGSB(),
                            bounds=(-50.0, -0.01), maxiter=500, tol=1e-12)
    oterms = [("SquaredError", [("EdgeLengthGoal", i, target, 1.0) for i in range(len(keys))], 1.0)]

    def f(q):
        st, _ = oracle.forward(q, prob.xyz_fixed, prob.loads, so)
        return og.loss(oterms, st, q=q, l2_alpha=1e-6)

    f0 = f(prob.q)
    assert res.fun <= 1e-2 * f0 and abs(f(q_opt) - res.fun) <= 1e-10 * max(1.0, abs(res.fun))
    assert np.all(q_opt <= -0.01 + 1e-12) and np.all(q_opt >= -50.0 - 1e-12)

    def f_and_fd(q):
        g = np.zeros_like(q)
        for k in range(q.size):
            d = np.zeros_like(q)
            d[k] = 1e-6
            g[k] = (f(q + d) - f(q - d)) / 2e-6
        return f(q), g

    ref = minimize(f_and_fd, prob.q, jac=True, method="L-BFGS-B", bounds=[(-50.0, -0.01)] * prob.q.size, tol=1e-12,
                   options={"maxiter": 500})
    assert abs(ref.fun - res.fun) <= 1e-6 * max(abs(ref.fun), 1e-12) + 1e-10
    assert np.abs(ref.x - q_opt).max() <= 1e-3 * np.abs(ref.x).max()


def test_batched_descent_step_uses_one_pass_for_all_designs():
    """vmap-style use (loss_plotter.py:98-103): B designs, B losses and B gradients from one pass."""
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS

    prob = datasets.pillow_problem(12)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    loss = LS.Loss(LS.MeanSquaredError([G.EdgeLengthGoal((int(u), int(v)), 0.8) for u, v in prob.edges]))
    model = eq.EquilibriumModelSparse(tmax=1)
    B = 16
    q = torch.as_tensor(datasets.pillow_batch_q(prob.edges.shape[0], 0, B), device="cuda:0").requires_grad_(True)
    xs = torch.as_tensor(prob.xyz_fixed, device="cuda:0")
    p = torch.as_tensor(prob.loads, device="cuda:0")
    v0 = loss(eq.EquilibriumParametersState(q, xs, eq.LoadState(p, 0.0, 0.0)), model, s)
    v0.sum().backward()
    step = 1e-2 * q.grad / q.grad.abs().amax(dim=1, keepdim=True)
    v1 = loss(eq.EquilibriumParametersState((q - step).detach(), xs, eq.LoadState(p, 0.0, 0.0)), model, s)
    assert v0.shape == (B,) and bool((v1 < v0).all())


def test_constrained_fdm_on_arrays():
    """The README flow of the reference (constrained_fdm(network, optimizer=LBFGSB(), loss=Loss(SquaredError(goals)))),
    COMPAS-free: the state at the optimum meets the length goals better than the start by two orders of magnitude."""
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS
    from jax_fdm_b200.optimization import LBFGSB

    prob = datasets.arch_problem(10)
    keys = [(int(u), int(v)) for u, v in prob.edges]
    loss = LS.Loss(LS.SquaredError([G.EdgeLengthGoal(k, 0.55) for k in keys]))
    s0, structure = eq.fdm(prob.nodes, prob.edges, prob.supports, prob.q, prob.xyz_fixed, prob.loads)
    state, structure, q_opt = eq.constrained_fdm(prob.nodes, prob.edges, prob.supports, prob.q, prob.xyz_fixed, prob.loads,
                                                 LBFGSB(), loss, bounds=(-50.0, -0.01), maxiter=300, tol=1e-12,
                                                 structure=structure)
    e0 = float(((s0.lengths - 0.55) ** 2).sum())
    e1 = float(((state.lengths - 0.55) ** 2).sum())
    assert e1 <= 1e-2 * e0 and q_opt.shape == prob.q.shape and np.all(q_opt < 0)


def test_inverse_design_over_q_supports_and_loads():
    """BASELINE config 3 in miniature (gridshell inverse design over q, loads and supports with vertex goals):
    the flat parameter vector [q, z of the supports, z of the free nodes' loads], point goals on the free
    nodes.  The gradient of the packed loss against central differences of the oracle, then L-BFGS-B."""
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS
    from jax_fdm_b200.optimization import LBFGSB

    prob = datasets.pillow_problem(6)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    free = np.flatnonzero(prob.supports == 0)
    rng = np.random.default_rng(3)
    target = prob.xyz0[free] + np.c_[np.zeros((free.size, 2)), 0.6 + 0.3 * np.sin(prob.xyz0[free, 0])]
    loss = LS.Loss(LS.MeanSquaredError([G.VertexPointGoal(int(prob.nodes[n]), t) for n, t in zip(free, target)]),
                   LS.L2Regularizer(1e-5))
    model = eq.EquilibriumModelSparse(tmax=1)
    mxs = np.zeros(prob.xyz_fixed.shape, dtype=bool); mxs[:, 2] = True
    mp = np.zeros(prob.loads.shape, dtype=bool); mp[free, 2] = True
    opt = LBFGSB()
    problem = opt.problem(model, s, loss, prob.q, prob.xyz_fixed, prob.loads, bounds=(-20.0, -0.05), maxiter=200, tol=1e-10,
                          opt_xyz_fixed=mxs, bounds_xyz_fixed=(-1.0, 2.0), opt_loads=mp, bounds_loads=(-1.0, 1.0))
    x0 = problem.x0
    assert x0.size == prob.q.size + mxs.sum() + mp.sum()
    oterms = [("MeanSquaredError", [("NodePointGoal", int(n), t, 1.0) for n, t in zip(free, target)], 1.0)]

    def f(x):
        q, xs, p = problem.unpack(x)
        st, _ = oracle.forward(q, xs, p, so)
        return og.loss(oterms, st, q=q, l2_alpha=1e-5)

    val, grad = problem.fun(x0)
    assert abs(val - f(x0)) <= 1e-12 * abs(val)
    scale = np.abs(grad).max()
    from conftest import fd5
    for k in rng.choice(x0.size, size=15, replace=False):
        fd = fd5(f, x0, k)
        assert abs(grad[k] - fd) <= 1e-8 * scale + 1e-11, (k, grad[k], fd)
    x = opt.solve(problem)
    assert opt.result.fun <= 0.05 * val
    lo = np.array([b[0] for b in problem.bounds]); hi = np.array([b[1] for b in problem.bounds])
    assert np.all(x >= lo - 1e-12) and np.all(x <= hi + 1e-12)


def test_liew_dome_inverse_design_steps():
    """BASELINE config 3 on the reference's own dome (tests/data/liew_dome.json through tests/golden/liew_dome.npz):
    vertex point goals 10 % above the form-found heights, parameters q + support z + nodal load z; a few
    L-BFGS-B iterations must cut the loss, and the packed gradient must match the oracle's finite differences."""
    import os
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS
    from jax_fdm_b200.optimization import LBFGSB

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "liew_dome.npz")))
    nodes, edges, supports = g["nodes"], g["edges"], g["supports"]
    s = eq.EquilibriumStructureSparse(nodes, edges, supports, device=0)
    so = oracle.build_structure(nodes, edges, supports)
    free = np.flatnonzero(supports == 0)
    xyz_ref = g["ref_state_xyz"]
    target = xyz_ref[free] * np.array([1.0, 1.0, 1.1])
    loss = LS.Loss(LS.MeanSquaredError([G.VertexPointGoal(int(nodes[n]), t) for n, t in zip(free, target)]))
    model = eq.EquilibriumModelSparse(tmax=1)
    mxs = np.zeros(g["xyz_fixed"].shape, dtype=bool); mxs[:, 2] = True
    mp = np.zeros(g["loads"].shape, dtype=bool); mp[free, 2] = True
    qsign = np.sign(g["q"][0])
    qb = (-1e3, -1e-3) if qsign < 0 else (1e-3, 1e3)
    opt = LBFGSB()
    problem = opt.problem(model, s, loss, g["q"], g["xyz_fixed"], g["loads"], bounds=qb, maxiter=15, tol=1e-12,
                          opt_xyz_fixed=mxs, opt_loads=mp)
    x0 = problem.x0
    oterms = [("MeanSquaredError", [("NodePointGoal", int(n), t, 1.0) for n, t in zip(free, target)], 1.0)]

    def f(x):
        q, xs, p = problem.unpack(x)
        st, _ = oracle.forward(q, xs, p, so)
        return og.loss(oterms, st)

    val, grad = problem.fun(x0)
    assert abs(val - f(x0)) <= 1e-10 * abs(val)
    rng = np.random.default_rng(0)
    scale = np.abs(grad).max()
    from conftest import fd5
    for k in rng.choice(x0.size, size=8, replace=False):
        fd = fd5(f, x0, k)
        assert abs(grad[k] - fd) <= 1e-8 * scale + 1e-11, (k, grad[k], fd)
    opt.solve(problem)
    assert opt.result.fun <= 0.5 * val
