"""
GPU: forward mode on the sparse path and its consumers.
  * `EquilibriumModel.jvp`: T tangents of (q, xyz_fixed, loads) at one primal -> tangents of every state field, the T
    tangent solves as right-hand sides of the primal factor (band Cholesky: one batched launch, reuse_factor = 2;
    iterative solvers: a loop on the stored coarse inverse).  Checked against central differences of the ORACLE's
    forward map and, exactly, against the reverse mode: <w, J v> = <J^T w, v>.
  * Constraint Jacobians (`jacfwd`, optimizers/constrained.py:38-102) for SLSQP / trust-constr, against central
    differences; SLSQP with edge-length constraints on the reference's Liew dome.
  * The Hessian handed to second-order optimisers (second_order.py:20-39), against second differences of the oracle.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import goals as og  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def T(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda:0")
    return t.requires_grad_(True) if grad else t


def N(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("case", ["pillow", "cablenet"])
def test_jvp_against_finite_differences_and_reverse_mode(case):
    import jax_fdm_b200.equilibrium as eq
    prob = datasets.pillow_problem(12) if case == "pillow" else datasets.cablenet_problem(40, seed=4)   # direct / iterative
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    model = eq.EquilibriumModelSparse(tmax=1, tol=1e-14)
    rng = np.random.default_rng(7)
    q = prob.q * (1.0 + 0.1 * rng.standard_normal(prob.q.size)) if case == "pillow" else prob.q
    nT = 4
    qd = rng.standard_normal((nT, q.size))
    xsd = rng.standard_normal((nT,) + prob.xyz_fixed.shape)
    pd = rng.standard_normal((nT,) + prob.loads.shape)
    params = eq.EquilibriumParametersState(T(q), T(prob.xyz_fixed), eq.LoadState(T(prob.loads), 0.0, 0.0))
    st, sd = model.jvp(params, s, q_dot=T(qd), xyz_fixed_dot=T(xsd), loads_dot=T(pd))
    ref, _ = oracle.forward(q, prob.xyz_fixed, prob.loads, so)
    assert np.abs(N(st.xyz) - ref["xyz"]).max() <= 1e-10 * np.abs(ref["xyz"]).max()
    h = 1e-6
    for t in range(nT):
        sp, _ = oracle.forward(q + h * qd[t], prob.xyz_fixed + h * xsd[t], prob.loads + h * pd[t], so)
        sm, _ = oracle.forward(q - h * qd[t], prob.xyz_fixed - h * xsd[t], prob.loads - h * pd[t], so)
        for f in ("xyz", "lengths", "forces", "residuals", "vectors"):
            fd = (sp[f] - sm[f]) / (2 * h)
            got = N(getattr(sd, f))[t]
            assert np.abs(got - fd).max() <= 1e-6 * max(np.abs(fd).max(), 1.0), (t, f)
    # a single unbatched tangent = the first of the batch
    _, s1 = model.jvp(params, s, q_dot=T(qd[0]), xyz_fixed_dot=T(xsd[0]), loads_dot=T(pd[0]))
    assert np.abs(N(s1.xyz) - N(sd.xyz)[0]).max() <= 1e-12 * np.abs(N(sd.xyz)[0]).max()
    # duality with the reverse mode: <w, J v> = <J^T w, v>
    w = {f: rng.standard_normal(N(getattr(st, f)).shape) for f in ("xyz", "lengths", "forces", "residuals")}
    qg, xg, pg = T(q, True), T(prob.xyz_fixed, True), T(prob.loads, True)
    st2 = model(eq.EquilibriumParametersState(qg, xg, eq.LoadState(pg, 0.0, 0.0)), s)
    sum((T(w[f]) * getattr(st2, f)).sum() for f in w).backward()
    for t in range(nT):
        lhs = sum(float((w[f] * N(getattr(sd, f))[t]).sum()) for f in w)
        rhs = float((N(qg.grad) * qd[t]).sum() + (N(xg.grad) * xsd[t]).sum() + (N(pg.grad) * pd[t]).sum())
        assert abs(lhs - rhs) <= 1e-9 * max(abs(lhs), abs(rhs), 1e-12), (t, lhs, rhs)


def test_constraint_jacobian_and_slsqp_on_the_liew_dome():
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import constraints as CT, goals as G, losses as LS
    from jax_fdm_b200.optimization import SLSQP
    g = dict(np.load(os.path.join(GOLDEN, "liew_dome.npz")))
    nodes, edges, supports = g["nodes"], g["edges"], g["supports"]
    s = eq.EquilibriumStructureSparse(nodes, edges, supports, device=0)
    so = oracle.build_structure(nodes, edges, supports)
    free = np.flatnonzero(supports == 0)
    target = g["ref_state_xyz"][free] * np.array([1.0, 1.0, 1.15])
    loss = LS.Loss(LS.MeanSquaredError([G.NodePointGoal(int(nodes[n]), t) for n, t in zip(free, target)]))
    model = eq.EquilibriumModelSparse(tmax=1)
    groups = [np.arange(k, edges.shape[0], 16) for k in range(16)]            # 16 shared force densities
    len0 = g["ref_state_lengths"].reshape(-1)
    picked = np.argsort(len0)[-40:]                                           # the 40 longest edges may not grow
    cons = [CT.EdgeLengthConstraint(tuple(int(k) for k in edges[e]), None, 1.002 * len0[e]) for e in picked]
    cons.append(CT.NodeZCoordinateConstraint(int(nodes[free[0]]), None, float(g["ref_state_xyz"][free[0], 2]) + 0.05))
    sign = np.sign(g["q"][0])
    bounds = (-1e3, -1e-3) if sign < 0 else (1e-3, 1e3)
    opt = SLSQP()
    problem = opt.problem(model, s, loss, g["q"], g["xyz_fixed"], g["loads"], bounds=bounds, maxiter=25, tol=1e-10,
                          q_groups=groups, constraints=cons)
    x0 = problem.x0
    assert x0.size == 16 and len(problem.constraints) == 2
    # Jacobians against central differences of the oracle
    for (fun, jac), field in zip(problem.constraint_functions, ("lengths", "xyz")):
        J = jac(x0)
        assert J.shape == (fun(x0).size, x0.size)

        def vals(x):
            q, xs, p = problem.unpack(x)
            st, _ = oracle.forward(q, xs, p, so)
            rows = np.concatenate([c.rows(s) for c in cons if c.field == field])
            return st[field].reshape(-1)[rows]
        assert np.abs(fun(x0) - vals(x0)).max() <= 1e-10 * np.abs(vals(x0)).max()
        for k in (0, 7, 15):
            d = np.zeros_like(x0); d[k] = 1e-6 * max(1.0, abs(x0[k]))
            fd = (vals(x0 + d) - vals(x0 - d)) / (2 * d[k])
            assert np.abs(J[:, k] - fd).max() <= 1e-6 * max(np.abs(fd).max(), 1e-3), k
    f0 = problem.fun(x0)[0]
    x = opt.solve(problem)
    assert opt.result.fun < f0                                                # the loss went down ...
    for c in problem.constraints:                                             # ... inside the constraints
        v = c.fun(x)
        assert np.all(v <= c.ub + 1e-4) and np.all(v >= c.lb - 1e-4)       # SLSQP stops at its own feasibility tolerance
    active = problem.constraints[0].fun(x) >= problem.constraints[0].ub - 1e-4
    assert active.any(), "the length bounds were chosen to bind"


def test_hessian_against_second_differences_and_newton_steps():
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS
    from jax_fdm_b200.optimization import NewtonCG, TrustRegionConstrained
    prob = datasets.arch_problem(10)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    keys = [(int(u), int(v)) for u, v in prob.edges]
    loss = LS.Loss(LS.SquaredError([G.EdgeLengthGoal(k, 0.55) for k in keys]), LS.L2Regularizer(1e-6))
    model = eq.EquilibriumModelSparse(tmax=1)
    opt = NewtonCG()
    problem = opt.problem(model, s, loss, prob.q, prob.xyz_fixed, prob.loads, maxiter=30, tol=1e-12, hessian=True)
    x0 = problem.x0
    H = problem.hess(x0)
    assert H.shape == (x0.size, x0.size) and np.abs(H - H.T).max() == 0.0
    oterms = [("SquaredError", [("EdgeLengthGoal", i, 0.55, 1.0) for i in range(len(keys))], 1.0)]

    def f(q):
        st, _ = oracle.forward(q, prob.xyz_fixed, prob.loads, so)
        return og.loss(oterms, st, q=q, l2_alpha=1e-6)

    h = 1e-4
    for i, j in ((0, 0), (2, 3), (5, 9), (9, 9)):
        ei, ej = np.eye(x0.size)[i] * h, np.eye(x0.size)[j] * h
        fd = (f(x0 + ei + ej) - f(x0 + ei - ej) - f(x0 - ei + ej) + f(x0 - ei - ej)) / (4 * h * h)
        assert abs(H[i, j] - fd) <= 1e-5 * max(np.abs(H).max(), 1.0), (i, j, H[i, j], fd)
    f0 = problem.fun(x0)[0]
    opt.solve(problem)
    assert opt.result.fun <= 1e-3 * f0                                        # Newton-CG with the Hessian converges
    tr = TrustRegionConstrained()
    p2 = tr.problem(model, s, loss, prob.q, prob.xyz_fixed, prob.loads, bounds=(-50.0, -0.01), maxiter=60, tol=1e-12, hessian=True)
    tr.solve(p2)
    assert tr.result.fun <= 1e-2 * f0
