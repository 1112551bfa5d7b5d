"""
GPU: BASELINE config 2 as the reference specifies it (tests/test_arch_benchmark.py:53-64, 100-166): the planar
arch with n_v in {5, 50, 100, 500} vertices, ONE grouped force density (EdgeGroupForceDensityParameter ->
`q_groups`), PredictionError on NetworkLoadPathGoal, L-BFGS-B with bounds (-1000, -1e-3), maxiter 100, tol 1e-9.
  * forward state and the q-gradient of the load path against the oracle at q = -10 (1e-10 / 1e-8),
  * the optimised rise against sqrt(3) d / 4 (0.1 %) and the load path against rho d^2 / sqrt(3)
    (1.1 % at n_v = 100, 0.3 % at n_v = 500), the reference's own tolerances,
  * the load-path error falling monotonically with n_v.
"""
from math import sqrt

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import goals as og  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402

SPAN, RHO = 10.0, 1.0
RISE = sqrt(3.0) * SPAN / 4.0
LOADPATH = RHO * SPAN ** 2 / sqrt(3.0)
TERMS = [("PredictionError", [("NetworkLoadPathGoal", 0, 0.0, 1.0)], 1.0)]


def _setup(nv):
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS
    prob = datasets.arch_benchmark_problem(nv)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    loss = LS.Loss(LS.PredictionError([G.NetworkLoadPathGoal()]))
    model = eq.EquilibriumModelSparse(tmax=1)
    return eq, prob, s, loss, model


@pytest.mark.parametrize("nv", [5, 50, 100, 500])
def test_arch_forward_and_q_gradient_vs_oracle(nv):
    eq, prob, s, loss, model = _setup(nv)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda:0")
    q = T(prob.q).requires_grad_(True)
    params = eq.EquilibriumParametersState(q, T(prob.xyz_fixed), eq.LoadState(T(prob.loads), 0.0, 0.0))
    state = model(params, s)
    val = loss(params, model, s)
    val.backward()

    def cot(st, xf):
        sg = np.sign(st["lengths"] * st["forces"])
        return {"lengths": sg * st["forces"], "forces": sg * st["lengths"]}

    ref = oracle.forward_adjoint(prob.q, prob.xyz_fixed, prob.loads, so, cot)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    assert rel(state.xyz.detach().cpu().numpy(), ref["state"]["xyz"]) <= 1e-10
    assert rel(state.forces.detach().cpu().numpy().reshape(-1), ref["state"]["forces"].reshape(-1)) <= 1e-10
    assert abs(float(val.detach()) - og.loss(TERMS, ref["state"])) <= 1e-10 * abs(float(val.detach()))
    assert rel(q.grad.cpu().numpy(), ref["q_bar"]) <= 1e-8


def _optimise(nv):
    from jax_fdm_b200.optimization import LBFGSB
    eq, prob, s, loss, model = _setup(nv)
    opt = LBFGSB()
    problem = opt.problem(model, s, loss, prob.q, prob.xyz_fixed, prob.loads, bounds=(-1000.0, -1e-3), maxiter=100,
                          tol=1e-9, q_groups=[np.arange(prob.q.size)])
    assert problem.x0.shape == (1,)                        # one grouped parameter
    x = opt.solve(problem)
    q, xs, p = problem.unpack(x)
    assert q.shape == prob.q.shape and np.all(q == x[0])
    st, _ = eq.fdm(prob.nodes, prob.edges, prob.supports, q, xs, p, structure=s)
    rise = float(np.abs(st.xyz.cpu().numpy()[:, 1]).max())
    return rise, float(opt.result.fun)


@pytest.mark.parametrize("nv,rise_tol,lp_tol", [(100, 1e-3, 1.1e-2), (500, 1e-3, 3e-3)])
def test_arch_matches_analytical(nv, rise_tol, lp_tol):
    rise, lp = _optimise(nv)
    assert abs(rise - RISE) / RISE < rise_tol
    assert abs(lp - LOADPATH) / LOADPATH < lp_tol


def test_arch_error_decreases_with_discretization():
    errs = [abs(_optimise(nv)[1] - LOADPATH) / LOADPATH for nv in (5, 50, 100, 500)]
    assert all(b < a for a, b in zip(errs, errs[1:])), errs
