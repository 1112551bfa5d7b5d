"""
Two procedures taken from the reference's own tests, run on the CUDA path:

* adjoint routing (tests/test_adjoint_routing.py:97-140): with tmax = 1 the iterative solver is never
  entered; with tmax > 1 and implicit_diff the model goes through `solver_fixedpoint_implicit` and the
  backward pass through `fixed_point_bwd_adjoint` -- spied on as module globals, as the reference does;
* Taylor-remainder convergence (tests/test_taylor_convergence.py:72-113): the second-order remainder
  |f(x0 + eps v) - f(x0) - eps <g, v>| of loss and gradient must fall with slope 2 as eps halves
  (the reference accepts 2.0 +- 0.1).
"""
from unittest import mock

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from jax_fdm_b200 import datasets  # noqa: E402


def T(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=torch.device("cuda", 0))
    return t.requires_grad_(True) if grad else t


def _problem(nx=8):
    prob = datasets.cablenet_problem(nx, seed=4)
    rng = np.random.default_rng(4)
    eload = np.zeros((prob.edges.shape[0], 3))
    eload[:, 2] = -0.05 * rng.uniform(0.5, 1.5, size=prob.edges.shape[0])
    return prob, eload


def _loss_and_grad(model, s, prob, eload, q_np):
    import jax_fdm_b200.equilibrium as eq

    q = T(q_np, True)
    st = model(eq.EquilibriumParametersState(q, T(prob.xyz_fixed), eq.LoadState(T(prob.loads), T(eload), 0.0)), s)
    loss = 0.5 * ((st.xyz - T(prob.xyz0)) ** 2).sum() + 0.05 * (st.forces ** 2).sum()
    loss.backward()
    return float(loss.detach()), q.grad.cpu().numpy()


@pytest.mark.parametrize("tmax", [1, 50])
def test_adjoint_routing(tmax):
    import jax_fdm_b200.equilibrium as eq
    import jax_fdm_b200.equilibrium.models as models_module

    prob, eload = _problem()
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    model = eq.EquilibriumModelSparse(tmax=tmax, eta=1e-10)
    with mock.patch.object(models_module, "solver_fixedpoint_implicit",
                           wraps=models_module.solver_fixedpoint_implicit) as spy_fwd, \
            mock.patch.object(models_module, "fixed_point_bwd_adjoint",
                              wraps=models_module.fixed_point_bwd_adjoint) as spy_adj:
        _loss_and_grad(model, s, prob, eload, prob.q)
    assert spy_fwd.called == (tmax > 1)
    assert spy_adj.called == (tmax > 1)


@pytest.mark.parametrize("tmax", [1, 100])
def test_taylor_remainder_slope_is_two(tmax):
    import jax_fdm_b200.equilibrium as eq

    prob, eload = _problem()
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    model = eq.EquilibriumModelSparse(tmax=tmax, eta=1e-14, tol=1e-14)
    f0, g0 = _loss_and_grad(model, s, prob, eload, prob.q)
    rng = np.random.default_rng(0)
    slopes = []
    for _ in range(3):
        v = rng.uniform(size=prob.q.shape)
        v /= np.linalg.norm(v)
        eps = [1e-2 * 0.5 ** i for i in range(7)]
        rem = np.array([abs(_loss_and_grad(model, s, prob, eload, prob.q + e * v)[0] - f0 - e * np.vdot(g0, v)) for e in eps])
        slopes.append(np.mean(np.log2(rem[:-1] / rem[1:])[:-2]))
    assert abs(np.mean(slopes) - 2.0) <= 0.1, slopes
