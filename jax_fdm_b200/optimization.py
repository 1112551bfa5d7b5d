"""
The seam between the device path and `scipy.optimize` (mirror of the reference's
optimization/optimizers/optimizer.py:283-415, 453-515: `problem()` builds `loss_and_grad(x)` over the flat
parameter vector, `solve()` hands it to `scipy.optimize.minimize(jac=True)`), for the common case where the
parameters are the force densities: every iteration is one forward + adjoint pass on the GPU (model, fused
goal / loss kernel, state VJP, adjoint solve with the stored factor), and only the scalar loss and the
gradient `q_bar` cross to the host.
"""
import numpy as np
import torch

from .equilibrium import EquilibriumParametersState, LoadState

__all__ = ["OptProblem", "Optimizer", "LBFGSB", "SLSQP", "TrustRegionConstrained", "optimize_q"]

F64 = torch.float64


class OptProblem:
    """optimizer.py:47-85: what scipy.optimize.minimize needs."""

    def __init__(self, fun, x0, bounds, method, tol, maxiter, callback=None):
        self.fun, self.x0, self.bounds, self.method = fun, x0, bounds, method
        self.tol, self.maxiter, self.callback = tol, maxiter, callback

    def to_kwargs(self):
        return dict(fun=self.fun, jac=True, x0=self.x0, method=self.method, bounds=self.bounds, tol=self.tol,
                    callback=self.callback, options={"maxiter": self.maxiter})


class Optimizer:
    """optimizer.py:88-515 for parameters = q (lower/upper bounds per edge or scalars)."""
    name = None

    def __init__(self, disp=False):
        self.disp = disp

    def problem(self, model, structure, loss, q0, xyz_fixed, loads, bounds=(None, None), maxiter=100, tol=1e-6,
                callback=None):
        dev = torch.device("cuda", structure.device) if isinstance(structure.device, int) else torch.device(structure.device)
        xs = torch.as_tensor(np.ascontiguousarray(xyz_fixed, dtype=np.float64), device=dev)
        p = torch.as_tensor(np.ascontiguousarray(loads, dtype=np.float64), device=dev)
        load_state = p if isinstance(loads, LoadState) else LoadState(p, 0.0, 0.0)

        def loss_and_grad(x):                            # optimizer.py:365-370: jit(value_and_grad(loss))
            q = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=dev).requires_grad_(True)
            val = loss(EquilibriumParametersState(q, xs, load_state), model, structure)
            val.backward()
            return float(val.item()), q.grad.detach().cpu().numpy()

        x0 = np.ascontiguousarray(q0, dtype=np.float64)
        val, grad = loss_and_grad(x0)                    # optimizer.py:372-380: warm start, no NaN allowed
        assert np.isfinite(val) and np.all(np.isfinite(grad)), "the loss or its gradient is not finite at x0"
        lo, hi = bounds
        lo = np.broadcast_to(-np.inf if lo is None else lo, x0.shape)
        hi = np.broadcast_to(np.inf if hi is None else hi, x0.shape)
        return OptProblem(loss_and_grad, x0, list(zip(lo, hi)), self.name, tol, maxiter, callback)

    def solve(self, opt_problem):
        from scipy.optimize import minimize
        res = minimize(**opt_problem.to_kwargs())
        self.result = res
        return res.x


class LBFGSB(Optimizer):
    """optimizers/gradient_based.py: L-BFGS-B."""
    name = "L-BFGS-B"


class SLSQP(Optimizer):
    """optimizers/gradient_based.py: SLSQP."""
    name = "SLSQP"


class TrustRegionConstrained(Optimizer):
    """optimizers/second_order.py: trust-constr (gradient only here)."""
    name = "trust-constr"


def optimize_q(model, structure, loss, q0, xyz_fixed, loads, optimizer=None, **kw):
    """constrained_fdm (fdm.py:170-330) on arrays, parameters = q: returns (q_opt, scipy result)."""
    opt = optimizer or LBFGSB()
    x = opt.solve(opt.problem(model, structure, loss, q0, xyz_fixed, loads, **kw))
    return x, opt.result
