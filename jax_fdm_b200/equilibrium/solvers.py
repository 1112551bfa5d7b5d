"""
Iterative equilibrium for shape-dependent loads (`tmax > 1`) -- mirror of
`jax_fdm.equilibrium.solvers.fixed_point` (fixed_point.py:139-240, 501-611) on the B200 kernels.

    solver_forward(f, a, x_init, solver_config)            fixed_point.py:139-198
    solver_fixedpoint_implicit(solver, solver_config, f, a, x_init)   fixed_point.py:206-240
    fixed_point_bwd_adjoint(...)                           fixed_point.py:501-604

What differs from the reference, on purpose: every linear solve of the loop and of the adjoint
reuses ONE factorisation / coarse inverse of K held in a device workspace (the reference
re-factorises on every iteration, models.py:591, and caches a SuperLU object in a process-global
dict for the adjoint, sparse.py:386-503); the adjoint system  w - J_P^T K^{-1} w = vec  is solved
by the transposed fixed-point iteration with the same kernels (the reference runs lineax CG on the
normal equations, rtol 1e-6) to a tighter tolerance.
"""

import torch

from . import ops

from typing import NamedTuple

__all__ = ["solver_forward", "solver_fixedpoint_implicit", "fixed_point_bwd_adjoint", "is_solver_fixedpoint",
           # the rest of the reference's solvers namespace (solvers/*.py __all__): importable, see the bottom
           "SolverIterParams", "fixed_point_bwd_adjoint_general", "fixed_point_bwd_fixedpoint",
           "fixed_point_bwd_materialize", "fixed_point_fwd", "solver_anderson", "solver_fixedpoint",
           "is_solver_leastsquares", "is_solver_root_finding", "solver_dogleg", "solver_gauss_newton",
           "solver_levenberg_marquardt", "solver_newton", "solver_jaxopt", "solver_optimistix",
           "nonlinear_bwd", "nonlinear_fwd", "solver_nonlinear_implicit"]


class SolverIterParams(NamedTuple):
    """solvers/types.py: the knobs every iterative solver takes."""
    tmax: int = 100
    eta: float = 1e-6
    implicit_diff: bool = True
    verbose: bool = False


def is_solver_fixedpoint(solver) -> bool:
    """fixed_point.py:48-63."""
    return solver in (None, solver_forward, solver_fixedpoint, solver_anderson)


def is_solver_leastsquares(solver) -> bool:
    """least_squares.py: the Gauss-Newton family."""
    return solver in (solver_gauss_newton, solver_levenberg_marquardt, solver_dogleg)


def is_solver_root_finding(solver) -> bool:
    """root_finding.py."""
    return solver in (solver_newton,)


def solver_forward(f, a, x_init, solver_config):
    """x <- f(a, x) until mean_n |x_prev - x|_2 <= eta or tmax steps (fixed_point.py:139-198)."""
    tmax, eta = solver_config["tmax"], solver_config["eta"]
    x_prev, x = x_init, f(a, x_init)
    steps = 0
    # the distance is a device scalar; one host read per iteration, as the reference's while_loop condition
    while steps < tmax and float(torch.linalg.norm((x_prev - x).detach(), dim=-1).mean()) > eta:
        x_prev, x = x, f(a, x)
        steps += 1
    solver_config["steps"] = steps
    return x


def fixed_point_bwd_adjoint(solve_fn, loads_vjp_x, vec, tol=1e-12, max_iters=200):
    """
    Solve  w - J_P^T K^{-1} w = vec  (fixed_point.py:573-590: A_fn(w) = w - vjp_x(K^{-1} w)) by the
    transposed iteration w <- vec + J_P^T K^{-1} w, and return (w, lam = K^{-1} w).
    `solve_fn(b)` = K^{-1} b with the forward factorisation; `loads_vjp_x(lam)` = J_P^T lam.
    """
    device = getattr(solve_fn, "device_adjoint", None)
    if device is not None:                                 # the same iteration as one CUDA graph (csrc/fixed_point.cu)
        return device(vec)
    w = vec
    scale = float(vec.abs().max())
    lam = solve_fn(w)
    for _ in range(max_iters):
        w_new = vec + loads_vjp_x(lam)
        done = float((w_new - w).abs().max()) <= tol * max(scale, 1e-300)
        w = w_new
        lam = solve_fn(w)
        if done:
            break
    return w, lam


def solver_fixedpoint_implicit(solver, solver_config, f, a, x_init):
    """fixed_point.py:206-240: run `solver`; the implicit backward rule lives in models._FixedPoint.  When the
    iteration function carries a device-resident loop (`f.device_loop`: a CUDA graph with a conditional WHILE node,
    csrc/fixed_point.cu) and the solver is the plain fixed-point iteration, that loop runs instead of the host's."""
    device = getattr(f, "device_loop", None)
    if device is not None and solver in (None, solver_forward):
        return device(x_init, solver_config)
    return (solver or solver_forward)(f, a, x_init, solver_config)


# ---- the remaining names of the reference's solvers namespace ------------------------------------------------------
# The equilibrium path of BASELINE.json runs on `solver_forward` (the reference's default, models.py:89) and its
# implicit adjoint.  The other entries are thin wrappers of third-party iteration libraries (jaxopt, optimistix:
# Anderson acceleration, Gauss-Newton / Levenberg-Marquardt / dogleg, Newton) or alternative backward rules of the
# same fixed point; they are kept importable so code that names them loads, and say what to use when called.
def _outside(name, use):
    def fn(*args, **kwargs):
        raise NotImplementedError(f"{name}: a wrapper of a third-party iteration library in the reference, outside this "
                                  f"build (SURVEY.md §8, out of scope); use {use}")
    fn.__name__ = name
    fn.__doc__ = f"Reference name kept importable; raises NotImplementedError (use {use})."
    return fn


def solver_fixedpoint(f, a, x_init, solver_config):
    """fixed_point.py:66-136 (jaxopt FixedPointIteration): the same iteration as `solver_forward`."""
    return solver_forward(f, a, x_init, solver_config)


def fixed_point_fwd(solver, solver_config, f, a, x_init):
    """fixed_point.py:243-290: forward pass of the implicit rule -> (x*, residuals)."""
    x_star = solver_fixedpoint_implicit(solver, solver_config, f, a, x_init)
    return x_star, (a, x_star)


solver_anderson = _outside("solver_anderson", "solver_forward")
solver_gauss_newton = _outside("solver_gauss_newton", "solver_forward")
solver_levenberg_marquardt = _outside("solver_levenberg_marquardt", "solver_forward")
solver_dogleg = _outside("solver_dogleg", "solver_forward")
solver_newton = _outside("solver_newton", "solver_forward")
solver_jaxopt = _outside("solver_jaxopt", "solver_forward")
solver_optimistix = _outside("solver_optimistix", "solver_forward")
solver_nonlinear_implicit = _outside("solver_nonlinear_implicit", "solver_fixedpoint_implicit")
nonlinear_fwd = _outside("nonlinear_fwd", "solver_fixedpoint_implicit")
nonlinear_bwd = _outside("nonlinear_bwd", "fixed_point_bwd_adjoint")
fixed_point_bwd_adjoint_general = _outside("fixed_point_bwd_adjoint_general", "fixed_point_bwd_adjoint")
fixed_point_bwd_fixedpoint = _outside("fixed_point_bwd_fixedpoint", "fixed_point_bwd_adjoint")
fixed_point_bwd_materialize = _outside("fixed_point_bwd_materialize", "fixed_point_bwd_adjoint")
