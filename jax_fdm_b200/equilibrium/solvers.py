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

__all__ = ["solver_forward", "solver_fixedpoint_implicit", "fixed_point_bwd_adjoint", "is_solver_fixedpoint"]


def is_solver_fixedpoint(solver) -> bool:
    return solver in (None, solver_forward)


def solver_forward(f, a, x_init, solver_config):
    """x <- f(a, x) until mean_n |x_prev - x|_2 <= eta or tmax steps (fixed_point.py:139-198)."""
    tmax, eta = solver_config["tmax"], solver_config["eta"]
    x_prev, x = x_init, f(a, x_init)
    steps = 0
    # the distance is a device scalar; one host read per iteration, as the reference's while_loop condition
    while steps < tmax and float(torch.linalg.norm(x_prev - x, dim=-1).mean()) > eta:
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
    """fixed_point.py:206-240: run `solver`; the implicit backward rule lives in models._FixedPoint."""
    return (solver or solver_forward)(f, a, x_init, solver_config)
