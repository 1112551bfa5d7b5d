"""
`fdm()` on arrays -- the COMPAS-free front door of `jax_fdm.equilibrium.fdm` (fdm.py:105-174).

The reference's `fdm(datastructure)` reads q / fixed coordinates / loads off a COMPAS
network or mesh, calls `model(params, structure)` and writes the state back
(fdm.py:99, 525-613).  COMPAS is a modelling layer outside the hot path; here the same call
takes the arrays the datastructure would have yielded.
"""

import numpy as np
import torch

from .models import EquilibriumModel, EquilibriumModelSparse
from .states import EquilibriumParametersState, LoadState
from .structures import EquilibriumStructure, EquilibriumStructureSparse

__all__ = ["fdm", "constrained_fdm", "model_from_sparsity", "structure_from_arrays", "arrays_validate"]


def arrays_validate(edges, supports, q):
    """fdm.py:491-522: at least one support, one edge, no zero force density."""
    assert np.count_nonzero(supports) > 0, "The FD datastructure has no supports"
    assert len(edges) > 0, "The FD datastructure has no edges"
    assert not np.any(np.asarray(q) == 0.0), "The FD datastructure has zero force densities"


def model_from_sparsity(sparse=True, **kwargs):
    """fdm.py:311-362."""
    return (EquilibriumModelSparse if sparse else EquilibriumModel)(**kwargs)


def structure_from_arrays(nodes, edges, supports, sparse=True, device=0):
    """fdm.py:365-442 on arrays; both flavours carry the sparse helpers (same plan)."""
    return EquilibriumStructureSparse(nodes, edges, supports, device=device)


def fdm(nodes, edges, supports, q, xyz_fixed, loads, sparse=True, tmax=1, eta=1e-6,
        is_load_local=False, implicit_diff=True, verbose=False, device=0, structure=None, **solve_opts):
    """
    One forward force-density solve.  Returns (EquilibriumState of torch tensors, structure).
    """
    arrays_validate(edges, supports, q.detach().cpu().numpy() if isinstance(q, torch.Tensor) else q)
    model = model_from_sparsity(sparse, tmax=tmax, eta=eta, is_load_local=is_load_local,
                                implicit_diff=implicit_diff, verbose=verbose, **solve_opts)
    if structure is None:
        structure = structure_from_arrays(nodes, edges, supports, sparse, device)
    dev = torch.device("cuda", device)
    as_t = lambda a: a.to(dev, torch.float64) if isinstance(a, torch.Tensor) else \
        torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)
    params = EquilibriumParametersState(as_t(q), as_t(xyz_fixed), LoadState(as_t(loads), 0.0, 0.0))
    return model(params, structure), structure


def constrained_fdm(nodes, edges, supports, q, xyz_fixed, loads, optimizer, loss, bounds=(None, None), maxiter=100,
                    tol=1e-6, sparse=True, tmax=1, eta=1e-6, is_load_local=False, device=0, structure=None,
                    callback=None, q_groups=None, **solve_opts):
    """
    `constrained_fdm` on arrays (fdm.py:177-308) with the force densities as the parameters: minimise `loss`
    over q within `bounds` with `optimizer` (jax_fdm_b200.optimization: LBFGSB, SLSQP, ...), then one forward
    solve at the optimum.  Returns (EquilibriumState, structure, q_opt).  Every optimiser iteration is one
    forward + adjoint pass on the device; constraints (which the reference routes through the dense model and
    forward mode, fdm.py:266-271) are outside this build.
    """
    q0 = q.detach().cpu().numpy() if isinstance(q, torch.Tensor) else np.asarray(q, dtype=np.float64)
    arrays_validate(edges, supports, q0)
    model = model_from_sparsity(sparse, tmax=tmax, eta=eta, is_load_local=is_load_local, **solve_opts)
    if structure is None:
        structure = structure_from_arrays(nodes, edges, supports, sparse, device)
    problem = optimizer.problem(model, structure, loss, q0, xyz_fixed, loads, bounds=bounds, maxiter=maxiter, tol=tol,
                                callback=callback, q_groups=q_groups)
    q_opt = problem.unpack(optimizer.solve(problem))[0]      # (E,): grouped parameters expanded to their edges
    state, _ = fdm(nodes, edges, supports, q_opt, xyz_fixed, loads, sparse=sparse, tmax=tmax, eta=eta,
                   is_load_local=is_load_local, device=device, structure=structure, **solve_opts)
    return state, structure, q_opt
