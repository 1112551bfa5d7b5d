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
from .states import DatastructureState, EquilibriumParametersState, EquilibriumState, LoadState, _is_mesh
from .structures import (EquilibriumMeshStructure, EquilibriumMeshStructureSparse, EquilibriumStructure,
                         EquilibriumStructureSparse)

__all__ = ["fdm", "constrained_fdm", "model_from_sparsity", "structure_from_arrays", "arrays_validate",
           "structure_from_datastructure", "structure_from_network", "structure_from_mesh", "datastructure_state",
           "datastructure_validate", "datastructure_updated", "datastructure_update", "datastructure_edges_update",
           "datastructure_nodes_update"]


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
    state = model(params, structure)
    _check_equilibrium(state, structure)
    return state, structure


def _check_equilibrium(state, structure, rtol=1e-6):
    """The front door fails LOUDLY where the kernels report through NaN or cannot certify: a solve whose status
    was not OK (NaN positions: K singular, or indefinite -- mixed-sign q -- on a band too wide for the signed
    factorisation, which the reference's pivoted LU would still solve), and a solution whose free-node residuals
    are not small (the unpivoted signed factorisation of an indefinite K can lose accuracy without a flag).  One
    device-to-host read; `fdm()` hands the state back to host code anyway."""
    xyz = state.xyz.detach()
    if not bool(torch.isfinite(xyz).all()):
        raise FloatingPointError("fdm: the equilibrium solve failed (non-finite positions): K(q) is singular, or it is "
                                 "indefinite (mixed-sign q) on a structure whose band is too wide for this build's signed "
                                 "factorisation (half bandwidth > 30); the iterative solvers need a definite K")
    free = torch.as_tensor(structure.indices_free, device=xyz.device)
    res = state.residuals.detach().index_select(-2, free).abs().amax()
    scale = torch.maximum(state.loads.detach().abs().amax(), (state.forces.detach().abs().amax()))
    if float(res) > rtol * max(float(scale), 1e-300):
        raise FloatingPointError(f"fdm: free-node residual {float(res):.3e} against loads / forces of {float(scale):.3e}: "
                                 "the solve is inaccurate (indefinite K factored without pivoting?)")


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


# ---- the COMPAS-facing helpers of fdm.py:365-700 ------------------------------------------------------------------
# COMPAS (FDNetwork / FDMesh) is a modelling layer above the path and is not installed here; these take ANY object
# with the accessors the reference calls (nodes() / edges() / is_node_support(), edges_forcedensities(),
# node_attribute(), ...), so a COMPAS datastructure works unchanged and the tests use a small array-backed stand-in.
def structure_from_network(network, sparse=True, device=0):
    """fdm.py:396-417."""
    return (EquilibriumStructureSparse if sparse else EquilibriumStructure).from_network(network, device=device)


def structure_from_mesh(mesh, sparse=True, device=0):
    """fdm.py:420-441."""
    return (EquilibriumMeshStructureSparse if sparse else EquilibriumMeshStructure).from_mesh(mesh, device=device)


def structure_from_datastructure(datastructure, sparse=True, device=0):
    """fdm.py:365-393."""
    return (structure_from_mesh if _is_mesh(datastructure) else structure_from_network)(datastructure, sparse, device)


def datastructure_state(datastructure, sparse=True, device=0):
    """fdm.py:444-488: structure, parameters and the STORED equilibrium state of a datastructure, without solving."""
    structure = structure_from_datastructure(datastructure, sparse, device)
    dev = None if device is None else torch.device("cuda", device)
    parameters = EquilibriumParametersState.from_datastructure(datastructure, device=dev)
    eq_state = EquilibriumState.from_datastructure(datastructure, structure, device=dev)
    return DatastructureState(eq_state, structure, parameters)


def datastructure_validate(datastructure):
    """fdm.py:491-522."""
    assert datastructure.number_of_supports() > 0, "The FD datastructure has no supports"
    assert datastructure.number_of_edges() > 0, "The FD datastructure has no edges"
    has_fd = np.abs(np.array(datastructure.edges_forcedensities())) > 0.0
    assert np.all(has_fd), f"The FD datastructure has {int(np.sum(~has_fd))} edges with zero force density"
    if _is_mesh(datastructure):
        assert datastructure.number_of_vertices() > 0, "The FD datastructure has no vertices"
    else:
        assert datastructure.number_of_nodes() > 0, "The FD datastructure has no nodes"


def datastructure_updated(datastructure, eq_state, params, use_loadsfromparams=False):
    """fdm.py:525-563: a copy of the datastructure carrying the equilibrium state."""
    datastructure = datastructure.copy()
    datastructure_update(datastructure, eq_state, params, use_loadsfromparams)
    return datastructure


def datastructure_update(datastructure, eq_state, params, use_loadsfromparams=False):
    """fdm.py:566-613: write the state onto the datastructure in place (one device-to-host copy per array)."""
    host = lambda t: t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    loads = host(params.loads.nodes) if use_loadsfromparams else host(eq_state.loads)
    datastructure_edges_update(datastructure, (host(eq_state.lengths), host(eq_state.forces), host(params.q)))
    datastructure_nodes_update(datastructure, (host(eq_state.xyz), host(eq_state.residuals), loads))


def datastructure_edges_update(datastructure, eqstate_edges):
    """fdm.py:616-637: per-edge length, force and force density."""
    lengths, forces, forcedensities = (np.asarray(a, dtype=np.float64).reshape(-1) for a in eqstate_edges)
    for idx, edge in datastructure.index_edge().items():
        datastructure.edge_attribute(edge, name="length", value=float(lengths[idx]))
        datastructure.edge_attribute(edge, name="force", value=float(forces[idx]))
        datastructure.edge_attribute(edge, name="q", value=float(forcedensities[idx]))


def datastructure_nodes_update(datastructure, eqstate_nodes):
    """fdm.py:640-672: per-node (vertex) coordinates, residuals and loads."""
    xyz, residuals, loads = (np.asarray(a, dtype=np.float64).reshape(-1, 3) for a in eqstate_nodes)
    mesh = _is_mesh(datastructure)
    update = datastructure.vertex_attribute if mesh else datastructure.node_attribute
    index_key = datastructure.index_vertex if mesh else datastructure.index_node
    for idx, key in index_key().items():
        for names, row in ((("x", "y", "z"), xyz[idx]), (("rx", "ry", "rz"), residuals[idx]), (("px", "py", "pz"), loads[idx])):
            for name, value in zip(names, row):
                update(key, name=name, value=float(value))
