"""
`EquilibriumModel` / `EquilibriumModelSparse` -- mirror of `jax_fdm.equilibrium.models`
(models.py:47-1079) for the tmax=1 path, on the B200 kernels.

Two routes through the same arithmetic:
  * method-by-method (`stiffness_matrix`, `load_matrix`, `linearsolve_fn`, `nodes_positions`,
    `equilibrium_state`): each a differentiable op with the reference's name and meaning;
  * fused (`nodes_free_positions`, `__call__`): one autograd node whose backward is the
    adjoint solve + the fused gradient kernel K3 (no K-bar materialised).
Both give the same numbers; tests compare them.
"""

import torch

from . import ops
from .sparse import CSC, sparse_solve
from .states import EquilibriumParametersState, EquilibriumState, LoadState

__all__ = ["EquilibriumModel", "EquilibriumModelSparse", "StiffnessMatrix"]

StiffnessMatrix = CSC


class _Stiffness(torch.autograd.Function):
    """models.py:1064-1079 and its transpose: q_bar = gather/row-sum transpose of K_bar.data."""

    @staticmethod
    def forward(ctx, q, structure):
        ctx.structure = structure
        K, _ = ops.assemble(structure, q, want_rhs=False)
        return K

    @staticmethod
    def backward(ctx, gK):
        s = ctx.structure
        dev = gK.device
        cache = getattr(s, "_torch_kmap", None)
        if cache is None or cache[0].device != dev:
            import numpy as np
            ia = s.index_array
            off = np.flatnonzero(ia.data > 0)
            cache = (torch.as_tensor(off, device=dev),
                     torch.as_tensor(ia.data[off] - 1, device=dev),
                     torch.as_tensor(s.diag_indices, device=dev),
                     torch.as_tensor(np.repeat(np.arange(s.num_free), np.diff(s.diags_indptr)), device=dev),
                     torch.as_tensor(s.diags_indices.astype("int64"), device=dev))
            s._torch_kmap = cache
        off, off_edge, dpos, drow, dedge = cache
        lead = gK.shape[:-1]
        qbar = torch.zeros(lead + (s.num_edges,), dtype=gK.dtype, device=dev)
        qbar.index_add_(-1, off_edge, -gK.index_select(-1, off))
        qbar.index_add_(-1, dedge, gK.index_select(-1, dpos).index_select(-1, drow))
        return qbar, None


class _LoadMatrix(torch.autograd.Function):
    """models.py:828-856, 949-976 and their transposes."""

    @staticmethod
    def forward(ctx, q, xyz_fixed, loads, structure):
        ctx.structure = structure
        ctx.save_for_backward(q, xyz_fixed)
        _, rhs = ops.assemble(structure, q, xyz_fixed, loads, want_K=False)
        return rhs

    @staticmethod
    def backward(ctx, g):
        q, xyz_fixed = ctx.saved_tensors
        s = ctx.structure
        # with lam := g the three outputs of K3 restricted to the rhs terms are exactly the
        # transposes of load_matrix: loads_bar[free] = g, xs_bar = sum q_e g_j, and
        # q_bar_e = -(C_f g)_e . (C_s x_s)_e  -> evaluate K3 on xyz := [0 on free, x_s on fixed].
        xyz0 = ops.positions(s, torch.zeros_like(g), xyz_fixed)
        qbar, pbar, xsbar = ops.adjoint_grads(s, q, xyz0, g.contiguous())
        # q_bar from K3 with x_free = 0 contains -(lam_h-lam_t).(x_h-x_t) with only support
        # coordinates non-zero: for a free-free edge that is 0, for a free-fixed edge it is the
        # wanted  -(C_f g)_e.(C_s x_s)_e.
        return qbar, xsbar, pbar, None


class _State(torch.autograd.Function):
    """models.py:753-794 (K4) and its VJP."""

    @staticmethod
    def forward(ctx, q, xyz, loads, structure):
        ctx.structure = structure
        ctx.save_for_backward(q, xyz)
        v, l, f, r = ops.state(structure, q, xyz, loads)
        return v, l, f, r

    @staticmethod
    def backward(ctx, gv, gl, gf, gr):
        q, xyz = ctx.saved_tensors
        qb, xb, pb = ops.state_vjp(ctx.structure, q, xyz, residuals_cot=gr, lengths_cot=gl,
                                   forces_cot=gf, vectors_cot=gv)
        return qb, xb, pb, None


class _Positions(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_free, xyz_fixed, structure):
        ctx.structure = structure
        return ops.positions(structure, x_free, xyz_fixed)

    @staticmethod
    def backward(ctx, g):
        xf, xs = ops.positions_vjp(ctx.structure, g.contiguous())
        return xf, xs, None


class _FusedFreePositions(torch.autograd.Function):
    """
    x_free = K(q)^{-1} (loads_f - C_f^T Q C_s x_s) as ONE node: forward = K1 + K2, backward =
    K2 (adjoint solve, reusing the Cholesky factor when that solver ran) + K3.
    Same algebra as sparse_solve_bwd + the transposes of stiffness_matrix / load_matrix
    (sparse.py:291-308, models.py:1069-1077, 856, 973-976); SURVEY.md Appendix B items 1-4.
    """

    @staticmethod
    def forward(ctx, q, xyz_fixed, loads, structure, opts):
        K, rhs = ops.assemble(structure, q, xyz_fixed, loads)
        ws = ops.Workspace()
        x = ops.solve(structure, K, rhs, workspace=ws, **opts)
        ctx.structure, ctx.opts, ctx.ws = structure, opts, ws
        ctx.save_for_backward(q, xyz_fixed, x, K)
        return x

    @staticmethod
    def backward(ctx, g):
        q, xyz_fixed, x, K = ctx.saved_tensors
        s = ctx.structure
        # same K, same workspace: the Cholesky factor / coarse inverse of the forward solve is reused
        lam = ops.solve(s, K, g.contiguous(), workspace=ctx.ws, reuse_factor=True, **ctx.opts)
        xyz = ops.positions(s, x, xyz_fixed)
        qbar, pbar, xsbar = ops.adjoint_grads(s, q, xyz, lam)
        return qbar, xsbar, pbar, None, None


class EquilibriumModel:
    """
    FDM model (models.py:47-91).  Only `tmax=1` (one linear FDM step with nodal loads) runs on
    the device path in this build; `tmax>1` raises NotImplementedError.
    The dense variant assembles the same K and solves with the same kernels: the "dense" vs
    "sparse" split of the reference is a storage choice, not different arithmetic.
    """

    def __init__(self, tmax=1, eta=1e-6, is_load_local=False, itersolve_fn=None, iterload_fn=None,
                 implicit_diff=True, verbose=False, solver="auto", tol=0.0, max_iters=0):
        self.tmax = tmax
        self.eta = eta
        self.is_load_local = is_load_local
        self.itersolve_fn = itersolve_fn
        self.iterload_fn = iterload_fn
        self.implicit_diff = implicit_diff
        self.verbose = verbose
        self.solve_opts = {"solver": solver, "tol": tol, "max_iters": max_iters}
        self.linearsolve_fn = lambda K, P: sparse_solve(K, P, **self.solve_opts)

    # ---- edges / nodes (models.py:99-193) -------------------------------------------
    @staticmethod
    def nodes_positions(xyz_free, xyz_fixed, structure):
        return _Positions.apply(xyz_free, xyz_fixed, structure)

    # ---- matrices -------------------------------------------------------------------------
    @staticmethod
    def stiffness_matrix(q, structure):
        return CSC(_Stiffness.apply(q, structure), structure)

    def load_matrix(self, q, xyz_fixed, load_nodes, structure):
        return _LoadMatrix.apply(q, xyz_fixed, load_nodes, structure)

    @staticmethod
    def residual_fixed_matrix(q, xyz_fixed, structure):
        zero = torch.zeros((structure.num_nodes, 3), dtype=q.dtype, device=q.device)
        return -_LoadMatrix.apply(q, xyz_fixed, zero, structure)

    def residual_free_matrix(self, params, xyz_free, structure):
        q, xyz_fixed, load_state = params
        K = self.stiffness_matrix(q, structure)
        P = self.load_matrix(q, xyz_fixed, load_state.nodes, structure)
        return K @ xyz_free - P

    # ---- equilibrium ------------------------------------------------------------------------
    def nodes_free_positions(self, q, xyz_fixed, loads, structure, fused=True):
        if fused:
            return _FusedFreePositions.apply(q, xyz_fixed, loads, structure, self.solve_opts)
        K = self.stiffness_matrix(q, structure)
        P = self.load_matrix(q, xyz_fixed, loads, structure)
        return self.linearsolve_fn(K, P)

    def nodes_equilibrium(self, q, xyz_fixed, loads_nodes, structure):
        return self.nodes_free_positions(q, xyz_fixed, loads_nodes, structure)

    def equilibrium(self, q, xyz_fixed, loads_nodes, structure):
        return self.nodes_equilibrium(q, xyz_fixed, loads_nodes, structure)

    def nodes_load(self, xyz, load_state, structure):
        nodes_load, edges_load, faces_load = load_state
        for extra in (edges_load, faces_load):
            if isinstance(extra, torch.Tensor) and extra.numel() > 1:
                raise NotImplementedError("edge/face tributary loads are outside this build (SURVEY.md §8f)")
        return nodes_load

    def equilibrium_state(self, q, xyz, loads_nodes, structure):
        vectors, lengths, forces, residuals = _State.apply(q, xyz, loads_nodes, structure)
        return EquilibriumState(xyz=xyz, residuals=residuals, lengths=lengths, forces=forces,
                                loads=loads_nodes, vectors=vectors)

    def __call__(self, params: EquilibriumParametersState, structure) -> EquilibriumState:
        q, xyz_fixed, load_state = params
        if not isinstance(load_state, LoadState):
            load_state = LoadState(load_state)
        if self.tmax > 1:
            raise NotImplementedError("tmax > 1 (shape-dependent loads) is outside this build; use tmax=1")
        xyz_free = self.equilibrium(q, xyz_fixed, load_state.nodes, structure)
        xyz = self.nodes_positions(xyz_free, xyz_fixed, structure)
        load_nodes = self.nodes_load(xyz, load_state, structure)
        return self.equilibrium_state(q, xyz, load_nodes, structure)


class EquilibriumModelSparse(EquilibriumModel):
    """models.py:1022-1079: the same model with `linearsolve_fn = sparse_solve`."""
