"""
Differentiable sparse linear solve -- mirror of `jax_fdm.equilibrium.sparse`
(sparse.py:54-308) on the B200 kernels.

`sparse_solve(A, b)` is the reference's `jax.custom_vjp` function restated as a
`torch.autograd.Function` with the same residuals (x, A, b) and the same cotangent
structure: the cotangent of `A` is a CSC whose `.data` alone is populated
(`Kbar.data[k] = -sum_xyz lam[row_k] * x[col_k]`), the cotangent of `b` is `lam`
(sparse.py:291-308).  Backends are registered through the same seam
(`register_sparse_solver`, sparse.py:167-199); only "gpu" exists here -- there is no CPU path.
"""

from typing import Callable, NamedTuple

import torch

from . import ops

__all__ = [
    "CSC",
    "register_sparse_solver",
    "sparse_solve",
    "sparse_solve_fwd",
    "sparse_solve_bwd",
    "spsolve",
    "spsolve_gpu",
]


class CSC(NamedTuple):
    """Stiffness matrix on the structure's pattern: `data` is a tensor, the pattern is static."""

    data: torch.Tensor     # (nnz,) or (B, nnz)
    structure: object      # EquilibriumStructureSparse (carries indices / indptr)

    @property
    def indices(self):
        return self.structure.index_array.indices

    @property
    def indptr(self):
        return self.structure.index_array.indptr

    @property
    def shape(self):
        return self.structure.index_array.shape

    def __matmul__(self, x):
        return _SpMM.apply(self.data, x, self.structure)


_DEFAULT_OPTS = {"solver": "auto", "tol": 0.0, "max_iters": 0}


def spsolve_gpu(A: CSC, b, **opts):
    """Solve on the B200 (replaces spsolve_gpu_ravel, sparse.py:54-84: no 3x block-diagonal copy)."""
    o = {**_DEFAULT_OPTS, **opts}
    return ops.solve(A.structure, A.data, b, **o)


def register_sparse_solver(solvers: dict) -> Callable:
    """sparse.py:167-192.  The active backend is the CUDA device; anything else is an error."""
    backend = "gpu" if torch.cuda.is_available() else "cpu"
    fn = solvers.get(backend)
    if not fn:
        raise ValueError(f"Default backend {backend} does not support a sparse solver")
    return fn


_solvers = {"gpu": spsolve_gpu}


def spsolve(A, b, **opts):
    return register_sparse_solver(_solvers)(A, b, **opts)


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, x, structure):
        ctx.structure = structure
        ctx.save_for_backward(data, x)
        return ops.spmm(structure, data, x)

    @staticmethod
    def backward(ctx, g):
        data, x = ctx.saved_tensors
        s = ctx.structure
        gd = _outer_on_pattern(s, g, x) if ctx.needs_input_grad[0] else None
        gx = ops.spmm(s, data, g.contiguous()) if ctx.needs_input_grad[1] else None   # K symmetric
        return gd, gx, None


def _pattern_rows_cols(s, device):
    cache = getattr(s, "_torch_pattern", None)
    if cache is None or cache[0].device != device:
        import numpy as np
        ia = s.index_array
        cols = np.repeat(np.arange(s.num_free), np.diff(ia.indptr))
        cache = (torch.as_tensor(ia.indices.astype("int64"), device=device),
                 torch.as_tensor(cols.astype("int64"), device=device))
        s._torch_pattern = cache
    return cache


def _outer_on_pattern(s, a, b):
    """sum_xyz a[row_k] * b[col_k] for every stored entry k (API-parity path, not the hot path)."""
    rows, cols = _pattern_rows_cols(s, a.device)
    return (a.index_select(-2, rows) * b.index_select(-2, cols)).sum(-1)


class _SparseSolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, b, structure, opts):
        x = ops.solve(structure, data, b, nan_on_failure=True, **opts)
        ctx.structure, ctx.opts = structure, opts
        ctx.save_for_backward(x, data, b)               # residuals (xk, A, b), sparse.py:264
        return x

    @staticmethod
    def backward(ctx, g):
        x, data, b = ctx.saved_tensors
        s = ctx.structure
        lam = ops.solve(s, data, g.contiguous(), **ctx.opts)      # sparse.py:296 (A symmetric)
        gd = -_outer_on_pattern(s, lam, x) if ctx.needs_input_grad[0] else None   # sparse.py:299-306
        gb = lam if ctx.needs_input_grad[1] else None
        return gd, gb, None, None


def sparse_solve(A: CSC, b, **opts):
    """sparse.py:207-230: x = A^{-1} b, reverse-mode differentiable w.r.t. A.data and b."""
    o = {**_DEFAULT_OPTS, **opts}
    return _SparseSolve.apply(A.data, b, A.structure, o)


def sparse_solve_fwd(A: CSC, b, **opts):
    """sparse.py:242-264."""
    xk = sparse_solve(A, b, **opts)
    return xk, (xk, A, b)


def sparse_solve_bwd(res, g, **opts):
    """sparse.py:267-308 -> (CSC cotangent with only .data populated, b cotangent)."""
    xk, A, b = res
    o = {**_DEFAULT_OPTS, **opts}
    lam = ops.solve(A.structure, A.data.detach(), g.contiguous(), **o)
    return CSC(-_outer_on_pattern(A.structure, lam, xk.detach()), A.structure), lam
