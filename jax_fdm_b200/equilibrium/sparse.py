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

import numpy as np
import torch

from . import ops

__all__ = [
    "CSC",
    "SparseSolveResidual",
    "SystemMatrixLHS",
    "SystemMatrixRHS",
    "SystemSolution",
    "blockdiag_matrix_sparse",
    "register_sparse_solver",
    "sparse_solve",
    "sparse_solve_fwd",
    "sparse_solve_bwd",
    "splu_clear",
    "splu_cpu",
    "splu_solve_cpu",
    "spsolve",
    "spsolve_cpu",
    "spsolve_gpu",
    "spsolve_gpu_ravel",
    "spsolve_gpu_stack",
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

    def todense(self):
        """The dense (Nf, Nf) matrix -- (B, Nf, Nf) for a batch -- that the reference's dense model builds as
        `c_free.T @ (q[:, None] * c_free)` (models.py:800-822): `data` scattered onto the static pattern.  For
        inspection, small problems and code written against the dense model; the solve path never forms it."""
        s = self.structure
        flat = getattr(s, "_dense_flat_index", None)
        if flat is None or flat.device != self.data.device:
            ia = s.index_array
            n = ia.shape[0]
            cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(ia.indptr))
            flat = torch.as_tensor(ia.indices.astype(np.int64) * n + cols, device=self.data.device)
            s._dense_flat_index = flat
        n = s.index_array.shape[0]
        out = self.data.new_zeros(self.data.shape[:-1] + (n * n,))
        return out.index_copy(-1, flat, self.data).reshape(self.data.shape[:-1] + (n, n))


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


def _outer_on_pattern(s, a, b, sign=1.0):
    """sign * sum_xyz a[row_k] * b[col_k] for every stored entry k: this library's kernel (fdm_kbar_pattern)."""
    return ops.kbar_pattern(s, a.contiguous(), b.contiguous(), sign)


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
        gd = _outer_on_pattern(s, lam, x, -1.0) if ctx.needs_input_grad[0] else None   # sparse.py:299-306
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
    return CSC(_outer_on_pattern(A.structure, lam, xk.detach(), -1.0), A.structure), lam


# ---- the remaining names of sparse.py:33-51 --------------------------------------------------------------------
SystemMatrixLHS = CSC                      # sparse.py:25-30: type aliases of the solve's operands
SystemMatrixRHS = torch.Tensor
SystemSolution = torch.Tensor
SparseSolveResidual = tuple                # (xk, A, b), sparse.py:239

# one 3-rhs kernel solve serves both GPU formulations of the reference (ravelled 3n x 3n block system,
# sparse.py:54-84; three stacked single-rhs solves, :86-96): same solution, no block-diagonal copy
spsolve_gpu_ravel = spsolve_gpu
spsolve_gpu_stack = spsolve_gpu


def spsolve_cpu(A, b):
    """sparse.py:130-159 (scipy SuperLU through a host callback).  This build has no CPU path by contract."""
    raise RuntimeError("spsolve_cpu: there is no CPU path in jax_fdm_b200 (the solve runs in the sm_100a kernels); "
                       "use spsolve / spsolve_gpu")


def blockdiag_matrix_sparse(A: CSC, num: int = 2, format=None):
    """sparse.py:161-205: `num` copies of A on the diagonal, as (data, indices, indptr, shape) in CSC order.  The
    reference needs it to solve three right-hand sides with a single-rhs solver; the kernels here take the three
    right-hand sides directly, so nothing on the path calls it."""
    import numpy as np

    assert num > 1, "Block diagonal matrix must have at least 2 blocks."
    n, m = A.shape
    assert n == m, "Supported sparse matrices must be square."
    nnz = int(A.indptr[-1])
    indptr = np.concatenate([(A.indptr + i * nnz)[(1 if i else 0):] for i in range(num)])
    indices = np.concatenate([A.indices + i * n for i in range(num)])
    data = torch.cat([A.data] * num, dim=-1)
    out = (data, indices, indptr)
    return format(out, shape=(n * num, n * num)) if format is not None else out + ((n * num, n * num),)


# sparse.py:386-503: a single-entry cache of one factorisation, keyed by a session id.  The reference keeps a scipy
# SuperLU object in a process-global dict on the host; here the entry is the matrix and the DEVICE workspace that
# holds its factor / coarse inverse (the names keep the reference's `_cpu` suffix so call sites import unchanged).
_SPLU_CACHE = {"factorization": None, "session_id": 0}


def splu_clear():
    _SPLU_CACHE["session_id"] = 0
    _SPLU_CACHE["factorization"] = None


def splu_cpu(A: CSC, session_id=None, **opts):
    """sparse.py:394-445: register A for repeated solves; the factorisation happens in the first
    `splu_solve_cpu` and stays in a device workspace."""
    sid = 0 if session_id is None else int(session_id)
    _SPLU_CACHE["factorization"] = {"A": A, "ws": ops.Workspace(), "factored": False, "opts": {**_DEFAULT_OPTS, **opts}}
    _SPLU_CACHE["session_id"] = sid
    return sid


def splu_solve_cpu(session_id, b):
    """sparse.py:448-503: solve with the cached factorisation; a session mismatch or an empty cache raises."""
    sid = int(session_id)
    if _SPLU_CACHE["session_id"] != sid:
        raise ValueError(f"Session mismatch: expected {sid}, but cache has {_SPLU_CACHE['session_id']}. "
                         "Make sure to call splu_cpu before splu_solve_cpu in the same session.")
    f = _SPLU_CACHE["factorization"]
    if f is None:
        raise ValueError("No factorization found in cache. Call splu_cpu first.")
    A = f["A"]
    x = ops.solve(A.structure, A.data.detach(), b.contiguous(), workspace=f["ws"], reuse_factor=f["factored"],
                  nan_on_failure=True, **f["opts"])
    f["factored"] = True
    return x
