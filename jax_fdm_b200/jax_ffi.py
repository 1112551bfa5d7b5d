"""
`jax.ffi` registration of the B200 kernels as XLA custom calls -- the binding the reference side
adds to run `jax_fdm.equilibrium` on this library.

JAX is NOT installed in this image (nor on the GPU box), so this module is source only here:
importing it raises ImportError with that explanation, and nothing else in the package depends
on it.  Where JAX and `lib/libfdm_b200_xla.so` (csrc/ffi/xla_ffi.cc) exist it provides

    Ops(structure)             custom-call wrappers bound to one structure: assemble, solve (+ the
                               factor-reusing solve), solve_from_q, spmm, positions, state,
                               adjoint_grads, edge_loads(+vjp), loss_sqdist, and
                               `nodes_free_positions`: the reference's jax.custom_vjp shape
                               (sparse.py:207-308) with the forward and the adjoint solve as custom
                               calls, the forward's workspace (Cholesky factor / coarse inverse)
                               carried as a residual
    spsolve_b200(structure)    -> a `(A: CSC, b) -> x` callable for the reference's own seam:
        jax_fdm.equilibrium.sparse.register_sparse_solver({"gpu": spsolve_b200(structure)})  (sparse.py:167-199)
"""

import ctypes
import os

try:
    import jax
    import jax.numpy as jnp
except ImportError as exc:  # pragma: no cover - the image has no jax
    raise ImportError("jax_fdm_b200.jax_ffi needs jax (absent from this image); the torch/ctypes "
                      "mirror in jax_fdm_b200.equilibrium drives the same C ABI") from exc

from . import _lib as L

_XLA_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libfdm_b200_xla.so")
_TARGETS = ["assemble", "solve", "solve_reuse", "solve_from_q", "spmm", "positions", "state", "state_vjp",
            "adjoint_grads", "edge_loads", "edge_loads_vjp", "loss_sqdist", "state_jvp", "kbar_pattern", "stiffness_vjp",
            "face_loads"]
_registered = False


def _register_targets():
    global _registered
    if _registered:
        return
    if not os.path.exists(_XLA_LIB):
        raise ImportError(f"{_XLA_LIB} not built (see the build line at the top of csrc/ffi/xla_ffi.cc)")
    lib = ctypes.CDLL(_XLA_LIB)
    for t in _TARGETS:
        jax.ffi.register_ffi_target(f"fdm_b200_{t}", jax.ffi.pycapsule(getattr(lib, f"fdm_b200_{t}")), platform="CUDA")
    _registered = True


def _f64(*shape):
    return jax.ShapeDtypeStruct(shape, jnp.float64)


class Ops:
    """Custom-call wrappers bound to one structure (its native plan handle is a static attribute)."""

    def __init__(self, structure):
        _register_targets()
        s = self.s = structure
        self.plan = int(s.plan.value)
        self.E, self.N, self.Nf, self.Ns, self.nnz = s.num_edges, s.num_nodes, s.num_free, s.num_supports, s.nnz

    def _call(self, name, outs, *args, **attrs):
        return jax.ffi.ffi_call(f"fdm_b200_{name}", outs, vmap_method="broadcast_all")(*args, plan=self.plan, **attrs)

    def assemble(self, q, xyz_fixed, loads):
        lead = q.shape[:-1]
        return self._call("assemble", (_f64(*lead, self.nnz), _f64(*lead, self.Nf, 3)), q, xyz_fixed, loads)

    def solve(self, Kdata, rhs, workspace=None, solver=L.SOLVER_AUTO, tol=0.0, max_iters=0):
        lead = rhs.shape[:-2]
        batch = 1
        for d in lead:
            batch *= d
        nbytes = int(L.lib().fdm_workspace_bytes(self.s.plan, batch, solver))
        outs = (_f64(*lead, self.Nf, 3), _f64(batch, L.INFO_STRIDE), jax.ShapeDtypeStruct((max(nbytes, 16),), jnp.uint8))
        if workspace is None:
            return self._call("solve", outs, Kdata, rhs, solver=solver, max_iters=max_iters, tol=tol, reuse_factor=0)
        # the backward pass: a second target with the workspace as a third OPERAND, aliased onto the workspace result
        return jax.ffi.ffi_call("fdm_b200_solve_reuse", outs, vmap_method="broadcast_all", input_output_aliases={2: 2})(
            Kdata, rhs, workspace, plan=self.plan, solver=solver, max_iters=max_iters, tol=tol)

    def solve_from_q(self, q, xyz_fixed, loads, solver=L.SOLVER_CHOLESKY_BATCHED):
        """K1 + K2 fused (fdm_solve_from_q): -> (x_free, info, workspace)."""
        lead = q.shape[:-1]
        batch = 1
        for d in lead:
            batch *= d
        nbytes = int(L.lib().fdm_workspace_bytes(self.s.plan, batch, solver))
        outs = (_f64(*lead, self.Nf, 3), _f64(batch, L.INFO_STRIDE), jax.ShapeDtypeStruct((max(nbytes, 16),), jnp.uint8))
        return self._call("solve_from_q", outs, q, xyz_fixed, loads, solver=solver)

    def spmm(self, Kdata, X):
        return self._call("spmm", _f64(*X.shape), Kdata, X)

    def edge_loads(self, xyz, loads_nodes, edges_load, is_local=False):
        return self._call("edge_loads", _f64(*xyz.shape), xyz, loads_nodes, edges_load, is_local=int(is_local))

    def edge_loads_vjp(self, xyz, edges_load, cot, is_local=False):
        lead = xyz.shape[:-2]
        return self._call("edge_loads_vjp", (_f64(*xyz.shape), _f64(*lead, self.E, 3)), xyz, edges_load, cot,
                          is_local=int(is_local))

    def face_loads(self, faces, edges_faces, xyz, loads_nodes, faces_load, is_local=False):
        """nodes_load with tributary face loads (mesh structures): returns (centroids, gl scratch, loads)."""
        lead = xyz.shape[:-2]
        F, D = faces.shape
        outs = (_f64(*lead, F, 3), _f64(*lead, F, D, 3), _f64(*xyz.shape))
        return self._call("face_loads", outs, faces, edges_faces, xyz, loads_nodes, faces_load, is_local=int(is_local))

    def state_jvp(self, q, xyz, q_dot, xyz_dot, loads_dot):
        lead = xyz_dot.shape[:-2]
        outs = (_f64(*lead, self.E, 3), _f64(*lead, self.E), _f64(*lead, self.E), _f64(*lead, self.N, 3))
        return self._call("state_jvp", outs, q, xyz, q_dot, xyz_dot, loads_dot)

    def kbar_pattern(self, a, b, sign=-1.0):
        """Kbar.data = sign * a[row] . b[col] on the pattern (sparse.py:299-306)."""
        return self._call("kbar_pattern", _f64(*a.shape[:-2], self.nnz), a, b, sign=float(sign))

    def stiffness_vjp(self, kbar_data):
        return self._call("stiffness_vjp", _f64(*kbar_data.shape[:-1], self.E), kbar_data)

    def loss_sqdist(self, x, x0):
        return self._call("loss_sqdist", (_f64(*x.shape), _f64(*x.shape[:-2])), x, x0)

    def positions(self, x_free, xyz_fixed):
        return self._call("positions", _f64(*x_free.shape[:-2], self.N, 3), x_free, xyz_fixed)

    def state(self, q, xyz, loads):
        lead = xyz.shape[:-2]
        outs = (_f64(*lead, self.E, 3), _f64(*lead, self.E, 1), _f64(*lead, self.E, 1), _f64(*lead, self.N, 3))
        return self._call("state", outs, q, xyz, loads)

    def adjoint_grads(self, q, xyz, lam):
        lead = xyz.shape[:-2]
        outs = (_f64(*lead, self.E), _f64(*lead, self.N, 3), _f64(*lead, self.Ns, 3))
        return self._call("adjoint_grads", outs, q, xyz, lam)

    # ---- the differentiable solve with the reference's custom_vjp shape (sparse.py:207-308) ----
    def nodes_free_positions(self, q, xyz_fixed, loads):
        ops = self

        @jax.custom_vjp
        def f(q, xyz_fixed, loads):
            K, rhs = ops.assemble(q, xyz_fixed, loads)
            return ops.solve(K, rhs)[0]

        def fwd(q, xyz_fixed, loads):
            K, rhs = ops.assemble(q, xyz_fixed, loads)
            x, _, ws = ops.solve(K, rhs)
            return x, (q, xyz_fixed, x, K, ws)

        def bwd(res, g):
            q, xyz_fixed, x, K, ws = res
            lam = ops.solve(K, g, workspace=ws)[0]            # same K: factor / coarse inverse reused
            xyz = ops.positions(x, xyz_fixed)
            q_bar, loads_bar, xs_bar = ops.adjoint_grads(q, xyz, lam)
            return q_bar, xs_bar, loads_bar

        f.defvjp(fwd, bwd)
        return f(q, xyz_fixed, loads)


def spsolve_b200(structure, **solve_kw):
    """A solver for the reference's backend seam (sparse.py:167-199): `(A, b) -> x` with A the reference's CSC
    StiffnessMatrix on this structure's pattern (only A.data is read) and b of shape (Nf, 3)."""
    ops = Ops(structure)

    def solve(A, b):
        return ops.solve(A.data, b, **solve_kw)[0]

    return solve
