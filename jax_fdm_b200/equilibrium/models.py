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
from .solvers import (fixed_point_bwd_adjoint, is_solver_fixedpoint, is_solver_leastsquares, is_solver_root_finding,
                      solver_fixedpoint_implicit, solver_forward)
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
        return ops.stiffness_vjp(ctx.structure, gK.contiguous()), None


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


class _NodesResiduals(torch.autograd.Function):
    """loads - C^T (q * vectors) for GIVEN edge vectors (models.py:164-193) and its transposes.  C^T w is the
    xyz-cotangent of the state map for a cotangent w of `vectors` alone (vectors = C xyz), so the segmented sum over
    each node's incident edges runs in fdm_state_vjp; C g, needed by the transposes, is fdm_state on g."""

    @staticmethod
    def forward(ctx, q, loads, vectors, structure):
        ctx.structure = structure
        ctx.save_for_backward(q, vectors)
        qv = (q.unsqueeze(-1) * vectors).contiguous()
        zero_xyz = torch.zeros_like(loads)
        # xyz_bar of the state VJP with only a `vectors` cotangent is C^T vectors_cot
        _, ct, _ = ops.state_vjp(structure, q.contiguous(), zero_xyz, vectors_cot=qv)
        return loads - ct

    @staticmethod
    def backward(ctx, g):
        q, vectors = ctx.saved_tensors
        s = ctx.structure
        # C g per edge = edge vectors of the "geometry" g
        Cg = ops.state(s, torch.zeros_like(q), g.contiguous(), torch.zeros_like(g))[0]
        gq = -(Cg * vectors).sum(-1) if ctx.needs_input_grad[0] else None
        gv = -(q.unsqueeze(-1) * Cg) if ctx.needs_input_grad[2] else None
        return gq, g, gv, None


def _Positions_free(loads_nodes, structure):
    """loads[indices_free]: the free rows of a nodal array (the transpose half of nodes_positions)."""
    return _PositionsT.apply(loads_nodes, structure)


class _PositionsT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz_like, structure):
        ctx.structure = structure
        ctx.ns = structure.num_supports
        return ops.positions_vjp(structure, xyz_like.contiguous())[0]

    @staticmethod
    def backward(ctx, g):
        s = ctx.structure
        zeros = torch.zeros(g.shape[:-2] + (ctx.ns, 3), dtype=g.dtype, device=g.device)
        return ops.positions(s, g.contiguous(), zeros), None


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
        ws = ops.Workspace()
        x = K = None
        if opts.get("fused_assembly", True) and opts.get("solver", "auto") in ("auto", "cholesky"):
            # batched designs on a narrow band: K is never materialised (None when this structure / batch is not served)
            x = ops.solve_from_q(structure, q, xyz_fixed, loads, workspace=ws, nan_on_failure=True)
        if x is None:
            opts = {k: v for k, v in opts.items() if k != "fused_assembly"}
            K, rhs = ops.assemble(structure, q, xyz_fixed, loads)
            x = ops.solve(structure, K, rhs, workspace=ws, nan_on_failure=True, **opts)
        else:
            K = torch.empty(0, dtype=q.dtype, device=q.device)
            opts = {k: v for k, v in opts.items() if k != "fused_assembly"}
            opts["solver"] = "cholesky"
        ctx.structure, ctx.opts, ctx.ws = structure, opts, ws
        ctx.save_for_backward(q, xyz_fixed, x, K)
        return x

    @staticmethod
    def backward(ctx, g):
        q, xyz_fixed, x, K = ctx.saved_tensors
        s = ctx.structure
        # same K, same workspace: the Cholesky factor / coarse inverse of the forward solve is reused
        # after a fused forward the stored factor is all the adjoint solve needs (K was never formed)
        lam = ops.solve(s, K if K.numel() else None, g.contiguous(), workspace=ctx.ws, reuse_factor=True, **ctx.opts)
        xyz = ops.positions(s, x, xyz_fixed)
        qbar, pbar, xsbar = ops.adjoint_grads(s, q, xyz, lam)
        return qbar, xsbar, pbar, None, None


class _NodesLoad(torch.autograd.Function):
    """nodes_load with tributary edge loads, global or edge-local frame (models.py:290-344, loads.py:296-465)."""

    @staticmethod
    def forward(ctx, xyz, loads_nodes, edges_load, structure, is_local):
        ctx.structure, ctx.is_local = structure, is_local
        ctx.save_for_backward(xyz, edges_load)
        return ops.edge_loads(structure, xyz, loads_nodes, edges_load, is_local=is_local)

    @staticmethod
    def backward(ctx, g):
        xyz, edges_load = ctx.saved_tensors
        xb, eb = ops.edge_loads_vjp(ctx.structure, xyz, edges_load, g.contiguous(), want_xyz=ctx.needs_input_grad[0],
                                    want_edges=ctx.needs_input_grad[2], is_local=ctx.is_local)
        return xb, g, eb, None, None


class _FacesLoad(torch.autograd.Function):
    """nodes_load with tributary face loads, global or face-local frame (models.py:333-344, loads.py:35-288)."""

    @staticmethod
    def forward(ctx, xyz, loads_nodes, faces_load, structure, is_local):
        out, aux = ops.face_loads(structure, xyz, loads_nodes, faces_load, is_local=is_local)
        ctx.structure, ctx.is_local = structure, is_local
        ctx.save_for_backward(xyz, faces_load, *aux)
        return out

    @staticmethod
    def backward(ctx, g):
        xyz, faces_load, cen, gl = ctx.saved_tensors
        xb, fb = ops.face_loads_vjp(ctx.structure, xyz, faces_load, (cen, gl), g.contiguous(),
                                    want_xyz=ctx.needs_input_grad[0], want_faces=ctx.needs_input_grad[2],
                                    is_local=ctx.is_local)
        return xb, g, fb, None, None


class _LazyCount:
    """The iteration count of a device-resident loop: stays on the device until somebody looks (comparison, int(),
    printing), so running the model never synchronises for it."""

    def __init__(self, read):
        self._read, self._value = read, None

    def __int__(self):
        if self._value is None:
            self._value = int(self._read())
        return self._value

    __index__ = __int__

    def __repr__(self):
        return str(int(self))

    def __eq__(self, other):
        return int(self) == other

    def __lt__(self, other):
        return int(self) < other

    def __le__(self, other):
        return int(self) <= other

    def __gt__(self, other):
        return int(self) > other

    def __ge__(self, other):
        return int(self) >= other


class _NativeLoops:
    """One `fdm_fp` object (csrc/fixed_point.cu): both loops of the tmax > 1 equilibrium as CUDA graphs with a
    conditional WHILE node.  Pooled per structure and configuration; an object is busy from a forward call until its
    backward has run (it keeps the factor and the converged geometry in between)."""

    def __init__(self, structure, key, has_edges, has_faces, device):
        import ctypes as C
        from .. import _lib as L
        tmax, eta, local = key[:3]
        self.lib, self.L, self.C = L.lib(), L, C
        self.handle = C.c_void_p()
        self.busy = False
        opts = L.FpOpts(int(tmax), int(bool(local)), 0, float(eta), 0.0)
        mesh = [None, 0, 0, None, None, None, None]
        if has_faces:
            d = structure.device_arrays(device)
            self._mesh_keepalive = d
            P = lambda t: C.c_void_p(t.data_ptr())
            mesh = [P(d["faces"]), structure.num_faces, structure.faces_indexed.shape[1], P(d["edges_faces"]),
                    P(d["faces_edges"]), P(d["node_faces_ptr"]), P(d["node_faces"])]
        L.check(self.lib.fdm_fp_create(structure.plan, C.byref(opts), int(has_edges), *mesh, C.byref(self.handle)))

    def __del__(self):
        try:
            if self.handle.value:
                self.lib.fdm_fp_destroy(self.handle)
                self.handle = self.C.c_void_p()
        except Exception:
            pass

    @staticmethod
    def acquire(structure, key, has_edges, has_faces, device):
        pool = structure.__dict__.setdefault("_fp_pool", {}).setdefault((key, has_edges, has_faces, str(device)), [])
        for obj in pool:
            if not obj.busy:
                obj.busy = True
                return obj
        obj = _NativeLoops(structure, key, has_edges, has_faces, device)
        obj.busy = True
        pool.append(obj)
        return obj

    def info(self, stream):
        import numpy as np
        out = np.zeros(5)
        self.L.check(self.lib.fdm_fp_info(self.handle, out.ctypes.data, stream))
        return out


class _Lease:
    """Frees a pooled loop object when the autograd context that used it goes away."""

    def __init__(self, obj):
        self.obj = obj

    def release(self):
        if self.obj is not None:
            self.obj.busy = False
            self.obj = None

    __del__ = release


class _FixedPoint(torch.autograd.Function):
    """
    x* = K(q)^{-1} P(x*),  P(x) = nodes_load(xyz(x))[free] - R_fixed(q, x_s): the fixed-point
    equilibrium of models.py:512-618 as ONE node.  Forward = solver_forward with a single
    factorisation of K reused by every iteration; backward = the implicit adjoint of
    fixed_point.py:501-604 with the same factor:  w - J_P^T K^{-1} w = x_bar,  lam = K^{-1} w,
    then the cotangents of f(theta, x*) = K^{-1} P(theta, x*) at w.
    """

    @staticmethod
    def _native_ok(q, x_init, structure, opts, config):
        """The device-resident loops serve one design on the band Cholesky with the default fixed-point solver."""
        return (q.dim() == 1 and x_init.dim() == 2 and q.is_cuda and config.get("solver") in (None, solver_forward)
                and opts.get("solver", "auto") in ("auto", "cholesky") and structure.half_bandwidth <= 31
                and not config.get("verbose"))

    @staticmethod
    def forward(ctx, q, xyz_fixed, loads_nodes, edges_load, faces_load, x_init, structure, opts, config):
        s = structure
        local = bool(config.get("is_load_local", False))
        has_faces = faces_load.numel() > 0
        host = {}                                           # the host-driven loop's pieces, built on first use

        def solve_fn(b):
            if "K" not in host:
                host["K"], host["rhs0"] = ops.assemble(s, q, xyz_fixed, torch.zeros_like(loads_nodes))   # rhs0 = -R_fixed
                host["ws"], host["first"] = ops.Workspace(), True
            x = ops.solve(s, host["K"], b.contiguous(), workspace=host["ws"], reuse_factor=not host["first"], **opts)
            host["first"] = False
            return x

        def f(_, x):
            xyz = ops.positions(s, x, xyz_fixed)
            loads = ops.edge_loads(s, xyz, loads_nodes, edges_load, is_local=local)
            if has_faces:
                loads, _ = ops.face_loads(s, xyz, loads, faces_load, is_local=local)
            if "K" not in host:
                host["K"], host["rhs0"] = ops.assemble(s, q, xyz_fixed, torch.zeros_like(loads_nodes))   # rhs0 = -R_fixed
                host["ws"], host["first"] = ops.Workspace(), True
            return solve_fn(ops.lincomb(ops.positions_vjp(s, loads.contiguous())[0], host["rhs0"]))

        ctx.native = None
        if _FixedPoint._native_ok(q, x_init, s, opts, config):
            import ctypes as C
            key = (int(config["tmax"]), float(config["eta"]), local)
            loops = _NativeLoops.acquire(s, key, True, has_faces, q.device)
            ctx.native, ctx.lease = loops, _Lease(loops)

            def device_loop(x0, cfg):
                P = lambda t: C.c_void_p(t.data_ptr())
                st = C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)
                x = torch.empty_like(x0)
                args = [t.contiguous() for t in (q, xyz_fixed, loads_nodes, edges_load, faces_load, x0)]
                loops.L.check(loops.lib.fdm_fp_forward(loops.handle, P(args[0]), P(args[1]), P(args[2]), P(args[3]),
                                                       P(args[4]) if has_faces else None, P(args[5]), P(x), None, st))
                cfg["steps"] = _LazyCount(lambda: loops.info(st)[0])
                return x
            f.device_loop = device_loop

        x = solver_fixedpoint_implicit(config.get("solver"), config, f, None, x_init)
        ctx.structure, ctx.solve_fn, ctx.has_faces, ctx.local = s, solve_fn, has_faces, local
        ctx.save_for_backward(q, xyz_fixed, edges_load, faces_load, x)
        return x

    @staticmethod
    def backward(ctx, g):
        q, xyz_fixed, edges_load, faces_load, x = ctx.saved_tensors
        s, solve_fn, has_faces, local = ctx.structure, ctx.solve_fn, ctx.has_faces, ctx.local
        xyz = ops.positions(s, x, xyz_fixed)
        aux = None
        if has_faces:   # centroids / global face loads of the converged geometry, needed by the face-load transpose
            _, aux = ops.face_loads(s, xyz, torch.zeros_like(xyz), faces_load, is_local=local)

        def loads_vjp_x(lam):                       # J_P^T lam, free rows
            c = ops.positions(s, lam.contiguous(), torch.zeros_like(xyz_fixed))     # lam on the free rows, 0 on supports
            xb, _ = ops.edge_loads_vjp(s, xyz, edges_load, c, want_edges=False, is_local=local)
            if has_faces:
                xb = ops.lincomb(xb, ops.face_loads_vjp(s, xyz, faces_load, aux, c, want_faces=False, is_local=local)[0])
            return ops.positions_vjp(s, xb)[0]

        if ctx.native is not None:
            import ctypes as C
            loops = ctx.native

            def device_adjoint(vec):                # the adjoint system as one graph launch (csrc/fixed_point.cu)
                st = C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)
                vec = vec.contiguous()
                w, lam = torch.empty_like(x), torch.empty_like(x)
                loops.L.check(loops.lib.fdm_fp_adjoint(loops.handle, C.c_void_p(vec.data_ptr()), C.c_void_p(w.data_ptr()),
                                                       C.c_void_p(lam.data_ptr()), st))
                return w, lam
            solve_fn.device_adjoint = device_adjoint

        w, lam = fixed_point_bwd_adjoint(solve_fn, loads_vjp_x, g.contiguous())
        if ctx.native is not None:
            ctx.lease.release()
        # cotangents of f(theta, x*) at w: K and R_fixed terms (K3), then the load terms
        qbar, pbar, xsbar = ops.adjoint_grads(s, q, xyz, lam)
        xb, ebar = ops.edge_loads_vjp(s, xyz, edges_load, pbar, is_local=local)
        fbar = None
        if has_faces:
            xb2, fbar = ops.face_loads_vjp(s, xyz, faces_load, aux, pbar, is_local=local)
            xb = ops.lincomb(xb, xb2)
        xsbar = ops.lincomb(xsbar, ops.positions_vjp(s, xb)[1])   # edge lengths / face areas at the supports depend on x_s
        return qbar, xsbar, pbar, ebar, fbar, None, None, None, None


class EquilibriumModel:
    """
    FDM model (models.py:47-91).  `tmax=1`: one linear FDM step with nodal loads.  `tmax>1`: the
    fixed-point iteration for shape-dependent loads (edge line loads and, on mesh structures, face area
    loads, in the global frame or -- is_load_local -- in the edge / face frames that follow the deformed
    geometry, geometry.py:581-657) with the implicit adjoint.
    The dense variant assembles the same K and solves with the same kernels: the "dense" vs
    "sparse" split of the reference is a storage choice, not different arithmetic.
    """

    def __init__(self, tmax=100, eta=1e-6, is_load_local=False, itersolve_fn=None, iterload_fn=None,
                 implicit_diff=True, verbose=False, solver="auto", tol=0.0, max_iters=0):
        self.tmax = tmax                                  # default 100, as models.py:77
        self.eta = eta
        self.is_load_local = is_load_local
        self.itersolve_fn = itersolve_fn or solver_forward
        self.iterload_fn = iterload_fn
        self.eq_iterative_fn = self.select_equilibrium_iterative_fn(self.itersolve_fn)
        self.implicit_diff = implicit_diff
        self.verbose = verbose
        self.solve_opts = {"solver": solver, "tol": tol, "max_iters": max_iters}
        self.linearsolve_fn = lambda K, P: sparse_solve(K, P, **self.solve_opts)

    # ---- edges / nodes (models.py:99-193) -------------------------------------------
    # The method-by-method route: the reference's one-liners with the same names and argument meaning.  `connectivity`
    # is `structure.connectivity` (scipy CSC / dense array carrying a reference to its structure) or the structure
    # itself; the products run in this library's kernels (fdm_state), not as matrix products.
    @staticmethod
    def _structure_of(connectivity):
        s = getattr(connectivity, "_fdm_structure", None)
        s = s() if callable(s) else s
        if s is None and hasattr(connectivity, "plan"):
            s = connectivity
        if s is None:
            raise TypeError("pass `structure.connectivity` (or the structure): the edge lists live in its plan")
        return s

    @classmethod
    def edges_vectors(cls, xyz, connectivity):
        """models.py:99-119: C @ xyz."""
        s = cls._structure_of(connectivity)
        q0 = torch.zeros(xyz.shape[:-2] + (s.num_edges,), dtype=xyz.dtype, device=xyz.device)
        return _State.apply(q0, xyz, torch.zeros_like(xyz), s)[0]

    @staticmethod
    def edges_lengths(vectors):
        """models.py:121-136: row norms, shape (E, 1)."""
        return torch.linalg.norm(vectors, dim=-1, keepdim=True)

    @staticmethod
    def edges_forces(q, lengths):
        """models.py:138-162: q * lengths, shape (E, 1)."""
        return q.unsqueeze(-1) * lengths

    @classmethod
    def nodes_residuals(cls, q, loads, vectors, connectivity):
        """models.py:164-193: loads - C^T (q * vectors)."""
        s = cls._structure_of(connectivity)
        return _NodesResiduals.apply(q, loads, vectors, s)

    @staticmethod
    def nodes_positions(xyz_free, xyz_fixed, structure):
        return _Positions.apply(xyz_free, xyz_fixed, structure)

    @staticmethod
    def faces_load(xyz, faces_load, structure, is_local=False):
        """models.py:348-375: tributary face loads on the nodes (nodes_load_from_faces)."""
        return _FacesLoad.apply(xyz, torch.zeros_like(xyz), faces_load, structure, is_local)

    @staticmethod
    def edges_load(xyz, edges_load, structure, is_local=False):
        """models.py:377-404: tributary edge loads on the nodes (nodes_load_from_edges)."""
        return _NodesLoad.apply(xyz, torch.zeros_like(xyz), edges_load, structure, is_local)

    # ---- matrices -------------------------------------------------------------------------
    @staticmethod
    def stiffness_matrix(q, structure):
        return CSC(_Stiffness.apply(q, structure), structure)

    @staticmethod
    def stiffness_matrix_dense(q, structure):
        """The (Nf, Nf) array the reference's dense `EquilibriumModel.stiffness_matrix` returns (models.py:800-822):
        the same `K.data` (fdm_assemble) scattered onto its pattern; differentiable with respect to q."""
        return CSC(_Stiffness.apply(q, structure), structure).todense()

    def load_matrix(self, q, xyz_fixed, load_nodes, structure):
        return _LoadMatrix.apply(q, xyz_fixed, load_nodes, structure)

    @staticmethod
    def residual_fixed_matrix(q, xyz_fixed, structure):
        zero = torch.zeros((structure.num_nodes, 3), dtype=q.dtype, device=q.device)
        return -_LoadMatrix.apply(q, xyz_fixed, zero, structure)

    def load_xyz_matrix(self, params, xyz_free, structure):
        """models.py:858-896: the load matrix with the shape-dependent loads evaluated at xyz_free."""
        q, xyz_fixed, load_state = params
        xyz = self.nodes_positions(xyz_free, xyz_fixed, structure)
        loads_nodes = self.nodes_load(xyz, load_state, structure)
        return self.load_matrix(q, xyz_fixed, loads_nodes, structure)

    def load_xyz_matrix_from_r_fixed(self, params, xyz_free, structure):
        """models.py:898-947: the same with the support term R_fixed precomputed: params = (K, R_fixed, xyz_fixed,
        load_state)."""
        _, R_fixed, xyz_fixed, load_state = params
        xyz = self.nodes_positions(xyz_free, xyz_fixed, structure)
        loads_nodes = self.nodes_load(xyz, load_state, structure)
        return _Positions_free(loads_nodes, structure) - R_fixed

    def residual_free_matrix(self, params, xyz_free, structure):
        """models.py:978-1014: K x_free - P(x_free)."""
        q = params[0]
        K = self.stiffness_matrix(q, structure)
        P = self.load_xyz_matrix(params, xyz_free, structure)
        return K @ xyz_free - P

    def select_equilibrium_iterative_fn(self, solver):
        """models.py:721-751.  Fixed-point solvers drive equilibrium_iterative_xyz; the least-squares and
        root-finding families (jaxopt / optimistix wrappers, solvers/least_squares.py, root_finding.py) drive the
        residual form, which is outside this build."""
        if is_solver_fixedpoint(solver):
            return self.equilibrium_iterative_xyz
        if is_solver_leastsquares(solver) or is_solver_root_finding(solver):
            return self.equilibrium_iterative_residual
        raise ValueError(f"Solver {solver} is not supported!")

    def equilibrium_iterative_residual(self, q, xyz_fixed, load_state, structure, xyz_free_init=None, tmax=100,
                                       eta=1e-6, solver=None, implicit_diff=True, verbose=False):
        """models.py:620-719: root of K x - P(x) by a least-squares / Newton solver.  Those solvers are wrappers
        of jaxopt / optimistix, which this build does not restate; the fixed-point form
        (`equilibrium_iterative_xyz`, the reference's default) reaches the same equilibrium."""
        raise NotImplementedError("the residual form needs a least-squares or root-finding solver (jaxopt / "
                                  "optimistix wrappers): outside this build; use a fixed-point solver "
                                  "(itersolve_fn=solver_forward, the default)")

    # ---- equilibrium ------------------------------------------------------------------------
    def nodes_free_positions(self, q, xyz_fixed, loads, structure, fused=True):
        if fused:
            return _FusedFreePositions.apply(q, xyz_fixed, loads, structure, self.solve_opts)
        K = self.stiffness_matrix(q, structure)
        P = self.load_matrix(q, xyz_fixed, loads, structure)
        return self.linearsolve_fn(K, P)

    def nodes_free_positions_jvp(self, q, xyz_fixed, loads, structure, q_dot=None, xyz_fixed_dot=None, loads_dot=None):
        """
        Forward-mode rule of the linear FDM step: returns (x_free, x_free_dot) with
            K x_dot = b_dot - K_dot x,   K_dot = K(q_dot)  (K is linear in q),
            b_dot = loads_dot_f - C_f^T (q_dot * C_s x_s) - C_f^T (q * C_s x_s_dot).
        The tangents may carry a leading axis T (T tangents at ONE primal, what `jacfwd` vmaps): the T tangent
        solves are more right-hand sides of the SAME factor -- one batched substitution launch on the band Cholesky
        (opts.reuse_factor = 2), a loop that reuses the coarse inverse on the iterative solvers.  The reference has
        no such rule: its sparse solve is reverse-only (sparse.py:1), which is why constraints (constrained.py:88,
        jacfwd) and Hessians force the dense model (fdm.py:266-271).
        """
        s = structure
        T = 0                                                              # 0: a single, unbatched tangent
        for d, base_rank in ((q_dot, 1), (xyz_fixed_dot, 2), (loads_dot, 2)):
            if d is not None and d.dim() > base_rank:
                T = max(T, int(d.shape[0]))
        lead = (T,) if T else ()
        zeros = lambda *shape: torch.zeros(lead + shape, dtype=q.dtype, device=q.device)
        q_dot = zeros(s.num_edges) if q_dot is None else q_dot
        loads_dot = zeros(s.num_nodes, 3) if loads_dot is None else loads_dot
        opts = {k: v for k, v in self.solve_opts.items() if k != "fused_assembly"}
        K, rhs = ops.assemble(s, q, xyz_fixed, loads)
        ws = ops.Workspace()
        ws.reserve(s, max(T, 1), opts.get("solver", "auto"), q.device)
        x = ops.solve(s, K, rhs, workspace=ws, **opts)
        K_dot, b_dot = ops.assemble(s, q_dot, xyz_fixed, loads_dot)        # loads_dot_f - C_f^T (q_dot * C_s x_s)
        b2 = None
        if xyz_fixed_dot is not None:                                      # - C_f^T (q * C_s x_s_dot)
            _, b2 = ops.assemble(s, q, xyz_fixed_dot, torch.zeros_like(loads), want_K=False)
        rhs_dot = ops.lincomb(b_dot, ops.spmm(s, K_dot, x), b2, 1.0, -1.0, 1.0)
        if T and s.half_bandwidth <= 31:
            x_dot = ops.solve(s, K, rhs_dot, workspace=ws, reuse_factor=2, **{**opts, "solver": "cholesky"})
        elif T:
            x_dot = torch.stack([ops.solve(s, K, rhs_dot[t], workspace=ws, reuse_factor=True, **opts) for t in range(T)])
        else:
            x_dot = ops.solve(s, K, rhs_dot, workspace=ws, reuse_factor=True, **opts)
        return x, x_dot

    def jvp(self, params, structure, q_dot=None, xyz_fixed_dot=None, loads_dot=None):
        """
        Forward mode through `__call__` for the linear model (tmax = 1, nodal loads): (state, state_dot), the
        tangents of every field of the equilibrium state for tangents of q, xyz_fixed and the nodal loads -- T of
        them at once with a leading axis on the tangents (`jax.jvp` / `jax.jacfwd` of `model(params, structure)`).
        """
        if self.tmax > 1:
            raise NotImplementedError("forward mode is built for the linear model (tmax = 1)")
        q, xyz_fixed, load_state = params
        loads = load_state.nodes if isinstance(load_state, LoadState) else load_state
        s = structure
        x, x_dot = self.nodes_free_positions_jvp(q, xyz_fixed, loads, s, q_dot, xyz_fixed_dot, loads_dot)
        xyz = ops.positions(s, x, xyz_fixed)
        xs_dot = xyz_fixed_dot if xyz_fixed_dot is not None else \
            torch.zeros(x_dot.shape[:-2] + (s.num_supports, 3), dtype=q.dtype, device=q.device)
        xyz_dot = ops.positions(s, x_dot, xs_dot)
        v, l, f, r = ops.state(s, q, xyz, loads)
        vd, ld, fd, rd = ops.state_jvp(s, q, xyz, xyz_dot, q_dot=q_dot, loads_dot=loads_dot)
        zero_loads = torch.zeros_like(xyz_dot) if loads_dot is None else loads_dot
        return (EquilibriumState(xyz=xyz, residuals=r, lengths=l, forces=f, loads=loads, vectors=v),
                EquilibriumState(xyz=xyz_dot, residuals=rd, lengths=ld, forces=fd, loads=zero_loads, vectors=vd))

    def nodes_equilibrium(self, q, xyz_fixed, loads_nodes, structure):
        return self.nodes_free_positions(q, xyz_fixed, loads_nodes, structure)

    def equilibrium(self, q, xyz_fixed, loads_nodes, structure):
        return self.nodes_equilibrium(q, xyz_fixed, loads_nodes, structure)

    def nodes_load(self, xyz, load_state, structure):
        """models.py:290-344: nodal loads plus tributary edge loads when an (E,3) edge load is present."""
        nodes_load, edges_load, faces_load = load_state
        has_edges = isinstance(edges_load, torch.Tensor) and edges_load.numel() > 1
        has_faces = isinstance(faces_load, torch.Tensor) and faces_load.numel() > 1 and hasattr(structure, "faces_indexed")
        if has_edges:
            nodes_load = _NodesLoad.apply(xyz, nodes_load, edges_load, structure, self.is_load_local)
        if has_faces:   # a non-scalar face load only takes effect on a mesh structure (models.py:333-344)
            nodes_load = _FacesLoad.apply(xyz, nodes_load, faces_load, structure, self.is_load_local)
        return nodes_load

    def equilibrium_iterative_xyz(self, q, xyz_fixed, load_state, structure, xyz_free_init=None, tmax=100,
                                  eta=1e-6, solver=None, implicit_diff=True, verbose=False):
        """models.py:512-618."""
        if xyz_free_init is None:
            xyz_free_init = self.equilibrium(q, xyz_fixed, load_state.nodes, structure)
        if not implicit_diff:
            # Unrolled differentiation (models.py:615, fixed_point.py:139-198 differentiated through the loop):
            # every iteration x <- K(q)^{-1} P(theta, x) is an autograd node of the same kernels, so the reverse
            # pass walks the iterations backwards.  Like the reference's loop it re-factorises K per iteration.
            def f(_, x):
                xyz = self.nodes_positions(x, xyz_fixed, structure)
                return _FusedFreePositions.apply(q, xyz_fixed, self.nodes_load(xyz, load_state, structure), structure,
                                                 self.solve_opts)
            config = {"tmax": tmax, "eta": eta, "verbose": verbose}
            self.last_solver_config = config
            return (solver or self.itersolve_fn)(f, None, xyz_free_init, config)
        edges_load = load_state.edges
        if not (isinstance(edges_load, torch.Tensor) and edges_load.numel() > 1):
            edges_load = torch.zeros((structure.num_edges, 3), dtype=q.dtype, device=q.device)
        faces_load = load_state.faces
        if not (isinstance(faces_load, torch.Tensor) and faces_load.numel() > 1 and hasattr(structure, "faces_indexed")):
            faces_load = torch.zeros((0, 3), dtype=q.dtype, device=q.device)
        config = {"tmax": tmax, "eta": eta, "verbose": verbose, "solver": solver or self.itersolve_fn,
                  "is_load_local": self.is_load_local}
        self.last_solver_config = config
        opts = {k: v for k, v in self.solve_opts.items() if k != "fused_assembly"}
        return _FixedPoint.apply(q, xyz_fixed, load_state.nodes, edges_load, faces_load, xyz_free_init.detach(),
                                 structure, opts, config)

    def equilibrium_state(self, q, xyz, loads_nodes, structure):
        vectors, lengths, forces, residuals = _State.apply(q, xyz, loads_nodes, structure)
        return EquilibriumState(xyz=xyz, residuals=residuals, lengths=lengths, forces=forces,
                                loads=loads_nodes, vectors=vectors)

    def __call__(self, params: EquilibriumParametersState, structure) -> EquilibriumState:
        q, xyz_fixed, load_state = params
        if not isinstance(load_state, LoadState):
            load_state = LoadState(load_state)
        xyz_free = self.equilibrium(q, xyz_fixed, load_state.nodes, structure)
        if self.tmax > 1:                                  # models.py:449-465
            if self.iterload_fn is not None:
                load_state = self.iterload_fn(load_state)
            xyz_free = self.equilibrium_iterative_xyz(q, xyz_fixed, load_state, structure, xyz_free_init=xyz_free,
                                                      tmax=self.tmax, eta=self.eta, solver=self.itersolve_fn,
                                                      implicit_diff=self.implicit_diff, verbose=self.verbose)
        xyz = self.nodes_positions(xyz_free, xyz_fixed, structure)
        load_nodes = self.nodes_load(xyz, load_state, structure)
        return self.equilibrium_state(q, xyz, load_nodes, structure)


class EquilibriumModelSparse(EquilibriumModel):
    """models.py:1022-1079: the same model with `linearsolve_fn = sparse_solve`."""
