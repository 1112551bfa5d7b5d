"""
`ForwardAdjoint`: the device pipeline behind `value_and_grad(loss)(params)` for a batch of B
designs on one structure -- what `jit(value_and_grad(loss_fn))` is to the reference
(optimization/optimizers/optimizer.py:365-370), with the quadratic benchmark loss of
SURVEY.md §8d:  L_b = 1/2 ||x_free_b - x_free0||^2.

All buffers are allocated once; one step is a fixed sequence of this library's kernels
(K1 assemble -> K2 solve -> positions -> K4 state -> loss/cotangent -> K2 adjoint solve ->
K3 gradients) launched through the C ABI on one stream, optionally replayed as a CUDA graph.
`step()` runs on device-resident inputs; `__call__` is the end-to-end call: host (pinned)
parameters in, host loss and gradients out.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib as L

F64 = torch.float64


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class ForwardAdjoint:
    def __init__(self, structure, batch=1, solver="auto", tol=1e-12, max_iters=0, with_state=True,
                 use_graph=True, device=None, fused_assembly=True):
        s = self.s = structure
        self.B = B = int(batch)
        self.dev = dev = torch.device("cuda", s.device if device is None else device)
        self.lib = L.lib()
        self.with_state = with_state
        E, N, Nf, Ns, nnz = s.num_edges, s.num_nodes, s.num_free, s.num_supports, s.nnz
        z = lambda *shape: torch.zeros(shape, dtype=F64, device=dev)
        # inputs
        self.q, self.xyz_fixed, self.loads, self.x0 = z(B, E), z(Ns, 3), z(N, 3), z(Nf, 3)
        # forward
        self.K, self.rhs, self.x, self.xyz = z(B, nnz), z(B, Nf, 3), z(B, Nf, 3), z(B, N, 3)
        if with_state:
            self.vectors, self.lengths, self.forces, self.residuals = z(B, E, 3), z(B, E), z(B, E), z(B, N, 3)
        # reverse
        self.g, self.lam, self.loss = z(B, Nf, 3), z(B, Nf, 3), z(B)
        self.q_bar, self.loads_bar, self.xs_bar = z(B, E), z(B, N, 3), z(B, Ns, 3)
        self.info_f, self.info_a = z(B, L.INFO_STRIDE), z(B, L.INFO_STRIDE)
        self.loss_sum = z(1)
        self.solver_id = L.SOLVERS[solver]
        need = int(self.lib.fdm_workspace_bytes(s.plan, B, self.solver_id))
        self.ws = torch.empty(max(need, 16), dtype=torch.uint8, device=dev)
        self.opts_f = L.SolveOpts(self.solver_id, int(max_iters), float(tol), 0, 0)
        # the adjoint system has the same K: reuse the Cholesky factor / coarse inverse left in ws
        self.opts_a = L.SolveOpts(self.solver_id, int(max_iters), float(tol), 0, 1)
        self.capturable = not (solver == "pcg" or (solver == "auto" and self._auto_is_multikernel()))
        # fused_assembly: fdm_solve_from_q (K.data never formed: 33 KB less memory per 30x30 design, one launch
        # less).  On by default for batches the tensor-core band kernels handle: 2 % faster per step than assemble
        # + solve once the rows' edge / support lists come as padded vectors (plan.band_row_*); bit-identical.
        is_chol = solver == "cholesky" or (solver == "auto" and self._auto_is_cholesky())
        self.is_cholesky = is_chol
        self.fused_assembly = bool(fused_assembly and is_chol and B > 1 and s.half_bandwidth <= 30)
        if self.fused_assembly:
            self.K = z(1)                                  # never formed; the adjoint solve only passes the pointer on
        self.graph = None
        self.use_graph = use_graph and self.capturable
        self.stream = torch.cuda.Stream(device=dev)
        # pinned host staging for the end-to-end call
        self.h_q = torch.zeros((B, E), dtype=F64).pin_memory()
        self.h_loss = torch.zeros((B,), dtype=F64).pin_memory()
        self.h_q_bar = torch.zeros((B, E), dtype=F64).pin_memory()
        self.launches_per_step = 0

    def _auto_is_cholesky(self):
        # mirrors pick_solver() in csrc/capi.cu: batches go to the band Cholesky when the plan supports it
        has_chol = int(self.lib.fdm_workspace_bytes(self.s.plan, self.B, L.SOLVER_CHOLESKY_BATCHED)) > 0
        has_pcg2 = int(self.lib.fdm_workspace_bytes(self.s.plan, self.B, L.SOLVER_PCG_PERSISTENT)) > 0
        return has_chol and (self.B > 1 or not has_pcg2 or self.s.half_bandwidth <= 31)

    def _auto_is_multikernel(self):
        return not self._auto_is_cholesky() and int(self.lib.fdm_workspace_bytes(self.s.plan, self.B, L.SOLVER_PCG_PERSISTENT)) == 0

    # ------------------------------------------------------------------------------
    def set_inputs(self, q, xyz_fixed, loads, x0):
        T = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
        self.q.copy_(T(q).reshape(self.B, -1))
        self.xyz_fixed.copy_(T(xyz_fixed))
        self.loads.copy_(T(loads))
        self.x0.copy_(T(x0))
        self.h_q.copy_(T(q).reshape(self.B, -1))
        torch.cuda.synchronize(self.dev)

    def _launch(self, st):
        """The kernel sequence of one forward+adjoint step on stream `st` (a raw cudaStream_t); returns the
        number of kernels launched (the forward band Cholesky solve is two: factor, substitution)."""
        s, lib, B = self.s, self.lib, self.B
        E, N, Nf, Ns = s.num_edges, s.num_nodes, s.num_free, s.num_supports
        ck = L.check
        n = 0
        if self.fused_assembly:
            # K1 + K2 in one call: the batched factorisation assembles its rows from q in place
            ck(lib.fdm_solve_from_q(s.plan, _p(self.q), _p(self.xyz_fixed), _p(self.loads), _p(self.x), _p(self.info_f),
                                    _p(self.ws), self.ws.numel(), B, E, 0, 0, C.byref(self.opts_f), st)); n += 2
        else:
            ck(lib.fdm_assemble(s.plan, _p(self.q), _p(self.xyz_fixed), _p(self.loads), _p(self.K),
                                _p(self.rhs), B, E, 0, 0, st)); n += 1
            ck(lib.fdm_solve(s.plan, _p(self.K), _p(self.rhs), _p(self.x), _p(self.info_f), _p(self.ws),
                             self.ws.numel(), B, C.byref(self.opts_f), st)); n += 2 if self.is_cholesky else 1
        ck(lib.fdm_positions(s.plan, _p(self.x), _p(self.xyz_fixed), _p(self.xyz), B, 0, 0, st)); n += 1
        if self.with_state:
            ck(lib.fdm_state(s.plan, _p(self.q), _p(self.xyz), _p(self.loads), _p(self.vectors),
                             _p(self.lengths), _p(self.forces), _p(self.residuals), B, E, 0, st)); n += 1
        ck(lib.fdm_loss_sqdist(_p(self.x), _p(self.x0), _p(self.g), _p(self.loss), Nf * 3, B, 0, st)); n += 1
        ck(lib.fdm_solve(s.plan, _p(self.K), _p(self.g), _p(self.lam), _p(self.info_a), _p(self.ws),
                         self.ws.numel(), B, C.byref(self.opts_a), st)); n += 1
        ck(lib.fdm_adjoint_grads(s.plan, _p(self.q), _p(self.xyz), _p(self.lam), _p(self.q_bar),
                                 _p(self.loads_bar), _p(self.xs_bar), B, E, 0, st)); n += 1
        self.launches_per_step = n
        return n

    def step(self):
        """One forward+adjoint on device-resident inputs, asynchronous on self.stream."""
        with torch.cuda.stream(self.stream):
            if self.use_graph:
                if self.graph is None:
                    self._launch(C.c_void_p(self.stream.cuda_stream))      # warm-up outside capture
                    self.stream.synchronize()
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph, stream=self.stream):
                        self._launch(C.c_void_p(torch.cuda.current_stream().cuda_stream))
                self.graph.replay()
            else:
                self._launch(C.c_void_p(self.stream.cuda_stream))

    def __call__(self, q_host=None):
        """End to end: pinned host q -> device, step, loss and q_bar back to pinned host."""
        self.enqueue_e2e(q_host)
        self.stream.synchronize()
        return self.h_loss, self.h_q_bar

    def enqueue_e2e(self, q_host=None):
        """Asynchronous end-to-end pass on self.stream: H2D q, step, D2H loss and q_bar (no sync)."""
        src = self.h_q if q_host is None else q_host
        with torch.cuda.stream(self.stream):
            self.q.copy_(src, non_blocking=True)
        self.step()
        with torch.cuda.stream(self.stream):
            self.h_loss.copy_(self.loss, non_blocking=True)
            self.h_q_bar.copy_(self.q_bar, non_blocking=True)

    @property
    def h2d_bytes(self):
        return self.h_q.numel() * 8

    @property
    def d2h_bytes(self):
        return (self.h_loss.numel() + self.h_q_bar.numel()) * 8


class ChunkedForwardAdjoint:
    """
    The end-to-end call for large batches: the batch is cut into `chunks` contiguous slices,
    each with its own ForwardAdjoint (own stream, buffers, CUDA graph), so the host->device copy
    of one slice, the kernels of another and the device->host copy of a third overlap on the
    copy engines and the SMs.  Results land in per-chunk pinned buffers; `__call__` returns
    them concatenated (views of one pinned allocation each).
    """

    def __init__(self, structure, batch, chunks=4, sizes=None, **kw):
        if sizes is None:
            chunks = max(1, min(int(chunks), int(batch)))
            sizes = [batch // chunks + (1 if i < batch % chunks else 0) for i in range(chunks)]
        sizes = [int(b) for b in sizes]
        assert sum(sizes) == int(batch) and min(sizes) > 0, "slice sizes must add up to the batch"
        chunks = len(sizes)
        self.offsets = [sum(sizes[:i]) for i in range(chunks)]
        self.parts = [ForwardAdjoint(structure, batch=b, **kw) for b in sizes]
        self.B = int(batch)
        # one loss vector for the whole batch; every slice writes its part (set before the graphs are captured)
        self.loss = torch.empty((self.B,), dtype=F64, device=self.parts[0].dev)
        for off, part in zip(self.offsets, self.parts):
            part.loss = self.loss[off:off + part.B]
        self.stream = torch.cuda.Stream(device=self.parts[0].dev)

    def step(self):
        """One forward+adjoint of the whole batch on device-resident inputs, ordered on `self.stream`: the slices
        run on their own streams between a fork and a join, so the bandwidth-bound substitution kernels of one
        slice overlap the factorisation of another.  Asynchronous; events recorded on self.stream bracket it."""
        for part in self.parts:
            part.stream.wait_stream(self.stream)
            part.step()
        for part in self.parts:
            self.stream.wait_stream(part.stream)

    def set_inputs(self, q, xyz_fixed, loads, x0):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(self.B, -1)
        for off, part in zip(self.offsets, self.parts):
            part.set_inputs(q[off:off + part.B], xyz_fixed, loads, x0)

    def enqueue(self, q_host=None):
        """Asynchronous end-to-end pass of every slice on its own stream (no synchronisation)."""
        for off, part in zip(self.offsets, self.parts):
            part.enqueue_e2e(None if q_host is None else q_host[off:off + part.B])

    def __call__(self, q_host=None):
        self.enqueue(q_host)
        self.synchronize()
        return [p.h_loss for p in self.parts], [p.h_q_bar for p in self.parts]

    def synchronize(self):
        for part in self.parts:
            part.stream.synchronize()

    @property
    def h2d_bytes(self):
        return sum(p.h2d_bytes for p in self.parts)

    @property
    def d2h_bytes(self):
        return sum(p.d2h_bytes for p in self.parts)

    @property
    def launches_per_step(self):
        return sum(p.launches_per_step for p in self.parts)


class _DevView:
    """A raw device pointer as a `__cuda_array_interface__` object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(d) for d in shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class NativeForwardAdjoint:
    """
    The same forward + adjoint pass as `ChunkedForwardAdjoint`, owned by the library (`fdm_fa_*`,
    include/fdm_b200.h, csrc/pipeline.cu): device buffers, chunk streams, events and CUDA graphs live in C++,
    so a step costs Python ONE ctypes call -- `step()` on device-resident q, `__call__` end to end (host q in,
    host loss and q_bar out; the pinned staging comes from fdm_host_alloc).  Loss: the quadratic benchmark loss
    by default, or any `jax_fdm_b200.losses.Loss` of goals (`set_loss`), compiled to a goal program.
    """
    ARRAYS = {"q": (L.FA_Q, "E"), "x": (L.FA_X_FREE, "Nf3"), "xyz": (L.FA_XYZ, "N3"), "loss": (L.FA_LOSS, ""),
              "q_bar": (L.FA_Q_BAR, "E"), "loads_bar": (L.FA_LOADS_BAR, "N3"), "xs_bar": (L.FA_XYZ_FIXED_BAR, "Ns3"),
              "info_f": (L.FA_INFO_FORWARD, "I"), "info_a": (L.FA_INFO_ADJOINT, "I"), "lengths": (L.FA_LENGTHS, "E"),
              "forces": (L.FA_FORCES, "E"), "residuals": (L.FA_RESIDUALS, "N3"), "vectors": (L.FA_VECTORS, "E3")}

    def __init__(self, structure, batch, chunks=1, solver="auto", with_state=True, use_graph=True, tol=1e-12,
                 max_iters=0, device=None):
        s = self.s = structure
        self.B = int(batch)
        self.lib = L.lib()
        self.dev = torch.device("cuda", s.device if device is None else device)
        if chunks == "auto":
            # device-resident step: slices of ~512 designs (measured on the 30x30 pillow: 4096 designs 8 slices
            # 1.46 ms against 1.52 in 2, 2048 in 4 slices 0.78 against 0.91 in one, 1024 in 2 slices 0.46 against
            # 0.50; below that one slice) -- a slice's bandwidth-bound kernels run beside the next slice's factorisation
            chunks = max(1, self.B // 512)
        opts = L.FaOpts(int(chunks), int(bool(with_state)), L.SOLVERS[solver], int(bool(use_graph)), int(max_iters), float(tol))
        self.handle = C.c_void_p()
        L.check(self.lib.fdm_fa_create(s.plan, self.B, C.byref(opts), C.byref(self.handle)))
        E = s.num_edges
        self._pinned = []
        self.h_q = self._host((self.B, E))
        self.h_loss = self._host((self.B,))
        self.h_q_bar = self._host((self.B, E))
        self.stream = torch.cuda.Stream(device=self.dev)
        self._loss_program = None

    def _host(self, shape):
        n = int(np.prod(shape))
        ptr = self.lib.fdm_host_alloc(n * 8)
        if not ptr:
            raise L.FdmError(-2, self.lib.fdm_last_error().decode())
        self._pinned.append(ptr)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,)).reshape(shape)

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.fdm_fa_destroy(self.handle)
            self.handle = C.c_void_p()
        for ptr in getattr(self, "_pinned", []):
            self.lib.fdm_host_free(ptr)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_inputs(self, q, xyz_fixed, loads, x0=None):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        q = f(q).reshape(self.B, -1)
        xs, p = f(xyz_fixed), f(loads)
        x0 = None if x0 is None else f(x0)
        L.check(self.lib.fdm_fa_set_constants(self.handle, xs.ctypes.data, p.ctypes.data,
                                              None if x0 is None else x0.ctypes.data))
        self.h_q[...] = q
        L.check(self.lib.fdm_fa_set_q(self.handle, self.h_q.ctypes.data))

    def set_loss(self, loss):
        """A `losses.Loss` (error terms only; regularisers act on the parameters above the path)."""
        prog = loss.program(self.s, None)
        self._loss_program = prog                           # keeps the host tables alive during the call
        L.check(self.lib.fdm_fa_set_goal_program(self.handle, C.byref(prog.host_struct)))

    def step(self, stream=None):
        st = self.stream if stream is None else stream
        L.check(self.lib.fdm_fa_step_device(self.handle, C.c_void_p(st.cuda_stream)))

    def enqueue(self, q_host=None, want_q_bar=True):
        src = self.h_q if q_host is None else q_host
        L.check(self.lib.fdm_fa_enqueue_host(self.handle, src.ctypes.data, self.h_loss.ctypes.data,
                                             self.h_q_bar.ctypes.data if want_q_bar else None))

    def synchronize(self):
        L.check(self.lib.fdm_fa_wait(self.handle))

    def enqueue_copies_only(self):
        """Measurement: the host <-> device copies of `enqueue` without the kernels."""
        L.check(self.lib.fdm_fa_enqueue_copies_only(self.handle, self.h_q.ctypes.data, self.h_loss.ctypes.data,
                                                    self.h_q_bar.ctypes.data))

    def __call__(self, q_host=None):
        src = self.h_q if q_host is None else q_host
        L.check(self.lib.fdm_fa_run_host(self.handle, src.ctypes.data, self.h_loss.ctypes.data, self.h_q_bar.ctypes.data))
        return self.h_loss, self.h_q_bar

    def array(self, name):
        """A device array of the object as a torch tensor (a view; valid while the object lives)."""
        which, kind = self.ARRAYS[name]
        ptr, per = C.c_void_p(), C.c_int64()
        L.check(self.lib.fdm_fa_device_array(self.handle, which, C.byref(ptr), C.byref(per)))
        s = self.s
        tail = {"E": (s.num_edges,), "Nf3": (s.num_free, 3), "N3": (s.num_nodes, 3), "Ns3": (s.num_supports, 3),
                "I": (L.INFO_STRIDE,), "E3": (s.num_edges, 3), "": ()}[kind]
        return torch.as_tensor(_DevView(ptr.value, (self.B,) + tail), device=self.dev)

    @property
    def launches_per_step(self):
        return int(self.lib.fdm_fa_info(self.handle, L.FA_INFO_LAUNCHES))

    @property
    def chunks(self):
        return int(self.lib.fdm_fa_info(self.handle, L.FA_INFO_CHUNKS))

    @property
    def h2d_bytes(self):
        return self.h_q.size * 8

    @property
    def d2h_bytes(self):
        return (self.h_loss.size + self.h_q_bar.size) * 8
