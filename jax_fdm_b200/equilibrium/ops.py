"""
Thin wrappers that launch the C-ABI entry points on torch CUDA tensors (device memory and
streams only -- no torch math on this path).  Every function accepts an optional leading
batch axis (B designs sharing the structure) and runs on torch's current stream.
"""

import ctypes as C
import threading

import torch

from .. import _lib as L

F64 = torch.float64


_tls = threading.local()


def _stream():
    """torch's current stream ON THE DEVICE OF THE TENSORS of this call (the last one `_chk` saw), not on the
    process's current device."""
    dev = getattr(_tls, "device", None)
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _chk(t, name, shape_tail):
    if t.dtype != F64 or not t.is_cuda:
        raise TypeError(f"{name}: expected a float64 CUDA tensor, got {t.dtype} on {t.device}")
    if tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise ValueError(f"{name}: trailing shape {tuple(t.shape)} != {tuple(shape_tail)}")
    _tls.device = t.device
    return t.contiguous()


def _batch_of(t, base_ndim):
    """(batch, stride_in_elements, is_batched) for a tensor whose unbatched rank is base_ndim."""
    if t.dim() == base_ndim:
        return 1, 0, False
    if t.dim() == base_ndim + 1:
        per = 1
        for s in t.shape[1:]:
            per *= s
        return int(t.shape[0]), per, True
    raise ValueError(f"unexpected rank {t.dim()}")


def _resolve_batch(*pairs):
    """pairs of (tensor, base_ndim) -> common batch size B (1 if none batched) and batched flag."""
    B, batched = 1, False
    for t, nd in pairs:
        if t is None:
            continue
        b, _, isb = _batch_of(t, nd)
        if isb:
            if batched and b != B:
                raise ValueError("inconsistent batch sizes")
            B, batched = b, True
    return B, batched


def _stride(t, base_ndim):
    b, per, isb = _batch_of(t, base_ndim)
    return per if isb else 0


def assemble(structure, q, xyz_fixed=None, loads=None, want_K=True, want_rhs=True):
    """K1 -> (Kdata (B?,nnz), rhs (B?,Nf,3)); models.py:1064-1079, 828-856, 949-976."""
    s = structure
    E, Ns, N, Nf, nnz = s.num_edges, s.num_supports, s.num_nodes, s.num_free, s.nnz
    q = _chk(q, "q", (E,))
    if want_rhs:
        xyz_fixed = _chk(xyz_fixed, "xyz_fixed", (Ns, 3))
        loads = _chk(loads, "loads", (N, 3))
    B, batched = _resolve_batch((q, 1), (xyz_fixed if want_rhs else None, 2), (loads if want_rhs else None, 2))
    lead = (B,) if batched else ()
    K = torch.empty(lead + (nnz,), dtype=F64, device=q.device) if want_K else None
    rhs = torch.empty(lead + (Nf, 3), dtype=F64, device=q.device) if want_rhs else None
    if batched and q.dim() == 1:
        q = q.expand(B, E).contiguous()
    L.check(L.lib().fdm_assemble(s.plan, _ptr(q), _ptr(xyz_fixed if want_rhs else None),
                                 _ptr(loads if want_rhs else None), _ptr(K), _ptr(rhs), B, E,
                                 _stride(xyz_fixed, 2) if want_rhs else 0,
                                 _stride(loads, 2) if want_rhs else 0, _stream()))
    return K, rhs


class Workspace:
    """Caller-owned scratch for fdm_solve, cached per (structure, batch, solver)."""

    def __init__(self):
        self.buf = None

    def reserve(self, structure, batch, solver, device):
        """Size the buffer for `batch` designs up front: a later call with more right-hand-side sets than the call
        that factored (reuse_factor=2: one factor, many right-hand sides) must not reallocate the factor away."""
        return self.get(structure, batch, L.SOLVERS[solver] if isinstance(solver, str) else int(solver), device)

    def get(self, structure, batch, solver, device):
        need = int(L.lib().fdm_workspace_bytes(structure.plan, batch, solver))
        if self.buf is None or self.buf.numel() < need or self.buf.device != device:
            self.buf = torch.empty(max(need, 16), dtype=torch.uint8, device=device)
        return self.buf, need


def solve(structure, Kdata, rhs, solver="auto", tol=0.0, max_iters=0, reuse_factor=False,
          workspace=None, return_info=False, check_every=0, stages=0, nan_on_failure=False):
    """K2: x = K^{-1} rhs (sparse.py:207-230 / 296).  Kdata (B?,nnz), rhs (B?,Nf,3).
    `stages` (profiling): bit mask of the band Cholesky launches to run, 0 = all (fdm_b200.h).
    `nan_on_failure`: designs whose solve status is not OK come back as NaN (what the reference's
    solvers hand to the optimiser's NaN check, optimizer.py:372-380) -- a device-side select, no sync."""
    s = structure
    Nf, nnz = s.num_free, s.nnz
    rhs = _chk(rhs, "rhs", (Nf, 3))
    if Kdata is None:
        # only the stored factor is read (band Cholesky with reuse_factor, after solve_from_q)
        if not reuse_factor:
            raise ValueError("Kdata may be omitted only with reuse_factor=True")
        B, batched = _resolve_batch((rhs, 2))
        Kdata = rhs                                   # placeholder pointer, never dereferenced
    else:
        Kdata = _chk(Kdata, "Kdata", (nnz,))
        B, batched = _resolve_batch((Kdata, 1), (rhs, 2))
        if batched:
            if Kdata.dim() == 1:
                Kdata = Kdata.expand(B, nnz).contiguous()
            if rhs.dim() == 2:
                rhs = rhs.expand(B, Nf, 3).contiguous()
    x = torch.empty_like(rhs)
    info = torch.zeros((B, L.INFO_STRIDE), dtype=F64, device=rhs.device)
    sid = L.SOLVERS[solver] if isinstance(solver, str) else int(solver)
    ws = workspace if workspace is not None else Workspace()
    buf, need = ws.get(s, B, sid, rhs.device)
    opts = L.SolveOpts(sid, int(max_iters), float(tol), int(check_every), int(reuse_factor), int(stages),
                       int(bool(nan_on_failure)))                 # the NaN select happens in the solve's own kernels
    L.check(L.lib().fdm_solve(s.plan, _ptr(Kdata), _ptr(rhs), _ptr(x), _ptr(info), _ptr(buf),
                              buf.numel(), B, C.byref(opts), _stream()))
    return (x, info) if return_info else x


def solve_from_q(structure, q, xyz_fixed, loads, workspace=None, return_info=False, nan_on_failure=False):
    """
    K1 + K2 fused: x_free = K(q)^{-1} (loads_f - C_f^T Q C_s xyz_fixed) without materialising K.data
    (models.py:222-256).  Returns None when this structure / batch is not served by the in-place
    assembling Cholesky (the caller then runs `assemble` + `solve`, another CUDA path).
    """
    s = structure
    E, Ns, N, Nf = s.num_edges, s.num_supports, s.num_nodes, s.num_free
    q = _chk(q, "q", (E,))
    xyz_fixed = _chk(xyz_fixed, "xyz_fixed", (Ns, 3))
    loads = _chk(loads, "loads", (N, 3))
    B, batched = _resolve_batch((q, 1), (xyz_fixed, 2), (loads, 2))
    if batched and q.dim() == 1:
        q = q.expand(B, E).contiguous()
    sid = L.SOLVER_CHOLESKY_BATCHED
    if B < 2 or int(L.lib().fdm_workspace_bytes(s.plan, B, sid)) == 0 or s.half_bandwidth > 30:
        return None
    ws = workspace if workspace is not None else Workspace()
    buf, _ = ws.get(s, B, sid, q.device)
    x = torch.empty(((B,) if batched else ()) + (Nf, 3), dtype=F64, device=q.device)
    info = torch.zeros((B, L.INFO_STRIDE), dtype=F64, device=q.device)
    opts = L.SolveOpts(sid, 0, 0.0, 0, 0, 0, int(bool(nan_on_failure)))
    rc = L.lib().fdm_solve_from_q(s.plan, _ptr(q), _ptr(xyz_fixed), _ptr(loads), _ptr(x), _ptr(info), _ptr(buf),
                                  buf.numel(), B, E, _stride(xyz_fixed, 2), _stride(loads, 2), C.byref(opts), _stream())
    if rc == -5:          # FDM_ERR_UNSUPPORTED
        return None
    L.check(rc)
    return (x, info) if return_info else x


def spmm(structure, Kdata, X):
    """Y = K X (models.py:1012, sparse.py:301)."""
    s = structure
    Kdata = _chk(Kdata, "Kdata", (s.nnz,))
    X = _chk(X, "X", (s.num_free, 3))
    B, batched = _resolve_batch((Kdata, 1), (X, 2))
    if batched:
        if Kdata.dim() == 1:
            Kdata = Kdata.expand(B, s.nnz).contiguous()
        if X.dim() == 2:
            X = X.expand(B, s.num_free, 3).contiguous()
    Y = torch.empty_like(X)
    L.check(L.lib().fdm_spmm(s.plan, _ptr(Kdata), _ptr(X), _ptr(Y), B, _stream()))
    return Y


def kbar_pattern(structure, a, b, sign=1.0):
    """Kbar.data[k] = sign * sum_xyz a[row_k] * b[col_k] on the structure's pattern (fdm_kbar_pattern): the
    cotangent of K of sparse_solve_bwd (sparse.py:299-306) and the VJP of `K @ x` w.r.t. K.data."""
    s = structure
    a = _chk(a, "a", (s.num_free, 3))
    b = _chk(b, "b", (s.num_free, 3))
    B, batched = _resolve_batch((a, 2), (b, 2))
    if batched:
        a = a if a.dim() == 3 else a.expand(B, -1, -1).contiguous()
        b = b if b.dim() == 3 else b.expand(B, -1, -1).contiguous()
    out = torch.empty(((B,) if batched else ()) + (s.nnz,), dtype=F64, device=a.device)
    L.check(L.lib().fdm_kbar_pattern(s.plan, _ptr(a), _ptr(b), _ptr(out), float(sign), B, _stream()))
    return out


def stiffness_vjp(structure, Kbar):
    """q_bar from Kbar.data: the transpose of stiffness_matrix (models.py:1069-1077; fdm_stiffness_vjp)."""
    s = structure
    Kbar = _chk(Kbar, "Kbar", (s.nnz,))
    B, batched = _resolve_batch((Kbar, 1))
    out = torch.empty(((B,) if batched else ()) + (s.num_edges,), dtype=F64, device=Kbar.device)
    L.check(L.lib().fdm_stiffness_vjp(s.plan, _ptr(Kbar), _ptr(out), B, _stream()))
    return out


def positions(structure, x_free, xyz_fixed):
    """xyz = concat(x_free, xyz_fixed)[indices_freefixed] (models.py:195-220)."""
    s = structure
    x_free = _chk(x_free, "x_free", (s.num_free, 3))
    xyz_fixed = _chk(xyz_fixed, "xyz_fixed", (s.num_supports, 3))
    B, batched = _resolve_batch((x_free, 2), (xyz_fixed, 2))
    if batched and x_free.dim() == 2:
        x_free = x_free.expand(B, s.num_free, 3).contiguous()
    lead = (B,) if batched else ()
    xyz = torch.empty(lead + (s.num_nodes, 3), dtype=F64, device=x_free.device)
    L.check(L.lib().fdm_positions(s.plan, _ptr(x_free), _ptr(xyz_fixed), _ptr(xyz), B,
                                  _stride(xyz_fixed, 2), 0, _stream()))
    return xyz


def positions_vjp(structure, xyz_bar):
    """x_free_bar = xyz_bar[indices_free], xyz_fixed_bar = xyz_bar[indices_fixed]."""
    s = structure
    xyz_bar = _chk(xyz_bar, "xyz_bar", (s.num_nodes, 3))
    B, batched = _resolve_batch((xyz_bar, 2))
    lead = (B,) if batched else ()
    xf = torch.empty(lead + (s.num_free, 3), dtype=F64, device=xyz_bar.device)
    xs = torch.empty(lead + (s.num_supports, 3), dtype=F64, device=xyz_bar.device)
    L.check(L.lib().fdm_positions(s.plan, _ptr(xf), _ptr(xs), _ptr(xyz_bar), B,
                                  s.num_supports * 3, 1, _stream()))
    return xf, xs


def state(structure, q, xyz, loads):
    """K4 -> vectors (E,3), lengths (E,1), forces (E,1), residuals (N,3) (models.py:753-794)."""
    s = structure
    E, N = s.num_edges, s.num_nodes
    q = _chk(q, "q", (E,))
    xyz = _chk(xyz, "xyz", (N, 3))
    loads = _chk(loads, "loads", (N, 3))
    B, batched = _resolve_batch((q, 1), (xyz, 2), (loads, 2))
    if batched and xyz.dim() == 2:
        xyz = xyz.expand(B, N, 3).contiguous()
    if batched and q.dim() == 1:
        q = q.expand(B, E).contiguous()
    lead = (B,) if batched else ()
    dev = q.device
    vectors = torch.empty(lead + (E, 3), dtype=F64, device=dev)
    lengths = torch.empty(lead + (E, 1), dtype=F64, device=dev)
    forces = torch.empty(lead + (E, 1), dtype=F64, device=dev)
    residuals = torch.empty(lead + (N, 3), dtype=F64, device=dev)
    L.check(L.lib().fdm_state(s.plan, _ptr(q), _ptr(xyz), _ptr(loads), _ptr(vectors), _ptr(lengths),
                              _ptr(forces), _ptr(residuals), B, E, _stride(loads, 2), _stream()))
    return vectors, lengths, forces, residuals


def state_vjp(structure, q, xyz, xyz_cot=None, residuals_cot=None, lengths_cot=None,
              forces_cot=None, loads_cot=None, vectors_cot=None):
    """VJP of `state` -> (q_bar, xyz_bar, loads_bar)."""
    s = structure
    E, N = s.num_edges, s.num_nodes
    q = _chk(q, "q", (E,))
    xyz = _chk(xyz, "xyz", (N, 3))
    cots = [xyz_cot, residuals_cot, lengths_cot, forces_cot, loads_cot, vectors_cot]
    tails = [(N, 3), (N, 3), (E, 1), (E, 1), (N, 3), (E, 3)]
    cots = [None if c is None else _chk(c, "cot", t) for c, t in zip(cots, tails)]
    B, batched = _resolve_batch((q, 1), (xyz, 2))
    if batched and q.dim() == 1:
        q = q.expand(B, E).contiguous()
    if batched:
        cots = [None if c is None else (c if c.dim() == 3 else c.expand((B,) + tuple(c.shape)).contiguous())
                for c in cots]
    lead = (B,) if batched else ()
    dev = q.device
    q_bar = torch.empty(lead + (E,), dtype=F64, device=dev)
    xyz_bar = torch.empty(lead + (N, 3), dtype=F64, device=dev)
    loads_bar = torch.empty(lead + (N, 3), dtype=F64, device=dev)
    L.check(L.lib().fdm_state_vjp(s.plan, _ptr(q), _ptr(xyz), *[_ptr(c) for c in cots],
                                  _ptr(q_bar), _ptr(xyz_bar), _ptr(loads_bar), B, E, _stream()))
    return q_bar, xyz_bar, loads_bar


def lincomb(x, y=None, z=None, a=1.0, b=1.0, c=1.0, out=None):
    """out = a x + b y + c z (fdm_lincomb): elementwise combinations between kernels, in this library."""
    x = x.contiguous()
    out = torch.empty_like(x) if out is None else out
    _tls.device = x.device
    L.check(L.lib().fdm_lincomb(_ptr(out), _ptr(x), _ptr(None if y is None else y.contiguous()),
                                _ptr(None if z is None else z.contiguous()), float(a), float(b), float(c), x.numel(), _stream()))
    return out


def state_jvp(structure, q, xyz, xyz_dot, q_dot=None, loads_dot=None):
    """JVP of `state` (fdm_state_jvp) -> (vectors_dot, lengths_dot, forces_dot, residuals_dot) for T tangents (leading
    axis of xyz_dot) at ONE primal (q (E,), xyz (N,3)) or at T primals."""
    s = structure
    E, N = s.num_edges, s.num_nodes
    q = _chk(q, "q", (E,))
    xyz = _chk(xyz, "xyz", (N, 3))
    xyz_dot = _chk(xyz_dot, "xyz_dot", (N, 3))
    q_dot = None if q_dot is None else _chk(q_dot, "q_dot", (E,))
    loads_dot = None if loads_dot is None else _chk(loads_dot, "loads_dot", (N, 3))
    T, batched = _resolve_batch((xyz_dot, 2), (q_dot, 1), (loads_dot, 2))
    if batched and xyz_dot.dim() == 2:
        xyz_dot = xyz_dot.expand(T, N, 3).contiguous()
    lead = (T,) if batched else ()
    dev = xyz.device
    vd = torch.empty(lead + (E, 3), dtype=F64, device=dev)
    ld = torch.empty(lead + (E, 1), dtype=F64, device=dev)
    fd = torch.empty(lead + (E, 1), dtype=F64, device=dev)
    rd = torch.empty(lead + (N, 3), dtype=F64, device=dev)
    L.check(L.lib().fdm_state_jvp(s.plan, _ptr(q), _ptr(xyz), _ptr(q_dot), _ptr(xyz_dot), _ptr(loads_dot), _ptr(vd), _ptr(ld),
                                  _ptr(fd), _ptr(rd), T, _stride(q, 1), _stride(xyz, 2), 0 if q_dot is None else _stride(q_dot, 1),
                                  0 if loads_dot is None else _stride(loads_dot, 2), _stream()))
    return vd, ld, fd, rd


def adjoint_grads(structure, q, xyz, lam, q_bar=None, loads_bar=None, xyz_fixed_bar=None):
    """
    K3: gradients of the implicit solve (see include/fdm_b200.h).  When output buffers are
    given they are accumulated into (direct terms + implicit terms); otherwise fresh.
    """
    s = structure
    E, N, Nf, Ns = s.num_edges, s.num_nodes, s.num_free, s.num_supports
    q = _chk(q, "q", (E,))
    xyz = _chk(xyz, "xyz", (N, 3))
    lam = _chk(lam, "lam", (Nf, 3))
    B, batched = _resolve_batch((q, 1), (xyz, 2), (lam, 2))
    if batched and q.dim() == 1:
        q = q.expand(B, E).contiguous()
    lead = (B,) if batched else ()
    dev = q.device
    acc = q_bar is not None or loads_bar is not None or xyz_fixed_bar is not None
    if acc:
        q_bar = torch.zeros(lead + (E,), dtype=F64, device=dev) if q_bar is None else q_bar.contiguous()
        loads_bar = torch.zeros(lead + (N, 3), dtype=F64, device=dev) if loads_bar is None else loads_bar.contiguous()
        xyz_fixed_bar = torch.zeros(lead + (Ns, 3), dtype=F64, device=dev) if xyz_fixed_bar is None else xyz_fixed_bar.contiguous()
    else:
        q_bar = torch.empty(lead + (E,), dtype=F64, device=dev)
        loads_bar = torch.empty(lead + (N, 3), dtype=F64, device=dev)
        xyz_fixed_bar = torch.empty(lead + (Ns, 3), dtype=F64, device=dev)
    L.check(L.lib().fdm_adjoint_grads(s.plan, _ptr(q), _ptr(xyz), _ptr(lam), _ptr(q_bar),
                                      _ptr(loads_bar), _ptr(xyz_fixed_bar), B, E, int(acc), _stream()))
    return q_bar, loads_bar, xyz_fixed_bar


def edge_loads(structure, xyz, loads_nodes, edges_load, is_local=False):
    """nodes_load with tributary edge line loads; is_local: loads follow the edge frame (models.py:290-344, loads.py:296-465)."""
    s = structure
    E, N = s.num_edges, s.num_nodes
    xyz = _chk(xyz, "xyz", (N, 3))
    loads_nodes = _chk(loads_nodes, "loads_nodes", (N, 3))
    edges_load = _chk(edges_load, "edges_load", (E, 3))
    B, batched = _resolve_batch((xyz, 2), (loads_nodes, 2), (edges_load, 2))
    if batched and xyz.dim() == 2:
        xyz = xyz.expand(B, N, 3).contiguous()
    out = torch.empty(((B,) if batched else ()) + (N, 3), dtype=F64, device=xyz.device)
    L.check(L.lib().fdm_edge_loads(s.plan, _ptr(xyz), _ptr(loads_nodes), _ptr(edges_load), _ptr(out), B,
                                   _stride(loads_nodes, 2), _stride(edges_load, 2), int(bool(is_local)), _stream()))
    return out


def edge_loads_vjp(structure, xyz, edges_load, cot, want_xyz=True, want_edges=True, is_local=False):
    """VJP of `edge_loads` -> (xyz_bar, edges_load_bar); the cotangent of loads_nodes is `cot` itself."""
    s = structure
    E, N = s.num_edges, s.num_nodes
    xyz = _chk(xyz, "xyz", (N, 3))
    edges_load = _chk(edges_load, "edges_load", (E, 3))
    cot = _chk(cot, "cot", (N, 3))
    B, batched = _resolve_batch((xyz, 2), (cot, 2))
    if batched and xyz.dim() == 2:
        xyz = xyz.expand(B, N, 3).contiguous()
    if batched and cot.dim() == 2:
        cot = cot.expand(B, N, 3).contiguous()
    lead = (B,) if batched else ()
    xyz_bar = torch.empty(lead + (N, 3), dtype=F64, device=xyz.device) if want_xyz else None
    e_bar = torch.empty(lead + (E, 3), dtype=F64, device=xyz.device) if want_edges else None
    L.check(L.lib().fdm_edge_loads_vjp(s.plan, _ptr(xyz), _ptr(edges_load), _ptr(cot), _ptr(xyz_bar), _ptr(e_bar), B,
                                       _stride(edges_load, 2), int(bool(is_local)), _stream()))
    if e_bar is not None and batched and edges_load.dim() == 2:
        e_bar = e_bar.sum(0)            # one edge load shared by the batch: its cotangent sums over designs
    return xyz_bar, e_bar


def face_loads(structure, xyz, loads_in, faces_load, is_local=False):
    """
    nodes_load with tributary face (area) loads (loads.py:35-71, 80-185, 187-288) -> (loads, aux); `aux` =
    (centroids, face loads in global coordinates) is what `face_loads_vjp` needs from the forward pass.
    """
    s = structure
    N, Fn = s.num_nodes, s.num_faces
    xyz = _chk(xyz, "xyz", (N, 3))
    loads_in = _chk(loads_in, "loads", (N, 3))
    faces_load = _chk(faces_load, "faces_load", (Fn, 3))
    B, batched = _resolve_batch((xyz, 2), (loads_in, 2), (faces_load, 2))
    if batched and xyz.dim() == 2:
        xyz = xyz.expand(B, N, 3).contiguous()
    lead = (B,) if batched else ()
    d = s.device_arrays(xyz.device)
    cen = torch.empty(lead + (Fn, 3), dtype=F64, device=xyz.device)
    gl = torch.empty(lead + (Fn, 3), dtype=F64, device=xyz.device)
    out = torch.empty(lead + (N, 3), dtype=F64, device=xyz.device)
    L.check(L.lib().fdm_face_loads(s.plan, _ptr(d["faces"]), Fn, s.faces_indexed.shape[1], _ptr(d["edges_faces"]),
                                   _ptr(xyz), _ptr(loads_in), _ptr(faces_load), _ptr(cen), _ptr(gl), _ptr(out), B,
                                   _stride(loads_in, 2), _stride(faces_load, 2), int(bool(is_local)), _stream()))
    return out, (cen, gl)


def face_loads_vjp(structure, xyz, faces_load, aux, cot, want_xyz=True, want_faces=True, is_local=False):
    """VJP of `face_loads` -> (xyz_bar, faces_load_bar); the cotangent of loads_in is `cot` itself."""
    s = structure
    N, Fn = s.num_nodes, s.num_faces
    D = s.faces_indexed.shape[1]
    xyz = _chk(xyz, "xyz", (N, 3))
    faces_load = _chk(faces_load, "faces_load", (Fn, 3))
    cot = _chk(cot, "cot", (N, 3))
    centroids = _chk(aux[0], "centroids", (Fn, 3))
    gl = _chk(aux[1], "faces_load_global", (Fn, 3))
    B, batched = _resolve_batch((xyz, 2), (cot, 2))
    if batched and xyz.dim() == 2:
        xyz = xyz.expand(B, N, 3).contiguous()
    if batched and cot.dim() == 2:
        cot = cot.expand(B, N, 3).contiguous()
    if batched and centroids.dim() == 2:
        centroids = centroids.expand(B, Fn, 3).contiguous()
        gl = gl.expand(B, Fn, 3).contiguous()
    lead = (B,) if batched else ()
    d = s.device_arrays(xyz.device)
    cen_bar = torch.empty(lead + (Fn, 3), dtype=F64, device=xyz.device)
    vert_bar = torch.empty(lead + (Fn, D, 3), dtype=F64, device=xyz.device) if is_local else None
    xyz_bar = torch.empty(lead + (N, 3), dtype=F64, device=xyz.device) if want_xyz else None
    fl_bar = torch.empty(lead + (Fn, 3), dtype=F64, device=xyz.device) if want_faces else None
    L.check(L.lib().fdm_face_loads_vjp(s.plan, _ptr(d["faces"]), Fn, D, _ptr(d["edges_faces"]),
                                       _ptr(d["faces_edges"]), _ptr(d["node_faces_ptr"]), _ptr(d["node_faces"]),
                                       _ptr(xyz), _ptr(faces_load), _ptr(centroids), _ptr(gl), _ptr(cot),
                                       _ptr(cen_bar), _ptr(vert_bar), _ptr(xyz_bar), _ptr(fl_bar), B,
                                       _stride(faces_load, 2), int(bool(is_local)), _stream()))
    if fl_bar is not None and batched and faces_load.dim() == 2:
        fl_bar = fl_bar.sum(0)
    return xyz_bar, fl_bar


def loss_sqdist(x, x0):
    """loss_b = 0.5 * sum (x - x0)^2 per design, g = x - x0 (benchmark loss)."""
    x = x.contiguous()
    batched = x.dim() == 3
    B = x.shape[0] if batched else 1
    n = x[0].numel() if batched else x.numel()
    x0 = x0.contiguous()
    x0s = n if (x0.dim() == 3) else 0
    g = torch.empty_like(x)
    loss = torch.empty((B,), dtype=F64, device=x.device)
    L.check(L.lib().fdm_loss_sqdist(_ptr(x), _ptr(x0), _ptr(g), _ptr(loss), n, B, x0s, _stream()))
    return loss, g
