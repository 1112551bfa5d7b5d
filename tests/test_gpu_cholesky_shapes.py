"""
Batched band Cholesky across the shapes its kernels specialise on (solver_chol_warp.cu /
solver_chol.cu): half bandwidth 1 .. 31 (FP64 tensor-core factor kernel up to 30, register
rank-1 kernel at 31, one template per 8 of bandwidth for the substitution kernel), orders that
are / are not multiples of 32 and smaller than one window, batch sizes that do not fill a CTA,
forward solve + adjoint solve with the stored factor, K = +SPD and K = -SPD, and the
one-CTA-per-design kernel for bands wider than 31.  Checked against the CPU oracle (SuperLU).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=torch.device("cuda", 0))


def strip_problem(nx, ny, sign, seed):
    """nx x ny quads, all boundary vertices fixed -> (nx-1)(ny-1) free nodes, half bandwidth min(nx,ny)-1."""
    m = datasets.meshgrid(nx, ny=ny)
    w = nx + 1
    i, j = np.divmod(np.arange(m.nodes.size), w)
    supports = ((i == 0) | (i == ny) | (j == 0) | (j == nx)).astype(np.int64)
    rng = np.random.default_rng(seed)
    q = sign * rng.uniform(0.5, 2.0, size=m.edges.shape[0])
    xyz = m.xyz.copy()
    xyz[:, 2] = 0.3 * np.sin(xyz[:, 0]) * np.cos(xyz[:, 1])
    loads = np.zeros_like(xyz)
    loads[:, 2] = -0.1
    loads[:, 0] = 0.02
    return m, supports, q, xyz[np.flatnonzero(supports)], loads


@pytest.mark.parametrize("nx,ny,batch,sign", [
    (3, 2, 1, 1.0),      # 2 free nodes, hbw 1
    (9, 4, 3, -1.0),     # hbw 3   (subst template 1)
    (40, 9, 5, 1.0),     # hbw 8   (template 2), n = 312
    (33, 17, 9, -1.0),   # hbw 16  (template 3), n = 512 = 16 * 32
    (12, 25, 2, 1.0),    # hbw 11, n = 264
    (31, 40, 11, 1.0),   # hbw 30  (largest tensor-core case)
    (32, 34, 4, -1.0),   # hbw 31  (rank-1 register kernel)
])
def test_band_cholesky_forward_and_stored_factor(nx, ny, batch, sign):
    import jax_fdm_b200.equilibrium as eq

    m, supports, q, xs, loads = strip_problem(nx, ny, sign, seed=nx * 100 + ny)
    s = eq.EquilibriumStructureSparse(m.nodes, m.edges, supports, device=0)
    so = oracle.build_structure(m.nodes, m.edges, supports)
    assert s.half_bandwidth == min(nx, ny) - 1
    rng = np.random.default_rng(7)
    qb = q[None, :] * rng.uniform(0.8, 1.25, size=(batch, q.size))
    K, rhs = eq.ops.assemble(s, T(qb), T(xs), T(loads))
    ws = eq.ops.Workspace()
    x, info = eq.ops.solve(s, K, rhs, solver="cholesky", return_info=True, workspace=ws)
    assert (info[:, 0] == 0).all() and (info[:, 3] == sign).all()
    g = rng.standard_normal((batch, s.num_free, 3))
    lam = eq.ops.solve(s, K, T(g), solver="cholesky", workspace=ws, reuse_factor=True)
    for b in range(batch):
        Kb = oracle.stiffness_matrix_sparse(qb[b], so)
        xo = oracle.sparse_solve(Kb, oracle.load_matrix(qb[b], xs, loads, so))
        lo = oracle.sparse_solve(Kb, g[b])
        assert np.abs(x[b].cpu().numpy() - xo).max() <= 1e-10 * np.abs(xo).max()
        assert np.abs(lam[b].cpu().numpy() - lo).max() <= 1e-10 * np.abs(lo).max()


@pytest.mark.parametrize("nx,ny,batch", [(9, 6, 3), (30, 30, 40), (12, 25, 5), (31, 40, 6)])
def test_mixed_sign_designs_are_solved_with_signed_pivots(nx, ny, batch):
    """Mixed-sign q makes K indefinite; the reference's LU (sparse.py:147-149) still solves it.  The plain
    Cholesky flags those designs and a second pass factors them as L D L^T (hbw <= 30): status 0, sign 0,
    positions and the adjoint solve with the stored factor against the oracle's SuperLU; the definite designs
    of the same batch are untouched (sign +-1)."""
    import jax_fdm_b200.equilibrium as eq

    m, supports, q, xs, loads = strip_problem(nx, ny, 1.0, seed=3 + nx)
    s = eq.EquilibriumStructureSparse(m.nodes, m.edges, supports, device=0)
    so = oracle.build_structure(m.nodes, m.edges, supports)
    assert s.half_bandwidth <= 30
    rng = np.random.default_rng(nx)
    qb = q[None, :] * rng.uniform(0.8, 1.25, size=(batch, q.size))
    mixed = np.arange(batch) % 2 == 1
    flip = np.where(rng.random(q.size) < 0.3, -1.0, 1.0)            # 30 % of the edges in compression
    qb[mixed] *= flip
    qb[0] *= -1.0                                                   # and one all-negative design (K = -SPD)
    K, rhs = eq.ops.assemble(s, T(qb), T(xs), T(loads))
    ws = eq.ops.Workspace()
    x, info = eq.ops.solve(s, K, rhs, solver="cholesky", return_info=True, workspace=ws)
    info = info.cpu().numpy()
    assert (info[:, 0] == 0).all()
    assert (info[mixed, 3] == 0).all() and info[0, 3] == -1 and (info[~mixed, 3] != 0).all()
    # the designs solved with signed pivots have their residual checked and reported (FDM_INFO_RELRES)
    assert (info[mixed, 2] > 0).all() and float(info[mixed, 2].max()) <= 1e-9      # no pivoting: ~1e-10 is what it leaves
    g = rng.standard_normal((batch, s.num_free, 3))
    lam = eq.ops.solve(s, K, T(g), solver="cholesky", workspace=ws, reuse_factor=True)
    for b in list(range(min(batch, 6))) + [batch - 1]:
        Kb = oracle.stiffness_matrix_sparse(qb[b], so)
        xo = oracle.sparse_solve(Kb, oracle.load_matrix(qb[b], xs, loads, so))
        lo = oracle.sparse_solve(Kb, g[b])
        cond_slack = 1e-9 if mixed[b] else 1e-10                    # no pivoting: a little element growth is allowed
        assert np.abs(x[b].cpu().numpy() - xo).max() <= cond_slack * np.abs(xo).max(), b
        assert np.abs(lam[b].cpu().numpy() - lo).max() <= cond_slack * np.abs(lo).max(), b
    # the fused route (K never formed) takes the same second pass: same bits
    x2, info2 = eq.ops.solve_from_q(s, T(qb), T(xs), T(loads), return_info=True)
    i2 = info2.cpu().numpy()
    assert torch.equal(x2, x) and np.array_equal(i2[:, [0, 1, 3]], info[:, [0, 1, 3]])
    # (the reported residual is formed from q there, from K.data here, and summed with atomics: equal to rounding)
    assert np.allclose(i2[:, 2], info[:, 2], rtol=1e-3, atol=1e-13)


def test_indefinite_design_is_flagged_at_half_bandwidth_31_without_the_fallback_workspace():
    """The signed second pass exists for the tensor-core kernels (hbw <= 30); the rank-1 kernel flags the design and
    the band L D L^T (solver_ldlt.cu, tests/test_gpu_ldlt.py) takes it over -- when the caller's workspace has the room
    fdm_workspace_bytes asks for.  A workspace sized for the Cholesky alone keeps the flag."""
    import ctypes as C
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import _lib as L

    m, supports, q, xs, loads = strip_problem(32, 34, 1.0, seed=3)
    s = eq.EquilibriumStructureSparse(m.nodes, m.edges, supports, device=0)
    qb = np.stack([q, q * np.where(np.arange(q.size) % 2, -1.0, 1.0), q])
    K, rhs = eq.ops.assemble(s, T(qb), T(xs), T(loads))
    lib = L.lib()
    need = int(lib.fdm_workspace_bytes(s.plan, 3, L.SOLVERS["cholesky"]))
    per = (s.num_free * (s.half_bandwidth + 1) + 4 * s.num_free) * 8              # solver_ldlt.cu: bytes per slot
    slots = min(3, (256 << 20) // per)
    chol_only = need - (256 + 256 + slots * per)                                  # what the Cholesky alone needs
    assert chol_only > 0
    P = lambda t: C.c_void_p(t.data_ptr())
    x, info = torch.empty_like(rhs), torch.zeros(3, 8, dtype=rhs.dtype, device=rhs.device)
    ws = torch.empty(chol_only, dtype=torch.uint8, device=rhs.device)
    opts = L.SolveOpts(L.SOLVERS["cholesky"], 0, 0.0, 0, 0, 0, 0)
    L.check(lib.fdm_solve(s.plan, P(K), P(rhs), P(x), P(info), P(ws), ws.numel(), 3, C.byref(opts), None))
    st = info[:, 0].cpu().numpy()
    assert st[0] == 0 and st[2] == 0 and st[1] == 2        # FDM_STATUS_INDEFINITE for the mixed-sign design only


def test_band_too_wide_is_refused_and_auto_uses_pcg():
    """hbw = 32, n = 1024: neither band kernel applies (295 KB of band > shared memory): the Cholesky
    solver returns FDM_ERR_UNSUPPORTED and the automatic choice is an iterative solver."""
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200._lib import FdmError

    m, supports, q, xs, loads = strip_problem(33, 33, 1.0, seed=1)
    s = eq.EquilibriumStructureSparse(m.nodes, m.edges, supports, device=0)
    so = oracle.build_structure(m.nodes, m.edges, supports)
    K, rhs = eq.ops.assemble(s, T(q), T(xs), T(loads))
    with pytest.raises(FdmError) as e:
        eq.ops.solve(s, K, rhs, solver="cholesky")
    assert e.value.code == -5
    x, info = eq.ops.solve(s, K, rhs, solver="auto", tol=1e-13, return_info=True)
    xo = oracle.sparse_solve(oracle.stiffness_matrix_sparse(q, so), oracle.load_matrix(q, xs, loads, so))
    assert info[0, 0] == 0 and np.abs(x.cpu().numpy() - xo).max() <= 1e-10 * np.abs(xo).max()


def test_fused_assembly_is_bit_identical_to_assemble_then_solve():
    """fdm_solve_from_q (K1 + K2 in one kernel, K.data never formed) against fdm_assemble + fdm_solve:
    same additions in the same order, so the positions must agree bit for bit; the adjoint solve then
    runs on the stored factor without K."""
    import jax_fdm_b200.equilibrium as eq

    for nx, sign in ((30, -1.0), (12, 1.0)):
        prob = datasets.pillow_problem(nx)
        s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
        rng = np.random.default_rng(nx)
        B = 37
        qb = sign * np.abs(datasets.pillow_batch_q(prob.edges.shape[0], 0, B)) * rng.uniform(0.7, 1.4, size=(B, 1))
        xs = prob.xyz_fixed + 0.05 * rng.standard_normal(prob.xyz_fixed.shape)
        q, xs_t, p = T(qb), T(xs), T(prob.loads)
        K, rhs = eq.ops.assemble(s, q, xs_t, p)
        x_ref = eq.ops.solve(s, K, rhs, solver="cholesky")
        ws = eq.ops.Workspace()
        x, info = eq.ops.solve_from_q(s, q, xs_t, p, workspace=ws, return_info=True)
        assert (info[:, 0] == 0).all() and (info[:, 3] == sign).all()
        assert torch.equal(x, x_ref)
        g = T(rng.standard_normal((B, s.num_free, 3)))
        lam = eq.ops.solve(s, None, g, solver="cholesky", workspace=ws, reuse_factor=True)
        lam_ref = eq.ops.solve(s, K, g, solver="cholesky")     # forward substitution fused in the factor kernel:
        err = (lam - lam_ref).abs().max() / lam_ref.abs().max()  # another summation order, not bit-identical
        assert err <= 1e-12


def test_fused_assembly_with_per_design_supports_and_loads():
    """fdm_solve_from_q with (B,Ns,3) support positions and (B,N,3) loads (non-zero strides), and a structure whose
    boundary rows have several support neighbours: bit-identical to fdm_assemble + fdm_solve."""
    import jax_fdm_b200.equilibrium as eq

    for prob in (datasets.pillow_problem(9), datasets.cablenet_problem(7, seed=3)):
        s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
        rng = np.random.default_rng(5)
        B = 19
        qb = prob.q[None, :] * rng.uniform(0.7, 1.4, size=(B, prob.q.size))
        xs = prob.xyz_fixed[None] + 0.05 * rng.standard_normal((B,) + prob.xyz_fixed.shape)
        ld = prob.loads[None] + 0.02 * rng.standard_normal((B,) + prob.loads.shape)
        q, xs_t, p = T(qb), T(xs), T(ld)
        K, rhs = eq.ops.assemble(s, q, xs_t, p)
        x_ref = eq.ops.solve(s, K, rhs, solver="cholesky")
        x = eq.ops.solve_from_q(s, q, xs_t, p)
        assert x is not None and torch.equal(x, x_ref)
