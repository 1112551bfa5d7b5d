"""
Band L D L^T in global memory (solver_ldlt.cu): the direct solver behind mixed-sign q (indefinite K) wherever the
other solvers stop with FDM_STATUS_INDEFINITE -- one larger mesh (the PCG solvers), batches of wide-band designs (the
shared-memory band Cholesky), half bandwidth 31 (the rank-1 warp kernel).  The reference's LU solves such systems
(sparse.py:147-149); checked against the CPU oracle (SuperLU), forward solve and adjoint solve with the stored factor.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from test_gpu_cholesky_shapes import T, strip_problem  # noqa: E402


def _mixed(q, rng, frac=0.3):
    return q * np.where(rng.random(q.size) < frac, -1.0, 1.0)


@pytest.mark.parametrize("nx,ny", [(12, 9), (30, 30), (40, 45)])
def test_explicit_ldlt_matches_the_oracle_on_definite_and_indefinite_designs(nx, ny):
    import jax_fdm_b200.equilibrium as eq

    m, supports, q, xs, loads = strip_problem(nx, ny, 1.0, seed=nx)
    s = eq.EquilibriumStructureSparse(m.nodes, m.edges, supports, device=0)
    so = oracle.build_structure(m.nodes, m.edges, supports)
    rng = np.random.default_rng(ny)
    qb = np.stack([q, -q, _mixed(q, rng), _mixed(q, rng, 0.5)])
    K, rhs = eq.ops.assemble(s, T(qb), T(xs), T(loads))
    ws = eq.ops.Workspace()
    x, info = eq.ops.solve(s, K, rhs, solver="ldlt", return_info=True, workspace=ws)
    assert (info[:, 0] == 0).all() and (info[:, 3] == 0).all()
    assert float(info[:, 2].max()) <= 1e-12               # FDM_INFO_RELRES: ||b - K x|| / ||b|| after the refinement step
    g = rng.standard_normal((4, s.num_free, 3))
    lam = eq.ops.solve(s, K, T(g), solver="ldlt", workspace=ws, reuse_factor=True)
    for b in range(4):
        Kb = oracle.stiffness_matrix_sparse(qb[b], so)
        xo = oracle.sparse_solve(Kb, oracle.load_matrix(qb[b], xs, loads, so))
        lo = oracle.sparse_solve(Kb, g[b])
        tol = 1e-10                                        # also on the indefinite ones: one refinement step
        assert np.abs(x[b].cpu().numpy() - xo).max() <= tol * np.abs(xo).max(), b
        assert np.abs(lam[b].cpu().numpy() - lo).max() <= tol * np.abs(lo).max(), b


def test_half_bandwidth_31_mixed_sign_design_is_solved_behind_the_rank1_kernel():
    import jax_fdm_b200.equilibrium as eq

    m, supports, q, xs, loads = strip_problem(32, 34, 1.0, seed=3)
    s = eq.EquilibriumStructureSparse(m.nodes, m.edges, supports, device=0)
    so = oracle.build_structure(m.nodes, m.edges, supports)
    assert s.half_bandwidth == 31
    qb = np.stack([q, q * np.where(np.arange(q.size) % 2, -1.0, 1.0), q])
    K, rhs = eq.ops.assemble(s, T(qb), T(xs), T(loads))
    ws = eq.ops.Workspace()
    x, info = eq.ops.solve(s, K, rhs, solver="cholesky", return_info=True, workspace=ws, nan_on_failure=True)
    info = info.cpu().numpy()
    assert (info[:, 0] == 0).all() and info[1, 3] == 0 and info[0, 3] == 1 and info[2, 3] == 1
    g = np.random.default_rng(0).standard_normal((3, s.num_free, 3))
    lam = eq.ops.solve(s, K, T(g), solver="cholesky", workspace=ws, reuse_factor=True, nan_on_failure=True)
    for b in range(3):
        Kb = oracle.stiffness_matrix_sparse(qb[b], so)
        xo = oracle.sparse_solve(Kb, oracle.load_matrix(qb[b], xs, loads, so))
        lo = oracle.sparse_solve(Kb, g[b])
        tol = 1e-10
        assert np.abs(x[b].cpu().numpy() - xo).max() <= tol * np.abs(xo).max(), b
        assert np.abs(lam[b].cpu().numpy() - lo).max() <= tol * np.abs(lo).max(), b
    # and the model level no longer hands NaN to the caller for that design
    model = eq.EquilibriumModelSparse(tmax=1, solver="cholesky")
    state = model(eq.EquilibriumParametersState(T(qb), T(xs), eq.LoadState(T(loads), 0.0, 0.0)), s)
    assert torch.isfinite(state.xyz).all()


def test_wide_band_batch_and_single_mesh_fall_back_design_by_design():
    """hbw 44: a batch takes the one-CTA-per-design Cholesky, a single design the persistent PCG; both flag the
    mixed-sign design and the band L D L^T solves it; the definite designs keep their solver's answer."""
    import jax_fdm_b200.equilibrium as eq

    m, supports, q, xs, loads = strip_problem(45, 50, 1.0, seed=5)
    s = eq.EquilibriumStructureSparse(m.nodes, m.edges, supports, device=0)
    so = oracle.build_structure(m.nodes, m.edges, supports)
    assert s.half_bandwidth > 31
    rng = np.random.default_rng(1)
    qb = np.stack([q, _mixed(q, rng), 1.1 * q, _mixed(q, rng, 0.4), -q])
    K, rhs = eq.ops.assemble(s, T(qb), T(xs), T(loads))
    ws = eq.ops.Workspace()
    x, info = eq.ops.solve(s, K, rhs, solver="auto", tol=1e-13, return_info=True, workspace=ws)
    info = info.cpu().numpy()
    assert (info[:, 0] == 0).all() and (info[[1, 3], 3] == 0).all() and (info[[0, 2], 3] == 1).all() and info[4, 3] == -1
    g = rng.standard_normal((5, s.num_free, 3))
    lam = eq.ops.solve(s, K, T(g), solver="auto", tol=1e-13, workspace=ws, reuse_factor=True)
    for b in range(5):
        Kb = oracle.stiffness_matrix_sparse(qb[b], so)
        xo = oracle.sparse_solve(Kb, oracle.load_matrix(qb[b], xs, loads, so))
        lo = oracle.sparse_solve(Kb, g[b])
        tol = 1e-10
        assert np.abs(x[b].cpu().numpy() - xo).max() <= tol * np.abs(xo).max(), b
        assert np.abs(lam[b].cpu().numpy() - lo).max() <= tol * np.abs(lo).max(), b
    # one design at a time: the iterative solvers in front
    for b in (1, 2):
        x1, i1 = eq.ops.solve(s, K[b], rhs[b], solver="auto", tol=1e-13, return_info=True)
        Kb = oracle.stiffness_matrix_sparse(qb[b], so)
        xo = oracle.sparse_solve(Kb, oracle.load_matrix(qb[b], xs, loads, so))
        assert i1.reshape(-1)[0] == 0
        assert np.abs(x1.cpu().numpy() - xo).max() <= 1e-10 * np.abs(xo).max()


def test_mixed_sign_q_through_the_model_and_its_gradient():
    """fdm-level: a mixed-sign design on a wide-band mesh, value and gradient of a loss against the oracle."""
    import jax_fdm_b200.equilibrium as eq

    m, supports, q, xs, loads = strip_problem(36, 40, 1.0, seed=7)
    s = eq.EquilibriumStructureSparse(m.nodes, m.edges, supports, device=0)
    so = oracle.build_structure(m.nodes, m.edges, supports)
    rng = np.random.default_rng(2)
    qm = _mixed(q, rng, 0.25)
    x0 = rng.standard_normal((s.num_free, 3))
    q_t = T(qm).requires_grad_(True)
    model = eq.EquilibriumModelSparse(tmax=1)
    state = model(eq.EquilibriumParametersState(q_t, T(xs), eq.LoadState(T(loads), 0.0, 0.0)), s)
    free = torch.as_tensor(np.asarray(s.indices_free), device=state.xyz.device)
    loss = 0.5 * ((state.xyz[free] - T(x0)) ** 2).sum()
    loss.backward()

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}
    ref = oracle.forward_adjoint(qm, xs, loads, so, cot)
    ref_loss = 0.5 * ((ref["x_free"] - x0) ** 2).sum()
    assert abs(loss.item() - ref_loss) <= 1e-8 * abs(ref_loss)
    assert np.abs(q_t.grad.cpu().numpy() - ref["q_bar"]).max() <= 1e-7 * np.abs(ref["q_bar"]).max()


def test_mixed_sign_q_on_the_200x200_net():
    """BASELINE config 4's mesh with a patch of edges in compression: the persistent PCG flags the design, the band
    L D L^T (40 000 columns, half bandwidth ~200) solves it; forward and adjoint solve against SuperLU."""
    import time
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import datasets

    prob = datasets.cablenet_problem(200)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    q = prob.q.copy()
    mid = prob.xyz0[prob.edges].mean(axis=1)
    patch = (np.abs(mid[:, 0] - 5.0) < 1.0) & (np.abs(mid[:, 1] - 5.0) < 1.0)
    q[patch] *= -0.5
    assert patch.sum() > 100
    K, rhs = eq.ops.assemble(s, T(q), T(prob.xyz_fixed), T(prob.loads))
    ws = eq.ops.Workspace()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, info = eq.ops.solve(s, K, rhs, solver="auto", return_info=True, workspace=ws)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    info = info.cpu().numpy().reshape(-1)
    assert info[0] == 0 and info[3] == 0, info
    ws2 = eq.ops.Workspace()
    t0 = time.perf_counter()
    x2 = eq.ops.solve(s, K, rhs, solver="ldlt", workspace=ws2)
    torch.cuda.synchronize()
    dt_direct = time.perf_counter() - t0
    assert torch.equal(x2, x)
    t0 = time.perf_counter()
    _, i3 = eq.ops.solve(s, K, rhs, solver="pcg_persistent", return_info=True, workspace=eq.ops.Workspace())
    torch.cuda.synchronize()
    print(f"ldlt alone {dt_direct * 1e3:.1f} ms; persistent PCG attempt {1e3 * (time.perf_counter() - t0):.1f} ms, status {i3.reshape(-1)[0].item()}")
    g = np.random.default_rng(0).standard_normal((s.num_free, 3))
    lam = eq.ops.solve(s, K, T(g), solver="auto", workspace=ws, reuse_factor=True)
    Kq = oracle.stiffness_matrix_sparse(q, so)
    xo = oracle.sparse_solve(Kq, oracle.load_matrix(q, prob.xyz_fixed, prob.loads, so))
    lo = oracle.sparse_solve(Kq, g)
    ex = np.abs(x.cpu().numpy() - xo).max() / np.abs(xo).max()
    el = np.abs(lam.cpu().numpy() - lo).max() / np.abs(lo).max()
    print(f"200x200 mixed-sign: solve {dt * 1e3:.1f} ms (cycles: load {info[4]:.3g}, factor {info[5]:.3g}, substitute {info[6]:.3g}), rel err x {ex:.2e}, lam {el:.2e}")
    assert ex <= 1e-10 and el <= 1e-10
