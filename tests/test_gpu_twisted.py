"""
GPU: the two-sided ("twisted") band factorisation -- two warps per design, one factoring [T | M] top-down, the
other B bottom-up, meeting on the middle block M -- that small batches take (batch <= 8 x SM count, half bandwidth
16..29).  Same L D L^T / L L^T arithmetic as the one-sided kernel up to the order of the updates of M, so results
agree with the oracle (1e-10 / 1e-8) and with the one-sided kernel (the same designs inside a batch large enough to
take it) to rounding.  Covered: both entry points (fdm_solve on K.data, fdm_solve_from_q), the adjoint solve with
the stored two-sided factor (forward + backward sweeps), odd block counts on either side, a mixed-sign design
(signed second pass), NaN-on-failure, and the whole forward + adjoint pipeline.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402

POS_RTOL, GRAD_RTOL = 1e-10, 1e-8
BIG = 1400            # > 8 x 148: the one-sided kernel


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda:0")


def N(t):
    return t.detach().cpu().numpy()


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("nx", [30, 24, 27, 18])       # hbw 29 / 23 / 26 / 17; 27, 17, 22, 10 blocks
def test_two_sided_solve_matches_oracle_and_one_sided(nx):
    import jax_fdm_b200.equilibrium as eq
    prob = datasets.pillow_problem(nx)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    assert 16 <= s.half_bandwidth <= 29
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    B = 7
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 11, BIG)
    q, xs, p = T(qb), T(prob.xyz_fixed), T(prob.loads)
    rng = np.random.default_rng(nx)
    gcot = rng.standard_normal((BIG, s.num_free, 3))
    out = {}
    for label, nb in (("two", B), ("one", BIG)):
        ws = eq.ops.Workspace()
        x = eq.ops.solve_from_q(s, q[:nb].contiguous(), xs, p, workspace=ws)
        lam = eq.ops.solve(s, None, T(gcot[:nb]), solver="cholesky", workspace=ws, reuse_factor=True)
        K, rhs = eq.ops.assemble(s, q[:nb].contiguous(), xs, p)
        ws2 = eq.ops.Workspace()
        x2, info = eq.ops.solve(s, K, rhs, solver="cholesky", workspace=ws2, return_info=True)
        lam2 = eq.ops.solve(s, K, T(gcot[:nb]), solver="cholesky", workspace=ws2, reuse_factor=True)
        assert np.all(N(info)[:, 0] == 0)
        out[label] = [N(t)[:B] for t in (x, lam, x2, lam2)]
    for b in range(B):
        Ko = oracle.stiffness_matrix_sparse(qb[b], so)
        xo = oracle.sparse_solve(Ko, oracle.load_matrix(qb[b], prob.xyz_fixed, prob.loads, so))
        lo = oracle.sparse_solve(Ko, gcot[b])
        for x, lam in ((out["two"][0][b], out["two"][1][b]), (out["two"][2][b], out["two"][3][b])):
            assert relerr(x, xo) <= POS_RTOL and relerr(lam, lo) <= POS_RTOL
    for a, c in zip(out["two"], out["one"]):
        assert relerr(a, c) <= 1e-12


def test_two_sided_pipeline_and_mixed_sign_design():
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import NativeForwardAdjoint
    prob = datasets.pillow_problem(30)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    B = 9
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    rng = np.random.default_rng(2)
    qb[3] = np.where(rng.random(qb.shape[1]) < 0.15, 0.7, -2.0) * (1.0 + 0.05 * rng.standard_normal(qb.shape[1]))   # indefinite K
    qb[6] = 0.0                                                                                         # no factorisation
    x0 = prob.xyz0[s.indices_free]
    fa = NativeForwardAdjoint(s, B, chunks=2, solver="cholesky")
    fa.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    h_loss, h_q_bar = fa()
    info = N(fa.array("info_f"))
    assert info[6, 0] != 0 and np.isnan(h_loss[6]) and np.isnan(N(fa.array("x"))[6]).all()
    assert info[3, 0] == 0 and info[3, 3] == 0.0          # solved with signed pivots
    x = N(fa.array("x"))

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    for b in (0, 3, 4, 5, 8):
        ref = oracle.forward_adjoint(qb[b], prob.xyz_fixed, prob.loads, so, cot)
        tol_x, tol_g = (1e-8, 1e-6) if b == 3 else (POS_RTOL, GRAD_RTOL)      # unpivoted LDL^T on an indefinite K
        assert relerr(x[b], ref["x_free"]) <= tol_x, b
        assert relerr(h_q_bar[b], ref["q_bar"]) <= tol_g, b
    fa.close()
