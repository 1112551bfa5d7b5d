"""
GPU: the library-owned forward + adjoint object (`fdm_fa_*`, csrc/pipeline.cu) through its ctypes wrapper
`NativeForwardAdjoint` -- the end-to-end call of bench.py.  Checked against the CPU oracle (positions 1e-10,
gradients 1e-8): device-resident step and host-in / host-out call, several chunkings (uneven slices included),
the quadratic benchmark loss and a compiled goal program (against the oracle's loss and the autograd route of the
model API), a single large mesh on the iterative solver, and NaN-on-failure for a design that cannot be factored.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import goals as og  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402

POS_RTOL, GRAD_RTOL = 1e-10, 1e-8


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def N(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("chunks", [1, 3, 16])
def test_quadratic_loss_device_step_and_host_call(chunks):
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import NativeForwardAdjoint

    prob = datasets.pillow_problem(12)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    B = 50
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    x0 = prob.xyz0[s.indices_free]
    fa = NativeForwardAdjoint(s, B, chunks=chunks, solver="cholesky")
    assert fa.chunks == chunks
    fa.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    for _ in range(3):
        fa.step()
    fa.stream.synchronize()
    assert fa.launches_per_step > 0
    assert bool((fa.array("info_f")[:, 0] == 0).all()) and bool((fa.array("info_a")[:, 0] == 0).all())
    dev = {k: N(fa.array(k)) for k in ("x", "loss", "q_bar", "loads_bar", "xs_bar", "xyz", "lengths")}
    # end to end with a DIFFERENT q: the host call must not reuse what the device step left behind
    q2 = qb[::-1].copy()
    fa.h_q[...] = q2
    h_loss, h_q_bar = fa()
    h_loss, h_q_bar = h_loss.copy(), h_q_bar.copy()

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    for b in (0, 1, 16, 17, 33, B - 1):
        ref = oracle.forward_adjoint(qb[b], prob.xyz_fixed, prob.loads, so, cot)
        lref = 0.5 * np.sum((ref["x_free"] - x0) ** 2)
        assert relerr(dev["x"][b], ref["x_free"]) <= POS_RTOL
        assert relerr(dev["xyz"][b], ref["state"]["xyz"]) <= POS_RTOL
        assert relerr(dev["lengths"][b], ref["state"]["lengths"].reshape(-1)) <= POS_RTOL
        assert abs(dev["loss"][b] - lref) <= POS_RTOL * lref
        assert relerr(dev["q_bar"][b], ref["q_bar"]) <= GRAD_RTOL
        assert relerr(dev["loads_bar"][b], ref["loads_bar"]) <= GRAD_RTOL
        assert relerr(dev["xs_bar"][b], ref["xyz_fixed_bar"]) <= GRAD_RTOL
        # the reversed batch through the host call: design b there is design B-1-b here
        assert abs(h_loss[B - 1 - b] - lref) <= POS_RTOL * lref
        assert relerr(h_q_bar[B - 1 - b], ref["q_bar"]) <= GRAD_RTOL
    fa.close()


def test_auto_slices_and_side_stream_give_the_bits_of_one_slice():
    """`chunks="auto"` cuts the batch into slices of ~512 designs, each on its own stream with positions + state on a
    side stream and the factor kernels at the low launch priority (csrc/pipeline.cu: prioritise).  Scheduling only:
    every output must be bit-identical to the one-slice step, also after repeated steps (the side stream of step k+1
    must not overtake the gradient kernel of step k, which reads xyz)."""
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import NativeForwardAdjoint

    prob = datasets.pillow_problem(30)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    B = 1536 + 7
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    x0 = prob.xyz0[s.indices_free]
    names = ("x", "xyz", "loss", "q_bar", "loads_bar", "xs_bar", "lengths", "forces", "residuals", "vectors")
    out = {}
    for chunks in (1, "auto"):
        fa = NativeForwardAdjoint(s, B, chunks=chunks, solver="cholesky")
        assert fa.chunks == (1 if chunks == 1 else B // 512)
        fa.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
        for _ in range(4):
            fa.step()
        fa.stream.synchronize()
        assert bool((fa.array("info_f")[:, 0] == 0).all()) and bool((fa.array("info_a")[:, 0] == 0).all())
        out[chunks] = {k: fa.array(k).clone() for k in names}
        fa.close()
    for k in names:
        assert torch.equal(out[1][k], out["auto"][k]), k


def test_goal_program_route_matches_oracle_and_autograd():
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import goals as G, losses as LS
    from jax_fdm_b200.pipeline import NativeForwardAdjoint

    prob = datasets.pillow_problem(10)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    B = 12
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 100, B)
    free = np.flatnonzero(prob.supports == 0)
    keys = [(int(u), int(v)) for u, v in prob.edges]
    pts = prob.xyz0[free[:7]] + np.array([0.0, 0.0, 0.4])
    loss = LS.Loss(LS.MeanSquaredError([G.EdgeLengthGoal(k, 0.9) for k in keys[::3]]),
                   LS.SquaredError([G.NodePointGoal(int(prob.nodes[n]), t) for n, t in zip(free[:7], pts)], alpha=0.5),
                   LS.PredictionError([G.NetworkLoadPathGoal()], alpha=1e-3),
                   LS.AbsoluteError([G.NodeResidualForceGoal(int(prob.nodes[np.flatnonzero(prob.supports)[2]]), 0.3)]))
    fa = NativeForwardAdjoint(s, B, chunks=3, solver="cholesky")
    fa.set_inputs(qb, prob.xyz_fixed, prob.loads)
    fa.set_loss(loss)
    h_loss, h_q_bar = fa()
    h_loss, h_q_bar = h_loss.copy(), h_q_bar.copy()
    fa.step()
    fa.stream.synchronize()
    assert np.array_equal(N(fa.array("loss")), h_loss) and np.array_equal(N(fa.array("q_bar")), h_q_bar)
    loads_bar, xs_bar = N(fa.array("loads_bar")), N(fa.array("xs_bar"))
    # the autograd route of the model API (what optimization.py drives)
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda:0")
    q = T(qb).requires_grad_(True)
    xs = T(prob.xyz_fixed).requires_grad_(True)
    p = T(prob.loads).requires_grad_(True)
    model = eq.EquilibriumModelSparse(tmax=1)
    val = loss(eq.EquilibriumParametersState(q, xs, eq.LoadState(p, 0.0, 0.0)), model, s)
    val.sum().backward()
    assert relerr(h_loss, N(val)) <= 1e-12
    assert relerr(h_q_bar, N(q.grad)) <= 1e-10
    assert relerr(loads_bar.sum(0), N(p.grad)) <= 1e-10 and relerr(xs_bar.sum(0), N(xs.grad)) <= 1e-10
    # and the oracle's loss, with a finite-difference check of q_bar on one design
    sup2 = int(np.flatnonzero(prob.supports)[2])
    oterms = [("MeanSquaredError", [("EdgeLengthGoal", i, 0.9, 1.0) for i in range(0, len(keys), 3)], 1.0),
              ("SquaredError", [("NodePointGoal", int(n), t, 1.0) for n, t in zip(free[:7], pts)], 0.5),
              ("PredictionError", [("NetworkLoadPathGoal", 0, 0.0, 1.0)], 1e-3),
              ("AbsoluteError", [("NodeResidualForceGoal", sup2, 0.3, 1.0)], 1.0)]

    def f(qv):
        st, _ = oracle.forward(qv, prob.xyz_fixed, prob.loads, so)
        return og.loss(oterms, st)

    for b in (0, 5, B - 1):
        assert abs(h_loss[b] - f(qb[b])) <= 1e-10 * abs(h_loss[b])
    rng = np.random.default_rng(1)
    scale = np.abs(h_q_bar[5]).max()
    from conftest import fd5
    for k in rng.choice(qb.shape[1], size=6, replace=False):
        fd = fd5(f, qb[5], k, h=1e-4)
        assert abs(h_q_bar[5, k] - fd) <= 1e-8 * scale, (k, h_q_bar[5, k], fd)
    fa.close()


def test_single_cablenet_on_the_iterative_solver():
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import NativeForwardAdjoint

    prob = datasets.cablenet_problem(60, seed=2)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    x0 = prob.xyz0[s.indices_free]
    fa = NativeForwardAdjoint(s, 1, solver="auto", tol=1e-13)
    fa.set_inputs(prob.q[None], prob.xyz_fixed, prob.loads, x0)
    h_loss, h_q_bar = fa()

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    ref = oracle.forward_adjoint(prob.q, prob.xyz_fixed, prob.loads, so, cot)
    assert N(fa.array("info_f"))[0, 0] == 0 and N(fa.array("info_a"))[0, 0] == 0
    assert relerr(N(fa.array("x"))[0], ref["x_free"]) <= POS_RTOL
    assert relerr(h_q_bar[0], ref["q_bar"]) <= GRAD_RTOL
    fa.close()


def test_failed_design_comes_back_as_nan_and_others_are_untouched():
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import NativeForwardAdjoint

    prob = datasets.pillow_problem(8)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    B = 9
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    good = NativeForwardAdjoint(s, B, chunks=2, solver="cholesky")
    good.set_inputs(qb, prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
    l0, g0 = good()
    l0, g0 = l0.copy(), g0.copy()
    qb[4] = 0.0                                       # K = 0: no factorisation exists
    bad = NativeForwardAdjoint(s, B, chunks=2, solver="cholesky")
    bad.set_inputs(qb, prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
    l1, g1 = bad()
    assert N(bad.array("info_f"))[4, 0] != 0
    touches_free = (prob.supports[prob.edges] == 0).any(axis=1)      # support-support edges have a structural zero
    assert np.isnan(l1[4]) and np.isnan(g1[4][touches_free]).all() and np.isnan(N(bad.array("x"))[4]).all()
    keep = [b for b in range(B) if b != 4]
    assert np.array_equal(l1[keep], l0[keep]) and np.array_equal(g1[keep], g0[keep])
    good.close(); bad.close()


@pytest.mark.parametrize("chunks", [1, 4])
def test_steps_enqueued_back_to_back_keep_their_own_losses(chunks):
    """fdm_fa_enqueue_host twice without a wait in between, different q and different host buffers per step: the
    per-design losses leave in ONE copy at the end of a step, parked in a staging row so that the next step's kernels
    cannot overwrite them first.  Both steps must come back exactly as two synchronous steps do."""
    import ctypes as C
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200 import _lib as L
    from jax_fdm_b200.pipeline import NativeForwardAdjoint

    prob = datasets.pillow_problem(12)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    B, E = 64, prob.edges.shape[0]
    fa = NativeForwardAdjoint(s, B, chunks=chunks, solver="cholesky")
    q1 = datasets.pillow_batch_q(E, 0, B)
    q2 = datasets.pillow_batch_q(E, 1000, B)
    fa.set_inputs(q1, prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
    ref = []
    for q in (q1, q2, q1):
        fa.h_q[...] = q
        loss, q_bar = fa()
        ref.append((loss.copy(), q_bar.copy()))
    assert not np.array_equal(ref[0][0], ref[1][0])
    lib = fa.lib
    bufs = []
    for q in (q1, q2, q1):                                  # three steps: the staging rows alternate and are reused
        hq, hl, hb = fa._host((B, E)), fa._host((B,)), fa._host((B, E))
        hq[...] = q
        hl[...] = -1.0
        bufs.append((hq, hl, hb))
    for hq, hl, hb in bufs:
        L.check(lib.fdm_fa_enqueue_host(fa.handle, hq.ctypes.data, hl.ctypes.data, hb.ctypes.data))
    L.check(lib.fdm_fa_wait(fa.handle))
    for (hq, hl, hb), (rl, rb) in zip(bufs, ref):
        assert np.array_equal(hl, rl) and np.array_equal(hb, rb)
    fa.close()
