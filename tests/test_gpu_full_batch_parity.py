"""
GPU: parity AT THE BENCHMARK'S OWN SIZE.  BASELINE config 5 -- 4096 pillow 30x30 designs -- through the very object
bench.py times (`NativeForwardAdjoint(chunks="auto")`: eight slices of 512 designs on their own streams inside one
fork/join CUDA graph, factor kernels at the low launch priority, positions + state on side streams, the factorisation
assembling its rows from q) and through the Python-level `ChunkedForwardAdjoint` (two slices / one), then 64+ sampled
designs against the CPU oracle:
positions 1e-10, loss 1e-10, gradients (q, loads, supports) 1e-8 -- relative, max norm.  The sample covers the
first and last design, both sides of the slice boundary (2047 | 2048), the designs around 2368 (the number of
designs resident at once in a slice-less launch of the one-warp-per-design factor kernel: 16 warps x 148 SMs, where
the tail wave starts), the ends of the first and last 8-design CTAs, both sides of every 512-design slice boundary, and a seeded random rest.
All 4096 status words must be OK and every output finite.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402

B = 4096
POS_RTOL, GRAD_RTOL = 1e-10, 1e-8


def sample_designs(batch, count=64, seed=0):
    fixed = [0, 1, 7, 8, batch // 2 - 1, batch // 2, batch // 2 + 7, batch // 2 + 8, 2367, 2368, 2369, 2375, 2376,
             batch - 9, batch - 8, batch - 2, batch - 1]
    fixed += [k * 512 + d for k in range(1, 8) for d in (-1, 0)]           # the slice boundaries of chunks="auto"
    fixed = [b for b in fixed if 0 <= b < batch]
    rng = np.random.default_rng(seed)
    rest = rng.choice(batch, size=max(0, count - len(fixed)), replace=False)
    return sorted(set(fixed) | set(int(b) for b in rest))


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("chunks", [2, 1])
def test_full_batch_sampled_against_the_oracle(chunks):
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import ChunkedForwardAdjoint

    prob = datasets.pillow_problem(30)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    x0 = prob.xyz0[s.indices_free]
    pipe = ChunkedForwardAdjoint(s, B, chunks=chunks, solver="cholesky", device=0)
    pipe.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    for _ in range(3):                                   # the third call is a CUDA-graph replay
        pipe.step()
    pipe.stream.synchronize()
    for part in pipe.parts:
        assert bool((part.info_f[:, 0] == 0).all()) and bool((part.info_a[:, 0] == 0).all())
    N = lambda t: t.detach().cpu().numpy()
    x = np.concatenate([N(p.x) for p in pipe.parts])
    q_bar = np.concatenate([N(p.q_bar) for p in pipe.parts])
    loads_bar = np.concatenate([N(p.loads_bar) for p in pipe.parts])
    xs_bar = np.concatenate([N(p.xs_bar) for p in pipe.parts])
    loss = N(pipe.loss)
    assert x.shape[0] == B and np.isfinite(x).all() and np.isfinite(q_bar).all()

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    worst = {"x": 0.0, "loss": 0.0, "q": 0.0, "p": 0.0, "xs": 0.0}
    for b in sample_designs(B):
        ref = oracle.forward_adjoint(qb[b], prob.xyz_fixed, prob.loads, so, cot)
        lref = 0.5 * np.sum((ref["x_free"] - x0) ** 2)
        worst["x"] = max(worst["x"], relerr(x[b], ref["x_free"]))
        worst["loss"] = max(worst["loss"], abs(loss[b] - lref) / lref)
        worst["q"] = max(worst["q"], relerr(q_bar[b], ref["q_bar"]))
        worst["p"] = max(worst["p"], relerr(loads_bar[b], ref["loads_bar"]))
        worst["xs"] = max(worst["xs"], relerr(xs_bar[b], ref["xyz_fixed_bar"]))
    assert worst["x"] <= POS_RTOL and worst["loss"] <= POS_RTOL, worst
    assert worst["q"] <= GRAD_RTOL and worst["p"] <= GRAD_RTOL and worst["xs"] <= GRAD_RTOL, worst


def test_full_batch_native_auto_slices_against_the_oracle():
    """The object bench.py times: fdm_fa_step_device on 8 slices (chunks="auto"), every output array of the step."""
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import NativeForwardAdjoint

    prob = datasets.pillow_problem(30)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    x0 = prob.xyz0[s.indices_free]
    fa = NativeForwardAdjoint(s, B, chunks="auto", solver="cholesky")
    assert fa.chunks == 8
    fa.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    for _ in range(3):
        fa.step()
    fa.stream.synchronize()
    assert bool((fa.array("info_f")[:, 0] == 0).all()) and bool((fa.array("info_a")[:, 0] == 0).all())
    out = {k: fa.array(k).cpu().numpy() for k in ("x", "loss", "q_bar", "loads_bar", "xs_bar", "lengths", "forces",
                                                  "residuals")}
    assert all(np.isfinite(v).all() for v in out.values())

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    worst = {"x": 0.0, "loss": 0.0, "q": 0.0, "p": 0.0, "xs": 0.0, "state": 0.0}
    for b in sample_designs(B, count=72):
        ref = oracle.forward_adjoint(qb[b], prob.xyz_fixed, prob.loads, so, cot)
        lref = 0.5 * np.sum((ref["x_free"] - x0) ** 2)
        worst["x"] = max(worst["x"], relerr(out["x"][b], ref["x_free"]))
        worst["loss"] = max(worst["loss"], abs(out["loss"][b] - lref) / lref)
        worst["q"] = max(worst["q"], relerr(out["q_bar"][b], ref["q_bar"]))
        worst["p"] = max(worst["p"], relerr(out["loads_bar"][b], ref["loads_bar"]))
        worst["xs"] = max(worst["xs"], relerr(out["xs_bar"][b], ref["xyz_fixed_bar"]))
        st = ref["state"]
        worst["state"] = max(worst["state"], relerr(out["lengths"][b], st["lengths"].reshape(-1)),
                             relerr(out["forces"][b], st["forces"].reshape(-1)))
        # residuals vanish at the free nodes (equilibrium): compare in absolute terms against the reaction scale
        scale = np.abs(st["residuals"]).max()
        worst["state"] = max(worst["state"], float(np.abs(out["residuals"][b] - st["residuals"]).max() / scale))
    fa.close()
    assert worst["x"] <= POS_RTOL and worst["loss"] <= POS_RTOL and worst["state"] <= POS_RTOL, worst
    assert worst["q"] <= GRAD_RTOL and worst["p"] <= GRAD_RTOL and worst["xs"] <= GRAD_RTOL, worst
