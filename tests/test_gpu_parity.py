"""
GPU parity tests: the CUDA path, called through the C ABI (jax_fdm_b200.equilibrium.ops ->
ctypes -> libfdm_b200.so), against the CPU oracle on the same seeded inputs and against the
golden fixtures produced by the reference's own code.

Tolerances (BASELINE.json north_star): bit-exact for index maps and for K.data (a gather and an
ordered sum); <= 1e-10 relative on positions / residual; <= 1e-8 on gradients.
"""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["arch10", "readme10", "cablenet12", "pillow30", "olympia", "stuttgart21", "liew_dome", "random40"]
POS_RTOL = 1e-10
GRAD_RTOL = 1e-8


def dev():
    return torch.device("cuda", 0)


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev())


def N(t):
    return t.detach().cpu().numpy()


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def load(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


@pytest.fixture(scope="module")
def eq():
    import jax_fdm_b200.equilibrium as e
    return e


def make(eq, g):
    s = eq.EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=0)
    so = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    return s, so


def solvers_for(s):
    out = ["pcg"]
    if s.num_free * (s.half_bandwidth + 1 + 3) * 8 <= 227 * 1024 - 1024:
        out.append("cholesky")
    return out


@pytest.mark.parametrize("name", CASES)
def test_assemble_bit_exact_and_rhs(eq, name):
    g = load(name)
    s, so = make(eq, g)
    K, rhs = eq.ops.assemble(s, T(g["q"]), T(g["xyz_fixed"]), T(g["loads"]))
    assert np.array_equal(N(K), g["ref_K_data"])                      # vs the reference's own output
    assert np.array_equal(N(K), oracle.stiffness_matrix_sparse(g["q"], so).data)
    np.testing.assert_allclose(N(rhs), g["ref_P"], rtol=0, atol=1e-13 * max(1, np.abs(g["ref_P"]).max()))


@pytest.mark.parametrize("name", CASES)
def test_forward_solve_and_state(eq, name):
    g = load(name)
    s, so = make(eq, g)
    q, xs, p = T(g["q"]), T(g["xyz_fixed"]), T(g["loads"])
    K, rhs = eq.ops.assemble(s, q, xs, p)
    for solver in solvers_for(s):
        x, info = eq.ops.solve(s, K, rhs, solver=solver, tol=1e-13, return_info=True)
        info = N(info)[0]
        assert info[0] == 0, (solver, info)
        assert relerr(N(x), g["ref_x_free"]) <= POS_RTOL, solver
        # residual of the linear system, relative
        r = oracle.stiffness_matrix_sparse(g["q"], so) @ N(x) - g["ref_P"]
        assert np.abs(r).max() <= POS_RTOL * max(np.abs(g["ref_P"]).max(), np.abs(g["ref_K_data"]).max() * np.abs(N(x)).max())
    xyz = eq.ops.positions(s, x, xs)
    assert relerr(N(xyz), g["ref_state_xyz"]) <= POS_RTOL
    v, l, f, r = eq.ops.state(s, q, xyz, p)
    scale = max(1.0, np.abs(g["ref_state_xyz"]).max())
    for mine, ref in [(v, "vectors"), (l, "lengths"), (f, "forces"), (r, "residuals")]:
        refa = g[f"ref_state_{ref}"]
        assert N(mine).shape == refa.shape
        np.testing.assert_allclose(N(mine), refa, rtol=1e-9, atol=1e-9 * scale * max(1, np.abs(g["q"]).max()), err_msg=ref)


@pytest.mark.parametrize("name", ["readme10", "liew_dome", "random40"])
def test_spmm(eq, name):
    g = load(name)
    s, so = make(eq, g)
    rng = np.random.default_rng(0)
    X = rng.normal(size=(s.num_free, 3))
    Y = eq.ops.spmm(s, T(g["ref_K_data"]), T(X))
    ref = oracle.stiffness_matrix_sparse(g["q"], so) @ X
    assert relerr(N(Y), ref) <= 1e-14


@pytest.mark.parametrize("name", ["cablenet12", "random40", "arch10", "pillow30"])
def test_state_vjp_and_adjoint_grads(eq, name):
    g = load(name)
    s, so = make(eq, g)
    rng = np.random.default_rng(5)
    st0, xf = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], so)
    w = {f: rng.normal(size=st0[f].shape) for f in st0}
    ref = oracle.forward_adjoint(g["q"], g["xyz_fixed"], g["loads"], so, lambda st, x: w)
    # K4 reverse
    qb_o, xb_o, pb_o = oracle.equilibrium_state_vjp(g["q"], st0["xyz"], so, w)
    qb, xb, pb = eq.ops.state_vjp(s, T(g["q"]), T(st0["xyz"]), xyz_cot=T(w["xyz"]),
                                  residuals_cot=T(w["residuals"]), lengths_cot=T(w["lengths"]),
                                  forces_cot=T(w["forces"]), loads_cot=T(w["loads"]),
                                  vectors_cot=T(w["vectors"]))
    assert relerr(N(qb), qb_o) <= 1e-12 and relerr(N(xb), xb_o) <= 1e-12 and relerr(N(pb), pb_o) <= 1e-13
    # positions transpose + adjoint solve + K3, accumulated on top of the direct terms
    gfree, xsb = eq.ops.positions_vjp(s, xb)
    K, _ = eq.ops.assemble(s, T(g["q"]), want_rhs=False)
    lam = eq.ops.solve(s, K, gfree, solver="pcg", tol=1e-14)
    assert relerr(N(lam), ref["lam"]) <= POS_RTOL
    qb2, pb2, xsb2 = eq.ops.adjoint_grads(s, T(g["q"]), T(st0["xyz"]), lam, q_bar=qb, loads_bar=pb,
                                          xyz_fixed_bar=xsb)
    assert relerr(N(qb2), ref["q_bar"]) <= GRAD_RTOL
    assert relerr(N(pb2), ref["loads_bar"]) <= GRAD_RTOL
    assert relerr(N(xsb2), ref["xyz_fixed_bar"]) <= GRAD_RTOL


@pytest.mark.parametrize("name,fused", [("cablenet12", True), ("cablenet12", False), ("random40", True),
                                        ("random40", False), ("liew_dome", True), ("pillow30", True)])
def test_model_autograd_matches_oracle(eq, name, fused):
    g = load(name)
    s, so = make(eq, g)
    rng = np.random.default_rng(11)
    st0, _ = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], so)
    w = {f: rng.normal(size=st0[f].shape) for f in st0}
    ref = oracle.forward_adjoint(g["q"], g["xyz_fixed"], g["loads"], so, lambda st, x: w)

    q = T(g["q"]).requires_grad_(True)
    xs = T(g["xyz_fixed"]).requires_grad_(True)
    p = T(g["loads"]).requires_grad_(True)
    model = eq.EquilibriumModelSparse(tmax=1, solver="pcg", tol=1e-14)
    if fused:
        state = model(eq.EquilibriumParametersState(q, xs, eq.LoadState(p)), s)
    else:  # method-by-method route, as user code composing the reference's methods would
        xf = model.nodes_free_positions(q, xs, p, s, fused=False)
        xyz = model.nodes_positions(xf, xs, s)
        state = model.equilibrium_state(q, xyz, p, s)
    for f in state._fields:
        np.testing.assert_allclose(N(getattr(state, f)), ref["state"][f], rtol=1e-9,
                                   atol=1e-9 * max(1, np.abs(ref["state"]["xyz"]).max()), err_msg=f)
    loss = sum((T(w[f]) * getattr(state, f)).sum() for f in state._fields)
    loss.backward()
    assert relerr(N(q.grad), ref["q_bar"]) <= GRAD_RTOL
    assert relerr(N(xs.grad), ref["xyz_fixed_bar"]) <= GRAD_RTOL
    assert relerr(N(p.grad), ref["loads_bar"]) <= GRAD_RTOL


def test_batched_pillow_designs(eq):
    """Config 5 shape at small batch: B designs, one topology, compression (K = -SPD)."""
    prob = datasets.pillow_problem(30)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    B = 6
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    q, xs, p = T(qb), T(prob.xyz_fixed), T(prob.loads)
    K, rhs = eq.ops.assemble(s, q, xs, p)
    ws = eq.ops.Workspace()
    x, info = eq.ops.solve(s, K, rhs, solver="cholesky", workspace=ws, return_info=True)
    assert np.all(N(info)[:, 0] == 0) and np.all(N(info)[:, 3] == -1.0)
    x0 = T(prob.xyz0[so.indices_free])
    loss, gcot = eq.ops.loss_sqdist(x, x0)
    lam = eq.ops.solve(s, K, gcot, solver="cholesky", workspace=ws, reuse_factor=True)
    xyz = eq.ops.positions(s, x, xs)
    qbar, pbar, xsbar = eq.ops.adjoint_grads(s, q, xyz, lam)
    for b in range(B):
        w = {"xyz": np.zeros((so.num_nodes, 3))}
        def cot(st, xf, b=b):
            c = np.zeros((so.num_nodes, 3))
            c[so.indices_free] = xf - prob.xyz0[so.indices_free]
            return {"xyz": c}
        ref = oracle.forward_adjoint(qb[b], prob.xyz_fixed, prob.loads, so, cot)
        assert relerr(N(x[b]), ref["x_free"]) <= POS_RTOL
        assert abs(float(loss[b]) - 0.5 * np.sum((ref["x_free"] - prob.xyz0[so.indices_free]) ** 2)) <= 1e-10 * float(loss[b])
        assert relerr(N(qbar[b]), ref["q_bar"]) <= GRAD_RTOL
        assert relerr(N(pbar[b]), ref["loads_bar"]) <= GRAD_RTOL
        assert relerr(N(xsbar[b]), ref["xyz_fixed_bar"]) <= GRAD_RTOL


def test_cablenet200_full_size_properties(eq):
    """Config 4 at full size: residual, symmetry of the adjoint identity, CPU oracle agreement."""
    prob = datasets.cablenet_problem(200)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    q, xs, p = T(prob.q), T(prob.xyz_fixed), T(prob.loads)
    K, rhs = eq.ops.assemble(s, q, xs, p)
    Ko = oracle.stiffness_matrix_sparse(prob.q, so)
    assert np.array_equal(N(K), Ko.data)
    x, info = eq.ops.solve(s, K, rhs, solver="auto", tol=1e-12, return_info=True)
    assert N(info)[0, 0] == 0
    r = Ko @ N(x) - N(rhs)
    assert np.linalg.norm(r) <= POS_RTOL * np.linalg.norm(N(rhs))
    xo = oracle.sparse_solve(Ko, N(rhs))
    assert relerr(N(x), xo) <= POS_RTOL
    # adjoint identity  <g, K^-1 b> == <K^-1 g, b>
    gcot = T(np.random.default_rng(1).normal(size=(s.num_free, 3)))
    lam = eq.ops.solve(s, K, gcot, solver="auto", tol=1e-12)
    lhs = float((gcot * x).sum())
    rhs_ = float((lam * rhs).sum())
    assert abs(lhs - rhs_) <= 1e-9 * abs(lhs)
    # gradient vs oracle
    xyz = eq.ops.positions(s, x, xs)
    qbar, pbar, xsbar = eq.ops.adjoint_grads(s, q, xyz, lam)
    lam_o = oracle.sparse_solve(Ko, N(gcot))
    full = np.zeros((so.num_nodes, 3)); full[so.indices_free] = lam_o
    xyz_o = oracle.nodes_positions(xo, prob.xyz_fixed, so)
    t, h = so.edges_indexed[:, 0], so.edges_indexed[:, 1]
    qbar_o = -np.sum((full[h] - full[t]) * (xyz_o[h] - xyz_o[t]), axis=1)
    assert relerr(N(qbar), qbar_o) <= GRAD_RTOL


def test_cablenet200_pipeline_forward_and_adjoint_converge(eq):
    """Config 4 through the benchmark's device pipeline: both solves of the persistent PCG (the adjoint
    reuses the coarse inverse) must CONVERGE to 1e-12 in a few hundred iterations, and q_bar must match
    the CPU oracle."""
    from jax_fdm_b200.pipeline import ForwardAdjoint

    prob = datasets.cablenet_problem(200)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    pipe = ForwardAdjoint(s, batch=1, solver="auto", tol=1e-12, device=0)
    x0 = prob.xyz0[s.indices_free]
    pipe.set_inputs(prob.q[None], prob.xyz_fixed, prob.loads, x0)
    pipe.step()
    pipe.stream.synchronize()
    for info in (N(pipe.info_f)[0], N(pipe.info_a)[0]):
        assert info[0] == 0, f"solve status {info[0]}"
        assert 0 < info[1] < 2000 and info[2] <= 1e-12

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    ref = oracle.forward_adjoint(prob.q, prob.xyz_fixed, prob.loads, so, cot)
    assert relerr(N(pipe.x)[0], ref["x_free"]) <= POS_RTOL
    assert relerr(N(pipe.q_bar)[0], ref["q_bar"]) <= GRAD_RTOL


def test_zero_rhs_column_and_errors(eq):
    g = load("arch10")      # y column of rhs is identically zero
    s, so = make(eq, g)
    K, rhs = eq.ops.assemble(s, T(g["q"]), T(g["xyz_fixed"]), T(g["loads"]))
    assert float(rhs[:, 1].abs().max()) == 0.0
    x = eq.ops.solve(s, K, rhs, solver="pcg")
    assert torch.isfinite(x).all() and float(x[:, 1].abs().max()) == 0.0
    with pytest.raises(TypeError):
        eq.ops.solve(s, K.float(), rhs)
    with pytest.raises(ValueError):
        eq.ops.solve(s, K[:-1], rhs)


@pytest.mark.parametrize("name", ["pillow30", "cablenet12", "liew_dome"])
def test_batched_assembly_and_spmm_are_bit_identical_to_the_per_design_kernels(eq, name):
    """A batch on one topology takes kernels that load the index streams once per CTA and stream the designs through
    (assemble_K_batch_kernel + assemble_rhs_tma_kernel / _stream_kernel / _pairs_kernel, spmm_batch_kernel); a single design takes the per-design kernels.  Same additions in the
    same order: the outputs must be the same bits, and K.data the reference's (ref_K_data) for the fixture's own q."""
    g = load(name)
    s, so = make(eq, g)
    rng = np.random.default_rng(3)
    B = 37
    qb = g["q"][None, :] * (1.0 + 0.2 * rng.random((B, g["q"].size)))
    qb[0] = g["q"]
    xsb = g["xyz_fixed"][None] + 0.1 * rng.standard_normal((B,) + g["xyz_fixed"].shape)
    pb = g["loads"][None] + 0.1 * rng.standard_normal((B,) + g["loads"].shape)
    K, rhs = eq.ops.assemble(s, T(qb), T(xsb), T(pb))
    assert np.array_equal(N(K)[0], oracle.stiffness_matrix_sparse(g["q"], so).data)
    X = T(rng.standard_normal((B, s.num_free, 3)))
    Y = eq.ops.spmm(s, K, X)
    for b in (0, 1, 17, B - 1):
        K1, rhs1 = eq.ops.assemble(s, T(qb[b]), T(xsb[b]), T(pb[b]))
        assert torch.equal(K1, K[b]) and torch.equal(rhs1, rhs[b])
        assert torch.equal(eq.ops.spmm(s, K1, X[b].contiguous()), Y[b])
    # shared supports / loads (stride 0) through the batched kernel
    K2, rhs2 = eq.ops.assemble(s, T(qb), T(g["xyz_fixed"]), T(g["loads"]))
    for b in (0, 5, B - 1):
        K3, rhs3 = eq.ops.assemble(s, T(qb[b]), T(g["xyz_fixed"]), T(g["loads"]))
        assert torch.equal(K2[b], K3) and torch.equal(rhs2[b], rhs3)
    # K.data alone and the rhs alone (NULL for the other output), straight through the C ABI
    import ctypes as C
    from jax_fdm_b200 import _lib as L
    P = lambda t: C.c_void_p(t.data_ptr())
    q_t, xs_t, p_t = T(qb), T(xsb), T(pb)
    Ko, ro = torch.full_like(K, float("nan")), torch.full_like(rhs, float("nan"))
    E = g["q"].size
    L.check(L.lib().fdm_assemble(s.plan, P(q_t), P(xs_t), P(p_t), P(Ko), None, B, E, xs_t[0].numel(), p_t[0].numel(), None))
    L.check(L.lib().fdm_assemble(s.plan, P(q_t), P(xs_t), P(p_t), None, P(ro), B, E, xs_t[0].numel(), p_t[0].numel(), None))
    torch.cuda.synchronize()
    assert torch.equal(Ko, K) and torch.equal(ro, rhs)
    # the batched squared-distance loss (persistent CTAs, next design prefetched) against the per-design kernel
    x0 = T(rng.standard_normal((s.num_free, 3)))
    n = s.num_free * 3
    gb, lb = torch.empty_like(X), torch.empty(B, dtype=X.dtype, device=X.device)
    L.check(L.lib().fdm_loss_sqdist(P(X), P(x0), P(gb), P(lb), n, B, 0, None))
    for b in (0, 9, B - 1):
        g1, l1 = torch.empty_like(X[b]), torch.empty(1, dtype=X.dtype, device=X.device)
        L.check(L.lib().fdm_loss_sqdist(P(X[b]), P(x0), P(g1), P(l1), n, 1, 0, None))
        assert torch.equal(g1, gb[b]) and torch.equal(l1[0], lb[b])
    assert np.allclose(N(lb), 0.5 * ((N(X) - N(x0)[None]) ** 2).sum(axis=(1, 2)), rtol=1e-13)
    # the state of a batch against the state of its designs one by one, with per-design loads, shared loads (stride 0)
    # and partial outputs (NULL arrays), straight through the C ABI; one design against the oracle
    Nn = s.num_nodes
    xyzb = T(rng.standard_normal((B, Nn, 3)))
    lib = L.lib()

    def state(q_, xyz_, loads_, nb, ls, want=(1, 1, 1, 1)):
        out = [torch.full((nb, E, 3), float("nan"), dtype=X.dtype, device=X.device),
               torch.full((nb, E), float("nan"), dtype=X.dtype, device=X.device),
               torch.full((nb, E), float("nan"), dtype=X.dtype, device=X.device),
               torch.full((nb, Nn, 3), float("nan"), dtype=X.dtype, device=X.device)]
        ptrs = [P(o) if w else None for o, w in zip(out, want)]
        L.check(lib.fdm_state(s.plan, P(q_), P(xyz_), P(loads_), *ptrs, nb, E, ls, None))
        torch.cuda.synchronize()
        return out

    full = state(q_t, xyzb, p_t, B, Nn * 3)
    shared = state(q_t, xyzb, p_t[0].contiguous(), B, 0)
    for b in (0, 1, 18, B - 1):
        one = state(q_t[b].contiguous(), xyzb[b].contiguous(), p_t[b].contiguous(), 1, 0)
        assert all(torch.equal(o[0], f[b]) for o, f in zip(one, full))
        one = state(q_t[b].contiguous(), xyzb[b].contiguous(), p_t[0].contiguous(), 1, 0)
        assert all(torch.equal(o[0], f[b]) for o, f in zip(one, shared))
    part = state(q_t, xyzb, p_t, B, Nn * 3, want=(0, 1, 0, 1))
    assert torch.equal(part[1], full[1]) and torch.equal(part[3], full[3])
    assert torch.isnan(part[0]).all() and torch.isnan(part[2]).all()
    ref = oracle.equilibrium_state(qb[3], N(xyzb)[3], pb[3], so)
    assert np.allclose(N(full[0])[3], ref["vectors"], rtol=1e-12, atol=1e-12)
    assert np.allclose(N(full[1])[3], ref["lengths"][:, 0], rtol=1e-12, atol=1e-12)
    assert np.allclose(N(full[2])[3], ref["forces"][:, 0], rtol=1e-12, atol=1e-12)
    assert np.allclose(N(full[3])[3], ref["residuals"], rtol=1e-12, atol=1e-11)


@pytest.mark.parametrize("B,G", [(8, 4096), (9, 4096), (37, 4096), (37, 4097)])
def test_batched_kernels_write_only_their_outputs(eq, B, G):
    """compute-sanitizer is closed on this pool, so the out-of-bounds check of the batched streaming kernels (TMA bulk
    stores of two-design images, whole-design CTAs, persistent loss CTAs) is done with guard bands: every output sits
    inside a larger buffer of sentinels, and the bands before and after it must come back untouched -- even and odd
    batch sizes (the bulk stores work on pairs of designs)."""
    import ctypes as C
    from jax_fdm_b200 import _lib as L
    g = load("pillow30")
    s, _ = make(eq, g)
    rng = np.random.default_rng(B)
    E, Nf, nnz = g["q"].size, s.num_free, s.nnz
    q = T(g["q"][None, :] * (1.0 + 0.2 * rng.random((B, E))))
    xs, p = T(g["xyz_fixed"]), T(g["loads"])
    # G guard doubles on each side; an odd G leaves the outputs 8-byte aligned only (no bulk stores, no 16-byte copies)
    sentinel = 1234.5

    def guarded(n):
        buf = torch.full((n + 2 * G,), sentinel, dtype=torch.float64, device=q.device)
        return buf, buf[G:G + n]

    def intact(buf, n):
        return bool((buf[:G] == sentinel).all()) and bool((buf[G + n:] == sentinel).all())

    P = lambda t: C.c_void_p(t.data_ptr())
    lib = L.lib()
    Kb, K = guarded(B * nnz)
    rb, rhs = guarded(B * Nf * 3)
    L.check(lib.fdm_assemble(s.plan, P(q), P(xs), P(p), P(K), P(rhs), B, E, 0, 0, None))
    X = T(rng.standard_normal((B, Nf, 3)))
    Yb, Y = guarded(B * Nf * 3)
    L.check(lib.fdm_spmm(s.plan, P(K), P(X), P(Y), B, None))
    x0 = T(rng.standard_normal((Nf, 3)))
    gb, gg = guarded(B * Nf * 3)
    lb, ll = guarded(B)
    L.check(lib.fdm_loss_sqdist(P(X), P(x0), P(gg), P(ll), Nf * 3, B, 0, None))
    torch.cuda.synchronize()
    assert intact(Kb, B * nnz) and intact(rb, B * Nf * 3) and intact(Yb, B * Nf * 3) and intact(gb, B * Nf * 3) and intact(lb, B)
    # and everything inside was written
    for t in (K, rhs, Y, gg, ll):
        assert not bool((t == sentinel).any())
    K2, rhs2 = eq.ops.assemble(s, q, xs, p)
    assert torch.equal(K.view(B, nnz), K2) and torch.equal(rhs.view(B, Nf, 3), rhs2)
