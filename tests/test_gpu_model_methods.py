"""
GPU: the model's method-by-method surface (models.py:99-193, 349-408, 858-1014) and the pieces around the solve that
round 1 left out -- against the reference's arrays (tests/golden: ref_state_*) and the oracle:
  * edges_vectors / edges_lengths / edges_forces / nodes_residuals with `structure.connectivity`, and their gradients
    against the fused state kernel's,
  * faces_load / edges_load, load_xyz_matrix / load_xyz_matrix_from_r_fixed / residual_free_matrix,
  * the default model (tmax = 100, models.py:77) on nodal loads = the single linear step,
  * UNROLLED differentiation (implicit_diff=False) against the implicit adjoint, 1e-8 -- the reference's own
    gradient test (tests/test_taylor_convergence.py:247-257),
  * splu_cpu / splu_solve_cpu (sparse.py:386-503): repeated solves with one device-resident factor,
  * the cotangent of K on the pattern (fdm_kbar_pattern) and its q-transpose (fdm_stiffness_vjp) against dense algebra.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from oracle import model as om  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def T(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda:0")
    return t.requires_grad_(True) if grad else t


def N(t):
    return t.detach().cpu().numpy()


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("name", ["cablenet12", "olympia"])
def test_state_methods_vs_reference_arrays(name):
    import jax_fdm_b200.equilibrium as eq
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    s = eq.EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=0)
    model = eq.EquilibriumModelSparse(tmax=1)
    q, xyz, loads = T(g["q"], True), T(g["ref_state_xyz"], True), T(g["loads"], True)
    C_ = s.connectivity
    vectors = model.edges_vectors(xyz, C_)
    lengths = model.edges_lengths(vectors)
    forces = model.edges_forces(q, lengths)
    residuals = model.nodes_residuals(q, loads, vectors, C_)
    tol = 1e-12 * max(1.0, np.abs(g["ref_state_xyz"]).max())
    for got, key in ((vectors, "vectors"), (lengths, "lengths"), (forces, "forces"), (residuals, "residuals")):
        assert got.shape == g[f"ref_state_{key}"].shape, key
        assert np.abs(N(got) - g[f"ref_state_{key}"]).max() <= tol, key
    # gradients of the composed route = gradients of the fused state kernel
    rng = np.random.default_rng(0)
    w = {k: T(rng.standard_normal(g[f"ref_state_{k}"].shape)) for k in ("vectors", "lengths", "forces", "residuals")}
    loss = (w["vectors"] * vectors).sum() + (w["lengths"] * lengths).sum() + (w["forces"] * forces).sum() + (w["residuals"] * residuals).sum()
    loss.backward()
    q2, xyz2, loads2 = T(g["q"], True), T(g["ref_state_xyz"], True), T(g["loads"], True)
    st = model.equilibrium_state(q2, xyz2, loads2, s)
    sum((w[k] * getattr(st, k)).sum() for k in w).backward()
    assert relerr(N(q.grad), N(q2.grad)) <= 1e-12 and relerr(N(xyz.grad), N(xyz2.grad)) <= 1e-12
    assert relerr(N(loads.grad), N(loads2.grad)) <= 1e-12
    # the structure itself is accepted where the reference takes the matrix
    assert torch.equal(model.edges_vectors(xyz, s), vectors)


@pytest.mark.parametrize("local", [False, True])
def test_load_methods_and_load_xyz_matrices(local):
    import jax_fdm_b200.equilibrium as eq
    g = dict(np.load(os.path.join(GOLDEN, "mesh_faces.npz")))
    tag = "mixed"
    s = eq.EquilibriumMeshStructureSparse(g[f"{tag}_nodes"], g[f"{tag}_faces"], g[f"{tag}_edges"], g[f"{tag}_supports"], device=0)
    so = oracle.build_structure(g[f"{tag}_nodes"], g[f"{tag}_edges"], g[f"{tag}_supports"])
    xyz = g[f"{tag}_{'local' if local else 'global'}_state_xyz"]
    rng = np.random.default_rng(2)
    eload = 0.05 * rng.standard_normal((s.num_edges, 3))
    fload = g[f"{tag}_faces_load"]
    model = eq.EquilibriumModelSparse(tmax=1, is_load_local=local)
    zero = np.zeros_like(xyz)
    want_e = om.nodes_load_edges(xyz, zero, eload, so, local)
    want_f = om.nodes_load_faces(xyz, zero, fload, so, s.faces_indexed, s.edges_faces_indexed, local, pad="first")
    assert relerr(N(model.edges_load(T(xyz), T(eload), s, local)), want_e) <= 1e-12
    assert relerr(N(model.faces_load(T(xyz), T(fload), s, local)), want_f) <= 1e-12
    assert relerr(N(eq.nodes_load_from_edges(T(xyz), T(eload), s, local)), want_e) <= 1e-12
    # load_xyz_matrix: P(x) = (loads + tributary loads at xyz(x))[free] - R_fixed
    q, xs, loads = g[f"{tag}_q"], g[f"{tag}_xyz_fixed"], g[f"{tag}_loads"]
    params = eq.EquilibriumParametersState(T(q), T(xs), eq.LoadState(T(loads), T(eload), T(fload)))
    x_free = xyz[so.indices_free]
    P = model.load_xyz_matrix(params, T(x_free), s)
    nodes_load = loads + want_e + want_f
    want_P = om.load_matrix(q, xs, nodes_load, so)
    assert relerr(N(P), want_P) <= 1e-12
    R_fixed = model.residual_fixed_matrix(T(q), T(xs), s)
    assert relerr(N(R_fixed), om.residual_fixed_matrix(q, xs, so)) <= 1e-12
    P2 = model.load_xyz_matrix_from_r_fixed((None, R_fixed, T(xs), params.loads), T(x_free), s)
    assert relerr(N(P2), want_P) <= 1e-12
    K = om.stiffness_matrix_sparse(q, so)
    assert relerr(N(model.residual_free_matrix(params, T(x_free), s)), K @ x_free - want_P) <= 1e-10


def test_default_model_on_nodal_loads_is_the_linear_step():
    import jax_fdm_b200.equilibrium as eq
    g = dict(np.load(os.path.join(GOLDEN, "cablenet12.npz")))
    s = eq.EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=0)
    params = eq.EquilibriumParametersState(T(g["q"]), T(g["xyz_fixed"]), eq.LoadState(T(g["loads"]), 0.0, 0.0))
    st = eq.EquilibriumModelSparse()(params, s)                    # tmax = 100 (models.py:77)
    assert relerr(N(st.xyz), g["ref_state_xyz"]) <= 1e-10 and relerr(N(st.residuals), g["ref_state_residuals"]) <= 1e-8


@pytest.mark.parametrize("local", [False, True])
def test_unrolled_differentiation_matches_the_implicit_adjoint(local):
    """tests/test_taylor_convergence.py:247-257: implicit vs unrolled gradients agree (1e-8 here as there)."""
    import jax_fdm_b200.equilibrium as eq
    prob = datasets.cablenet_problem(6, seed=5)
    rng = np.random.default_rng(5)
    E = prob.edges.shape[0]
    eload = np.zeros((E, 3))
    eload[:, 2] = -0.05 * rng.uniform(0.5, 1.5, size=E)
    eload[:, 0] = 0.01 * rng.standard_normal(E)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    grads = {}
    for implicit in (True, False):
        q, xs, p, w = T(prob.q, True), T(prob.xyz_fixed, True), T(prob.loads, True), T(eload, True)
        model = eq.EquilibriumModelSparse(tmax=200, eta=1e-13, is_load_local=local, implicit_diff=implicit, tol=1e-14)
        st = model(eq.EquilibriumParametersState(q, xs, eq.LoadState(p, w, 0.0)), s)
        loss = 0.5 * ((st.xyz - T(prob.xyz0)) ** 2).sum() + 0.1 * (st.lengths ** 2).sum()
        loss.backward()
        assert model.last_solver_config["steps"] > 1
        grads[implicit] = [N(t.grad) for t in (q, xs, p, w)] + [float(loss.detach())]
    assert abs(grads[True][-1] - grads[False][-1]) <= 1e-12 * abs(grads[True][-1])
    for a, b, nm in zip(grads[True][:4], grads[False][:4], ("q", "xyz_fixed", "loads", "edges_load")):
        assert relerr(b, a) <= 1e-8, nm


@pytest.mark.parametrize("name", ["arch10", "cablenet12", "olympia", "random40", "readme10", "stuttgart21"])
def test_dense_stiffness_matrix_is_the_reference_dense_model_array(name):
    """`EquilibriumModel.stiffness_matrix` of the reference's dense model returns the (Nf, Nf) array
    (models.py:800-822); `ref_dense_K` is that array, computed by the reference's own models.py."""
    import jax_fdm_b200.equilibrium as eq
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    s = eq.EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=0)
    model = eq.EquilibriumModel(tmax=1)
    q = T(g["q"]).requires_grad_(True)
    Kd = model.stiffness_matrix_dense(q, s)
    assert Kd.shape == g["ref_dense_K"].shape
    assert np.allclose(N(Kd), g["ref_dense_K"], rtol=1e-13, atol=1e-13)
    assert torch.equal(Kd, model.stiffness_matrix(q, s).todense())
    # its pull-back is the transpose of the assembly: d/dq sum(W * K) against the oracle's basis matrices
    so = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    W = np.random.default_rng(0).standard_normal(g["ref_dense_K"].shape)
    (Kd * T(W)).sum().backward()
    E = g["q"].size
    want = np.array([(om.stiffness_matrix_dense(np.eye(E)[e], so) * W).sum() for e in range(E)])
    assert relerr(N(q.grad), want) <= 1e-12
    # a batch scatters design by design
    qb = T(np.stack([g["q"], 2.0 * g["q"]]))
    Kb = model.stiffness_matrix(qb, s).todense()
    assert Kb.shape == (2,) + g["ref_dense_K"].shape and np.allclose(N(Kb)[1], 2.0 * g["ref_dense_K"], rtol=1e-13, atol=1e-13)


def test_splu_session_reuses_one_device_factor():
    import jax_fdm_b200.equilibrium as eq
    g = dict(np.load(os.path.join(GOLDEN, "pillow30.npz")))
    s = eq.EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=0)
    so = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    model = eq.EquilibriumModelSparse(tmax=1)
    K = model.stiffness_matrix(T(g["q"]), s)
    Ko = om.stiffness_matrix_sparse(g["q"], so)
    eq.splu_clear()
    sid = eq.splu_cpu(K, session_id=7)
    rng = np.random.default_rng(3)
    for _ in range(3):
        b = rng.standard_normal((s.num_free, 3))
        x = eq.splu_solve_cpu(sid, T(b))
        assert relerr(N(x), om.sparse_solve(Ko, b)) <= 1e-10
    with pytest.raises(ValueError):
        eq.splu_solve_cpu(8, T(b))
    eq.splu_clear()
    with pytest.raises(ValueError):
        eq.splu_solve_cpu(0, T(b))
    # the GPU aliases of the reference give the same solution
    assert torch.equal(eq.spsolve_gpu_ravel(K, T(b)), eq.spsolve_gpu_stack(K, T(b)))


def test_kbar_on_pattern_and_its_q_transpose_vs_dense_algebra():
    import jax_fdm_b200.equilibrium as eq
    g = dict(np.load(os.path.join(GOLDEN, "random40.npz")))
    s = eq.EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=0)
    so = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    rng = np.random.default_rng(4)
    B = 3
    a, b = rng.standard_normal((B, s.num_free, 3)), rng.standard_normal((B, s.num_free, 3))
    Kbar = N(eq.ops.kbar_pattern(s, T(a), T(b), -1.0))
    ia = s.index_array
    cols = np.repeat(np.arange(s.num_free), np.diff(ia.indptr))
    for k in range(B):
        assert np.allclose(Kbar[k], -np.sum(a[k][ia.indices] * b[k][cols], axis=1), rtol=1e-14, atol=1e-14)
    # q_bar = d <Kbar, K(q)> / dq, with K(q) linear in q: finite-difference free check through the oracle's assembly
    qbar = N(eq.ops.stiffness_vjp(s, T(Kbar)))
    E = s.num_edges
    basis = np.stack([om.stiffness_matrix_sparse(np.eye(E)[e], so).data for e in range(E)])      # (E, nnz)
    assert np.allclose(qbar, Kbar @ basis.T, rtol=1e-13, atol=1e-13)


def test_fdm_front_door_solves_mixed_sign_q_and_fails_loudly_when_a_design_stays_unsolved():
    """Mixed-sign q on a band too wide for the signed warp factorisation: the solver in front flags the design and the
    band L D L^T solves it, as the reference's LU does (checked against SuperLU).  When a design does come back
    unsolved (here: the same system with the fallback switched off) the kernels hand back NaN and `fdm()` turns that
    into an exception instead of returning NaN positions."""
    import jax_fdm_b200.equilibrium as eq
    import oracle
    prob = datasets.cablenet_problem(40, seed=1)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    assert s.half_bandwidth > 30
    st, _ = eq.fdm(prob.nodes, prob.edges, prob.supports, prob.q, prob.xyz_fixed, prob.loads, structure=s)
    assert bool(torch.isfinite(st.xyz).all())
    q = prob.q.copy()
    q[::3] *= -1.0
    st, _ = eq.fdm(prob.nodes, prob.edges, prob.supports, q, prob.xyz_fixed, prob.loads, structure=s)
    xo = oracle.sparse_solve(oracle.stiffness_matrix_sparse(q, so), oracle.load_matrix(q, prob.xyz_fixed, prob.loads, so))
    x = st.xyz.cpu().numpy()[np.asarray(so.indices_free)]
    assert np.abs(x - xo).max() <= 1e-8 * np.abs(xo).max()
    # with the fallback switched off (a process-wide switch, hence the child process) the kernels flag the design and
    # hand back NaN, and the front door raises instead of returning NaN positions
    import os
    import subprocess
    import sys
    code = """
import numpy as np, torch
from jax_fdm_b200 import datasets
import jax_fdm_b200.equilibrium as eq
prob = datasets.cablenet_problem(40, seed=1)
q = prob.q.copy(); q[::3] *= -1.0
try:
    eq.fdm(prob.nodes, prob.edges, prob.supports, q, prob.xyz_fixed, prob.loads)
except FloatingPointError as e:
    print("raised:", e)
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FDM_LDLT_FALLBACK="0", PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=300)
    assert "raised:" in r.stdout, r.stdout + r.stderr
