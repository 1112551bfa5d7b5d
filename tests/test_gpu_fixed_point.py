"""
tmax > 1 on the GPU: fixed-point equilibrium with shape-dependent (tributary edge) loads and its
implicit adjoint, against the CPU oracle (models.py:410-478, 512-618; fixed_point.py:139-240,
501-604; loads.py:296-331, 407-465).

Forward parity: <= 1e-10 relative on x* (both iterate to a tight eta).  Gradient parity: the
reference pins its implicit adjoint against unrolled AD at 1e-7 (tests/test_taylor_convergence.py:
265-304); JAX cannot run here, so the implicit gradients are checked against central finite
differences of the ORACLE's forward solve (fourth-order stencil) at the north star's 1e-8 relative.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402


GRAD_RTOL = 1e-8      # BASELINE.json north star: 1e-8 on gradients


def fd5(f, arr, k, h=1e-3):
    """d f / d arr.flat[k] by the fourth-order five-point stencil: truncation O(h^4), round-off O(eps |f| / h) -- what
    lets a finite-difference check reach the 1e-8 the north star asks of gradients (two-point differences stop at ~1e-6)."""
    step = h * max(1.0, abs(arr.flat[k]))
    d = np.zeros(arr.size)
    d[k] = step
    d = d.reshape(arr.shape)
    return (-f(arr + 2 * d) + 8.0 * f(arr + d) - 8.0 * f(arr - d) + f(arr - 2 * d)) / (12.0 * step)


def T(a, grad=False):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=torch.device("cuda", 0))
    return t.requires_grad_(True) if grad else t


def problem(nx=6, seed=5):
    prob = datasets.cablenet_problem(nx, seed=seed)
    rng = np.random.default_rng(seed)
    E = prob.edges.shape[0]
    edges_load = np.zeros((E, 3))
    edges_load[:, 2] = -0.05 * rng.uniform(0.5, 1.5, size=E)       # self-weight-like line load
    edges_load[:, 0] = 0.01 * rng.standard_normal(E)
    return prob, edges_load


def oracle_loss(q, xs, loads, eload, so, x0):
    x, steps = oracle.fixed_point_forward(q, xs, loads, eload, so, tmax=200, eta=1e-13)
    xyz = oracle.nodes_positions(x, xs, so)
    ei = so.edges_indexed
    lengths = np.linalg.norm(xyz[ei[:, 1]] - xyz[ei[:, 0]], axis=1)
    return 0.5 * np.sum((xyz - x0) ** 2) + 0.1 * np.sum(lengths ** 2), x, steps


@pytest.mark.parametrize("solver", ["cholesky", "pcg"])
def test_fixed_point_forward_and_implicit_gradients(solver):
    import jax_fdm_b200.equilibrium as eq

    prob, eload = problem()
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    x0 = prob.xyz0

    q, xs, p, w = T(prob.q, True), T(prob.xyz_fixed, True), T(prob.loads, True), T(eload, True)
    model = eq.EquilibriumModelSparse(tmax=200, eta=1e-13, solver=solver, tol=1e-14)
    state = model(eq.EquilibriumParametersState(q, xs, eq.LoadState(p, w, 0.0)), s)
    loss = 0.5 * ((state.xyz - T(x0)) ** 2).sum() + 0.1 * (state.lengths ** 2).sum()
    loss.backward()

    ref_loss, x_ref, steps = oracle_loss(prob.q, prob.xyz_fixed, prob.loads, eload, so, x0)
    assert 1 < steps < 200, "the oracle's fixed point must converge in a few iterations"
    assert model.last_solver_config["steps"] > 1
    xf = state.xyz.detach().cpu().numpy()[so.indices_free]
    assert np.abs(xf - x_ref).max() <= 1e-10 * np.abs(x_ref).max()
    assert abs(loss.item() - ref_loss) <= 1e-10 * abs(ref_loss)
    # loads reported by the state include the tributary edge loads at the converged geometry
    xyz_ref = oracle.nodes_positions(x_ref, prob.xyz_fixed, so)
    loads_ref = oracle.nodes_load_edges(xyz_ref, prob.loads, eload, so)
    assert np.abs(state.loads.detach().cpu().numpy() - loads_ref).max() <= 1e-10 * np.abs(loads_ref).max()

    rng = np.random.default_rng(0)

    def fd(arr, make_args, n_probe=6):
        flat_idx = rng.choice(arr.size, size=min(n_probe, arr.size), replace=False)
        return [(k, fd5(lambda a: oracle_loss(*make_args(a), so, x0)[0], arr, k)) for k in flat_idx]

    checks = [
        (q.grad, prob.q, lambda a: (a, prob.xyz_fixed, prob.loads, eload)),
        (xs.grad, prob.xyz_fixed, lambda a: (prob.q, a, prob.loads, eload)),
        (p.grad, prob.loads, lambda a: (prob.q, prob.xyz_fixed, a, eload)),
        (w.grad, eload, lambda a: (prob.q, prob.xyz_fixed, prob.loads, a)),
    ]
    for grad, arr, mk in checks:
        gnp = grad.cpu().numpy()
        scale = np.abs(gnp).max()
        for k, ref in fd(arr, mk):
            assert abs(gnp.flat[k] - ref) <= GRAD_RTOL * scale + 1e-11, (k, gnp.flat[k], ref)


def test_tmax1_ignores_edge_loads_in_the_solve_but_reports_them():
    """models.py:447, 475-478: with tmax = 1 edge loads do not enter the solve
    but are added to the reported loads / residuals."""
    import jax_fdm_b200.equilibrium as eq

    prob, eload = problem()
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    model = eq.EquilibriumModelSparse(tmax=1)
    st = model(eq.EquilibriumParametersState(T(prob.q), T(prob.xyz_fixed), eq.LoadState(T(prob.loads), T(eload), 0.0)), s)
    _, xf = oracle.forward(prob.q, prob.xyz_fixed, prob.loads, so)
    xyz = oracle.nodes_positions(xf, prob.xyz_fixed, so)
    assert np.abs(st.xyz.cpu().numpy() - xyz).max() <= 1e-10 * np.abs(xyz).max()
    loads_ref = oracle.nodes_load_edges(xyz, prob.loads, eload, so)
    assert np.abs(st.loads.cpu().numpy() - loads_ref).max() <= 1e-12 * np.abs(loads_ref).max()


@pytest.mark.parametrize("solver", ["cholesky", "pcg"])
def test_forward_mode_rule_matches_finite_differences(solver):
    """nodes_free_positions_jvp (tangent solve with the primal factor) against central differences of the
    CPU oracle's forward solve, along a random direction in (q, xyz_fixed, loads)."""
    import jax_fdm_b200.equilibrium as eq

    prob, _ = problem(nx=7, seed=11)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    rng = np.random.default_rng(2)
    dq, dxs, dp = (rng.standard_normal(a.shape) for a in (prob.q, prob.xyz_fixed, prob.loads))
    model = eq.EquilibriumModelSparse(tmax=1, solver=solver, tol=1e-14)
    x, x_dot = model.nodes_free_positions_jvp(T(prob.q), T(prob.xyz_fixed), T(prob.loads), s, T(dq), T(dxs), T(dp))
    eps = 1e-3
    xat = lambda t: oracle.forward(prob.q + t * eps * dq, prob.xyz_fixed + t * eps * dxs, prob.loads + t * eps * dp, so)[1]
    fd = (-xat(2) + 8.0 * xat(1) - 8.0 * xat(-1) + xat(-2)) / (12.0 * eps)          # fourth-order stencil
    x0 = oracle.forward(prob.q, prob.xyz_fixed, prob.loads, so)[1]
    assert np.abs(x.cpu().numpy() - x0).max() <= 1e-10 * np.abs(x0).max()
    assert np.abs(x_dot.cpu().numpy() - fd).max() <= GRAD_RTOL * np.abs(fd).max()


def _oracle_edges_faces(faces_indexed, edges_indexed):
    """structures/meshes.py:146-204 restated: first two incident faces of every edge over both windings."""
    half = {}
    for f, face in enumerate(faces_indexed):
        clean = [int(v) for v in face if v >= 0]
        for u, v in zip(clean, clean[1:] + clean[:1]):
            half.setdefault((u, v), []).append(f)
    out = np.full((len(edges_indexed), 2), -1, dtype=np.int64)
    for e, (u, v) in enumerate(edges_indexed):
        fs = half.get((int(u), int(v)), []) + half.get((int(v), int(u)), [])
        out[e, :len(fs[:2])] = fs[:2]
    return out


def test_face_loads_fixed_point_and_gradients():
    """Area (face) loads on a mesh structure: kernel vs oracle for nodes_load, the tmax > 1 fixed point, and
    the implicit gradients w.r.t. q, xyz_fixed and the face loads against finite differences of the oracle."""
    import jax_fdm_b200.equilibrium as eq

    nx = 5
    m = datasets.meshgrid(nx)
    prob = datasets.cablenet_problem(nx, seed=9)
    # quads plus two degenerate paddings to exercise the -1 handling: turn two quads into triangles + pad
    faces = np.concatenate([m.faces, -np.ones((m.faces.shape[0], 1), dtype=np.int64)], axis=1)
    s = eq.EquilibriumMeshStructureSparse(prob.nodes, faces, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    ef = _oracle_edges_faces(s.faces_indexed, so.edges_indexed)
    assert np.array_equal(ef, s.edges_faces_indexed)
    rng = np.random.default_rng(3)
    fload = np.zeros((faces.shape[0], 3))
    fload[:, 2] = -0.08 * rng.uniform(0.5, 1.5, size=faces.shape[0])
    fload[:, 1] = 0.01 * rng.standard_normal(faces.shape[0])
    x0 = prob.xyz0

    # nodes_load alone, at a non-flat geometry
    xyz = prob.xyz0.copy()
    xyz[:, 2] += 0.3 * np.sin(xyz[:, 0]) * np.cos(0.7 * xyz[:, 1])
    got, _ = eq.ops.face_loads(s, T(xyz), T(prob.loads), T(fload))
    ref = oracle.nodes_load_faces(xyz, prob.loads, fload, so, s.faces_indexed, ef)
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-13 * np.abs(ref).max()

    def oracle_loss(q, xs, fl):
        x, steps = oracle.fixed_point_forward(q, xs, prob.loads, np.zeros((so.num_edges, 3)), so, tmax=200, eta=1e-13,
                                              faces=(fl, s.faces_indexed, ef))
        xyz_ = oracle.nodes_positions(x, xs, so)
        return 0.5 * np.sum((xyz_ - x0) ** 2), x, steps

    q, xs, fl = T(prob.q, True), T(prob.xyz_fixed, True), T(fload, True)
    model = eq.EquilibriumModelSparse(tmax=200, eta=1e-13, tol=1e-14)
    state = model(eq.EquilibriumParametersState(q, xs, eq.LoadState(T(prob.loads), 0.0, fl)), s)
    loss = 0.5 * ((state.xyz - T(x0)) ** 2).sum()
    loss.backward()
    ref_loss, x_ref, steps = oracle_loss(prob.q, prob.xyz_fixed, fload)
    assert 1 < steps < 200
    xf = state.xyz.detach().cpu().numpy()[so.indices_free]
    assert np.abs(xf - x_ref).max() <= 1e-10 * np.abs(x_ref).max()

    def fd(arr, mk, n_probe=5):
        return [(k, fd5(lambda a: oracle_loss(*mk(a))[0], arr, k))
                for k in rng.choice(arr.size, size=min(n_probe, arr.size), replace=False)]

    for grad, arr, mk in [(q.grad, prob.q, lambda a: (a, prob.xyz_fixed, fload)),
                          (xs.grad, prob.xyz_fixed, lambda a: (prob.q, a, fload)),
                          (fl.grad, fload, lambda a: (prob.q, prob.xyz_fixed, a))]:
        gnp = grad.cpu().numpy()
        scale = np.abs(gnp).max()
        for k, refv in fd(arr, mk):
            assert abs(gnp.flat[k] - refv) <= GRAD_RTOL * scale + 1e-11, (k, gnp.flat[k], refv)


# ---------------------------------------------------------------------------------------------------
# local (follower) load frames: is_load_local=True (geometry.py:583-657, loads.py:80-185, 340-405)
# ---------------------------------------------------------------------------------------------------

def _golden_loads():
    import os
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loads_frames.npz")))


@pytest.mark.parametrize("tag", ["quad", "mixed"])
def test_local_frame_loads_vs_reference_arrays_and_fd(tag):
    """nodes_load with edge / face loads given in the element frames: kernels against arrays the reference's
    loads.py produced (quad mesh) and the oracle, then both VJPs against central differences of the oracle."""
    import jax_fdm_b200.equilibrium as eq

    g = _golden_loads()
    s = eq.EquilibriumMeshStructureSparse(g[f"{tag}_nodes"], g[f"{tag}_faces"], g[f"{tag}_edges"], g[f"{tag}_supports"], device=0)
    so = oracle.build_structure(g[f"{tag}_nodes"], g[f"{tag}_edges"], g[f"{tag}_supports"])
    assert np.array_equal(s.faces_indexed, g[f"{tag}_faces_indexed"])
    fi, ef = s.faces_indexed, s.edges_faces_indexed
    xyz, el, fl = g[f"{tag}_xyz"], g[f"{tag}_edges_load"], g[f"{tag}_faces_load"]
    zero = np.zeros_like(xyz)

    got_e = eq.ops.edge_loads(s, T(xyz), T(zero), T(el), is_local=True).cpu().numpy()
    ref_e = g[f"{tag}_ref_nodes_load_edges_local"]
    assert np.abs(got_e - ref_e).max() <= 1e-13 * np.abs(ref_e).max()
    got_f = eq.ops.face_loads(s, T(xyz), T(zero), T(fl), is_local=True)[0].cpu().numpy()
    # padded faces: the kernels read the polygon without its pads (oracle pad="first"); loads.py:170 reads xyz[-1]
    ref_f = oracle.nodes_load_faces(xyz, zero, fl, so, fi, ef, is_local=True, pad="first")
    assert np.abs(got_f - ref_f).max() <= 1e-13 * np.abs(ref_f).max()
    if tag == "quad":
        ref = g["quad_ref_nodes_load_faces_local"]
        assert np.abs(got_f - ref).max() <= 1e-13 * np.abs(ref).max()
    # global mode is untouched by the frame code
    got = eq.ops.face_loads(s, T(xyz), T(zero), T(fl))[0].cpu().numpy()
    ref = g[f"{tag}_ref_nodes_load_faces_global"]
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()

    rng = np.random.default_rng(4)
    cot = rng.standard_normal(xyz.shape)

    def phi_e(x, w):
        return np.sum(cot * oracle.nodes_load_edges(x, zero, w, so, is_local=True))

    def phi_f(x, w):
        return np.sum(cot * oracle.nodes_load_faces(x, zero, w, so, fi, ef, is_local=True, pad="first"))

    xb_e, wb_e = eq.ops.edge_loads_vjp(s, T(xyz), T(el), T(cot), is_local=True)
    _, aux = eq.ops.face_loads(s, T(xyz), T(zero), T(fl), is_local=True)
    xb_f, wb_f = eq.ops.face_loads_vjp(s, T(xyz), T(fl), aux, T(cot), is_local=True)
    for phi, xb, wb, w in ((phi_e, xb_e, wb_e, el), (phi_f, xb_f, wb_f, fl)):
        xb, wb = xb.cpu().numpy(), wb.cpu().numpy()
        for arr, grad, mk in ((xyz, xb, lambda a: (a, w)), (w, wb, lambda a: (xyz, a))):
            scale = np.abs(grad).max()
            for k in rng.choice(arr.size, size=12, replace=False):
                d = np.zeros(arr.size)
                d[k] = 1e-6
                d = d.reshape(arr.shape)
                fdv = (phi(*mk(arr + d)) - phi(*mk(arr - d))) / 2e-6
                assert abs(grad.flat[k] - fdv) <= 2e-8 * scale + 1e-10, (k, grad.flat[k], fdv)


def test_follower_face_loads_fixed_point_and_gradients():
    """After the reference's tests/test_adjoint_routing.py:40-75 (meshgrid, q = -2, face loads [0, 0, -0.1] in the
    FACE frames, is_load_local=True, tmax > 1), here on the corner-supported 5x5 cable net with follower edge
    loads on top: forward vs the oracle, implicit gradients vs central differences of the oracle."""
    import jax_fdm_b200.equilibrium as eq

    nx = 5
    m = datasets.meshgrid(nx)
    prob = datasets.cablenet_problem(nx, seed=2)
    s = eq.EquilibriumMeshStructureSparse(prob.nodes, m.faces, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    fi, ef = s.faces_indexed, s.edges_faces_indexed
    rng = np.random.default_rng(8)
    q0 = -2.0 * np.ones(so.num_edges)
    fload = np.tile([0.0, 0.0, -0.1], (s.num_faces, 1)) + 0.01 * rng.standard_normal((s.num_faces, 3))
    eload = 0.02 * rng.standard_normal((so.num_edges, 3))
    xs0 = prob.xyz_fixed.copy()
    xs0[:, 2] += 0.5 * np.sin(xs0[:, 0])              # non-flat boundary so that the frames rotate
    loads0 = np.zeros_like(prob.loads)
    loads0[:, 2] = -0.05

    def oracle_loss(q, xs, fl, el):
        x, steps = oracle.fixed_point_forward(q, xs, loads0, el, so, tmax=300, eta=1e-13, faces=(fl, fi, ef), is_local=True)
        xyz_ = oracle.nodes_positions(x, xs, so)
        ei = so.edges_indexed
        lengths = np.linalg.norm(xyz_[ei[:, 1]] - xyz_[ei[:, 0]], axis=1)
        return 0.5 * np.sum((lengths - 1.0) ** 2), x, steps

    q, xs, fl, el = T(q0, True), T(xs0, True), T(fload, True), T(eload, True)
    model = eq.EquilibriumModelSparse(tmax=300, eta=1e-13, tol=1e-14, is_load_local=True)
    state = model(eq.EquilibriumParametersState(q, xs, eq.LoadState(T(loads0), el, fl)), s)
    loss = 0.5 * ((state.lengths - 1.0) ** 2).sum()
    loss.backward()
    ref_loss, x_ref, steps = oracle_loss(q0, xs0, fload, eload)
    assert 1 < steps < 300
    xf = state.xyz.detach().cpu().numpy()[so.indices_free]
    assert np.abs(xf - x_ref).max() <= 1e-10 * np.abs(x_ref).max()
    assert abs(loss.item() - ref_loss) <= 1e-10 * abs(ref_loss)
    # the loads reported in the state are the follower loads at the converged geometry
    xyz_ref = oracle.nodes_positions(x_ref, xs0, so)
    loads_ref = oracle.nodes_load_faces(xyz_ref, oracle.nodes_load_edges(xyz_ref, loads0, eload, so, is_local=True),
                                        fload, so, fi, ef, is_local=True)
    assert np.abs(state.loads.detach().cpu().numpy() - loads_ref).max() <= 1e-10 * np.abs(loads_ref).max()

    for grad, arr, mk in [(q.grad, q0, lambda a: (a, xs0, fload, eload)),
                          (xs.grad, xs0, lambda a: (q0, a, fload, eload)),
                          (fl.grad, fload, lambda a: (q0, xs0, a, eload)),
                          (el.grad, eload, lambda a: (q0, xs0, fload, a))]:
        gnp = grad.cpu().numpy()
        scale = np.abs(gnp).max()
        for k in rng.choice(arr.size, size=5, replace=False):
            refv = fd5(lambda a: oracle_loss(*mk(a))[0], arr, k)
            assert abs(gnp.flat[k] - refv) <= GRAD_RTOL * scale + 1e-11, (k, gnp.flat[k], refv)


@pytest.mark.parametrize("tag", ["net6", "net9"])
@pytest.mark.parametrize("local", [False, True])
def test_fixed_point_state_vs_reference_arrays(tag, local):
    """The CUDA fixed point against states the reference's own models.py / fixed_point.py / loads.py produced
    (tests/golden/make_golden_fixed_point.py), same tmax and eta: positions to 1e-9 (both stop on the same rule,
    mean nodal move <= 1e-10), every field of the reported state to 1e-8."""
    import os
    import jax_fdm_b200.equilibrium as eq

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixed_point.npz")))
    s = eq.EquilibriumStructureSparse(g[f"{tag}_nodes"], g[f"{tag}_edges"], g[f"{tag}_supports"], device=0)
    model = eq.EquilibriumModelSparse(tmax=100, eta=1e-10, is_load_local=local, tol=1e-14)
    params = eq.EquilibriumParametersState(T(g[f"{tag}_q"]), T(g[f"{tag}_xyz_fixed"]),
                                           eq.LoadState(T(g[f"{tag}_loads"]), T(g[f"{tag}_edges_load"]), 0.0))
    st = model(params, s)
    key = f"{tag}_{'local' if local else 'global'}"
    ref = g[f"{key}_state_xyz"]
    assert np.abs(st.xyz.cpu().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()
    for f in ("residuals", "lengths", "forces", "loads", "vectors"):
        r = g[f"{key}_state_{f}"]
        got = getattr(st, f).cpu().numpy().reshape(r.shape)
        assert np.abs(got - r).max() <= 1e-8 * max(1.0, np.abs(r).max()), f


@pytest.mark.parametrize("tag,local", [("quad", False), ("quad", True), ("mixed", False)])
def test_face_load_fixed_point_vs_reference_arrays(tag, local):
    """The CUDA fixed point under face (area) loads against states the reference's own mesh structure, models.py,
    fixed_point.py and loads.py produced (tests/golden/make_golden_mesh.py).  (Local loads on the mixed mesh are the
    one documented divergence: the reference reads xyz[-1] for padded faces.)"""
    import os
    import jax_fdm_b200.equilibrium as eq

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_faces.npz")))
    s = eq.EquilibriumMeshStructureSparse(g[f"{tag}_nodes"], g[f"{tag}_faces"], g[f"{tag}_edges"], g[f"{tag}_supports"], device=0)
    model = eq.EquilibriumModelSparse(tmax=100, eta=1e-10, is_load_local=local, tol=1e-14)
    params = eq.EquilibriumParametersState(T(g[f"{tag}_q"]), T(g[f"{tag}_xyz_fixed"]),
                                           eq.LoadState(T(g[f"{tag}_loads"]), 0.0, T(g[f"{tag}_faces_load"])))
    st = model(params, s)
    key = f"{tag}_{'local' if local else 'global'}"
    ref = g[f"{key}_state_xyz"]
    assert np.abs(st.xyz.cpu().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()
    for f in ("residuals", "lengths", "forces", "loads", "vectors"):
        r = g[f"{key}_state_{f}"]
        got = getattr(st, f).cpu().numpy().reshape(r.shape)
        assert np.abs(got - r).max() <= 1e-8 * max(1.0, np.abs(r).max()), f


def test_device_resident_loops_are_pooled_and_match_the_host_driven_loops():
    """The tmax > 1 equilibrium on the band Cholesky runs as two CUDA graphs with conditional WHILE nodes
    (csrc/fixed_point.cu).  Two forwards before either backward must not share one loop object (it keeps the factor and
    the converged geometry between the passes), and the result must match the host-driven loops of the iterative
    solver path (same kernels, convergence read on the host every iteration)."""
    import jax_fdm_b200.equilibrium as eq

    prob, eload = problem(nx=6, seed=5)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    outs = {}
    for solver in ("cholesky", "pcg"):                    # device-resident loops / host-driven loops
        model = eq.EquilibriumModelSparse(tmax=200, eta=1e-13, solver=solver, tol=1e-14)
        leaves, states = [], []
        for scale in (1.0, 1.7):                          # two forwards, then two backwards
            q, w = T(prob.q * scale, True), T(eload, True)
            st = model(eq.EquilibriumParametersState(q, T(prob.xyz_fixed), eq.LoadState(T(prob.loads), w, 0.0)), s)
            leaves.append((q, w))
            states.append(st)
        steps = model.last_solver_config["steps"]
        for st in states:
            (0.5 * (st.xyz ** 2).sum() + 0.1 * (st.lengths ** 2).sum()).backward()
        assert int(steps) > 1
        outs[solver] = [states[0].xyz.detach().cpu().numpy(), states[1].xyz.detach().cpu().numpy()] + \
                       [t.grad.cpu().numpy() for pair in leaves for t in pair]
    if True:
        pool = s.__dict__["_fp_pool"]
        assert sum(len(v) for v in pool.values()) == 2 and not any(o.busy for v in pool.values() for o in v)
    for a, b in zip(outs["cholesky"], outs["pcg"]):
        assert np.abs(a - b).max() <= 1e-9 * max(np.abs(b).max(), 1e-300)
