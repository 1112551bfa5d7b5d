"""
oracle/model.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Forward FDM solve, equilibrium state and the reverse pass, restated on CPU.

Follows (reference file:line):
  models.py:1064-1079   sparse K: data = -q[index_array.data - 1]; diag <- diags @ q
  models.py:800-822     dense  K: C_f^T (q[:,None] * C_f)
  models.py:949-976     R_fixed = C_f^T (q[:,None] * (C_s @ xyz_fixed))
  models.py:828-856     P = loads[indices_free] - R_fixed
  sparse.py:147-149     x = scipy spsolve(csc_matrix((data, indices, indptr)), b)
  models.py:88          dense: jnp.linalg.solve  (LAPACK gesv)  -> numpy.linalg.solve
  models.py:195-220     xyz = concat(x_free, x_fixed)[indices_freefixed]
  models.py:753-794     vectors, lengths, residuals, forces
  sparse.py:267-308     adjoint: lam = solve(K, g); vjp of (b - K x) at lam
The reverse pass of the assembly expressions has no source in the reference (JAX
autodiff generates it); `sparse_solve_bwd` + `forward_adjoint` write out the
transposes of the lines cited above by hand.  They are checked against central
finite differences and Taylor-remainder slopes in tests/ (the reference's own
procedure, tests/test_taylor_convergence.py:72-113).
"""

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve as _spsolve


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------

def stiffness_matrix_sparse(q, s):
    """models.py:1064-1079 -> scipy csc_matrix sharing index_array's pattern."""
    q = np.asarray(q, dtype=np.float64)
    data = -q[s.index_data - 1]                # diag slots read q[-1], overwritten below
    data[s.diag_indices] = s.diags @ q
    n = s.num_free
    return sp.csc_matrix((data, s.index_indices, s.index_indptr), shape=(n, n))


def stiffness_matrix_dense(q, s):
    """models.py:800-822."""
    cf = s.connectivity_free.toarray()
    return cf.T @ (np.asarray(q)[:, None] * cf)


def residual_fixed_matrix(q, xyz_fixed, s):
    """models.py:949-976."""
    q = np.asarray(q, dtype=np.float64)
    return s.connectivity_free.T @ (q[:, None] * (s.connectivity_fixed @ xyz_fixed))


def load_matrix(q, xyz_fixed, loads, s):
    """models.py:828-856."""
    return np.asarray(loads)[s.indices_free, :] - residual_fixed_matrix(q, xyz_fixed, s)


def sparse_solve(K, b):
    """sparse.py:147-149 (SuperLU, COLAMD).  b:(n,3) -> (n,3)."""
    K = sp.csc_matrix((K.data, K.indices, K.indptr), shape=K.shape)
    x = _spsolve(K, np.asarray(b, dtype=np.float64))
    return np.asarray(x).reshape(np.asarray(b).shape)


def dense_solve(K, b):
    """models.py:88."""
    return np.linalg.solve(K, b)


def nodes_positions(xyz_free, xyz_fixed, s):
    """models.py:195-220."""
    return np.concatenate((xyz_free, xyz_fixed))[s.indices_freefixed, :]


def equilibrium_state(q, xyz, loads, s):
    """models.py:753-794 -> dict with the six EquilibriumState fields (states.py:25-50)."""
    q = np.asarray(q, dtype=np.float64)
    C = s.connectivity
    vectors = C @ xyz
    lengths = np.linalg.norm(vectors, axis=1, keepdims=True)
    residuals = loads - C.T @ (q[:, None] * vectors)
    forces = q.reshape(-1, 1) * lengths
    return dict(xyz=xyz, residuals=residuals, lengths=lengths, forces=forces,
                loads=loads, vectors=vectors)


def forward(q, xyz_fixed, loads, s, sparse=True):
    """
    models.py:410-478 at tmax=1 with node loads only: returns (state dict, x_free).
    """
    if sparse:
        K = stiffness_matrix_sparse(q, s)
        b = load_matrix(q, xyz_fixed, loads, s)
        xf = sparse_solve(K, b)
    else:
        K = stiffness_matrix_dense(q, s)
        b = load_matrix(q, xyz_fixed, loads, s)
        xf = dense_solve(K, b)
    xyz = nodes_positions(xf, xyz_fixed, s)
    return equilibrium_state(q, xyz, np.asarray(loads, dtype=np.float64), s), xf


# ----------------------------------------------------------------------------
# reverse
# ----------------------------------------------------------------------------

def equilibrium_state_vjp(q, xyz, s, cot):
    """
    Transpose of models.py:780-794.  `cot` maps any subset of
    {xyz, residuals, lengths, forces, loads, vectors} to cotangent arrays.
    Returns (q_bar (E,), xyz_bar (N,3), loads_bar (N,3)).
    A zero-length edge with a nonzero length/force cotangent yields NaN, as
    jnp.linalg.norm's gradient does.
    """
    q = np.asarray(q, dtype=np.float64)
    C = s.connectivity
    E, N = C.shape
    v = C @ xyz
    ell = np.linalg.norm(v, axis=1, keepdims=True)

    z3 = lambda n: np.zeros((n, 3))
    vbar = np.array(cot.get("vectors", z3(E)), dtype=np.float64)
    lbar = np.array(cot.get("lengths", np.zeros((E, 1))), dtype=np.float64).reshape(E, 1)
    fbar = np.array(cot.get("forces", np.zeros((E, 1))), dtype=np.float64).reshape(E, 1)
    rbar = np.array(cot.get("residuals", z3(N)), dtype=np.float64)
    xbar = np.array(cot.get("xyz", z3(N)), dtype=np.float64)
    pbar = np.array(cot.get("loads", z3(N)), dtype=np.float64)

    qbar = np.zeros(E)
    # forces = q * lengths
    qbar += (fbar * ell)[:, 0]
    ltot = lbar + q[:, None] * fbar
    # residuals = loads - C^T (q * v)
    Crb = C @ rbar
    qbar += -np.sum(Crb * v, axis=1)
    vtot = vbar - q[:, None] * Crb
    # lengths = |v|
    if np.any(ltot != 0.0):
        with np.errstate(invalid="ignore", divide="ignore"):
            vtot = vtot + np.where(ltot != 0.0, ltot * (v / ell), 0.0)
    # vectors = C xyz
    xbar = xbar + C.T @ vtot
    pbar = pbar + rbar
    return qbar, xbar, pbar


def sparse_solve_bwd(K, x, g):
    """
    sparse.py:291-308: lam = solve(K, g); cotangents of residual_fn(A,b) = b - A@x at lam:
    Kbar.data[k] = -sum_xyz lam[row_k] * x[col_k],  bbar = lam.
    """
    lam = sparse_solve(K, g)
    n = K.shape[0]
    cols = np.repeat(np.arange(n), np.diff(K.indptr))
    rows = K.indices
    Kbar = -np.sum(lam[rows] * x[cols], axis=1)
    return lam, Kbar, lam


def forward_adjoint(q, xyz_fixed, loads, s, cot_fn):
    """
    One forward solve + one reverse pass (sparse model, tmax=1).
    cot_fn(state, x_free) -> dict of cotangents on the state fields.
    Returns dict(state, x_free, lam, q_bar, xyz_fixed_bar, loads_bar).
    """
    q = np.asarray(q, dtype=np.float64)
    loads = np.asarray(loads, dtype=np.float64)
    xyz_fixed = np.asarray(xyz_fixed, dtype=np.float64)
    K = stiffness_matrix_sparse(q, s)
    b = load_matrix(q, xyz_fixed, loads, s)
    xf = sparse_solve(K, b)
    xyz = nodes_positions(xf, xyz_fixed, s)
    state = equilibrium_state(q, xyz, loads, s)

    cot = cot_fn(state, xf)
    qbar, xbar, pbar = equilibrium_state_vjp(q, xyz, s, cot)
    # transpose of nodes_positions (models.py:218-220)
    g = xbar[s.indices_free]
    xsbar = xbar[s.indices_fixed].copy()

    lam, Kbar, bbar = sparse_solve_bwd(K, xf, g)

    # transpose of stiffness_matrix (models.py:1069-1077)
    kb_off = Kbar.copy()
    kb_diag = Kbar[s.diag_indices]
    kb_off[s.diag_indices] = 0.0                     # .at[diag].set overwrites those slots
    np.add.at(qbar, s.index_data - 1, -kb_off)       # gather transpose (diag slots add 0 to q[-1])
    qbar += s.diags.T @ kb_diag

    # transpose of load_matrix / residual_fixed_matrix (models.py:856, 973-976)
    pbar[s.indices_free] += bbar
    Cf, Cs = s.connectivity_free, s.connectivity_fixed
    Cf_lam = Cf @ bbar                                # (E,3)
    qbar += -np.sum(Cf_lam * (Cs @ xyz_fixed), axis=1)
    xsbar += -(Cs.T @ (q[:, None] * Cf_lam))

    return dict(state=state, x_free=xf, lam=lam, q_bar=qbar,
                xyz_fixed_bar=xsbar, loads_bar=pbar)


# ----------------------------------------------------------------------------
# shape-dependent loads and the fixed-point equilibrium (tmax > 1)
# ----------------------------------------------------------------------------

WORLD_X, WORLD_Y, WORLD_Z = np.eye(3)


def _allclose(a, b):
    """jnp.allclose defaults (rtol 1e-5, atol 1e-8), as the frame code uses them."""
    return bool(np.all(np.abs(np.asarray(a) - b) <= 1e-8 + 1e-5 * np.abs(b)))


def normal_polygon(polygon):
    """geometry.py:383-425 (unitized; a near-zero normal comes back as the zero vector, :138-168)."""
    r = np.asarray(polygon, dtype=np.float64)
    r = r - r.mean(axis=0)
    n = (0.5 * np.cross(np.roll(r, 1, axis=0), r)).sum(axis=0)
    if _allclose(n, 0.0):
        return np.zeros(3)
    return n / np.linalg.norm(n)


def line_lcs(line):
    """geometry.py:583-618: rows U, V, W of the edge frame."""
    line = np.asarray(line, dtype=np.float64)
    u = line[1] - line[0]
    u = u / np.linalg.norm(u)
    vperp = WORLD_Y if _allclose(abs(WORLD_Z @ u), 1.0) else WORLD_Z
    v = np.cross(vperp, u)
    v = v / np.linalg.norm(v)
    w = np.cross(u, v)
    if w @ vperp < 0.0:
        w = -w
    w = w / np.linalg.norm(w)
    return np.vstack((u, v, w))


def polygon_lcs(polygon):
    """geometry.py:621-657: rows U, V, W of the face frame."""
    w = normal_polygon(polygon)
    vperp = WORLD_Y if _allclose(abs(WORLD_X @ w), 1.0) else WORLD_X
    v = np.cross(w, vperp)
    v = v / np.linalg.norm(v)
    u = np.cross(w, v)
    if u @ vperp < 0.0:
        u = -u
    u = u / np.linalg.norm(u)
    return np.vstack((u, v, w))


def edges_load_global(xyz, edges_load, s, is_local):
    """loads.py:340-405: edge loads given in the edge frames rotated to global coordinates."""
    edges_load = np.asarray(edges_load, dtype=np.float64)
    if not is_local:
        return edges_load
    return np.stack([load @ line_lcs(xyz[edge]) for edge, load in zip(s.edges_indexed, edges_load)])


def faces_load_global(xyz, faces_load, faces_indexed, is_local, pad="wrap"):
    """
    loads.py:80-185: face loads given in the face frames rotated to global coordinates; a face with a
    (near) zero normal keeps its load unrotated.  `pad`: how the -1 padding of faces with fewer
    vertices than the widest face is read.  "wrap" is what loads.py:170 does (xyz[face] with -1 ->
    the LAST node of the structure, an artefact); "first" is the face_xyz helper of loads.py:108-145
    the reference leaves commented out at :169 (pads repeat the first vertex: same normal as the
    unpadded polygon).  The CUDA path drops the pads, which equals "first".
    """
    faces_load = np.asarray(faces_load, dtype=np.float64)
    if not is_local:
        return faces_load
    out = np.empty_like(faces_load)
    for i, (face, load) in enumerate(zip(faces_indexed, faces_load)):
        face = np.asarray(face)
        if pad == "first":
            face = np.where(face >= 0, face, face[0])
        fxyz = xyz[face]
        lcs = np.eye(3) if _allclose(normal_polygon(fxyz), 0.0) else polygon_lcs(fxyz)
        out[i] = load @ lcs
    return out


def nodes_load_edges(xyz, loads_nodes, edges_load, s, is_local=False):
    """
    models.py:290-344 with loads.py:296-331, 340-465: line loads (rotated from the edge frames when
    `is_local`) times edge length, half to each end node, added to the nodal loads.
    """
    ei = s.edges_indexed
    lengths = np.linalg.norm(xyz[ei[:, 1]] - xyz[ei[:, 0]], axis=1, keepdims=True)
    total = edges_load_global(xyz, edges_load, s, is_local) * lengths
    out = np.array(loads_nodes, dtype=np.float64, copy=True)
    np.add.at(out, ei[:, 0], 0.5 * total)
    np.add.at(out, ei[:, 1], 0.5 * total)
    return out


def nodes_load_faces(xyz, loads_nodes, faces_load, s, faces_indexed, edges_faces_indexed, is_local=False,
                     pad="wrap"):
    """
    models.py:333-344 with loads.py:35-71, 80-288: every edge takes, from each of its (first two)
    incident faces, the face load (rotated from the face frame when `is_local`) times the area of the
    triangle (edge, face centroid) (loads.py:226-288); half of the edge total goes to each end node
    (loads.py:435-465).  `faces_indexed` (F, maxdeg) padded -1 and `edges_faces_indexed` (E, 2) padded -1
    as the reference defines them (structures/meshes.py:108-204).
    """
    faces_load = faces_load_global(xyz, faces_load, faces_indexed, is_local, pad)
    mask = faces_indexed >= 0
    cen = np.stack([xyz[face[m]].mean(axis=0) for face, m in zip(faces_indexed, mask)])
    ei = s.edges_indexed
    a, b = xyz[ei[:, 0]], xyz[ei[:, 1]]
    edge_load = np.zeros((ei.shape[0], 3))
    for slot in range(2):
        f = edges_faces_indexed[:, slot]
        ok = f >= 0
        c = cen[np.where(ok, f, 0)]
        area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
        edge_load += np.where(ok, area, 0.0)[:, None] * faces_load[np.where(ok, f, 0)]
    out = np.array(loads_nodes, dtype=np.float64, copy=True)
    np.add.at(out, ei[:, 0], 0.5 * edge_load)
    np.add.at(out, ei[:, 1], 0.5 * edge_load)
    return out


def fixed_point_forward(q, xyz_fixed, loads_nodes, edges_load, s, tmax=100, eta=1e-6, faces=None, is_local=False):
    """
    models.py:410-478, 512-618 + fixed_point.py:139-198: one linear step with the nodal loads, then
    x <- K^{-1} (nodes_load(xyz(x))[free] - R_fixed) until the mean nodal move <= eta or tmax steps.
    Returns (x_free, steps).  The reference re-solves with SuperLU every iteration; so does this.
    """
    K = stiffness_matrix_sparse(q, s)
    r_fixed = residual_fixed_matrix(q, xyz_fixed, s)

    def f(x):
        xyz = nodes_positions(x, xyz_fixed, s)
        loads = nodes_load_edges(xyz, loads_nodes, edges_load, s, is_local)
        if faces is not None:      # (faces_load, faces_indexed, edges_faces_indexed)
            loads = nodes_load_faces(xyz, loads, faces[0], s, faces[1], faces[2], is_local)
        return sparse_solve(K, loads[s.indices_free, :] - r_fixed)

    x_prev = sparse_solve(K, np.asarray(loads_nodes)[s.indices_free, :] - r_fixed)
    x = f(x_prev)
    steps = 0
    while steps < tmax and np.mean(np.linalg.norm(x_prev - x, axis=1)) > eta:
        x_prev, x = x, f(x)
        steps += 1
    return x, steps
