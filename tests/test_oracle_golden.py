"""
Pin the CPU oracle (oracle/) against
  (1) arrays produced by the reference's own code (tests/golden/*.npz, made by
      tests/golden/make_golden.py), bit-exact for integers, 1e-12 for floats;
  (2) the reference's known-answer tests: arch baseline JSON
      (tests/baselines/arch_coordinates.json, rtol 1e-6/atol 1e-9), olympia and
      stuttgart21 (tests/test_reference_solver.py:90-102, atol 1e-8), Liew dome
      statistics (tests/test_liew_dome.py:45,152-165), README cable-net table
      (docs/examples/cablenet.md:92-95), arch invariants (tests/test_solver.py:11-56);
  (3) finite differences + Taylor-remainder slope for the hand-written reverse pass
      (procedure of tests/test_taylor_convergence.py:72-113).
CPU only.
"""

import os

import numpy as np
import pytest

import oracle
from jax_fdm_b200 import datasets

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["arch10", "readme10", "cablenet12", "pillow30", "olympia", "stuttgart21",
         "liew_dome", "random40"]


def load(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


@pytest.mark.parametrize("name", CASES)
def test_structure_bit_exact(name):
    g = load(name)
    s = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    for mine, ref in [
        (s.edges_indexed, "ref_edges_indexed"), (s.indices_free, "ref_indices_free"),
        (s.indices_fixed, "ref_indices_fixed"), (s.indices_freefixed, "ref_indices_freefixed"),
        (s.index_data, "ref_index_data"), (s.index_indices, "ref_index_indices"),
        (s.index_indptr, "ref_index_indptr"), (s.diag_indices, "ref_diag_indices"),
    ]:
        assert np.array_equal(np.asarray(mine), g[ref]), ref
    d = s.diags.copy()
    d.sort_indices()
    assert np.array_equal(d.indices, g["ref_diags_indices"])
    assert np.array_equal(d.indptr, g["ref_diags_indptr"])
    assert np.all(d.data == 1.0)
    for m, nm in [(s.connectivity_free, "cfree"), (s.connectivity_fixed, "cfixed"),
                  (s.connectivity, "conn")]:
        assert np.array_equal(m.data, g[f"ref_{nm}_data"])
        assert np.array_equal(m.indices, g[f"ref_{nm}_indices"])
        assert np.array_equal(m.indptr, g[f"ref_{nm}_indptr"])
    # properties SURVEY.md Appendix A claims for the pattern
    n = s.num_free
    for i in range(n):
        row = s.index_indices[s.index_indptr[i]:s.index_indptr[i + 1]]
        assert np.all(np.diff(row) > 0), "indices sorted within each column"
    assert s.nnz == n + 2 * int(np.sum((s.supports[s.edges_indexed] == 0).all(axis=1)))


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    g = load(name)
    s = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    K = oracle.stiffness_matrix_sparse(g["q"], s)
    assert np.array_equal(K.data, g["ref_K_data"])            # same gather + same row sums
    P = oracle.load_matrix(g["q"], g["xyz_fixed"], g["loads"], s)
    np.testing.assert_allclose(P, g["ref_P"], rtol=0, atol=1e-13)
    st, xf = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], s, sparse=True)
    scale = max(1.0, np.abs(g["ref_x_free"]).max())
    np.testing.assert_allclose(xf, g["ref_x_free"], rtol=0, atol=1e-12 * scale)
    for f in ("xyz", "residuals", "lengths", "forces", "loads", "vectors"):
        ref = g[f"ref_state_{f}"]
        np.testing.assert_allclose(st[f], ref, rtol=1e-12, atol=1e-11 * scale, err_msg=f)
        assert st[f].shape == ref.shape
    if "ref_dense_K" in g:
        np.testing.assert_allclose(oracle.stiffness_matrix_dense(g["q"], s), g["ref_dense_K"],
                                   rtol=0, atol=1e-13)
        std, _ = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], s, sparse=False)
        for f in ("xyz", "residuals", "lengths", "forces"):
            np.testing.assert_allclose(std[f], g[f"ref_dense_state_{f}"], rtol=1e-9,
                                       atol=1e-9 * scale, err_msg=f)
        # tests/test_sparse_parity.py:44-73: dense == sparse to 1e-6
        np.testing.assert_allclose(std["xyz"], st["xyz"], atol=1e-6)


def test_kat_arch_baseline_and_invariants():
    g = load("arch10")
    s = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    st, _ = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], s)
    np.testing.assert_allclose(st["xyz"], g["kat_arch_coordinates"], rtol=1e-6, atol=1e-9)
    free = s.indices_free
    assert np.allclose(st["residuals"][free], 0.0, atol=1e-9)                 # test_solver.py:11-20
    reactions = st["residuals"][s.indices_fixed]                               # reaction = residual at supports
    assert np.allclose(np.abs(reactions.sum(0)), np.abs(g["loads"][free].sum(0)), atol=1e-9)
    z = st["xyz"][:, 2]
    assert np.allclose(z, z[::-1], atol=1e-9)                                  # test_solver.py:38-47


@pytest.mark.parametrize("name", ["olympia", "stuttgart21"])
def test_kat_reference_solver(name):
    g = load(name)
    s = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    st, _ = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], s, sparse=False)
    np.testing.assert_allclose(st["xyz"], g["xyz0"], atol=1e-8)


def test_kat_liew_dome():
    g = load("liew_dome")
    s = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    st, _ = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], s)
    ell = st["lengths"][:, 0]
    load_path = np.sum(np.abs(st["forces"][:, 0]) * ell)
    assert abs(load_path - 402.0) <= 0.5
    assert abs(ell.min() - 0.153) <= 1e-3
    assert abs(ell.max() - 2.143) <= 1e-3
    assert abs(ell.mean() - 0.960) <= 1e-3
    assert abs(np.mean(np.abs(ell - 1.0)) - 0.320) <= 1e-3


def test_kat_readme_cablenet_table():
    p = datasets.cablenet_readme_problem(10)
    s = oracle.build_structure(p.nodes, p.edges, p.supports)
    st, _ = oracle.forward(p.q, p.xyz_fixed, p.loads, s)
    w = 11
    i, j = np.divmod(p.edges, w)
    bnd = ((i[:, 0] == i[:, 1]) & ((i[:, 0] == 0) | (i[:, 0] == 10))) | \
          ((j[:, 0] == j[:, 1]) & ((j[:, 0] == 0) | (j[:, 0] == 10)))
    f, ell = st["forces"][:, 0], st["lengths"][:, 0]
    assert round(f[bnd].min(), 2) == 15.16 and round(f[bnd].max(), 2) == 18.27
    assert abs(ell[~bnd].min() - 0.78) < 0.015 and round(ell[~bnd].max(), 2) == 1.80
    assert np.all(f > 0)                                       # "all of its cables in tension"


# ---------------------------------------------------------------------------
# reverse pass: FD + Taylor slope (the reference has no source for it)
# ---------------------------------------------------------------------------

def _loss_and_cot(weights):
    def loss(st, xf):
        return sum(float(np.sum(weights[f] * st[f])) for f in weights)

    def cot(st, xf):
        return {f: weights[f] for f in weights}
    return loss, cot


@pytest.mark.parametrize("name", ["cablenet12", "random40", "arch10"])
def test_adjoint_finite_differences(name):
    g = load(name)
    s = oracle.build_structure(g["nodes"], g["edges"], g["supports"])
    rng = np.random.default_rng(3)
    st0, _ = oracle.forward(g["q"], g["xyz_fixed"], g["loads"], s)
    weights = {f: rng.normal(size=st0[f].shape) for f in st0}
    loss, cot = _loss_and_cot(weights)
    out = oracle.forward_adjoint(g["q"], g["xyz_fixed"], g["loads"], s, cot)

    def L(q, xs, p):
        st, xf = oracle.forward(q, xs, p, s)
        return loss(st, xf)

    dq = rng.normal(size=g["q"].shape)
    dxs = rng.normal(size=g["xyz_fixed"].shape)
    dp = rng.normal(size=g["loads"].shape)
    lin = float(np.sum(out["q_bar"] * dq) + np.sum(out["xyz_fixed_bar"] * dxs)
                + np.sum(out["loads_bar"] * dp))
    h = 1e-5
    fd = (L(g["q"] + h * dq, g["xyz_fixed"] + h * dxs, g["loads"] + h * dp)
          - L(g["q"] - h * dq, g["xyz_fixed"] - h * dxs, g["loads"] - h * dp)) / (2 * h)
    assert abs(fd - lin) <= 1e-6 * max(1.0, abs(lin))

    # Taylor remainder: |L(p + h d) - L(p) - h <grad, d>| = O(h^2)  -> slope 2.0 +- 0.1
    L0 = L(g["q"], g["xyz_fixed"], g["loads"])
    hs = np.array([1e-2, 5e-3, 2.5e-3, 1.25e-3])
    rem = [abs(L(g["q"] + t * dq, g["xyz_fixed"] + t * dxs, g["loads"] + t * dp) - L0 - t * lin)
           for t in hs]
    slope = np.polyfit(np.log(hs), np.log(rem), 1)[0]
    assert abs(slope - 2.0) <= 0.1

    # negative control (tests/test_taylor_convergence.py:312-329): a wrong gradient gives slope 1
    rem_bad = [abs(L(g["q"] + t * dq, g["xyz_fixed"] + t * dxs, g["loads"] + t * dp) - L0 - t * 0.5 * lin)
               for t in hs]
    assert abs(np.polyfit(np.log(hs), np.log(rem_bad), 1)[0] - 1.0) <= 0.1


def test_adjoint_closed_form_identities():
    """SURVEY.md Appendix B items 2-4 for a loss on x_free only."""
    p = datasets.cablenet_problem(9, seed=1)
    s = oracle.build_structure(p.nodes, p.edges, p.supports)
    rng = np.random.default_rng(0)
    w = rng.normal(size=(s.num_nodes, 3))
    w[s.indices_fixed] = 0.0
    out = oracle.forward_adjoint(p.q, p.xyz_fixed, p.loads, s, lambda st, xf: {"xyz": w})
    lam = np.zeros((s.num_nodes, 3))
    lam[s.indices_free] = out["lam"]
    xyz = out["state"]["xyz"]
    t, h = s.edges_indexed[:, 0], s.edges_indexed[:, 1]
    qbar = -np.sum((lam[h] - lam[t]) * (xyz[h] - xyz[t]), axis=1)
    np.testing.assert_allclose(out["q_bar"], qbar, atol=1e-12)
    np.testing.assert_allclose(out["loads_bar"], lam, atol=1e-14)
    xs = np.zeros((s.num_nodes, 3))
    for e in range(s.num_edges):
        a, b = t[e], h[e]
        if s.supports[a] and not s.supports[b]:
            xs[a] += p.q[e] * lam[b]
        if s.supports[b] and not s.supports[a]:
            xs[b] += p.q[e] * lam[a]
    np.testing.assert_allclose(out["xyz_fixed_bar"], xs[s.indices_fixed], atol=1e-12)
