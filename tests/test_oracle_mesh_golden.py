"""
CPU: mesh structures and face (area) loads against arrays produced by the reference's OWN structures/meshes.py,
structures.py, models.py, fixed_point.py and loads.py (tests/golden/make_golden_mesh.py -> mesh_faces.npz):
the host mirror's face tables bit for bit, and the oracle's tmax > 1 fixed point under face loads in the global
and in the local (follower) frames, including the mixed triangle / quad mesh whose padded faces the reference
reads with xyz[-1] (oracle pad="wrap").
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from jax_fdm_b200.equilibrium.structures import EquilibriumMeshStructureSparse

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_faces.npz")


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


@pytest.mark.parametrize("tag", ["quad", "mixed"])
def test_mesh_tables_bit_exact_vs_reference(g, tag):
    s = EquilibriumMeshStructureSparse(g[f"{tag}_nodes"], g[f"{tag}_faces"], g[f"{tag}_edges"], g[f"{tag}_supports"], device=None)
    assert np.array_equal(s.faces_indexed, g[f"{tag}_ref_faces_indexed"])
    assert np.array_equal(s.edges_faces_indexed, g[f"{tag}_ref_edges_faces_indexed"])
    cfv = sp.csr_matrix(s.connectivity_faces_vertices)
    cfv.sort_indices()
    assert np.array_equal(cfv.indices, g[f"{tag}_ref_cfv_indices"]) and np.array_equal(cfv.indptr, g[f"{tag}_ref_cfv_indptr"])
    assert np.array_equal(cfv.data, g[f"{tag}_ref_cfv_data"])


@pytest.mark.parametrize("tag", ["quad", "mixed"])
@pytest.mark.parametrize("local", [False, True])
def test_face_load_fixed_point_vs_reference(g, tag, local):
    so = oracle.build_structure(g[f"{tag}_nodes"], g[f"{tag}_edges"], g[f"{tag}_supports"])
    q, xs, loads, fload = g[f"{tag}_q"], g[f"{tag}_xyz_fixed"], g[f"{tag}_loads"], g[f"{tag}_faces_load"]
    fi, ef = g[f"{tag}_ref_faces_indexed"], g[f"{tag}_ref_edges_faces_indexed"]
    key = f"{tag}_{'local' if local else 'global'}"
    x, steps = oracle.fixed_point_forward(q, xs, loads, np.zeros((so.num_edges, 3)), so, tmax=100, eta=1e-10,
                                          faces=(fload, fi, ef), is_local=local)
    assert steps == int(g[f"{key}_steps"])
    xyz = oracle.nodes_positions(x, xs, so)
    ref = g[f"{key}_state_xyz"]
    assert np.abs(xyz - ref).max() <= 1e-12 * np.abs(ref).max()
    loads_star = oracle.nodes_load_faces(xyz, loads, fload, so, fi, ef, is_local=local)
    st = oracle.equilibrium_state(q, xyz, loads_star, so)
    for f in ("residuals", "lengths", "forces", "loads", "vectors"):
        r = g[f"{key}_state_{f}"]
        assert np.abs(np.asarray(st[f]) - r).max() <= 1e-11 * max(1.0, np.abs(r).max()), f
