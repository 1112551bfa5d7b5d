"""
CPU: the oracle's restatement of the local coordinate systems and of the shape-dependent (tributary)
loads against (1) the reference's own known-answer tests (tests/test_lcs.py:19-128) and (2) arrays
produced by the reference's real geometry.py / loads.py (tests/golden/make_golden_loads.py).
"""
import os

import numpy as np
import pytest

import oracle
from oracle import model as om

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loads_frames.npz")
X, Y, Z = np.eye(3)


@pytest.fixture(scope="module")
def g():
    return dict(np.load(GOLDEN))


def _square_xy():
    # compas Polygon.from_sides_and_radius_xy(4, 1.0): unit-radius square, counter-clockwise from +Y
    return np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [1.0, 0.0, 0.0]])


def test_reference_kat_polygon_frames():
    sq = _square_xy()
    assert np.allclose(om.polygon_lcs(sq), np.eye(3))                                  # test_lcs.py:19-36
    assert np.allclose(np.array([0, 0, 2.0]) @ om.polygon_lcs(sq), 2.0 * Z)
    zx = np.c_[sq[:, 1], 0 * sq[:, 0], sq[:, 0]]                                       # worldXY -> worldZX: x->Z, y->X
    assert np.allclose(om.polygon_lcs(zx[::-1]), np.vstack((X, Z, -Y)))                # test_lcs.py:39-60
    assert np.allclose(np.array([0, 0, 2.0]) @ om.polygon_lcs(zx[::-1]), -2.0 * Y)
    yz = np.c_[0 * sq[:, 0], sq[:, 0], sq[:, 1]]                                       # worldXY -> worldYZ: x->Y, y->Z
    assert np.allclose(om.polygon_lcs(yz), np.vstack((Y, Z, X)))                       # test_lcs.py:63-83
    assert np.allclose(np.array([0, 0, -2.0]) @ om.polygon_lcs(yz), -2.0 * X)


def test_reference_kat_line_frames():
    assert np.allclose(om.line_lcs([[0, 0, 0], [1.0, 0, 0]]), np.eye(3))               # test_lcs.py:86-94
    assert np.allclose(om.line_lcs([[0, 0, 0], [0, 1.0, 0]]), [[0, 1, 0], [-1, 0, 0], [0, 0, 1]])   # :97-106
    assert np.allclose(om.line_lcs([[0, 0, 0], [0, 0, 1.0]]), [[0, 0, 1], [1, 0, 0], [0, 1, 0]])    # :109-118


def test_frames_vs_reference_arrays(g):
    got = np.stack([om.line_lcs(l) for l in g["lines"]])
    assert np.abs(got - g["ref_line_lcs"]).max() <= 1e-15
    for deg in (3, 4, 6):
        polys, ref_n, ref_f = g[f"polys{deg}"], g[f"ref_normal{deg}"], g[f"ref_polygon_lcs{deg}"]
        for p, n, f in zip(polys, ref_n, ref_f):
            assert np.abs(om.normal_polygon(p) - n).max() <= 1e-15
            if np.any(n != 0.0):                 # the reference's frame of a zero normal is NaN; loads.py:178-179 masks it
                assert np.abs(om.polygon_lcs(p) - f).max() <= 1e-15
        assert np.all(ref_n[4] == 0.0)


@pytest.mark.parametrize("tag", ["quad", "mixed"])
@pytest.mark.parametrize("local", [False, True])
def test_nodes_load_vs_reference_arrays(g, tag, local):
    so = oracle.build_structure(g[f"{tag}_nodes"], g[f"{tag}_edges"], g[f"{tag}_supports"])
    xyz, k = g[f"{tag}_xyz"], "local" if local else "global"
    zero = np.zeros_like(xyz)
    got = oracle.nodes_load_edges(xyz, zero, g[f"{tag}_edges_load"], so, is_local=local)
    ref = g[f"{tag}_ref_nodes_load_edges_{k}"]
    assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()
    got = oracle.nodes_load_faces(xyz, zero, g[f"{tag}_faces_load"], so, g[f"{tag}_faces_indexed"],
                                  g[f"{tag}_edges_faces_indexed"], is_local=local)
    ref = g[f"{tag}_ref_nodes_load_faces_{k}"]
    assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()


def test_padded_faces_read_the_last_node_in_the_reference(g):
    """loads.py:170 indexes xyz with the -1 padding; the geometric reading ('first') differs on the padded faces."""
    so = oracle.build_structure(g["mixed_nodes"], g["mixed_edges"], g["mixed_supports"])
    xyz, fl, fi = g["mixed_xyz"], g["mixed_faces_load"], g["mixed_faces_indexed"]
    wrap = om.faces_load_global(xyz, fl, fi, True, pad="wrap")
    first = om.faces_load_global(xyz, fl, fi, True, pad="first")
    padded = (fi < 0).any(axis=1)
    assert np.abs(wrap[~padded] - first[~padded]).max() == 0.0
    assert np.abs(wrap[padded] - first[padded]).max() > 1e-3
    for face, load, out in zip(fi[padded], fl[padded], first[padded]):     # 'first' == the polygon without its pads
        assert np.abs(load @ om.polygon_lcs(xyz[face[face >= 0]]) - out).max() <= 1e-14
