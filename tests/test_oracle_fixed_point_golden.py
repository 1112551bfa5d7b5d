"""
CPU: the oracle's tmax > 1 fixed-point equilibrium with edge loads in the global and in the local (follower)
frames against states produced by the reference's OWN models.py / solvers/fixed_point.py / loads.py
(tests/golden/make_golden_fixed_point.py -> tests/golden/fixed_point.npz): converged positions, iteration
counts, and the six fields of the final equilibrium state.
"""
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixed_point.npz")


@pytest.mark.parametrize("tag", ["net6", "net9"])
@pytest.mark.parametrize("local", [False, True])
def test_fixed_point_forward_vs_reference(tag, local):
    g = dict(np.load(GOLDEN))
    so = oracle.build_structure(g[f"{tag}_nodes"], g[f"{tag}_edges"], g[f"{tag}_supports"])
    q, xs, loads, eload = g[f"{tag}_q"], g[f"{tag}_xyz_fixed"], g[f"{tag}_loads"], g[f"{tag}_edges_load"]
    key = f"{tag}_{'local' if local else 'global'}"
    x, steps = oracle.fixed_point_forward(q, xs, loads, eload, so, tmax=100, eta=1e-10, is_local=local)
    assert steps == int(g[f"{key}_steps"])
    xyz = oracle.nodes_positions(x, xs, so)
    ref = g[f"{key}_state_xyz"]
    assert np.abs(xyz - ref).max() <= 1e-12 * np.abs(ref).max()
    # models.py:466-478: the reported loads are the shape-dependent loads at the converged geometry
    loads_star = oracle.nodes_load_edges(xyz, loads, eload, so, is_local=local)
    st = oracle.equilibrium_state(q, xyz, loads_star, so)
    for f in ("residuals", "lengths", "forces", "loads", "vectors"):
        r = g[f"{key}_state_{f}"]
        assert np.abs(np.asarray(st[f]) - r).max() <= 1e-11 * max(1.0, np.abs(r).max()), f
