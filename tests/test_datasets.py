"""The synthetic meshgrid generator (ours) against a literal halfedge-dict emulation."""

import itertools

import numpy as np
import pytest

from jax_fdm_b200 import datasets


def halfedge_edges(nx, ny):
    n = (nx + 1) * (ny + 1)
    he = {u: {} for u in range(n)}
    for i, j in itertools.product(range(ny), range(nx)):
        a = i * (nx + 1) + j
        f = [a, a + 1, a + nx + 2, a + nx + 1]
        for u, v in zip(f, f[1:] + f[:1]):
            he[u][v] = 1
            if u not in he[v]:
                he[v][u] = None
    seen, out = set(), []
    for u in he:
        for v in he[u]:
            if (u, v) in seen or (v, u) in seen:
                continue
            seen.add((u, v))
            seen.add((v, u))
            out.append((u, v))
    return np.asarray(out, dtype=np.int64)


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 2), (3, 3), (5, 5), (10, 10), (4, 7), (7, 3)])
def test_meshgrid_edges_closed_form(nx, ny):
    assert np.array_equal(datasets.meshgrid_edges(nx, ny), halfedge_edges(nx, ny))


def test_config_sizes():
    # SURVEY.md §8 size table
    p = datasets.cablenet_readme_problem(10)
    assert (p.nodes.size, p.edges.shape[0], int(p.supports.sum())) == (121, 220, 4)
    p = datasets.pillow_problem(30)
    assert (p.nodes.size, p.edges.shape[0], int(p.supports.sum())) == (961, 1860, 120)
    p = datasets.cablenet_problem(200)
    assert (p.nodes.size, p.edges.shape[0], int(p.supports.sum())) == (40401, 80400, 4)
    assert p.q.min() >= 0.5 and p.q.max() <= 2.0


def test_pillow_batch_is_shardable():
    a = datasets.pillow_batch_q(1860, 0, 8)
    b = datasets.pillow_batch_q(1860, 4, 4)
    assert np.array_equal(a[4:], b)
    assert np.all(a < 0)
