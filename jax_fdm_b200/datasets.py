"""
Synthetic array-level inputs (the COMPAS-free front door).

The reference builds its inputs from COMPAS datastructures
(`FDMesh.from_meshgrid`, `datastructures/mesh.py`), which are third-party and not
part of the equilibrium hot path.  The generator below is OURS by definition
(SURVEY.md §8d): it yields plain numpy arrays (`nodes`, `edges`, `supports`, `xyz`,
`q`, `loads`) that are fed identically to the CUDA path and to the CPU oracle.

Numbering: vertices row-major over an (ny+1) x (nx+1) lattice; quads listed as
(a, a+1, a+nx+2, a+nx+1); edges enumerated the way a halfedge dictionary built
from those quads would list them (`for u in halfedge: for v in halfedge[u]`,
skipping pairs already seen).  `meshgrid_edges` is the closed form of that
enumeration; `tests/test_datasets.py` checks it against a literal dictionary
emulation.
"""

from typing import NamedTuple

import numpy as np

__all__ = [
    "MeshgridArrays",
    "meshgrid",
    "meshgrid_edges",
    "cablenet_problem",
    "cablenet_readme_problem",
    "pillow_problem",
    "pillow_batch_q",
    "arch_problem",
    "arch_benchmark_problem",
]


class MeshgridArrays(NamedTuple):
    nodes: np.ndarray      # (N,) int64 keys 0..N-1
    xyz: np.ndarray        # (N,3) float64
    edges: np.ndarray      # (E,2) int64 (tail, head)
    faces: np.ndarray      # (F,4) int64
    nx: int
    ny: int


def meshgrid_edges(nx: int, ny: int) -> np.ndarray:
    """Closed form of the halfedge-dictionary edge enumeration (see module doc)."""
    w = nx + 1
    u = np.arange((nx + 1) * (ny + 1), dtype=np.int64)
    i, j = np.divmod(u, w)
    has_r = j < nx
    has_d = i < ny
    # slot 0 / slot 1 per vertex; row 0 (except vertex 0) lists "down" before "right"
    swap = (i == 0) & (j >= 1)
    first_is_r = ~swap
    heads = np.full((u.size, 2), -1, dtype=np.int64)
    right = np.where(has_r, u + 1, -1)
    down = np.where(has_d, u + w, -1)
    heads[:, 0] = np.where(first_is_r, right, down)
    heads[:, 1] = np.where(first_is_r, down, right)
    tails = np.repeat(u[:, None], 2, axis=1)
    keep = heads >= 0
    return np.stack((tails[keep], heads[keep]), axis=1)


def meshgrid(nx: int, length: float = 10.0, ny: int | None = None) -> MeshgridArrays:
    """nx x ny quads over [0, length]^2, z = 0."""
    ny = nx if ny is None else ny
    w = nx + 1
    xs = np.linspace(0.0, length, nx + 1)
    ys = np.linspace(0.0, length, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")          # row-major: y outer, x inner
    xyz = np.stack((X.ravel(), Y.ravel(), np.zeros(X.size)), axis=1)
    a = (np.arange(ny, dtype=np.int64)[:, None] * w + np.arange(nx, dtype=np.int64)[None, :]).ravel()
    faces = np.stack((a, a + 1, a + w + 1, a + w), axis=1)
    edges = meshgrid_edges(nx, ny)
    nodes = np.arange(xyz.shape[0], dtype=np.int64)
    return MeshgridArrays(nodes, xyz, edges, faces, nx, ny)


class Problem(NamedTuple):
    nodes: np.ndarray
    edges: np.ndarray
    supports: np.ndarray      # (N,) int64 0/1
    q: np.ndarray             # (E,) float64
    xyz_fixed: np.ndarray     # (Ns,3) rows in ascending node order (structures.py:163-172)
    loads: np.ndarray         # (N,3)
    xyz0: np.ndarray          # (N,3) initial geometry


def _corner_ids(nx, ny):
    w = nx + 1
    return np.array([0, nx, ny * w, ny * w + nx], dtype=np.int64)


def cablenet_problem(nx: int = 200, seed: int = 0, length: float = 10.0,
                     pz: float = -0.01) -> Problem:
    """
    BASELINE config 4: four degree-2 corners fixed, first and last
    lifted to z = 5 (README.md:88-92), q ~ U[0.5, 2.0] (tension => K SPD), uniform
    nodal load pz.
    """
    m = meshgrid(nx, length)
    N = m.nodes.size
    supports = np.zeros(N, dtype=np.int64)
    corners = _corner_ids(nx, nx)
    supports[corners] = 1
    xyz = m.xyz.copy()
    xyz[corners[0], 2] = 5.0
    xyz[corners[-1], 2] = 5.0
    rng = np.random.default_rng(seed)
    q = rng.uniform(0.5, 2.0, size=m.edges.shape[0])
    loads = np.zeros((N, 3))
    loads[:, 2] = pz
    return Problem(m.nodes, m.edges, supports, q, xyz[np.flatnonzero(supports)], loads, xyz)


def cablenet_readme_problem(nx: int = 10, length: float = 10.0) -> Problem:
    """
    README.md:85-101: q = force / rest_length with force 20 on boundary edges and 1
    inside; corners fixed, two opposite ones lifted by 5; no loads.
    """
    m = meshgrid(nx, length)
    N = m.nodes.size
    w = nx + 1
    supports = np.zeros(N, dtype=np.int64)
    corners = _corner_ids(nx, nx)
    supports[corners] = 1
    xyz = m.xyz.copy()
    xyz[corners[0], 2] = 5.0
    xyz[corners[-1], 2] = 5.0
    i, j = np.divmod(m.edges, w)
    on_bnd = ((i[:, 0] == i[:, 1]) & ((i[:, 0] == 0) | (i[:, 0] == nx))) | \
             ((j[:, 0] == j[:, 1]) & ((j[:, 0] == 0) | (j[:, 0] == nx)))
    rest = np.linalg.norm(xyz[m.edges[:, 1]] - xyz[m.edges[:, 0]], axis=1)
    q = np.where(on_bnd, 20.0, 1.0) / rest
    loads = np.zeros((N, 3))
    return Problem(m.nodes, m.edges, supports, q, xyz[np.flatnonzero(supports)], loads, xyz)


def pillow_problem(nx: int = 30, length: float = 10.0, q0: float = -2.0,
                   pz_total: float = -100.0) -> Problem:
    """
    BASELINE config 5 base design (examples/pillow/pillow.py:42-43,
    90-98 simplified): every boundary vertex fixed, uniform compression q0, uniform
    nodal load pz_total / N.
    """
    m = meshgrid(nx, length)
    N = m.nodes.size
    w = nx + 1
    i, j = np.divmod(m.nodes, w)
    supports = ((i == 0) | (i == nx) | (j == 0) | (j == nx)).astype(np.int64)
    q = np.full(m.edges.shape[0], q0)
    loads = np.zeros((N, 3))
    loads[:, 2] = pz_total / N
    return Problem(m.nodes, m.edges, supports, q, m.xyz[np.flatnonzero(supports)], loads, m.xyz)


def pillow_batch_q(num_edges: int, start: int, count: int, q0: float = -2.0,
                   amp: float = 0.1) -> np.ndarray:
    """
    Designs start..start+count-1 of config 5: q_b = q0 + amp * (U[0,1) - 0.5) with
    default_rng(seed=b), one independent stream per design so any shard of the batch
    can be generated locally.
    """
    out = np.empty((count, num_edges))
    for k in range(count):
        out[k] = q0 + amp * (np.random.default_rng(start + k).random(num_edges) - 0.5)
    return out


def arch_problem(num_segments: int = 10, span: float = 5.0, q0: float = -1.0,
                 pz: float = -0.2) -> Problem:
    """
    The reference's arch fixture (tests/conftest.py:50-68): a polyline from
    (-span/2,0,0) to (span/2,0,0) in num_segments pieces, both ends anchored,
    q = q0 everywhere, load (0,0,pz) on every free node.
    """
    n = num_segments + 1
    nodes = np.arange(n, dtype=np.int64)
    xyz = np.zeros((n, 3))
    xyz[:, 0] = -span / 2 + span * np.arange(n) / num_segments
    edges = np.stack((nodes[:-1], nodes[1:]), axis=1)
    supports = np.zeros(n, dtype=np.int64)
    supports[[0, n - 1]] = 1
    q = np.full(n - 1, q0)
    loads = np.zeros((n, 3))
    loads[supports == 0, 2] = pz
    return Problem(nodes, edges, supports, q, xyz[np.flatnonzero(supports)], loads, xyz)


def arch_benchmark_problem(num_vertices: int, span: float = 10.0, rho: float = 1.0, q0: float = -10.0) -> Problem:
    """
    BASELINE config 2, the analytical arch benchmark (tests/test_arch_benchmark.py:73-97): `num_vertices`
    vertices evenly spaced on [-span/2, span/2] along x, the two ends anchored, every edge at q0, the uniform
    load rho lumped into equal point loads (0, -rho span / n_v, 0) on EVERY vertex (supports included).
    """
    n = int(num_vertices)
    nodes = np.arange(n, dtype=np.int64)
    xyz = np.zeros((n, 3))
    xyz[:, 0] = -span / 2.0 + span * np.arange(n) / (n - 1)
    edges = np.stack((nodes[:-1], nodes[1:]), axis=1)
    supports = np.zeros(n, dtype=np.int64)
    supports[[0, n - 1]] = 1
    q = np.full(n - 1, q0)
    loads = np.zeros((n, 3))
    loads[:, 1] = (-rho * span) / n
    return Problem(nodes, edges, supports, q, xyz[np.flatnonzero(supports)], loads, xyz)
