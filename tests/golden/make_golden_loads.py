#!/usr/bin/env python
"""
Golden vectors for the shape-dependent loads (global and local / follower frames), made by
EXECUTING THE REFERENCE'S OWN CODE:

    src/jax_fdm/geometry/geometry.py     (line_lcs, polygon_lcs, normal_polygon, area_triangle, ...)
    src/jax_fdm/equilibrium/loads.py     (nodes_load_from_edges, nodes_load_from_faces)

loaded unmodified from /root/reference under the numpy stand-in for `jax` of make_golden.py, plus
a loop-based `vmap`.  The mesh index arrays (`faces_indexed`, `edges_faces_indexed`,
`connectivity_faces_vertices`) fed to loads.py here come from this repo's host mirror of EquilibriumMeshStructure;
make_golden_mesh.py runs the reference's real mesh structure class and pins those tables (and the face-load
fixed point) separately.

Run here (needs /root/reference):   python tests/golden/make_golden_loads.py
Output: tests/golden/loads_frames.npz
"""

import importlib
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402

G.REAL |= {"jax_fdm.geometry", "jax_fdm.geometry.geometry", "jax_fdm.equilibrium.loads"}


def vmap(fn, in_axes=0):
    """jax.vmap over leading axes, as a Python loop (outputs stacked)."""
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.asarray(a).shape[0] for a, ax in zip(args, axes) if ax is not None)
        rows = [fn(*[a if ax is None else np.asarray(a)[i].view(G.JArr) for a, ax in zip(args, axes)]) for i in range(n)]
        return G.jarr(np.stack([np.asarray(r) for r in rows]))
    return mapped


_build_fake = G._build_fake


def build_fake(name):
    m = _build_fake(name)
    if name == "jax":
        m.vmap = vmap
    elif name == "jax.numpy":
        m.zeros = lambda shape, dtype=None: G.jarr(np.zeros(shape))
        m.eye = lambda n: G.jarr(np.eye(n))
        m.vstack = lambda xs: G.jarr(np.vstack([np.asarray(x) for x in xs]))
        m.where = lambda c, a, b: G.jarr(np.where(np.asarray(c), np.asarray(a), np.asarray(b)))
        m.cross = lambda a, b: G.jarr(np.cross(np.asarray(a), np.asarray(b)))
    return m


G._build_fake = build_fake


def mesh_case(nx, seed, pad):
    from jax_fdm_b200 import datasets
    from jax_fdm_b200.equilibrium.structures import EquilibriumMeshStructureSparse

    m = datasets.meshgrid(nx)
    prob = datasets.cablenet_problem(nx, seed=seed)
    faces, edges = m.faces, prob.edges
    if pad:     # a mixed mesh: split the first quad into two triangles (new diagonal edge), pad the face table with -1
        a, b, c, d = faces[0]
        faces = np.concatenate([np.array([[a, b, c, -1], [a, c, d, -1]], dtype=np.int64), faces[1:]])
        edges = np.concatenate([edges, np.array([[a, c]], dtype=np.int64)])
    s = EquilibriumMeshStructureSparse(prob.nodes, faces, edges, prob.supports, device=None)
    rng = np.random.default_rng(seed)
    xyz = prob.xyz0.copy()
    xyz[:, 2] += 0.4 * np.sin(xyz[:, 0]) * np.cos(0.7 * xyz[:, 1]) + 0.05 * rng.standard_normal(len(xyz))
    E, F = s.num_edges, s.num_faces
    eload = rng.standard_normal((E, 3)) * 0.1
    fload = rng.standard_normal((F, 3)) * 0.1
    return s, xyz, eload, fload


def main():
    G.load_reference()
    geom = importlib.import_module("jax_fdm.geometry")
    loads = importlib.import_module("jax_fdm.equilibrium.loads")
    out = {}
    rng = np.random.default_rng(11)

    lines = rng.standard_normal((24, 2, 3))
    lines[0] = [[0, 0, 0], [1, 0, 0]]
    lines[1] = [[0, 0, 0], [0, 1, 0]]
    lines[2] = [[0, 0, 0], [0, 0, 1]]
    lines[3] = [[1, 2, 3], [1, 2, -4]]                       # along -Z
    lines[4] = [[0, 0, 0], [1e-4, 0, 1]]                     # within the allclose band of Z
    lines[5] = [[0, 0, 0], [1e-2, 0, 1]]                     # just outside it
    out["lines"] = lines
    out["ref_line_lcs"] = np.stack([np.asarray(geom.line_lcs(G.jarr(l))) for l in lines])

    for deg in (3, 4, 6):
        polys = rng.standard_normal((12, deg, 3))
        t = np.linspace(0, 2 * np.pi, deg, endpoint=False)
        polys[0] = np.c_[np.cos(t), np.sin(t), 0 * t]         # XY, counter-clockwise
        polys[1] = np.c_[0 * t, np.cos(t), np.sin(t)]         # YZ: normal along X
        polys[2] = np.c_[np.sin(t), 0 * t, np.cos(t)]         # ZX
        polys[3] = polys[0][::-1] + 100.0                     # clockwise, far from the origin
        polys[4] = np.c_[t, 2 * t, -t]                        # collinear: zero normal
        out[f"polys{deg}"] = polys
        out[f"ref_normal{deg}"] = np.stack([np.asarray(geom.normal_polygon(G.jarr(p))) for p in polys])
        out[f"ref_polygon_lcs{deg}"] = np.stack([np.asarray(geom.polygon_lcs(G.jarr(p))) for p in polys])

    for tag, pad in (("quad", False), ("mixed", True)):
        s, xyz, eload, fload = mesh_case(5, 9, pad)
        F, D = s.faces_indexed.shape
        mask = s.faces_indexed >= 0
        rows = np.repeat(np.arange(F), mask.sum(1))
        cfv = sp.coo_matrix((np.ones(rows.size), (rows, s.faces_indexed[mask])), shape=(F, s.num_nodes)).tocsr()
        cfv = sp.diags(1.0 / mask.sum(1)) @ cfv
        ei = s.edges_indexed
        E = len(ei)
        conn = sp.coo_matrix((np.r_[-np.ones(E), np.ones(E)], (np.r_[np.arange(E), np.arange(E)], np.r_[ei[:, 0], ei[:, 1]])),
                             shape=(E, s.num_nodes)).tocsr()
        ns = types.SimpleNamespace(edges_indexed=G.jarr(ei), faces_indexed=G.jarr(s.faces_indexed),
                                   edges_faces_indexed=G.jarr(s.edges_faces_indexed),
                                   connectivity_faces_vertices=G.BCOO(cfv), connectivity=G.BCOO(conn),
                                   num_nodes=s.num_nodes)
        out[f"{tag}_nodes"] = np.asarray(s.nodes if hasattr(s, "nodes") else np.arange(s.num_nodes))
        out[f"{tag}_faces"] = np.asarray(s.faces)
        out[f"{tag}_edges"] = np.asarray(s.edges)
        out[f"{tag}_supports"] = np.asarray(s.supports)
        out[f"{tag}_faces_indexed"] = s.faces_indexed
        out[f"{tag}_edges_faces_indexed"] = s.edges_faces_indexed
        out[f"{tag}_xyz"] = xyz
        out[f"{tag}_edges_load"] = eload
        out[f"{tag}_faces_load"] = fload
        for local in (False, True):
            k = "local" if local else "global"
            out[f"{tag}_ref_nodes_load_edges_{k}"] = np.asarray(loads.nodes_load_from_edges(G.jarr(xyz), G.jarr(eload), ns, local))
            out[f"{tag}_ref_nodes_load_faces_{k}"] = np.asarray(loads.nodes_load_from_faces(G.jarr(xyz), G.jarr(fload), ns, local))

    path = os.path.join(HERE, "loads_frames.npz")
    np.savez_compressed(path, **out)
    print(f"loads_frames.npz: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
