#!/usr/bin/env python
"""
Golden vectors for mesh structures and face (area) loads, made by EXECUTING THE REFERENCE'S OWN CODE:

    src/jax_fdm/equilibrium/structures/meshes.py, structures.py   (EquilibriumMeshStructureSparse: faces_indexed,
                                                                   edges_faces_indexed, connectivity_faces_vertices)
    src/jax_fdm/equilibrium/models.py, solvers/fixed_point.py, loads.py, geometry/geometry.py
                                                                  (tmax > 1 fixed point under face loads, both frames)

loaded unmodified from /root/reference under the numpy stand-in for `jax` of make_golden.py (COMPAS only appears in
type annotations of meshes.py, so a placeholder serves).  Forward pass only.

Run here (needs /root/reference):   python tests/golden/make_golden_mesh.py
Output: tests/golden/mesh_faces.npz
"""
import importlib
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import make_golden_loads as GL  # noqa: E402,F401
import make_golden_fixed_point as GF  # noqa: E402,F401  (while_loop, lax.cond, the solvers package)

G.REAL |= {"jax_fdm.equilibrium.structures.meshes"}
_prev_late = G._populate_late


def _populate_late(m):
    _prev_late(m)
    if m.__name__ == "jax_fdm.equilibrium.structures":
        s = importlib.import_module("jax_fdm.equilibrium.structures.structures")
        m.EquilibriumMeshStructure = s.EquilibriumMeshStructure
        m.EquilibriumMeshStructureSparse = s.EquilibriumMeshStructureSparse


G._populate_late = _populate_late


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from jax_fdm_b200 import datasets

    structures, states, models = G.load_reference()
    fp = importlib.import_module("jax_fdm.equilibrium.solvers.fixed_point")
    eqx_internal = importlib.import_module("equinox.internal")
    out = {}
    nx = 5
    m = datasets.meshgrid(nx)
    prob = datasets.cablenet_problem(nx, seed=9)
    for tag in ("quad", "mixed"):
        faces, edges = m.faces, prob.edges
        q = prob.q
        if tag == "mixed":      # first quad split into two triangles along a new edge; face table padded with -1
            a, b, c, d = faces[0]
            faces = np.concatenate([np.array([[a, b, c, -1], [a, c, d, -1]], dtype=np.int64), faces[1:]])
            edges = np.concatenate([edges, np.array([[a, c]], dtype=np.int64)])
            q = np.concatenate([q, [q.mean()]])
        s = structures.EquilibriumMeshStructureSparse(prob.nodes, faces, edges, prob.supports)
        rng = np.random.default_rng(3)
        fload = np.zeros((faces.shape[0], 3))
        fload[:, 2] = -0.08 * rng.uniform(0.5, 1.5, size=faces.shape[0])
        fload[:, 1] = 0.01 * rng.standard_normal(faces.shape[0])
        out[f"{tag}_nodes"], out[f"{tag}_faces"], out[f"{tag}_edges"] = prob.nodes, faces, edges
        out[f"{tag}_supports"], out[f"{tag}_q"], out[f"{tag}_xyz_fixed"], out[f"{tag}_loads"] = prob.supports, q, prob.xyz_fixed, prob.loads
        out[f"{tag}_faces_load"] = fload
        out[f"{tag}_ref_faces_indexed"] = np.asarray(s.faces_indexed)
        out[f"{tag}_ref_edges_faces_indexed"] = np.asarray(s.edges_faces_indexed)
        cfv = sp.csr_matrix(s.connectivity_faces_vertices.m)
        cfv.sort_indices()
        out[f"{tag}_ref_cfv_data"], out[f"{tag}_ref_cfv_indices"], out[f"{tag}_ref_cfv_indptr"] = cfv.data, cfv.indices, cfv.indptr
        cef = sp.csr_matrix(s.connectivity_edges_faces.m)
        cef.sort_indices()
        out[f"{tag}_ref_cef_data"], out[f"{tag}_ref_cef_indices"], out[f"{tag}_ref_cef_indptr"] = cef.data, cef.indices, cef.indptr
        qj, xs, loads = G.jarr(q), G.jarr(prob.xyz_fixed), G.jarr(prob.loads)
        for local in (False, True):
            model = models.EquilibriumModelSparse(tmax=100, eta=1e-10, is_load_local=local, itersolve_fn=fp.solver_forward,
                                                  implicit_diff=True, verbose=False)
            params = states.EquilibriumParametersState(qj, xs, states.LoadState(loads, 0.0, G.jarr(fload)))
            st = model(params, s)
            key = f"{tag}_{'local' if local else 'global'}"
            out[f"{key}_steps"] = np.asarray(eqx_internal.while_loop.steps)
            for f in st._fields:
                out[f"{key}_state_{f}"] = np.asarray(getattr(st, f))
            print(key, "steps", out[f"{key}_steps"], "max |z|", np.abs(out[f"{key}_state_xyz"][:, 2]).max())
    path = os.path.join(HERE, "mesh_faces.npz")
    np.savez_compressed(path, **out)
    print(f"mesh_faces.npz: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
