"""
Host-side mirror of `jax_fdm.equilibrium.structures` for the equilibrium hot path.

Same class names and attribute names as the reference
(structures/graphs.py:27-168, structures/structures.py:32-350); arrays are numpy
(static topology, never differentiated).  All index arrays come from the native
plan builder (csrc/plan.cpp) through the C ABI, not from scipy: the plan is the
single source of truth shared with the kernels, and the tests compare it
bit-for-bit with what the reference's own code produced (tests/golden/).
"""

import ctypes as C
from typing import NamedTuple

import numpy as np

from .. import _lib as L

__all__ = [
    "Graph",
    "GraphSparse",
    "EquilibriumStructure",
    "EquilibriumStructureSparse",
    "EquilibriumMeshStructure",
    "EquilibriumMeshStructureSparse",
    "IndexArray",
]


class IndexArray(NamedTuple):
    """The labelled CSC `index_array` of structures.py:248-288 (data/indices/indptr/shape)."""

    data: np.ndarray      # (nnz,)  int64, e+1 off-diagonal, 0 on the diagonal
    indices: np.ndarray   # (nnz,)  int32
    indptr: np.ndarray    # (Nf+1,) int32
    shape: tuple


class Graph:
    """
    Nodes and edges as static index arrays (structures/graphs.py:27-116).

    `nodes` are integer keys, `edges` pairs of keys (tail, head).  `edges_indexed`
    rewrites keys to positions in `nodes` (graphs.py:86-97).
    """

    def __init__(self, nodes, edges, **kwargs):
        self.nodes = np.asarray(nodes, dtype=np.int64)
        edges = np.asarray(edges, dtype=np.int64)
        assert edges.ndim == 2 and edges.shape[1] == 2, "Edges in graph must connect exactly 2 nodes"
        self.edges = edges
        self.edges_indexed = self._edges_indexed()

    @property
    def num_nodes(self):
        return int(self.nodes.size)

    @property
    def num_edges(self):
        return int(self.edges.shape[0])

    @property
    def node_index(self):
        return {int(n): i for i, n in enumerate(self.nodes)}

    @property
    def edge_index(self):
        return {(int(u), int(v)): i for i, (u, v) in enumerate(self.edges)}

    def _edges_indexed(self):
        # vectorised equivalent of the per-edge dictionary lookup in graphs.py:86-97
        if np.unique(self.nodes).size != self.nodes.size:
            raise ValueError("duplicate node keys")
        order = np.argsort(self.nodes, kind="stable")
        pos = np.searchsorted(self.nodes[order], self.edges.ravel())
        pos = np.clip(pos, 0, self.nodes.size - 1)
        idx = order[pos]
        if not np.array_equal(self.nodes[idx], self.edges.ravel()):
            raise KeyError("edge references a node key that is not in `nodes`")
        return idx.reshape(-1, 2).astype(np.int64)


class GraphSparse(Graph):
    """Sparse flavour (graphs.py:124-168): connectivity is kept implicit (edge lists)."""

    @property
    def connectivity_scipy(self):
        """Signed incidence matrix as scipy CSC (graphs.py:176-211); built on demand, host only."""
        import scipy.sparse as sp

        m = self.num_edges
        ei = self.edges_indexed
        data = np.concatenate((-np.ones(m), np.ones(m)))
        rows = np.concatenate((np.arange(m), np.arange(m)))
        cols = np.concatenate((ei[:, 0], ei[:, 1]))
        return sp.coo_matrix((data, (rows, cols)), shape=(m, self.num_nodes)).tocsc()


class EquilibriumStructure(Graph):
    """
    Graph + supports + free/fixed index maps (structures.py:32-188), backed by a native plan.

    Parameters mirror the reference: `nodes` (N,), `edges` (E,2) keys, `supports` (N,) 0/1.
    `device`: CUDA device index for the plan; `None` builds a host-only plan (index maps
    available, compute calls refused) -- used by CPU tests.
    """

    def __init__(self, nodes, edges, supports, device=0, **kwargs):
        super().__init__(nodes=nodes, edges=edges, **kwargs)
        self.supports = np.asarray(supports, dtype=np.int64)
        assert self.supports.shape == self.nodes.shape
        self.device = device
        self._plan = C.c_void_p()
        lib = L.lib()
        ei = np.ascontiguousarray(self.edges_indexed, dtype=np.int64)
        sup = np.ascontiguousarray(self.supports, dtype=np.int64)
        if device is None:
            rc = lib.fdm_plan_create_host(ei.ctypes.data, ei.shape[0], self.nodes.size,
                                          sup.ctypes.data, C.byref(self._plan))
        else:
            rc = lib.fdm_plan_create(ei.ctypes.data, ei.shape[0], self.nodes.size,
                                     sup.ctypes.data, int(device), C.byref(self._plan))
        L.check(rc)
        self.indices_free = self._get(L.ARR_INDICES_FREE, np.int64, self.num_free)
        self.indices_fixed = self._get(L.ARR_INDICES_FIXED, np.int64, self.num_supports)
        self.indices_freefixed = self._get(L.ARR_INDICES_FREEFIXED, np.int64, self.num_nodes)

    # -- plan plumbing ---------------------------------------------------------------
    @property
    def plan(self):
        return self._plan

    def _size(self, which):
        return int(L.lib().fdm_plan_size(self._plan, which))

    def _get(self, which, dtype, count):
        out = np.empty(count, dtype=dtype)
        n = L.lib().fdm_plan_get_array(self._plan, which, out.ctypes.data, out.nbytes)
        if n < 0:
            L.check(int(n))
        assert n == out.nbytes, (which, n, out.nbytes)
        return out

    def __del__(self):
        try:
            if getattr(self, "_plan", None) and self._plan.value:
                L.lib().fdm_plan_destroy(self._plan)
                self._plan = C.c_void_p()
        except Exception:
            pass

    # -- reference attribute surface ------------------------------------------------------
    @classmethod
    def from_arrays(cls, nodes, edges, supports, **kw):
        return cls(nodes, edges, supports, **kw)

    @property
    def num_supports(self):
        return self._size(L.SIZE_FIXED)

    @property
    def num_free(self):
        return self._size(L.SIZE_FREE)

    @property
    def support_index(self):
        return {int(k): i for i, k in enumerate(self.nodes_fixed)}

    @property
    def nodes_free(self):
        return self.nodes[self.indices_free]

    @property
    def nodes_fixed(self):
        return self.nodes[self.indices_fixed]


class EquilibriumStructureSparse(EquilibriumStructure, GraphSparse):
    """
    Adds the stiffness-assembly helpers of structures.py:196-350: `index_array`,
    `diag_indices`, `diags` (here: CSR arrays `diags_indptr`, `diags_indices`, all ones).
    """

    def __init__(self, nodes, edges, supports, device=0, **kwargs):
        super().__init__(nodes=nodes, edges=edges, supports=supports, device=device, **kwargs)
        nnz, nf = self._size(L.SIZE_NNZ), self.num_free
        self.index_array = IndexArray(
            data=self._get(L.ARR_INDEX_DATA, np.int64, nnz),
            indices=self._get(L.ARR_INDEX_INDICES, np.int32, nnz),
            indptr=self._get(L.ARR_INDEX_INDPTR, np.int32, nf + 1),
            shape=(nf, nf),
        )
        self.diag_indices = self._get(L.ARR_DIAG_INDICES, np.int64, nf)
        self.diags_indptr = self._get(L.ARR_DIAGS_INDPTR, np.int32, nf + 1)
        self.diags_indices = self._get(L.ARR_DIAGS_INDICES, np.int32, self._size(L.SIZE_DIAGS_NNZ))

    @property
    def nnz(self):
        return self._size(L.SIZE_NNZ)

    @property
    def half_bandwidth(self):
        return self._size(L.SIZE_HBW)

    @property
    def diags(self):
        """`diags` as scipy CSR (Nf x E) of ones (structures.py:316-350); host only."""
        import scipy.sparse as sp

        return sp.csr_matrix((np.ones(self.diags_indices.size), self.diags_indices, self.diags_indptr),
                             shape=(self.num_free, self.num_edges))

    @property
    def band_perm(self):
        return self._get(L.ARR_BAND_PERM, np.int32, self.num_free)

    @property
    def part_of_row(self):
        return self._get(L.ARR_PART_OF_ROW, np.int32, self.num_free)


class EquilibriumMeshStructureSparse(EquilibriumStructureSparse):
    """
    Mesh flavour (structures/structures.py:358-505, structures/meshes.py:29-293): the graph structure plus
    faces, for tributary area loads.  `faces` is (F, maxdeg) of vertex KEYS padded with -1.

    Attributes named as in the reference: `vertices`, `faces`, `faces_indexed`, `edges_faces_indexed`
    (first two incident faces of every edge over both windings, padded -1, meshes.py:146-204),
    `connectivity_faces_vertices` (row-normalised face-vertex matrix, meshes.py:361-399), `face_keys`,
    `num_faces`.  `faces_edges` and `node_faces` are the transposed adjacencies the reverse kernels need.
    """

    def __init__(self, vertices, faces, edges, supports, face_keys=None, device=0, **kwargs):
        super().__init__(nodes=vertices, edges=edges, supports=supports, device=device, **kwargs)
        self.vertices = self.nodes
        self.faces = np.asarray(faces, dtype=np.int64)
        assert self.faces.ndim == 2 and self.faces.shape[1] > 2, "The mesh faces must have at least 3 vertices each"
        self.face_keys = np.arange(self.faces.shape[0]) if face_keys is None else np.asarray(face_keys, dtype=np.int64)
        vi = self.node_index
        self.faces_indexed = np.array([[vi[int(v)] if v >= 0 else int(v) for v in face] for face in self.faces],
                                      dtype=np.int64)
        self.edges_faces_indexed = self._edges_faces_indexed()
        self.faces_edges, self.node_faces_ptr, self.node_faces = self._transposed_adjacency()
        self._dev = None

    @property
    def num_faces(self):
        return int(self.faces.shape[0])

    @property
    def vertex_index(self):
        return self.node_index

    @property
    def connectivity_faces_vertices(self):
        import scipy.sparse as sp

        mask = self.faces_indexed >= 0
        counts = mask.sum(axis=1)
        rows = np.repeat(np.arange(self.num_faces), counts)
        return sp.coo_matrix((np.repeat(1.0 / counts, counts), (rows, self.faces_indexed[mask])),
                             shape=(self.num_faces, self.num_nodes)).tocsc()

    def _edges_faces_indexed(self):
        half = {}
        for f, face in enumerate(self.faces_indexed):
            clean = [int(v) for v in face if v >= 0]
            for u, v in zip(clean, clean[1:] + clean[:1]):
                half.setdefault((u, v), []).append(f)
        out = np.full((self.num_edges, 2), -1, dtype=np.int64)
        for e, (u, v) in enumerate(self.edges_indexed):
            fs = half.get((int(u), int(v)), []) + half.get((int(v), int(u)), [])
            out[e, :len(fs[:2])] = fs[:2]
        return out

    def _transposed_adjacency(self):
        eidx = {}
        for e, (u, v) in enumerate(self.edges_indexed):
            eidx[(int(u), int(v))] = e
            eidx[(int(v), int(u))] = e
        F, D = self.faces_indexed.shape
        fe = np.full((F, D), -1, dtype=np.int32)
        nf = [[] for _ in range(self.num_nodes)]
        for f, face in enumerate(self.faces_indexed):
            clean = [int(v) for v in face if v >= 0]
            for i, (u, v) in enumerate(zip(clean, clean[1:] + clean[:1])):
                e = eidx.get((u, v), -1)
                if e >= 0 and f in self.edges_faces_indexed[e]:
                    fe[f, i] = e
            for v in clean:
                nf[v].append(f)
        ptr = np.zeros(self.num_nodes + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(x) for x in nf])
        flat = np.array([f for x in nf for f in x], dtype=np.int32)
        return fe, ptr, flat

    def device_arrays(self, device):
        """int32 device copies of the mesh adjacency (built once)."""
        import torch

        if self._dev is None or self._dev["faces"].device != device:
            t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=device)
            self._dev = {"faces": t(self.faces_indexed), "edges_faces": t(self.edges_faces_indexed),
                         "faces_edges": t(self.faces_edges), "node_faces_ptr": t(self.node_faces_ptr),
                         "node_faces": t(self.node_faces if self.node_faces.size else np.zeros(1))}
        return self._dev


EquilibriumMeshStructure = EquilibriumMeshStructureSparse
