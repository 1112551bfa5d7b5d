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
    "Mesh",
    "MeshSparse",
    "EquilibriumStructure",
    "EquilibriumStructureSparse",
    "EquilibriumMeshStructure",
    "EquilibriumMeshStructureSparse",
    "IndexArray",
    "adjacency_matrix",
    "connectivity_matrix",
    "face_matrix",
    "mesh_edges_faces",
    "mesh_connectivity_edges_faces",
]


# ==========================================================================
# Helper functions (structures/graphs.py:176-250, structures/meshes.py:301-399)
# ==========================================================================
def connectivity_matrix(edges, num_nodes=None):
    """Signed edge-node incidence matrix as scipy CSC (graphs.py:176-211): -1 at the tail, +1 at the head."""
    import scipy.sparse as sp

    ei = np.asarray(edges)
    m = len(ei)
    data = np.concatenate((-np.ones(m), np.ones(m)))
    rows = np.concatenate((np.arange(m), np.arange(m)))
    cols = np.concatenate((ei[:, 0], ei[:, 1]))
    n = num_nodes if num_nodes is not None else int(np.max(ei)) + 1
    return sp.coo_matrix((data, (rows, cols)), shape=(m, n)).tocsc()


def adjacency_matrix(edges, num_nodes=None):
    """Symmetric node-node adjacency as scipy CSC (graphs.py:214-250): a one wherever two nodes share an edge."""
    import scipy.sparse as sp

    ei = np.asarray(edges)
    n = num_nodes if num_nodes is not None else int(np.max(np.ravel(ei))) + 1
    rows = np.concatenate((ei[:, 0], ei[:, 1]))
    cols = np.concatenate((ei[:, 1], ei[:, 0]))
    return sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n)).tocsc()


def face_matrix(face_vertices, normalize=True):
    """Face-vertex incidence as scipy CSC, -1 padding ignored (meshes.py:361-399); rows sum to one when normalised."""
    import scipy.sparse as sp

    faces = np.asarray(face_vertices)
    mask = faces >= 0
    counts = mask.sum(axis=1)
    rows = np.repeat(np.arange(faces.shape[0]), counts)
    data = np.repeat(1.0 / counts, counts) if normalize else np.ones(rows.size)
    return sp.coo_matrix((data, (rows, faces[mask]))).tocsc()


def mesh_edges_faces(mesh):
    """Face indices incident to each edge of a COMPAS-style mesh (meshes.py:301-331): anything with `faces()`,
    `edges()` and `edge_faces((u, v))`."""
    face_index = {fkey: idx for idx, fkey in enumerate(mesh.faces())}
    out = []
    for u, v in mesh.edges():
        found = tuple(face_index[f] for f in mesh.edge_faces((u, v)) if f is not None)
        assert len(found) <= 2
        out.append(found)
    return out


def mesh_connectivity_edges_faces(mesh):
    """Dense edge-face incidence of a COMPAS-style mesh (meshes.py:334-358)."""
    ef = mesh_edges_faces(mesh)
    C_ = np.zeros((len(ef), mesh.number_of_faces()))
    for e, fs in enumerate(ef):
        C_[e, list(fs)] = 1.0
    return C_


class _TaggedArray(np.ndarray):
    """A dense connectivity matrix that remembers its structure (see `_tag`)."""
    _fdm_structure = None


def _tag(matrix, structure):
    """`structure.connectivity` is what the reference's model methods take (models.py:99-193: `connectivity @ xyz`);
    the kernels need the structure's plan instead, so the matrix carries a weak reference to its owner."""
    import weakref

    if isinstance(matrix, np.ndarray):
        matrix = matrix.view(_TaggedArray)
    matrix._fdm_structure = weakref.ref(structure)
    return matrix


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

    # The reference stores these two as arrays built in __init__ (graphs.py:54-55, 99-116); here they are built on
    # first access and cached: the kernels never read them (connectivity is implicit in the plan's edge lists), and
    # the DENSE flavour is E x N doubles -- 26 GB at 200x200 -- so nothing pays for it unless it asks.
    @property
    def connectivity(self):
        if getattr(self, "_connectivity", None) is None:
            self._connectivity = _tag(self._connectivity_matrix(), self)
        return self._connectivity

    @property
    def adjacency(self):
        if getattr(self, "_adjacency", None) is None:
            self._adjacency = self._adjacency_matrix()
        return self._adjacency

    def _connectivity_matrix(self):
        return connectivity_matrix(self.edges_indexed, self.num_nodes).toarray()

    def _adjacency_matrix(self):
        return adjacency_matrix(self.edges_indexed, self.num_nodes).toarray()

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
        """Signed incidence matrix as scipy CSC (graphs.py:140-158, 176-211); rebuilt on every access like the
        reference's, host only."""
        return connectivity_matrix(self.edges_indexed, self.num_nodes)

    def _connectivity_matrix(self):
        """Sparse flavour (graphs.py:134-138): scipy CSC where the reference holds a BCOO of the same matrix."""
        return self.connectivity_scipy

    def _adjacency_matrix(self):
        return adjacency_matrix(self.edges_indexed, self.num_nodes)


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
            self.__dict__.pop("_fp_pool", None)            # loop objects built on this plan go first
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
    def connectivity_free(self):
        """C[:, indices_free] (structures.py:46, 190-194 dense; :236-240 sparse), built on first access."""
        if getattr(self, "_connectivity_free", None) is None:
            self._connectivity_free = self.connectivity[:, self.indices_free]
        return self._connectivity_free

    @property
    def connectivity_fixed(self):
        """C[:, indices_fixed] (structures.py:47, 242-246)."""
        if getattr(self, "_connectivity_fixed", None) is None:
            self._connectivity_fixed = self.connectivity[:, self.indices_fixed]
        return self._connectivity_fixed

    @classmethod
    def from_network(cls, network, **kw):
        """structures.py:71-112 for any COMPAS-style network: `nodes()`, `edges()`, `is_node_support(node)`."""
        nodes = np.array(list(network.nodes()), dtype=np.int64)
        edges = np.array([tuple(e) for e in network.edges()], dtype=np.int64)
        supports = np.array([1 if network.is_node_support(n) else 0 for n in nodes], dtype=np.int64)
        return cls(nodes, edges, supports, **kw)

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


class Mesh(Graph):
    """
    A graph with faces (structures/meshes.py:29-259).  `faces` is (F, maxdeg) of vertex KEYS padded with -1;
    edges are derived from the faces when not given.  Attributes as in the reference: `vertices`, `faces`,
    `faces_indexed`, `edges_faces_indexed` (first two incident faces of every edge over both windings, padded
    -1, meshes.py:146-204), `connectivity_faces_vertices` (row-normalised face-vertex matrix, the centroid
    operator, meshes.py:247-259), `connectivity_edges_faces` (meshes.py:135-144), `face_keys`.
    """

    def __init__(self, vertices, faces, edges=None, face_keys=None, **kwargs):
        faces = np.asarray(faces, dtype=np.int64)
        assert faces.ndim == 2 and faces.shape[1] > 2, "Mesh faces must connect at least 3 vertices"
        if edges is None:
            edges = self._edges_from_faces(faces)
        kwargs.pop("nodes", None)
        self.vertices = np.asarray(vertices, dtype=np.int64)
        self.faces = faces
        self.face_keys = np.arange(len(faces)) if face_keys is None else np.asarray(face_keys, dtype=np.int64)
        super().__init__(nodes=self.vertices, edges=edges, **kwargs)
        vi = self.node_index
        self.faces_indexed = np.array([[vi[int(v)] if v >= 0 else int(v) for v in face] for face in self.faces],
                                      dtype=np.int64)
        self.edges_faces_indexed = self._edges_faces_indexed()

    @property
    def num_vertices(self):
        return int(self.vertices.size)

    @property
    def num_faces(self):
        return int(self.faces.shape[0])

    @property
    def vertex_index(self):
        return self.node_index

    @property
    def face_index(self):
        return {int(f): i for i, f in enumerate(self.face_keys)}

    @property
    def connectivity_faces_vertices(self):
        return face_matrix(self.faces_indexed, normalize=True).toarray()

    @property
    def connectivity_edges_faces(self):
        return self._connectivity_edges_faces_scipy().toarray()

    def _connectivity_edges_faces_scipy(self):
        import scipy.sparse as sp

        ef = self.edges_faces_indexed
        mask = ef >= 0
        rows = np.repeat(np.arange(self.num_edges), mask.sum(axis=1))
        return sp.coo_matrix((np.ones(rows.size), (rows, ef[mask])), shape=(self.num_edges, self.num_faces)).tocsc()

    def _edges_faces(self):
        """Face indices incident to each edge over both windings (meshes.py:162-204)."""
        half = {}
        for f, face in enumerate(self.faces_indexed):
            clean = [int(v) for v in face if v >= 0]
            for u, v in zip(clean, clean[1:] + clean[:1]):
                half.setdefault((u, v), []).append(f)
        out = []
        for u, v in self.edges_indexed:
            fs = half.get((int(u), int(v)), []) + half.get((int(v), int(u)), [])
            if len(fs) > 2:
                print(f"Warning: Edge {(int(u), int(v))} is non-manifold, it's shared by ({len(fs)}) faces. "
                      "This might lead to unexpected behavior in e.g. in area load calculations.")
            out.append(tuple(fs))
        return tuple(out)

    def _edges_faces_indexed(self):
        out = np.full((self.num_edges, 2), -1, dtype=np.int64)
        for e, fs in enumerate(self._edges_faces()):
            out[e, :len(fs[:2])] = fs[:2]
        return out

    @staticmethod
    def _edges_from_faces(faces):
        """Unique undirected edges in order of first appearance around the faces (meshes.py:207-245)."""
        edges, seen = [], set()
        for face in np.asarray(faces):
            loop = [int(v) for v in face] + [int(face[0])]
            for u, v in zip(loop, loop[1:]):
                if (u, v) in seen or (v, u) in seen:
                    continue
                seen.add((u, v))
                edges.append((u, v))
        return np.asarray(edges, dtype=np.int64)


class MeshSparse(Mesh, GraphSparse):
    """Sparse flavour (meshes.py:261-293): the two incidence matrices as scipy CSC (the reference holds BCOO)."""

    @property
    def connectivity_faces_vertices(self):
        return face_matrix(self.faces_indexed, normalize=True)

    @property
    def connectivity_edges_faces(self):
        return self._connectivity_edges_faces_scipy()


class _MeshStructureMixin:
    """What both mesh structures add to their graph structure (structures.py:358-505): the classmethod
    `from_mesh`, the vertex aliases, and the transposed adjacencies the reverse kernels of the face loads read."""

    def _init_mesh_tables(self):
        self.faces_edges, self.node_faces_ptr, self.node_faces = self._transposed_adjacency()
        self._dev = None

    @classmethod
    def from_mesh(cls, mesh, **kw):
        """structures.py:394-452 for any COMPAS-style mesh: `vertices()`, `edges()`, `faces()`,
        `face_vertices(fkey)`, `is_vertex_support(vertex)`; faces padded with -1 to a common length."""
        vertices = list(mesh.vertices())
        edges = [tuple(e) for e in mesh.edges()]
        supports = [1 if mesh.is_vertex_support(v) else 0 for v in vertices]
        face_keys = np.asarray(list(mesh.faces()), dtype=np.int64)
        faces = [list(mesh.face_vertices(f)) for f in face_keys]
        deg = max(len(f) for f in faces)
        assert deg > 2, "The mesh faces must have at least 3 vertices each"
        faces = np.asarray([f + [-1] * (deg - len(f)) for f in faces], dtype=np.int64)
        return cls(np.asarray(vertices, dtype=np.int64), faces, edges=np.asarray(edges, dtype=np.int64),
                   supports=np.asarray(supports, dtype=np.int64), face_keys=face_keys, **kw)

    @property
    def vertices_free(self):
        return self.vertices[self.indices_free]

    @property
    def vertices_fixed(self):
        return self.vertices[self.indices_fixed]

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


class EquilibriumMeshStructure(_MeshStructureMixin, EquilibriumStructure, Mesh):
    """Mesh flavour of the structure (structures.py:358-476): vertices play the role of nodes, faces carry the
    tributary area loads."""

    def __init__(self, vertices, faces, edges, supports, face_keys=None, device=0, **kwargs):
        super().__init__(nodes=vertices, edges=edges, supports=supports, device=device, vertices=vertices, faces=faces,
                         face_keys=face_keys, **kwargs)
        self._init_mesh_tables()


class EquilibriumMeshStructureSparse(_MeshStructureMixin, EquilibriumStructureSparse, MeshSparse):
    """Sparse mesh structure (structures.py:479-505): EquilibriumStructureSparse + MeshSparse."""

    def __init__(self, vertices, faces, edges, supports, face_keys=None, device=0, **kwargs):
        super().__init__(nodes=vertices, edges=edges, supports=supports, device=device, vertices=vertices, faces=faces,
                         face_keys=face_keys, **kwargs)
        self._init_mesh_tables()
