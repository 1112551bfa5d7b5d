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
