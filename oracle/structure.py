"""
oracle/structure.py -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Topology arrays of the sparse equilibrium structure, restated with the same
scipy.sparse calls the reference makes, so integer outputs are bit-identical.

Follows (reference file:line):
  structures/graphs.py:86-97      edges keyed -> edges indexed
  structures/graphs.py:176-211    signed incidence matrix C (COO -> CSC)
  structures/structures.py:152-188   indices_free / indices_fixed / indices_freefixed
  structures/structures.py:236-246   C_f = C[:, free], C_s = C[:, fixed]
  structures/structures.py:248-288   index_array  (labelled CSC, Nf x Nf)
  structures/structures.py:290-314   diag_indices
  structures/structures.py:316-350   diags (Nf x E selector of ones)
"""

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp


@dataclass
class OracleStructure:
    nodes: np.ndarray            # (N,) int64 keys
    edges: np.ndarray            # (E,2) int64 keys
    supports: np.ndarray         # (N,) int64 0/1
    edges_indexed: np.ndarray    # (E,2) int64
    connectivity: sp.csc_matrix  # (E,N)
    connectivity_free: sp.csc_matrix
    connectivity_fixed: sp.csc_matrix
    indices_free: np.ndarray     # (Nf,) int64
    indices_fixed: np.ndarray    # (Ns,) int64
    indices_freefixed: np.ndarray  # (N,) int64
    index_data: np.ndarray       # (nnz,) int64   index_array.data
    index_indices: np.ndarray    # (nnz,) int32   index_array.indices
    index_indptr: np.ndarray     # (Nf+1,) int32  index_array.indptr
    diag_indices: np.ndarray     # (Nf,) int64
    diags: sp.csr_matrix         # (Nf,E) ones

    @property
    def num_nodes(self):
        return int(self.nodes.size)

    @property
    def num_edges(self):
        return int(self.edges.shape[0])

    @property
    def num_free(self):
        return int(self.indices_free.size)

    @property
    def num_fixed(self):
        return int(self.indices_fixed.size)

    @property
    def nnz(self):
        return int(self.index_data.size)


def edges_indexed_from_keys(nodes, edges):
    """graphs.py:72-97: node key -> position in `nodes`, applied per endpoint."""
    lookup = {int(k): i for i, k in enumerate(np.asarray(nodes).tolist())}
    out = np.empty((len(edges), 2), dtype=np.int64)
    for e, (u, v) in enumerate(np.asarray(edges).tolist()):
        out[e, 0] = lookup[int(u)]
        out[e, 1] = lookup[int(v)]
    return out


def connectivity_matrix(edges_indexed, num_nodes):
    """graphs.py:176-211: -1 at the tail column, +1 at the head column, COO->CSC."""
    ei = np.asarray(edges_indexed)
    m = ei.shape[0]
    vals = np.concatenate((np.full(m, -1.0), np.full(m, 1.0)))
    rr = np.concatenate((np.arange(m), np.arange(m)))
    cc = np.concatenate((ei[:, 0], ei[:, 1]))
    return sp.coo_matrix((vals, (rr, cc)), shape=(m, num_nodes)).tocsc()


def index_array_arrays(c_free):
    """
    structures.py:248-288.  Every stored value of C_f is scaled by (edge id + 1),
    the product -(C_f^T @ scaled) labels each off-diagonal (i,j) with e+1, the
    diagonal is forced to an explicit 0, then cast to int.  Returned as the three
    raw arrays (data int64, indices int32, indptr int32).
    """
    scaled = c_free.copy()
    scaled.data *= np.take(np.arange(c_free.shape[0]) + 1, c_free.indices)
    lab = -(c_free.T @ scaled)
    lab.setdiag(0)
    lab = lab.astype(int)
    return lab.data, lab.indices, lab.indptr


def diag_positions(indices, indptr, n):
    """structures.py:290-314: storage positions whose row index equals their column."""
    cols = np.repeat(np.arange(n), np.diff(indptr))
    return np.flatnonzero(indices == cols).astype(np.int64)


def diag_selector(c_free):
    """structures.py:316-350: ones on the pattern of C_f, transposed to (Nf, E), CSR."""
    ones = np.ones_like(c_free.data)
    m = sp.csc_matrix((ones, c_free.indices, c_free.indptr), c_free.shape).tocsr().T
    return sp.csr_matrix(m)


def build_structure(nodes, edges, supports):
    nodes = np.asarray(nodes, dtype=np.int64)
    edges = np.asarray(edges, dtype=np.int64)
    supports = np.asarray(supports, dtype=np.int64)
    assert edges.ndim == 2 and edges.shape[1] == 2

    ei = edges_indexed_from_keys(nodes, edges)
    n = nodes.size
    C = connectivity_matrix(ei, n)

    ns = int(np.count_nonzero(supports))
    free = np.flatnonzero(supports == 0).astype(np.int64)      # structures.py:152-161
    fixed = np.flatnonzero(supports).astype(np.int64)          # structures.py:163-172
    assert fixed.size == ns
    freefixed = np.argsort(np.concatenate([free, fixed])).astype(np.int64)  # :174-188

    Cf = C[:, free]
    Cs = C[:, fixed]
    data, indices, indptr = index_array_arrays(Cf)
    dpos = diag_positions(indices, indptr, free.size)
    diags = diag_selector(Cf)

    return OracleStructure(
        nodes=nodes, edges=edges, supports=supports, edges_indexed=ei,
        connectivity=C, connectivity_free=Cf, connectivity_fixed=Cs,
        indices_free=free, indices_fixed=fixed, indices_freefixed=freefixed,
        index_data=data, index_indices=indices, index_indptr=indptr,
        diag_indices=dpos, diags=diags,
    )
