"""
CPU: the `jax_fdm.equilibrium` surface of SURVEY.md §8(b).  (1) Every name the reference's
`equilibrium/__init__.py:3-16` re-exports (the `__all__` lists of fdm.py, indexing.py, loads.py, models.py,
solvers/*.py, sparse.py, states.py, structures/*.py, copied below with their files) is importable from the mirror.
(2) The structure attributes goals and parameters read -- connectivity, connectivity_free / _fixed, adjacency,
connectivity_faces_vertices, connectivity_edges_faces -- against arrays produced by the reference's own classes
(tests/golden/*.npz: ref_conn_*, ref_cfree_*, ref_cfixed_*, ref_adj_*, *_ref_cfv_*, *_ref_cef_*).  (3) The load
helpers of loads.py against the oracle (itself pinned to the reference's loads.py), on CPU tensors.  (4) The
COMPAS-facing helpers of fdm.py on an array-backed stand-in for FDNetwork.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")

import jax_fdm_b200.equilibrium as eq  # noqa: E402
import oracle  # noqa: E402
from oracle import model as om  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

REFERENCE_ALL = {
    "fdm.py": ["fdm", "constrained_fdm", "model_from_sparsity", "structure_from_datastructure", "structure_from_network",
               "structure_from_mesh", "datastructure_state", "datastructure_validate", "datastructure_updated",
               "datastructure_update", "datastructure_edges_update", "datastructure_nodes_update"],
    "indexing.py": ["indices_from_keys"],
    "loads.py": ["calculate_edges_load", "calculate_faces_load", "edge_load_lcs", "edge_tributary_face_area",
                 "edge_tributary_faces_load", "edges_tributary_edges_load", "edges_tributary_faces_load", "face_load_lcs",
                 "face_xyz", "nodes_load_from_edges", "nodes_load_from_faces", "nodes_tributary_edges_load"],
    "models.py": ["EquilibriumModel", "EquilibriumModelSparse", "StiffnessMatrix"],
    "sparse.py": ["SparseSolveResidual", "SystemMatrixLHS", "SystemMatrixRHS", "SystemSolution", "blockdiag_matrix_sparse",
                  "register_sparse_solver", "sparse_solve", "sparse_solve_bwd", "sparse_solve_fwd", "splu_clear",
                  "splu_cpu", "splu_solve_cpu", "spsolve", "spsolve_cpu", "spsolve_gpu", "spsolve_gpu_ravel",
                  "spsolve_gpu_stack"],
    "states.py": ["DatastructureState", "EquilibriumParametersState", "EquilibriumState", "LoadState"],
    "solvers/fixed_point.py": ["fixed_point_bwd_adjoint", "fixed_point_bwd_adjoint_general", "fixed_point_bwd_fixedpoint",
                               "fixed_point_bwd_materialize", "fixed_point_fwd", "is_solver_fixedpoint", "solver_anderson",
                               "solver_fixedpoint", "solver_fixedpoint_implicit", "solver_forward"],
    "solvers/jaxopt.py": ["solver_jaxopt"],
    "solvers/least_squares.py": ["is_solver_leastsquares", "solver_dogleg", "solver_gauss_newton",
                                 "solver_levenberg_marquardt"],
    "solvers/nonlinear.py": ["nonlinear_bwd", "nonlinear_fwd", "solver_nonlinear_implicit"],
    "solvers/optimistix.py": ["solver_optimistix"],
    "solvers/root_finding.py": ["is_solver_root_finding", "solver_newton"],
    "solvers/types.py": ["SolverIterParams"],
    "structures/graphs.py": ["Graph", "GraphSparse", "adjacency_matrix", "connectivity_matrix"],
    "structures/meshes.py": ["Mesh", "MeshSparse", "face_matrix", "mesh_connectivity_edges_faces", "mesh_edges_faces"],
    "structures/structures.py": ["EquilibriumMeshStructure", "EquilibriumMeshStructureSparse", "EquilibriumStructure",
                                 "EquilibriumStructureSparse"],
}


def test_every_reference_name_is_importable():
    missing = [f"{f}:{n}" for f, names in REFERENCE_ALL.items() for n in names if not hasattr(eq, n)]
    assert not missing, missing
    ref = "/root/reference/src/jax_fdm/equilibrium"
    if os.path.isdir(ref):                      # in the build container: the list above IS the reference's
        import re
        for f, names in REFERENCE_ALL.items():
            m = re.search(r"__all__\s*=\s*\[(.*?)\]", open(os.path.join(ref, f)).read(), re.S)
            assert sorted(re.findall(r'"([A-Za-z_0-9]+)"', m.group(1))) == sorted(names), f


def test_model_defaults_and_attributes():
    m = eq.EquilibriumModelSparse()
    assert m.tmax == 100 and m.eta == 1e-6 and m.implicit_diff is True and m.is_load_local is False   # models.py:75-91
    assert m.itersolve_fn is eq.solver_forward and m.eq_iterative_fn == m.equilibrium_iterative_xyz
    for name in ("stiffness_matrix", "load_matrix", "residual_fixed_matrix", "residual_free_matrix", "nodes_free_positions",
                 "equilibrium", "equilibrium_state", "edges_vectors", "edges_lengths", "edges_forces", "nodes_residuals",
                 "nodes_positions", "nodes_equilibrium", "nodes_load", "faces_load", "edges_load", "load_xyz_matrix",
                 "load_xyz_matrix_from_r_fixed", "equilibrium_iterative_xyz", "equilibrium_iterative_residual",
                 "select_equilibrium_iterative_fn", "linearsolve_fn"):
        assert hasattr(m, name), name
    with pytest.raises(ValueError):
        m.select_equilibrium_iterative_fn(print)
    with pytest.raises(NotImplementedError):
        eq.solver_anderson(None, None, None, None)
    with pytest.raises(RuntimeError):
        eq.spsolve_cpu(None, None)


def _csc(g, tag):
    return g[f"ref_{tag}_data"], g[f"ref_{tag}_indices"], g[f"ref_{tag}_indptr"]


@pytest.mark.parametrize("name", ["arch10", "cablenet12", "olympia", "random40"])
def test_structure_connectivity_surface_vs_reference(name):
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    s = eq.EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=None)
    for tag, mat in (("conn", s.connectivity), ("cfree", s.connectivity_free), ("cfixed", s.connectivity_fixed),
                     ("adj", s.adjacency)):
        m = sp.csc_matrix(mat)
        m.sort_indices()
        d, i, p = _csc(g, tag)
        assert np.array_equal(m.data, d) and np.array_equal(m.indices, i) and np.array_equal(m.indptr, p), tag
    # the dense flavour holds the same matrices as arrays (graphs.py:99-116, structures.py:190-194)
    sd = eq.EquilibriumStructure(g["nodes"], g["edges"], g["supports"], device=None)
    assert isinstance(sd.connectivity, np.ndarray) and np.array_equal(sd.connectivity, s.connectivity.toarray())
    assert np.array_equal(sd.connectivity_free, s.connectivity_free.toarray())
    assert np.array_equal(sd.adjacency, s.adjacency.toarray())
    assert sd.connectivity._fdm_structure() is sd
    # module-level builders (graphs.py:176-250)
    assert (eq.connectivity_matrix(s.edges_indexed, s.num_nodes) != s.connectivity).nnz == 0
    assert (eq.adjacency_matrix(s.edges_indexed, s.num_nodes) != s.adjacency).nnz == 0


@pytest.mark.parametrize("tag", ["quad", "mixed"])
def test_mesh_surface_vs_reference(tag):
    g = dict(np.load(os.path.join(GOLDEN, "mesh_faces.npz")))
    args = (g[f"{tag}_nodes"], g[f"{tag}_faces"], g[f"{tag}_edges"], g[f"{tag}_supports"])
    s = eq.EquilibriumMeshStructureSparse(*args, device=None)
    assert isinstance(s, eq.MeshSparse) and isinstance(s, eq.EquilibriumStructureSparse)
    assert np.array_equal(s.faces_indexed, g[f"{tag}_ref_faces_indexed"])
    assert np.array_equal(s.edges_faces_indexed, g[f"{tag}_ref_edges_faces_indexed"])
    for key, mat in (("cfv", s.connectivity_faces_vertices), ("cef", s.connectivity_edges_faces)):
        m = sp.csr_matrix(mat)
        m.sort_indices()
        assert np.array_equal(m.data, g[f"{tag}_ref_{key}_data"]) and np.array_equal(m.indices, g[f"{tag}_ref_{key}_indices"])
        assert np.array_equal(m.indptr, g[f"{tag}_ref_{key}_indptr"])
    assert s.num_faces == len(args[1]) and s.num_vertices == len(args[0]) and s.face_index[0] == 0
    sd = eq.EquilibriumMeshStructure(*args, device=None)
    assert np.array_equal(sd.connectivity_edges_faces, s.connectivity_edges_faces.toarray())
    # a plain Mesh derives its edges from the faces when none are given (meshes.py:207-245)
    quads = g["quad_faces"]
    m = eq.Mesh(g["quad_nodes"], quads)
    und = {tuple(sorted(e)) for e in m.edges.tolist()}
    assert und == {tuple(sorted(e)) for e in g["quad_edges"].tolist()} and len(und) == m.num_edges
    assert np.array_equal(eq.face_matrix(m.faces_indexed).toarray(), m.connectivity_faces_vertices)


@pytest.mark.parametrize("tag", ["quad", "mixed"])
@pytest.mark.parametrize("is_local", [False, True])
def test_load_helpers_vs_oracle_on_cpu_tensors(tag, is_local):
    g = dict(np.load(os.path.join(GOLDEN, "mesh_faces.npz")))
    s = eq.EquilibriumMeshStructureSparse(g[f"{tag}_nodes"], g[f"{tag}_faces"], g[f"{tag}_edges"], g[f"{tag}_supports"],
                                          device=None)
    so = oracle.build_structure(g[f"{tag}_nodes"], g[f"{tag}_edges"], g[f"{tag}_supports"])
    xyz = g[f"{tag}_{'local' if is_local else 'global'}_state_xyz"]
    rng = np.random.default_rng(5)
    eload = 0.1 * rng.standard_normal((s.num_edges, 3))
    fload = g[f"{tag}_faces_load"]
    T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
    zero = np.zeros_like(xyz)
    got_e = eq.nodes_load_from_edges(T(xyz), T(eload), s, is_local).numpy()
    assert np.allclose(got_e, om.nodes_load_edges(xyz, zero, eload, so, is_local), rtol=1e-12, atol=1e-14)
    got_f = eq.nodes_load_from_faces(T(xyz), T(fload), s, is_local).numpy()
    want_f = om.nodes_load_faces(xyz, zero, fload, so, s.faces_indexed, s.edges_faces_indexed, is_local, pad="wrap")
    assert np.allclose(got_f, want_f, rtol=1e-12, atol=1e-14)
    # the per-element helpers on single elements
    e0 = s.edges_indexed[3]
    assert np.allclose(eq.edge_load_lcs(T(xyz), e0, T(eload[3])).numpy(), eload[3] @ om.line_lcs(xyz[e0]), atol=1e-14)
    f0 = s.faces_indexed[1]
    assert np.allclose(eq.face_load_lcs(T(xyz), f0, T(fload[1])).numpy(),
                       om.faces_load_global(xyz, fload[1:2], s.faces_indexed[1:2], True)[0], atol=1e-14)
    fx = eq.face_xyz(T(xyz), s.faces_indexed[0]).numpy()
    face = s.faces_indexed[0]
    assert np.array_equal(fx, xyz[np.where(face >= 0, face, face[0])])
    tri = T(np.array([[0.0, 0, 0], [2.0, 0, 0]]))
    assert float(eq.edge_tributary_face_area(tri, T([0.0, 3.0, 0.0]))) == 3.0
    assert np.allclose(eq.nodes_tributary_edges_load(T(eload), s).numpy().sum(0), eload.sum(0), atol=1e-14)


def test_indices_from_keys():
    nodes = np.array([10, 4, 7, 22])
    assert eq.indices_from_keys(nodes, [7, 10]).tolist() == [2, 0] and eq.indices_from_keys(nodes, 22).tolist() == [3]
    edges = np.array([[10, 4], [4, 7], [7, 22]])
    assert eq.indices_from_keys(edges, [(7, 22), (10, 4)]).tolist() == [2, 0] and eq.indices_from_keys(edges, (4, 7)).tolist() == [1]
    with pytest.raises(KeyError):
        eq.indices_from_keys(edges, (4, 10))          # directed: the reverse of an edge is absent
    with pytest.raises(KeyError):
        eq.indices_from_keys(nodes, 5)
    with pytest.raises(KeyError):
        eq.indices_from_keys(edges, (99, 4))


def test_blockdiag_matrix_sparse():
    g = dict(np.load(os.path.join(GOLDEN, "arch10.npz")))
    s = eq.EquilibriumStructureSparse(g["nodes"], g["edges"], g["supports"], device=None)
    A = eq.StiffnessMatrix(torch.as_tensor(g["ref_K_data"]), s)
    data, indices, indptr, shape = eq.blockdiag_matrix_sparse(A, 3)
    n = s.num_free
    one = sp.csc_matrix((g["ref_K_data"], s.index_array.indices, s.index_array.indptr), shape=(n, n))
    got = sp.csc_matrix((data.numpy(), indices, indptr), shape=shape)
    assert shape == (3 * n, 3 * n) and (got != sp.block_diag([one] * 3, format="csc")).nnz == 0


class _ArrayNetwork:
    """Array-backed stand-in with the accessors of FDNetwork that equilibrium/fdm.py and states.py call."""

    def __init__(self, xyz, edges, supports, q, loads):
        self.xyz, self._edges, self.sup, self.q, self.loads = (np.array(a, dtype=float) if i != 1 else [tuple(e) for e in a]
                                                               for i, a in enumerate((xyz, edges, supports, q, loads)))
        self.node_attr = {k: {} for k in range(len(self.xyz))}
        self.edge_attr = {e: {} for e in self._edges}

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def nodes(self): return iter(range(len(self.xyz)))
    def edges(self): return iter(self._edges)
    def is_node_support(self, n): return bool(self.sup[n])
    def number_of_supports(self): return int(self.sup.sum())
    def number_of_edges(self): return len(self._edges)
    def number_of_nodes(self): return len(self.xyz)
    def edges_forcedensities(self): return list(self.q)
    def nodes_coordinates(self): return self.xyz.tolist()
    def nodes_fixedcoordinates(self): return self.xyz[self.sup > 0].tolist()
    def nodes_loads(self): return self.loads.tolist()
    def nodes_residual(self): return np.zeros_like(self.xyz).tolist()
    def edges_loads(self): return np.zeros((len(self._edges), 3)).tolist()
    def edges_lengths(self): return [float(np.linalg.norm(self.xyz[v] - self.xyz[u])) for u, v in self._edges]
    def edges_forces(self): return [qi * l for qi, l in zip(self.q, self.edges_lengths())]
    def index_edge(self): return dict(enumerate(self._edges))
    def index_node(self): return dict(enumerate(range(len(self.xyz))))
    def node_attribute(self, key, name, value): self.node_attr[key][name] = value
    def edge_attribute(self, key, name, value): self.edge_attr[key][name] = value


def test_datastructure_helpers_on_a_network_stand_in():
    from jax_fdm_b200 import datasets
    prob = datasets.arch_problem(6)
    net = _ArrayNetwork(prob.xyz0, prob.edges, prob.supports, prob.q, prob.loads)
    eq.datastructure_validate(net)
    ds = eq.datastructure_state(net, sparse=True, device=None)
    assert isinstance(ds, eq.DatastructureState) and isinstance(ds.structure, eq.EquilibriumStructureSparse)
    assert np.array_equal(ds.structure.supports, prob.supports) and np.array_equal(ds.structure.edges, prob.edges)
    assert np.allclose(ds.parameters.q.numpy(), prob.q) and np.allclose(ds.parameters.xyz_fixed.numpy(), prob.xyz_fixed)
    assert ds.parameters.loads.edges == 0.0 and ds.parameters.loads.faces == 0.0
    assert np.allclose(ds.eq_state.vectors.numpy(), prob.xyz0[prob.edges[:, 1]] - prob.xyz0[prob.edges[:, 0]])
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    st, _ = oracle.forward(prob.q, prob.xyz_fixed, prob.loads, so)
    state = eq.EquilibriumState(**{k: torch.as_tensor(v) for k, v in st.items()})
    new = eq.datastructure_updated(net, state, ds.parameters)
    assert net.node_attr[2] == {} and new.node_attr[2]["z"] == st["xyz"][2, 2] and new.node_attr[2]["rz"] == st["residuals"][2, 2]
    assert new.edge_attr[(1, 2)]["length"] == st["lengths"][1, 0] and new.edge_attr[(1, 2)]["q"] == prob.q[1]
    bad = _ArrayNetwork(prob.xyz0, prob.edges, prob.supports, np.where(np.arange(prob.q.size) == 1, 0.0, prob.q), prob.loads)
    with pytest.raises(AssertionError):
        eq.datastructure_validate(bad)
