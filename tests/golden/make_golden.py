#!/usr/bin/env python
"""
Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE'S OWN
CODE for the hot path:

    src/jax_fdm/equilibrium/structures/graphs.py
    src/jax_fdm/equilibrium/structures/structures.py
    src/jax_fdm/equilibrium/states.py
    src/jax_fdm/equilibrium/sparse.py      (forward: spsolve_cpu -> scipy SuperLU)
    src/jax_fdm/equilibrium/models.py      (forward: EquilibriumModel[Sparse].__call__, tmax=1)

`jax`, `equinox`, `jaxtyping`, `compas`, `lineax`, ... are not installed in this
image, so the reference cannot be imported as a package.  This script loads the
five files above *unmodified, from where they lie under /root/reference* with a
small numpy stand-in for the parts of `jax` they touch on the forward path
(jnp -> numpy, BCOO/BCSR/CSC -> thin scipy.sparse wrappers, pure_callback -> direct
call, eqx.Module -> plain class).  Everything else they import (solvers, loads,
geometry, datastructures, compas) is replaced by inert placeholder modules; none
of it executes at tmax=1 with nodal loads.  The reverse pass is JAX autodiff and
cannot run here -- gradients are pinned by finite differences in tests/ instead.

Run here (needs /root/reference):   python tests/golden/make_golden.py
Outputs: tests/golden/*.npz  (inputs + every array the reference produced).
No reference source is copied; only arrays it computed are stored.
"""

import csv
import importlib.abc
import importlib.machinery
import importlib.util
import json
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

REF = os.environ.get("JAX_FDM_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))  # repo root, for jax_fdm_b200.datasets

REAL = {
    "jax_fdm.equilibrium.structures.graphs",
    "jax_fdm.equilibrium.structures.structures",
    "jax_fdm.equilibrium.states",
    "jax_fdm.equilibrium.sparse",
    "jax_fdm.equilibrium.models",
}
FAKE_ROOTS = ("jax", "jaxtyping", "equinox", "compas", "jax_fdm", "lineax", "jaxopt", "optimistix")


# --------------------------------------------------------------------------
# numpy stand-ins
# --------------------------------------------------------------------------

class JArr(np.ndarray):
    """ndarray with jax's functional `.at[idx].set(v)`."""

    @property
    def at(self):
        arr = self

        class _At:
            def __getitem__(self, idx):
                class _Ref:
                    def set(self, v):
                        out = arr.copy()
                        out[idx] = v
                        return out

                    def add(self, v):
                        out = arr.copy()
                        np.add.at(out, idx, v)
                        return out
                return _Ref()
        return _At()


def jarr(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(JArr)


class _Sparse:
    """BCOO / BCSR stand-in backed by a scipy matrix."""

    def __init__(self, m):
        self.m = m
        self.shape = m.shape

    @classmethod
    def from_scipy_sparse(cls, m):
        return cls(m)

    def __matmul__(self, other):
        return jarr(self.m @ np.asarray(other))

    @property
    def T(self):
        return type(self)(self.m.T)

    def transpose(self):
        return self.T

    def todense(self):
        return jarr(self.m.toarray())


class BCOO(_Sparse):
    pass


class BCSR(_Sparse):
    pass


class CSC:
    """jax.experimental.sparse.CSC stand-in: just the three arrays and a shape."""

    def __init__(self, args, shape=None):
        self.data, self.indices, self.indptr = args
        self.shape = tuple(shape)

    def __matmul__(self, other):
        m = sp.csc_matrix((np.asarray(self.data), self.indices, self.indptr), shape=self.shape)
        return jarr(m @ np.asarray(other))


class _Anything:
    """Placeholder: subscriptable, callable, or-able; used for type annotations."""

    def __init__(self, name="?"):
        self._n = name

    def __getitem__(self, k):
        return self

    def __class_getitem__(cls, k):
        return cls()

    def __call__(self, *a, **k):
        return self

    def __or__(self, o):
        return self

    __ror__ = __or__

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything(self._n + "." + k)

    def __mro_entries__(self, bases):
        return (object,)

    def __bool__(self):
        return True


class _JnpModule(types.ModuleType):
    def __getattr__(self, name):
        return getattr(np, name)


def _make_jnp():
    m = _JnpModule("jax.numpy")
    m.float64 = np.float64
    m.int64 = np.int64
    m.asarray = lambda x, dtype=None: jarr(x, dtype)
    m.array = lambda x, dtype=None: jarr(np.array(x, dtype=dtype))
    m.flatnonzero = lambda x, size=None: jarr(np.flatnonzero(np.asarray(x)))
    m.count_nonzero = lambda x: np.count_nonzero(np.asarray(x))
    m.concatenate = lambda xs, axis=0: jarr(np.concatenate([np.asarray(x) for x in xs], axis=axis))
    m.reshape = lambda x, s: jarr(np.reshape(np.asarray(x), s))
    m.ones_like = lambda x: jarr(np.ones_like(np.asarray(x)))
    lin = types.ModuleType("jax.numpy.linalg")
    lin.solve = lambda a, b: jarr(np.linalg.solve(np.asarray(a), np.asarray(b)))
    lin.norm = lambda x, axis=None, keepdims=False: jarr(np.linalg.norm(np.asarray(x), axis=axis, keepdims=keepdims))
    m.linalg = lin
    return m


class _CustomVJP:
    def __init__(self, fn, nondiff_argnums=()):
        self.fn = fn
        self.__doc__ = fn.__doc__

    def __call__(self, *a, **k):
        return self.fn(*a, **k)

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd


class _PlaceholderModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(self.__name__ + "." + name)


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root not in FAKE_ROOTS:
            return None
        if fullname in REAL:
            base = os.path.join(SRC, fullname.replace(".", os.sep))
            if os.path.isdir(base):     # a real package: its own __init__.py
                return importlib.util.spec_from_file_location(fullname, os.path.join(base, "__init__.py"),
                                                              submodule_search_locations=[base])
            return importlib.util.spec_from_file_location(fullname, base + ".py")
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        return _build_fake(spec.name)

    def exec_module(self, module):
        # runs after the module is in sys.modules, so submodule imports can see their parent
        _populate_late(module)


def _build_fake(name):
    if name == "jax.numpy":
        return _make_jnp()
    m = _PlaceholderModule(name)
    m.__path__ = []
    if name == "jax":
        m.Array = np.ndarray
        m.custom_vjp = _CustomVJP
        m.pure_callback = lambda cb, shape, *args, **kw: jarr(cb(*[np.asarray(a) for a in args]))
        m.default_backend = lambda: "cpu"
    elif name == "jax.experimental.sparse":
        m.BCOO, m.BCSR, m.CSC = BCOO, BCSR, CSC
    elif name == "equinox":
        m.Module = type("Module", (), {})
    elif name == "jax_fdm":
        m.DTYPE_NP = np.float64
        m.DTYPE_JAX = np.float64
        m.DTYPE_INT_NP = np.int64
        m.DTYPE_INT_JAX = np.int64
    elif name == "jax_fdm.datastructures":
        m.FDMesh = type("FDMesh", (), {})
        m.FDNetwork = type("FDNetwork", (), {})
    return m


def _populate_late(m):
    name = m.__name__
    if name == "jax_fdm.equilibrium.structures":
        # the package __init__ star-imports graphs, meshes, structures; meshes needs COMPAS
        # and is not on the path, so expose only what models.py/states.py ask for.
        g = importlib.import_module("jax_fdm.equilibrium.structures.graphs")
        s = importlib.import_module("jax_fdm.equilibrium.structures.structures")
        for mod in (g, s):
            for k in mod.__all__:
                setattr(m, k, getattr(mod, k))
        m.EquilibriumMeshStructure = type("EquilibriumMeshStructure", (s.EquilibriumStructure,), {})
    elif name == "jax_fdm.equilibrium.structures.meshes":
        g = importlib.import_module("jax_fdm.equilibrium.structures.graphs")
        m.Mesh = type("Mesh", (g.Graph,), {})
        m.MeshSparse = type("MeshSparse", (g.GraphSparse,), {})


def load_reference():
    if not os.path.isdir(SRC):
        raise SystemExit(f"reference tree not found at {REF}")
    sys.meta_path.insert(0, _Finder())
    structures = importlib.import_module("jax_fdm.equilibrium.structures.structures")
    states = importlib.import_module("jax_fdm.equilibrium.states")
    models = importlib.import_module("jax_fdm.equilibrium.models")
    return structures, states, models


# --------------------------------------------------------------------------
# cases
# --------------------------------------------------------------------------

def case_arch():
    from jax_fdm_b200.datasets import arch_problem
    return arch_problem(10)._asdict()


def case_meshgrid(kind, nx):
    from jax_fdm_b200 import datasets as D
    p = {"readme": D.cablenet_readme_problem, "cablenet": D.cablenet_problem,
         "pillow": D.pillow_problem}[kind](nx)
    return p._asdict()


def case_csv(name):
    """tests/test_reference_solver.py:23-65, without COMPAS."""
    d = os.path.join(REF, "tests", "data", "reference")
    keys, xyz, sup = [], [], []
    with open(os.path.join(d, f"{name}-nodes.csv")) as f:
        for i, row in enumerate(csv.reader(f)):
            if i == 0:
                continue
            keys.append(int(row[0]) - 1)
            p = [float(row[1]) + float(row[5]), float(row[2]) + float(row[6]), float(row[3]) + float(row[7])]
            xyz.append(p)
            sup.append(1 if row[4] == "0" else 0)
    keys = np.asarray(keys, dtype=np.int64)
    xyz = np.asarray(xyz)
    pos = {int(k): i for i, k in enumerate(keys)}
    # Graph.add_edge / edges(): `for u in edge: for v in edge[u]` where edge[u] exists for
    # every node in insertion order -> group edges by tail in node order, keep file order inside.
    adj = {int(k): {} for k in keys}
    with open(os.path.join(d, f"{name}-edges.csv")) as f:
        for i, row in enumerate(csv.reader(f)):
            if i == 0:
                continue
            u, v = int(row[0]) - 1, int(row[1]) - 1
            adj[u][v] = float(row[4])
    edges, q = [], []
    for u in adj:
        for v, force in adj[u].items():
            length = np.linalg.norm(xyz[pos[u]] - xyz[pos[v]])
            edges.append((u, v))
            q.append(force / length)
    supports = np.asarray(sup, dtype=np.int64)
    return dict(nodes=keys, edges=np.asarray(edges, dtype=np.int64), supports=supports,
                q=np.asarray(q), xyz_fixed=xyz[np.flatnonzero(supports)],
                loads=np.zeros_like(xyz), xyz0=xyz)


def case_liew():
    """tests/data/liew_dome.json (COMPAS graph JSON; iteration order = JSON order)."""
    with open(os.path.join(REF, "tests", "data", "liew_dome.json")) as f:
        d = json.load(f)["data"]
    dn, de = d["default_node_attributes"], d["default_edge_attributes"]
    keys, xyz, sup, loads = [], [], [], []
    for k, a in d["node"].items():
        a = {**dn, **a}
        keys.append(int(k))
        xyz.append([a["x"], a["y"], a["z"]])
        loads.append([a["px"], a["py"], a["pz"]])
        sup.append(1 if a["is_support"] else 0)
    edges, q = [], []
    for u, nbrs in d["edge"].items():
        for v, a in nbrs.items():
            a = {**de, **a}
            edges.append((int(u), int(v)))
            q.append(a["q"])
    supports = np.asarray(sup, dtype=np.int64)
    xyz = np.asarray(xyz)
    return dict(nodes=np.asarray(keys, dtype=np.int64), edges=np.asarray(edges, dtype=np.int64),
                supports=supports, q=np.asarray(q), xyz_fixed=xyz[np.flatnonzero(supports)],
                loads=np.asarray(loads), xyz0=xyz)


def case_random(seed, n=40, extra=50):
    """Irregular simple graph, shuffled non-contiguous keys, random edge directions, mixed-sign-free q."""
    rng = np.random.default_rng(seed)
    keys = rng.permutation(np.arange(100, 100 + 3 * n, 3))[:n].astype(np.int64)
    pairs = set()
    order = rng.permutation(n)
    for a, b in zip(order[:-1], order[1:]):         # spanning path keeps it connected
        pairs.add((int(min(a, b)), int(max(a, b))))
    while len(pairs) < n - 1 + extra:
        a, b = rng.integers(0, n, size=2)
        if a != b:
            pairs.add((int(min(a, b)), int(max(a, b))))
    pairs = [p if rng.random() < 0.5 else (p[1], p[0]) for p in sorted(pairs)]
    pairs = [pairs[i] for i in rng.permutation(len(pairs))]
    edges = np.asarray([(keys[a], keys[b]) for a, b in pairs], dtype=np.int64)
    supports = np.zeros(n, dtype=np.int64)
    supports[rng.choice(n, size=max(3, n // 8), replace=False)] = 1
    xyz = rng.normal(size=(n, 3))
    q = rng.uniform(0.5, 3.0, size=len(pairs))
    loads = rng.normal(size=(n, 3)) * 0.1
    return dict(nodes=keys, edges=edges, supports=supports, q=q,
                xyz_fixed=xyz[np.flatnonzero(supports)], loads=loads, xyz0=xyz)


# --------------------------------------------------------------------------
# run the reference and dump
# --------------------------------------------------------------------------

def run_reference(case, structures, states, models, dense_too):
    nodes, edges, supports = case["nodes"], case["edges"], case["supports"]
    out = {k: np.asarray(v) for k, v in case.items()}

    s = structures.EquilibriumStructureSparse(nodes, edges, supports)
    out["ref_edges_indexed"] = np.asarray(s.edges_indexed)
    out["ref_indices_free"] = np.asarray(s.indices_free)
    out["ref_indices_fixed"] = np.asarray(s.indices_fixed)
    out["ref_indices_freefixed"] = np.asarray(s.indices_freefixed)
    ia = s.index_array
    out["ref_index_data"] = np.asarray(ia.data)
    out["ref_index_indices"] = np.asarray(ia.indices)
    out["ref_index_indptr"] = np.asarray(ia.indptr)
    out["ref_diag_indices"] = np.asarray(s.diag_indices)
    dg = sp.csr_matrix(s.diags.m)
    dg.sort_indices()
    out["ref_diags_indices"] = dg.indices
    out["ref_diags_indptr"] = dg.indptr
    for nm, mat in (("cfree", s.connectivity_free.m), ("cfixed", s.connectivity_fixed.m),
                    ("conn", s.connectivity.m), ("adj", s.adjacency.m)):
        mat = sp.csc_matrix(mat)
        out[f"ref_{nm}_data"] = mat.data
        out[f"ref_{nm}_indices"] = mat.indices
        out[f"ref_{nm}_indptr"] = mat.indptr

    q = jarr(case["q"], np.float64)
    xs = jarr(case["xyz_fixed"], np.float64)
    loads = jarr(case["loads"], np.float64)
    params = states.EquilibriumParametersState(q, xs, states.LoadState(loads, 0.0, 0.0))

    model = models.EquilibriumModelSparse(tmax=1)
    K = model.stiffness_matrix(q, s)
    out["ref_K_data"] = np.asarray(K.data)
    out["ref_P"] = np.asarray(model.load_matrix(q, xs, loads, s))
    out["ref_x_free"] = np.asarray(model.equilibrium(q, xs, loads, s))
    st = model(params, s)
    for f in st._fields:
        out[f"ref_state_{f}"] = np.asarray(getattr(st, f))

    if dense_too:
        sd = structures.EquilibriumStructure(nodes, edges, supports)
        md = models.EquilibriumModel(tmax=1)
        std = md(params, sd)
        out["ref_dense_K"] = np.asarray(md.stiffness_matrix(q, sd))
        for f in std._fields:
            out[f"ref_dense_state_{f}"] = np.asarray(getattr(std, f))
    return out


def main():
    structures, states, models = load_reference()
    cases = {
        "arch10": (case_arch(), True),
        "readme10": (case_meshgrid("readme", 10), True),
        "cablenet12": (case_meshgrid("cablenet", 12), True),
        "pillow30": (case_meshgrid("pillow", 30), False),
        "olympia": (case_csv("olympia"), True),
        "stuttgart21": (case_csv("stuttgart21"), True),
        "liew_dome": (case_liew(), False),
        "random40": (case_random(7), True),
    }
    for name, (case, dense_too) in cases.items():
        out = run_reference(case, structures, states, models, dense_too)
        if name == "arch10":
            # the reference's own captured baseline (tests/baselines/arch_coordinates.json,
            # compared at rtol=1e-6, atol=1e-9 by tests/conftest.py:24-42)
            with open(os.path.join(REF, "tests", "baselines", "arch_coordinates.json")) as f:
                out["kat_arch_coordinates"] = np.asarray(json.load(f))
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name:12s} N={case['nodes'].size:5d} E={case['edges'].shape[0]:5d} "
              f"nnz={out['ref_index_data'].size:6d} -> {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
