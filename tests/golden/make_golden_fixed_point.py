#!/usr/bin/env python
"""
Golden vectors for the tmax > 1 fixed-point equilibrium with shape-dependent edge loads (global and local
frames), made by EXECUTING THE REFERENCE'S OWN CODE:

    src/jax_fdm/equilibrium/models.py                 (EquilibriumModelSparse.__call__, equilibrium_iterative_xyz)
    src/jax_fdm/equilibrium/solvers/fixed_point.py    (solver_fixedpoint_implicit -> solver_forward)
    src/jax_fdm/equilibrium/loads.py, geometry/geometry.py, structures/*.py, sparse.py (spsolve_cpu), states.py

loaded unmodified from /root/reference under the numpy stand-in for `jax` of make_golden.py, with
`equinox.internal.while_loop` as a Python loop.  Only the forward pass runs (the reverse pass is JAX autodiff).

Run here (needs /root/reference):   python tests/golden/make_golden_fixed_point.py
Output: tests/golden/fixed_point.npz
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402
import make_golden_loads as GL  # noqa: E402,F401

G.REAL |= {"jax_fdm.equilibrium.solvers.fixed_point"}
_prev = G._build_fake


def _build_fake(name):
    m = _prev(name)
    if name == "equinox.internal":
        def while_loop(cond_fun, body_fun, init_val, max_steps=None, kind=None):
            val, steps = init_val, 0
            while bool(cond_fun(val)) and (max_steps is None or steps < max_steps):
                val = body_fun(val)
                steps += 1
            while_loop.steps = steps
            return val
        m.while_loop = while_loop
    elif name == "jax.lax":
        m.cond = lambda pred, t, f: t() if bool(pred) else f()
    return m


G._build_fake = _build_fake
_prev_late = G._populate_late


def _populate_late(m):
    _prev_late(m)
    if m.__name__ == "jax_fdm.equilibrium.solvers":      # models.py imports the fixed-point solvers from the package
        real = importlib.import_module("jax_fdm.equilibrium.solvers.fixed_point")
        for k in ("solver_fixedpoint_implicit", "solver_forward"):
            setattr(m, k, getattr(real, k))


G._populate_late = _populate_late


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from jax_fdm_b200 import datasets

    structures, states, models = G.load_reference()
    fp = importlib.import_module("jax_fdm.equilibrium.solvers.fixed_point")
    eqx_internal = importlib.import_module("equinox.internal")
    out = {}
    for tag, nx, seed in (("net6", 6, 5), ("net9", 9, 1)):
        prob = datasets.cablenet_problem(nx, seed=seed)
        rng = np.random.default_rng(seed)
        E = prob.edges.shape[0]
        eload = np.zeros((E, 3))
        eload[:, 2] = -0.05 * rng.uniform(0.5, 1.5, size=E)
        eload[:, 0] = 0.01 * rng.standard_normal(E)
        s = structures.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports)
        q, xs, loads = G.jarr(prob.q), G.jarr(prob.xyz_fixed), G.jarr(prob.loads)
        for k in ("nodes", "edges", "supports", "q", "xyz_fixed", "loads"):
            out[f"{tag}_{k}"] = np.asarray(getattr(prob, k))
        out[f"{tag}_edges_load"] = eload
        for local in (False, True):
            model = models.EquilibriumModelSparse(tmax=100, eta=1e-10, is_load_local=local, itersolve_fn=fp.solver_forward,
                                                  implicit_diff=True, verbose=False)
            params = states.EquilibriumParametersState(q, xs, states.LoadState(loads, G.jarr(eload), 0.0))
            st = model(params, s)
            key = f"{tag}_{'local' if local else 'global'}"
            out[f"{key}_steps"] = np.asarray(getattr(eqx_internal.while_loop, "steps", -1))
            for f in st._fields:
                out[f"{key}_state_{f}"] = np.asarray(getattr(st, f))
            print(key, "steps", out[f"{key}_steps"], "max |xyz|", np.abs(out[f"{key}_state_xyz"]).max())
    path = os.path.join(HERE, "fixed_point.npz")
    np.savez_compressed(path, **out)
    print(f"fixed_point.npz: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
