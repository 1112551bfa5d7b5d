"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + scipy) of the JAX FDM equilibrium hot path: the forward
force-density solve and its implicit-differentiation adjoint.  All `file:line`
citations are into the reference tree (`src/jax_fdm/equilibrium/...`).

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` -- as the checker or as the
timed CPU baseline, never as the product.  Nothing under `jax_fdm_b200/` imports it.

Why numpy/scipy and not C: the reference is pure Python and its own CPU sparse
arithmetic *is* `scipy.sparse.linalg.spsolve` (SuperLU) reached through
`jax.pure_callback` (`sparse.py:130-159`); the topology arrays are built with
scipy.sparse (`structures/graphs.py:176-211`, `structures/structures.py:248-350`).
The same scipy calls here reproduce the same bits.

Parity pinning: `tests/golden/make_golden.py` executes the reference's *real*
`structures/graphs.py`, `structures/structures.py` and the forward half of
`models.py` (under a numpy shim that stands in for the absent `jax` package) and
commits the resulting arrays under `tests/golden/`; `tests/test_oracle_golden.py`
checks this restatement against them bit-for-bit (integers) / to 1e-12 (floats),
and against the reference's own KATs (arch baseline JSON, olympia / stuttgart21
CSVs, Liew dome statistics, README cable-net table).  The same is done for the local
frames and tributary loads (make_golden_loads.py), the goals / error terms
(make_golden_goals.py), the tmax > 1 fixed point with edge loads
(make_golden_fixed_point.py) and the mesh structures + face loads (make_golden_mesh.py), each with its own tests/test_oracle_*_golden.py.
"""

from .structure import OracleStructure, build_structure  # noqa: F401
from .model import (  # noqa: F401
    stiffness_matrix_sparse,
    stiffness_matrix_dense,
    residual_fixed_matrix,
    load_matrix,
    sparse_solve,
    dense_solve,
    nodes_positions,
    equilibrium_state,
    equilibrium_state_vjp,
    sparse_solve_bwd,
    forward,
    forward_adjoint,
    nodes_load_edges,
    nodes_load_faces,
    fixed_point_forward,
)
from . import goals  # noqa: F401
