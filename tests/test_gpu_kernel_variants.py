"""
The factor-kernel variants behind the A/B switches (read once per process, hence subprocesses): the
two-column tensor-core kernel and the rank-1 kernel at bandwidths where the four-column kernel is the
default, and other CTA sizes.  Each variant must reproduce the oracle's SuperLU positions, the adjoint solve
with the stored factor, and (tensor-core kernels) the signed second pass on a mixed-sign design.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %(root)r)
import oracle
from jax_fdm_b200 import datasets
import jax_fdm_b200.equilibrium as eq
T = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device="cuda:0")
signed = %(signed)d
for nx, B in ((12, 9), (30, 21)):
    p = datasets.pillow_problem(nx)
    s = eq.EquilibriumStructureSparse(p.nodes, p.edges, p.supports, device=0)
    so = oracle.build_structure(p.nodes, p.edges, p.supports)
    rng = np.random.default_rng(nx)
    qb = datasets.pillow_batch_q(p.edges.shape[0], 0, B)
    if signed:
        qb[1] *= np.where(rng.random(qb.shape[1]) < 0.3, -1.0, 1.0)
    K, rhs = eq.ops.assemble(s, T(qb), T(p.xyz_fixed), T(p.loads))
    ws = eq.ops.Workspace()
    x, info = eq.ops.solve(s, K, rhs, solver="cholesky", return_info=True, workspace=ws)
    g = rng.standard_normal((B, s.num_free, 3))
    lam = eq.ops.solve(s, K, T(g), solver="cholesky", workspace=ws, reuse_factor=True)
    info = info.cpu().numpy()
    assert (info[:, 0] == 0).all(), info[:, 0]
    assert info[1, 3] == (0 if signed else -1) and info[0, 3] == -1
    for b in (0, 1, B - 1):
        Kb = oracle.stiffness_matrix_sparse(qb[b], so)
        xo = oracle.sparse_solve(Kb, oracle.load_matrix(qb[b], p.xyz_fixed, p.loads, so))
        lo = oracle.sparse_solve(Kb, g[b])
        tol = 1e-9 if (signed and b == 1) else 1e-10
        assert np.abs(x[b].cpu().numpy() - xo).max() <= tol * np.abs(xo).max(), (nx, b)
        assert np.abs(lam[b].cpu().numpy() - lo).max() <= tol * np.abs(lo).max(), (nx, b)
print("variant ok")
"""


@pytest.mark.parametrize("env,signed", [({"FDM_CHOL_PANEL2": "1"}, 1), ({"FDM_CHOL_SIMT": "1"}, 0),
                                        ({"FDM_CHOL_FACWARPS": "5"}, 1), ({"FDM_CHOL_FACWARPS": "1"}, 1)])
def test_factor_kernel_variant(env, signed):
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "signed": signed}], env=e, capture_output=True,
                       text=True, timeout=280)
    assert r.returncode == 0 and "variant ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
