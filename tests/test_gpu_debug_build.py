"""
GPU: small cases on the -DFDM_DEBUG build of the library (lib/libfdm_b200_debug.so, built by
`__graft_entry__.build()`): every index the factor kernels take from the plan's tables is range-checked where it is
consumed (FDM_ASSERT in csrc/launch.h: a failed check prints file:line and traps, so the launch fails).  The pool has
no compute-sanitizer; this is what stands in for it.  Runs in a subprocess because the library path is fixed at
import (FDM_B200_LIB).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEBUG_LIB = os.path.join(ROOT, "jax_fdm_b200", "lib", "libfdm_b200_debug.so")

SCRIPT = r"""
import numpy as np, torch
import oracle
from jax_fdm_b200 import datasets, _lib
import jax_fdm_b200.equilibrium as eq
from jax_fdm_b200.pipeline import NativeForwardAdjoint
assert _lib.LIB_PATH.endswith("libfdm_b200_debug.so")
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
for nx, B in ((30, 6), (30, 1300), (24, 5), (12, 9), (30, 1)):       # two-sided, one-sided, NBL = 3, narrow band, single
    prob = datasets.pillow_problem(nx)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    x0 = prob.xyz0[s.indices_free]
    fa = NativeForwardAdjoint(s, B, chunks=min(2, B), solver="cholesky")
    fa.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    loss, q_bar = fa()
    def cot(st, xf):
        c = np.zeros_like(st["xyz"]); c[so.indices_free] = xf - x0
        return {"xyz": c}
    for b in (0, B - 1):
        ref = oracle.forward_adjoint(qb[b], prob.xyz_fixed, prob.loads, so, cot)
        assert rel(fa.array("x").cpu().numpy()[b], ref["x_free"]) <= 1e-10
        assert rel(q_bar[b], ref["q_bar"]) <= 1e-8
    fa.close()
print("debug build ok")
"""


def test_small_cases_on_the_bounds_checked_build():
    if not os.path.exists(DEBUG_LIB):
        pytest.fail(f"{DEBUG_LIB} missing: run __graft_entry__.build()")
    env = dict(os.environ, FDM_B200_LIB=DEBUG_LIB, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "debug build ok" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
    assert "FDM_ASSERT" not in r.stdout + r.stderr
