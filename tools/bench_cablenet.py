#!/usr/bin/env python
"""Experiment driver: forward + adjoint of the 200x200 cable-net (BASELINE config 4) through fdm_fa_step_device, for the
persistent-PCG switches FDM_PCG2_CHEB_K / FDM_PCG2_CHEB_ALPHA / FDM_PCG2_PARTS (read once per process)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import timed_steps  # noqa: E402
from jax_fdm_b200 import datasets  # noqa: E402
import jax_fdm_b200.equilibrium as eq  # noqa: E402
from jax_fdm_b200.pipeline import NativeForwardAdjoint  # noqa: E402

nx = int(os.environ.get("NX", "200"))
prob = datasets.cablenet_problem(nx)
s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
fa = NativeForwardAdjoint(s, 1, solver="auto", tol=1e-12)
fa.set_inputs(prob.q[None], prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
for _ in range(3):
    fa.step()
fa.stream.synchronize()
buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")


def flush():
    with torch.cuda.stream(fa.stream):
        buf.fill_(1)


ts = timed_steps(fa.step, fa.stream, 8, flush)
f, a = fa.array("info_f").cpu().numpy()[0], fa.array("info_a").cpu().numpy()[0]
env = {k: v for k, v in os.environ.items() if k.startswith("FDM_")}
print(f"nx={nx} {np.median(ts):.3f} ms  iters {int(f[1])}+{int(a[1])}  relres {f[2]:.1e} {a[2]:.1e}  status {int(f[0])},{int(a[0])}  {env}", flush=True)
