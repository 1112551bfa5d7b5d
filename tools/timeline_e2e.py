#!/usr/bin/env python
"""Experiment driver: CUPTI timeline (torch.profiler) of one fdm_fa_run_host step: when each slice's copies and
kernels start and end, relative to the first device activity."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_fdm_b200 import datasets  # noqa: E402
import jax_fdm_b200.equilibrium as eq  # noqa: E402
from jax_fdm_b200.pipeline import NativeForwardAdjoint  # noqa: E402

B, chunks = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 16
prob = datasets.pillow_problem(30)
s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
fa = NativeForwardAdjoint(s, B, chunks=chunks, solver="cholesky")
fa.set_inputs(datasets.pillow_batch_q(prob.edges.shape[0], 0, B), prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
for _ in range(3):
    fa()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fa()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in evs)
rows = sorted(evs, key=lambda e: e.time_range.start)
print(f"{len(rows)} device activities; times in us from the first one")
for e in rows:
    name = e.name[:44]
    print(f"{(e.time_range.start - t0):9.1f} {(e.time_range.end - t0):9.1f}  {e.time_range.end - e.time_range.start:8.1f}  {name}")
