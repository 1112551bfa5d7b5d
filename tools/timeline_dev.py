#!/usr/bin/env python
"""Experiment driver: CUPTI timeline (torch.profiler) of one device-resident fdm_fa_step_device step at a given batch
and slice count: when every kernel of the step starts and ends. `timeline_dev.py B chunks`."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_fdm_b200 import datasets  # noqa: E402
import jax_fdm_b200.equilibrium as eq  # noqa: E402
from jax_fdm_b200.pipeline import NativeForwardAdjoint  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
prob = datasets.pillow_problem(30)
s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
fa = NativeForwardAdjoint(s, B, chunks=chunks, solver="cholesky")
fa.set_inputs(datasets.pillow_batch_q(prob.edges.shape[0], 0, B), prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    fa.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(10):
    flush.fill_(1)
    fa.stream.wait_stream(torch.cuda.current_stream())
    e0.record(fa.stream)
    fa.step()
    e1.record(fa.stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"B={B} chunks={chunks}: step {sum(ts) / len(ts):.4f} ms (min {min(ts):.4f})")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    flush.fill_(1)
    torch.cuda.synchronize()
    fa.step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "fill" not in e.name.lower()]
t0 = min(e.time_range.start for e in evs)
rows = sorted(evs, key=lambda e: e.time_range.start)
print(f"{len(rows)} device activities; times in us from the first one")
for e in rows:
    print(f"{(e.time_range.start - t0):9.1f} {(e.time_range.end - t0):9.1f}  {e.time_range.end - e.time_range.start:8.1f}  {e.name[:70]}")
