#!/usr/bin/env python
"""Experiment driver: kernel-level breakdown (torch.profiler, CUDA activity) of one model-API forward + backward on
the 4096-design pillow batch -- where the API path's time goes compared with the library-owned pipeline."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_fdm_b200 import datasets, goals as G, losses as LS  # noqa: E402
import jax_fdm_b200.equilibrium as eq  # noqa: E402

B = int(os.environ.get("B", "4096"))
prob = datasets.pillow_problem(30)
s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
x0 = prob.xyz0[s.indices_free]
loss = LS.Loss(LS.SquaredError([G.NodePointGoal(int(prob.nodes[n]), t) for n, t in zip(s.indices_free, x0)], alpha=0.5))
model = eq.EquilibriumModelSparse(tmax=1)
dev = torch.device("cuda", 0)
q_t = torch.as_tensor(datasets.pillow_batch_q(prob.edges.shape[0], 0, B), device=dev)
xs_t, p_t = torch.as_tensor(prob.xyz_fixed, device=dev), torch.as_tensor(prob.loads, device=dev)


def step():
    q = q_t.detach().requires_grad_(True)
    val = loss(eq.EquilibriumParametersState(q, xs_t, eq.LoadState(p_t, 0.0, 0.0)), model, s)
    val.backward(torch.ones_like(val))


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
