#!/usr/bin/env python
"""Experiment driver: the goal kernel (fdm_goals_loss) alone on the 4096-design pillow batch with 841 NodePointGoals:
forward-only and reverse-mode launches timed with CUDA events."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_fdm_b200 import datasets, goals as G, losses as LS  # noqa: E402
from jax_fdm_b200.losses import _goals_call  # noqa: E402
import jax_fdm_b200.equilibrium as eq  # noqa: E402

B = int(os.environ.get("B", "4096"))
prob = datasets.pillow_problem(30)
s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
x0 = prob.xyz0[s.indices_free]
loss = LS.Loss(LS.SquaredError([G.NodePointGoal(int(prob.nodes[n]), t) for n, t in zip(s.indices_free, x0)], alpha=0.5))
dev = torch.device("cuda", 0)
prog = loss.program(s, dev)
N, E = s.num_nodes, s.num_edges
xyz = torch.randn(B, N, 3, dtype=torch.float64, device=dev)
lengths, forces = torch.rand(B, E, 1, dtype=torch.float64, device=dev), torch.rand(B, E, 1, dtype=torch.float64, device=dev)
res = torch.randn(B, N, 3, dtype=torch.float64, device=dev)
out = torch.empty(B, dtype=torch.float64, device=dev)
bars = [torch.empty_like(t) for t in (xyz, lengths, forces, res)]
cot = torch.ones(B, dtype=torch.float64, device=dev)
for name, fn in (("forward", lambda: _goals_call(s, prog, xyz, lengths, forces, res, out, None, [None] * 4)),
                 ("reverse", lambda: _goals_call(s, prog, xyz, lengths, forces, res, None, cot, bars))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(name, a.elapsed_time(b) / 10, "ms")
