#!/usr/bin/env python
"""Experiment driver: end-to-end time (host q in -> host loss + q_bar out, synchronised every step) of the 4096-design
pillow batch against the number of slices of fdm_fa_run_host.  One line per setting."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_fdm_b200 import datasets  # noqa: E402
import jax_fdm_b200.equilibrium as eq  # noqa: E402
from jax_fdm_b200.pipeline import NativeForwardAdjoint  # noqa: E402

B = 4096
prob = datasets.pillow_problem(30)
s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
for chunks in [int(c) for c in (sys.argv[1:] or "2 4 6 8 12 16 24 32".split())]:
    fa = NativeForwardAdjoint(s, B, chunks=chunks, solver="cholesky")
    fa.set_inputs(qb, prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
    for _ in range(3):
        fa()
    ts = []
    for _ in range(12):
        flush_buf.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fa()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"chunks={chunks:3d}: e2e median {np.median(ts):.3f} ms  min {min(ts):.3f} ms", flush=True)
    fa.close()
    del fa
    torch.cuda.empty_cache()
