#!/usr/bin/env python
"""Experiment driver: the factor kernel alone (fdm_solve_from_q, stages = 1) on B pillow 30x30 designs, timed with
CUDA events (L2 flushed) -- for ncu captures and A/B runs.   B=148 FDM_CHOL_FACWARPS=1: one lone warp per SM."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import timed_steps  # noqa: E402
from jax_fdm_b200 import _lib as L, datasets  # noqa: E402
import jax_fdm_b200.equilibrium as eq  # noqa: E402

B = int(os.environ.get("B", "4096"))
NX = int(os.environ.get("NX", "30"))
stages = int(os.environ.get("STAGES", "1"))
prob = datasets.pillow_problem(NX)
s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
dev = torch.device("cuda", 0)
T = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=dev)
q, xs, p = T(datasets.pillow_batch_q(prob.edges.shape[0], 0, B)), T(prob.xyz_fixed), T(prob.loads)
x = torch.empty(B, s.num_free, 3, dtype=torch.float64, device=dev)
info = torch.zeros(B, 8, dtype=torch.float64, device=dev)
lib = L.lib()
ws = torch.empty(int(lib.fdm_workspace_bytes(s.plan, B, 3)), dtype=torch.uint8, device=dev)
opts = L.SolveOpts(3, 0, 0.0, 0, 0, stages, 0)
stream = torch.cuda.Stream(device=dev)
st = C.c_void_p(stream.cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def flush():
    with torch.cuda.stream(stream):
        flush_buf.fill_(1)


def fn():
    L.check(lib.fdm_solve_from_q(s.plan, P(q), P(xs), P(p), P(x), P(info), P(ws), ws.numel(), B, s.num_edges, 0, 0,
                                 C.byref(opts), st))


for _ in range(3):
    fn()
stream.synchronize()
ts = timed_steps(fn, stream, int(os.environ.get("REPS", "10")), flush)
print(f"B={B} nx={NX} stages={stages} hbw={s.half_bandwidth}: {np.median(ts):.4f} ms (min {min(ts):.4f})")
