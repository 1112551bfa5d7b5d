#!/usr/bin/env python
"""Experiment driver: the device-resident step (fdm_fa_step_device) and the host step (fdm_fa_run_host) at a list of
slice counts, in ONE process (the library's A/B switches -- FDM_FA_PRIORITY, FDM_CHOL_SUBWARPS, FDM_CHOL_FACWARPS,
FDM_FA_TW_FIRST -- are read once per process).  `sweep_dev_slices.py B dev=2,4,8 e2e=16`."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_fdm_b200 import datasets  # noqa: E402
import jax_fdm_b200.equilibrium as eq  # noqa: E402
from jax_fdm_b200.pipeline import NativeForwardAdjoint  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lists = {"dev": [], "e2e": []}
for a in sys.argv[2:]:
    k, v = a.split("=")
    lists[k] = [int(x) for x in v.split(",") if x]
prob = datasets.pillow_problem(30)
s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tag = " ".join(f"{k}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("FDM_"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = None
for kind in ("dev", "e2e"):
    for chunks in lists[kind]:
        fa = NativeForwardAdjoint(s, B, chunks=chunks, solver="cholesky")
        fa.set_inputs(qb, prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
        run = fa.step if kind == "dev" else fa
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.fill_(1)
            if kind == "dev":
                fa.stream.wait_stream(torch.cuda.current_stream())
                e0.record(fa.stream)
                fa.step()
                e1.record(fa.stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            else:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fa()
                ts.append((time.perf_counter() - t0) * 1e3)
        qbar = fa.array("q_bar").clone()
        if ref is None:
            ref = qbar
        same = bool(torch.equal(ref, qbar))
        print(f"[{tag}] B={B} {kind} chunks={chunks}: {sum(ts) / len(ts):.4f} ms (min {min(ts):.4f}) q_bar identical to first: {same}", flush=True)
        fa.close()
