#!/usr/bin/env python
"""Experiment driver (not part of the product): device-resident step time of the 4096-design pillow batch for a
grid of co-residency settings -- slices of the batch, designs per CTA of the factor / substitution kernels, extra
shared memory that caps the factor kernel's CTAs per SM.  Every setting runs in its own process (the library reads
its A/B switches once).  One JSON line per setting on stdout.

    python tools/sweep_overlap.py            # the grid
    python tools/sweep_overlap.py --one      # measure with the current environment (used by the grid)
"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torch
    from bench import timed_steps
    from jax_fdm_b200 import datasets
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import NativeForwardAdjoint
    B = int(os.environ.get("SWEEP_B", "4096"))
    chunks = int(os.environ.get("SWEEP_CHUNKS", "2"))
    prob = datasets.pillow_problem(30)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    qb = datasets.pillow_batch_q(prob.edges.shape[0], 0, B)
    fa = NativeForwardAdjoint(s, B, chunks=chunks, solver="cholesky")
    fa.set_inputs(qb, prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")

    def flush():
        with torch.cuda.stream(fa.stream):
            flush_buf.fill_(1)
    for _ in range(4):
        fa.step()
    fa.stream.synchronize()
    ts = timed_steps(fa.step, fa.stream, 12, flush)
    ok = bool((fa.array("info_f")[:, 0] == 0).all().item())
    print(json.dumps({"B": B, "chunks": chunks, "ms": float(np.median(ts)), "min": float(min(ts)), "ok": ok,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("FDM_")}}), flush=True)


def grid():
    cfgs = []
    for chunks, pad, sub, fac in itertools.product([1, 2, 4, 8], [0, 32768], [11, 5], [8]):
        cfgs.append((chunks, pad, sub, fac))
    cfgs += [(4, 36864, 3, 4), (8, 36864, 3, 4), (4, 0, 5, 4), (8, 0, 11, 4), (16, 32768, 5, 8), (16, 0, 11, 8)]
    for chunks, pad, sub, fac in cfgs:
        env = dict(os.environ, SWEEP_CHUNKS=str(chunks), FDM_CHOL_FAC_SMEM_PAD=str(pad), FDM_CHOL_SUBWARPS=str(sub),
                   FDM_CHOL_FACWARPS=str(fac))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, capture_output=True, text=True)
        sys.stdout.write(r.stdout if r.returncode == 0 else json.dumps({"cfg": [chunks, pad, sub, fac], "error": r.stderr[-300:]}) + "\n")
        sys.stdout.flush()


if __name__ == "__main__":
    one() if "--one" in sys.argv else grid()
