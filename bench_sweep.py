#!/usr/bin/env python
"""
Size sweep of the HBM-bound kernels (SURVEY.md §8d "size sweep"): assembly (K1), SpMM, state (K4),
adjoint gradients (K3), positions, on single cable-net meshes whose working sets grow past the
126 MB L2 (nx = 200 ... 2000) and on batches of 200x200 meshes.  Every launch is timed alone with
CUDA events on the launching stream, L2 flushed before it; achieved = algorithmic bytes / time
against the measured HBM peak (MEASURED_PEAKS.json).  Prints one JSON line per case.

    python bench_sweep.py [--nx 200 500 1000 2000] [--batch 1 64] [--reps 10]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, nargs="+", default=[200, 500, 1000, 2000])
    ap.add_argument("--batch", type=int, nargs="+", default=[1])
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from bench import peaks, timed_steps
    from jax_fdm_b200 import _lib as L
    from jax_fdm_b200 import datasets
    import jax_fdm_b200.equilibrium as eq

    dev = torch.device("cuda", 0)
    hbm, src = peaks()
    lib = L.lib()
    stream = torch.cuda.Stream(device=dev)
    st = C.c_void_p(stream.cuda_stream)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    P = lambda t: C.c_void_p(t.data_ptr())

    def flush():
        with torch.cuda.stream(stream):
            flush_buf.fill_(1)

    for nx in args.nx:
        t0 = time.perf_counter()
        prob = datasets.cablenet_problem(nx)
        s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
        t_plan = time.perf_counter() - t0
        E, N, Nf, Ns, nnz = s.num_edges, s.num_nodes, s.num_free, s.num_supports, s.nnz
        for B in args.batch:
            if B * (nnz * 8 + E * 80 + N * 150) > 60e9:
                continue
            z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
            rng = np.random.default_rng(1)
            q = torch.as_tensor(np.tile(prob.q, (B, 1)) * rng.uniform(0.9, 1.1, size=(B, 1)), device=dev)
            xs = torch.as_tensor(prob.xyz_fixed, device=dev)
            loads = torch.as_tensor(prob.loads, device=dev)
            xyz = torch.as_tensor(np.tile(prob.xyz0, (B, 1, 1)), device=dev).contiguous()
            K, rhs, x, y = z(B, nnz), z(B, Nf, 3), z(B, Nf, 3), z(B, Nf, 3)
            lam = torch.as_tensor(rng.standard_normal((B, Nf, 3)), device=dev)
            vec, lng, frc, res = z(B, E, 3), z(B, E), z(B, E), z(B, N, 3)
            qb, lb, xb = z(B, E), z(B, N, 3), z(B, Ns, 3)
            x.copy_(xyz[:, torch.as_tensor(s.indices_free, device=dev)])
            calls = [
                ("assemble (K1)", 16 * E + 12 * nnz + 52 * Nf + 24 * Ns,
                 lambda: lib.fdm_assemble(s.plan, P(q), P(xs), P(loads), P(K), P(rhs), B, E, 0, 0, st)),
                ("SpMM K X (3 rhs)", 12 * nnz + 52 * Nf,
                 lambda: lib.fdm_spmm(s.plan, P(K), P(x), P(y), B, st)),
                ("positions", 24 * Nf + 24 * N + 4 * N,
                 lambda: lib.fdm_positions(s.plan, P(x), P(xs), P(xyz), B, 0, 0, st)),
                ("state (K4)", 56 * E + 72 * N,
                 lambda: lib.fdm_state(s.plan, P(q), P(xyz), P(loads), P(vec), P(lng), P(frc), P(res), B, E, 0, st)),
                ("adjoint gradients (K3)", 24 * E + 24 * Nf + 48 * N + 24 * Ns,
                 lambda: lib.fdm_adjoint_grads(s.plan, P(q), P(xyz), P(lam), P(qb), P(lb), P(xb), B, E, 0, st)),
            ]
            out = {"workload": f"cable-net {nx}x{nx}", "batch": B, "nodes": N, "edges": E, "free": Nf, "nnz": nnz,
                   "plan_build_s": t_plan, "hbm_peak_gbs": hbm, "peak_source": src, "kernels": []}
            for name, per, fn in calls:
                # algorithmic bytes per design exactly as SURVEY.md §8d defines them (FP64 values + int32
                # index streams, every array touched once)
                for _ in range(3):
                    L.check(fn())
                ts = timed_steps(lambda: L.check(fn()), stream, args.reps, flush)
                ms = float(np.median(ts))
                gbs = per * B / (ms * 1e-3) / 1e9
                out["kernels"].append({"kernel": name, "ms": ms, "bytes": per * B, "achieved_gbs": gbs,
                                       "frac_of_measured_hbm": gbs / hbm})
            print(json.dumps(out), flush=True)
            del q, xyz, K, rhs, x, y, lam, vec, lng, frc, res, qb, lb, xb
            torch.cuda.empty_cache()
        del s


if __name__ == "__main__":
    main()
