#!/usr/bin/env python
"""
bench.py -- forward+adjoint FDM solves/s (BASELINE.json metric).

Headline workload (config.workload): BASELINE config 5 -- a batch of 4096 synthetic 30x30
pillow designs (841 free nodes, 1860 edges each, q_b = -2 + 0.1 (U-0.5), seed b), forward
solve + adjoint + gradients of L_b = 1/2||x_free - x_free0||^2.  Designs are independent, so
at N GPUs every rank owns 4096 designs (weak scaling, no data-path collective) and the scalar
loss is all-reduced with NCCL; `config5_fixed_global_batch` additionally reports the fixed
4096-design batch cut over the N ranks.  A "step" is one pass over a rank's batch.  The same
JSON line carries `kernels` (every kernel of the step timed alone against the HBM roofline)
and `single_mesh`: BASELINE config 4 (200x200 cable-net, one design), ms per forward+adjoint.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...          (N > 1: one rank per GPU)

Timing: W >= 3 warm-up steps; every timed step is bracketed by CUDA events on the launching
stream; an L2 flush (write of a 256 MiB buffer) runs between timed steps, outside the events;
barrier + synchronize on both sides of the timed region; max over ranks.
`--impl reference` times the CPU restatement of the reference path (oracle/: scipy SuperLU
forward + adjoint, exactly what sparse.py:147-149, 296 run) on the host cores -- jax is not
installable in this image, so the oracle port is the reference arm.
"""

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fwd+adjoint FDM solves/sec"
UNIT = "solves/s"
PILLOW_NX = 30
PER_GPU_BATCH = 4096          # designs per GPU (weak scaling); N=1 is BASELINE config 5
CONFIG5_BATCH = 4096          # BASELINE config 5's fixed global batch (reported beside the headline at N > 1)
GLOBAL_BATCH = PER_GPU_BATCH  # the CPU arm times a bounded sample of one GPU's share
CABLENET_NX = 200
E2E_CHUNKS = int(os.environ.get("FDM_BENCH_E2E_CHUNKS", "16"))
DEV_CHUNKS = int(os.environ.get("FDM_BENCH_DEV_CHUNKS", "2"))   # slices of the device-resident step (own streams)
# dram__bytes_read.sum + dram__bytes_write.sum of one chol_factor_mma_kernel launch (B = 4096), from
# profiles/r01k_chol_final_ncu.txt (0.325 GB read + 0.960 GB written)
NCU_TRAFFIC_CHOL_FWD = 1.295e9   # dram read+write of chol_factor_mma_kernel<0,4>, ncu --set full (profiles/r01n_chol_panel4_ncu.txt)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/fdm_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle) -- cpu_baseline leg and --impl reference
# --------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(kind, nx):
    import oracle
    from jax_fdm_b200 import datasets
    prob = datasets.pillow_problem(nx) if kind == "pillow" else datasets.cablenet_problem(nx)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    _CPU.update(prob=prob, so=so, oracle=oracle, datasets=datasets,
                x0=prob.xyz0[so.indices_free])


def _cpu_one(b):
    """forward + adjoint of design b on one core: 2 SuperLU spsolves + closed-form gradients."""
    o, prob, so = _CPU["oracle"], _CPU["prob"], _CPU["so"]
    if b is None:
        q = prob.q
    else:
        q = _CPU["datasets"].pillow_batch_q(prob.edges.shape[0], b, 1)[0]
    x0 = _CPU["x0"]

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    out = o.forward_adjoint(q, prob.xyz_fixed, prob.loads, so, cot)
    return float(0.5 * np.sum((out["x_free"] - x0) ** 2))


def cpu_time_designs(designs, procs):
    """seconds to run `designs` pillow designs with `procs` worker processes."""
    if procs <= 1:
        _cpu_init("pillow", PILLOW_NX)
        _cpu_one(0)
        t = time.perf_counter()
        for b in designs:
            _cpu_one(b)
        return time.perf_counter() - t
    import multiprocessing as mp
    with mp.get_context("fork").Pool(procs, initializer=_cpu_init, initargs=("pillow", PILLOW_NX)) as pool:
        pool.map(_cpu_one, range(procs))                       # warm every worker
        t = time.perf_counter()
        pool.map(_cpu_one, list(designs), chunksize=max(1, len(designs) // (4 * procs)))
        return time.perf_counter() - t


def cpu_time_cablenet(reps):
    _cpu_init("cablenet", CABLENET_NX)
    _cpu_one(None)
    t = time.perf_counter()
    for _ in range(reps):
        _cpu_one(None)
    return (time.perf_counter() - t) / reps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = len(os.sched_getaffinity(0))
    sample = min(GLOBAL_BATCH, max(cores * 64, 512))
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_time_designs(range(sample), cores)
        if i >= args.warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    value = sample / per_step
    ms_cable = cpu_time_cablenet(2) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3 * GLOBAL_BATCH / sample,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{GLOBAL_BATCH} x pillow {PILLOW_NX}x{PILLOW_NX} designs (841 free nodes, 1860 edges, "
                               "q_b = -2 + 0.1(U-0.5) seed b): forward solve + adjoint solve + gradients on the host "
                               "cores (rank 0 only; the reference has no multi-device path)",
                   "per_gpu_batch": GLOBAL_BATCH, "timed_sample_designs": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of {GLOBAL_BATCH} designs per step, one process per core; "
                                   "numpy/scipy restatement of the reference CPU path (SuperLU fwd + adjoint)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "single_mesh": {"workload": f"cable-net {CABLENET_NX}x{CABLENET_NX}", "ms_per_solve": ms_cable,
                        "solves_per_s": 1e3 / ms_cable, "cores": 1},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# ours
# --------------------------------------------------------------------------------------------
def timed_steps(fn, stream, steps, flush, dist=None):
    """per-step CUDA-event times (ms) of fn() on `stream`, L2 flushed between steps."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s0, s1 in ev:
        flush()                 # on `stream`: the start event below is ordered after it, no host sync needed
        s0.record(stream)       # (a host sync per step would add the ranks' host-side start skew to every
        fn()                    #  step that ends in a collective)
        s1.record(stream)
    stream.synchronize()
    return [a.elapsed_time(b) for a, b in ev]


def kernel_table(pipe, s, B, steps, flush, hbm):
    """Each C-ABI call of the step timed alone (CUDA events on pipe.stream, L2 flushed before every
    launch): algorithmic bytes per launch, time, achieved GB/s and fraction of the measured HBM peak."""
    import ctypes as C
    from jax_fdm_b200 import _lib as L
    lib = L.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(pipe.stream.cuda_stream)
    E, N, Nf, Ns, nnz = s.num_edges, s.num_nodes, s.num_free, s.num_supports, s.nnz
    fac = 8 * 1024 * ((Nf + 31) // 32)                       # factor columns, 32 doubles each, rows padded to 32
    Y = pipe.rhs.clone()
    stage_opts = {m: L.SolveOpts(pipe.solver_id, 0, 0.0, 0, 0, m) for m in (1, 4)}
    calls = [
        ("assemble (K1)", 8 * E + 8 * nnz + 24 * Nf,
         lambda: lib.fdm_assemble(s.plan, P(pipe.q), P(pipe.xyz_fixed), P(pipe.loads), P(pipe.K), P(pipe.rhs), B, E, 0, 0, st)),
        ("band Cholesky forward solve (K2b: the two launches below)", 8 * nnz + 48 * Nf + fac,
         lambda: lib.fdm_solve(s.plan, P(pipe.K), P(pipe.rhs), P(pipe.x), P(pipe.info_f), P(pipe.ws), pipe.ws.numel(), B,
                               C.byref(pipe.opts_f), st)),
        ("  chol_factor_mma_kernel (factorisation, fused forward substitution, diagonal blocks inverted)",
         8 * nnz + 24 * Nf + fac + fac // 8,
         lambda: lib.fdm_solve(s.plan, P(pipe.K), P(pipe.rhs), P(pipe.x), P(pipe.info_f), P(pipe.ws), pipe.ws.numel(), B,
                               C.byref(stage_opts[1]), st)),
        ("  chol_subst_kernel (backward substitution, factor streamed by TMA)", fac + fac // 8 + 24 * Nf,
         lambda: lib.fdm_solve(s.plan, P(pipe.K), P(pipe.rhs), P(pipe.x), P(pipe.info_f), P(pipe.ws), pipe.ws.numel(), B,
                               C.byref(stage_opts[4]), st)),
        ("positions", 24 * Nf + 24 * N,
         lambda: lib.fdm_positions(s.plan, P(pipe.x), P(pipe.xyz_fixed), P(pipe.xyz), B, 0, 0, st)),
        ("state (K4)", 8 * E + 24 * N + 40 * E + 24 * N,
         lambda: lib.fdm_state(s.plan, P(pipe.q), P(pipe.xyz), P(pipe.loads), P(pipe.vectors), P(pipe.lengths),
                               P(pipe.forces), P(pipe.residuals), B, E, 0, st)),
        ("loss + cotangent", 48 * Nf + 8,
         lambda: lib.fdm_loss_sqdist(P(pipe.x), P(pipe.x0), P(pipe.g), P(pipe.loss), Nf * 3, B, 0, st)),
        ("chol_subst_kernel, adjoint solve: forward + backward sweep over the stored factor", 2 * (fac + fac // 8) + 48 * Nf,
         lambda: lib.fdm_solve(s.plan, P(pipe.K), P(pipe.g), P(pipe.lam), P(pipe.info_a), P(pipe.ws), pipe.ws.numel(), B,
                               C.byref(pipe.opts_a), st)),
        ("adjoint gradients (K3)", 8 * E + 24 * N + 24 * Nf + 8 * E + 24 * N + 24 * Ns,
         lambda: lib.fdm_adjoint_grads(s.plan, P(pipe.q), P(pipe.xyz), P(pipe.lam), P(pipe.q_bar), P(pipe.loads_bar),
                                       P(pipe.xs_bar), B, E, 0, st)),
        ("SpMM K X (3 rhs)", 8 * nnz + 48 * Nf,
         lambda: lib.fdm_spmm(s.plan, P(pipe.K), P(pipe.x), P(Y), B, st)),
    ]
    out = []
    for name, per_design, fn in calls:
        ts = timed_steps(lambda: L.check(fn()), pipe.stream, steps, flush)
        ms = float(np.mean(ts))
        gbs = per_design * B / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "ms": ms, "algorithmic_bytes_per_launch": per_design * B,
                    "achieved_gbs": gbs, "frac_of_measured_hbm": gbs / hbm})
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    from jax_fdm_b200 import datasets
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.distributed import shard_range, total_loss
    from jax_fdm_b200.pipeline import ChunkedForwardAdjoint, ForwardAdjoint

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    B = PER_GPU_BATCH                                    # weak scaling: every rank owns 4096 designs
    steps, warmup = args.steps, max(3, args.warmup)
    hbm, hbm_src = peaks()

    # ---------------- headline: batched pillow designs ----------------
    prob = datasets.pillow_problem(PILLOW_NX)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=local)
    lo, hi = shard_range(B * world, rank, world)
    qb = datasets.pillow_batch_q(prob.edges.shape[0], lo, hi - lo)
    x0 = prob.xyz0[s.indices_free]
    # launch-by-launch table: the unfused sequence (assembly kernel, then the solve on K.data)
    pipe = ForwardAdjoint(s, batch=B, solver="cholesky", device=local, fused_assembly=False)
    pipe.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    e2e_pipe = ChunkedForwardAdjoint(s, B, chunks=E2E_CHUNKS, solver="cholesky", device=local)
    e2e_pipe.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    # the timed device-resident step: the batch in DEV_CHUNKS slices on their own streams (fork/join on
    # dev_pipe.stream), so one slice's bandwidth-bound substitutions overlap another's factorisation
    dev_pipe = ChunkedForwardAdjoint(s, B, chunks=DEV_CHUNKS, solver="cholesky", device=local)
    dev_pipe.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        with torch.cuda.stream(pipe.stream):
            flush_buf.fill_(1)

    total = torch.zeros(1, dtype=torch.float64, device=dev)

    def reduce_loss(p):
        if world > 1:
            with torch.cuda.stream(p.stream):
                total_loss(p.loss, out=total)             # local sum + one 8-byte NCCL all-reduce

    def flush_dev():
        with torch.cuda.stream(dev_pipe.stream):
            flush_buf.fill_(1)

    def step_dev():
        dev_pipe.step()
        reduce_loss(dev_pipe)

    def step_e2e():
        e2e_pipe()
        if world > 1:
            with torch.cuda.stream(pipe.stream):
                total_loss(e2e_pipe.loss, out=total)
            pipe.stream.synchronize()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    t_dev = timed_steps(step_dev, dev_pipe.stream, steps, flush_dev)
    barrier()
    wall = time.perf_counter() - wall0
    # end to end: pinned host q in, loss + q_bar out, every step
    for _ in range(3):
        step_e2e()
    barrier()
    t_e2e = []
    for _ in range(steps):
        flush()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_e2e()
        torch.cuda.synchronize(dev)
        t_e2e.append((time.perf_counter() - t0) * 1e3)
    barrier()
    # the same steps enqueued back to back (throughput mode for independent batches): every step still copies
    # its q in and its loss + q_bar out, but step k+1's copy-in overlaps step k's kernels and copy-out
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_pipe.enqueue()
    e2e_pipe.synchronize()
    t_pipelined = (time.perf_counter() - t0) * 1e3 / steps
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sum_dev = reduce_max(sum(t_dev))
    sum_e2e = reduce_max(sum(t_e2e))
    t_pipelined = reduce_max(t_pipelined)
    status_ok = all(bool((p_.info_f[:, 0] == 0).all().item() and (p_.info_a[:, 0] == 0).all().item())
                    for p_ in dev_pipe.parts)

    # diagnostic at N > 1: the same steps without the loss all-reduce (what the collective + the lock-step
    # it imposes on the ranks cost), and the spread of the per-rank times
    scaling_diag = None
    if world > 1:
        barrier()
        t_local = timed_steps(dev_pipe.step, dev_pipe.stream, steps, flush_dev)
        barrier()
        mine = torch.tensor([sum(t_dev) / steps, sum(t_local) / steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = torch.stack(allr).cpu().numpy()
        scaling_diag = {"ms_per_step_with_allreduce_per_rank": [float(v) for v in per_rank[:, 0]],
                        "ms_per_step_without_allreduce_per_rank": [float(v) for v in per_rank[:, 1]]}

    # every kernel of the step alone, for the roofline table (rank 0's numbers are reported)
    kernels = kernel_table(pipe, s, B, max(3, min(steps, 10)), flush, hbm)

    # BASELINE config 5 exactly (fixed global batch 4096 cut over the ranks), strong scaling, N > 1 only
    fixed = None
    if world > 1:
        lo5, hi5 = shard_range(CONFIG5_BATCH, rank, world)
        p5 = ForwardAdjoint(s, batch=hi5 - lo5, solver="cholesky", device=local)
        p5.set_inputs(datasets.pillow_batch_q(prob.edges.shape[0], lo5, hi5 - lo5), prob.xyz_fixed, prob.loads, x0)

        def step5():
            p5.step()
            reduce_loss(p5)
        for _ in range(warmup):
            step5()
        barrier()

        def flush5():
            with torch.cuda.stream(p5.stream):
                flush_buf.fill_(1)
        t5 = reduce_max(sum(timed_steps(step5, p5.stream, steps, flush5)))
        fixed = {"workload": f"{CONFIG5_BATCH} designs in total, {hi5 - lo5} per GPU", "scaling": "strong",
                 "ms_per_step": t5 / steps, "value": CONFIG5_BATCH / (t5 / steps * 1e-3), "unit": UNIT}

    # ---------------- second workload: one 200x200 cable-net (rank 0 only) ----------------
    single = None
    if rank == 0:
        cprob = datasets.cablenet_problem(CABLENET_NX)
        cs = eq.EquilibriumStructureSparse(cprob.nodes, cprob.edges, cprob.supports, device=local)
        cpipe = ForwardAdjoint(cs, batch=1, solver="auto", tol=1e-12, device=local)
        cpipe.set_inputs(cprob.q[None], cprob.xyz_fixed, cprob.loads, cprob.xyz0[cs.indices_free])
        for _ in range(3):
            cpipe.step()
        cpipe.stream.synchronize()
        k = max(3, min(steps, 10))

        def cflush():
            with torch.cuda.stream(cpipe.stream):
                flush_buf.fill_(1)
        tc = timed_steps(cpipe.step, cpipe.stream, k, cflush)
        inf_f, inf_a = cpipe.info_f.cpu().numpy()[0], cpipe.info_a.cpu().numpy()[0]
        single = {"workload": f"cable-net {CABLENET_NX}x{CABLENET_NX}, 4 corners fixed, q~U[0.5,2], fwd+adjoint",
                  "ms_per_solve": float(np.mean(tc)), "solves_per_s": 1e3 / float(np.mean(tc)),
                  "solver": "two-level PCG (persistent)" if cpipe.capturable else "Jacobi-PCG (multi-kernel)",
                  "iters_fwd": int(inf_f[1]), "iters_adj": int(inf_a[1]),
                  "relres_fwd": float(inf_f[2]), "relres_adj": float(inf_a[2]),
                  "status": [int(inf_f[0]), int(inf_a[0])]}

    if world > 1:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None

    dom = kernels[2]                                     # chol_factor_mma_kernel, the largest share of the step
    ms_per_step = sum_dev / steps
    value = B * world / (ms_per_step * 1e-3)
    e2e_value = B * world / (sum_e2e / steps * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{B} x pillow {PILLOW_NX}x{PILLOW_NX} designs per GPU (841 free nodes, 1860 edges, "
                               "q_b = -2 + 0.1(U-0.5) seed b): forward solve + state + loss + adjoint solve + gradients; "
                               "BASELINE config 5 at N=1",
                   "per_gpu_batch": B, "global_batch": B * world, "solver": "batched band Cholesky, one warp per design",
                   "l2": "flushed between timed steps (256 MiB write, outside the events)",
                   "device_step": f"{DEV_CHUNKS} slices of the batch on their own streams, forked from and joined to the "
                                  "timed stream (CUDA events on it bracket the whole step); the factorisation "
                                  "assembles its rows from q in place (fdm_solve_from_q), the `kernels` table times "
                                  "the unfused launches",
                   "parallelism": f"designs sharded x{world}, no data-path collective"
                                  + (", 8-byte loss all-reduce (NCCL) per step" if world > 1 else "")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_pipe.h2d_bytes * world,
                "d2h_bytes_per_step": e2e_pipe.d2h_bytes * world, "ms_per_step": sum_e2e / steps,
                "how": f"pinned host q -> device, step, loss + q_bar -> pinned host, {E2E_CHUNKS} chunks on "
                       "separate streams so copies overlap kernels; host wall clock, synchronised both sides of "
                       "EVERY step",
                "pipelined": {"value": B * world / (t_pipelined * 1e-3), "unit": UNIT, "ms_per_step": t_pipelined,
                              "how": "the same K steps enqueued back to back and synchronised once: step k+1's copy-in "
                                     "overlaps step k's kernels and copy-out (independent batches; an optimiser loop "
                                     "that needs step k's gradient before step k+1 gets the synchronous number)"}},
        "gpu_launches": dev_pipe.launches_per_step * steps,
        "roofline": {"kernel": dom["kernel"].strip(), "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": hbm,
                     "unit": "GB/s", "frac": dom["achieved_gbs"] / hbm, "traffic": NCU_TRAFFIC_CHOL_FWD,
                     "peak_source": hbm_src, "kernel_ms": dom["ms"],
                     "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                     "share_of_step": dom["ms"] / ms_per_step,
                     "note": "timed alone on K.data (template instance <FROM_Q=0>); inside the step the <FROM_Q=1> "
                             "instance runs, which reads q (15 KB per design) instead of K.data (33 KB)"},
        "kernels": kernels,
        "wall_s_timed_region": wall,
        "solve_status_ok": status_ok and (single is None or single["status"] == [0, 0]),
        "single_mesh": single,
    }
    if fixed:
        line["config5_fixed_global_batch"] = fixed
    if scaling_diag:
        line["scaling_diagnostic"] = scaling_diag
    if world == 1:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        sample = 2048
        cores = len(os.sched_getaffinity(0))
        dt = cpu_time_designs(range(sample), cores)
        line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{sample} of {B} designs, one process per core (scipy SuperLU is "
                                          "single-threaded); numpy/scipy restatement of sparse.py:147-149,296"}
    return line


class _StdoutToStderr:
    """Library banners (NCCL prints its version on stdout) must not pollute the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        with _StdoutToStderr():
            line = run_ours(args)
        if line is not None:
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
