#!/usr/bin/env python
"""
bench.py -- forward+adjoint FDM solves/s (BASELINE.json metric).

Headline workload (config.workload): BASELINE config 5 -- a batch of 4096 synthetic 30x30
pillow designs (841 free nodes, 1860 edges each, q_b = -2 + 0.1 (U-0.5), seed b), forward
solve + state + adjoint + gradients of L_b = 1/2||x_free - x_free0||^2.  Designs are independent, so
at N GPUs every rank owns 4096 designs (weak scaling, no data-path collective) and the scalar
loss is all-reduced with NCCL; `config5_fixed_global_batch` reports, at every N, BASELINE's fixed
4096-design batch cut over the N ranks (strong scaling).  A "step" is one pass over a rank's batch.

One JSON line.  Beside the contract keys it carries
  value / ms_per_step      the device-resident step of the library-owned pipeline (fdm_fa_step_device), CUDA events
  e2e                      the same through fdm_fa_run_host: pinned host q in, host loss + q_bar out, every step;
                           + `copy_roofline` (the same copies without kernels: what the links deliver),
                           `loss_only_readback`, `pipelined`
  roofline                 the dominant kernel -- the factor kernel INSTANCE THE STEP LAUNCHES, timed alone through
                           fdm_solve_from_q(stages = 1); `traffic` and `counters` from the committed ncu capture
                           (profiles/ncu_counters.json); `strict`: SURVEY §8d's compulsory bytes only
  step_roofline            SURVEY §8d's compulsory bytes of the whole step over the timed step
  kernels                  every C-ABI call of the step timed alone (+ write-adjusted roofline, `stream_bandwidths`)
  sweep_large_mesh         the HBM-bound kernels on one 2000x2000 mesh (working set far beyond L2)
  api_path                 the public model API: EquilibriumModelSparse.__call__ + losses.Loss + .backward()
  parity                   sampled designs of the timed batch against the CPU oracle (1e-10 / 1e-8)
  single_mesh              BASELINE config 4 (200x200 cable-net, one design), ms per forward+adjoint
  cpu_baseline             the oracle port on the host cores (N = 1)

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...          (N > 1: one rank per GPU)

Timing: W >= 3 warm-up steps; every timed step is bracketed by CUDA events on the launching
stream; an L2 flush (write of a 256 MiB buffer) runs between timed steps, outside the events;
barrier + synchronize on both sides of the timed region; max over ranks.
`--impl reference` times the CPU restatement of the reference path (oracle/: scipy SuperLU
forward + adjoint, exactly what sparse.py:147-149, 296 run) on ALL 4096 designs per step on the host
cores -- jax is not installable in this image, so the oracle port is the reference arm.
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
DEV_CHUNKS = int(os.environ.get("FDM_BENCH_DEV_CHUNKS", "8"))   # slices of the device-resident step (own streams)
SWEEP_NX = 2000                # one mesh far beyond L2 for the HBM-bound kernels' fractions (SURVEY.md §8d size sweep)
WORKLOAD = (f"{PER_GPU_BATCH} x pillow {PILLOW_NX}x{PILLOW_NX} designs per GPU (841 free nodes, 1860 edges, "
            "q_b = -2 + 0.1(U-0.5) seed b): forward solve + state + loss + adjoint solve + gradients; "
            "BASELINE config 5 at N=1")


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
    """The reference arm: the CPU restatement of the reference path on ALL 4096 designs of the workload per step
    (one process per host core), the step's own wall time as ms_per_step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = len(os.sched_getaffinity(0))
    sample = GLOBAL_BATCH
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
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": GLOBAL_BATCH, "global_batch": GLOBAL_BATCH,
                   "where": "host cores, rank 0 only (the reference has no multi-device path): scipy SuperLU forward "
                            "+ adjoint solve + closed-form gradients per design, one process per core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"all {sample} designs of the workload per step, one process per core; "
                                   "numpy/scipy restatement of the reference CPU path (sparse.py:147-149, 296)"},
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


def stream_bandwidths(dev, stream):
    """What this GPU moves for a pure WRITE stream and for a copy, measured here with torch fills / copies of 1 GiB
    (CUDA events on `stream`): MEASURED_PEAKS.json's number is a copy (read + write).  Round 1 blamed the assembly
    kernel's fraction on a write stream at half the copy rate; measured at this size the write stream is no slower
    than the copy, so `frac_of_write_adjusted_roofline` (time >= max(all bytes / copy peak, written bytes / write
    peak)) equals the plain fraction and the blame does not hold."""
    import torch
    n = 1 << 27                                               # 1 GiB of doubles
    a = torch.empty(n, dtype=torch.float64, device=dev)
    b = torch.empty(n, dtype=torch.float64, device=dev)
    out = {}
    with torch.cuda.stream(stream):
        for name, fn, nbytes in (("write_gbs", lambda: a.fill_(1.0), 8 * n), ("copy_gbs", lambda: b.copy_(a), 16 * n)):
            best = float("inf")
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                stream.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name] = nbytes / (best * 1e-3) / 1e9
    del a, b
    torch.cuda.empty_cache()
    return out


def ncu_counters():
    """Counters of the dominant kernel from the committed ncu capture (profiles/ncu_counters.json, written by
    profiles/summarize.py from `ncu --set full` of the same instance the step launches)."""
    path = os.path.join(ROOT, "profiles", "ncu_counters.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def strict_bytes(s):
    """SURVEY.md §8d's compulsory traffic per design of the whole forward + adjoint step (inputs read once,
    outputs written once; the factor, intermediate right-hand sides and the state re-reads are NOT compulsory):
    read q, write x_free, xyz and the state, write q_bar, loads_bar, xyz_fixed_bar, loss."""
    E, N, Nf, Ns = s.num_edges, s.num_nodes, s.num_free, s.num_supports
    fwd = 8 * E + 24 * Nf + 24 * N + (24 * E + 8 * E + 8 * E + 24 * N)
    rev = 8 * E + 24 * N + 24 * Ns + 8
    return fwd + rev


def kernel_table(pipe, s, B, steps, flush, hbm, write_gbs=None):
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
    stage_opts = {m: L.SolveOpts(pipe.solver_id, 0, 0.0, 0, 0, m, 0) for m in (1, 4)}
    # bytes WRITTEN per design by each call (a mostly-writing kernel is bounded by the write stream, see stream_bandwidths)
    written = {"assemble (K1)": 8 * nnz + 24 * Nf, "positions": 24 * N, "state (K4)": 40 * E + 24 * N,
               "loss + cotangent": 24 * Nf + 8, "adjoint gradients (K3)": 8 * E + 24 * N + 24 * Ns, "SpMM K X (3 rhs)": 24 * Nf}
    calls = [
        ("assemble (K1)", 8 * E + 8 * nnz + 24 * Nf,
         lambda: lib.fdm_assemble(s.plan, P(pipe.q), P(pipe.xyz_fixed), P(pipe.loads), P(pipe.K), P(pipe.rhs), B, E, 0, 0, st)),
        ("chol_factor_mma_kernel<FROM_Q=1,PW=4> (the instance the step launches: rows assembled from q, factorisation, "
         "fused forward substitution, diagonal blocks inverted)", 8 * E + fac + fac // 8,
         lambda: lib.fdm_solve_from_q(s.plan, P(pipe.q), P(pipe.xyz_fixed), P(pipe.loads), P(pipe.x), P(pipe.info_f), P(pipe.ws),
                                      pipe.ws.numel(), B, E, 0, 0, C.byref(stage_opts[1]), st)),
        ("chol_factor_mma_kernel<FROM_Q=0,PW=4> (the same on K.data)", 8 * nnz + 24 * Nf + fac + fac // 8,
         lambda: lib.fdm_solve(s.plan, P(pipe.K), P(pipe.rhs), P(pipe.x), P(pipe.info_f), P(pipe.ws), pipe.ws.numel(), B,
                               C.byref(stage_opts[1]), st)),
        ("chol_subst_kernel (backward substitution, factor streamed by TMA)", fac + fac // 8 + 24 * Nf,
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
        for _ in range(2):                                   # warm-up: lazy module load, function attributes
            L.check(fn())
        ts = timed_steps(lambda: L.check(fn()), pipe.stream, steps, flush)
        ms = float(np.mean(ts))
        gbs = per_design * B / (ms * 1e-3) / 1e9
        row = {"kernel": name, "ms": ms, "algorithmic_bytes_per_launch": per_design * B,
               "achieved_gbs": gbs, "frac_of_measured_hbm": gbs / hbm}
        if write_gbs and name in written:
            # roofline with the write stream accounted for: time >= max(all bytes / copy peak, written bytes / write peak)
            floor_ms = max(per_design * B / hbm, written[name] * B / write_gbs) / 1e6
            row["write_bytes_per_launch"] = written[name] * B
            row["frac_of_write_adjusted_roofline"] = floor_ms / ms
        out.append(row)
    return out


def sweep_large_mesh(nx, steps, flush_buf, hbm):
    """The HBM-bound kernels (assembly, SpMM, state, gradients) on ONE nx x nx cable-net whose working set is far
    beyond L2 -- SURVEY.md §8d's size sweep, the point where the >= 60 % of HBM target is judged."""
    import ctypes as C
    import torch
    from jax_fdm_b200 import _lib as L
    from jax_fdm_b200 import datasets
    import jax_fdm_b200.equilibrium as eq
    lib = L.lib()
    dev = flush_buf.device
    prob = datasets.cablenet_problem(nx)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=dev.index)
    E, N, Nf, Ns, nnz = s.num_edges, s.num_nodes, s.num_free, s.num_supports, s.nnz
    stream = torch.cuda.Stream(device=dev)
    st = C.c_void_p(stream.cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())
    z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
    q, xs, loads = (torch.as_tensor(a, device=dev) for a in (prob.q, prob.xyz_fixed, prob.loads))
    xyz = torch.as_tensor(prob.xyz0, device=dev).contiguous()
    K, rhs, y = z(nnz), z(Nf, 3), z(Nf, 3)
    x = xyz[torch.as_tensor(s.indices_free, device=dev)].contiguous()
    lam = torch.as_tensor(np.random.default_rng(1).standard_normal((Nf, 3)), device=dev)
    vec, lng, frc, res, qb, lb, xb = z(E, 3), z(E), z(E), z(N, 3), z(E), z(N, 3), z(Ns, 3)

    def flush():
        with torch.cuda.stream(stream):
            flush_buf.fill_(1)
    calls = [
        ("assemble (K1)", 16 * E + 12 * nnz + 52 * Nf + 24 * Ns,
         lambda: lib.fdm_assemble(s.plan, P(q), P(xs), P(loads), P(K), P(rhs), 1, E, 0, 0, st)),
        ("SpMM K X (3 rhs)", 12 * nnz + 52 * Nf, lambda: lib.fdm_spmm(s.plan, P(K), P(x), P(y), 1, st)),
        ("state (K4)", 56 * E + 72 * N,
         lambda: lib.fdm_state(s.plan, P(q), P(xyz), P(loads), P(vec), P(lng), P(frc), P(res), 1, E, 0, st)),
        ("adjoint gradients (K3)", 24 * E + 24 * Nf + 48 * N + 24 * Ns,
         lambda: lib.fdm_adjoint_grads(s.plan, P(q), P(xyz), P(lam), P(qb), P(lb), P(xb), 1, E, 0, st)),
    ]
    out = []
    for name, per, fn in calls:
        for _ in range(3):
            L.check(fn())
        ms = float(np.median(timed_steps(lambda: L.check(fn()), stream, steps, flush)))
        gbs = per / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "ms": ms, "algorithmic_bytes_per_launch": per, "achieved_gbs": gbs,
                    "frac_of_measured_hbm": gbs / hbm})
    return {"workload": f"one cable-net {nx}x{nx} ({N} nodes, {E} edges): working set beyond L2", "kernels": out}


def parity_sample(fa, prob, so, qb, x0, designs):
    """Sampled designs of the timed batch against the CPU oracle (the checker, never the thing measured)."""
    import oracle
    x = fa.array("x").cpu().numpy()
    q_bar = fa.array("q_bar").cpu().numpy()
    loss = fa.array("loss").cpu().numpy()

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    ex = eq_ = el = 0.0
    for b in designs:
        ref = oracle.forward_adjoint(qb[b], prob.xyz_fixed, prob.loads, so, cot)
        lref = 0.5 * np.sum((ref["x_free"] - x0) ** 2)
        ex, eq_ = max(ex, rel(x[b], ref["x_free"])), max(eq_, rel(q_bar[b], ref["q_bar"]))
        el = max(el, abs(loss[b] - lref) / lref)
    return {"designs_checked": len(designs), "designs": [int(b) for b in designs], "max_rel_err_x_free": ex,
            "max_rel_err_loss": el, "max_rel_err_q_bar": eq_, "tolerance": {"x_free": 1e-10, "q_bar": 1e-8},
            "ok": bool(ex <= 1e-10 and el <= 1e-10 and eq_ <= 1e-8), "against": "oracle/ (numpy/scipy SuperLU)"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from jax_fdm_b200 import datasets
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.distributed import shard_range, total_loss
    from jax_fdm_b200.pipeline import ForwardAdjoint, NativeForwardAdjoint

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
    # the timed device-resident step: the library-owned pipeline (fdm_fa_*), the batch in DEV_CHUNKS slices on their
    # own streams inside ONE CUDA graph (fork / join), so one slice's bandwidth-bound substitutions overlap
    # another's factorisation; one ctypes call per step
    dev_fa = NativeForwardAdjoint(s, B, chunks=DEV_CHUNKS, solver="cholesky", device=local)
    dev_fa.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    # end to end: host q in, host loss + q_bar out through the same object with more, smaller slices
    e2e_fa = NativeForwardAdjoint(s, B, chunks=E2E_CHUNKS, solver="cholesky", device=local)
    e2e_fa.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    main_stream = dev_fa.stream

    def flush_on(stream):
        def f():
            with torch.cuda.stream(stream):
                flush_buf.fill_(1)
        return f

    flush_dev = flush_on(main_stream)
    # for the per-kernel table: the same write, followed by a READ of another buffer larger than L2 -- the write alone
    # leaves L2 full of DIRTY lines whose write-back (up to 126 MB of DRAM traffic that is not the kernel's) lands inside
    # the timed launch; after the read L2 is cold AND clean.  The timed STEP keeps the plain write flush.
    drain_buf = torch.ones(48 << 20, dtype=torch.float64, device=dev)         # 384 MB
    drain_out = torch.zeros(1, dtype=torch.float64, device=dev)

    def flush_clean_on(stream):
        def f():
            with torch.cuda.stream(stream):
                flush_buf.fill_(1)
                torch.sum(drain_buf, dim=(0,), keepdim=True, out=drain_out)
        return f
    total = torch.zeros(1, dtype=torch.float64, device=dev)

    def reduce_loss(loss_vec, stream):
        if world > 1:
            with torch.cuda.stream(stream):
                total_loss(loss_vec, out=total)           # fdm_batch_sum + one 8-byte NCCL all-reduce
    dev_loss, e2e_loss = dev_fa.array("loss"), e2e_fa.array("loss")

    def step_dev():
        dev_fa.step(main_stream)
        reduce_loss(dev_loss, main_stream)

    def step_e2e():
        e2e_fa()
        if world > 1:
            reduce_loss(e2e_loss, main_stream)
            main_stream.synchronize()

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
    t_dev = timed_steps(step_dev, main_stream, steps, flush_dev)
    barrier()
    wall = time.perf_counter() - wall0
    # end to end: pinned host q in, loss + q_bar out, every step
    for _ in range(3):
        step_e2e()
    barrier()
    t_e2e = []
    for _ in range(steps):
        flush_dev()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_e2e()
        t_e2e.append((time.perf_counter() - t0) * 1e3)
    barrier()
    # the same steps enqueued back to back (throughput mode for independent batches): every step still copies
    # its q in and its loss + q_bar out, but step k+1's copy-in overlaps step k's kernels and copy-out
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_fa.enqueue()
    e2e_fa.synchronize()
    t_pipelined = (time.perf_counter() - t0) * 1e3 / steps
    barrier()
    # the same call when the gradient stays on the device (an optimiser on the GPU): only the losses come back
    t_lossonly = []
    for _ in range(steps):
        flush_dev()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_fa.enqueue(want_q_bar=False)
        e2e_fa.synchronize()
        t_lossonly.append((time.perf_counter() - t0) * 1e3)
    barrier()
    # what the link alone can do for this call: the same copies, same slices and streams, no kernels
    t_copy = []
    for _ in range(steps):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_fa.enqueue_copies_only()
        e2e_fa.synchronize()
        t_copy.append((time.perf_counter() - t0) * 1e3)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sum_dev = reduce_max(sum(t_dev))
    sum_e2e = reduce_max(sum(t_e2e))
    sum_copy = reduce_max(sum(t_copy))
    sum_lossonly = reduce_max(sum(t_lossonly))
    t_pipelined = reduce_max(t_pipelined)
    dev_fa.step(main_stream)                              # leave the device arrays of a plain step for the checks below
    main_stream.synchronize()
    status_ok = bool((dev_fa.array("info_f")[:, 0] == 0).all().item() and (dev_fa.array("info_a")[:, 0] == 0).all().item())

    # diagnostic at N > 1: the same steps without the loss all-reduce (what the collective + the lock-step
    # it imposes on the ranks cost), and the spread of the per-rank times
    scaling_diag = None
    if world > 1:
        barrier()
        t_local = timed_steps(lambda: dev_fa.step(main_stream), main_stream, steps, flush_dev)
        barrier()
        mine = torch.tensor([sum(t_dev) / steps, sum(t_local) / steps, sum(t_e2e) / steps, sum(t_copy) / steps],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = torch.stack(allr).cpu().numpy()
        scaling_diag = {"ms_per_step_with_allreduce_per_rank": [float(v) for v in per_rank[:, 0]],
                        "ms_per_step_without_allreduce_per_rank": [float(v) for v in per_rank[:, 1]],
                        "e2e_ms_per_step_per_rank": [float(v) for v in per_rank[:, 2]],
                        "copies_only_ms_per_step_per_rank": [float(v) for v in per_rank[:, 3]]}

    # BASELINE config 5 exactly (fixed global batch 4096 cut over the ranks), strong scaling -- at every N
    lo5, hi5 = shard_range(CONFIG5_BATCH, rank, world)
    if world > 1:
        p5 = NativeForwardAdjoint(s, hi5 - lo5, chunks="auto", solver="cholesky", device=local)   # slices of ~512 designs
        p5.set_inputs(datasets.pillow_batch_q(prob.edges.shape[0], lo5, hi5 - lo5), prob.xyz_fixed, prob.loads, x0)
        p5_loss = p5.array("loss")

        def step5():
            p5.step(main_stream)
            reduce_loss(p5_loss, main_stream)
        for _ in range(warmup):
            step5()
        barrier()
        t5 = reduce_max(sum(timed_steps(step5, main_stream, steps, flush_dev)))
        p5.close()
    else:
        t5 = sum_dev
    fixed = {"workload": f"{CONFIG5_BATCH} designs in total, {hi5 - lo5} per GPU", "scaling": "strong",
             "ms_per_step": t5 / steps, "value": CONFIG5_BATCH / (t5 / steps * 1e-3), "unit": UNIT,
             "slices_per_gpu": max(1, (hi5 - lo5) // 512) if world > 1 else DEV_CHUNKS}

    line_extra = {}
    if rank == 0:
        # ---- parity at the timed size: sampled designs of rank 0's shard against the oracle
        import oracle
        so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
        rng = np.random.default_rng(0)
        sample = sorted({0, 1, B // 2 - 1, B // 2, 2367, 2368, B - 2, B - 1} | set(int(b) for b in rng.choice(B, 8, replace=False)))
        line_extra["parity"] = parity_sample(dev_fa, prob, so, qb, x0, sample)

        # ---- the public model API on the same batch: EquilibriumModelSparse.__call__ + Loss + backward (autograd
        #      bookkeeping, per-call allocations, no CUDA graph): what API-level overhead costs
        from jax_fdm_b200 import goals as G, losses as LS
        free = s.indices_free
        api_loss = LS.Loss(LS.SquaredError([G.NodePointGoal(int(prob.nodes[n]), t) for n, t in zip(free, x0)], alpha=0.5))
        model = eq.EquilibriumModelSparse(tmax=1)
        q_t = torch.as_tensor(qb, device=dev)
        xs_t, p_t = torch.as_tensor(prob.xyz_fixed, device=dev), torch.as_tensor(prob.loads, device=dev)
        api_out = {}

        def api_step():
            q_leaf = q_t.detach().requires_grad_(True)
            val = api_loss(eq.EquilibriumParametersState(q_leaf, xs_t, eq.LoadState(p_t, 0.0, 0.0)), model, s)
            val.backward(torch.ones_like(val))
            api_out["loss"], api_out["q_bar"] = val.detach(), q_leaf.grad
        cur = torch.cuda.current_stream(dev)
        for _ in range(3):
            api_step()
        t_api = timed_steps(api_step, cur, max(3, min(steps, 10)), flush_on(cur))
        ref_loss, ref_qbar = dev_fa.array("loss"), dev_fa.array("q_bar")
        line_extra["api_path"] = {
            "ms_per_step": float(np.mean(t_api)), "value": B / (float(np.mean(t_api)) * 1e-3), "unit": UNIT,
            "how": "EquilibriumModelSparse(tmax=1)(params, structure) + losses.Loss(SquaredError(NodePointGoal x "
                   f"{len(free)}, alpha=0.5)) + .backward() on the same {B} designs, device-resident, torch autograd "
                   "nodes over the same kernels, per-call allocations, no CUDA graph; a loss of node goals reads positions only, "
                   "so the state kernel and its VJP are not launched on this route",
            "max_rel_diff_loss_vs_pipeline": float(((api_out["loss"] - ref_loss).abs() / ref_loss.abs()).max().item()),
            "max_rel_diff_q_bar_vs_pipeline": float(((api_out["q_bar"] - ref_qbar).abs().max() / ref_qbar.abs().max()).item())}
        del q_t, api_out
        torch.cuda.empty_cache()

    # every kernel of the step alone, for the roofline table (rank 0's numbers are reported)
    pipe = ForwardAdjoint(s, batch=B, solver="cholesky", device=local, fused_assembly=False)
    pipe.set_inputs(qb, prob.xyz_fixed, prob.loads, x0)
    pipe.step()
    pipe.stream.synchronize()
    streams = stream_bandwidths(dev, pipe.stream)
    kernels = kernel_table(pipe, s, B, max(3, min(steps, 10)), flush_clean_on(pipe.stream), hbm, streams["write_gbs"])
    del pipe
    torch.cuda.empty_cache()

    # ---------------- second workload: one 200x200 cable-net (rank 0 only) ----------------
    single = sweep = None
    if rank == 0:
        cprob = datasets.cablenet_problem(CABLENET_NX)
        cs = eq.EquilibriumStructureSparse(cprob.nodes, cprob.edges, cprob.supports, device=local)
        cfa = NativeForwardAdjoint(cs, 1, solver="auto", tol=1e-12, device=local)
        cfa.set_inputs(cprob.q[None], cprob.xyz_fixed, cprob.loads, cprob.xyz0[cs.indices_free])
        for _ in range(3):
            cfa.step(main_stream)
        main_stream.synchronize()
        k = max(3, min(steps, 10))
        tc = timed_steps(lambda: cfa.step(main_stream), main_stream, k, flush_dev)
        inf_f, inf_a = cfa.array("info_f").cpu().numpy()[0], cfa.array("info_a").cpu().numpy()[0]
        single = {"workload": f"cable-net {CABLENET_NX}x{CABLENET_NX}, 4 corners fixed, q~U[0.5,2], fwd+adjoint",
                  "ms_per_solve": float(np.mean(tc)), "solves_per_s": 1e3 / float(np.mean(tc)),
                  "solver": "two-level PCG (persistent)",
                  "iters_fwd": int(inf_f[1]), "iters_adj": int(inf_a[1]),
                  "relres_fwd": float(inf_f[2]), "relres_adj": float(inf_a[2]),
                  "status": [int(inf_f[0]), int(inf_a[0])]}
        cfa.close()
        # one design of the headline topology: the latency an optimiser loop over a single structure sees per
        # value-and-gradient (device-resident, one CUDA graph; the two-sided factorisation: a warp pair per design)
        ofa = NativeForwardAdjoint(s, 1, solver="cholesky", device=local)
        ofa.set_inputs(qb[:1], prob.xyz_fixed, prob.loads, x0)
        for _ in range(3):
            ofa.step(main_stream)
        main_stream.synchronize()
        t1 = timed_steps(lambda: ofa.step(main_stream), main_stream, k, flush_dev)
        single["one_design_30x30"] = {"ms_per_value_and_grad": float(np.mean(t1)),
                                      "how": "fdm_fa_step_device at batch 1 (pillow 30x30): latency, not throughput"}
        ofa.close()
        sweep = sweep_large_mesh(SWEEP_NX, 5, flush_buf, hbm)

    if world > 1:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None

    dom = kernels[1]                                     # the factor kernel instance the step launches (FROM_Q = 1)
    ms_per_step = sum_dev / steps
    value = B * world / (ms_per_step * 1e-3)
    e2e_ms = sum_e2e / steps
    e2e_value = B * world / (e2e_ms * 1e-3)
    copy_ms = sum_copy / steps
    ctr = ncu_counters().get("chol_factor_mma_kernel<1,4,0,0>", {})   # <FROM_Q, PW, SIGNED, TW>: the one-sided instance of the step
    strict = strict_bytes(s)
    E, Nf = s.num_edges, s.num_free
    fac_strict = 8 * E + 24 * Nf                         # the factor kernel's compulsory share: q in (x out comes later)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "per_gpu_batch": B, "global_batch": B * world, "solver": "batched band Cholesky, one warp per design",
                   "l2": "flushed between timed steps (256 MiB write, outside the events)",
                   "device_step": f"library-owned pipeline (fdm_fa_step_device): {DEV_CHUNKS} slices of the batch on their "
                                  "own streams inside one CUDA graph, forked from and joined to the timed stream (CUDA "
                                  "events on it bracket the whole step); the factorisation assembles its rows from q in "
                                  "place (fdm_solve_from_q); the factor kernels carry the LOW launch priority and every other "
                                  "kernel of a slice the high one, so one slice's bandwidth-bound kernels take freed CTA "
                                  "slots ahead of the queued factor CTAs of the next",
                   "parallelism": f"designs sharded x{world}, no data-path collective"
                                  + (", 8-byte loss all-reduce (NCCL) per step" if world > 1 else "")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_fa.h2d_bytes * world,
                "d2h_bytes_per_step": e2e_fa.d2h_bytes * world, "ms_per_step": e2e_ms,
                "how": f"fdm_fa_run_host: pinned host q -> device, forward + adjoint, loss + q_bar -> pinned host, in "
                       f"{E2E_CHUNKS} slices on their own streams so copies overlap kernels (the per-design losses leave in "
                       "one copy at the end); one C call per step; host wall clock, synchronised both sides of EVERY step",
                "copy_roofline": {"ms_per_step": copy_ms,
                                  "gbs_per_gpu_each_way": e2e_fa.h2d_bytes / (copy_ms * 1e-3) / 1e9,
                                  "e2e_over_copy": e2e_ms / copy_ms,
                                  "how": "the same cudaMemcpyAsync calls (same slices, streams, pinned buffers) with no "
                                         "kernels between them, all ranks at once: what the host <-> device links deliver"},
                "loss_only_readback": {"value": B * world / (sum_lossonly / steps * 1e-3), "unit": UNIT,
                                       "ms_per_step": sum_lossonly / steps,
                                       "d2h_bytes_per_step": e2e_fa.h_loss.size * 8 * world,
                                       "how": "the same call with q_bar left on the device (h_q_bar = NULL): host q in, "
                                              "per-design losses out"},
                "pipelined": {"value": B * world / (t_pipelined * 1e-3), "unit": UNIT, "ms_per_step": t_pipelined,
                              "how": "the same K steps enqueued back to back and synchronised once: step k+1's copy-in "
                                     "overlaps step k's kernels and copy-out (independent batches; an optimiser loop "
                                     "that needs step k's gradient before step k+1 gets the synchronous number)"}},
        "gpu_launches": dev_fa.launches_per_step * steps,
        "roofline": {"kernel": "chol_factor_mma_kernel<FROM_Q=1,PW=4> (factorisation from q, fused forward substitution)",
                     "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": hbm,
                     "unit": "GB/s", "frac": dom["achieved_gbs"] / hbm, "traffic": ctr.get("dram_bytes"),
                     "traffic_source": ctr.get("source"),
                     "peak_source": hbm_src, "kernel_ms": dom["ms"],
                     "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                     "algorithmic_bytes_note": "q read + the factor (32 doubles per column) and forward-substituted "
                                               "right-hand sides written: what THIS design must move; SURVEY §8d's "
                                               "compulsory input is the q read alone (`strict`)",
                     "strict": {"bytes_per_launch": fac_strict * B,
                                "achieved": fac_strict * B / (dom["ms"] * 1e-3) / 1e9,
                                "frac": fac_strict * B / (dom["ms"] * 1e-3) / 1e9 / hbm},
                     "share_of_step": dom["ms"] / ms_per_step,
                     "limiter": "latency / issue: the 841-pivot dependency chain of one warp per design at "
                                "register-capped occupancy; neither HBM nor a math pipe is near its limit (ncu `counters`)",
                     "counters": {k: ctr.get(k) for k in ("warps_active_pct", "l1_data_pipe_pct", "dmma_pipe_pct",
                                                          "fp64_pipe_pct", "dram_pct", "duration_us", "source")}},
        "step_roofline": {"strict_bytes_per_step": strict * B,
                          "achieved": strict * B / (ms_per_step * 1e-3) / 1e9, "unit": "GB/s",
                          "frac": strict * B / (ms_per_step * 1e-3) / 1e9 / hbm,
                          "note": "SURVEY.md §8d compulsory bytes of the whole forward + adjoint step (q in; x_free, xyz, "
                                  "state, loss, q_bar, loads_bar, xyz_fixed_bar out) over the timed step"},
        "kernels": kernels,
        "kernels_l2": "cold AND clean before every launch of the table: a 256 MiB write (the flush of the timed step) followed "
                      "by a 384 MB read, so that the write-back of the flush buffer's dirty lines is not charged to the kernel",
        "stream_bandwidths": dict(streams, note="measured in this run (1 GiB fill / copy, best of 5); "
                                                  "`frac_of_write_adjusted_roofline` in `kernels` = max(all bytes / copy "
                                                  "peak, written bytes / write peak) / time"),
        "sweep_large_mesh": sweep,
        "config5_fixed_global_batch": fixed,
        "wall_s_timed_region": wall,
        "solve_status_ok": status_ok and (single is None or single["status"] == [0, 0]),
        "single_mesh": single,
    }
    line.update(line_extra)
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
