"""Assembly (K1) and SpMM timed alone at the bench batch, L2 flushed before each launch (CUDA events).
Usage: python tools/time_stream_kernels.py [B]   -- environment knobs FDM_ASM_BATCH / FDM_ASM_DPC / FDM_SPMM_BATCH /
FDM_SPMM_DPC select the variant."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jax_fdm_b200 import _lib as L  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    from jax_fdm_b200 import datasets
    import jax_fdm_b200.equilibrium as eq
    from jax_fdm_b200.pipeline import ForwardAdjoint
    prob = datasets.pillow_problem(bench.PILLOW_NX)
    s = eq.EquilibriumStructureSparse(prob.nodes, prob.edges, prob.supports, device=0)
    pipe = ForwardAdjoint(s, batch=B, solver="cholesky", device=0, fused_assembly=False)
    pipe.set_inputs(datasets.pillow_batch_q(prob.edges.shape[0], 0, B), prob.xyz_fixed, prob.loads, prob.xyz0[s.indices_free])
    pipe.step()
    pipe.stream.synchronize()
    lib = L.lib()
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(pipe.stream.cuda_stream)
    E, Nf, nnz, N, Ns = s.num_edges, s.num_free, s.nnz, s.num_nodes, s.num_supports
    fb = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    rb = torch.ones(64 << 20, dtype=torch.float64, device=dev)       # 512 MB, read after the write-flush
    mode = os.environ.get("FLUSH", "write")

    def flush():
        with torch.cuda.stream(pipe.stream):
            if mode != "none":
                fb.fill_(1)
            if mode == "write+read":                                   # leaves L2 full of CLEAN lines: no write-back of
                rb.sum()                                               # the flush buffer is charged to the timed kernel
    Y = pipe.rhs.clone()
    K2 = pipe.K.clone()
    calls = [
        ("assemble", 8 * E + 8 * nnz + 24 * Nf,
         lambda: lib.fdm_assemble(s.plan, P(pipe.q), P(pipe.xyz_fixed), P(pipe.loads), P(pipe.K), P(pipe.rhs), B, E, 0, 0, st)),
        ("asm K only", 8 * E + 8 * nnz,
         lambda: lib.fdm_assemble(s.plan, P(pipe.q), P(pipe.xyz_fixed), P(pipe.loads), P(pipe.K), None, B, E, 0, 0, st)),
        ("asm rhs only", 8 * E + 24 * Nf,
         lambda: lib.fdm_assemble(s.plan, P(pipe.q), P(pipe.xyz_fixed), P(pipe.loads), None, P(pipe.rhs), B, E, 0, 0, st)),
        ("fill K", 8 * nnz, lambda: (pipe.K.fill_(1.0), 0)[1]),
        ("fill K+8B", 8 * nnz, lambda: (pipe.K.view(-1)[1:].fill_(1.0), 0)[1]),
        ("fill rhs", 24 * Nf, lambda: (pipe.rhs.fill_(1.0), 0)[1]),
        ("fill rhs+8B", 24 * Nf, lambda: (pipe.rhs.view(-1)[1:].fill_(1.0), 0)[1]),
        ("fill rhs f32x2", 24 * Nf, lambda: (pipe.rhs.view(torch.float32).fill_(1.0), 0)[1]),
        ("copy K->K2", 16 * nnz, lambda: (K2.copy_(pipe.K), 0)[1]),
        ("state", 8 * E + 24 * N + 40 * E + 24 * N,
         lambda: lib.fdm_state(s.plan, P(pipe.q), P(pipe.xyz), P(pipe.loads), P(pipe.vectors), P(pipe.lengths),
                               P(pipe.forces), P(pipe.residuals), B, E, 0, st)),
        ("grads", 8 * E + 24 * N + 24 * Nf + 8 * E + 24 * N + 24 * Ns,
         lambda: lib.fdm_adjoint_grads(s.plan, P(pipe.q), P(pipe.xyz), P(pipe.lam), P(pipe.q_bar), P(pipe.loads_bar),
                                       P(pipe.xs_bar), B, E, 0, st)),
        ("loss", 48 * Nf + 8,
         lambda: lib.fdm_loss_sqdist(P(pipe.x), P(pipe.x0), P(pipe.g), P(pipe.loss), Nf * 3, B, 0, st)),
        ("spmm", 8 * nnz + 48 * Nf, lambda: lib.fdm_spmm(s.plan, P(pipe.K), P(pipe.x), P(Y), B, st)),
    ]
    def run(fn):
        with torch.cuda.stream(pipe.stream):
            L.check(fn())

    for name, per, fn in calls:
        ts = bench.timed_steps(lambda: run(fn), pipe.stream, 20, flush)
        ms = float(np.median(ts))
        print(f"[{mode}] {name:10s} {ms:.4f} ms  {per * B / ms / 1e6:.0f} GB/s  env={ {k: v for k, v in os.environ.items() if k.startswith('FDM_')} }", flush=True)


if __name__ == "__main__":
    main()
