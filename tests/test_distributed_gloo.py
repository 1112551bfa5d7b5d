"""World-size-2 `gloo` test of the N>1 host logic (shard + loss / shared-gradient all-reduce), CPU only.

Each rank evaluates its slice of a small pillow batch with the CPU oracle (the GPU kernels are
covered by the -m gpu parity tests); the all-reduced totals must equal the single-process sums.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jax_fdm_b200 import datasets
from jax_fdm_b200.distributed import shard_range, shared_parameter_grad, total_loss, world_info


def test_shard_range_partitions_exactly():
    for total in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_range(total, r, world)
                assert 0 <= lo <= hi <= total
                cover.extend(range(lo, hi))
            assert cover == list(range(total))
            sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_world_info_without_process_group():
    assert world_info() == (0, 1)


def _design_loss_and_grad(b, prob, so, x0):
    import oracle

    q = datasets.pillow_batch_q(prob.edges.shape[0], b, 1)[0]

    def cot(st, xf):
        c = np.zeros_like(st["xyz"])
        c[so.indices_free] = xf - x0
        return {"xyz": c}

    out = oracle.forward_adjoint(q, prob.xyz_fixed, prob.loads, so, cot)
    return 0.5 * np.sum((out["x_free"] - x0) ** 2), out["q_bar"]


def _worker(rank, world, port, total, q):
    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = datasets.pillow_problem(6)
        so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
        x0 = prob.xyz0[so.indices_free]
        lo, hi = shard_range(total, rank, world)
        res = [_design_loss_and_grad(b, prob, so, x0) for b in range(lo, hi)]
        loss = torch.tensor([r[0] for r in res], dtype=torch.float64)
        grads = torch.tensor(np.stack([r[1] for r in res]), dtype=torch.float64)
        tot = total_loss(loss)
        g = shared_parameter_grad(grads)
        q.put((rank, world_info(), (lo, hi), float(tot.item()), g.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_loss_and_shared_gradient_allreduce():
    import oracle

    total, world = 5, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob = datasets.pillow_problem(6)
    so = oracle.build_structure(prob.nodes, prob.edges, prob.supports)
    x0 = prob.xyz0[so.indices_free]
    ref = [_design_loss_and_grad(b, prob, so, x0) for b in range(total)]
    ref_loss = sum(r[0] for r in ref)
    ref_grad = np.sum([r[1] for r in ref], axis=0)
    slices = sorted(g[2] for g in got)
    assert slices == [(0, 3), (3, 5)]
    for rank, winfo, _, tot, grad in got:
        assert winfo == (rank, world)
        assert abs(tot - ref_loss) <= 1e-12 * abs(ref_loss)
        np.testing.assert_allclose(grad, ref_grad, rtol=1e-12, atol=1e-14)
