"""
Batch sharding of independent designs across the GPUs of one box.

Designs that share one topology are independent units: rank r of W takes a contiguous slice of
the batch axis of `q` (and of every per-design output / cotangent); the plan (CSR pattern,
index maps, band order) is replicated.  The data path has NO collective.  The only exchange is
a sum over ranks of (a) the scalar loss and (b) the gradient of a parameter shared by all
designs when the solver sits as a layer inside a network -- one NCCL all-reduce over NVLink
(`gloo` in the CPU tests), latency-bound (8 bytes / size(theta) * 8 bytes).

The reference has no multi-device path (no pmap / sharding / psum anywhere); its only batched
form is `jax.vmap(model, in_axes=(0, None))` on one device (visualization/plotters/
loss_plotter.py:98-103).  This module is what replaces "vmap on one device" at 2-8 GPUs.
"""

import ctypes as C

import torch
import torch.distributed as dist

__all__ = ["shard_range", "world_info", "allreduce_sum_", "total_loss", "shared_parameter_grad"]


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of `total` designs owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def world_info():
    """(rank, world) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks (no-op in a single-process run).  Runs on the current CUDA stream."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def _batch_sum(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[i] = sum_b x[b, i]: this library's fixed-order reduction kernel (fdm_batch_sum) for device tensors; the
    world-size-2 `gloo` tests drive the host logic with CPU tensors, which take torch.sum."""
    if not x.is_cuda:
        return out.copy_(torch.sum(x, dim=0).reshape(out.shape))
    from . import _lib as L
    x = x.contiguous()
    n = x[0].numel() if x.dim() > 1 else 1
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    L.check(L.lib().fdm_batch_sum(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), x.shape[0], n, st))
    return out


def total_loss(loss_per_design: torch.Tensor, out: torch.Tensor = None, group=None) -> torch.Tensor:
    """sum_b L_b over the local shard, then over ranks: the scalar a shared optimiser sees."""
    if out is None:
        out = torch.zeros(1, dtype=loss_per_design.dtype, device=loss_per_design.device)
    _batch_sum(loss_per_design, out)
    return allreduce_sum_(out, group)


def shared_parameter_grad(grad_per_design: torch.Tensor, out: torch.Tensor = None, group=None) -> torch.Tensor:
    """
    Gradient of a parameter shared by every design of the batch (e.g. one q vector broadcast to
    all load cases): sum of the per-design gradients over the local shard and over ranks.
    """
    if out is None:
        out = torch.zeros(grad_per_design.shape[1:], dtype=grad_per_design.dtype, device=grad_per_design.device)
    _batch_sum(grad_per_design, out)
    return allreduce_sum_(out, group)
