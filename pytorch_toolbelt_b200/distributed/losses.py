"""Data-parallel losses: the batch is sharded over ranks, one tiny all-reduce makes the loss global.

New relative to the reference, whose Dice / Jaccard losses under DDP are computed per rank on the local batch (the
region losses are NOT linear in the batch, so the average of per-rank Dice differs from the Dice of the whole batch).

* ``SyncDiceLoss`` / ``SyncJaccardLoss``: per-rank statistics kernel -> ``all_reduce`` of the float64 ``[C,3]`` sums
  (I, S, T per class; 24*C bytes) -> local loss + gradient kernel fed the global sums. Every rank returns the
  global-batch loss; ``backward`` yields d(global loss)/d(local logits) exactly, with no second collective.
* ``global_mean``: for losses that ARE linear in the batch (focal, wing with ``reduction="mean"`` on equal shards):
  all-reduce of the scalar for reporting, gradient 1/world to the local term.

GRADIENT CONVENTION -- read this before combining with DistributedDataParallel. ``backward`` hands every rank the exact
d(global loss)/d(its local logits). The gradient of the global loss w.r.t. a PARAMETER is therefore the SUM over ranks of
the per-rank parameter gradients. DDP, however, AVERAGES parameter gradients over ranks, which would make the effective
gradient ``world_size`` times too small (``world_size**2`` for ``global_mean``, whose backward already carries 1/world).
Pick one:

* ``grad_reduction="sum"`` (default): you reduce parameter gradients with a SUM yourself (manual ``all_reduce``, or DDP
  with a comm hook that sums) -- nothing is rescaled;
* ``grad_reduction="ddp_mean"``: the backward is multiplied by ``world_size`` (SyncBatchNorm-style), so that DDP's averaging
  reproduces the single-process whole-batch gradient exactly. ``global_mean(..., grad_reduction="ddp_mean")`` likewise
  back-propagates 1 instead of 1/world.
``tests/test_distributed_cpu.py::test_sync_loss_gradient_conventions`` checks both conventions against the whole-batch
parameter gradient over gloo.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from ..losses.dice import DiceLoss
from ..losses.jaccard import JaccardLoss

__all__ = ["SyncDiceLoss", "SyncJaccardLoss", "global_mean"]


_GRAD_REDUCTIONS = ("sum", "ddp_mean")


def _check_grad_reduction(grad_reduction: str) -> str:
    if grad_reduction not in _GRAD_REDUCTIONS:
        raise ValueError(f"grad_reduction must be one of {_GRAD_REDUCTIONS}, got {grad_reduction!r}")
    return grad_reduction


class _SyncMixin:
    process_group = None
    grad_reduction = "sum"

    def _sync_cfg(self) -> dict:
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError(f"{type(self).__name__} needs an initialised torch.distributed process group")
        mul = float(dist.get_world_size(self.process_group)) if self.grad_reduction == "ddp_mean" else 1.0
        return {"sync": True, "group": self.process_group, "grad_mul": mul}


class SyncDiceLoss(_SyncMixin, DiceLoss):
    """``DiceLoss`` (same constructor) over the GLOBAL batch of a data-parallel group; see the module docstring for
    ``grad_reduction`` ("sum" | "ddp_mean")."""

    def __init__(self, *args, process_group=None, grad_reduction: str = "sum", **kwargs):
        super().__init__(*args, **kwargs)
        self.process_group = process_group
        self.grad_reduction = _check_grad_reduction(grad_reduction)


class SyncJaccardLoss(_SyncMixin, JaccardLoss):
    """``JaccardLoss`` (same constructor) over the GLOBAL batch of a data-parallel group; see the module docstring for
    ``grad_reduction`` ("sum" | "ddp_mean")."""

    def __init__(self, *args, process_group=None, grad_reduction: str = "sum", **kwargs):
        super().__init__(*args, **kwargs)
        self.process_group = process_group
        self.grad_reduction = _check_grad_reduction(grad_reduction)


class _GlobalMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, group, back):
        world = dist.get_world_size(group)
        ctx.back = back if back is not None else 1.0 / world
        total = value.detach().clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return total / world

    @staticmethod
    def backward(ctx, upstream):
        return upstream * ctx.back, None, None


def global_mean(local_mean: torch.Tensor, group=None, grad_reduction: str = "sum") -> torch.Tensor:
    """Mean over ranks of a per-rank mean loss (equal shards): value identical on all ranks. Gradient to the local term:
    1/world with ``grad_reduction="sum"`` (parameter gradients summed over ranks), 1 with ``"ddp_mean"`` (DDP averages)."""
    return _GlobalMean.apply(local_mean, group, 1.0 if _check_grad_reduction(grad_reduction) == "ddp_mean" else None)
