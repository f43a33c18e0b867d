"""Data-parallel training over independent mesh graphs (one process per GPU).

The reference is single-device.  The path shards naturally over graphs:
every rank runs the whole fwd+bwd on its own batch of meshes (graph-LayerNorm statistics
are per rank = the reference run with batch B_local) and the only exchange is ONE NCCL
all-reduce of the flat 167 299-float gradient buffer ``pdg_backward`` writes -- it is
averaged in place inside the autograd backward, before torch sees the per-parameter views.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """torchrun-style initialisation; returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_indices(n_items: int, rank: int, world: int):
    """Round-robin graph ownership: item i belongs to rank i % world."""
    return list(range(rank, n_items, world))


def allreduce_flat_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over ranks of a flat gradient buffer.

    NCCL: ONE collective with ``ReduceOp.AVG`` (the 1/G scale happens inside the reduction kernel -- no separate
    elementwise launch between the backward and the optimizer).  Other backends (gloo in the CPU tests): sum, then
    1/G (exact for G = 2^k)."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            if flat.is_cuda and dist.get_backend(group) == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
                flat.mul_(1.0 / world)
    return flat


def enable_data_parallel(model, group=None):
    """Average ``pdg_backward``'s flat gradient across ranks inside backward."""
    model._pdg_dp = (True, group)
    return model


def broadcast_parameters(model, src: int = 0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src, group=group)
    return model
