"""Multi-GPU plumbing: the batch of image pairs shards embarrassingly (every sample's VGG + LM
loop is independent: no BatchNorm, per-sample norms and normal equations — SURVEY.md section 8e),
one process per GPU, ONE collective: an all-gather of the final [B/G, 3] poses.

The reference has no distributed code at all (train_kitti.py:526-529 uses cuda:0 only).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group when
    WORLD_SIZE > 1.  NCCL on GPUs, gloo otherwise."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of n samples owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_reset_draws(draws: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """The reference draws [B,1] reset samples for the WHOLE batch from the CPU generator
    (models_kitti.py:1028-1029).  Every rank makes the same full-batch draws (same seed) and keeps
    its slice, so an N-GPU run consumes the RNG stream exactly like the single-GPU run."""
    lo, hi = shard_bounds(draws.shape[-1], rank, world)
    return draws[..., lo:hi].contiguous()


def gather_poses(local: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather the per-rank final poses [b, 3] into [world * b, 3] (equal shards).  This is the
    only collective of the path; 12 KB at B = 1024, latency bound."""
    if world == 1 or not dist.is_initialized():
        return local
    out = torch.empty(world * local.shape[0], *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out
