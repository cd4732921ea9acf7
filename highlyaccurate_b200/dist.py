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


class PoseComm:
    """The library's own communicator for the one collective of the path (ha_comm_* / ha_pose_allgather in
    include/ha_b200.h): NCCL bound by libha_b200.so at run time, no torch types on the data path.  The 128-byte
    NCCL id travels from rank 0 to the others through the already-initialised torch.distributed group (any
    out-of-band channel would do); after that torch.distributed is not involved in the gather any more."""

    def __init__(self, rank: int, world: int, device: torch.device):
        import ctypes as C
        from . import _lib
        self._lib, self.rank, self.world, self.device = _lib, rank, world, device
        L = _lib.lib()
        ident = (C.c_ubyte * _lib.HA_COMM_ID_BYTES)()
        if rank == 0:
            _lib.check(L.ha_comm_unique_id(ident), "ha_comm_unique_id")
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0)
        ident = (C.c_ubyte * _lib.HA_COMM_ID_BYTES).from_buffer_copy(box[0])
        handle = C.c_void_p()
        _lib.check(L.ha_comm_init(C.byref(handle), world, rank, ident, device.index or 0), "ha_comm_init")
        self.handle = handle

    def allgather(self, local: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """local [b, 3] fp32 (may be the rank's slice of `out`: in place) -> [world * b, 3]."""
        assert local.is_cuda and local.dtype == torch.float32 and local.is_contiguous() and local.shape[-1] == 3
        b = local.shape[0]
        if out is None:
            out = torch.empty(self.world * b, 3, dtype=torch.float32, device=local.device)
        st = torch.cuda.current_stream(local.device).cuda_stream
        self._lib.check(self._lib.lib().ha_pose_allgather(self.handle, local.data_ptr(), out.data_ptr(), b, st),
                        "ha_pose_allgather")
        return out

    def close(self):
        if self.handle is not None:
            self._lib.lib().ha_comm_destroy(self.handle)
            self.handle = None


def gather_poses(local: torch.Tensor, world: int, comm: "PoseComm | None" = None) -> torch.Tensor:
    """All-gather the per-rank final poses [b, 3] into [world * b, 3] (equal shards).  This is the only collective of
    the path; 12 KB at B = 1024, latency bound.  With a PoseComm it goes through the C ABI (ha_pose_allgather);
    without one (the gloo CPU tests of the host logic) through torch.distributed."""
    if world == 1:
        return local
    if comm is not None:
        return comm.allgather(local.contiguous())
    if not dist.is_initialized():
        return local
    out = torch.empty(world * local.shape[0], *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out
