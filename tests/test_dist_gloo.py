"""CPU, world_size 2 over gloo: batch sharding + the single pose all-gather of the N-GPU path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from highlyaccurate_b200 import dist as hd
from highlyaccurate_b200 import engine


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = hd.init_from_env("gloo")
    B = 6
    torch.manual_seed(123)                                   # same seed everywhere: identical full-batch draws
    draws = engine.draw_reset_uv(4, B)
    mine = hd.shard_reset_draws(draws, r, w)
    lo, hi = hd.shard_bounds(B, r, w)
    poses = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3)[lo:hi] + 0.5
    allp = hd.gather_poses(poses, w)
    q.put((rank, mine, allp, draws))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    full = torch.arange(18, dtype=torch.float32).reshape(6, 3) + 0.5
    for rank, mine, allp, draws in res:
        assert torch.equal(allp, full)
        assert torch.equal(mine, draws[..., rank * 3:(rank + 1) * 3])
    assert torch.equal(res[0][3], res[1][3])


def test_shard_bounds_cover_batch():
    for n in (1, 7, 32, 1024):
        for w in (1, 2, 4, 8):
            spans = [hd.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
