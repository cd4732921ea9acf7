#!/usr/bin/env python
"""Micro-benchmark of the fused LM step per pyramid level (random features, B pairs, KITTI shapes).
    python tools/bench_lm.py [B] [reps] [levels] [variants, e.g. 0,1,2]
Prints per-level time, algorithmic GB/s and fraction of the measured HBM peak for every kernel variant
(HaLmParams.kernel_variant: 0 = default ring kernel, whole loop with chained launches; 1 = register-staged validation
kernel; 2 = default kernel, whole loop with plain stream-ordered launches)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HA_QUIET", "1")
import torch
from highlyaccurate_b200 import engine
from highlyaccurate_b200.models_kitti import LM_S2GP
from bench import ref_args, PYR_C, peaks
from bench import SAT_TEXELS_TOUCHED as _STT
SAT_TEXELS_TOUCHED = _STT['kitti']

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = int(sys.argv[3]) if len(sys.argv) > 3 else 3
VARIANTS = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [None]
dev = torch.device("cuda:0")
net = LM_S2GP(ref_args(5, L)).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
sat = engine.Pyramid([torch.randn(B, 512 >> (3 - l), 512 >> (3 - l), PYR_C[l], device=dev, generator=g) for l in range(L)], [None] * L)
grd = engine.Pyramid([torch.randn(B, 256 >> (3 - l), 1024 >> (3 - l), PYR_C[l], device=dev, generator=g) for l in range(L)], [None] * L)
setup = engine.setup_from_args(net.args, "kitti", 0)
tabs = net._tables(dev)
pose = (torch.rand(B, 3, device=dev) - 0.5) * 0.4
zeros = torch.zeros(2, B)
pk = peaks()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for variant in VARIANTS:
    if variant is not None:
        setup.kernel_variant = variant
        print('--- kernel_variant=%d  B=%d' % (variant, B))
    for lv in range(L):
        for _ in range(3):
            engine.lm_step(setup, lv, sat, grd, tabs, [0.1] * 3, pose, reset_uv=zeros)
        ts = []
        for _ in range(reps):
            flush.zero_()                                # flush L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # time only the kernel: call the C entry directly through engine.lm_step's inner launch
            torch.cuda.synchronize()
            e0.record()
            engine.lm_step(setup, lv, sat, grd, tabs, [0.1] * 3, pose, reset_uv=zeros)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        t = ts[len(ts) // 2]
        h, w = 256 >> (3 - lv), 1024 >> (3 - lv)
        byt = 4 * PYR_C[lv] * ((h // 2) * w + SAT_TEXELS_TOUCHED[lv]) * B
        print("level %d C=%3d: %8.1f us (incl. ~10 us of host-side staging)  %7.1f GB/s  %.3f of %s HBM peak"
              % (lv, PYR_C[lv], t, byt / t / 1e3, byt / t / 1e3 / pk["hbm"], pk["src"]))

    # whole loop (5 iterations x L levels): exercises the cached-|g|^2 (FAST) launches too
    kv = {} if variant is None else {"kernel_variant": variant}
    for _ in range(2):
        res = net.refine(sat, grd, reset_uv=torch.zeros(5 * L, 2, B), **kv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    draws = torch.zeros(5 * L, 2, B, device=dev)
    e0.record()
    for _ in range(reps):
        res = net.refine(sat, grd, reset_uv=draws, **kv)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byt = sum(4 * PYR_C[l] * (((256 >> (3 - l)) // 2) * (1024 >> (3 - l)) + SAT_TEXELS_TOUCHED[l]) for l in range(L)) * 5 * B
    print("whole LM loop (5 iters x %d levels, B=%d): %.3f ms  -> %.1f GB/s algorithmic = %.3f of HBM peak  (status %d)"
          % (L, B, ms, byt / ms / 1e6, byt / ms / 1e6 / pk["hbm"], int(res.status.item())))
