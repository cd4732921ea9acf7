#!/usr/bin/env python
"""Which batch sizes the fused LM loop is run at (ha_lm_run, KITTI pyramid, 5 x 3 steps): time and fraction of the HBM roofline."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HA_QUIET", "1")
import torch
from highlyaccurate_b200 import engine
from highlyaccurate_b200.models_kitti import LM_S2GP
from bench import ref_args, PYR_C, peaks, SAT_TEXELS_TOUCHED
dev = torch.device("cuda:0")
pk = peaks()
L = 3
VARIANTS = [int(v) for v in os.environ.get("HA_LM_VARIANTS", "0,2").split(",")]
for B in [int(b) for b in sys.argv[1:]] or [32, 64, 128, 256, 512]:
    net = LM_S2GP(ref_args(5, L)).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    sat = engine.Pyramid([torch.randn(B, 512 >> (3 - l), 512 >> (3 - l), PYR_C[l], device=dev, generator=g) for l in range(L)], [None] * L)
    grd = engine.Pyramid([torch.randn(B, 256 >> (3 - l), 1024 >> (3 - l), PYR_C[l], device=dev, generator=g) for l in range(L)], [None] * L)
    draws = torch.zeros(5 * L, 2, B, device=dev)
    for v in VARIANTS:
        for _ in range(3):
            res = net.refine(sat, grd, reset_uv=draws, kernel_variant=v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            res = net.refine(sat, grd, reset_uv=draws, kernel_variant=v)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        byt = sum(4 * PYR_C[l] * (((256 >> (3 - l)) // 2) * (1024 >> (3 - l)) + SAT_TEXELS_TOUCHED["kitti"][l]) for l in range(L)) * 5 * B
        print("B=%4d variant %d: %.3f ms  %.1f GB/s = %.3f of HBM peak (status %d)" % (B, v, ms, byt / ms / 1e6, byt / ms / 1e6 / pk["hbm"], int(res.status.item())))
    del sat, grd
    torch.cuda.empty_cache()
