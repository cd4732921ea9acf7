#!/usr/bin/env python
"""Launch the fused LM step once per pyramid level (after one warm-up sweep) for an ncu capture:
    ncu --set full --import-source on -k regex:lm_step -s L -c L -o out python tools/ncu_lm.py [B] [variant] [levels]
(levels = 3: BASELINE config 2 pyramid; 4: config 5, adds the C = 16 full-resolution level)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HA_QUIET", "1")
import torch
from highlyaccurate_b200 import engine
from highlyaccurate_b200.models_kitti import LM_S2GP
from bench import ref_args, PYR_C

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
net = LM_S2GP(ref_args(5, L)).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
sat = engine.Pyramid([torch.randn(B, 512 >> (3 - l), 512 >> (3 - l), PYR_C[l], device=dev, generator=g) for l in range(L)], [None] * L)
grd = engine.Pyramid([torch.randn(B, 256 >> (3 - l), 1024 >> (3 - l), PYR_C[l], device=dev, generator=g) for l in range(L)], [None] * L)
setup = engine.setup_from_args(net.args, "kitti", 0)
if len(sys.argv) > 2:
    setup.kernel_variant = int(sys.argv[2])
tabs = net._tables(dev)
pose = (torch.rand(B, 3, device=dev) - 0.5) * 0.4
zeros = torch.zeros(2, B)
for rep in range(2):
    for lv in range(L):
        engine.lm_step(setup, lv, sat, grd, tabs, [0.1] * 3, pose, reset_uv=zeros)
torch.cuda.synchronize()
