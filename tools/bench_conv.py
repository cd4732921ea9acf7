#!/usr/bin/env python
"""A/B of the U-Net schedules on one GPU: CTA-pair (cta_group::2) halo kernels vs the single-CTA kernels.
    python tools/bench_conv.py [B] [reps]
Checks the two against each other on the same input, then times one branch (512 x 512) per precision."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HA_QUIET", "1")
import torch
from highlyaccurate_b200.VGG import VGGUnet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = VGGUnet(3).to(dev)
x = torch.rand(B, 3, 512, 512, device=dev)
out = {}
for prec in ("f16x3_1cta", "f16x3", "f16"):
    net.precision = prec
    for _ in range(2):
        p = net.pyramid(x, want_conf=False, want_scale=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        p = net.pyramid(x, want_conf=False, want_scale=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[prec] = [f.clone() for f in p.feats]
    gf = 520056 * 262144 * B / 1e9
    print("%-11s %8.3f ms per branch (B=%d)  %.0f TFLOP/s algorithmic" % (prec, ms, B, gf / ms))
for l in range(3):
    a, b = out["f16x3_1cta"][l], out["f16x3"][l]
    print("level %d: pair vs single-CTA max|d| / max|x| = %.2e" % (l, float((a - b).abs().max() / a.abs().max())))
