"""One training step (forward(mode='train') + loss.backward(), train_kitti.py:354-365) on the GPU:
native (tcgen05 U-Net forward / data / weight gradients + fused LM loop forward / backward) vs the same module with the
U-Nets on torch autograd / cuDNN (TF32 off and on).  usage: python tools/bench_train.py [B=3] [reps=5] [only: native|torch_cudnn_fp32|torch_cudnn_tf32]  -> one JSON line"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HA_QUIET", "1")
from highlyaccurate_b200 import _lib  # noqa: E402
from highlyaccurate_b200.models_kitti import LM_S2GP  # noqa: E402
from tests.cases import ref_args  # noqa: E402


def step(net, sat, grd, gt):
    net.zero_grad(set_to_none=True)
    out = net(sat, grd, gt[:, 0:1], gt[:, 1:2], gt[:, 2:3], mode="train")
    out[0].backward()
    return out[0]


def timed(net, sat, grd, gt, reps):
    for _ in range(2):
        step(net, sat, grd, gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.lib().ha_launch_count()
    e0.record()
    for _ in range(reps):
        loss = step(net, sat, grd, gt)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, float(loss.detach()), (_lib.lib().ha_launch_count() - n0) // reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = LM_S2GP(ref_args()).to(dev)
    g = torch.Generator().manual_seed(1)
    sat = torch.rand(B, 3, 512, 512, generator=g).to(dev)
    grd = torch.rand(B, 3, 256, 1024, generator=g).to(dev)
    gt = (torch.rand(B, 3, generator=g) * 2 - 1).to(dev)
    res = {"what": "LM_S2GP train step: forward(mode='train') + backward, KITTI shapes, level 3, 5 LM iterations", "batch": B}
    only = sys.argv[3] if len(sys.argv) > 3 else None
    for name, native, tf32 in (("native", True, False), ("torch_cudnn_fp32", False, False), ("torch_cudnn_tf32", False, True)):
        if only and name != only:
            continue
        net.SatFeatureNet.native_train = net.GrdFeatureNet.native_train = native
        torch.backends.cudnn.allow_tf32 = tf32
        torch.manual_seed(7)
        ms, loss, launches = timed(net, sat, grd, gt, reps)
        res[name] = {"ms_per_step": ms, "pairs_per_s": B / ms * 1e3, "loss": loss, "library_launches_per_step": int(launches)}
    torch.backends.cudnn.allow_tf32 = False
    # 2 U-Nets x (forward 136.3 + data gradients ~133 + weight gradients 136.3 GFLOP per image), f16x3 issues 3x
    res["unet_gflop_per_pair_fwd_bwd"] = 2 * (136.3 * 3 - 3.6)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
