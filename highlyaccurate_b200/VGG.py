"""Mirror of the reference's VGG.py public surface for the hot path: `VGGUnet` and `L2_norm`.

Same constructor, same parameter names/shapes (so reference checkpoints load unchanged:
VGG.py:23-81 -> 24 state-dict keys per U-Net), same forward contract (VGG.py:121-203) — but the
convolutions run in libha_b200.so.  The nn.Conv2d children below are parameter containers only.
"""
from __future__ import annotations

import os
import warnings

import torch
import torch.nn as nn

from . import engine


def _conv(cin, cout, bias):
    return nn.Conv2d(cin, cout, kernel_size=(3, 3), stride=(1, 1), padding=1, bias=bias)


def _dec(cin, cmid, cout):
    return nn.Sequential(nn.ReLU(inplace=True), _conv(cin, cmid, False), nn.ReLU(inplace=True), _conv(cmid, cout, False))


def _conf(cin):
    return nn.Sequential(nn.ReLU(), _conv(cin, 1, False), nn.Sigmoid())


class VGGUnet(nn.Module):
    """VGG.py:13.  `level` selects the returned pyramid: 3 -> [x15,x18,x21], 4 -> +x24 (VGG.py:192-203)."""

    G2S = False       # VGGUnet_G2S: decoders on the folded [2H, W/2] maps

    def __init__(self, level, estimate_depth=0):
        super().__init__()
        if estimate_depth:
            raise NotImplementedError("estimate_depth heads are outside the accelerated path (SURVEY.md section 2, row 1)")
        self.level = level
        self.estimate_depth = 0
        self.conv0 = _conv(3, 64, True)
        self.conv2 = _conv(64, 64, True)
        self.conv5 = _conv(64, 128, True)
        self.conv7 = _conv(128, 128, True)
        self.conv10 = _conv(128, 256, True)
        self.conv12 = _conv(256, 256, True)
        self.conv14 = _conv(256, 256, True)
        self.conv_dec1 = _dec(384, 128, 128)
        self.conv_dec2 = _dec(192, 64, 64)
        self.conv_dec3 = _dec(128, 32, 16)
        self.conf0, self.conf1, self.conf2, self.conf3 = _conf(256), _conf(128), _conf(64), _conf(16)
        self._load_pretrained_encoder()
        self.precision = os.environ.get("HA_VGG_PRECISION", "f16x3")
        self._runner = engine.VggRunner()
        self.native_train = True      # train mode: U-Net backward in libha_b200.so (engine.VggTrain); False = torch autograd / cuDNN
        self._named = None            # name -> Parameter, built once (Parameter objects are stable across .to() / load_state_dict)

    def _load_pretrained_encoder(self):
        """VGG.py:20-29 takes the encoder from torchvision's ImageNet VGG16.  Offline boxes have no
        checkpoint: keep the default initialisation and say so (weights normally come from
        load_state_dict anyway, train_kitti.py:546)."""
        path = os.environ.get("HA_VGG16_PTH") or os.path.join(torch.hub.get_dir(), "checkpoints", "vgg16-397923af.pth")
        if not os.path.exists(path):     # never try to download: no network on the target boxes
            if os.environ.get("HA_QUIET", "0") != "1":
                warnings.warn("VGG16 ImageNet checkpoint not found at %s; encoder keeps its random init" % path)
            return
        sd = torch.load(path, map_location="cpu")
        for idx in (0, 2, 5, 7, 10, 12, 14):
            getattr(self, "conv%d" % idx).load_state_dict({"weight": sd["features.%d.weight" % idx],
                                                           "bias": sd["features.%d.bias" % idx]})

    def _apply(self, fn, recurse=True):
        """.to() / .cuda() / .half(): the Parameter objects survive, but drop the cached dict anyway (cheap, and correct
        under torch.__future__.set_overwrite_module_params_on_conversion)."""
        out = super()._apply(fn, recurse)
        self._named = None
        return out

    # VGG.py:192-203: level -> (pyramid levels the U-Net has to compute, slice of [x15, x18, x21, x24] that is returned)
    _LEVELS = {3: (3, slice(0, 3)), 4: (4, slice(0, 4)), -1: (3, slice(0, 1)), 2: (3, slice(1, 3))}

    def n_levels(self) -> int:
        """Pyramid levels the U-Net computes (3, or 4 with the full-resolution decoder)."""
        if self.level in self._LEVELS:
            return self._LEVELS[self.level][0]
        raise NotImplementedError("VGGUnet level %r: the reference defines 3, 4, -1 and 2 (VGG.py:192-203)" % (self.level,))

    def level_slice(self) -> slice:
        self.n_levels()
        return self._LEVELS[self.level][1]

    def pyramid(self, x: torch.Tensor, want_conf: bool = True, want_scale: bool = True) -> engine.Pyramid:
        """Engine-layout output (NHWC raw features + lazy L2 scale + confidences) of the levels `level` selects.
        `want_scale=False` leaves the L2-norm scales out (callers whose LM step renormalises: they cancel)."""
        if self._named is None:
            self._named = dict(self.named_parameters())
        p = self._runner(self._named, x, self.n_levels(), want_conf, self.precision, want_scale, g2s=self.G2S)
        sl = self.level_slice()
        return engine.Pyramid(p.feats[sl], p.scales[sl], p.confs[sl])

    def forward(self, x):
        """Reference contract: ([B,C,H,W] L2-normalised features], [B,1,H,W] confidences])."""
        p = self.pyramid(x, want_conf=True)
        feats = [p.nchw(l, normalised=True) for l in range(len(p.feats))]
        confs = [c[:, None] for c in p.confs]
        return feats, confs

    def forward_train(self, x):
        """Train-mode features with the NATIVE backward (engine.VggTrain: tcgen05 forward keeping its activations, tcgen05
        data / weight gradients) where it applies — CUDA, three computed levels, f16x3 — else `forward_autograd`.
        Returns ([L2-normalised NHWC features], [confidences [B,1,H,W]]) for `level`."""
        if not (self.native_train and engine.VggTrain.supports(x, self.n_levels(), self.precision)):
            feats, confs = self.forward_autograd(x)
            return [f.permute(0, 2, 3, 1).contiguous() for f in feats], confs
        if self._named is None:
            self._named = dict(self.named_parameters())
        names = engine.VGG_CONV_NAMES
        params = [self._named[n + ".weight"] for n in names[:engine.N_FEATURE_CONVS]] + \
                 [self._named[n + ".bias"] for n in names[:engine.N_BIASED_CONVS]]
        n = self.n_levels()
        out = engine.VggTrain.apply(self._runner, self._named, n, x, *params)
        raw, confs = out[:n], out[n:]
        feats = []
        for f in raw:                                                  # VGG.py:172-175 / :511-514, on the NHWC tensor
            nrm = f.reshape(f.shape[0], -1).norm(p=2, dim=-1).clamp_min(1e-12)
            feats.append(f / nrm[:, None, None, None])
        sl = self.level_slice()
        return feats[sl], [c[:, None] for c in confs][sl]

    def forward_autograd(self, x):
        """The same network (VGG.py:121-203) evaluated with torch ops through the parameter containers, so that autograd
        reaches the weights: used by `forward(mode='train')` of the LM models until the fused backward
        exists.  Returns the reference's ([L2-normalised features], [confidences]) for `level`."""
        F = torch.nn.functional
        pool = lambda t: F.max_pool2d(t, 2, 2)
        up = lambda t, like: F.interpolate(t, size=like.shape[-2:], mode="nearest")
        x1 = F.relu(self.conv0(x))
        x2 = self.conv2(x1)
        x4 = F.relu(pool(x2))                                   # the reference's in-place ReLU makes the skip post-ReLU
        x7 = self.conv7(F.relu(self.conv5(x4)))
        x9 = F.relu(pool(x7))
        x15 = pool(self.conv14(F.relu(self.conv12(F.relu(self.conv10(x9))))))
        x18 = self.conv_dec1(torch.cat([up(x15, x9), x9], dim=1))
        x21 = self.conv_dec2(torch.cat([up(x18, x4), x4], dim=1))
        feats = [x15, x18, x21]
        if self.n_levels() == 4:
            feats.append(self.conv_dec3(torch.cat([up(x21, x2), x2], dim=1)))
        heads = [self.conf0, self.conf1, self.conf2, self.conf3]
        confs = [torch.sigmoid(-heads[i](f)) for i, f in enumerate(feats)]      # VGG.py:160-163
        sl = self.level_slice()
        return [L2_norm(f) for f in feats][sl], confs[sl]


class VGGUnet_G2S(VGGUnet):
    """VGG.py:206-345: the ground branch of `LM_G2SP --proj nn`.  Same parameters (and state-dict keys) as VGGUnet; every map
    the decoders see is folded from [H, W] to [2H, W/2], so a 256 x 1024 image yields square 64 / 128 / 256 (/ 512) feature
    maps.  Runs ha_vgg_g2s_forward (same tcgen05 kernels, the decoder schedule on the folded geometry).  Levels 3 and 4."""

    G2S = True

    def __init__(self, level):
        super().__init__(level)
        if level not in (3, 4):
            raise NotImplementedError("VGGUnet_G2S: levels 3 and 4 are on the accelerated path")

    def forward_autograd(self, x):
        raise NotImplementedError("VGGUnet_G2S runs in test mode only on the accelerated path")

    def forward_train(self, x):
        raise NotImplementedError("VGGUnet_G2S runs in test mode only on the accelerated path")


def L2_norm(x):
    """VGG.py:511-514."""
    B = x.shape[0]
    return torch.nn.functional.normalize(x.reshape(B, -1), p=2, dim=-1).view(x.shape)
