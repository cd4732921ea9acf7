"""CPU oracle for the cross-view pose-refinement hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional torch-CPU restatement of the reference's algorithm for the one path
this repo accelerates (two-branch VGG U-Net -> iterative LM pose refinement).  It is the checker
the CUDA path is compared against; it is never imported by the product package
(`highlyaccurate_b200/`).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.

Third-party arithmetic: the reference is pure PyTorch; its conv / norm / bmm / 3x3 inverse
arithmetic lives in torch (ATen + oneDNN on CPU), which the reference does not pin.  The image
pins torch 2.11.0+cu128; the oracle calls the same torch CPU operators so both sides execute
the same library arithmetic.

Parity status: PINNED.  `oracle/make_golden.py` imports the unmodified reference from
/root/reference (possible only in the build container), runs it and this file on identical
seeded inputs, asserts agreement, and writes the fixtures in tests/golden/ that
tests/test_oracle_golden.py re-checks everywhere (no /root/reference needed at test time).

Every function cites the reference file:line it restates.  Layouts are the reference's
(NCHW features, [B,1] poses).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- constants
# utils.py:7,11,17,142-146
CAMERA_HEIGHT = 1.65
SAT_PROCESS_SIDE = 512
EPS = 1e-7


def meter_per_pixel_base() -> float:
    """utils.py:142-146 with the default lat=49.015, zoom=18, scale=1."""
    mpp = 156543.03392 * np.cos(49.015 * np.pi / 180.0) / (2 ** 18)
    mpp /= 2
    mpp /= 1.0
    return float(mpp)


@dataclass
class LMArgs:
    """The subset of the argparse Namespace the path reads (train_kitti.py:426-485)."""
    level: int = 3
    N_iters: int = 5
    using_weight: int = 0
    damping: float = 0.1
    train_damping: int = 0
    rotation_range: float = 10.0
    shift_range_lat: float = 20.0
    shift_range_lon: float = 20.0
    use_hessian: int = 0
    level_first: int = 0
    Optimizer: str = "LM"          # 'LM' | 'SGD' | 'ADAM' (LM_S2GP) | 'GN' (LM_S2GP_Ford)
    proj: str = "geo"              # anything else: polar ground table, residual over the whole ground image
    beta1: float = 0.9             # train_kitti.py:480-481
    beta2: float = 0.999


# ----------------------------------------------------------------------------- ground tables
_KITTI_K = [[582.9802, 0.0, 496.2420], [0.0, 482.7076, 125.0034], [0.0, 0.0, 1.0]]
_FORD_K = [945.391406, 0.0, 855.502825, 0.0, 945.668274, 566.372868, 0.0, 0.0, 1.0]


def _lift_to_ground(k0: torch.Tensor, gh: float, gw: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Shared tail of models_kitti.py:663-682 / models_ford.py:133-155: scale K to the level,
    back-project every pixel centre and intersect with the plane y = camera height."""
    k = k0.clone()
    k[:, :1, :] = k0[:, :1, :] * gw / 1024
    k[:, 1:2, :] = k0[:, 1:2, :] * gh / 256
    kinv = torch.inverse(k)
    v, u = torch.meshgrid(torch.arange(0, gh, dtype=torch.float32),
                          torch.arange(0, gw, dtype=torch.float32), indexing="ij")
    uv1 = torch.stack([u, v, torch.ones_like(u)], dim=-1).unsqueeze(0)
    ray = torch.sum(kinv[:, None, None, :, :] * uv1[:, :, :, None, :], dim=-1)
    y = ray[..., 1:2]
    w = CAMERA_HEIGHT / torch.where(torch.abs(y) > EPS, y, EPS * torch.ones_like(y))
    xyz = ray * w
    mask = (xyz[..., -1] > 0).float()
    return xyz[0].contiguous(), mask[0].contiguous()


def kitti_ground_table(level: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """models_kitti.py:622-635,655-682.  Returns xyz[H,W,3], mask[H,W] (fp32)."""
    gh, gw = 256 / (2 ** (3 - level)), 1024 / (2 ** (3 - level))
    k0 = torch.tensor([_KITTI_K], dtype=torch.float32)
    return _lift_to_ground(k0, gh, gw)


def ford_ground_table(level: int, n_levels: int = 3) -> Tuple[torch.Tensor, torch.Tensor]:
    """models_ford.py:41-65,110-155.  `level` indexes the pyramid; args.level==2 uses /4,/2."""
    top = 2 if n_levels == 2 else 3
    gh, gw = 256 / (2 ** (top - level)), 1024 / (2 ** (top - level))
    kfl = torch.tensor(_FORD_K, dtype=torch.float32).reshape(1, 3, 3)
    k0 = torch.zeros_like(kfl)
    k0[0, 0] = kfl[0, 0] / 1656 * 1024
    k0[0, 1] = kfl[0, 1] / 860 * 256
    k0[0, 2] = kfl[0, 2]
    return _lift_to_ground(k0, gh, gw)


def polar_ground_table(level: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """models_kitti.py:684-698 / models_ford.py:156-170 (grd_img2cam_polar), used for every proj != 'geo'."""
    gh, gw = 256 / (2 ** (3 - level)), 1024 / (2 ** (3 - level))
    v, u = torch.meshgrid(torch.arange(0, gh, dtype=torch.float32), torch.arange(0, gw, dtype=torch.float32), indexing="ij")
    theta = u / gw * np.pi / 4
    radius = (1 - v / gh) * 30
    z = radius * torch.cos(np.pi / 4 - theta)
    x = -radius * torch.sin(np.pi / 4 - theta)
    y = CAMERA_HEIGHT * torch.ones_like(z)
    return torch.stack([x, y, z], dim=-1).contiguous(), torch.ones_like(z)


def ground_table(kind: str, level: int, n_levels: int, proj: str = "geo"):
    if proj != "geo":
        return polar_ground_table(level)
    return kitti_ground_table(level) if kind == "kitti" else ford_ground_table(level, n_levels)


# ----------------------------------------------------------------------------- geometry
def kitti_sat_uv(xyz_tab: torch.Tensor, mask_tab: torch.Tensor, su: torch.Tensor, sv: torch.Tensor,
                 th: torch.Tensor, A: int, args: LMArgs, want_jac: bool = True):
    """models_kitti.py:700-801 (grd2cam2world2sat).  su/sv/th are [B,1] normalised poses.
    Returns uv[B,H,W,2], mask[B,H,W], (duv/dsu, duv/dsv, duv/dth) each [B,H,W,2] or Nones."""
    dt = su.dtype
    B = th.shape[0]
    heading = th * args.rotation_range / 180 * np.pi
    shu = su * args.shift_range_lon
    shv = sv * args.shift_range_lat
    c, s = torch.cos(heading), torch.sin(heading)
    z0, o1 = torch.zeros_like(c), torch.ones_like(c)
    R = torch.cat([c, z0, -s, z0, o1, z0, s, z0, c], dim=-1).view(B, 3, 3)
    T0 = torch.cat([shv, CAMERA_HEIGHT * torch.ones_like(shu[:, :1]), -shu], dim=-1)
    T = torch.sum(-R * T0[:, None, :], dim=-1)
    xyz_grd = xyz_tab.to(dt)[None].repeat(B, 1, 1, 1)
    mask = mask_tab.to(dt)[None].repeat(B, 1, 1)
    H, W = xyz_grd.shape[1:3]
    xyz = torch.sum(R[:, None, None, :, :] * xyz_grd[:, :, :, None, :], dim=-1) + T[:, None, None, :]
    R_sat = torch.tensor([0, 0, 1, 1, 0, 0], dtype=dt).reshape(2, 3)
    zx = torch.sum(R_sat[None, None, None, :, :] * xyz[:, :, :, None, :], dim=-1)
    mpp = meter_per_pixel_base()
    mpp *= SAT_PROCESS_SIDE / A
    uv = zx / mpp + A / 2
    if not want_jac:
        return uv, mask, None, None, None
    k = args.rotation_range / 180 * np.pi
    dR = (k * torch.cat([-s, z0, -c, z0, z0, z0, c, z0, -s], dim=-1)).view(B, 3, 3)
    dT0u = args.shift_range_lon * torch.tensor([0.0, 0.0, -1.0], dtype=dt).view(1, 3).repeat(B, 1)
    dT0v = args.shift_range_lat * torch.tensor([1.0, 0.0, 0.0], dtype=dt).view(1, 3).repeat(B, 1)
    dxyz_u = torch.sum(-R * dT0u[:, None, :], dim=-1)[:, None, None, :].repeat(1, H, W, 1)
    dxyz_v = torch.sum(-R * dT0v[:, None, :], dim=-1)[:, None, None, :].repeat(1, H, W, 1)
    dxyz_t = torch.sum(dR[:, None, None, :, :] * xyz_grd[:, :, :, None, :], dim=-1) + \
        torch.sum(-dR * T0[:, None, :], dim=-1)[:, None, None, :]
    proj = lambda d: 1 / mpp * torch.sum(R_sat[None, None, None, :, :] * d[:, :, :, None, :], dim=-1)
    return uv, mask, proj(dxyz_u), proj(dxyz_v), proj(dxyz_t)


def ford_sat_uv(xyz_tab: torch.Tensor, mask_tab: torch.Tensor, R_FL: torch.Tensor, T_FL: torch.Tensor,
                su: torch.Tensor, sv: torch.Tensor, th: torch.Tensor, A: int, side_m: float,
                args: LMArgs, want_jac: bool = True):
    """models_ford.py:173-264 (cam2body2world2sat), depth estimation off."""
    dt = su.dtype
    B = su.shape[0]
    Xc = xyz_tab.to(dt)[None].repeat(B, 1, 1, 1)
    mask = mask_tab.to(dt)[None].repeat(B, 1, 1)
    Xb = torch.sum(R_FL[:, None, None, :, :] * Xc[:, :, :, None, :], dim=-1) + T_FL[:, None, None, :]
    H, W = Xb.shape[1:3]
    um = args.shift_range_lat * su
    vm = args.shift_range_lon * sv
    Tw = torch.cat([vm, -um, torch.zeros_like(vm)], dim=-1)
    yaw = th * args.rotation_range / 180 * np.pi
    c, s = torch.cos(yaw), torch.sin(yaw)
    z0, o1 = torch.zeros_like(c), torch.ones_like(c)
    Rw = torch.cat([c, s, z0, -s, c, z0, z0, z0, o1], dim=-1).view(B, 3, 3)
    Xw = torch.sum(Rw[:, None, None, :, :] * (Xb[:, :, :, None, :] + Tw[:, None, None, None, :]), dim=-1)
    Rs = torch.tensor([0, 1, 0, -1, 0, 0, 0, 0, 1], dtype=dt).reshape(3, 3)[None].repeat(B, 1, 1)
    Xs = torch.sum(Rs[:, None, None, :, :] * Xw[:, :, :, None, :], dim=-1)
    mpp = side_m / A
    uv = Xs[..., :2] / mpp + A // 2
    if not want_jac:
        return uv, mask, None, None, None
    k = args.rotation_range / 180 * np.pi
    dRw = (k * torch.cat([-s, c, z0, -c, -s, z0, z0, z0, z0], dim=-1)).view(B, 3, 3)
    dTu = args.shift_range_lat * torch.tensor([0.0, -1.0, 0.0], dtype=dt).view(1, 3).repeat(B, 1)
    dTv = args.shift_range_lon * torch.tensor([1.0, 0.0, 0.0], dtype=dt).view(1, 3).repeat(B, 1)
    dXw_t = torch.sum(dRw[:, None, None, :, :] * (Xb[:, :, :, None, :] + Tw[:, None, None, None, :]), dim=-1)
    dXw_u = torch.sum(Rw * dTu[:, None, :], dim=-1)
    dXw_v = torch.sum(Rw * dTv[:, None, :], dim=-1)
    dXs_t = torch.sum(Rs[:, None, None, :, :] * dXw_t[:, :, :, None, :], dim=-1)
    dXs_u = torch.sum(Rs * dXw_u[:, None, :], dim=-1)[:, None, None, :].repeat(1, H, W, 1)
    dXs_v = torch.sum(Rs * dXw_v[:, None, :], dim=-1)[:, None, None, :].repeat(1, H, W, 1)
    return uv, mask, dXs_u[..., 0:2] / mpp, dXs_v[..., 0:2] / mpp, dXs_t[..., 0:2] / mpp


# ----------------------------------------------------------------------------- sampler
def bilinear_sample(image: torch.Tensor, uv: torch.Tensor, jac: Optional[torch.Tensor] = None):
    """jacobian.py:138-205 (grid_sample).  image[N,C,IH,IW]; uv[N,H,W,2] in pixel units;
    jac[M,N,H,W,2] = d(uv)/d(pose_m).  Returns out[N,C,H,W], jac_out[M,N,C,H,W] or None."""
    N, C, IH, IW = image.shape
    _, H, W, _ = uv.shape
    x = uv[..., 0].reshape(N, 1, H, W)
    y = uv[..., 1].reshape(N, 1, H, W)
    x0, y0 = torch.floor(x), torch.floor(y)
    # four corners, each clamped independently (jacobian.py:147-166)
    xw = torch.clamp(x0, 0, IW - 1)
    xe = torch.clamp(x0 + 1, 0, IW - 1)
    yn = torch.clamp(y0, 0, IH - 1)
    ys = torch.clamp(y0 + 1, 0, IH - 1)
    m = ((x >= 0) & (x <= IW - 1)) * ((y >= 0) & (y <= IH - 1))
    assert torch.sum(m) > 0
    w_nw = (xe - x) * (ys - y) * m
    w_ne = (x - xw) * (ys - y) * m
    w_sw = (xe - x) * (y - yn) * m
    w_se = (x - xw) * (y - yn) * m
    flat = image.reshape(N, C, IH * IW)

    def tap(yy, xx):
        idx = (yy * IW + xx).long().reshape(N, 1, H * W).expand(N, C, H * W)
        return torch.gather(flat, 2, idx).reshape(N, C, H, W)

    t_nw, t_ne, t_sw, t_se = tap(yn, xw), tap(yn, xe), tap(ys, xw), tap(ys, xe)
    out = t_nw * w_nw + t_ne * w_ne + t_sw * w_sw + t_se * w_se
    if jac is None:
        return out, None
    d_dx = t_nw * (-(ys - y) * m) + t_ne * (ys - y) * m + t_sw * (-(y - yn) * m) + t_se * (y - yn) * m
    d_dy = t_nw * (-(xe - x) * m) + t_ne * (-(x - xw) * m) + t_sw * (xe - x) * m + t_se * (x - xw) * m
    dxy = torch.stack([d_dx, d_dy], dim=-1)
    jac_out = torch.sum(dxy[None] * jac[:, :, None], dim=-1)
    return out, jac_out


# ----------------------------------------------------------------------------- one LM step
@dataclass
class StepStats:
    """Per-sample diagnostics of one LM step (not returned by the reference; used for parity)."""
    hessian: torch.Tensor      # [B,N,N]   J~^T W J~
    grad: torch.Tensor         # [B,N]     J~^T W r
    sat_norm: torch.Tensor     # [B]
    grd_norm: torch.Tensor     # [B]
    res_sq: torch.Tensor       # [B]       ||r||^2
    delta: torch.Tensor        # [B,N]
    n_inrange: int = 0


def lm_update(su, sv, th, sat_proj, grd_feat, grd_conf, dfeat, args: LMArgs, damping: torch.Tensor,
              rand_uv: Optional[Tuple[torch.Tensor, torch.Tensor]], always_3dof: bool = False):
    """models_kitti.py:939-1041 / models_ford.py:380-466 (LM_update), dropout off.
    `damping` is the already-resolved [1,N] (or scalar) lambda.  `rand_uv` are the two [B,1]
    U(-1,1) draws the reference takes from the CPU generator on every 3-DOF call."""
    dof = "full"
    if not always_3dof:
        if args.rotation_range == 0:
            dfeat, dof = dfeat[:2], "shift"
        elif args.shift_range_lat == 0 and args.shift_range_lon == 0:
            dfeat, dof = dfeat[2:], "rot"
    N, B, C, H, W = dfeat.shape
    J = dfeat.reshape(N, B, -1)
    s = sat_proj.reshape(B, -1)
    g = grd_feat.reshape(B, -1)
    sn = torch.maximum(torch.norm(s, p=2, dim=-1), 1e-6 * torch.ones(B, dtype=s.dtype))
    s = s / sn[:, None]
    J = J / sn[None, :, None]
    gn = torch.maximum(torch.norm(g, p=2, dim=-1), 1e-6 * torch.ones(B, dtype=s.dtype))
    g = g / gn[:, None]
    r = s - g
    if args.using_weight:
        w = grd_conf.reshape(B, -1)[:, None, :].repeat(1, C, 1).reshape(B, -1)
    else:
        w = torch.ones([B, g.shape[-1]], dtype=s.dtype)
    Jm = J.permute(1, 2, 0)
    JtW = Jm.transpose(1, 2) * w.unsqueeze(1)
    Hm = JtW @ Jm
    if args.use_hessian:
        Dm = torch.diag_embed(torch.diagonal(Hm, dim1=1, dim2=2))
    else:
        Dm = torch.eye(N, dtype=s.dtype).unsqueeze(0).repeat(B, 1, 1)
    grad = JtW @ r.reshape(B, -1, 1)                      # diagnostic only
    # the reference's expression associates left to right: ((-inv) @ JtW) @ r  (models_kitti.py:1012)
    delta = -torch.inverse(Hm + damping * Dm) @ JtW @ r.reshape(B, -1, 1)
    if dof == "shift":
        su_n, sv_n, th_n = su + delta[:, 0:1, 0], sv + delta[:, 1:2, 0], th
    elif dof == "rot":
        su_n, sv_n, th_n = su, sv, th + delta[:, 0:1, 0]
    else:
        su_n, sv_n, th_n = su + delta[:, 0:1, 0], sv + delta[:, 1:2, 0], th + delta[:, 2:3, 0]
        ru, rv = rand_uv
        su_n = torch.where((su_n > -2.5) & (su_n < 2.5), su_n, ru.to(su_n.dtype))
        sv_n = torch.where((sv_n > -2.5) & (sv_n < 2.5), sv_n, rv.to(sv_n.dtype))
    stats = StepStats(Hm, grad[:, :, 0], sn, gn, torch.sum(r * r, dim=-1), delta[:, :, 0])
    return su_n, sv_n, th_n, stats


def pose_gradient(sat_proj, grd_feat, dfeat):
    """The l2-loss gradient both first-order updates use (models_kitti.py:1072-1078, :1106-1112): sum 2 (s - g) J."""
    r = sat_proj - grd_feat
    return torch.sum((2 * r)[None, ...] * dfeat, dim=[2, 3, 4]).transpose(0, 1)       # [B, 3]


def sgd_update(su, sv, th, sat_proj, grd_feat, dfeat):
    """LM_S2GP.SGD_update, models_kitti.py:1056-1084: all three components, step 0.01, no reset, no RNG draw."""
    d = pose_gradient(sat_proj, grd_feat, dfeat)
    zero = torch.zeros(d.shape[0], dtype=d.dtype)
    return su - 0.01 * d[:, 0:1], sv - 0.01 * d[:, 1:2], th - 0.01 * d[:, 2:3], StepStats(
        torch.zeros(d.shape[0], 3, 3, dtype=d.dtype), d, zero, zero, zero, -0.01 * d)


def adam_update(su, sv, th, sat_proj, grd_feat, dfeat, m, v, t, args: LMArgs):
    """LM_S2GP.ADAM_update, models_kitti.py:1086-1124."""
    d = pose_gradient(sat_proj, grd_feat, dfeat)
    m = args.beta1 * m + (1 - args.beta1) * d
    v = args.beta2 * v + (1 - args.beta2) * (d * d)
    m_hat = m / (1 - args.beta1 ** (t + 1))
    v_hat = v / (1 - args.beta2 ** (t + 1))
    step = m_hat / (v_hat ** 0.5 + 1e-8)
    zero = torch.zeros(d.shape[0], dtype=d.dtype)
    st = StepStats(torch.zeros(d.shape[0], 3, 3, dtype=d.dtype), d, zero, zero, zero, -0.01 * step)
    return su - 0.01 * step[:, 0:1], sv - 0.01 * step[:, 1:2], th - 0.01 * step[:, 2:3], m, v, st


def gn_update(su, sv, th, sat_proj, grd_feat, grd_conf, dfeat, args: LMArgs, rand_uv):
    """LM_S2GP_Ford.GN_update, models_ford.py:534-598: s / ||s|| (no clamp), g as it is, inverse of the undamped
    J^T W J, the LM update's reset rule and RNG draws."""
    N, B, C, H, W = dfeat.shape
    sn = torch.norm(sat_proj.reshape(B, -1), p=2, dim=-1)
    s = sat_proj / sn[:, None, None, None]
    dfeat = dfeat / sn[None, :, None, None, None]
    r = s - grd_feat
    if args.using_weight:
        w = grd_conf.repeat(1, C, 1, 1).reshape(B, C * H * W)
    else:
        w = torch.ones([B, C * H * W], dtype=s.dtype)
    J = dfeat.flatten(start_dim=2).permute(1, 2, 0)
    JtW = J.transpose(1, 2) * w.unsqueeze(1)
    Hm = JtW @ J
    grad = JtW @ r.reshape(B, -1, 1)
    delta = -torch.inverse(Hm) @ JtW @ r.reshape(B, C * H * W, 1)
    su_n, sv_n, th_n = su + delta[:, 0:1, 0], sv + delta[:, 1:2, 0], th + delta[:, 2:3, 0]
    ru, rv = rand_uv
    su_n = torch.where((su_n > -2.5) & (su_n < 2.5), su_n, ru.to(su_n.dtype))
    sv_n = torch.where((sv_n > -2.5) & (sv_n < 2.5), sv_n, rv.to(sv_n.dtype))
    gn = torch.norm(grd_feat.reshape(B, -1), p=2, dim=-1)
    return su_n, sv_n, th_n, StepStats(Hm, grad[:, :, 0], sn, gn, torch.sum(r.reshape(B, -1) ** 2, dim=-1), delta[:, :, 0])


def nnrefine_state_dict(seed: int, dtype=torch.float32) -> dict:
    """Seeded weights with the shapes of RNNs.NNrefine (RNNs.py:98-116); keys as in its state dict."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for i, c in enumerate((256, 128, 64, 16)):
        sd["linear%d.1.weight" % i] = (torch.randn(64, c, 3, 3, generator=g) * (2.0 / (9 * c)) ** 0.5).to(dtype)
        sd["linear%d.1.bias" % i] = (torch.randn(64, generator=g) * 0.01).to(dtype)
    sd["mapping.1.weight"] = (torch.randn(16, 64, generator=g) * 0.5).to(dtype)
    sd["mapping.1.bias"] = (torch.randn(16, generator=g) * 0.1).to(dtype)
    sd["mapping.3.weight"] = (torch.randn(3, 16, generator=g) * 0.1).to(dtype)       # steps of ~0.1: tanh far from saturation
    sd["mapping.3.bias"] = (torch.randn(3, generator=g) * 0.02).to(dtype)
    return sd


def nn_update(su, sv, th, sat_proj, grd_feat, nn_sd: dict):
    """LM_S2GP.NN_update (models_kitti.py:1043-1054) with RNNs.NNrefine.forward (RNNs.py:118-126)."""
    r = sat_proj - grd_feat
    k = {256: 0, 128: 1, 64: 2, 16: 3}[r.shape[1]]
    x = F.conv2d(F.relu(r), nn_sd["linear%d.1.weight" % k].to(r.dtype), nn_sd["linear%d.1.bias" % k].to(r.dtype), padding=1)
    x = torch.mean(x, dim=[2, 3])
    h = F.relu(F.linear(F.relu(x), nn_sd["mapping.1.weight"].to(r.dtype), nn_sd["mapping.1.bias"].to(r.dtype)))
    y = torch.tanh(F.linear(h, nn_sd["mapping.3.weight"].to(r.dtype), nn_sd["mapping.3.bias"].to(r.dtype)))
    zero = torch.zeros(y.shape[0], dtype=y.dtype)
    return su + y[:, 0:1], sv + y[:, 1:2], th + y[:, 2:3], StepStats(torch.zeros(y.shape[0], 3, 3, dtype=y.dtype), y, zero, zero, zero, y)


def resolve_damping(args: LMArgs, damping_param: Optional[torch.Tensor], n_dof: int, dtype=torch.float32):
    """models_kitti.py:958-966: trained lambda = 10^(-6 + 11*sigmoid(p)), else args.damping."""
    if args.train_damping:
        return 10.0 ** (-6 + damping_param.to(dtype).sigmoid() * 11)
    return args.damping * torch.ones(size=(1, n_dof), dtype=dtype)


def draw_reset(B: int):
    """models_kitti.py:1028-1029: two [B,1] draws from the default CPU generator per step."""
    u = torch.distributions.uniform.Uniform(-1, 1).sample([B, 1])
    v = torch.distributions.uniform.Uniform(-1, 1).sample([B, 1])
    return u, v


def lm_one_step(kind: str, sf, gf, gc, tab, su, sv, th, args: LMArgs, lam, draws, ford: Optional[dict] = None,
                adam: Optional[dict] = None):
    """One (iteration, level) body of the reference's forward loop: project_map_to_grd ->
    masking / bottom-half crop (proj 'geo' only) -> the update of args.Optimizer
    (models_kitti.py:1187-1260 / models_ford.py:700-800).  `adam` = dict(m, v, t), updated in place."""
    B = sf.shape[0]
    dt = sf.dtype
    A = sf.shape[-1]
    if kind == "kitti":
        uv, mask, ju, jv, jt = kitti_sat_uv(tab[0], tab[1], su, sv, th, A, args)
    else:
        uv, mask, ju, jv, jt = ford_sat_uv(tab[0], tab[1], ford["R_FL"].to(dt), ford["T_FL"].to(dt),
                                           su, sv, th, A, ford["side_m"], args)
    jac = torch.stack([ju, jv, jt], dim=0)
    sp, dj = bilinear_sample(sf, uv, jac)                       # models_kitti.py:924
    sp = sp * mask[:, None]                                     # :927
    dj = dj * mask[None, :, None]                               # :929
    gfm = gf * mask[:, None]                                    # :1191
    gcm = (gc * mask[:, None]) if gc is not None else torch.ones(B, 1, *gf.shape[-2:], dtype=dt) * mask[:, None]
    h2 = gf.shape[-2] // 2 if args.proj == "geo" else 0         # :1194-1205 bottom half only for the ground-plane lift
    sp, gfm, gcm, dj = sp[:, :, h2:], gfm[:, :, h2:], gcm[:, :, h2:], dj[:, :, :, h2:]
    if args.Optimizer == "SGD":
        return sgd_update(su, sv, th, sp, gfm, dj)
    if args.Optimizer == "ADAM":
        su, sv, th, adam["m"], adam["v"], st = adam_update(su, sv, th, sp, gfm, dj, adam["m"], adam["v"], adam["t"], args)
        return su, sv, th, st
    if args.Optimizer == "GN":
        return gn_update(su, sv, th, sp, gfm, gcm, dj, args, draws)
    if args.Optimizer == "NN":
        return nn_update(su, sv, th, sp, gfm, adam["nn"])
    return lm_update(su, sv, th, sp, gfm, gcm, dj, args, lam, draws, always_3dof=(kind == "ford"))


def n_dof(kind: str, args: LMArgs) -> int:
    if kind == "ford" or not (args.rotation_range == 0 or (args.shift_range_lat == 0 and args.shift_range_lon == 0)):
        return 3
    return 2 if args.rotation_range == 0 else 1


# ----------------------------------------------------------------------------- LM loops
@dataclass
class LoopResult:
    lats: torch.Tensor            # [B,N_iters,L]  (KITTI: sv ; Ford: su)
    lons: torch.Tensor            # [B,N_iters,L]  (KITTI: su ; Ford: sv)
    thetas: torch.Tensor          # [B,N_iters,L]
    stats: List[List[Optional[StepStats]]] = field(default_factory=list)   # [iter][level]
    pose_in: List[List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]] = field(default_factory=list)


def _step_order(n_iters: int, n_levels: int, level_first: int):
    """models_kitti.py:1176-1180 (iter-first) vs :1349-1353 (level-first)."""
    if level_first:
        return [(it, lv) for lv in range(n_levels) for it in range(n_iters)]
    return [(it, lv) for it in range(n_iters) for lv in range(n_levels)]


def lm_loop(kind: str, sat_feats: Sequence[torch.Tensor], grd_feats: Sequence[torch.Tensor],
            grd_confs: Sequence[Optional[torch.Tensor]], args: LMArgs,
            damping_param: Optional[torch.Tensor] = None, reset_draws=None,
            ford: Optional[dict] = None, pose0=None, nn_sd: Optional[dict] = None) -> LoopResult:
    """models_kitti.py:1141-1316 / :1318-1492 and models_ford.py:652-866 / :868-1026, mode='test'.
    kind: 'kitti' | 'ford'.  ford = dict(R_FL[B,3,3], T_FL[B,3], side_m).  `reset_draws`
    (optional) = list of (u,v) per executed step; default draws them from the CPU generator in
    the reference's order."""
    L = len(sat_feats)
    B = sat_feats[0].shape[0]
    dt = sat_feats[0].dtype
    if pose0 is None:
        su = torch.zeros([B, 1], dtype=dt); sv = torch.zeros([B, 1], dtype=dt); th = torch.zeros([B, 1], dtype=dt)
    else:
        su, sv, th = [p.clone().to(dt) for p in pose0]
    always3 = kind == "ford"
    if always3 or not (args.rotation_range == 0 or (args.shift_range_lat == 0 and args.shift_range_lon == 0)):
        ndof = 3
    else:
        ndof = 2 if args.rotation_range == 0 else 1
    lam = resolve_damping(args, damping_param, ndof, dt)
    draws_rng = ndof == 3 and args.Optimizer in ("LM", "GN")     # SGD_update / ADAM_update draw nothing
    adam = dict(m=0, v=0, t=0, nn=nn_sd)                        # per-run optimiser state (Adam moments, NNrefine weights)
    tabs = []
    for lv in range(L):
        tabs.append(tuple(t.to(dt) for t in ground_table(kind, lv, L, args.proj)))
    rec_u = torch.zeros(B, args.N_iters, L, dtype=dt)
    rec_v = torch.zeros(B, args.N_iters, L, dtype=dt)
    rec_t = torch.zeros(B, args.N_iters, L, dtype=dt)
    stats = [[None] * L for _ in range(args.N_iters)]
    pin = [[None] * L for _ in range(args.N_iters)]
    for k, (it, lv) in enumerate(_step_order(args.N_iters, L, args.level_first)):
        draws = None
        if draws_rng:
            draws = reset_draws[k] if reset_draws is not None else draw_reset(B)
        pin[it][lv] = (su.clone(), sv.clone(), th.clone())
        adam["t"] = it * args.level + lv                          # models_kitti.py:1241 (it multiplies by args.level)
        if adam["t"] == 0:
            adam["m"], adam["v"] = 0, 0                           # :1242-1244
        su, sv, th, st = lm_one_step(kind, sat_feats[lv], grd_feats[lv], grd_confs[lv], tabs[lv], su, sv, th, args, lam,
                                     draws, ford, adam)
        stats[it][lv] = st
        rec_u[:, it, lv], rec_v[:, it, lv], rec_t[:, it, lv] = su[:, 0], sv[:, 0], th[:, 0]
    if kind == "kitti":      # models_kitti.py:1281-1283: lats = shift_v, lons = shift_u
        return LoopResult(rec_v, rec_u, rec_t, stats, pin)
    return LoopResult(rec_u, rec_v, rec_t, stats, pin)               # models_ford.py:823-825



# ----------------------------------------------------------------------------- G2SP (ground -> satellite plane)
def g2sp_sat_table(A: int, dtype=torch.float32) -> torch.Tensor:
    """models_kitti.py:54-84 (get_warp_sat2real): homogeneous ground-plane point of every satellite
    pixel, [A, A, 4] = (X south, Y = 0, Z east, 1); the map centre is A // 2."""
    i = torch.arange(0, A)
    ii, jj = torch.meshgrid(i, i, indexing="ij")
    uv = torch.stack([jj, ii], dim=-1).float()
    uv_c = uv - torch.tensor([A // 2, A // 2])
    mpp = meter_per_pixel_base()
    mpp *= SAT_PROCESS_SIDE / A
    aff = mpp * torch.tensor([[0, 1], [1, 0]]).float()
    xz = torch.einsum("ij, hwj -> hwi", aff, uv_c)
    y = torch.zeros_like(xz[..., 0:1])
    return torch.cat([xz[:, :, :1], y, xz[:, :, 1:], torch.ones_like(y)], dim=-1).to(dtype)


def g2sp_cam_uv(xyz1: torch.Tensor, su, sv, th, cam_k: torch.Tensor, gh: int, gw: int, ori_h: int, ori_w: int,
                args: LMArgs):
    """models_kitti.py:86-160 (seq_warp_real2camera): project the satellite-plane points into the
    ground camera, P = K_l [R(-heading) | T], perspective divide with max(w, 1e-6), quotient-rule
    Jacobians zeroed where w <= 1e-6.  Returns uv[B,A,A,2], (duv/dsu, duv/dsv, duv/dth), mask[B,A,A,1]."""
    dt = su.dtype
    B = th.shape[0]
    shu = args.shift_range_lon * su
    shv = args.shift_range_lat * sv
    heading = th * args.rotation_range / 180 * np.pi
    c, s = torch.cos(-heading), torch.sin(-heading)
    z0, o1 = torch.zeros_like(c), torch.ones_like(c)
    R = torch.cat([c, z0, -s, z0, o1, z0, s, z0, c], dim=-1).view(B, 3, 3)
    T = torch.cat([shv, CAMERA_HEIGHT * torch.ones_like(shu), -shu], dim=-1).unsqueeze(-1)
    k = cam_k.clone()
    k[:, :1, :] = cam_k[:, :1, :] * gw / ori_w
    k[:, 1:2, :] = cam_k[:, 1:2, :] * gh / ori_h
    P = k @ torch.cat([R, T], dim=-1)
    uv1 = torch.sum(P[:, None, None, :, :] * xyz1[None, :, :, None, :], dim=-1)
    w = torch.maximum(uv1[..., 2:], torch.ones_like(uv1[..., 2:]) * 1e-6)
    uv = uv1[..., :2] / w
    mask = torch.greater(w, torch.ones_like(w) * 1e-6)
    kk = args.rotation_range / 180 * np.pi
    dT_u = args.shift_range_lon * torch.tensor([0.0, 0.0, -1.0], dtype=dt).view(1, 3, 1).repeat(B, 1, 1)
    dT_v = args.shift_range_lat * torch.tensor([1.0, 0.0, 0.0], dtype=dt).view(1, 3, 1).repeat(B, 1, 1)
    dR = (kk * torch.cat([s, z0, c, z0, z0, z0, -c, z0, s], dim=-1)).view(B, 3, 3)
    Rz, Tz = torch.zeros(B, 3, 3, dtype=dt), torch.zeros(B, 3, 1, dtype=dt)
    out = []
    for dP in (k @ torch.cat([Rz, dT_u], dim=-1), k @ torch.cat([Rz, dT_v], dim=-1), k @ torch.cat([dR, Tz], dim=-1)):
        d1 = torch.sum(dP[:, None, None, :, :] * xyz1[None, :, :, None, :], dim=-1)
        d = d1[..., 0:2] / w - uv1[..., :2] * d1[..., 2:] / (w ** 2)
        out.append(torch.where(mask, d, torch.zeros_like(d)))
    return uv, out[0], out[1], out[2], mask


def g2sp_inplane_uv(A: int, su, sv, th, args: LMArgs):
    """models_kitti.py:289-331 (inplane_grd_to_map, proj == 'nn'): the ground features live on a square map of the
    satellite's size already (VGGUnet_G2S), so the warp is an in-plane rotation about the map centre A / 2 plus a shift in
    pixels.  Returns uv[B,A,A,2] and (duv/dsu, duv/dsv, duv/dth); the mask is all ones."""
    dt = su.dtype
    B = th.shape[0]
    mpp = meter_per_pixel_base()
    mpp *= SAT_PROCESS_SIDE / A
    shu_px = args.shift_range_lon * su / mpp
    shv_px = args.shift_range_lat * sv / mpp
    T = torch.cat([-shu_px, shv_px], dim=-1)
    heading = th * args.rotation_range / 180 * np.pi
    c, s = torch.cos(heading), torch.sin(heading)
    R = torch.cat([c, -s, s, c], dim=-1).view(B, 2, 2)
    i = torch.arange(0, A)
    vv, uu = torch.meshgrid(i, i, indexing="ij")
    uv2 = torch.stack([uu, vv], dim=-1).unsqueeze(0).repeat(B, 1, 1, 1).to(dt)
    uv2 = uv2 - A / 2
    uv = torch.einsum("bij, bhwj->bhwi", R, uv2) + T[:, None, None, :] + A / 2
    ju = (args.shift_range_lon / mpp * torch.tensor([-1.0, 0.0], dtype=dt).view(1, 2).repeat(B, 1))[:, None, None, :].repeat(1, A, A, 1)
    jv = (args.shift_range_lat / mpp * torch.tensor([0.0, 1.0], dtype=dt).view(1, 2).repeat(B, 1))[:, None, None, :].repeat(1, A, A, 1)
    dR = args.rotation_range / 180 * np.pi * torch.cat([-s, -c, c, -s], dim=-1).view(B, 2, 2)
    jt = torch.einsum("bij, bhwj->bhwi", dR, uv2)
    return uv, ju, jv, jt


def g2sp_lm_update(su, sv, th, grd_proj, grd_conf_proj, sat_feat, dfeat, args: LMArgs, damping: torch.Tensor):
    """models_kitti.py:333-379 (LM_G2SP.LM_update): r = grd_proj - sat, NO renormalisation, identity
    damping, always 3 DOF, no out-of-range reset."""
    N, B, C, H, W = dfeat.shape
    r = grd_proj - sat_feat
    if args.using_weight:
        w = grd_conf_proj.repeat(1, C, 1, 1).reshape(B, C * H * W)
    else:
        w = torch.ones([B, C * H * W], dtype=sat_feat.dtype)
    J = dfeat.flatten(start_dim=2).permute(1, 2, 0)
    JtW = J.transpose(1, 2) * w.unsqueeze(1)
    Hm = JtW @ J
    eye = torch.eye(N, dtype=sat_feat.dtype).unsqueeze(0).repeat(B, 1, 1)
    grad = JtW @ r.reshape(B, C * H * W, 1)
    delta = -torch.inverse(Hm + damping * eye) @ JtW @ r.reshape(B, C * H * W, 1)
    st = StepStats(Hm, grad[:, :, 0], torch.zeros(B), torch.zeros(B), torch.sum(r.reshape(B, -1) ** 2, dim=-1), delta[:, :, 0])
    return su + delta[:, 0:1, 0], sv + delta[:, 1:2, 0], th + delta[:, 2:, 0], st


def g2sp_one_step(sf, gf, gc, cam_k, su, sv, th, ori_h: int, ori_w: int, args: LMArgs, lam):
    """One (iteration, level) body of LM_G2SP.forward (models_kitti.py:432-470):
    project_grd_to_map (:163-177,276-287) -> LM_update."""
    A = sf.shape[-1]
    gh, gw = gf.shape[-2:]
    if args.proj == "nn":                                   # models_kitti.py:232-233
        uv, ju, jv, jt = g2sp_inplane_uv(A, su, sv, th, args)
    else:
        xyz1 = g2sp_sat_table(A, sf.dtype)
        uv, ju, jv, jt, mask = g2sp_cam_uv(xyz1, su, sv, th, cam_k.to(sf.dtype), gh, gw, ori_h, ori_w, args)
    gp, dj = bilinear_sample(gf, uv, torch.stack([ju, jv, jt], dim=0))
    gcp = bilinear_sample(gc, uv)[0] if gc is not None else None
    return g2sp_lm_update(su, sv, th, gp, gcp, sf, dj, args, lam)


def lm_loop_g2sp(sat_feats, grd_feats, grd_confs, cam_k, args: LMArgs, damping_param=None, ori_hw=(256, 1024),
                 pose0=None) -> LoopResult:
    """models_kitti.py:381-499 (LM_G2SP.forward, mode='test'), iteration-first only."""
    L = len(sat_feats)
    B = sat_feats[0].shape[0]
    dt = sat_feats[0].dtype
    if pose0 is None:
        su = torch.zeros([B, 1], dtype=dt); sv = torch.zeros([B, 1], dtype=dt); th = torch.zeros([B, 1], dtype=dt)
    else:
        su, sv, th = [p.clone().to(dt) for p in pose0]
    lam = damping_param.to(dt) if args.train_damping else args.damping * torch.ones(1, 3, dtype=dt)   # :356-359
    rec = torch.zeros(3, B, args.N_iters, L, dtype=dt)
    stats = [[None] * L for _ in range(args.N_iters)]
    pin = [[None] * L for _ in range(args.N_iters)]
    for it in range(args.N_iters):
        for lv in range(L):
            pin[it][lv] = (su.clone(), sv.clone(), th.clone())
            su, sv, th, st = g2sp_one_step(sat_feats[lv], grd_feats[lv], grd_confs[lv], cam_k, su, sv, th, ori_hw[0],
                                           ori_hw[1], args, lam)
            stats[it][lv] = st
            rec[0][:, it, lv], rec[1][:, it, lv], rec[2][:, it, lv] = su[:, 0], sv[:, 0], th[:, 0]
    return LoopResult(rec[1], rec[0], rec[2], stats, pin)           # :472-474: lats = shift_v, lons = shift_u


# ----------------------------------------------------------------------------- VGG U-Net
def l2_norm(x: torch.Tensor) -> torch.Tensor:
    """VGG.py:511-514: per-sample L2 normalisation over C*H*W (F.normalize, eps 1e-12)."""
    B = x.shape[0]
    return F.normalize(x.reshape(B, -1), p=2, dim=-1).view(x.shape)


def vgg_unet(sd: dict, x: torch.Tensor, level: int, prefix: str = "", keep: Optional[dict] = None):
    """VGG.py:121-203 (VGGUnet.forward, estimate_depth off) driven by a state-dict.
    Returns ([features], [confidences]) for `level` in {3,4,-1,2}.  `keep` (optional dict)
    receives the intermediate activations by the reference's variable names."""
    w = lambda n: sd[prefix + n]
    conv = lambda t, n, bias=True: F.conv2d(t, w(n + ".weight"), w(n + ".bias") if bias else None, padding=1)
    pool = lambda t: F.max_pool2d(t, 2, 2)
    x1 = F.relu(conv(x, "conv0"))
    x2 = conv(x1, "conv2")
    x4 = F.relu(pool(x2))                       # x3 is overwritten in place by ReLU (VGG.py:126-128)
    x7 = conv(F.relu(conv(x4, "conv5")), "conv7")
    x9 = F.relu(pool(x7))                       # == x8 after the in-place ReLU
    x14 = conv(F.relu(conv(F.relu(conv(x9, "conv10")), "conv12")), "conv14")
    x15 = pool(x14)
    up = lambda t, ref: F.interpolate(t, [ref.shape[2], ref.shape[3]], mode="nearest")
    dec = lambda t, n: F.conv2d(F.relu(F.conv2d(F.relu(t), w(n + ".1.weight"), None, padding=1)),
                                w(n + ".3.weight"), None, padding=1)
    x18 = dec(torch.cat([up(x15, x9), x9], dim=1), "conv_dec1")
    x21 = dec(torch.cat([up(x18, x4), x4], dim=1), "conv_dec2")
    feats, confs = [x15, x18, x21], []
    if level == 4:
        x24 = dec(torch.cat([up(x21, x2), x2], dim=1), "conv_dec3")
        feats.append(x24)
    for i, f in enumerate(feats):
        c = torch.sigmoid(F.conv2d(F.relu(f), w("conf%d.1.weight" % i), None, padding=1))
        confs.append(torch.sigmoid(-c))        # VGG.py:160-163
    if keep is not None:
        keep.update(x1=x1, x2=x2, x4=x4, x7=x7, x9=x9, x14=x14, x15=x15, x18=x18, x21=x21)
    feats = [l2_norm(f) for f in feats]
    if level in (3, 4):
        return feats, confs
    if level == -1:
        return feats[:1], confs[:1]
    if level == 2:
        return feats[1:3], confs[1:3]
    raise ValueError("unsupported level %r" % level)


def vgg_unet_g2s(sd: dict, x: torch.Tensor, level: int, prefix: str = ""):
    """VGG.py:275-345 (VGGUnet_G2S.forward): the encoder of VGGUnet, but every skip / bottleneck map is re-shaped from
    [H, W] to [2H, W/2] (a reinterpretation of the row-major pixel order, :283-299) before the decoders run on it; the
    features come out square (64 / 128 / 256 / 512 for a 256 x 1024 image).  c0 is computed on the un-reshaped x15 (:326)."""
    w = lambda n: sd[prefix + n]
    conv = lambda t, n, bias=True: F.conv2d(t, w(n + ".weight"), w(n + ".bias") if bias else None, padding=1)
    pool = lambda t: F.max_pool2d(t, 2, 2)
    fold = lambda t: t.reshape(t.shape[0], t.shape[1], t.shape[2] * 2, t.shape[3] // 2)
    x1 = F.relu(conv(x, "conv0"))
    x2 = conv(x1, "conv2")
    x4 = F.relu(pool(x2))                       # x3 / x3_ share storage with the in-place ReLU's output (:281,289)
    x7 = conv(F.relu(conv(x4, "conv5")), "conv7")
    x9 = F.relu(pool(x7))
    x15 = pool(conv(F.relu(conv(F.relu(conv(x9, "conv10")), "conv12")), "conv14"))
    x2_, x3_, x8_, x15_ = fold(x2), fold(x4), fold(x9), fold(x15)
    up = lambda t, ref: F.interpolate(t, [ref.shape[2], ref.shape[3]], mode="nearest")
    dec = lambda t, n: F.conv2d(F.relu(F.conv2d(F.relu(t), w(n + ".1.weight"), None, padding=1)),
                                w(n + ".3.weight"), None, padding=1)
    x18 = dec(torch.cat([up(x15_, x8_), x8_], dim=1), "conv_dec1")
    x21 = dec(torch.cat([up(x18, x3_), x3_], dim=1), "conv_dec2")
    feats, heads = [x15_, x18, x21], [x15, x18, x21]
    if level == 4:
        x24 = dec(torch.cat([up(x21, x2_), x2_], dim=1), "conv_dec3")
        feats.append(x24); heads.append(x24)
    confs = [torch.sigmoid(-torch.sigmoid(F.conv2d(F.relu(h), w("conf%d.1.weight" % i), None, padding=1))) for i, h in enumerate(heads)]
    feats = [l2_norm(f) for f in feats]
    if level in (3, 4):
        return feats, confs
    raise ValueError("unsupported level %r" % level)


def vgg_state_dict(seed: int, prefix: str = "", dtype=torch.float32, scale: float = 1.0) -> dict:
    """Deterministic random weights with the reference's 24 per-U-Net keys/shapes (VGG.py:23-81).
    He-normal so activations keep O(1) scale through the 11-conv stack."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def mk(name, co, ci, bias):
        std = scale * math.sqrt(2.0 / (9 * ci))
        sd[prefix + name + ".weight"] = (torch.randn(co, ci, 3, 3, generator=g) * std).to(dtype)
        if bias:
            sd[prefix + name + ".bias"] = (torch.randn(co, generator=g) * 0.05).to(dtype)

    for name, co, ci in [("conv0", 64, 3), ("conv2", 64, 64), ("conv5", 128, 64), ("conv7", 128, 128),
                         ("conv10", 256, 128), ("conv12", 256, 256), ("conv14", 256, 256)]:
        mk(name, co, ci, True)
    for name, co, ci in [("conv_dec1.1", 128, 384), ("conv_dec1.3", 128, 128), ("conv_dec2.1", 64, 192),
                         ("conv_dec2.3", 64, 64), ("conv_dec3.1", 32, 128), ("conv_dec3.3", 16, 32),
                         ("conf0.1", 1, 256), ("conf1.1", 1, 128), ("conf2.1", 1, 64), ("conf3.1", 1, 16)]:
        mk(name, co, ci, False)
    return sd


# ----------------------------------------------------------------------------- whole forwards
def forward_kitti(sd: dict, sat_map: torch.Tensor, grd_img: torch.Tensor, args: LMArgs, reset_draws=None) -> LoopResult:
    """models_kitti.py:1126-1316, mode='test': returns the full trajectory; the reference's
    return value is (lats[:, -1, -1], lons[:, -1, -1], thetas[:, -1, -1])."""
    sf, _ = vgg_unet(sd, sat_map, args.level, "SatFeatureNet.")
    gf, gc = vgg_unet(sd, grd_img, args.level, "GrdFeatureNet.")
    return lm_loop("kitti", sf, gf, gc, args, sd.get("damping"), reset_draws)


def forward_ford(sd: dict, sat_map, grd_img, side_m: float, R_FL, T_FL, args: LMArgs, reset_draws=None) -> LoopResult:
    """models_ford.py:1028-1036,652-866, mode='test'."""
    sf, _ = vgg_unet(sd, sat_map, args.level, "SatFeatureNet.")
    gf, gc = vgg_unet(sd, grd_img, args.level, "GrdFeatureNet.")
    return lm_loop("ford", sf, gf, gc, args, sd.get("damping"), reset_draws,
                   ford=dict(R_FL=R_FL, T_FL=T_FL, side_m=side_m))


# ----------------------------------------------------------------------------- synthetic inputs
def smooth_pyramid(B: int, A: int, n_levels: int, seed: int, dtype=torch.float32):
    """smooth random satellite features, bicubic x8 upsampled noise."""
    g = torch.Generator().manual_seed(seed)
    chans = [256, 128, 64, 16][:n_levels]
    out = []
    for lv, C in enumerate(chans):
        a = A // (2 ** (3 - lv))
        base = torch.randn(B, C, max(a // 8, 2), max(a // 8, 2), generator=g)
        out.append(F.interpolate(base, size=(a, a), mode="bicubic", align_corners=True).to(dtype).contiguous())
    return out


def planted_case(kind: str, B: int, A: int, n_levels: int, seed: int, gt, args: LMArgs, ford: Optional[dict] = None,
                 l2: bool = False):
    """KAT-4: ground features are the satellite features warped at a planted pose `gt`
    ([B,3] = su,sv,th), so the LM loop contracts onto gt.  `l2`: both pyramids L2-normalised per sample like the
    U-Net outputs (VGG.py:172-175) — the first-order updates take the features as they are."""
    sat = smooth_pyramid(B, A, n_levels, seed)
    if l2:
        sat = [l2_norm(x) for x in sat]
    gt = torch.as_tensor(gt, dtype=torch.float32).reshape(B, 3)
    su, sv, th = gt[:, 0:1], gt[:, 1:2], gt[:, 2:3]
    grd = []
    for lv in range(n_levels):
        tab = ground_table(kind, lv, n_levels, args.proj)
        if kind == "kitti":
            uv, mask, *_ = kitti_sat_uv(tab[0], tab[1], su, sv, th, sat[lv].shape[-1], args, want_jac=False)
        else:
            uv, mask, *_ = ford_sat_uv(tab[0], tab[1], ford["R_FL"], ford["T_FL"], su, sv, th, sat[lv].shape[-1],
                                       ford["side_m"], args, want_jac=False)
        f, _ = bilinear_sample(sat[lv], uv)
        f = (f * mask[:, None]).contiguous()
        grd.append(l2_norm(f) if l2 else f)
    return sat, grd


def random_pyramid(B: int, A: int, n_levels: int, seed: int):
    """Unstructured (non-contractive) features for per-step parity: randn, L2-normalised."""
    g = torch.Generator().manual_seed(seed)
    chans = [256, 128, 64, 16][:n_levels]
    sat, grd, conf = [], [], []
    for lv, C in enumerate(chans):
        a = A // (2 ** (3 - lv))
        h, w = 256 // (2 ** (3 - lv)), 1024 // (2 ** (3 - lv))
        sat.append(l2_norm(torch.randn(B, C, a, a, generator=g)))
        grd.append(l2_norm(torch.randn(B, C, h, w, generator=g)))
        conf.append(torch.sigmoid(-torch.sigmoid(torch.randn(B, 1, h, w, generator=g))))
    return sat, grd, conf
