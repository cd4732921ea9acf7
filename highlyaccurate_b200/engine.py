"""Host side of the pose-refinement engine: everything between the reference-shaped
`nn.Module.forward()` mirrors (models_kitti.py / models_ford.py in this package) and the C ABI
(include/ha_b200.h).  PyTorch is used for device memory, streams and the CPU RNG only — all
arithmetic of the hot path happens inside libha_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import HaLevel, HaLmParams, HaVggGrads, HaVggStateDict, check

CAMERA_HEIGHT = 1.65           # utils.py:7
SAT_PROCESS_SIDE = 512         # utils.py:11
GROUND_EPS = 1e-7              # utils.py:17
PYRAMID_CHANNELS = (256, 128, 64, 16)
KITTI_K = ((582.9802, 0.0, 496.2420), (0.0, 482.7076, 125.0034), (0.0, 0.0, 1.0))   # models_kitti.py:657-659
FORD_K_FL = (945.391406, 0.0, 855.502825, 0.0, 945.668274, 566.372868, 0.0, 0.0, 1.0)  # models_ford.py:116-118


def kitti_meter_per_pixel() -> float:
    """utils.py:142-146 at the defaults (lat 49.015, zoom 18, scale 1) in python doubles."""
    mpp = 156543.03392 * np.cos(49.015 * np.pi / 180.0) / (2 ** 18)
    return float(mpp / 2 / 1.0)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise _lib.HaError("%s must be a CUDA tensor: the engine has no CPU path" % what)


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------- ground tables
def polar_ground_table(level: int) -> torch.Tensor:
    """The ground table of every `proj` other than 'geo' (models_kitti.py:684-698 / models_ford.py:156-170,
    grd_img2cam_polar): image column -> bearing over a 45 degree fan, image row -> range up to 30 m, every pixel valid.
    [H, W, 4] fp32 = (x, y, z, 1), same arithmetic as the reference so the values are identical."""
    gh, gw = 256 / (2 ** (3 - level)), 1024 / (2 ** (3 - level))
    vv, uu = torch.meshgrid(torch.arange(0, gh, dtype=torch.float32), torch.arange(0, gw, dtype=torch.float32), indexing="ij")
    bearing = uu / gw * np.pi / 4
    rng = (1 - vv / gh) * 30
    z = rng * torch.cos(np.pi / 4 - bearing)
    x = -rng * torch.sin(np.pi / 4 - bearing)
    y = CAMERA_HEIGHT * torch.ones_like(z)
    return torch.stack([x, y, z, torch.ones_like(z)], dim=-1).contiguous()


def ground_table(kind: str, level: int, n_levels: int = 3, proj: str = "geo") -> torch.Tensor:
    """Ground-plane lift of every ground-image pixel at a pyramid level, in the camera frame:
    [H, W, 4] fp32 = (x, y, z, mask).  Init-time CPU work, same arithmetic as
    models_kitti.py:655-682 (grd_img2cam) / models_ford.py:110-155 so the values are identical."""
    if proj != "geo":
        return polar_ground_table(level)
    if kind == "kitti":
        top = 3
        k0 = torch.tensor([KITTI_K], dtype=torch.float32)
    elif kind == "ford":
        top = 2 if n_levels == 2 else 3
        raw = torch.tensor(FORD_K_FL, dtype=torch.float32).reshape(1, 3, 3)
        k0 = torch.zeros_like(raw)
        k0[0, 0] = raw[0, 0] / 1656 * 1024        # models_ford.py:121-130: sensor 1656x860 -> 1024x256
        k0[0, 1] = raw[0, 1] / 860 * 256
        k0[0, 2] = raw[0, 2]
    else:
        raise ValueError(kind)
    gh, gw = 256 / (2 ** (top - level)), 1024 / (2 ** (top - level))
    k = k0.clone()
    k[:, :1, :] = k0[:, :1, :] * gw / 1024
    k[:, 1:2, :] = k0[:, 1:2, :] * gh / 256
    kinv = torch.inverse(k)
    vv, uu = torch.meshgrid(torch.arange(0, gh, dtype=torch.float32), torch.arange(0, gw, dtype=torch.float32),
                            indexing="ij")
    pix = torch.stack([uu, vv, torch.ones_like(uu)], dim=-1).unsqueeze(0)
    ray = torch.sum(kinv[:, None, None, :, :] * pix[:, :, :, None, :], dim=-1)
    ry = ray[..., 1:2]
    depth = CAMERA_HEIGHT / torch.where(torch.abs(ry) > GROUND_EPS, ry, GROUND_EPS * torch.ones_like(ry))
    xyz = ray * depth
    mask = (xyz[..., -1:] > 0).float()
    return torch.cat([xyz, mask], dim=-1)[0].contiguous()


# ------------------------------------------------------------------------------- pyramids
@dataclass
class Pyramid:
    """One branch's feature pyramid in the engine's layout: NHWC fp32, raw (not L2-normalised)
    features plus the per-sample 1/||x|| the reference would have applied (VGG.py:172-175)."""
    feats: List[torch.Tensor]
    scales: List[Optional[torch.Tensor]]
    confs: List[Optional[torch.Tensor]] = field(default_factory=list)    # [B,H,W] each

    @property
    def batch(self) -> int:
        return self.feats[0].shape[0]

    @staticmethod
    def from_nchw(feats: Sequence[torch.Tensor], confs: Optional[Sequence[Optional[torch.Tensor]]] = None) -> "Pyramid":
        """Reference-layout ([B,C,H,W], already normalised) features -> engine layout."""
        out = [nchw_to_nhwc(f) for f in feats]
        cf = [None if c is None else c.reshape(c.shape[0], c.shape[-2], c.shape[-1]).contiguous().float()
              for c in (confs or [None] * len(out))]
        return Pyramid(out, [None] * len(out), cf)

    def nchw(self, level: int, normalised: bool = True) -> torch.Tensor:
        f = nhwc_to_nchw(self.feats[level])
        if normalised and self.scales[level] is not None:
            f = f * self.scales[level][:, None, None, None]
        return f


def nchw_to_nhwc(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, "nchw_to_nhwc input")
    x = x.contiguous().float()
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, Cc, device=x.device, dtype=torch.float32)
    check(_lib.lib().ha_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), B, Cc, H, W, _stream_ptr()), "ha_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, "nhwc_to_nchw input")
    x = x.contiguous().float()
    B, H, W, Cc = x.shape
    out = torch.empty(B, Cc, H, W, device=x.device, dtype=torch.float32)
    check(_lib.lib().ha_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), B, Cc, H, W, _stream_ptr()), "ha_nhwc_to_nchw")
    return out


# ------------------------------------------------------------------------------- LM parameters
@dataclass
class LmSetup:
    """Everything about one LM run that does not depend on the batch content."""
    kind: str                  # 'kitti' | 'ford' | 'g2sp' | 'g2sp_nn' (LM_G2SP --proj nn: in-plane warp, models_kitti.py:289-331)
    n_iters: int
    level_first: int
    dof: int
    using_weight: int
    use_hessian: int
    rotation_range: float
    shift_range_lat: float
    shift_range_lon: float
    kernel_variant: int = 0    # HaLmParams.kernel_variant: 0 = default kernels (lm_run chains its step launches), 1 = register-staged
                               # validation kernel, 2 = default kernel with plain stream-ordered launches
    optimizer: str = "LM"      # args.Optimizer: 'LM' | 'SGD' | 'ADAM' (LM_S2GP) | 'GN' (LM_S2GP_Ford)
    full_height: int = 0       # 1: residual over the whole ground image (args.proj != 'geo', models_kitti.py:1200-1205)
    adam_level_mult: int = 0   # args.level: the reference's Adam step count is iter * args.level + level (:1241)
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999


def dof_of(args, kind: str) -> int:
    """models_kitti.py:954-957; the Ford model always refines all three (models_ford.py:380)."""
    if kind in ("ford", "g2sp", "g2sp_nn"):    # LM_G2SP.LM_update always refines all three (models_kitti.py:375-377)
        return 3
    if args.rotation_range == 0:
        return 2
    if args.shift_range_lat == 0 and args.shift_range_lon == 0:
        return 1
    return 3


def setup_from_args(args, kind: str, level_first: int = 0) -> LmSetup:
    opt = getattr(args, "Optimizer", "LM")
    first_order = opt in ("SGD", "ADAM", "NN")           # SGD_update / ADAM_update / NN_update ignore the confidence weights and the DOF switch
    return LmSetup(kind=kind, n_iters=int(args.N_iters), level_first=int(level_first), dof=3 if first_order else dof_of(args, kind),
                   using_weight=0 if first_order else int(bool(args.using_weight)),
                   use_hessian=int(bool(getattr(args, "use_hessian", 0))),
                   rotation_range=float(args.rotation_range), shift_range_lat=float(args.shift_range_lat),
                   shift_range_lon=float(args.shift_range_lon), optimizer=opt,
                   full_height=int(not kind.startswith("g2sp") and getattr(args, "proj", "geo") != "geo"),
                   adam_level_mult=int(args.level), adam_beta1=float(getattr(args, "beta1", 0.9)),
                   adam_beta2=float(getattr(args, "beta2", 0.999)))


def resolve_damping(args, damping_param: Optional[torch.Tensor], dof: int, kind: str = "kitti") -> List[float]:
    """models_kitti.py:958-966: trained lambda = 10^(-6 + 11*sigmoid(p)) else args.damping.
    LM_G2SP uses the trained parameter as is (models_kitti.py:356-359)."""
    if kind.startswith("g2sp"):
        if getattr(args, "train_damping", 0):
            return [float(v) for v in damping_param.detach().float().reshape(-1).tolist()]
        return [float(np.float32(args.damping))] * 3
    if getattr(args, "train_damping", 0):
        lam = (10.0 ** (-6 + damping_param.detach().float().sigmoid() * 11)).reshape(-1).tolist()
        if len(lam) == 1:
            lam = lam * dof
        if len(lam) != dof:
            raise _lib.HaError("damping parameter has %d entries for %d degrees of freedom" % (len(lam), dof))
        return [float(v) for v in lam]
    return [float(np.float32(args.damping))] * dof


def execution_order(n_iters: int, n_levels: int, level_first: int) -> List[Tuple[int, int]]:
    """(iteration, level) in the order the reference executes them (models_kitti.py:1176-1180 / :1349-1353)."""
    if level_first:
        return [(it, lv) for lv in range(n_levels) for it in range(n_iters)]
    return [(it, lv) for it in range(n_iters) for lv in range(n_levels)]


def draw_reset_uv(n_steps: int, B: int) -> torch.Tensor:
    """The reference draws two [B,1] Uniform(-1,1) samples from the CPU default generator on every
    3-DOF LM step whether or not they are used (models_kitti.py:1028-1029).  Make the same draws in
    the same order so the global RNG stream (and any reset sample) stays identical."""
    dist = torch.distributions.uniform.Uniform(-1, 1)
    out = torch.empty(n_steps, 2, B, dtype=torch.float32)
    for k in range(n_steps):
        out[k, 0] = dist.sample([B, 1])[:, 0]
        out[k, 1] = dist.sample([B, 1])[:, 0]
    return out


def _levels(p: Pyramid, n: int):
    arr = (HaLevel * n)()
    for i in range(n):
        f = p.feats[i]
        if f.dtype != torch.float32 or not f.is_contiguous():
            raise _lib.HaError("pyramid level %d must be contiguous fp32 NHWC" % i)
        arr[i].data = f.data_ptr()
        arr[i].scale = p.scales[i].data_ptr() if p.scales[i] is not None else None
        arr[i].H, arr[i].W, arr[i].C = f.shape[1], f.shape[2], f.shape[3]
    return arr


def make_params(setup: LmSetup, sat: Pyramid, damping: Sequence[float], side_m: Optional[float],
                ori_grd_hw: Tuple[int, int] = (256, 1024)) -> HaLmParams:
    n = len(sat.feats)
    p = HaLmParams()
    p.geometry = {"kitti": _lib.HA_GEOM_KITTI, "ford": _lib.HA_GEOM_FORD, "g2sp": _lib.HA_GEOM_G2SP,
                  "g2sp_nn": _lib.HA_GEOM_G2SP_NN}[setup.kind]
    p.ori_grd_h, p.ori_grd_w = int(ori_grd_hw[0]), int(ori_grd_hw[1])
    p.kernel_variant, p.reserved = int(setup.kernel_variant), 0
    p.optimizer = OPTIMIZERS[setup.optimizer]
    p.full_height, p.adam_level_mult, p.adam_iter = int(setup.full_height), int(setup.adam_level_mult), 0
    p.adam_beta1, p.adam_beta2 = setup.adam_beta1, setup.adam_beta2
    p.n_levels, p.n_iters, p.level_first, p.dof = n, setup.n_iters, setup.level_first, setup.dof
    p.using_weight, p.use_hessian, p.batch = setup.using_weight, setup.use_hessian, sat.batch
    p.rotation_range, p.shift_range_lat, p.shift_range_lon = setup.rotation_range, setup.shift_range_lat, setup.shift_range_lon
    for i, v in enumerate(damping):
        p.damping[i] = v
    for lv in range(n):
        A = sat.feats[lv].shape[1]
        if setup.kind == "kitti":
            mpp = kitti_meter_per_pixel() * (SAT_PROCESS_SIDE / A)      # models_kitti.py:757-758
            center = A / 2                                              # :765
        elif setup.kind == "g2sp":
            mpp = kitti_meter_per_pixel() * (SAT_PROCESS_SIDE / A)      # models_kitti.py:69-70
            center = A // 2                                             # :65
        elif setup.kind == "g2sp_nn":
            mpp = kitti_meter_per_pixel() * (SAT_PROCESS_SIDE / A)      # models_kitti.py:291-292
            center = A / 2                                              # :307,312
        else:
            mpp = side_m / A                                            # models_ford.py:230
            center = A // 2                                             # :231
        p.meter_per_pixel[lv] = mpp
        p.inv_meter_per_pixel[lv] = 1.0 / mpp
        p.sat_center[lv] = center
    return p


OPTIMIZERS = {"LM": _lib.HA_OPT_LM, "SGD": _lib.HA_OPT_SGD, "ADAM": _lib.HA_OPT_ADAM, "GN": _lib.HA_OPT_GN,
              "NN": _lib.HA_OPT_LM}      # 'NN' never reaches the step kernel (engine.nn_run): the field is unused there


def draws_reset(setup: LmSetup) -> bool:
    """Whether a step consumes the two [B,1] CPU-RNG draws: the 3-DOF LM / GN updates of the S2GP models do
    (models_kitti.py:1028-1029, models_ford.py:453-454, :583-584); SGD / ADAM and LM_G2SP never draw."""
    return setup.dof == 3 and not setup.kind.startswith("g2sp") and setup.optimizer in ("LM", "GN")


class LmWorkspace:
    """Scratch of the LM loop (partials, tickets, zero vector, |g|^2 cache), reused across calls."""

    def __init__(self):
        self.buf = None

    def get(self, B: int, device) -> torch.Tensor:
        need = _lib.lib().ha_lm_workspace_bytes(B)
        if self.buf is None or self.buf.numel() < need or self.buf.device != device:
            self.buf = torch.empty(need, dtype=torch.uint8, device=device)
        return self.buf


_WS = {}


def _workspace(device) -> LmWorkspace:
    """One workspace per (device, stream): the C ABI is re-entrant per (device, stream) with caller-owned scratch, so two
    models (or threads) running on different streams of one device must not share tickets and partials."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    if key not in _WS:
        if len(_WS) > 64:                       # short-lived streams: do not grow without bound
            _WS.clear()
        _WS[key] = LmWorkspace()
    return _WS[key]


def _new_status(device) -> torch.Tensor:
    """A status word of its own for every call (ha_lm_run clears it on the stream before OR-ing bits into it), so that
    results of different calls never alias each other's flags."""
    return torch.empty(1, dtype=torch.int32, device=device)


class LmStatusError(AssertionError):
    """The reference's `assert torch.sum(mask) > 0` (jacobian.py:172) as seen through the device status word."""


def check_status(status: torch.Tensor, where: str = "LM loop") -> int:
    """Reads a status word (ONE device->host sync; the reference syncs on every step: jacobian.py:172,200,
    models_kitti.py:1037) and applies the reference's error convention: AssertionError when a step found no satellite
    sample point in range anywhere in the batch, a printed note for NaN poses.  Returns the bits."""
    bits = int(status.item())
    if bits & _lib.HA_STATUS_NO_INRANGE:
        raise LmStatusError("%s: no sample point of the batch lands inside the satellite map (jacobian.py:172 asserts "
                            "torch.sum(mask) > 0)" % where)
    if bits & _lib.HA_STATUS_TIMEOUT:
        # the chained step launches of ha_lm_run gave up waiting for a sample's previous step: the poses are invalid
        raise RuntimeError("%s: a step of the chained LM loop timed out waiting for its predecessor (HA_STATUS_TIMEOUT)" % where)
    if bits & _lib.HA_STATUS_NAN_POSE:
        print('theta_new is nan')               # models_kitti.py:1037-1039 prints and carries on
    return bits


@dataclass
class LmResult:
    traj: torch.Tensor                    # [B, n_iters, n_levels, 3] (su, sv, theta) after every step
    pose: torch.Tensor                    # [B, 3] final
    stats: Optional[torch.Tensor]         # [n_iters, n_levels, B, HA_STATS] or None
    status: torch.Tensor                  # device int32[1], OR of HA_STATUS_* bits (not synced here)


def lm_run(setup: LmSetup, sat: Pyramid, grd: Pyramid, tables: Sequence[torch.Tensor], damping: Sequence[float],
           extrinsics: Optional[torch.Tensor] = None, side_m: Optional[float] = None,
           pose0: Optional[torch.Tensor] = None, reset_uv: Optional[torch.Tensor] = None,
           want_stats: bool = False, ori_grd_hw: Tuple[int, int] = (256, 1024)) -> LmResult:
    """The whole LM loop on the current CUDA stream, no host synchronisation.
    `extrinsics`: Ford [B,12] (R_FL | T_FL); G2SP [B,9] (left_camera_k); None for KITTI S2GP."""
    L = _lib.lib()
    n = len(sat.feats)
    B = sat.batch
    dev = sat.feats[0].device
    _require_cuda(sat.feats[0], "satellite features")
    pose = torch.zeros(B, 3, dtype=torch.float32, device=dev) if pose0 is None else pose0.to(dev, torch.float32).clone()
    traj = torch.empty(B, setup.n_iters, n, 3, dtype=torch.float32, device=dev)
    stats = torch.empty(setup.n_iters, n, B, _lib.HA_STATS, dtype=torch.float32, device=dev) if want_stats else None
    n_steps = setup.n_iters * n
    if draws_reset(setup):                            # LM_G2SP / SGD / ADAM have no out-of-range reset, hence no RNG draws
        if reset_uv is None:
            reset_uv = draw_reset_uv(n_steps, B)
        reset_uv = reset_uv.to(dev, torch.float32, non_blocking=True).contiguous()
        assert reset_uv.shape == (n_steps, 2, B)
    else:
        reset_uv = None
    if setup.kind == "ford":
        if extrinsics is None or side_m is None:
            raise _lib.HaError("Ford geometry needs R_FL/T_FL and satmap_sidelength_meters")
    if setup.kind == "g2sp" and extrinsics is None:
        raise _lib.HaError("G2SP geometry needs left_camera_k")
    if extrinsics is not None:
        extrinsics = extrinsics.to(dev, torch.float32).contiguous()
    params = make_params(setup, sat, damping, side_m, ori_grd_hw)
    confs = (C.c_void_p * n)()
    tabs = (C.c_void_p * n)()
    for i in range(n):
        c = grd.confs[i] if grd.confs else None
        if setup.using_weight and c is None:
            raise _lib.HaError("using_weight needs ground confidence maps")
        confs[i] = c.data_ptr() if (c is not None and setup.using_weight) else None
        if setup.kind.startswith("g2sp"):
            tabs[i] = None                            # the satellite-plane points are computed in the kernel
        else:
            assert tables[i].is_cuda and tables[i].shape[:2] == grd.feats[i].shape[1:3], "ground table / feature shape mismatch"
            tabs[i] = tables[i].data_ptr()
    ws, status = _workspace(dev).get(B, dev), _new_status(dev)
    rc = L.ha_lm_run(C.byref(params), _levels(sat, n), _levels(grd, n), confs, tabs,
                     extrinsics.data_ptr() if extrinsics is not None else None, pose.data_ptr(),
                     reset_uv.data_ptr() if reset_uv is not None else None, traj.data_ptr(),
                     stats.data_ptr() if stats is not None else None, status.data_ptr(), ws.data_ptr(), ws.numel(),
                     _stream_ptr())
    check(rc, "ha_lm_run")
    return LmResult(traj, pose, stats, status)


def lm_step(setup: LmSetup, level: int, sat: Pyramid, grd: Pyramid, tables: Sequence[torch.Tensor],
            damping: Sequence[float], pose: torch.Tensor, extrinsics: Optional[torch.Tensor] = None,
            side_m: Optional[float] = None, reset_uv: Optional[torch.Tensor] = None,
            ori_grd_hw: Tuple[int, int] = (256, 1024), adam_iter: int = 0):
    """One fused LM step at `level` from `pose` [B,3]; returns (new_pose [B,3], stats [B,HA_STATS]).
    `adam_iter`: iteration index of the step (Optimizer 'ADAM' only: its moments persist in the stream's workspace)."""
    L = _lib.lib()
    n = len(sat.feats)
    B = sat.batch
    dev = sat.feats[0].device
    pose = pose.to(dev, torch.float32).clone().contiguous()
    stats = torch.empty(B, _lib.HA_STATS, dtype=torch.float32, device=dev)
    if draws_reset(setup):
        if reset_uv is None:
            reset_uv = draw_reset_uv(1, B)[0]
        reset_uv = reset_uv.to(dev, torch.float32).contiguous()
    else:
        reset_uv = None
    if extrinsics is not None:
        extrinsics = extrinsics.to(dev, torch.float32).contiguous()
    params = make_params(setup, sat, damping, side_m, ori_grd_hw)
    params.adam_iter = int(adam_iter)
    c = grd.confs[level] if (grd.confs and setup.using_weight) else None
    ws, status = _workspace(dev).get(B, dev), _new_status(dev)
    sl, gl = _levels(sat, n), _levels(grd, n)
    rc = L.ha_lm_step(C.byref(params), level, C.byref(sl[level]), C.byref(gl[level]),
                      c.data_ptr() if c is not None else None,
                      tables[level].data_ptr() if not setup.kind.startswith("g2sp") else None,
                      extrinsics.data_ptr() if extrinsics is not None else None, pose.data_ptr(),
                      reset_uv.data_ptr() if reset_uv is not None else None, stats.data_ptr(), status.data_ptr(),
                      ws.data_ptr(), ws.numel(), _stream_ptr())
    check(rc, "ha_lm_step")
    lm_step.last_status = status
    return pose, stats


class FusedLmLoop(torch.autograd.Function):
    """The whole LM loop with a native backward (first slice of SURVEY.md section 8 f-1): forward = ha_lm_run with the
    per-step diagnostics kept, backward = ha_lm_step_backward over the steps in reverse execution order — the adjoint of
    project_map_to_grd -> jacobian.grid_sample -> LM_update that `loss.backward()` computes in the reference
    (train_kitti.py:365 through models_kitti.py:1176-1260).  Differentiable inputs: the damping columns `lam` [3] and the
    L2-normalised NHWC feature pyramids; output: the pose trajectory [B, N_iters, L, 3].  Scope of the slice: S2GP
    geometries, 3 degrees of freedom, unweighted residuals (`supports()`); everything else keeps the torch path."""

    @staticmethod
    def supports(setup: LmSetup) -> bool:
        return setup.kind in ("kitti", "ford") and setup.dof == 3 and not setup.using_weight and \
            setup.optimizer == "LM" and not setup.full_height

    @staticmethod
    def forward(ctx, setup, tables, extrinsics, side_m, reset_uv, lam, n_levels, *feats):
        sat = Pyramid([f.detach().contiguous() for f in feats[:n_levels]], [None] * n_levels)
        grd = Pyramid([f.detach().contiguous() for f in feats[n_levels:]], [None] * n_levels)
        lam_list = [float(v) for v in lam.detach().reshape(-1).tolist()]
        res = lm_run(setup, sat, grd, tables, lam_list, extrinsics=extrinsics, side_m=side_m, reset_uv=reset_uv, want_stats=True)
        ctx.setup, ctx.tables, ctx.extrinsics, ctx.side_m, ctx.lam_list, ctx.n = setup, tables, extrinsics, side_m, lam_list, n_levels
        ctx.sat, ctx.grd, ctx.res = sat, grd, res
        ctx.lam_shape = lam.shape
        return res.traj

    @staticmethod
    def backward(ctx, gtraj):
        L = _lib.lib()
        setup, sat, grd, res, n = ctx.setup, ctx.sat, ctx.grd, ctx.res, ctx.n
        B = sat.batch
        dev = sat.feats[0].device
        gtraj = gtraj.contiguous().float()
        gsat = [torch.zeros_like(f) for f in sat.feats]
        ggrd = [torch.zeros_like(f) for f in grd.feats]
        glam = torch.zeros(B, 3, dtype=torch.float32, device=dev)
        params = make_params(setup, sat, ctx.lam_list, ctx.side_m)
        sl, gl = _levels(sat, n), _levels(grd, n)
        need = L.ha_lm_backward_workspace_bytes(B)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        ext = ctx.extrinsics.to(dev, torch.float32).contiguous() if ctx.extrinsics is not None else None
        order = execution_order(setup.n_iters, n, setup.level_first)
        carry = torch.zeros(B, 3, dtype=torch.float32, device=dev)
        zero_pose = torch.zeros(B, 3, dtype=torch.float32, device=dev)
        st = _stream_ptr()
        for k in range(len(order) - 1, -1, -1):
            it, lv = order[k]
            pose_in = zero_pose if k == 0 else res.traj[:, order[k - 1][0], order[k - 1][1]].contiguous()
            gout = (gtraj[:, it, lv] + carry).contiguous()
            gin = torch.empty(B, 3, dtype=torch.float32, device=dev)
            check(L.ha_lm_step_backward(C.byref(params), lv, C.byref(sl[lv]), C.byref(gl[lv]), ctx.tables[lv].data_ptr(),
                                        ext.data_ptr() if ext is not None else None, pose_in.data_ptr(),
                                        res.stats[it, lv].data_ptr(), gout.data_ptr(), gin.data_ptr(), gsat[lv].data_ptr(),
                                        ggrd[lv].data_ptr(), glam.data_ptr(), ws.data_ptr(), need, st), "ha_lm_step_backward")
            carry = gin
        return (None, None, None, None, None, glam.sum(dim=0).reshape(ctx.lam_shape), None, *gsat, *ggrd)


class PoseLoss(torch.autograd.Function):
    """loss_func with loss_method 0 (models_ford.py:1041-1093) as one kernel forward (ha_pose_loss) and one backward
    (ha_pose_loss_backward).  traj [B, N_iters, L, 3] and gt [B, 3] share a component order; coe = the three loss
    coefficients in that order.  Returns (loss scalar, err [N_iters, L, 3] = mean_b |traj - gt|)."""

    @staticmethod
    def forward(ctx, traj, gt, coe):
        _require_cuda(traj, "trajectory")
        traj = traj.contiguous().float()
        gt = gt.to(traj.device, torch.float32).contiguous()
        B, N, Lv, _ = traj.shape
        err = torch.empty(N, Lv, 3, dtype=torch.float32, device=traj.device)
        loss = torch.empty((), dtype=torch.float32, device=traj.device)
        c = (C.c_float * 3)(*[float(v) for v in coe])
        check(_lib.lib().ha_pose_loss(traj.data_ptr(), gt.data_ptr(), B, N, Lv, c, err.data_ptr(), loss.data_ptr(), _stream_ptr()),
              "ha_pose_loss")
        ctx.save_for_backward(traj, gt)
        ctx.coe = [float(v) for v in coe]
        return loss, err

    @staticmethod
    def backward(ctx, gloss, gerr):
        traj, gt = ctx.saved_tensors
        B, N, Lv, _ = traj.shape
        gtraj = torch.empty_like(traj)
        c = (C.c_float * 3)(*ctx.coe)
        gl = gloss.contiguous().float() if gloss is not None else None
        ge = gerr.contiguous().float() if gerr is not None else None
        check(_lib.lib().ha_pose_loss_backward(traj.data_ptr(), gt.data_ptr(), B, N, Lv, c, ge.data_ptr() if ge is not None else None,
                                               gl.data_ptr() if gl is not None else None, gtraj.data_ptr(), _stream_ptr()),
              "ha_pose_loss_backward")
        return gtraj, None, None


def ford_extrinsics(R_FL: torch.Tensor, T_FL: torch.Tensor) -> torch.Tensor:
    """[B,3,3] + [B,3] -> [B,12] in the layout ha_lm_* expects."""
    B = R_FL.shape[0]
    return torch.cat([R_FL.reshape(B, 9), T_FL.reshape(B, 3)], dim=1).float().contiguous()


# ------------------------------------------------------------------------------- VGG
VGG_CONV_NAMES = ["conv0", "conv2", "conv5", "conv7", "conv10", "conv12", "conv14", "conv_dec1.1", "conv_dec1.3",
                  "conv_dec2.1", "conv_dec2.3", "conv_dec3.1", "conv_dec3.3", "conf0.1", "conf1.1", "conf2.1", "conf3.1"]
PRECISIONS = {"fp32": _lib.HA_CONV_FP32_SIMT, "f16x3": _lib.HA_CONV_F16X3, "f16": _lib.HA_CONV_F16,
              "f16x3_1cta": _lib.HA_CONV_F16X3_1CTA}


class VggRunner:
    """Packs one U-Net's weights for the kernels (re-packed when a parameter changes) and runs the
    feature extractor through ha_vgg_forward, chunking the batch to bound the workspace."""

    def __init__(self, max_ws_bytes: int = 24 << 30):
        _lib.lib()                      # fail loudly at construction if libha_b200.so is missing: nothing here degrades to torch
        self.packed = None
        self.key = None
        self.ws = None
        self.max_ws_bytes = max_ws_bytes

    def _pack(self, named: dict, device) -> None:
        # `named` is the module's own (cached) name -> Parameter dict: storage address + version counter of every
        # parameter tell whether a load_state_dict / optimiser step / .to() happened since the weights were packed
        key = tuple((p.data_ptr(), p._version) for p in named.values())
        if self.packed is not None and key == self.key and self.packed.device == device:
            return
        L = _lib.lib()
        sd = HaVggStateDict()
        keep = []
        for i, n in enumerate(VGG_CONV_NAMES):
            w = named[n + ".weight"].detach()
            _require_cuda(w, n + ".weight")
            w = w.float().contiguous()
            keep.append(w)
            sd.weight[i] = w.data_ptr()
            b = named.get(n + ".bias")
            if b is not None:
                b = b.detach().float().contiguous()
                keep.append(b)
                sd.bias[i] = b.data_ptr()
        nbytes = L.ha_vgg_packed_weight_bytes()
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
        check(L.ha_vgg_pack_weights(C.byref(sd), self.packed.data_ptr(), nbytes, _stream_ptr()), "ha_vgg_pack_weights")
        self.key = key
        self._keep = keep

    def __call__(self, named: dict, img: torch.Tensor, n_levels: int, want_conf: bool, precision: str,
                 want_scale: bool = True, g2s: bool = False) -> Pyramid:
        """`want_scale=False` skips the L2-norm pass (VGG.py:172-175): the S2GP LM step renormalises the sampled and the
        ground vectors itself (models_kitti.py:982-989), so the per-sample scale cancels there exactly."""
        _require_cuda(img, "image")
        L = _lib.lib()
        dev = img.device
        img = img.float().contiguous()
        B, c3, H, W = img.shape
        if c3 != 3:
            raise _lib.HaError("image must be [B,3,H,W]")
        prec = PRECISIONS[precision]
        self._pack(named, dev)
        per1 = L.ha_vgg_workspace_bytes(1, H, W, n_levels, prec)
        if per1 == 0:
            raise _lib.HaError("unsupported VGG shape B=%d H=%d W=%d levels=%d" % (B, H, W, n_levels))
        chunk = max(1, min(B, int(self.max_ws_bytes // per1)))
        need = L.ha_vgg_workspace_bytes(chunk, H, W, n_levels, prec)
        if self.ws is None or self.ws.numel() < need or self.ws.device != dev:
            self.ws = None
            self.ws = torch.empty(need, dtype=torch.uint8, device=dev)
        feats, scales, confs = [], [], []
        for l in range(n_levels):
            h, w, ch = H >> (3 - l), W >> (3 - l), PYRAMID_CHANNELS[l]
            feats.append(torch.empty(B, h, w, ch, dtype=torch.float32, device=dev))
            scales.append(torch.empty(B, dtype=torch.float32, device=dev) if want_scale else None)
            confs.append(torch.empty(B, h, w, dtype=torch.float32, device=dev) if want_conf else None)
        st = _stream_ptr()
        for b0 in range(0, B, chunk):
            nb = min(chunk, B - b0)
            pf, ps, pc = (C.c_void_p * n_levels)(), (C.c_void_p * n_levels)(), (C.c_void_p * n_levels)()
            for l in range(n_levels):
                pf[l] = feats[l][b0:].data_ptr()
                ps[l] = scales[l][b0:].data_ptr() if want_scale else None
                pc[l] = confs[l][b0:].data_ptr() if want_conf else None
            fn, what = (L.ha_vgg_g2s_forward, "ha_vgg_g2s_forward") if g2s else (L.ha_vgg_forward, "ha_vgg_forward")
            check(fn(self.packed.data_ptr(), img[b0:].data_ptr(), nb, H, W, n_levels, prec, pf, ps, pc,
                     self.ws.data_ptr(), self.ws.numel(), st), what)
        if g2s:      # VGG.py:283-299: the features are the [h, w] maps read as [2h, w/2]; c0 stays un-folded (:326)
            feats = [f.view(B, 2 * f.shape[1], f.shape[2] // 2, f.shape[3]) for f in feats]
            confs = [c if (c is None or l == 0) else c.view(B, 2 * c.shape[1], c.shape[2] // 2) for l, c in enumerate(confs)]
        return Pyramid(feats, scales, confs)


N_FEATURE_CONVS = 13            # conv0 ... conv_dec3.3: the convolutions the pyramid features depend on
N_BIASED_CONVS = 7              # the VGG16 encoder convolutions carry a bias (VGG.py:23-29)


class VggTrain(torch.autograd.Function):
    """The U-Net with a native backward (second slice): forward = ha_vgg_forward_train (the tcgen05
    schedule of the eval path, keeping the activations in its workspace), backward = ha_vgg_backward (data gradients on the
    same tcgen05 convolution kernels with flipped weights, weight gradients on a tcgen05 split-K GEMM over the pixels).
    apply(runner, named, n_levels, img, *params) with params = the 13 feature-conv weights then the 7 encoder biases
    (autograd inputs); returns the raw NHWC feature maps x15 / x18 / x21 (/ x24) (VGG.py:141,147,152,157) followed by the
    confidence maps (not differentiated); the caller L2-normalises and slices."""

    @staticmethod
    def supports(img: torch.Tensor, n_levels: int, precision: str) -> bool:
        return img.is_cuda and n_levels in (3, 4) and precision == "f16x3" and img.shape[-1] % 64 == 0 and img.shape[-2] % 32 == 0

    @staticmethod
    def forward(ctx, runner, named, n_levels, img, *params):
        L = _lib.lib()
        dev = img.device
        img = img.float().contiguous()
        B, _, H, W = img.shape
        n = int(n_levels)
        runner._pack(named, dev)
        need = L.ha_vgg_train_workspace_bytes(B, H, W, n)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)          # owned by this call's graph until backward
        feats = [torch.empty(B, H >> (3 - l), W >> (3 - l), PYRAMID_CHANNELS[l], dtype=torch.float32, device=dev) for l in range(n)]
        confs = [torch.empty(B, H >> (3 - l), W >> (3 - l), dtype=torch.float32, device=dev) for l in range(n)]
        pf, pc = (C.c_void_p * n)(*[f.data_ptr() for f in feats]), (C.c_void_p * n)(*[c.data_ptr() for c in confs])
        check(L.ha_vgg_forward_train(runner.packed.data_ptr(), img.data_ptr(), B, H, W, n, pf, None, pc, ws.data_ptr(), need,
                                     _stream_ptr()), "ha_vgg_forward_train")
        ctx.ws, ctx.img, ctx.named, ctx.shape, ctx.n = ws, img, named, (B, H, W), n
        ctx.mark_non_differentiable(*confs)
        return (*feats, *confs)

    @staticmethod
    def backward(ctx, *grads_out):
        L = _lib.lib()
        B, H, W = ctx.shape
        n = ctx.n
        dev = ctx.img.device
        named = ctx.named
        gs = []
        for l, g in enumerate(grads_out[:n]):
            shape = (B, H >> (3 - l), W >> (3 - l), PYRAMID_CHANNELS[l])
            gs.append(torch.zeros(shape, dtype=torch.float32, device=dev) if g is None else g.float().contiguous())
        sd, gd = HaVggStateDict(), HaVggGrads()
        keep, gw, gb = [], [], []
        n_convs = 11 if n == 3 else N_FEATURE_CONVS                  # conv_dec3 only exists in the 4-level pyramid
        for i, name in enumerate(VGG_CONV_NAMES):
            wt = named[name + ".weight"].detach().float().contiguous()
            keep.append(wt)
            sd.weight[i] = wt.data_ptr()
            if i < n_convs:
                g = torch.empty_like(wt)
                gw.append(g)
                gd.weight[i] = g.data_ptr()
            elif i < N_FEATURE_CONVS:
                gw.append(None)
            if i < N_BIASED_CONVS:
                g = torch.empty(wt.shape[0], dtype=torch.float32, device=dev)
                gb.append(g)
                gd.bias[i] = g.data_ptr()
        need = L.ha_vgg_backward_workspace_bytes(B, H, W, n)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        pg = (C.c_void_p * n)(*[g.data_ptr() for g in gs])
        check(L.ha_vgg_backward(C.byref(sd), ctx.img.data_ptr(), B, H, W, n, ctx.ws.data_ptr(), pg, C.byref(gd), ws.data_ptr(), need,
                                _stream_ptr()), "ha_vgg_backward")
        ctx.ws = None
        return (None, None, None, None, *gw, *gb)


def conv3x3(x_nhwc: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], precision: str) -> torch.Tensor:
    """One 3x3/pad-1 convolution through ha_conv3x3_nhwc: fp32 NHWC in -> fp32 NHWC out (bias, no activation)."""
    _require_cuda(x_nhwc, "conv3x3 input")
    L = _lib.lib()
    x = x_nhwc.float().contiguous()
    w = weight.detach().float().contiguous().to(x.device)
    b = None if bias is None else bias.detach().float().contiguous().to(x.device)
    B, H, W, cin = x.shape
    cout = w.shape[0]
    out = torch.empty(B, H, W, cout, dtype=torch.float32, device=x.device)
    need = L.ha_conv3x3_workspace_bytes(cin, cout, B, H, W)
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    check(L.ha_conv3x3_nhwc(x.data_ptr(), cin, w.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(), cout,
                            B, H, W, PRECISIONS[precision], ws.data_ptr(), need, _stream_ptr()), "ha_conv3x3_nhwc")
    return out


def conv3x3_backward(x_nhwc: torch.Tensor, weight: torch.Tensor, dy_nhwc: torch.Tensor, want_dx: bool = True):
    """Backward of one 3x3 / pad-1 convolution through ha_conv3x3_backward_nhwc (tcgen05 data and weight gradients):
    returns (dx NHWC or None, dw OIHW, db)."""
    _require_cuda(x_nhwc, "conv3x3_backward input")
    L = _lib.lib()
    x = x_nhwc.float().contiguous()
    dy = dy_nhwc.float().contiguous()
    w = weight.detach().float().contiguous().to(x.device)
    B, H, W, cin = x.shape
    cout = w.shape[0]
    dx = torch.empty_like(x) if want_dx else None
    dw = torch.empty_like(w)
    db = torch.empty(cout, dtype=torch.float32, device=x.device)
    need = L.ha_conv3x3_backward_workspace_bytes(cin, cout, B, H, W)
    if need == 0:
        raise _lib.HaError("unsupported conv3x3_backward shape")
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    check(L.ha_conv3x3_backward_nhwc(x.data_ptr(), cin, w.data_ptr(), dy.data_ptr(), cout, B, H, W, dx.data_ptr() if want_dx else None,
                                     dw.data_ptr(), db.data_ptr(), ws.data_ptr(), need, _stream_ptr()), "ha_conv3x3_backward_nhwc")
    return dx, dw, db


def nn_run(setup: LmSetup, sat: Pyramid, grd: Pyramid, tables: Sequence[torch.Tensor], nn_params: dict, precision: str = "f16x3",
           extrinsics: Optional[torch.Tensor] = None, side_m: Optional[float] = None, pose0: Optional[torch.Tensor] = None) -> LmResult:
    """The refinement loop with Optimizer 'NN' (models_kitti.py:1233-1239, NN_update :1043-1054, RNNs.NNrefine): per step
    ha_lm_residual (materialised relu(sat_proj - grd)) -> ha_conv3x3_nhwc (linear_k: Conv2d(C, 64) + bias on tcgen05) ->
    ha_nn_pose_update (mean, 64-16-3 mapping, tanh, pose += delta).  `nn_params`: the NNrefine state dict
    ('linear{k}.1.weight/bias' for C = 256 / 128 / 64 / 16, 'mapping.1.*', 'mapping.3.*') as CUDA tensors."""
    L = _lib.lib()
    n = len(sat.feats)
    B = sat.batch
    dev = sat.feats[0].device
    _require_cuda(sat.feats[0], "satellite features")
    pose = torch.zeros(B, 3, dtype=torch.float32, device=dev) if pose0 is None else pose0.to(dev, torch.float32).clone()
    traj = torch.empty(B, setup.n_iters, n, 3, dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    params = make_params(setup, sat, [0.0] * 3, side_m)
    sl, gl = _levels(sat, n), _levels(grd, n)
    ext = extrinsics.to(dev, torch.float32).contiguous() if extrinsics is not None else None
    head = {256: "linear0", 128: "linear1", 64: "linear2", 16: "linear3"}            # RNNs.py:118-125: chosen by channel count
    w0, b0 = nn_params["mapping.1.weight"].float().contiguous(), nn_params["mapping.1.bias"].float().contiguous()
    w1, b1 = nn_params["mapping.3.weight"].float().contiguous(), nn_params["mapping.3.bias"].float().contiguous()
    st = _stream_ptr()
    for it, lv in execution_order(setup.n_iters, n, setup.level_first):
        g = grd.feats[lv]
        rows = g.shape[1] if setup.full_height else g.shape[1] - g.shape[1] // 2
        res = torch.empty(B, rows, g.shape[2], g.shape[3], dtype=torch.float32, device=dev)
        check(L.ha_lm_residual(C.byref(params), lv, C.byref(sl[lv]), C.byref(gl[lv]), tables[lv].data_ptr(),
                               ext.data_ptr() if ext is not None else None, pose.data_ptr(), 1, res.data_ptr(), st), "ha_lm_residual")
        k = head[g.shape[3]]
        x = conv3x3(res, nn_params[k + ".1.weight"], nn_params[k + ".1.bias"], precision)
        step = traj[:, it, lv]
        check(L.ha_nn_pose_update(x.data_ptr(), B, rows * g.shape[2], w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr(),
                                  pose.data_ptr(), step.data_ptr(), setup.n_iters * n * 3, status.data_ptr(), st), "ha_nn_pose_update")
    return LmResult(traj, pose, None, status)
