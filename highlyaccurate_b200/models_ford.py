"""Drop-in for the reference's models_ford.py on the accelerated path: `LM_S2GP_Ford` and
`loss_func`.  Citations are into the upstream models_ford.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import compat, engine
from ._lib import HaError
from .VGG import VGGUnet


class _TrajectoryOutputs(torch.autograd.Function):
    """Attaches the engine's pose trajectories to the autograd graph.

    The reference's eval loops call loss.backward() on the outputs "just to release graph"
    (train_kitti.py:60-64, train_ford.py:78-79), so test-mode outputs must require grad and accept a
    backward; it contributes nothing.  The fused engine has no backward yet:
    `forward(mode='train')` of the S2GP models therefore takes the differentiable torch path
    (`train_forward` below) and never reaches this function with train=True; anything else that would
    fails loudly instead of silently producing zero gradients."""

    @staticmethod
    def forward(ctx, anchor, train, *trajs):
        ctx.train = bool(train)
        return tuple(t.clone() for t in trajs)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.train:
            raise HaError("backward through the fused LM loop is not implemented yet (SURVEY.md section 8 f-1); "
                          "train with the reference model and evaluate with this engine")
        return (None, None) + tuple(None for _ in grads)


def train_forward_fused(net, kind, sat_map, grd_img, gt_lat, gt_lon, gt_theta, level_first, coe_theta, ford=None):
    """`forward(mode='train')` with the native LM backward (engine.FusedLmLoop): the U-Nets run through torch autograd
    (`VGGUnet.forward_autograd`), the whole LM loop — forward and backward — runs in libha_b200.so, the loss is
    `loss_func` method 0.  Gradients reach both U-Nets and `damping` exactly as in the reference
    (tests/test_gpu_parity.py::test_fused_lm_backward_* compare with the reference's autograd, KAT-8 / KAT-9)."""
    a = net.args
    sat_feats, _ = net.SatFeatureNet.forward_train(sat_map)            # NHWC, L2-normalised; native U-Net backward where it applies
    grd_feats, grd_confs = net.GrdFeatureNet.forward_train(grd_img)
    L = len(sat_feats)
    dev = sat_map.device
    setup = engine.setup_from_args(a, kind, level_first)
    lam = compat.resolve_damping_tensor(a, net.damping, 3, dev).reshape(3)
    nhwc = [f.contiguous() for f in (*sat_feats, *grd_feats)]
    ext = engine.ford_extrinsics(ford["R_FL"], ford["T_FL"]) if kind == "ford" else None
    side_m = ford["side_m"] if kind == "ford" else None
    reset_uv = engine.draw_reset_uv(a.N_iters * L, sat_map.shape[0])          # same CPU-RNG consumption as the reference
    t = engine.FusedLmLoop.apply(setup, net._tables(dev), ext, side_m, reset_uv, lam, L, *nhwc)
    if a.loss_method != 0:
        raise NotImplementedError("loss_method %r is outside the accelerated path" % (a.loss_method,))
    # the trajectory is already [B, N_iters, L, (su, sv, theta)]: KITTI lat = sv, lon = su (models_kitti.py:1281-1283);
    # Ford lat = su, lon = sv (models_ford.py:823-825) -> the fused pose loss takes it as is, no re-stacking
    if kind == "kitti":
        gt = torch.stack([gt_lon, gt_lat, gt_theta], dim=-1)
        r = _loss_tuple(t, gt, (a.coe_shift_lon, a.coe_shift_lat, coe_theta), (1, 0, 2))
    else:
        gt = torch.stack([gt_lat, gt_lon, gt_theta], dim=-1)
        r = _loss_tuple(t, gt, (a.coe_shift_lat, a.coe_shift_lon, coe_theta), (0, 1, 2))
    return (*r, grd_confs)


def train_forward(net, kind, sat_map, grd_img, gt_lat, gt_lon, gt_theta, level_first, coe_theta, ford=None):
    """`forward(mode='train')` of LM_S2GP / LM_S2GP_Ford: on CUDA tensors, for the configurations the native LM backward
    covers (engine.FusedLmLoop.supports: 3 degrees of freedom, unweighted), the fused path above; otherwise the reference's own
    computation (models_kitti.py:1141-1314 / models_ford.py:652-866) — U-Nets through `VGGUnet.forward_autograd`,
    then `project_map_to_grd` -> mask -> bottom-half crop -> `LM_update` chained over (iteration, level) without
    detaching — so that `loss.backward()` reaches both U-Nets and `damping` exactly as in the reference
    (tests/test_compat_surface.py checks loss and gradients against the reference's autograd).  It runs at the
    reference's speed: the accelerated engine is the test / eval path.  Returns the reference's 14-tuple."""
    a = net.args
    if sat_map.is_cuda and getattr(net, "fused_backward", True) and \
            engine.FusedLmLoop.supports(engine.setup_from_args(a, kind, level_first)):
        return train_forward_fused(net, kind, sat_map, grd_img, gt_lat, gt_lon, gt_theta, level_first, coe_theta, ford)
    sat_feats, _ = net.SatFeatureNet.forward_autograd(sat_map)
    grd_feats, grd_confs = net.GrdFeatureNet.forward_autograd(grd_img)
    B, L = sat_map.shape[0], len(sat_feats)
    dev = sat_map.device
    su, sv, th = (torch.zeros(B, 1, device=dev) for _ in range(3))
    traj = [[None] * L for _ in range(a.N_iters)]
    order = [(it, lv) for lv in range(L) for it in range(a.N_iters)] if level_first else \
            [(it, lv) for it in range(a.N_iters) for lv in range(L)]
    for it, lv in order:
        if kind == "kitti":
            sp, _, dj, _, mask = net.project_map_to_grd(sat_feats[lv], None, su, sv, th, lv)
        else:
            sp, _, dj, _, mask = net.project_map_to_grd(sat_feats[lv], None, ford["R_FL"], ford["T_FL"], su, sv, th, lv,
                                                        ford["side_m"])
        gf = grd_feats[lv] * mask[:, None]                   # models_kitti.py:1191-1199: mask, then the bottom half only
        gc = grd_confs[lv] * mask[:, None]
        h2 = gf.shape[-2] // 2
        su, sv, th = net.LM_update(su, sv, th, sp[:, :, h2:], None, gf[:, :, h2:], gc[:, :, h2:], dj[:, :, :, h2:])
        traj[it][lv] = torch.cat([su, sv, th], dim=1)
    t = torch.stack([torch.stack(row, dim=1) for row in traj], dim=1)             # [B, N_iters, L, (su, sv, th)]
    # KITTI: shift_lats = shift_vs, shift_lons = shift_us (models_kitti.py:1281-1283); Ford: lats = us, lons = vs (:823-825)
    lats, lons = (t[..., 1], t[..., 0]) if kind == "kitti" else (t[..., 0], t[..., 1])
    r = loss_func(a.loss_method, None, None, None, lats, lons, t[..., 2], gt_lat, gt_lon, gt_theta, None, None,
                  a.coe_shift_lat, a.coe_shift_lon, coe_theta, a.coe_L1, a.coe_L2, a.coe_L3, a.coe_L4)
    return (*r, grd_confs)


def loss_func(loss_method, ref_feat_list, pred_feat_dict, gt_feat_dict, shift_lats, shift_lons, thetas,
              gt_shift_lat, gt_shift_lon, gt_theta, pred_uv_dict, gt_uv_dict,
              coe_shift_lat=100, coe_shift_lon=100, coe_theta=100, coe_L1=100, coe_L2=100, coe_L3=100, coe_L4=100):
    """models_ford.py:1041-1093, loss_method 0 (direct pose supervision).  Inputs are the
    [B, N_iters, Level] trajectories; returns the reference's 13-tuple.  Methods 1-3 need the
    materialised warped features the fused engine never builds and are out of scope."""
    if loss_method != 0:
        raise NotImplementedError("loss_method %r is outside the accelerated path" % (loss_method,))
    if shift_lats.is_cuda:
        # one kernel forward, one backward (engine.PoseLoss -> ha_pose_loss); the stack is the only torch op left
        traj = torch.stack([shift_lats, shift_lons, thetas], dim=-1)
        gt = torch.stack([gt_shift_lat, gt_shift_lon, gt_theta], dim=-1)
        return _loss_tuple(traj, gt, (coe_shift_lat, coe_shift_lon, coe_theta), (0, 1, 2))
    err = [torch.abs(t - g[:, None, None]).mean(dim=0)
           for t, g in ((shift_lats, gt_shift_lat), (shift_lons, gt_shift_lon), (thetas, gt_theta))]   # [N_iters, Level] each
    lat_e, lon_e, th_e = err
    total = coe_shift_lat * lat_e + coe_shift_lon * lon_e + coe_theta * th_e
    return (total.mean(), total[0] - total[-1], lat_e[0] - lat_e[-1], lon_e[0] - lon_e[-1], th_e[0] - th_e[-1],
            total[-1], lat_e[-1], lon_e[-1], th_e[-1], None, None, None, None)


def _loss_tuple(traj, gt, coe, order):
    """The reference's 13-tuple from the fused pose-loss kernel.  traj [B, N_iters, L, 3] / gt [B, 3] / coe share one
    component order; `order` = positions of (lat, lon, theta) in it."""
    loss, err = engine.PoseLoss.apply(traj, gt, coe)
    ilat, ilon, ith = order
    lat_e, lon_e, th_e = err[..., ilat], err[..., ilon], err[..., ith]
    total = coe[ilat] * lat_e + coe[ilon] * lon_e + coe[ith] * th_e
    return (loss, total[0] - total[-1], lat_e[0] - lat_e[-1], lon_e[0] - lon_e[-1], th_e[0] - th_e[-1],
            total[-1], lat_e[-1], lon_e[-1], th_e[-1], None, None, None, None)


class LM_S2GP_Ford(nn.Module):
    """models_ford.py:21.  forward(sat_map, grd_img_left, satmap_sidelength_meters, R_FL, T_FL, ...)."""

    KIND = "ford"

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.level = args.level
        self.N_iters = args.N_iters
        self.using_weight = args.using_weight
        self.loss_method = args.loss_method
        self.estimate_depth = getattr(args, "estimate_depth", 0)
        if self.estimate_depth:
            raise NotImplementedError("estimate_depth is outside the accelerated path")
        self.optimizer = getattr(args, "Optimizer", "LM")
        self.proj = getattr(args, "proj", "geo")
        if self.optimizer not in ("LM", "GN"):
            # models_ford.py:609-628 SGD_update indexes its [B,3] gradient with three subscripts and :600-607 NN_update adds
            # a [B] row to a [B,1] column: neither runs in the reference itself; GN (:534-598) does
            raise NotImplementedError("--Optimizer %s: LM and GN are on the accelerated path" % self.optimizer)
        if getattr(args, "dropout", 0):
            # models_ford.py:406-412 subsamples half of the pixels with a numpy permutation when dropout > 0
            raise NotImplementedError("dropout is outside the accelerated path")
        self.SatFeatureNet = VGGUnet(self.level)
        self.GrdFeatureNet = VGGUnet(self.level)
        self.damping = nn.Parameter(torch.zeros(size=(1, 3), dtype=torch.float32, requires_grad=True))   # :38-39
        self.check_status = True
        self.ori_grdH, self.ori_grdW = 256, 1024
        if self.level == 2:                                                            # :59-65: [x18, x21] with the /4 and /2 grids
            self._tables_cpu = [engine.ground_table("ford", lv, n_levels=2) for lv in range(2)]
        else:                                                                          # :45-58, polar fan (:156-170) unless 'geo'
            self._tables_cpu = [engine.ground_table("ford", lv, proj=self.proj) for lv in range(4)]
        self._tables_dev = {}
        self.last_result = None

    def _tables(self, device):
        key = (device.type, device.index)
        if key not in self._tables_dev:
            self._tables_dev[key] = [t.to(device) for t in self._tables_cpu]
        return self._tables_dev[key]

    def extract(self, sat_map, grd_img, want_conf):
        # the L2 norm of VGG.py:172-175 cancels in LM_update's renormalisation (:420-426): not computed on that path;
        # GN_update (:549-566) keeps the ground features as normalised by the U-Net
        scale = self.optimizer != "LM"
        sat = self.SatFeatureNet.pyramid(sat_map, want_conf=False, want_scale=scale)
        grd = self.GrdFeatureNet.pyramid(grd_img, want_conf=want_conf, want_scale=scale)
        return sat, grd

    def refine(self, sat, grd, satmap_sidelength_meters, R_FL, T_FL, level_first=0, pose0=None, reset_uv=None,
               want_stats=False, kernel_variant=0) -> engine.LmResult:
        setup = engine.setup_from_args(self.args, self.KIND, level_first)
        setup.kernel_variant = kernel_variant
        lam = engine.resolve_damping(self.args, self.damping, setup.dof)
        ext = engine.ford_extrinsics(R_FL, T_FL)
        res = engine.lm_run(setup, sat, grd, self._tables(sat.feats[0].device), lam, extrinsics=ext,
                            side_m=float(satmap_sidelength_meters), pose0=pose0, reset_uv=reset_uv, want_stats=want_stats)
        self.last_result = res
        return res

    def project_map_to_grd(self, sat_f, sat_c, R_FL, T_FL, shift_u, shift_v, theta, level, satmap_sidelength_meters,
                           require_jac=True, depth=None):
        """models_ford.py:266-378: materialised warp of the satellite features into the front-left camera view; returns
        (sat_f_trans, sat_c_trans, new_jac [3,B,C,H,W], uv * mask, mask).  Compatibility surface only — forward()
        fuses this into the LM step kernel and never builds these tensors."""
        if depth is not None:
            raise NotImplementedError("estimate_depth is outside the accelerated path")
        A = sat_f.shape[-1]
        a = self.args
        uv, mask, jac = compat.sat_uv_ford(self._tables(sat_f.device)[level].to(sat_f.dtype), R_FL, T_FL, shift_u, shift_v, theta,
                                           A, a.rotation_range, a.shift_range_lat, a.shift_range_lon,
                                           float(satmap_sidelength_meters) / A)
        return compat.project_map_to_grd(uv, mask, jac, sat_f, sat_c, require_jac)

    def LM_update(self, shift_u, shift_v, theta, sat_feat_proj, sat_conf_proj, grd_feat, grd_conf, dfeat_dpose):
        """models_ford.py:380-466 on materialised tensors (always 3-DOF; compatibility surface)."""
        lam = compat.resolve_damping_tensor(self.args, self.damping, 3, dfeat_dpose.device)
        return compat.lm_update_dense(shift_u, shift_v, theta, sat_feat_proj, grd_feat, grd_conf, dfeat_dpose, lam, "full",
                                      bool(self.using_weight), bool(self.args.use_hessian), redraw=True)

    def forward(self, sat_map, grd_img_left, satmap_sidelength_meters, R_FL, T_FL, gt_shift_u=None, gt_shift_v=None,
                gt_theta=None, mode='train', file_name=None, level_first=0, loop=0):
        """models_ford.py:1028-1036 -> forward_iters_level (:652-866) / forward_level_iters (:868-1026)."""
        if self.optimizer == "GN" and level_first:
            raise NotImplementedError("forward_level_iters has no GN branch (models_ford.py:933-955)")
        if mode == 'train':
            if self.optimizer != "LM" or self.proj != "geo":
                raise NotImplementedError("train mode covers --Optimizer LM --proj geo; the ablation flags run in test mode")
            ford = dict(R_FL=R_FL, T_FL=T_FL, side_m=float(satmap_sidelength_meters))
            coe_heading = 0 if self.args.rotation_range == 0 else self.args.coe_heading        # :843-846, :1000-1003
            return train_forward(self, "ford", sat_map, grd_img_left, gt_shift_u, gt_shift_v, gt_theta, level_first,
                                 coe_heading, ford)
        want_conf = bool(self.using_weight)
        sat, grd = self.extract(sat_map, grd_img_left, want_conf)
        res = self.refine(sat, grd, satmap_sidelength_meters, R_FL, T_FL, level_first)
        if self.check_status:                 # the reference's error convention, one host sync per forward (jacobian.py:172)
            engine.check_status(res.status, "LM_S2GP_Ford.forward")
        traj = res.traj
        # :823-825: shift_lats = shift_us, shift_lons = shift_vs
        shift_lats, shift_lons, thetas = _TrajectoryOutputs.apply(self.damping, False, traj[..., 0], traj[..., 1], traj[..., 2])
        return shift_lats[:, -1, -1], shift_lons[:, -1, -1], thetas[:, -1, -1]
