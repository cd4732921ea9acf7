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
    backward; it contributes nothing.  Differentiating the fused LM loop for training is the next
    scope row: until it lands a train-mode backward fails loudly instead
    of silently producing zero gradients."""

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


def loss_func(loss_method, ref_feat_list, pred_feat_dict, gt_feat_dict, shift_lats, shift_lons, thetas,
              gt_shift_lat, gt_shift_lon, gt_theta, pred_uv_dict, gt_uv_dict,
              coe_shift_lat=100, coe_shift_lon=100, coe_theta=100, coe_L1=100, coe_L2=100, coe_L3=100, coe_L4=100):
    """models_ford.py:1041-1093, loss_method 0 (direct pose supervision).  Inputs are the
    [B, N_iters, Level] trajectories; returns the reference's 13-tuple.  Methods 1-3 need the
    materialised warped features the fused engine never builds and are out of scope."""
    if loss_method != 0:
        raise NotImplementedError("loss_method %r is outside the accelerated path" % (loss_method,))
    err = [torch.abs(t - g[:, None, None]).mean(dim=0)
           for t, g in ((shift_lats, gt_shift_lat), (shift_lons, gt_shift_lon), (thetas, gt_theta))]   # [N_iters, Level] each
    lat_e, lon_e, th_e = err
    total = coe_shift_lat * lat_e + coe_shift_lon * lon_e + coe_theta * th_e
    return (total.mean(), total[0] - total[-1], lat_e[0] - lat_e[-1], lon_e[0] - lon_e[-1], th_e[0] - th_e[-1],
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
        if getattr(args, "Optimizer", "LM") != "LM" or getattr(args, "proj", "geo") != "geo":
            raise NotImplementedError("only --Optimizer LM --proj geo is on the accelerated path")
        self.SatFeatureNet = VGGUnet(self.level)
        self.GrdFeatureNet = VGGUnet(self.level)
        self.damping = nn.Parameter(torch.zeros(size=(1, 3), dtype=torch.float32, requires_grad=True))   # :38-39
        self.ori_grdH, self.ori_grdW = 256, 1024
        self._tables_cpu = [engine.ground_table("ford", lv) for lv in range(4)]       # :45-58
        self._tables_dev = {}
        self.last_result = None

    def _tables(self, device):
        key = (device.type, device.index)
        if key not in self._tables_dev:
            self._tables_dev[key] = [t.to(device) for t in self._tables_cpu]
        return self._tables_dev[key]

    def extract(self, sat_map, grd_img, want_conf):
        sat = self.SatFeatureNet.pyramid(sat_map, want_conf=False)
        grd = self.GrdFeatureNet.pyramid(grd_img, want_conf=want_conf)
        return sat, grd

    def refine(self, sat, grd, satmap_sidelength_meters, R_FL, T_FL, level_first=0, pose0=None, reset_uv=None,
               want_stats=False) -> engine.LmResult:
        setup = engine.setup_from_args(self.args, self.KIND, level_first)
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
        want_conf = bool(self.using_weight) or mode == 'train'
        sat, grd = self.extract(sat_map, grd_img_left, want_conf)
        res = self.refine(sat, grd, satmap_sidelength_meters, R_FL, T_FL, level_first)
        traj = res.traj
        # :823-825: shift_lats = shift_us, shift_lons = shift_vs
        shift_lats, shift_lons, thetas = _TrajectoryOutputs.apply(self.damping, mode == 'train', traj[..., 0],
                                                                  traj[..., 1], traj[..., 2])
        if mode == 'train':
            r = loss_func(self.args.loss_method, None, None, None, shift_lats, shift_lons, thetas,
                          gt_shift_u, gt_shift_v, gt_theta, None, None,
                          self.args.coe_shift_lat, self.args.coe_shift_lon, self.args.coe_heading,
                          self.args.coe_L1, self.args.coe_L2, self.args.coe_L3, self.args.coe_L4)
            return (*r, [c[:, None] for c in grd.confs])
        return shift_lats[:, -1, -1], shift_lons[:, -1, -1], thetas[:, -1, -1]
