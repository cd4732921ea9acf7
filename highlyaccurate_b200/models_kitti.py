"""Drop-in for the reference's models_kitti.py on the accelerated path: `LM_S2GP` and `loss_func`
with the reference's constructor / forward signatures and state-dict keys, running the fused
sm_100a engine.  Citations are into the upstream models_kitti.py.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import compat, engine
from .VGG import VGGUnet, VGGUnet_G2S
from .models_ford import loss_func, train_forward, _TrajectoryOutputs  # noqa: F401  (the reference re-exports loss_func too, :16)


class NNrefine(nn.Module):
    """RNNs.py:98-126: parameter container with the reference's names (`NNrefine.linear{0..3}.1.*`, `NNrefine.mapping.{1,3}.*`
    in the model's state dict); the computation runs in engine.nn_run."""

    def __init__(self):
        super().__init__()
        for i, c in enumerate((256, 128, 64, 16)):
            setattr(self, "linear%d" % i, nn.Sequential(nn.ReLU(inplace=True),
                                                        nn.Conv2d(c, 64, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1))))
        self.mapping = nn.Sequential(nn.ReLU(inplace=True), nn.Linear(64, 16), nn.ReLU(inplace=True), nn.Linear(16, 3), nn.Tanh())


class LM_S2GP(nn.Module):
    """models_kitti.py:598.  forward(sat_map, grd_img_left, ..., mode='test') -> (lat, lon, theta)."""

    KIND = "kitti"

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.level = args.level
        self.N_iters = args.N_iters
        self.using_weight = args.using_weight
        self.loss_method = args.loss_method
        self.optimizer = getattr(args, "Optimizer", "LM")
        self.proj = getattr(args, "proj", "geo")
        if self.optimizer not in ("LM", "SGD", "ADAM", "NN"):
            # anything else leaves the reference's own loop without an update (:1207-1254)
            raise NotImplementedError("--Optimizer %s: LM, SGD, ADAM and NN are on the accelerated path" % self.optimizer)
        if getattr(args, "dropout", 0) or getattr(args, "use_gt_depth", 0):
            raise NotImplementedError("dropout / use_gt_depth are outside the accelerated path")
        if self.level == 2:
            # models_kitti.py:624-635 indexes xyz_grds 0 -> /8 ... 3 -> /1 whatever `level` is, so the reference's own
            # level-2 run pairs the /4 features with the /8 grid and fails on shapes (SURVEY.md 8a "level support caveat")
            raise NotImplementedError("LM_S2GP level 2 is shape-inconsistent in the reference itself; use 3, 4 or -1")
        self.SatFeatureNet = VGGUnet(self.level)
        self.GrdFeatureNet = VGGUnet(self.level)
        if args.rotation_range > 0:                                   # :615-620
            self.damping = nn.Parameter(torch.zeros(size=(1, 3), dtype=torch.float32, requires_grad=True))
        else:
            self.damping = nn.Parameter(torch.zeros(size=(), dtype=torch.float32, requires_grad=True))
        # :622-635: the ground-plane lift for proj 'geo', the polar fan of grd_img2cam_polar (:684-698) for any other value
        self._tables_cpu = [engine.ground_table("kitti", lv, proj=self.proj) for lv in range(4)]
        self._tables_dev = {}
        self.meters_per_pixel = [engine.kitti_meter_per_pixel() * (2 ** (3 - lv)) for lv in range(4)]   # :637-640
        if self.optimizer == "NN":                                   # :648-649
            self.NNrefine = NNrefine()
        self.last_result = None
        # forward(mode='test') reads the device status word once per call and applies the reference's error convention
        # (AssertionError of jacobian.py:172, NaN note of :1037); set to False for a fully asynchronous forward and call
        # engine.check_status(net.last_result.status) when convenient
        self.check_status = True

    # -- helpers ---------------------------------------------------------------------------
    def _tables(self, device):
        key = (device.type, device.index)
        if key not in self._tables_dev:
            self._tables_dev[key] = [t.to(device) for t in self._tables_cpu]
        return self._tables_dev[key]

    def extract(self, sat_map, grd_img, want_conf):
        # the L2 norm of VGG.py:172-175 cancels in LM_update's renormalisation (:982-989): not computed on that path;
        # SGD_update / ADAM_update (:1056-1124) take the normalised features as they are
        scale = self.optimizer != "LM"
        sat = self.SatFeatureNet.pyramid(sat_map, want_conf=False, want_scale=scale)
        grd = self.GrdFeatureNet.pyramid(grd_img, want_conf=want_conf, want_scale=scale)
        return sat, grd

    def refine(self, sat: engine.Pyramid, grd: engine.Pyramid, level_first=0, pose0=None, reset_uv=None,
               want_stats=False, kernel_variant=0) -> engine.LmResult:
        """The LM loop on already-extracted pyramids (used by forward and by the parity tests)."""
        setup = engine.setup_from_args(self.args, self.KIND, level_first)
        setup.kernel_variant = kernel_variant
        if self.optimizer == "NN":
            nn_params = {k: v.detach() for k, v in self.NNrefine.state_dict().items()}
            res = engine.nn_run(setup, sat, grd, self._tables(sat.feats[0].device), nn_params, self.SatFeatureNet.precision,
                                pose0=pose0)
            self.last_result = res
            return res
        lam = engine.resolve_damping(self.args, self.damping, setup.dof)
        res = engine.lm_run(setup, sat, grd, self._tables(sat.feats[0].device), lam, pose0=pose0, reset_uv=reset_uv,
                            want_stats=want_stats)
        self.last_result = res
        return res

    # -- reference surface --------------------------------------------------------------------
    def project_map_to_grd(self, sat_f, sat_c, shift_u, shift_v, heading, level, require_jac=True, gt_depth=None):
        """models_kitti.py:803-937: materialised warp of the satellite features (and confidence) into the ground view,
        returns (sat_f_trans, sat_c_trans, new_jac [3,B,C,H,W], uv * mask, mask).  Compatibility surface only — the
        accelerated forward() fuses this into the LM step kernel and never builds these tensors."""
        if gt_depth is not None:
            raise NotImplementedError("use_gt_depth is outside the accelerated path")
        A = sat_f.shape[-1]
        a = self.args
        mpp = engine.kitti_meter_per_pixel() * (engine.SAT_PROCESS_SIDE / A)
        uv, mask, jac = compat.sat_uv_kitti(self._tables(sat_f.device)[level].to(sat_f.dtype), shift_u, shift_v, heading, A,
                                            a.rotation_range, a.shift_range_lat, a.shift_range_lon, mpp)
        return compat.project_map_to_grd(uv, mask, jac, sat_f, sat_c, require_jac)

    def LM_update(self, shift_u, shift_v, theta, sat_feat_proj, sat_conf_proj, grd_feat, grd_conf, dfeat_dpose):
        """models_kitti.py:939-1041 on materialised tensors (compatibility surface, see project_map_to_grd)."""
        dof = compat.dof_name(self.args, False)
        lam = compat.resolve_damping_tensor(self.args, self.damping, {"full": 3, "shift": 2, "rot": 1}[dof], dfeat_dpose.device)
        return compat.lm_update_dense(shift_u, shift_v, theta, sat_feat_proj, grd_feat, grd_conf, dfeat_dpose, lam, dof,
                                      bool(self.using_weight), bool(self.args.use_hessian), redraw=True)

    def forward(self, sat_map, grd_img_left, gt_shiftu=None, gt_shiftv=None, gt_heading=None, mode='train',
                file_name=None, gt_depth=None, loop=0, level_first=0):
        """models_kitti.py:1126-1316 (iter-first) / :1318-1492 (level-first)."""
        if mode == 'train':
            if self.optimizer != "LM" or self.proj != "geo":
                raise NotImplementedError("train mode covers --Optimizer LM --proj geo; the ablation flags run in test mode")
            coe_heading = 0 if self.args.rotation_range == 0 else self.args.coe_heading      # :1298-1301
            return train_forward(self, "kitti", sat_map, grd_img_left, gt_shiftv[:, 0], gt_shiftu[:, 0], gt_heading[:, 0],
                                 level_first, coe_heading)
        want_conf = bool(self.using_weight) and self.optimizer == "LM"
        sat, grd = self.extract(sat_map, grd_img_left, want_conf)
        res = self.refine(sat, grd, level_first)
        if self.check_status:
            engine.check_status(res.status, "LM_S2GP.forward")
        traj = res.traj
        # :1281-1283: shift_lats = shift_vs, shift_lons = shift_us
        shift_lats, shift_lons, thetas = traj[..., 1], traj[..., 0], traj[..., 2]
        shift_lats, shift_lons, thetas = _TrajectoryOutputs.apply(self.damping, False, shift_lats, shift_lons, thetas)
        return shift_lats[:, -1, -1], shift_lons[:, -1, -1], thetas[:, -1, -1]


class LM_G2SP(nn.Module):
    """models_kitti.py:22.  Ground features are warped onto the satellite plane with the full K[R|T]
    projection (:86-160) and the residual lives on the whole satellite map (:333-379).
    forward(sat_map, grd_img_left, left_camera_k, ..., mode='test') -> (lat, lon, theta)."""

    KIND = "g2sp"

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.level = args.level
        self.N_iters = args.N_iters
        self.using_weight = args.using_weight
        self.loss_method = args.loss_method
        self.proj = getattr(args, "proj", "geo")
        if self.proj not in ("geo", "nn"):
            # models_kitti.py:176-233: project_grd_to_map knows 'geo' and 'nn' only ('polar' leaves uv undefined there)
            raise NotImplementedError("LM_G2SP: --proj geo and nn are defined by the reference (models_kitti.py:176,232)")
        if self.level not in (3, 4):
            raise NotImplementedError("LM_G2SP: levels 3 and 4 are on the accelerated path")
        self.SatFeatureNet = VGGUnet(self.level)
        self.GrdFeatureNet = VGGUnet_G2S(self.level) if self.proj == "nn" else VGGUnet(self.level)     # :36-39
        self.KIND = "g2sp_nn" if self.proj == "nn" else "g2sp"
        if self.proj == "nn" and self.using_weight:
            # VGG.py:326 takes c0 from the UN-folded x15 ([B,1,32,128]) while the level-0 feature is the folded [B,256,64,64]; the
            # reference then samples that 32 x 128 map with the 64 x 64 map's coordinates (models_kitti.py:280-282)
            raise NotImplementedError("LM_G2SP --proj nn --using_weight 1: the reference pairs a 32x128 confidence with a 64x64 feature")
        self.damping = nn.Parameter(args.damping * torch.ones(size=(1, 3), dtype=torch.float32, requires_grad=True))   # :41
        self.meters_per_pixel = [engine.kitti_meter_per_pixel() * (2 ** (3 - lv)) for lv in range(4)]                     # :43-46
        self.last_result = None
        self.check_status = True

    def extract(self, sat_map, grd_img, want_conf):
        sat = self.SatFeatureNet.pyramid(sat_map, want_conf=False)
        grd = self.GrdFeatureNet.pyramid(grd_img, want_conf=want_conf)
        return sat, grd

    def refine(self, sat: engine.Pyramid, grd: engine.Pyramid, left_camera_k, ori_grd_hw=(256, 1024), pose0=None,
               want_stats=False) -> engine.LmResult:
        setup = engine.setup_from_args(self.args, self.KIND, 0)
        lam = engine.resolve_damping(self.args, self.damping, 3, self.KIND)
        B = sat.batch
        ext = left_camera_k.reshape(B, 9) if self.KIND == "g2sp" else None       # the in-plane warp of --proj nn ignores camera_k
        res = engine.lm_run(setup, sat, grd, [None] * len(sat.feats), lam, extrinsics=ext,
                            pose0=pose0, want_stats=want_stats, ori_grd_hw=ori_grd_hw)
        self.last_result = res
        return res

    def project_grd_to_map(self, grd_f, grd_c, shift_u, shift_v, heading, camera_k, satmap_sidelength, ori_grdH, ori_grdW):
        """models_kitti.py:163-287: materialised warp of the ground features (and confidence) onto the satellite plane;
        returns (grd_f_trans, grd_c_trans, new_jac [3,B,C,A,A]).  Compatibility surface only (see LM_S2GP)."""
        if self.proj != "geo":
            raise NotImplementedError("the materialising compatibility method covers --proj geo")
        A = int(satmap_sidelength)
        a = self.args
        mpp = engine.kitti_meter_per_pixel() * (engine.SAT_PROCESS_SIDE / A)
        uv, jac = compat.cam_uv_g2sp(A, shift_u, shift_v, heading, camera_k, grd_f.shape[-2], grd_f.shape[-1], ori_grdH, ori_grdW,
                                     a.rotation_range, a.shift_range_lat, a.shift_range_lon, mpp)
        f, nj = compat.sample_with_jacobian(grd_f, uv, jac)
        c = compat.sample_with_jacobian(grd_c, uv)[0] if grd_c is not None else None
        return f, c, nj

    def LM_update(self, shift_u, shift_v, heading, grd_feat_proj, grd_conf_proj, sat_feat, sat_conf, dfeat_dpose):
        """models_kitti.py:333-379 on materialised tensors (compatibility surface)."""
        lam = self.damping if getattr(self.args, "train_damping", 0) else \
            self.args.damping * torch.ones(1, 3, dtype=torch.float32, device=dfeat_dpose.device)
        return compat.lm_update_g2sp(shift_u, shift_v, heading, grd_feat_proj, grd_conf_proj, sat_feat, dfeat_dpose, lam,
                                     bool(self.using_weight))

    def _train_forward(self, sat_map, grd_img, cam_k, gt_u, gt_v, gt_h):
        """`forward(mode='train')` until the fused backward exists: the reference's computation (models_kitti.py:381-499)
        on the differentiable torch path, so that loss.backward() reaches both U-Nets and `damping`."""
        a = self.args
        sat_feats, _ = self.SatFeatureNet.forward_autograd(sat_map)
        grd_feats, grd_confs = self.GrdFeatureNet.forward_autograd(grd_img)
        B = sat_map.shape[0]
        oh, ow = grd_img.shape[-2:]
        su, sv, th = (torch.zeros(B, 1, device=sat_map.device) for _ in range(3))
        rows = []
        for it in range(a.N_iters):
            row = []
            for lv, (sf, gf, gc) in enumerate(zip(sat_feats, grd_feats, grd_confs)):
                gp, gcp, dj = self.project_grd_to_map(gf, gc, su, sv, th, cam_k, sf.shape[-1], oh, ow)
                su, sv, th = self.LM_update(su, sv, th, gp, gcp, sf, None, dj)
                row.append(torch.cat([su, sv, th], dim=1))
            rows.append(torch.stack(row, dim=1))
        t = torch.stack(rows, dim=1)                                       # [B, N_iters, L, (su, sv, th)]
        r = loss_func(a.loss_method, None, None, None, t[..., 1], t[..., 0], t[..., 2], gt_v[:, 0], gt_u[:, 0], gt_h[:, 0],
                      None, None, a.coe_shift_lat, a.coe_shift_lon, a.coe_heading, a.coe_L1, a.coe_L2, a.coe_L3, a.coe_L4)
        return (*r, grd_confs)

    def forward(self, sat_map, grd_img_left, left_camera_k, gt_shift_u=None, gt_shift_v=None, gt_heading=None,
                mode='train', file_name=None, gt_depth=None):
        """models_kitti.py:381-499."""
        if mode == 'train':
            if self.proj != "geo":
                raise NotImplementedError("LM_G2SP --proj nn runs in test mode on the accelerated path")
            return self._train_forward(sat_map, grd_img_left, left_camera_k, gt_shift_u, gt_shift_v, gt_heading)
        want_conf = bool(self.using_weight)
        sat, grd = self.extract(sat_map, grd_img_left, want_conf)
        res = self.refine(sat, grd, left_camera_k, ori_grd_hw=tuple(grd_img_left.shape[-2:]))
        if self.check_status:
            engine.check_status(res.status, "LM_G2SP.forward")
        traj = res.traj
        shift_lats, shift_lons, thetas = _TrajectoryOutputs.apply(self.damping, False, traj[..., 1], traj[..., 0],
                                                                  traj[..., 2])       # :472-474
        return shift_lats[:, -1, -1], shift_lons[:, -1, -1], thetas[:, -1, -1]
