"""Pin the oracle against the UNMODIFIED reference and write tests/golden/*.npz.

Runs only in the build container (needs /root/reference; read-only, nothing is written there).
For every known-answer test it executes the reference's own functions
(`jacobian.grid_sample`, `LM_S2GP.grd2cam2world2sat`, `LM_S2GP.LM_update`, `LM_S2GP.forward`,
`LM_S2GP_Ford.*`, `VGGUnet.forward`) and oracle/oracle.py on identical seeded inputs, asserts
they agree, and stores the REFERENCE outputs (small) as fixtures.  Inputs are regenerated from
seeds at test time; each fixture carries an input checksum so an RNG drift is detected.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    """SURVEY 8c obstacles: vgg16(pretrained=True) needs the network -> weights=None."""
    sys.dont_write_bytecode = True
    import torchvision
    orig = torchvision.models.vgg16
    torchvision.models.vgg16 = lambda pretrained=False, **kw: orig(weights=None)
    sys.path.insert(0, REF)
    import jacobian as ref_jac          # noqa
    import models_kitti as ref_kitti    # noqa
    import models_ford as ref_ford      # noqa
    import VGG as ref_vgg               # noqa
    torch.autograd.set_detect_anomaly(False)
    return ref_jac, ref_kitti, ref_ford, ref_vgg


def ref_args(**kw):
    d = dict(level=3, N_iters=5, using_weight=0, loss_method=0, rotation_range=10.0, proj="geo", Optimizer="LM",
             damping=0.1, train_damping=0, shift_range_lat=20.0, shift_range_lon=20.0, use_hessian=0, dropout=0,
             use_gt_depth=0, visualize=0, coe_shift_lat=100.0, coe_shift_lon=100.0, coe_heading=100.0,
             coe_L1=100.0, coe_L2=100.0, coe_L3=100.0, coe_L4=100.0, estimate_depth=0, level_first=0,
             beta1=0.9, beta2=0.999)
    d.update(kw)
    return types.SimpleNamespace(**d)


def o_args(a) -> O.LMArgs:
    return O.LMArgs(level=a.level, N_iters=a.N_iters, using_weight=a.using_weight, damping=a.damping,
                    train_damping=a.train_damping, rotation_range=a.rotation_range,
                    shift_range_lat=a.shift_range_lat, shift_range_lon=a.shift_range_lon,
                    use_hessian=a.use_hessian, level_first=a.level_first, Optimizer=a.Optimizer, proj=a.proj,
                    beta1=a.beta1, beta2=a.beta2)


def csum(*ts) -> np.ndarray:
    return np.array([float(t.double().sum()) for t in ts] + [float(t.double().abs().sum()) for t in ts])


def close(a, b, tol, what):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    d = float((a - b).abs().max())
    scale = max(float(b.abs().max()), 1e-30)
    assert d <= tol * max(scale, 1.0), "%s: oracle vs reference max|d|=%g (scale %g, tol %g)" % (what, d, scale, tol)
    return d


# --------------------------------------------------------------------------------------- KATs
def kat1_sampler(rj):
    """jacobian.py:216-239 (the reference's own commented test) + border cases."""
    g = torch.Generator().manual_seed(11)
    img = torch.rand(1, 3, 32, 32, generator=g)
    grid = torch.rand(1, 32, 32, 2, generator=g) * 2 - 1
    uv = (grid + 1) / 2 * 31
    edge = torch.tensor([-0.5, 0.0, 1e-3, 15.5, 30.0, 31 - 1e-3, 31.0, 31.5])
    ex, ey = torch.meshgrid(edge, edge, indexing="ij")
    uv_edge = torch.stack([ex, ey], dim=-1)[None]
    jac = torch.randn(3, 1, 32, 32, 2, generator=g)
    jac_e = torch.randn(3, 1, 8, 8, 2, generator=g)
    out = {}
    for name, u, j in (("rand", uv, jac), ("edge", uv_edge, jac_e)):
        r_out, r_jac = rj.grid_sample(img, u, j)
        o_out, o_jac = O.bilinear_sample(img, u, j)
        close(o_out, r_out, 0, "kat1 %s out" % name)
        close(o_jac, r_jac, 0, "kat1 %s jac" % name)
        out[name + "_uv"], out[name + "_jacin"] = u.numpy(), j.numpy()
        out[name + "_out"], out[name + "_jac"] = r_out.numpy(), r_jac.numpy()
    f_out = torch.nn.functional.grid_sample(img, grid, align_corners=True)
    close(out["rand_out"], f_out, 1e-6, "kat1 vs F.grid_sample")
    out["img"] = img.numpy()
    np.savez_compressed(os.path.join(GOLD, "kat1_sampler.npz"), **out)
    print("KAT-1 sampler ok")


POSES = [(0.0, 0.0, 0.0), (1.0, -1.0, 1.0), (0.3, -0.25, 0.5), (-0.7, 0.9, -0.4)]


def kat2_geometry(rk, rf):
    """models_kitti.py:700-801 and models_ford.py:173-264 at several poses, all levels.
    Stores a strided subsample of uv + Jacobians (full-field equality is asserted here)."""
    a = ref_args(shift_range_lat=20.0, shift_range_lon=15.0, level=4)
    net = rk.LM_S2GP(a); torch.autograd.set_detect_anomaly(False)
    netf = rf.LM_S2GP_Ford(a); torch.autograd.set_detect_anomaly(False)
    oa = o_args(a)
    pose = torch.tensor(POSES, dtype=torch.float32)
    su, sv, th = pose[:, 0:1], pose[:, 1:2], pose[:, 2:3]
    B = pose.shape[0]
    R_FL = torch.tensor([[0., 0., 1.], [1., 0., 0.], [0., 1., 0.]])[None].repeat(B, 1, 1)
    T_FL = torch.tensor([1.7, -0.3, -1.5])[None].repeat(B, 1)
    out = {"poses": pose.numpy(), "R_FL": R_FL.numpy(), "T_FL": T_FL.numpy()}
    for lv in range(4):
        A = 512 // (2 ** (3 - lv))
        r = net.grd2cam2world2sat(su, sv, th, lv, A, require_jac=True)
        tab = O.kitti_ground_table(lv)
        close(tab[0], net.xyz_grds[lv][0][0], 0, "kitti table")
        o = O.kitti_sat_uv(tab[0], tab[1], su, sv, th, A, oa)
        for i, nm in enumerate(["uv", "mask", "ju", "jv", "jt"]):
            close(o[i], r[i], 0, "kat2 kitti L%d %s" % (lv, nm))
            out["kitti_L%d_%s" % (lv, nm)] = r[i].detach()[:, ::4, ::8].numpy()
        out["kitti_L%d_tab" % lv] = tab[0][::4, ::8].numpy()
        Af = 1280 // (2 ** (3 - lv))
        side = 1280 * 0.22
        rfo = netf.cam2body2world2sat(R_FL, T_FL, su, sv, th, lv, side, Af, require_jac=True)
        tabf = O.ford_ground_table(lv)
        close(tabf[0], netf.xyz_grds[lv][0][0], 0, "ford table")
        of = O.ford_sat_uv(tabf[0], tabf[1], R_FL, T_FL, su, sv, th, Af, side, oa)
        for i, nm in enumerate(["uv", "mask", "ju", "jv", "jt"]):
            close(of[i], rfo[i], 0, "kat2 ford L%d %s" % (lv, nm))
            out["ford_L%d_%s" % (lv, nm)] = rfo[i].detach()[:, ::4, ::8].numpy()
        out["ford_L%d_tab" % lv] = tabf[0][::4, ::8].numpy()
    np.savez_compressed(os.path.join(GOLD, "kat2_geometry.npz"), **out)
    print("KAT-2 geometry ok")


def run_ref_loop(net, kind, sat, grd, conf, a, ford=None, pose0=None):
    """Drive the reference's own project_map_to_grd + LM_update exactly as its forward does
    (models_kitti.py:1176-1260), recording pose_in/pose_out per step."""
    L = len(sat)
    B = sat[0].shape[0]
    if pose0 is None:
        su = torch.zeros(B, 1); sv = torch.zeros(B, 1); th = torch.zeros(B, 1)
    else:
        su, sv, th = [p.clone() for p in pose0]
    traj = torch.zeros(B, a.N_iters, L, 3)
    pin = torch.zeros(B, a.N_iters, L, 3)
    order = O._step_order(a.N_iters, L, a.level_first)
    for it, lv in order:
        if kind == "kitti":
            sp, sc, dj, uv, mask = net.project_map_to_grd(sat[lv], None, su, sv, th, lv)
        else:
            sp, sc, dj, uv, mask = net.project_map_to_grd(sat[lv], None, ford["R_FL"], ford["T_FL"], su, sv, th, lv,
                                                           ford["side_m"], require_jac=True)
        gf = grd[lv] * mask[:, None]
        gc = conf[lv] * mask[:, None]
        h2 = gf.shape[-2] // 2 if a.proj == "geo" else 0            # models_kitti.py:1194-1205
        pin[:, it, lv] = torch.cat([su, sv, th], dim=1)
        step_in = (su, sv, th, sp[:, :, h2:], gc[:, :, h2:], gf[:, :, h2:], gc[:, :, h2:], dj[:, :, :, h2:])
        if a.Optimizer == "LM":
            su, sv, th = net.LM_update(*step_in)
        elif a.Optimizer == "SGD":                                   # :1214-1232
            su, sv, th = net.SGD_update(*step_in)
        elif a.Optimizer == "ADAM":                                  # :1240-1253
            t = it * a.level + lv
            if t == 0:
                adam_m, adam_v = 0, 0
            su, sv, th, adam_m, adam_v = net.ADAM_update(*step_in, adam_m, adam_v, t)
        elif a.Optimizer == "GN":                                    # models_ford.py:775-781
            su, sv, th = net.GN_update(*step_in)
        elif a.Optimizer == "NN":                                    # :1233-1239
            su, sv, th = net.NN_update(*step_in)
        su, sv, th = su.detach(), sv.detach(), th.detach()
        traj[:, it, lv] = torch.cat([su, sv, th], dim=1)
    return traj, pin


def stats_arrays(res: O.LoopResult, n_iters, L):
    ks = ["hessian", "grad", "sat_norm", "grd_norm", "res_sq", "delta"]
    out = {}
    for k in ks:
        out[k] = np.stack([np.stack([getattr(res.stats[it][lv], k).numpy() for lv in range(L)], 0)
                           for it in range(n_iters)], 0)
    return out


def fp64_truth(kind, sat, grd, conf, oa, damp, ford, pose0, r_pin, nn_sd=None):
    """The same algorithm in float64 = the truth the fp32 reference itself deviates from
    (SURVEY 8c noise floor).  traj64: whole loop with the same draws; step64/hess64/grad64/delta64:
    every step restarted from the REFERENCE's fp32 pose state."""
    dd = lambda xs: [x.double() for x in xs]
    s64, g64, c64 = dd(sat), dd(grd), dd(conf)
    f64 = None if ford is None else dict(R_FL=ford["R_FL"].double(), T_FL=ford["T_FL"].double(), side_m=ford["side_m"])
    p64 = None if pose0 is None else tuple(p.double() for p in pose0)
    torch.manual_seed(4242)
    nn64 = None if nn_sd is None else {k: v.double() for k, v in nn_sd.items()}
    res = O.lm_loop(kind, s64, g64, c64, oa, None if damp is None else damp.double(), None, f64, p64, nn_sd=nn64)
    if kind == "kitti":
        traj64 = torch.stack([res.lons, res.lats, res.thetas], dim=-1)
    else:
        traj64 = torch.stack([res.lats, res.lons, res.thetas], dim=-1)
    L, B = len(sat), sat[0].shape[0]
    if oa.Optimizer == "ADAM":          # the Adam moments make a step depend on the whole history: no per-step restart
        return dict(traj64=traj64.numpy())
    nd = O.n_dof(kind, oa) if oa.Optimizer in ("LM", "GN") else 3
    lam = O.resolve_damping(oa, None if damp is None else damp.double(), nd, torch.float64)
    step = torch.zeros(B, oa.N_iters, L, 3, dtype=torch.float64)
    hess = torch.zeros(oa.N_iters, L, B, nd, nd, dtype=torch.float64)
    grad = torch.zeros(oa.N_iters, L, B, nd, dtype=torch.float64)
    delta = torch.zeros(oa.N_iters, L, B, nd, dtype=torch.float64)
    zero = (torch.zeros(B, 1, dtype=torch.float64), torch.zeros(B, 1, dtype=torch.float64))
    for it in range(oa.N_iters):
        for lv in range(L):
            pin = r_pin[:, it, lv].double()
            tab = tuple(t.double() for t in O.ground_table(kind, lv, L, oa.proj))
            su, sv, th, st = O.lm_one_step(kind, s64[lv], g64[lv], c64[lv], tab, pin[:, 0:1], pin[:, 1:2], pin[:, 2:3],
                                           oa, lam, zero, f64, dict(m=0, v=0, t=0, nn=nn64))
            step[:, it, lv] = torch.cat([su, sv, th], dim=1)
            hess[it, lv], grad[it, lv], delta[it, lv] = st.hessian, st.grad, st.delta
    return dict(traj64=traj64.numpy(), step64=step.numpy(), hess64=hess.numpy(), grad64=grad.numpy(),
                delta64=delta.numpy())


def kat_loop(rk, rf, name, kind, make_inputs, tol=2e-6, pose0=None, **akw):
    """Whole-loop KAT: reference trajectory vs oracle trajectory on identical inputs and
    identical CPU-RNG state; stores the reference trajectory + oracle per-step stats."""
    a = ref_args(**akw)
    net = (rk.LM_S2GP if kind == "kitti" else rf.LM_S2GP_Ford)(a)
    torch.autograd.set_detect_anomaly(False)
    oa = o_args(a)
    sat, grd, conf, ford, meta = make_inputs(oa)
    damp = None
    if a.train_damping:
        damp = net.damping.detach().clone()
    nn_sd = None
    if a.Optimizer == "NN":                                # seeded NNrefine weights (RNNs.py:98-116) loaded into the reference
        nn_sd = O.nnrefine_state_dict(int(meta["seed"]) + 1000)
        net.NNrefine.load_state_dict(nn_sd)
    torch.manual_seed(4242)
    with torch.no_grad():
        r_traj, r_pin = run_ref_loop(net, kind, sat, grd, conf, a, ford, pose0)
    torch.manual_seed(4242)
    res = O.lm_loop(kind, sat, grd, conf, oa, damp, None, ford, pose0, nn_sd=nn_sd)
    # reference loop records (su,sv,th); oracle LoopResult is (lat,lon,theta) in the model's convention
    if kind == "kitti":
        o_traj = torch.stack([res.lons, res.lats, res.thetas], dim=-1)
    else:
        o_traj = torch.stack([res.lats, res.lons, res.thetas], dim=-1)
    d = close(o_traj, r_traj, tol, name + " trajectory")
    out = dict(traj=r_traj.numpy(), pose_in=r_pin.numpy(), in_csum=csum(*sat, *grd), **stats_arrays(res, a.N_iters, len(sat)))
    out.update(fp64_truth(kind, sat, grd, conf, oa, damp, ford, pose0, r_pin, nn_sd))
    out.update(meta)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("%s ok  (oracle vs reference max|d| = %.2e; final pose sample0 = %s)" % (name, d, r_traj[0, -1, -1].tolist()))



def kat_g2sp(rk, name, make_inputs, **akw):
    """LM_G2SP (models_kitti.py:22-499): the reference hard-codes .cuda() (:59,68,73); run it on CPU
    with Tensor.cuda patched to the identity (SURVEY 8c obstacle 2).  Reference trajectory + fp64 truth."""
    a = ref_args(**akw)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *x, **k: self
    try:
        net = rk.LM_G2SP(a)
        torch.autograd.set_detect_anomaly(False)
        oa = o_args(a)
        sat, grd, conf, meta = make_inputs(oa)
        B, L = sat[0].shape[0], len(sat)
        cam_k = torch.tensor([O._KITTI_K], dtype=torch.float32).repeat(B, 1, 1)
        su = torch.zeros(B, 1); sv = torch.zeros(B, 1); th = torch.zeros(B, 1)
        traj = torch.zeros(B, a.N_iters, L, 3)
        pin = torch.zeros(B, a.N_iters, L, 3)
        with torch.no_grad():
            for it in range(a.N_iters):
                for lv in range(L):
                    A = sat[lv].shape[-1]
                    gp, gcp, dj = net.project_grd_to_map(grd[lv], conf[lv], su, sv, th, cam_k, A, 256, 1024)
                    pin[:, it, lv] = torch.cat([su, sv, th], dim=1)
                    su, sv, th = net.LM_update(su, sv, th, gp, gcp, sat[lv], None, dj)
                    traj[:, it, lv] = torch.cat([su, sv, th], dim=1)
    finally:
        torch.Tensor.cuda = orig_cuda
    res = O.lm_loop_g2sp(sat, grd, conf, cam_k, oa)
    o_traj = torch.stack([res.lons, res.lats, res.thetas], dim=-1)
    d = close(o_traj, traj, 2e-6, name + " trajectory")
    dd = lambda xs: [x.double() for x in xs]
    r64 = O.lm_loop_g2sp(dd(sat), dd(grd), dd(conf), cam_k.double(), oa)
    traj64 = torch.stack([r64.lons, r64.lats, r64.thetas], dim=-1)
    lam = oa.damping * torch.ones(1, 3, dtype=torch.float64)
    step64 = torch.zeros(B, a.N_iters, L, 3, dtype=torch.float64)
    hess64 = torch.zeros(a.N_iters, L, B, 3, 3, dtype=torch.float64)
    grad64 = torch.zeros(a.N_iters, L, B, 3, dtype=torch.float64)
    for it in range(a.N_iters):
        for lv in range(L):
            p = pin[:, it, lv].double()
            u, v, t, st = O.g2sp_one_step(sat[lv].double(), grd[lv].double(), conf[lv].double(), cam_k.double(),
                                          p[:, 0:1], p[:, 1:2], p[:, 2:3], 256, 1024, oa, lam)
            step64[:, it, lv] = torch.cat([u, v, t], dim=1)
            hess64[it, lv], grad64[it, lv] = st.hessian, st.grad
    out = dict(traj=traj.numpy(), pose_in=pin.numpy(), in_csum=csum(*sat, *grd), traj64=traj64.numpy(), step64=step64.numpy(),
               hess64=hess64.numpy(), grad64=grad64.numpy(), cam_k=cam_k.numpy(), **stats_arrays(res, a.N_iters, L))
    out.update(meta)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("%s ok  (oracle vs reference max|d| = %.2e; final pose sample0 = %s)" % (name, d, traj[0, -1, -1].tolist()))


FORD_EXT = dict(R=[[0., 0., 1.], [1., 0., 0.], [0., 1., 0.]], T=[1.7, -0.3, -1.5])


def kat8_train_gradients(rk):
    """KAT-8 (groundwork for SURVEY 8 f-1, the backward of the fused loop): the reference's own train-mode loop —
    project_map_to_grd + LM_update chained WITHOUT detaching (models_kitti.py:1176-1260), loss_func method 0
    (:1296-1314) — differentiated by autograd w.r.t. both feature pyramids and the trained damping.  Stored: the
    loss, per-tensor gradient sums / abs-sums and the gradient at 64 seeded positions per tensor."""
    a = ref_args(N_iters=2, train_damping=1)
    oa = o_args(a)
    gt = [[0.3, -0.25, 0.5]]
    seed, B, A, L = 77, 1, 512, 3
    sat, grd = O.planted_case("kitti", B, A, L, seed, gt, oa)
    sat = [s.clone().requires_grad_(True) for s in sat]
    grd = [g.clone().requires_grad_(True) for g in grd]
    torch.manual_seed(0)
    net = rk.LM_S2GP(a)
    torch.autograd.set_detect_anomaly(False)
    torch.manual_seed(4242)
    su = torch.zeros(B, 1); sv = torch.zeros(B, 1); th = torch.zeros(B, 1)
    us_all, vs_all, th_all = [], [], []
    for it in range(a.N_iters):
        us, vs, ths = [], [], []
        for lv in range(L):
            sp, _, dj, _, mask = net.project_map_to_grd(sat[lv], None, su, sv, th, lv)
            gf = grd[lv] * mask[:, None]
            gc = torch.ones(B, 1, *gf.shape[-2:]) * mask[:, None]
            h2 = gf.shape[-2] // 2
            su, sv, th = net.LM_update(su, sv, th, sp[:, :, h2:], gc[:, :, h2:], gf[:, :, h2:], gc[:, :, h2:], dj[:, :, :, h2:])
            us.append(su[:, 0]); vs.append(sv[:, 0]); ths.append(th[:, 0])
        us_all.append(torch.stack(us, dim=1)); vs_all.append(torch.stack(vs, dim=1)); th_all.append(torch.stack(ths, dim=1))
    lats, lons, thetas = torch.stack(vs_all, dim=1), torch.stack(us_all, dim=1), torch.stack(th_all, dim=1)
    g = torch.tensor(gt)
    out = rk.loss_func(0, None, None, None, lats, lons, thetas, g[:, 1], g[:, 0], g[:, 2], None, None,
                       a.coe_shift_lat, a.coe_shift_lon, a.coe_heading, a.coe_L1, a.coe_L2, a.coe_L3, a.coe_L4)
    loss = out[0]
    loss.backward()
    rec = dict(seed=seed, B=B, A=A, L=L, gt=np.array(gt, dtype=np.float32), loss=np.float32(loss.item()),
               traj=torch.stack([lons, lats, thetas], dim=-1).detach().numpy(), damping_grad=net.damping.grad.numpy(),
               in_csum=csum(*[t.detach() for t in sat], *[t.detach() for t in grd]))
    gidx = torch.Generator().manual_seed(99)
    for name, ts in (("sat", sat), ("grd", grd)):
        for lv, t in enumerate(ts):
            gflat = t.grad.reshape(-1)
            idx = torch.randint(0, gflat.numel(), (64,), generator=gidx)
            top = torch.topk(gflat.abs(), 32).indices                      # the largest entries carry the signal
            idx = torch.cat([idx, top])
            rec["%s%d_idx" % (name, lv)] = idx.numpy()
            rec["%s%d_val" % (name, lv)] = gflat[idx].numpy()
            rec["%s%d_sum" % (name, lv)] = np.array([float(gflat.double().sum()), float(gflat.double().abs().sum())])
    np.savez_compressed(os.path.join(GOLD, "kat8_train_grad.npz"), **rec)
    print("kat8_train_grad: loss %.6f  |dL/ddamping| %s" % (loss.item(), np.abs(rec["damping_grad"]).ravel()))


E2E_TRAIN_PARAMS = ["SatFeatureNet.conv0.weight", "SatFeatureNet.conv_dec2.3.weight", "SatFeatureNet.conv14.bias",
                    "GrdFeatureNet.conv14.weight", "GrdFeatureNet.conv_dec1.1.weight", "GrdFeatureNet.conv2.bias"]


def kat9_train_e2e(rk, name="kat9_train_e2e", level_first=0, **akw):
    """KAT-9: the reference's LM_S2GP.forward(mode='train') end to end (both U-Nets + N_iters x 3 levels + loss_func
    method 0) and its autograd gradients into U-Net weights of both branches: what train_kitti.py:354-365 computes."""
    a = ref_args(**dict(dict(N_iters=1), **akw))
    torch.manual_seed(0)
    net = rk.LM_S2GP(a)
    sd = {}
    sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
    sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
    sd["damping"] = torch.zeros(1, 3)
    net.load_state_dict(sd)
    torch.autograd.set_detect_anomaly(False)
    g = torch.Generator().manual_seed(2022)
    sat = torch.rand(1, 3, 512, 512, generator=g)
    grd = torch.rand(1, 3, 256, 1024, generator=g)
    gt = torch.tensor([[0.3, -0.25, 0.5]])
    torch.manual_seed(4242)
    out = net(sat, grd, gt[:, 0:1], gt[:, 1:2], gt[:, 2:3], mode="train", level_first=level_first)
    out[0].backward()
    rec = dict(gt=gt.numpy(), loss=np.float32(out[0].item()), loss_last=out[5].detach().numpy(),
               lat_last=out[6].detach().numpy(), lon_last=out[7].detach().numpy(), theta_last=out[8].detach().numpy(),
               loss_decrease=out[1].detach().numpy(), damping_grad=net.damping.grad.numpy() if net.damping.grad is not None else np.zeros(3))
    gidx = torch.Generator().manual_seed(123)
    params = dict(net.named_parameters())
    rec["n_conf"] = len(out[13])
    for k, pname in enumerate(E2E_TRAIN_PARAMS):
        if params[pname].grad is None:
            continue
        gflat = params[pname].grad.reshape(-1)
        idx = torch.cat([torch.randint(0, gflat.numel(), (32,), generator=gidx), torch.topk(gflat.abs(), min(32, gflat.numel())).indices])
        rec["p%d_idx" % k] = idx.numpy()
        rec["p%d_val" % k] = gflat[idx].numpy()
        rec["p%d_sum" % k] = np.array([float(gflat.double().sum()), float(gflat.double().abs().sum())])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print("%s: loss %.6f, last (lat, lon, theta) errors %s %s %s" % (name, out[0].item(), rec["lat_last"], rec["lon_last"], rec["theta_last"]))


def kat9_train_e2e_ford(rf, name="kat9_train_e2e_ford", **akw):
    """KAT-9 for LM_S2GP_Ford.forward(mode='train') (train_ford.py:229-240): losses + autograd gradients."""
    a = ref_args(**dict(dict(N_iters=1), **akw))
    torch.manual_seed(0)
    net = rf.LM_S2GP_Ford(a)
    sd = {}
    sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
    sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
    sd["damping"] = torch.zeros(1, 3)
    net.load_state_dict(sd)
    torch.autograd.set_detect_anomaly(False)
    g = torch.Generator().manual_seed(2023)
    sat = torch.rand(1, 3, 512, 512, generator=g)
    grd = torch.rand(1, 3, 256, 1024, generator=g)
    fd = ford_dict(1, 0.22 * 512)
    gt = torch.tensor([[0.2, -0.3, 0.4]])
    torch.manual_seed(4242)
    out = net(sat, grd, fd["side_m"], fd["R_FL"], fd["T_FL"], gt[:, 0], gt[:, 1], gt[:, 2], mode="train")
    out[0].backward()
    rec = dict(gt=gt.numpy(), side_m=np.float32(fd["side_m"]), loss=np.float32(out[0].item()), loss_last=out[5].detach().numpy(),
               lat_last=out[6].detach().numpy(), lon_last=out[7].detach().numpy(), theta_last=out[8].detach().numpy())
    gidx = torch.Generator().manual_seed(123)
    params = dict(net.named_parameters())
    for k, pname in enumerate(E2E_TRAIN_PARAMS):
        if params[pname].grad is None:
            continue
        gflat = params[pname].grad.reshape(-1)
        idx = torch.cat([torch.randint(0, gflat.numel(), (32,), generator=gidx), torch.topk(gflat.abs(), min(32, gflat.numel())).indices])
        rec["p%d_idx" % k] = idx.numpy()
        rec["p%d_val" % k] = gflat[idx].numpy()
        rec["p%d_sum" % k] = np.array([float(gflat.double().sum()), float(gflat.double().abs().sum())])
    rec["n_conf"] = len(out[13])
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **rec)
    print("%s: loss %.6f" % (name, out[0].item()))


def kat9_train_e2e_g2sp(rk):
    """KAT-9 for LM_G2SP.forward(mode='train') (train_kitti.py:363-365) on CPU with Tensor.cuda patched to the identity."""
    a = ref_args(N_iters=1)
    g = torch.Generator().manual_seed(2024)
    sat = torch.rand(1, 3, 512, 512, generator=g)
    grd = torch.rand(1, 3, 256, 1024, generator=g)
    cam_k = torch.tensor([O._KITTI_K], dtype=torch.float32)
    gt = torch.tensor([[0.25, -0.2, 0.35]])
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *x, **k: self
    try:
        torch.manual_seed(0)
        net = rk.LM_G2SP(a)
        torch.autograd.set_detect_anomaly(False)
        sd = {}
        sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
        sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
        sd["damping"] = net.damping.detach().clone()
        net.load_state_dict(sd)
        out = net(sat, grd, cam_k, gt[:, 0:1], gt[:, 1:2], gt[:, 2:3], mode="train")
        out[0].backward()
    finally:
        torch.Tensor.cuda = orig_cuda
    rec = dict(gt=gt.numpy(), cam_k=cam_k.numpy(), loss=np.float32(out[0].item()), loss_last=out[5].detach().numpy(),
               lat_last=out[6].detach().numpy(), lon_last=out[7].detach().numpy(), theta_last=out[8].detach().numpy())
    gidx = torch.Generator().manual_seed(123)
    params = dict(net.named_parameters())
    for k, name in enumerate(E2E_TRAIN_PARAMS):
        gflat = params[name].grad.reshape(-1)
        idx = torch.cat([torch.randint(0, gflat.numel(), (32,), generator=gidx), torch.topk(gflat.abs(), min(32, gflat.numel())).indices])
        rec["p%d_idx" % k] = idx.numpy()
        rec["p%d_val" % k] = gflat[idx].numpy()
    np.savez_compressed(os.path.join(GOLD, "kat9_train_e2e_g2sp.npz"), **rec)
    print("kat9_train_e2e_g2sp: loss %.6f" % out[0].item())


def signatures(rk, rf, rv):
    """Parameter names and defaults of the reference's public surface for the hot path, as a JSON fixture:
    the drop-in modules must expose the same callables with the same parameters."""
    import inspect
    import json

    def sig(f):
        return [[n, None if p.default is inspect._empty else repr(p.default)] for n, p in inspect.signature(f).parameters.items()]

    out = {}
    for cls_name, cls in (("LM_S2GP", rk.LM_S2GP), ("LM_G2SP", rk.LM_G2SP), ("LM_S2GP_Ford", rf.LM_S2GP_Ford)):
        for m in ("__init__", "forward", "LM_update", "project_map_to_grd", "project_grd_to_map"):
            if hasattr(cls, m):
                out["%s.%s" % (cls_name, m)] = sig(getattr(cls, m))
    out["models_kitti.loss_func"] = sig(rk.loss_func)
    out["models_ford.loss_func"] = sig(rf.loss_func)
    out["VGGUnet.__init__"] = sig(rv.VGGUnet.__init__)
    out["VGGUnet.forward"] = sig(rv.VGGUnet.forward)
    out["L2_norm"] = sig(rv.L2_norm)
    with open(os.path.join(GOLD, "signatures.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("signatures: %d callables" % len(out))


def ford_dict(B, side_m):
    return dict(R_FL=torch.tensor(FORD_EXT["R"])[None].repeat(B, 1, 1),
                T_FL=torch.tensor(FORD_EXT["T"])[None].repeat(B, 1), side_m=side_m)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    opt = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    rj, rk, rf, rv = import_reference()
    want = lambda k: (not opt.only) or (k in opt.only.split(","))

    if want("signatures"):
        signatures(rk, rf, rv)
    if want("kat1"):
        kat1_sampler(rj)
    if want("kat2"):
        kat2_geometry(rk, rf)
    if want("kat8"):
        kat8_train_gradients(rk)
    if want("kat9"):
        kat9_train_e2e(rk)
        kat9_train_e2e(rk, "kat9_train_e2e_weighted_levelfirst", level_first=1, N_iters=2, using_weight=1, train_damping=1)
        kat9_train_e2e(rk, "kat9_train_e2e_level_m1", level=-1, N_iters=2)
        kat9_train_e2e_ford(rf)
        kat9_train_e2e_ford(rf, "kat9_train_e2e_ford_level2", level=2, N_iters=2)
        kat9_train_e2e_ford(rf, "kat9_train_e2e_ford_rot0", rotation_range=0.0)      # coe_heading = 0 (models_ford.py:843-846)
        kat9_train_e2e_g2sp(rk)

    GT2 = [[0.3, -0.25, 0.5], [-0.2, 0.4, -0.3]]

    def rand_inputs(seed, B=2, A=512, L=3):
        def f(oa):
            sat, grd, conf = O.random_pyramid(B, A, L, seed)
            return sat, grd, conf, None, dict(seed=seed, B=B, A=A, L=L)
        return f

    def planted_inputs(kind, seed, gt, A=512, L=3, side_m=None, l2=False):
        def f(oa):
            B = len(gt)
            fd = ford_dict(B, side_m) if kind == "ford" else None
            sat, grd = O.planted_case(kind, B, A, L, seed, gt, oa, fd, l2=l2)
            conf = [torch.ones(B, 1, *g.shape[-2:]) for g in grd]
            return sat, grd, conf, fd, dict(seed=seed, B=B, A=A, L=L, gt=np.array(gt, dtype=np.float32),
                                            side_m=np.float32(side_m or 0))
        return f

    if want("kat3"):   # per-step parity on non-contractive features
        kat_loop(rk, rf, "kat3_random_kitti", "kitti", rand_inputs(31))
    if want("kat4"):   # planted pose, whole trajectory
        kat_loop(rk, rf, "kat4_planted_kitti", "kitti", planted_inputs("kitti", 41, GT2))
        kat_loop(rk, rf, "kat4_planted_ford", "ford", planted_inputs("ford", 42, GT2, A=512, side_m=512 * 0.22))
    if want("kat5"):   # modes
        kat_loop(rk, rf, "kat5_weight", "kitti", rand_inputs(51), using_weight=1)
        kat_loop(rk, rf, "kat5_hessian", "kitti", rand_inputs(52), use_hessian=1)
        kat_loop(rk, rf, "kat5_traindamp", "kitti", planted_inputs("kitti", 53, GT2), train_damping=1)
        kat_loop(rk, rf, "kat5_levelfirst", "kitti", planted_inputs("kitti", 54, GT2), level_first=1)
        kat_loop(rk, rf, "kat5_shiftonly", "kitti", planted_inputs("kitti", 55, [[0.3, -0.25, 0.0], [-0.2, 0.4, 0.0]]),
                 rotation_range=0.0)
        kat_loop(rk, rf, "kat5_rotonly", "kitti", planted_inputs("kitti", 56, [[0.0, 0.0, 0.5], [0.0, 0.0, -0.3]]),
                 shift_range_lat=0.0, shift_range_lon=0.0)
        kat_loop(rk, rf, "kat5_level4", "kitti", planted_inputs("kitti", 57, GT2[:1], L=4), level=4, N_iters=2)
        kat_loop(rk, rf, "kat5_ford1280", "ford", planted_inputs("ford", 58, GT2[:1], A=1280, side_m=1280 * 0.22),
                 N_iters=3)
        kat_loop(rk, rf, "kat5_anisotropic", "kitti", planted_inputs("kitti", 59, GT2), shift_range_lat=20.0,
                 shift_range_lon=12.0, rotation_range=15.0)
    if want("kat10"):  # optimiser ablations and the polar ground table (non-default flags)
        kat_loop(rk, rf, "kat10_sgd", "kitti", planted_inputs("kitti", 101, GT2, l2=True), Optimizer="SGD", N_iters=3)
        kat_loop(rk, rf, "kat10_adam", "kitti", planted_inputs("kitti", 102, GT2, l2=True), Optimizer="ADAM", N_iters=3)
        kat_loop(rk, rf, "kat10_adam_level4", "kitti", planted_inputs("kitti", 103, GT2[:1], L=4, l2=True), Optimizer="ADAM",
                 level=4, N_iters=2)
        # GN_update ends in `if torch.isnan(theta_new):` (models_ford.py:594), which only evaluates for a batch of one
        kat_loop(rk, rf, "kat10_gn_ford", "ford", planted_inputs("ford", 104, GT2[:1], A=512, side_m=512 * 0.22, l2=True),
                 Optimizer="GN", N_iters=3)
        kat_loop(rk, rf, "kat10_polar_kitti", "kitti", planted_inputs("kitti", 105, GT2), proj="polar", N_iters=3)
        kat_loop(rk, rf, "kat10_polar_ford", "ford", planted_inputs("ford", 106, GT2, A=512, side_m=512 * 0.22), proj="nn",
                 N_iters=2)
        kat_loop(rk, rf, "kat10_polar_sgd", "kitti", rand_inputs(107), proj="polar", Optimizer="SGD", N_iters=2)
    if want("kat10nn"):
        kat_loop(rk, rf, "kat10_nn", "kitti", planted_inputs("kitti", 108, GT2, l2=True), Optimizer="NN", N_iters=2, tol=5e-6)
        kat_loop(rk, rf, "kat10_nn_level4_polar", "kitti", planted_inputs("kitti", 109, GT2[:1], L=4, l2=True), Optimizer="NN", level=4,
                 proj="polar", N_iters=1, tol=5e-6)
    if want("kat6"):   # forced out-of-range reset (models_kitti.py:1028-1033): start outside (-2.5, 2.5)
        p0 = torch.tensor([[3.0, 0.1, 0.2], [0.1, -2.8, -0.1]])
        kat_loop(rk, rf, "kat6_reset", "kitti", rand_inputs(61), N_iters=2,
                 pose0=(p0[:, 0:1], p0[:, 1:2], p0[:, 2:3]))

    if want("g2sp"):   # LM_G2SP: ground features warped to the satellite plane
        def g2sp_planted(seed, gt, using_conf=False):
            def f(oa):
                B = len(gt)
                sat, grd = O.planted_case("kitti", B, 512, 3, seed, gt, oa)
                g = torch.Generator().manual_seed(seed + 1)
                conf = [torch.sigmoid(-torch.sigmoid(torch.randn(B, 1, *x.shape[-2:], generator=g))) for x in grd]
                return sat, grd, conf, dict(seed=seed, B=B, A=512, L=3, gt=np.array(gt, dtype=np.float32))
            return f

        def g2sp_rand(seed, B=2):
            def f(oa):
                sat, grd, conf = O.random_pyramid(B, 512, 3, seed)
                return sat, grd, conf, dict(seed=seed, B=B, A=512, L=3)
            return f
        kat_g2sp(rk, "g2sp_planted", g2sp_planted(81, GT2), N_iters=3)
        kat_g2sp(rk, "g2sp_random", g2sp_rand(82), N_iters=2)
        kat_g2sp(rk, "g2sp_weight", g2sp_planted(83, GT2), N_iters=2, using_weight=1)

    if want("g2spnn"):  # LM_G2SP --proj nn: in-plane warp of square ground features (models_kitti.py:289-331)
        def nn_planted(seed, gt, A=512, L=3):
            def f(oa):
                B = len(gt)
                grd = [O.l2_norm(x) for x in O.smooth_pyramid(B, A, L, seed)]
                g = torch.as_tensor(gt, dtype=torch.float32).reshape(B, 3)
                sat = []
                for lv in range(L):
                    uv, *_ = O.g2sp_inplane_uv(grd[lv].shape[-1], g[:, 0:1], g[:, 1:2], g[:, 2:3], oa)
                    sat.append(O.bilinear_sample(grd[lv], uv)[0].contiguous())
                gg = torch.Generator().manual_seed(seed + 1)
                conf = [torch.sigmoid(-torch.sigmoid(torch.randn(B, 1, *x.shape[-2:], generator=gg))) for x in grd]
                return sat, grd, conf, dict(seed=seed, B=B, A=A, L=L, gt=np.array(gt, dtype=np.float32))
            return f
        kat_g2sp(rk, "g2sp_nn_planted", nn_planted(91, [[0.1, -0.08, 0.3], [-0.06, 0.12, -0.2]]), N_iters=3, proj="nn")
        kat_g2sp(rk, "g2sp_nn_weight", nn_planted(92, [[0.05, 0.1, -0.25]]), N_iters=2, proj="nn", using_weight=1)
        sd = O.vgg_state_dict(7)
        for level in (3, 4):           # VGGUnet_G2S (VGG.py:206-345) on a small image
            net = rv.VGGUnet_G2S(level)
            net.load_state_dict(sd)
            net.eval()
            x = torch.rand(2, 3, 64, 128, generator=torch.Generator().manual_seed(170 + level))
            with torch.no_grad():
                rfe, rco = net(x)
            ofe, oco = O.vgg_unet_g2s(sd, x, level)
            out = {"in_csum": csum(x)}
            for i in range(len(rfe)):
                close(ofe[i], rfe[i], 1e-6, "g2s feat %d" % i)
                close(oco[i], rco[i], 1e-6, "g2s conf %d" % i)
                out["feat%d" % i] = rfe[i].numpy()
                out["conf%d" % i] = rco[i].numpy()
            np.savez_compressed(os.path.join(GOLD, "kat7_vgg_g2s_level%d.npz" % level), **out)
        e2e_g2sp(rk, proj="nn", name="e2e_g2sp_nn")
        print("G2SP nn ok")

    if want("kat7"):   # VGG U-Net: small image, all intermediate activations
        sd = O.vgg_state_dict(7)
        for level in (3, 4):
            net = rv.VGGUnet(level)
            net.load_state_dict(sd)
            net.eval()
            g = torch.Generator().manual_seed(70 + level)
            x = torch.rand(2, 3, 64, 128, generator=g)
            with torch.no_grad():
                rfe, rco = net(x)
            ofe, oco = O.vgg_unet(sd, x, level)
            out = {"in_csum": csum(x), "w_csum": csum(*[sd[k] for k in sorted(sd)])}
            for i in range(len(rfe)):
                close(ofe[i], rfe[i], 1e-6, "kat7 feat %d" % i)
                close(oco[i], rco[i], 1e-6, "kat7 conf %d" % i)
                out["feat%d" % i] = rfe[i].numpy()
                out["conf%d" % i] = rco[i].numpy()
            np.savez_compressed(os.path.join(GOLD, "kat7_vgg_level%d.npz" % level), **out)
        print("KAT-7 VGG ok")

    if want("e2eg2sp"):
        e2e_g2sp(rk)
    if want("e2ex"):   # round-2 end-to-end goldens: other level selections, config-3 shapes, 8 pairs
        e2e_more(rk, rf, "e2e_kitti_level_m1", "kitti", 2, 512, seed=2031, level=-1)
        e2e_more(rk, rf, "e2e_ford_level2", "ford", 2, 512, seed=2032, level=2)
        e2e_more(rk, rf, "e2e_ford1280", "ford", 2, 1280, seed=2033)
        e2e_more(rk, rf, "e2e_kitti8", "kitti", 8, 512, seed=2034)
    if want("b32"):    # config-2 sized planted-pose trajectory (B = 32) from the reference, in chunks
        planted_b32(rk)

    if want("e2e"):    # whole forward through the reference nn.Module (VGG + LM), KITTI + Ford
        sd = {}
        sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
        sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
        sd["damping"] = torch.zeros(1, 3)
        g = torch.Generator().manual_seed(2022)
        sat = torch.rand(2, 3, 512, 512, generator=g)
        grd = torch.rand(2, 3, 256, 1024, generator=g)
        for kind in ("kitti", "ford"):
            a = ref_args()
            net = (rk.LM_S2GP if kind == "kitti" else rf.LM_S2GP_Ford)(a)
            torch.autograd.set_detect_anomaly(False)
            net.load_state_dict(sd)
            net.eval()
            torch.manual_seed(999)
            with torch.no_grad():
                if kind == "kitti":
                    r = net(sat, grd, mode="test")
                    torch.manual_seed(999)
                    o = O.forward_kitti(sd, sat, grd, o_args(a))
                else:
                    fd = ford_dict(2, 512 * 0.22)
                    r = net(sat, grd, fd["side_m"], fd["R_FL"], fd["T_FL"], mode="test")
                    torch.manual_seed(999)
                    o = O.forward_ford(sd, sat, grd, fd["side_m"], fd["R_FL"], fd["T_FL"], o_args(a))
            r = torch.stack(r, dim=-1)
            of = torch.stack([o.lats[:, -1, -1], o.lons[:, -1, -1], o.thetas[:, -1, -1]], dim=-1)
            d = close(of, r, 1e-5, "e2e " + kind)
            np.savez_compressed(os.path.join(GOLD, "e2e_%s.npz" % kind), final=r.numpy(),
                                lats=o.lats.numpy(), lons=o.lons.numpy(), thetas=o.thetas.numpy(),
                                in_csum=csum(sat, grd))
            print("e2e %s ok (max|d| %.2e) final=%s" % (kind, d, r.tolist()))


def e2e_more(rk, rf, name, kind, B, A, seed, **akw):
    """Whole forward(mode='test') of the UNMODIFIED reference module on B seeded pairs (satellite side A), with the oracle
    run next to it on the same CPU-RNG state; stores the reference's final poses and the oracle's trajectories."""
    sd = {}
    sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
    sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
    sd["damping"] = torch.zeros(1, 3)
    g = torch.Generator().manual_seed(seed)
    sat = torch.rand(B, 3, A, A, generator=g)
    grd = torch.rand(B, 3, 256, 1024, generator=g)
    a = ref_args(**akw)
    net = (rk.LM_S2GP if kind == "kitti" else rf.LM_S2GP_Ford)(a)
    torch.autograd.set_detect_anomaly(False)
    net.load_state_dict(sd)
    net.eval()
    outs, lats, lons, ths, t64 = [], [], [], [], []
    sd64 = {k: v.double() for k, v in sd.items()}
    chunk = 2                                  # bounds the reference's [3,B,C,H,W] Jacobian tensors
    torch.manual_seed(999)
    state = torch.get_rng_state()
    for b0 in range(0, B, chunk):
        sl = slice(b0, b0 + chunk)
        with torch.no_grad():
            torch.set_rng_state(state)         # every chunk sees the draws a 2-pair batch would (no reset fires: checked below)
            if kind == "kitti":
                r = net(sat[sl], grd[sl], mode="test")
                torch.set_rng_state(state)
                o = O.forward_kitti(sd, sat[sl], grd[sl], o_args(a))
            else:
                fd = ford_dict(chunk, A * 0.22)
                r = net(sat[sl], grd[sl], fd["side_m"], fd["R_FL"], fd["T_FL"], mode="test")
                torch.set_rng_state(state)
                o = O.forward_ford(sd, sat[sl], grd[sl], fd["side_m"], fd["R_FL"], fd["T_FL"], o_args(a))
            # the same algorithm end to end in float64: how far the fp32 reference itself is from the truth on this
            # (non-contractive, random-weight) input is the noise floor any fp32 implementation is held to
            torch.set_rng_state(state)
            if kind == "kitti":
                o64 = O.forward_kitti(sd64, sat[sl].double(), grd[sl].double(), o_args(a))
            else:
                o64 = O.forward_ford(sd64, sat[sl].double(), grd[sl].double(), fd["side_m"], fd["R_FL"].double(), fd["T_FL"].double(), o_args(a))
        outs.append(torch.stack(r, dim=-1))
        lats.append(o.lats); lons.append(o.lons); ths.append(o.thetas)
        t64.append(torch.stack([o64.lats, o64.lons, o64.thetas], dim=-1))
    r = torch.cat(outs)
    lats, lons, ths = torch.cat(lats), torch.cat(lons), torch.cat(ths)
    of = torch.stack([lats[:, -1, -1], lons[:, -1, -1], ths[:, -1, -1]], dim=-1)
    d = close(of, r, 1e-5, name)
    assert float(torch.stack([lats, lons]).abs().max()) < 2.4, "a reset fired: the chunked RNG replay is not valid for this case"
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), final=r.numpy(), lats=lats.numpy(), lons=lons.numpy(),
                        thetas=ths.numpy(), traj64=torch.cat(t64).numpy(), in_csum=csum(sat, grd), seed=seed, B=B, A=A)
    n64 = (torch.stack([lats, lons, ths], dim=-1) - torch.cat(t64)).abs().amax(dim=(1, 2, 3))
    print("%s ok (max|d| %.2e) final[0]=%s; reference fp32 vs fp64 per pair: %s" % (name, d, r[0].tolist(), ["%.1e" % v for v in n64.tolist()]))


def planted_b32(rk):
    """BASELINE config-2 size on contractive inputs: 32 planted-pose KITTI pairs through the reference's own
    project_map_to_grd + LM_update loop (chunks of 4 bound the [3,B,C,H,W] Jacobians) and the float64 oracle."""
    B, seed = 32, 77
    a = ref_args()
    oa = o_args(a)
    net = rk.LM_S2GP(a)
    torch.autograd.set_detect_anomaly(False)
    gen = torch.Generator().manual_seed(5)
    gt = (torch.rand(B, 3, generator=gen) - 0.5) * 0.8
    sat, grd = O.planted_case("kitti", B, 512, 3, seed, gt, oa)
    traj, traj64 = [], []
    for b0 in range(0, B, 4):
        s4, g4 = [x[b0:b0 + 4] for x in sat], [x[b0:b0 + 4] for x in grd]
        c4 = [torch.ones(4, 1, *x.shape[-2:]) for x in g4]
        torch.manual_seed(4242)
        with torch.no_grad():
            t, _ = run_ref_loop(net, "kitti", s4, g4, c4, a)
        torch.manual_seed(4242)
        r64 = O.lm_loop("kitti", [x.double() for x in s4], [x.double() for x in g4], [x.double() for x in c4], oa)
        traj.append(t)
        traj64.append(torch.stack([r64.lons, r64.lats, r64.thetas], dim=-1))
        print("  planted_b32 chunk %d: |final - gt| %.2e" % (b0 // 4, float((t[:, -1, -1] - gt[b0:b0 + 4]).abs().max())), flush=True)
    traj, traj64 = torch.cat(traj), torch.cat(traj64)
    assert float(traj[..., :2].abs().max()) < 2.4, "a reset fired"
    np.savez_compressed(os.path.join(GOLD, "kat4_planted_b32.npz"), traj=traj.numpy(), traj64=traj64.numpy(), gt=gt.numpy(),
                        in_csum=csum(*sat, *grd), seed=seed, B=B)
    print("kat4_planted_b32 ok: max |final - gt| %.2e" % float((traj[:, -1, -1] - gt).abs().max()))


def e2e_g2sp(rk, proj="geo", name="e2e_g2sp"):
    """Whole LM_G2SP.forward (VGG + LM) through the reference nn.Module on CPU (Tensor.cuda patched)."""
    sd = {}
    sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
    sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
    g = torch.Generator().manual_seed(2022)
    sat = torch.rand(2, 3, 512, 512, generator=g)
    grd = torch.rand(2, 3, 256, 1024, generator=g)
    cam_k = torch.tensor([O._KITTI_K], dtype=torch.float32).repeat(2, 1, 1)
    a = ref_args(proj=proj)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *x, **k: self
    try:
        net = rk.LM_G2SP(a)
        torch.autograd.set_detect_anomaly(False)
        sd["damping"] = net.damping.detach().clone()
        net.load_state_dict(sd)
        net.eval()
        with torch.no_grad():
            r = torch.stack(net(sat, grd, cam_k, mode="test"), dim=-1)
    finally:
        torch.Tensor.cuda = orig_cuda
    oa = o_args(a)
    sf, _ = O.vgg_unet(sd, sat, 3, "SatFeatureNet.")
    gf, gc = (O.vgg_unet_g2s if proj == "nn" else O.vgg_unet)(sd, grd, 3, "GrdFeatureNet.")
    res = O.lm_loop_g2sp(sf, gf, gc, cam_k, oa)
    of = torch.stack([res.lats[:, -1, -1], res.lons[:, -1, -1], res.thetas[:, -1, -1]], dim=-1)
    d = close(of, r, 1e-5, "e2e g2sp")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), final=r.numpy(), lats=res.lats.numpy(), lons=res.lons.numpy(),
                        thetas=res.thetas.numpy(), in_csum=csum(sat, grd), cam_k=cam_k.numpy())
    print("e2e g2sp ok (max|d| %.2e) final=%s" % (d, r.tolist()))


if __name__ == "__main__":
    main()
