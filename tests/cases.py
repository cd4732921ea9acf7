"""Shared definitions of the known-answer cases (inputs are regenerated from seeds; the
expected outputs are the REFERENCE's, stored in tests/golden/*.npz by oracle/make_golden.py)."""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORD_R = [[0., 0., 1.], [1., 0., 0.], [0., 1., 0.]]
FORD_T = [1.7, -0.3, -1.5]
GT2 = [[0.3, -0.25, 0.5], [-0.2, 0.4, -0.3]]

# name -> (kind, input family, LMArgs overrides, extra)
LOOP_CASES = {
    "kat3_random_kitti": ("kitti", "rand", {}, {}),
    "kat4_planted_kitti": ("kitti", "planted", {}, {}),
    "kat4_planted_ford": ("ford", "planted", {}, {}),
    "kat5_weight": ("kitti", "rand", dict(using_weight=1), {}),
    "kat5_hessian": ("kitti", "rand", dict(use_hessian=1), {}),
    "kat5_traindamp": ("kitti", "planted", dict(train_damping=1), dict(damping_param=torch.zeros(1, 3))),
    "kat5_levelfirst": ("kitti", "planted", dict(level_first=1), {}),
    "kat5_shiftonly": ("kitti", "planted", dict(rotation_range=0.0), {}),
    "kat5_rotonly": ("kitti", "planted", dict(shift_range_lat=0.0, shift_range_lon=0.0), {}),
    "kat5_level4": ("kitti", "planted", dict(level=4, N_iters=2), {}),
    "kat5_ford1280": ("ford", "planted", dict(N_iters=3), {}),
    "kat5_anisotropic": ("kitti", "planted", dict(shift_range_lat=20.0, shift_range_lon=12.0, rotation_range=15.0), {}),
    "kat6_reset": ("kitti", "rand", dict(N_iters=2), dict(pose0_from_golden=True)),
    # the optimiser ablations (models_kitti.py:1056-1124, models_ford.py:534-598) and the polar ground
    # table of every proj != 'geo' (models_kitti.py:684-698); "planted_l2" = planted pose on L2-normalised pyramids
    "kat10_sgd": ("kitti", "planted_l2", dict(Optimizer="SGD", N_iters=3), {}),
    "kat10_adam": ("kitti", "planted_l2", dict(Optimizer="ADAM", N_iters=3), {}),
    "kat10_adam_level4": ("kitti", "planted_l2", dict(Optimizer="ADAM", level=4, N_iters=2), {}),
    "kat10_gn_ford": ("ford", "planted_l2", dict(Optimizer="GN", N_iters=3), {}),
    "kat10_polar_kitti": ("kitti", "planted", dict(proj="polar", N_iters=3), {}),
    "kat10_polar_ford": ("ford", "planted", dict(proj="nn", N_iters=2), {}),
    "kat10_polar_sgd": ("kitti", "rand", dict(proj="polar", Optimizer="SGD", N_iters=2), {}),
    # Optimizer 'NN' (RNNs.NNrefine with seeded weights, O.nnrefine_state_dict(seed + 1000))
    "kat10_nn": ("kitti", "planted_l2", dict(Optimizer="NN", N_iters=2), dict(nn=True)),
    "kat10_nn_level4_polar": ("kitti", "planted_l2", dict(Optimizer="NN", level=4, proj="polar", N_iters=1), dict(nn=True)),
}
# the cases cheap enough for the CPU suite (the rest are exercised by the gpu parity tests)
CPU_LOOP_CASES = ["kat3_random_kitti", "kat4_planted_kitti", "kat4_planted_ford", "kat5_weight", "kat5_shiftonly",
                  "kat5_rotonly", "kat5_level4", "kat6_reset", "kat10_sgd", "kat10_adam_level4", "kat10_gn_ford",
                  "kat10_polar_ford", "kat10_nn_level4_polar"]


def csum(*ts) -> np.ndarray:
    return np.array([float(t.double().sum()) for t in ts] + [float(t.double().abs().sum()) for t in ts])


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def ford_dict(B, side_m):
    return dict(R_FL=torch.tensor(FORD_R)[None].repeat(B, 1, 1), T_FL=torch.tensor(FORD_T)[None].repeat(B, 1),
                side_m=float(side_m))


def build_loop_case(name):
    """Returns dict(kind, args, sat, grd, conf, ford, damping_param, pose0, gold)."""
    kind, fam, akw, extra = LOOP_CASES[name]
    gold = load_golden(name)
    args = O.LMArgs(**akw)
    seed, B, A, L = int(gold["seed"]), int(gold["B"]), int(gold["A"]), int(gold["L"])
    ford = None
    if fam == "rand":
        sat, grd, conf = O.random_pyramid(B, A, L, seed)
    else:
        gt = gold["gt"]
        ford = ford_dict(B, float(gold["side_m"])) if kind == "ford" else None
        sat, grd = O.planted_case(kind, B, A, L, seed, gt, args, ford, l2=(fam == "planted_l2"))
        conf = [torch.ones(B, 1, *g.shape[-2:]) for g in grd]
    np.testing.assert_allclose(csum(*sat, *grd), gold["in_csum"], rtol=1e-6, err_msg="input regeneration drifted")
    pose0 = None
    if extra.get("pose0_from_golden"):
        p0 = torch.from_numpy(gold["pose_in"][:, 0, 0])
        pose0 = (p0[:, 0:1].clone(), p0[:, 1:2].clone(), p0[:, 2:3].clone())
    nn_sd = O.nnrefine_state_dict(seed + 1000) if extra.get("nn") else None
    return dict(kind=kind, args=args, sat=sat, grd=grd, conf=conf, ford=ford,
                damping_param=extra.get("damping_param"), pose0=pose0, gold=gold, B=B, A=A, L=L, nn_sd=nn_sd)


RESET_SEED = 4242    # oracle/make_golden.py seeds the CPU generator with this before every loop


def ref_args(**kw):
    """The argparse Namespace of train_kitti.py:426-485 / train_ford.py:343-412 at its defaults."""
    import types
    d = dict(level=3, N_iters=5, using_weight=0, loss_method=0, rotation_range=10.0, proj="geo", Optimizer="LM",
             damping=0.1, train_damping=0, shift_range_lat=20.0, shift_range_lon=20.0, use_hessian=0, dropout=0,
             use_gt_depth=0, visualize=0, coe_shift_lat=100.0, coe_shift_lon=100.0, coe_heading=100.0,
             coe_L1=100.0, coe_L2=100.0, coe_L3=100.0, coe_L4=100.0, estimate_depth=0, beta1=0.9, beta2=0.999)
    d.update(kw)
    return types.SimpleNamespace(**d)


def args_from_lmargs(a: "O.LMArgs"):
    return ref_args(level=a.level, N_iters=a.N_iters, using_weight=a.using_weight, damping=a.damping,
                    train_damping=a.train_damping, rotation_range=a.rotation_range, shift_range_lat=a.shift_range_lat,
                    shift_range_lon=a.shift_range_lon, use_hessian=a.use_hessian, Optimizer=a.Optimizer, proj=a.proj,
                    beta1=a.beta1, beta2=a.beta2)


G2SP_CASES = {"g2sp_planted": ("planted", dict(N_iters=3)), "g2sp_random": ("rand", dict(N_iters=2)),
              "g2sp_weight": ("planted", dict(N_iters=2, using_weight=1)),
              # LM_G2SP --proj nn (models_kitti.py:289-331): square ground features, in-plane warp
              "g2sp_nn_planted": ("nn_planted", dict(N_iters=3, proj="nn")),
              "g2sp_nn_weight": ("nn_planted", dict(N_iters=2, proj="nn", using_weight=1))}


def build_g2sp_case(name):
    fam, akw = G2SP_CASES[name]
    gold = load_golden(name)
    args = O.LMArgs(**akw)
    seed, B, A, L = int(gold["seed"]), int(gold["B"]), int(gold["A"]), int(gold["L"])
    if fam == "rand":
        sat, grd, conf = O.random_pyramid(B, A, L, seed)
    elif fam == "nn_planted":          # square L2-normalised ground pyramids; the satellite side is their in-plane warp at gt
        grd = [O.l2_norm(x) for x in O.smooth_pyramid(B, A, L, seed)]
        gt = torch.as_tensor(gold["gt"], dtype=torch.float32).reshape(B, 3)
        sat = []
        for lv in range(L):
            uv, *_ = O.g2sp_inplane_uv(grd[lv].shape[-1], gt[:, 0:1], gt[:, 1:2], gt[:, 2:3], args)
            sat.append(O.bilinear_sample(grd[lv], uv)[0].contiguous())
        g = torch.Generator().manual_seed(seed + 1)
        conf = [torch.sigmoid(-torch.sigmoid(torch.randn(B, 1, *x.shape[-2:], generator=g))) for x in grd]
    else:
        sat, grd = O.planted_case("kitti", B, A, L, seed, gold["gt"], args)
        g = torch.Generator().manual_seed(seed + 1)
        conf = [torch.sigmoid(-torch.sigmoid(torch.randn(B, 1, *x.shape[-2:], generator=g))) for x in grd]
    np.testing.assert_allclose(csum(*sat, *grd), gold["in_csum"], rtol=1e-6, err_msg="input regeneration drifted")
    return dict(args=args, sat=sat, grd=grd, conf=conf, cam_k=torch.from_numpy(gold["cam_k"]), gold=gold, B=B, A=A, L=L)
