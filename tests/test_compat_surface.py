"""CPU: the materialising compatibility methods (`project_map_to_grd`, `LM_update`; SURVEY.md 8b "LM-step
signatures to keep") against the oracle's restatement of the same reference functions.  Tolerances: the closed-form
geometry differs from the reference's matrix chain by <= 1 ulp of |uv|, so warped features agree to
1e-4 relative and one dense LM update to 2e-5 absolute on the pose."""
import numpy as np
import pytest
import torch

from highlyaccurate_b200.models_ford import LM_S2GP_Ford
from highlyaccurate_b200.models_kitti import LM_S2GP
from oracle import oracle as O
from tests import cases as K


def _rand_case(B, C, A, level, seed):
    g = torch.Generator().manual_seed(seed)
    h, w = 256 >> (3 - level), 1024 >> (3 - level)
    sat = torch.randn(B, C, A, A, generator=g)
    conf = torch.rand(B, 1, A, A, generator=g)
    grd = torch.randn(B, C, h, w, generator=g)
    gconf = torch.rand(B, 1, h, w, generator=g)
    pose = (torch.rand(B, 3, generator=g) - 0.5) * 0.6
    return sat, conf, grd, gconf, pose[:, 0:1], pose[:, 1:2], pose[:, 2:3]


@pytest.mark.parametrize("kind", ["kitti", "ford"])
def test_project_map_to_grd_matches_oracle(kind):
    B, C, A, level = 2, 8, 64, 0
    sat, conf, grd, gconf, su, sv, th = _rand_case(B, C, A, level, 11)
    a = O.LMArgs()
    if kind == "kitti":
        net = LM_S2GP(K.args_from_lmargs(a))
        f, c, jac, uvm, mask = net.project_map_to_grd(sat, conf, su, sv, th, level)
        uv, m, ju, jv, jt = O.kitti_sat_uv(*O.kitti_ground_table(level), su, sv, th, A, a)
    else:
        ford = K.ford_dict(B, 0.22 * 512)
        net = LM_S2GP_Ford(K.args_from_lmargs(a))
        f, c, jac, uvm, mask = net.project_map_to_grd(sat, conf, ford["R_FL"], ford["T_FL"], su, sv, th, level, ford["side_m"])
        uv, m, ju, jv, jt = O.ford_sat_uv(*O.ford_ground_table(level), ford["R_FL"], ford["T_FL"], su, sv, th, A, ford["side_m"], a)
    want_f, want_j = O.bilinear_sample(sat, uv, torch.stack([ju, jv, jt], dim=0))
    want_c, _ = O.bilinear_sample(conf, uv)
    np.testing.assert_allclose(mask.numpy(), m.numpy())
    np.testing.assert_allclose(uvm.numpy(), (uv * m[..., None]).numpy(), rtol=1e-5, atol=2e-4)
    # compare where the sample point is safely inside a texel cell (a 1-ulp difference in uv can flip floor() elsewhere)
    frac = uv - torch.floor(uv)
    safe = ((frac > 1e-3) & (frac < 1 - 1e-3)).all(dim=-1)[:, None].expand_as(want_f)
    got, want = (f * safe).numpy(), (want_f * m[:, None] * safe).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose((c * safe[:, :1]).numpy(), (want_c * m[:, None] * safe[:, :1]).numpy(), rtol=1e-4, atol=2e-4)
    gj, wj = (jac * safe[None]).numpy(), (want_j * m[None, :, None] * safe[None]).numpy()
    assert np.abs(gj - wj).max() <= 1e-4 * np.abs(wj).max() + 1e-5
    assert jac.shape == (3, B, C, grd.shape[-2], grd.shape[-1])
    f2, _, j2, _, _ = (net.project_map_to_grd(sat, None, su, sv, th, level, require_jac=False) if kind == "kitti" else
                       net.project_map_to_grd(sat, None, ford["R_FL"], ford["T_FL"], su, sv, th, level, ford["side_m"], require_jac=False))
    assert j2 is None and torch.equal(f2, f)


@pytest.mark.parametrize("mode", ["full", "shift", "rot", "weight", "hessian", "traindamp"])
def test_lm_update_matches_oracle(mode):
    B, C, A, level = 2, 8, 64, 0
    sat, conf, grd, gconf, su, sv, th = _rand_case(B, C, A, level, 23)
    kw = {"shift": dict(rotation_range=0.0), "rot": dict(shift_range_lat=0.0, shift_range_lon=0.0), "weight": dict(using_weight=1),
          "hessian": dict(use_hessian=1), "traindamp": dict(train_damping=1)}.get(mode, {})
    a = O.LMArgs(**kw)
    net = LM_S2GP(K.args_from_lmargs(a))
    f, c, jac, uvm, mask = net.project_map_to_grd(sat, conf, su, sv, th, level)
    h2 = grd.shape[-2] // 2
    gm, gcm = grd * mask[:, None], gconf * mask[:, None]
    n = O.n_dof("kitti", a)
    lam = O.resolve_damping(a, net.damping.detach() if a.train_damping else None, n)
    torch.manual_seed(5)
    want = O.lm_update(su, sv, th, f[:, :, h2:], gm[:, :, h2:], gcm[:, :, h2:], jac[:, :, :, h2:], a, lam, O.draw_reset(B))
    torch.manual_seed(5)
    got = net.LM_update(su, sv, th, f[:, :, h2:], c[:, :, h2:], gm[:, :, h2:], gcm[:, :, h2:], jac[:, :, :, h2:])
    for g_, w_ in zip(got, want[:3]):
        np.testing.assert_allclose(g_.detach().numpy(), w_.numpy(), atol=2e-5, rtol=1e-4)


def test_lm_update_is_differentiable():
    B, C, A, level = 1, 4, 64, 0
    sat, conf, grd, gconf, su, sv, th = _rand_case(B, C, A, level, 31)
    net = LM_S2GP(K.args_from_lmargs(O.LMArgs()))
    sat.requires_grad_(True)
    f, c, jac, _, mask = net.project_map_to_grd(sat, conf, su, sv, th, level)
    out = net.LM_update(su, sv, th, f, c, grd * mask[:, None], gconf * mask[:, None], jac)
    sum(o.sum() for o in out).backward()
    assert sat.grad is not None and torch.isfinite(sat.grad).all() and float(sat.grad.abs().sum()) > 0


def test_train_gradients_match_reference_autograd():
    """KAT-8: the reference's train-mode loop (project_map_to_grd + LM_update chained without detach, loss_func method 0)
    differentiated by autograd, stored by oracle/make_golden.py.  The compatibility methods are differentiable closed
    forms of the same functions, so chaining them must reproduce the loss, the trajectory and the gradients w.r.t. both
    feature pyramids and the trained damping.  This is the executable specification for the fused backward."""
    from highlyaccurate_b200.models_kitti import loss_func
    gold = K.load_golden("kat8_train_grad")
    a = O.LMArgs(N_iters=2, train_damping=1)
    B, A, L = int(gold["B"]), int(gold["A"]), int(gold["L"])
    sat, grd = O.planted_case("kitti", B, A, L, int(gold["seed"]), gold["gt"], a)
    np.testing.assert_allclose(K.csum(*sat, *grd), gold["in_csum"], rtol=1e-6, err_msg="input regeneration drifted")
    sat = [s.clone().requires_grad_(True) for s in sat]
    grd = [g.clone().requires_grad_(True) for g in grd]
    net = LM_S2GP(K.args_from_lmargs(a))
    torch.manual_seed(4242)
    su, sv, th = torch.zeros(B, 1), torch.zeros(B, 1), torch.zeros(B, 1)
    rows = []
    for it in range(a.N_iters):
        row = []
        for lv in range(L):
            sp, _, dj, _, mask = net.project_map_to_grd(sat[lv], None, su, sv, th, lv)
            gf = grd[lv] * mask[:, None]
            gc = torch.ones(B, 1, *gf.shape[-2:]) * mask[:, None]
            h2 = gf.shape[-2] // 2
            su, sv, th = net.LM_update(su, sv, th, sp[:, :, h2:], gc[:, :, h2:], gf[:, :, h2:], gc[:, :, h2:], dj[:, :, :, h2:])
            row.append(torch.cat([su, sv, th], dim=1))
        rows.append(torch.stack(row, dim=1))
    traj = torch.stack(rows, dim=1)                                       # [B, N_iters, L, (su, sv, th)]
    np.testing.assert_allclose(traj.detach().numpy(), gold["traj"], atol=2e-5, rtol=1e-4)
    g = torch.from_numpy(gold["gt"])
    loss = loss_func(0, None, None, None, traj[..., 1], traj[..., 0], traj[..., 2], g[:, 1], g[:, 0], g[:, 2], None, None)[0]
    np.testing.assert_allclose(float(loss), float(gold["loss"]), rtol=1e-4)
    loss.backward()
    np.testing.assert_allclose(net.damping.grad.numpy(), gold["damping_grad"], rtol=2e-3, atol=1e-3)
    for name, ts in (("sat", sat), ("grd", grd)):
        for lv, t in enumerate(ts):
            gflat = t.grad.reshape(-1)
            want = gold["%s%d_val" % (name, lv)]
            got = gflat[torch.from_numpy(gold["%s%d_idx" % (name, lv)])].numpy()
            scale = np.abs(want).max()
            assert np.abs(got - want).max() <= 5e-3 * scale, "%s level %d: %g of %g" % (name, lv, np.abs(got - want).max(), scale)
            sums = np.array([float(gflat.double().sum()), float(gflat.double().abs().sum())])
            np.testing.assert_allclose(sums[1], gold["%s%d_sum" % (name, lv)][1], rtol=5e-3)


@pytest.mark.parametrize("gold_name,kw,level_first,gtol", [
    ("kat9_train_e2e", dict(N_iters=1), 0, 1e-4),                       # 3 steps: measured 2e-6 .. 6e-6 of the largest entry
    ("kat9_train_e2e_weighted_levelfirst", dict(N_iters=2, using_weight=1, train_damping=1), 1, 2e-3),    # 6 steps: 5e-4
    ("kat9_train_e2e_level_m1", dict(N_iters=2, level=-1), 0, 1e-3)])   # VGG.py:198-199: only x15, one pyramid level
def test_train_mode_forward_and_gradients_match_reference(gold_name, kw, level_first, gtol):
    """KAT-9: `LM_S2GP.forward(mode='train')` (the differentiable path used until the fused backward exists) against the
    reference's own forward + autograd on the same seeded weights and images: the 14-tuple's losses and the gradients
    into U-Net weights of both branches and `damping` (what train_kitti.py:354-365 consumes).  The second case covers
    the level-first order, confidence weighting and the trained damping."""
    from oracle.make_golden import E2E_TRAIN_PARAMS
    gold = K.load_golden(gold_name)
    net = LM_S2GP(K.ref_args(**kw))
    sd = {}
    sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
    sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
    sd["damping"] = torch.zeros(1, 3)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(2022)
    sat = torch.rand(1, 3, 512, 512, generator=g)
    grd = torch.rand(1, 3, 256, 1024, generator=g)
    gt = torch.from_numpy(gold["gt"])
    torch.manual_seed(4242)
    out = net(sat, grd, gt[:, 0:1], gt[:, 1:2], gt[:, 2:3], mode="train", level_first=level_first)
    assert len(out) == 14 and len(out[13]) == int(gold["n_conf"]) and out[13][0].shape == (1, 1, 32, 128)
    np.testing.assert_allclose(float(out[0].detach()), float(gold["loss"]), rtol=1e-5)
    for i, key in ((5, "loss_last"), (6, "lat_last"), (7, "lon_last"), (8, "theta_last")):
        np.testing.assert_allclose(out[i].detach().numpy(), gold[key], rtol=1e-4, atol=1e-5)
    # a difference of two ~100-scale losses (coe = 100): 2e-3 absolute = 2e-5 in pose units
    np.testing.assert_allclose(out[1].detach().numpy(), gold["loss_decrease"], rtol=1e-4, atol=2e-3)
    out[0].backward()
    if kw.get("train_damping"):
        np.testing.assert_allclose(net.damping.grad.numpy(), gold["damping_grad"], rtol=5e-3, atol=1e-4)
    params = dict(net.named_parameters())
    for k, name in enumerate(E2E_TRAIN_PARAMS):
        if "p%d_val" % k not in gold.files:            # the reference gave this weight no gradient (decoder unused at level -1)
            assert params[name].grad is None or float(params[name].grad.abs().max()) == 0.0, name
            continue
        gflat = params[name].grad.reshape(-1)
        want = gold["p%d_val" % k]
        got = gflat[torch.from_numpy(gold["p%d_idx" % k])].numpy()
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= gtol * scale, "%s: %g of %g" % (name, np.abs(got - want).max(), scale)
        np.testing.assert_allclose(float(gflat.double().abs().sum()), gold["p%d_sum" % k][1], rtol=10 * gtol)


@pytest.mark.parametrize("gold_name,kw,gtol", [("kat9_train_e2e_ford", dict(N_iters=1), 1e-4),
                                               ("kat9_train_e2e_ford_level2", dict(N_iters=2, level=2), 1e-3),    # models_ford.py:59-65
                                               # rotation_range == 0: coe_heading is forced to 0 (models_ford.py:843-846)
                                               ("kat9_train_e2e_ford_rot0", dict(N_iters=1, rotation_range=0.0), 1e-4)])
def test_ford_train_mode_forward_and_gradients_match_reference(gold_name, kw, gtol):
    """KAT-9 (Ford): `LM_S2GP_Ford.forward(mode='train')` against the reference's forward + autograd (train_ford.py:229-240),
    at the default level 3 and at level 2 ([x18, x21] with the /4 and /2 ground grids)."""
    from oracle.make_golden import E2E_TRAIN_PARAMS
    gold = K.load_golden(gold_name)
    net = LM_S2GP_Ford(K.ref_args(**kw))
    sd = {}
    sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
    sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
    sd["damping"] = torch.zeros(1, 3)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(2023)
    sat = torch.rand(1, 3, 512, 512, generator=g)
    grd = torch.rand(1, 3, 256, 1024, generator=g)
    fd = K.ford_dict(1, float(gold["side_m"]))
    gt = torch.from_numpy(gold["gt"])
    torch.manual_seed(4242)
    out = net(sat, grd, fd["side_m"], fd["R_FL"], fd["T_FL"], gt[:, 0], gt[:, 1], gt[:, 2], mode="train")
    assert len(out) == 14 and len(out[13]) == int(gold["n_conf"])
    np.testing.assert_allclose(float(out[0].detach()), float(gold["loss"]), rtol=1e-5)
    for i, key in ((5, "loss_last"), (6, "lat_last"), (7, "lon_last"), (8, "theta_last")):
        np.testing.assert_allclose(out[i].detach().numpy(), gold[key], rtol=1e-4, atol=1e-5)
    out[0].backward()
    params = dict(net.named_parameters())
    for k, name in enumerate(E2E_TRAIN_PARAMS):
        if "p%d_val" % k not in gold.files:
            continue
        gflat = params[name].grad.reshape(-1)
        want = gold["p%d_val" % k]
        got = gflat[torch.from_numpy(gold["p%d_idx" % k])].numpy()
        assert np.abs(got - want).max() <= gtol * np.abs(want).max(), "%s: %g of %g" % (name, np.abs(got - want).max(), np.abs(want).max())


def test_g2sp_compat_step_matches_oracle():
    """LM_G2SP.project_grd_to_map + LM_update (materialising compatibility methods) against the oracle's restatement of
    models_kitti.py:163-287 / :333-379 on random features."""
    from highlyaccurate_b200.models_kitti import LM_G2SP
    B, C, A = 2, 8, 64
    g = torch.Generator().manual_seed(41)
    sf, gf = torch.randn(B, C, A, A, generator=g), torch.randn(B, C, 32, 128, generator=g)
    gc = torch.rand(B, 1, 32, 128, generator=g)
    pose = (torch.rand(B, 3, generator=g) - 0.5) * 0.4
    su, sv, th = pose[:, 0:1], pose[:, 1:2], pose[:, 2:3]
    cam_k = torch.tensor([O._KITTI_K], dtype=torch.float32).repeat(B, 1, 1)
    for uw in (0, 1):
        a = O.LMArgs(using_weight=uw)
        net = LM_G2SP(K.args_from_lmargs(a))
        gp, gcp, dj = net.project_grd_to_map(gf, gc, su, sv, th, cam_k, A, 256, 1024)
        got = net.LM_update(su, sv, th, gp, gcp, sf, None, dj)
        want = O.g2sp_one_step(sf, gf, gc, cam_k, su, sv, th, 256, 1024, a, a.damping * torch.ones(1, 3))
        for x, y in zip(got, want[:3]):
            np.testing.assert_allclose(x.detach().numpy(), y.numpy(), atol=2e-5, rtol=1e-4)


def test_g2sp_train_mode_forward_and_gradients_match_reference():
    """KAT-9 (G2SP): `LM_G2SP.forward(mode='train')` against the reference's forward + autograd (train_kitti.py:363-365)."""
    from highlyaccurate_b200.models_kitti import LM_G2SP
    from oracle.make_golden import E2E_TRAIN_PARAMS
    gold = K.load_golden("kat9_train_e2e_g2sp")
    net = LM_G2SP(K.ref_args(N_iters=1))
    sd = {}
    sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
    sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
    sd["damping"] = net.damping.detach().clone()
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(2024)
    sat = torch.rand(1, 3, 512, 512, generator=g)
    grd = torch.rand(1, 3, 256, 1024, generator=g)
    gt = torch.from_numpy(gold["gt"])
    out = net(sat, grd, torch.from_numpy(gold["cam_k"]), gt[:, 0:1], gt[:, 1:2], gt[:, 2:3], mode="train")
    assert len(out) == 14
    np.testing.assert_allclose(float(out[0].detach()), float(gold["loss"]), rtol=1e-5)
    for i, key in ((5, "loss_last"), (6, "lat_last"), (7, "lon_last"), (8, "theta_last")):
        np.testing.assert_allclose(out[i].detach().numpy(), gold[key], rtol=1e-4, atol=1e-5)
    out[0].backward()
    params = dict(net.named_parameters())
    for k, name in enumerate(E2E_TRAIN_PARAMS):
        gflat = params[name].grad.reshape(-1)
        want = gold["p%d_val" % k]
        got = gflat[torch.from_numpy(gold["p%d_idx" % k])].numpy()
        # measured 4e-3 .. 3e-2 of the largest entry.  One step on given features agrees to 2e-5 (checked when this test was
        # written); end to end, the perspective projection puts uv at ~1e3 pixels where a 1-ulp difference between the
        # closed-form projection and the reference's matrix chain flips floor() for a few pixels, and the bilinear
        # *gradient* is discontinuous across texel boundaries.  The S2GP paths (uv <= 512) agree to 6e-6.
        assert np.abs(got - want).max() <= 5e-2 * np.abs(want).max(), "%s: %g of %g" % (name, np.abs(got - want).max(), np.abs(want).max())
