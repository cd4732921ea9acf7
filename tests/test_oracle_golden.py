"""CPU: the oracle restatement reproduces the REFERENCE outputs stored in tests/golden/.
(The fixtures were produced by oracle/make_golden.py from the unmodified reference.)"""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tests import cases as K


def test_kat1_sampler():
    g = K.load_golden("kat1_sampler")
    img = torch.from_numpy(g["img"])
    for name in ("rand", "edge"):
        out, jac = O.bilinear_sample(img, torch.from_numpy(g[name + "_uv"]), torch.from_numpy(g[name + "_jacin"]))
        np.testing.assert_array_equal(out.numpy(), g[name + "_out"])
        np.testing.assert_array_equal(jac.numpy(), g[name + "_jac"])
    # the reference's own commented check (jacobian.py:223-225): == F.grid_sample(align_corners=True)
    grid = torch.from_numpy(g["rand_uv"]) / 31 * 2 - 1
    f = torch.nn.functional.grid_sample(img, grid, align_corners=True)
    np.testing.assert_allclose(g["rand_out"], f.numpy(), atol=2e-6)


def test_kat1_sampler_jacobian_matches_autograd():
    g = K.load_golden("kat1_sampler")
    img = torch.from_numpy(g["img"]).double()
    uv = torch.from_numpy(g["rand_uv"]).double()[:, :4, :4].clone().requires_grad_(True)
    eye = torch.zeros(2, 1, 4, 4, 2, dtype=torch.float64)
    eye[0, ..., 0] = 1
    eye[1, ..., 1] = 1
    out, jac = O.bilinear_sample(img, uv, eye)
    for c in range(3):
        gr, = torch.autograd.grad(out[0, c].sum(), uv, retain_graph=True)
        np.testing.assert_allclose(jac[0, 0, c].detach().numpy(), gr[0, ..., 0].numpy(), atol=1e-12)
        np.testing.assert_allclose(jac[1, 0, c].detach().numpy(), gr[0, ..., 1].numpy(), atol=1e-12)


@pytest.mark.parametrize("kind", ["kitti", "ford"])
def test_kat2_geometry(kind):
    g = K.load_golden("kat2_geometry")
    args = O.LMArgs(level=4, shift_range_lat=20.0, shift_range_lon=15.0)
    pose = torch.from_numpy(g["poses"])
    su, sv, th = pose[:, 0:1], pose[:, 1:2], pose[:, 2:3]
    for lv in range(4):
        if kind == "kitti":
            tab = O.kitti_ground_table(lv)
            o = O.kitti_sat_uv(tab[0], tab[1], su, sv, th, 512 // 2 ** (3 - lv), args)
        else:
            tab = O.ford_ground_table(lv)
            o = O.ford_sat_uv(tab[0], tab[1], torch.from_numpy(g["R_FL"]), torch.from_numpy(g["T_FL"]), su, sv, th,
                              1280 // 2 ** (3 - lv), 1280 * 0.22, args)
        np.testing.assert_array_equal(tab[0][::4, ::8].numpy(), g["%s_L%d_tab" % (kind, lv)])
        for i, nm in enumerate(["uv", "mask", "ju", "jv", "jt"]):
            np.testing.assert_array_equal(o[i][:, ::4, ::8].numpy(), g["%s_L%d_%s" % (kind, lv, nm)])


@pytest.mark.parametrize("name", K.CPU_LOOP_CASES)
def test_lm_loop_matches_reference(name):
    c = K.build_loop_case(name)
    torch.manual_seed(K.RESET_SEED)
    res = O.lm_loop(c["kind"], c["sat"], c["grd"], c["conf"], c["args"], c["damping_param"], None, c["ford"], c["pose0"],
                    nn_sd=c["nn_sd"])
    if c["kind"] == "kitti":
        traj = torch.stack([res.lons, res.lats, res.thetas], dim=-1)
    else:
        traj = torch.stack([res.lats, res.lons, res.thetas], dim=-1)
    np.testing.assert_allclose(traj.numpy(), c["gold"]["traj"], rtol=0, atol=5e-6 if c["nn_sd"] else 2e-6)
    if "gt" in c["gold"].files and name not in ("kat5_ford1280",) and (not name.startswith("kat10") or name == "kat10_gn_ford"):
        # planted pose: the reference itself converges onto gt (contractive input)
        np.testing.assert_allclose(c["gold"]["traj"][:, -1, -1], c["gold"]["gt"], atol=5e-5)


@pytest.mark.parametrize("level", [3, 4])
def test_kat7_vgg(level):
    g = K.load_golden("kat7_vgg_level%d" % level)
    sd = O.vgg_state_dict(7)
    x = torch.rand(2, 3, 64, 128, generator=torch.Generator().manual_seed(70 + level))
    np.testing.assert_allclose(K.csum(x), g["in_csum"], rtol=1e-6)
    feats, confs = O.vgg_unet(sd, x, level)
    for i in range(len(feats)):
        np.testing.assert_allclose(feats[i].numpy(), g["feat%d" % i], atol=1e-6)
        np.testing.assert_allclose(confs[i].numpy(), g["conf%d" % i], atol=1e-6)


@pytest.mark.parametrize("level", [3, 4])
def test_kat7_vgg_g2s(level):
    """VGGUnet_G2S (VGG.py:206-345) restated: folded decoder maps, c0 from the un-folded x15."""
    g = K.load_golden("kat7_vgg_g2s_level%d" % level)
    x = torch.rand(2, 3, 64, 128, generator=torch.Generator().manual_seed(170 + level))
    np.testing.assert_allclose(K.csum(x), g["in_csum"], rtol=1e-6)
    feats, confs = O.vgg_unet_g2s(O.vgg_state_dict(7), x, level)
    for i in range(len(feats)):
        np.testing.assert_allclose(feats[i].numpy(), g["feat%d" % i], atol=1e-6)
        np.testing.assert_allclose(confs[i].numpy(), g["conf%d" % i], atol=1e-6)
    assert feats[0].shape[-2:] == (16, 8) and confs[0].shape[-2:] == (8, 16)


@pytest.mark.parametrize("name", ["g2sp_random", "g2sp_weight", "g2sp_nn_weight"])
def test_g2sp_loop_matches_reference(name):
    c = K.build_g2sp_case(name)
    res = O.lm_loop_g2sp(c["sat"], c["grd"], c["conf"], c["cam_k"], c["args"])
    traj = torch.stack([res.lons, res.lats, res.thetas], dim=-1)
    np.testing.assert_allclose(traj.numpy(), c["gold"]["traj"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("name,kind,akw", [("e2e_kitti_level_m1", "kitti", dict(level=-1)), ("e2e_ford_level2", "ford", dict(level=2))])
def test_oracle_whole_forward_at_other_level_selections(name, kind, akw):
    """VGG.py:192-203 level -1 ([x15]) and models_ford.py:59-65 level 2 ([x18, x21] with the /4 and /2 grids): the oracle's
    whole forward against the unmodified reference's test-mode output (tests/golden, oracle/make_golden.py e2e_more)."""
    g = K.load_golden(name)
    B, A = int(g["B"]), int(g["A"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    sat = torch.rand(B, 3, A, A, generator=gen)
    grd = torch.rand(B, 3, 256, 1024, generator=gen)
    np.testing.assert_allclose(K.csum(sat, grd), g["in_csum"], rtol=1e-6)
    sd = {}
    sd.update(O.vgg_state_dict(100, "SatFeatureNet."))
    sd.update(O.vgg_state_dict(101, "GrdFeatureNet."))
    sd["damping"] = torch.zeros(1, 3)
    a = O.LMArgs(**akw)
    torch.manual_seed(999)
    with torch.no_grad():
        if kind == "kitti":
            o = O.forward_kitti(sd, sat, grd, a)
        else:
            f = K.ford_dict(B, A * 0.22)
            o = O.forward_ford(sd, sat, grd, f["side_m"], f["R_FL"], f["T_FL"], a)
    got = torch.stack([o.lats[:, -1, -1], o.lons[:, -1, -1], o.thetas[:, -1, -1]], dim=-1).numpy()
    np.testing.assert_allclose(got, g["final"], rtol=0, atol=1e-5)
