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
