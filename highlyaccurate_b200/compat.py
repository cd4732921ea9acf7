"""Materialising compatibility surface: `project_map_to_grd` / `LM_update` with the reference's signatures.

The accelerated `forward()` never calls these — the fused engine does not materialise the warped
features or `dfeat_dpose` (SURVEY.md section 8b, row "LM-step signatures to keep").  They exist for
callers of the reference that use the two methods directly (visualisation, ablation scripts): same
arguments, same returned tensors, differentiable, written as closed-form tensor algebra on the
caller's device.  They are not a fallback for the engine: the model constructors still require
libha_b200.so, and `forward()` raises on CPU tensors.

Geometry (closed forms of models_kitti.py:700-801 / models_ford.py:173-264, derived in SURVEY.md 8a):
  KITTI  u = ( s (X - sv_m) + c (Z + su_m)) / mpp + A/2,   v = (c (X - sv_m) - s (Z + su_m)) / mpp + A/2
  Ford   u = (-s X' + c Y') / mpp + A//2,                  v = -(c X' + s Y') / mpp + A//2,
         X' = xb + sv range_lon,  Y' = yb - su range_lat,  (xb, yb) = (R_FL p + T_FL)[:2]
  both   d(u,v)/dtheta = (rot pi/180) (v - centre, -(u - centre))
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def _pose_cols(*xs):
    return [x.reshape(-1, 1, 1) for x in xs]


def sat_uv_kitti(table: torch.Tensor, su, sv, th, A: int, rot: float, lat: float, lon: float, mpp: float):
    """table [H,W,4] = (x, y, z, mask) -> uv [B,H,W,2], mask [B,H,W], jac [3,B,H,W,2] (d/dsu, d/dsv, d/dtheta)."""
    su, sv, th = _pose_cols(su, sv, th)
    X, Z, m = table[None, ..., 0], table[None, ..., 2], table[None, ..., 3]
    k = rot * math.pi / 180.0
    c, s = torch.cos(th * k), torch.sin(th * k)
    dx, dz = X - sv * lat, Z + su * lon
    half = A / 2
    u = (s * dx + c * dz) / mpp + half
    v = (c * dx - s * dz) / mpp + half
    ones = torch.ones_like(u)
    j_su = torch.stack([c * lon / mpp * ones, -s * lon / mpp * ones], dim=-1)
    j_sv = torch.stack([-s * lat / mpp * ones, -c * lat / mpp * ones], dim=-1)
    j_th = torch.stack([k * (v - half), -k * (u - half)], dim=-1)
    B = u.shape[0]
    return torch.stack([u, v], dim=-1), m.expand(B, -1, -1), torch.stack([j_su, j_sv, j_th], dim=0)


def sat_uv_ford(table: torch.Tensor, R_FL, T_FL, su, sv, th, A: int, rot: float, lat: float, lon: float, mpp: float):
    """As sat_uv_kitti for the Ford front-left camera: camera -> body (R_FL, T_FL) -> world yaw -> satellite."""
    su, sv, th = _pose_cols(su, sv, th)
    p = table[..., :3]
    body = torch.einsum("bij,hwj->bhwi", R_FL.to(p.dtype), p) + T_FL.to(p.dtype)[:, None, None, :]
    m = table[None, ..., 3]
    k = rot * math.pi / 180.0
    c, s = torch.cos(th * k), torch.sin(th * k)
    Xp, Yp = body[..., 0] + sv * lon, body[..., 1] - su * lat
    half = float(A // 2)
    u = (-s * Xp + c * Yp) / mpp + half
    v = -(c * Xp + s * Yp) / mpp + half
    ones = torch.ones_like(u)
    j_su = torch.stack([-c * lat / mpp * ones, s * lat / mpp * ones], dim=-1)
    j_sv = torch.stack([-s * lon / mpp * ones, -c * lon / mpp * ones], dim=-1)
    j_th = torch.stack([k * (v - half), -k * (u - half)], dim=-1)
    B = u.shape[0]
    return torch.stack([u, v], dim=-1), m.expand(B, -1, -1), torch.stack([j_su, j_sv, j_th], dim=0)


def sample_with_jacobian(image: torch.Tensor, uv: torch.Tensor, jac: Optional[torch.Tensor] = None):
    """jacobian.py:138-205 semantics: pixel-unit coordinates, four independently clamped corners, inclusive range
    mask, weights from the clamped corners.  image [B,C,IH,IW], uv [B,H,W,2], jac [N,B,H,W,2] or None
    -> out [B,C,H,W], new_jac [N,B,C,H,W] or None."""
    B, C, IH, IW = image.shape
    x, y = uv[..., 0], uv[..., 1]
    x0, y0 = torch.floor(x), torch.floor(y)
    xw, xe = x0.clamp(0, IW - 1), (x0 + 1).clamp(0, IW - 1)
    yn, ys = y0.clamp(0, IH - 1), (y0 + 1).clamp(0, IH - 1)
    inside = ((x >= 0) & (x <= IW - 1) & (y >= 0) & (y <= IH - 1)).to(image.dtype)
    ex, wx, sy, ny = (xe - x) * inside, (x - xw) * inside, (ys - y) * inside, (y - yn) * inside
    nhwc = image.permute(0, 2, 3, 1)
    bi = torch.arange(B, device=image.device)[:, None, None]
    tap = lambda yy, xx: nhwc[bi, yy.long(), xx.long()]                       # [B,H,W,C]
    nw, ne, sw, se = tap(yn, xw), tap(yn, xe), tap(ys, xw), tap(ys, xe)
    e = lambda t: t[..., None]
    out = nw * e(ex * sy) + ne * e(wx * sy) + sw * e(ex * ny) + se * e(wx * ny)
    out = out.permute(0, 3, 1, 2)
    if jac is None:
        return out, None
    d_dx = (ne - nw) * e(sy * inside) + (se - sw) * e(ny * inside)           # weights already carry `inside` once:
    d_dy = (sw - nw) * e(ex * inside) + (se - ne) * e(wx * inside)           # inside is 0/1, so squaring it is harmless
    new_jac = d_dx[None] * jac[..., 0:1] + d_dy[None] * jac[..., 1:2]         # [N,B,H,W,C]
    return out, new_jac.permute(0, 1, 4, 2, 3)


def project_map_to_grd(uv, mask, jac, sat_f, sat_c, require_jac=True):
    """Tail of models_kitti.py:914-937: sample, multiply by the geometric mask, same for the confidence map."""
    f, nj = sample_with_jacobian(sat_f, uv, jac if require_jac else None)
    f = f * mask[:, None]
    if nj is not None:
        nj = nj * mask[None, :, None]
    c = None
    if sat_c is not None:
        c, _ = sample_with_jacobian(sat_c, uv)
        c = c * mask[:, None]
    return f, c, nj, uv * mask[..., None], mask


def lm_update_dense(su, sv, th, sat_proj, grd_feat, grd_conf, dfeat, damping: torch.Tensor, dof: str, using_weight: bool,
                    use_hessian: bool, redraw: bool):
    """models_kitti.py:939-1041 on materialised tensors.  dof: 'full' | 'shift' | 'rot'; `damping` [1,N] or scalar;
    `redraw`: apply the (-2.5, 2.5) shift reset with two draws from the CPU default generator (3-DOF only)."""
    if dof == "shift":
        dfeat = dfeat[:2]
    elif dof == "rot":
        dfeat = dfeat[2:]
    N, B = dfeat.shape[:2]
    C = sat_proj.shape[1]
    J = dfeat.reshape(N, B, -1)
    s = sat_proj.reshape(B, -1)
    g = grd_feat.reshape(B, -1)
    ns = torch.linalg.vector_norm(s, dim=-1).clamp_min(1e-6)
    ng = torch.linalg.vector_norm(g, dim=-1).clamp_min(1e-6)
    r = s / ns[:, None] - g / ng[:, None]
    Jn = (J / ns[None, :, None]).permute(1, 0, 2)                              # [B,N,D]
    if using_weight:
        w = grd_conf.reshape(B, 1, -1).expand(-1, C, -1).reshape(B, 1, -1)
        JW = Jn * w
    else:
        JW = Jn
    Hm = JW @ Jn.transpose(1, 2)
    eye = torch.eye(N, dtype=Hm.dtype, device=Hm.device)[None]
    M = torch.diag_embed(torch.diagonal(Hm, dim1=1, dim2=2)) if use_hessian else eye.expand(B, -1, -1)
    delta = -(torch.linalg.inv(Hm + damping * M) @ (JW @ r[:, :, None]))[..., 0]   # [B,N]
    if dof == "shift":
        return su + delta[:, 0:1], sv + delta[:, 1:2], th
    if dof == "rot":
        return su, sv, th + delta[:, 0:1]
    su_n, sv_n, th_n = su + delta[:, 0:1], sv + delta[:, 1:2], th + delta[:, 2:3]
    if redraw:
        uni = torch.distributions.uniform.Uniform(-1, 1)
        ru, rv = uni.sample([B, 1]).to(su.device), uni.sample([B, 1]).to(su.device)   # two CPU draws per call, in this order
        su_n = torch.where((su_n > -2.5) & (su_n < 2.5), su_n, ru.to(su_n.dtype))
        sv_n = torch.where((sv_n > -2.5) & (sv_n < 2.5), sv_n, rv.to(sv_n.dtype))
    return su_n, sv_n, th_n


def cam_uv_g2sp(A: int, su, sv, th, cam_k: torch.Tensor, gh: int, gw: int, ori_h: int, ori_w: int, rot: float, lat: float,
                lon: float, mpp: float):
    """models_kitti.py:54-160 (get_warp_sat2real + seq_warp_real2camera): every satellite pixel (row i, column j) is the
    ground-plane point (X south, 0, Z east) = mpp (i - A//2, 0, j - A//2); it is projected with P = K_l [R(-heading) | T],
    T = (sv lat, 1.65, -su lon), perspective divide by max(w, 1e-6); quotient-rule Jacobians, zero where w <= 1e-6.
    Returns uv [B,A,A,2] (ground-image pixel units) and jac [3,B,A,A,2]."""
    dt, dev = su.dtype, su.device
    B = su.shape[0]
    k = cam_k.to(dt).clone()
    k[:, 0] = cam_k[:, 0].to(dt) * gw / ori_w
    k[:, 1] = cam_k[:, 1].to(dt) * gh / ori_h
    kk = rot * math.pi / 180.0
    h = -(th.reshape(B) * kk)
    c, s = torch.cos(h), torch.sin(h)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    R = torch.stack([c, z, -s, z, o, z, s, z, c], dim=-1).reshape(B, 3, 3)
    dR = kk * torch.stack([s, z, c, z, z, z, -c, z, s], dim=-1).reshape(B, 3, 3)        # dR/dtheta (heading enters as -theta k)
    T = torch.stack([sv.reshape(B) * lat, 1.65 * o, -su.reshape(B) * lon], dim=-1)
    idx = torch.arange(A, device=dev, dtype=dt) - (A // 2)
    X, Z = (mpp * idx)[:, None].expand(A, A), (mpp * idx)[None, :].expand(A, A)           # rows -> X (south), columns -> Z (east)
    pts = torch.stack([X, torch.zeros_like(X), Z], dim=-1)                                  # [A,A,3]
    KR, KT = k @ R, (k @ T[:, :, None])[..., 0]
    uv1 = torch.einsum("bij,hwj->bhwi", KR, pts) + KT[:, None, None, :]
    w = uv1[..., 2:3].clamp_min(1e-6)
    uv = uv1[..., :2] / w
    vis = (uv1[..., 2:3] > 1e-6).to(dt)
    e_u = torch.tensor([0.0, 0.0, -lon], dtype=dt, device=dev)
    e_v = torch.tensor([lat, 0.0, 0.0], dtype=dt, device=dev)
    d_u = (k @ e_u)[:, None, None, :].expand_as(uv1)                                        # dP/dsu touches only the last column
    d_v = (k @ e_v)[:, None, None, :].expand_as(uv1)
    d_t = torch.einsum("bij,hwj->bhwi", k @ dR, pts)
    quot = lambda d: (d[..., :2] / w - uv1[..., :2] * d[..., 2:3] / (w * w)) * vis
    return uv, torch.stack([quot(d_u), quot(d_v), quot(d_t)], dim=0)


def lm_update_g2sp(su, sv, th, grd_proj, grd_conf_proj, sat_feat, dfeat, damping: torch.Tensor, using_weight: bool):
    """models_kitti.py:333-379: r = grd_proj - sat with no renormalisation, identity damping, always 3-DOF, no reset."""
    N, B = dfeat.shape[:2]
    C = sat_feat.shape[1]
    J = dfeat.reshape(N, B, -1).permute(1, 0, 2)                                            # [B,3,D]
    JW = J * grd_conf_proj.expand(-1, C, -1, -1).reshape(B, 1, -1) if using_weight else J
    r = (grd_proj - sat_feat).reshape(B, -1, 1)
    eye = torch.eye(N, dtype=J.dtype, device=J.device)[None]
    delta = -(torch.linalg.inv(JW @ J.transpose(1, 2) + damping * eye) @ (JW @ r))[..., 0]
    return su + delta[:, 0:1], sv + delta[:, 1:2], th + delta[:, 2:3]


def resolve_damping_tensor(args, damping_param: torch.Tensor, n: int, device) -> torch.Tensor:
    """models_kitti.py:960-966: trained 10^(-6 + 11 sigmoid(p)) or the fixed args.damping, shape [1,N]."""
    if getattr(args, "train_damping", 0):
        return 10.0 ** (-6.0 + damping_param.sigmoid() * 11.0)
    return args.damping * torch.ones(1, n, dtype=torch.float32, device=device)


def dof_name(args, always_full: bool) -> str:
    if always_full:
        return "full"
    if args.rotation_range == 0:
        return "shift"
    if args.shift_range_lat == 0 and args.shift_range_lon == 0:
        return "rot"
    return "full"
