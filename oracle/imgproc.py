"""CPU oracle for the input pipeline.  TEST INFRASTRUCTURE ONLY.

The reference prepares every sample on the CPU with PIL + torchvision (dataLoader/KITTI_dataset.py:128-157, :256-288;
dataLoader/Ford_dataset.py:178-209): `Image.rotate` (nearest), two or three `Image.transform(AFFINE, BILINEAR)`,
`TF.center_crop`, `transforms.Resize` (PIL's antialiased bilinear resample) and `ToTensor`.  That arithmetic lives in
Pillow's C library (libImaging Geometry.c / Resample.c; the image pins Pillow 12.2.0, the reference pins nothing).
This file restates it in numpy integer / float64 arithmetic; tests/test_imgproc.py pins the restatement against PIL
itself on every stage (bit-exact uint8), so `parity: PINNED` against the reference's own dependency.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import math

import numpy as np

NEAREST, BILINEAR = 0, 2          # PIL.Image.Resampling values


# ----------------------------------------------------------------------------- affine coefficient builders (PIL Image.py)
def rotate_matrix(angle_deg: float, w: int, h: int):
    """PIL.Image.Image.rotate (expand=False, centre = image centre): the destination->source affine `data` it hands to
    transform(); None when PIL takes its exact fast paths (0 / 180 / 90 / 270 degrees), returned as ('copy'|'r180'|...)."""
    angle = angle_deg % 360.0
    if angle == 0:
        return "copy"
    if angle == 180:
        return "r180"
    if angle in (90, 270) and w == h:
        return "r90" if angle == 90 else "r270"
    cx, cy = w / 2, h / 2
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return m


def translate_matrix(tx: float, ty: float):
    """`transform(size, Image.AFFINE, (1, 0, tx, 0, 1, ty))` of KITTI_dataset.py:131-135."""
    return [1.0, 0.0, float(tx), 0.0, 1.0, float(ty)]


# ----------------------------------------------------------------------------- Geometry.c
def _fix(v: float) -> int:
    """#define FIX(v) FLOOR((v) * 65536.0 + 0.5) with FLOOR(v) = v >= 0 ? (int)v : (int)floor(v)."""
    x = v * 65536.0 + 0.5
    return int(x) if x >= 0.0 else int(math.floor(x))


def affine_nearest(img: np.ndarray, m) -> np.ndarray:
    """affine_fixed(): nearest neighbour in 16.16 fixed point, zero fill outside.  img [H,W,C] uint8."""
    H, W = img.shape[:2]
    a0, a1, a3, a4 = _fix(m[0]), _fix(m[1]), _fix(m[3]), _fix(m[4])
    a2 = _fix(m[2] + m[0] * 0.5 + m[1] * 0.5)
    a5 = _fix(m[5] + m[3] * 0.5 + m[4] * 0.5)
    x = np.arange(W, dtype=np.int64)[None, :]
    y = np.arange(H, dtype=np.int64)[:, None]
    xx = a2 + y * a1 + x * a0
    yy = a5 + y * a4 + x * a3
    # the C code accumulates in 32-bit ints; the check_fixed() guard keeps every value inside that range
    xin, yin = xx >> 16, yy >> 16
    ok = (xin >= 0) & (xin < W) & (yin >= 0) & (yin < H)
    out = np.zeros_like(img)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def affine_bilinear(img: np.ndarray, m) -> np.ndarray:
    """ImagingGenericTransform + affine_transform + bilinear_filter32RGB: double arithmetic, truncation to uint8,
    clamped neighbours, zero fill where the source point lies outside [0, size)."""
    H, W = img.shape[:2]
    xs = np.arange(W, dtype=np.float64)[None, :] + 0.5
    ys = np.arange(H, dtype=np.float64)[:, None] + 0.5
    xin = m[0] * xs + m[1] * ys + m[2]
    yin = m[3] * xs + m[4] * ys + m[5]
    inside = ~((xin < 0.0) | (xin >= W) | (yin < 0.0) | (yin >= H))
    xin = xin - 0.5
    yin = yin - 0.5
    x = np.floor(xin).astype(np.int64)
    y = np.floor(yin).astype(np.int64)
    dx = (xin - x)[..., None]
    dy = (yin - y)[..., None]
    x0 = np.clip(x, 0, W - 1)
    x1 = np.clip(x + 1, 0, W - 1)
    yc = np.clip(y, 0, H - 1)
    f = img.astype(np.int64)
    r0a, r0b = f[yc, x0], f[yc, x1]
    v1 = r0a + (r0b - r0a) * dx
    has2 = ((y + 1 >= 0) & (y + 1 < H))[..., None]
    y1 = np.clip(y + 1, 0, H - 1)
    r1a, r1b = f[y1, x0], f[y1, x1]
    v2 = np.where(has2, r1a + (r1b - r1a) * dx, v1)
    v = v1 + (v2 - v1) * dy
    out = v.astype(np.int64).astype(np.uint8)          # (UINT8)v1: truncation
    out[~inside] = 0
    return out


def affine(img: np.ndarray, m, resample: int) -> np.ndarray:
    if isinstance(m, str):
        return {"copy": img.copy(), "r180": img[::-1, ::-1].copy(), "r90": np.rot90(img, 1).copy(),
                "r270": np.rot90(img, 3).copy()}[m]
    return affine_nearest(img, m) if resample == NEAREST else affine_bilinear(img, m)


# ----------------------------------------------------------------------------- Resample.c (8 bits per channel, bilinear)
PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size: int, out_size: int):
    """precompute_coeffs() + normalize_coeffs_8bpc() for the bilinear filter (support 1) over the whole input range."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize, dtype=np.float64)
        for x in range(xmax):
            t = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - t if t < 1.0 else 0.0
        ww = 0.0
        for x in range(xmax):
            ww += w[x]
        if ww != 0.0:
            w[:xmax] = w[:xmax] / ww
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return kk, bounds, ksize


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    kk, bounds, ksize = resample_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Image.resize((out_w, out_h), BILINEAR) = ImagingResample: horizontal pass, then vertical pass, uint8 in between.
    A pass whose size does not change is skipped; equal sizes return a copy (Image.resize's early exit)."""
    H, W = img.shape[:2]
    out = img
    if W != out_w:
        out = _resample_axis(out, out_w, 1)
    if H != out_h:
        out = _resample_axis(out, out_h, 0)
    return out.copy() if out is img else out


def center_crop(img: np.ndarray, side: int) -> np.ndarray:
    """torchvision.transforms.functional.center_crop for an image at least `side` large in both directions."""
    H, W = img.shape[:2]
    top = int(round((H - side) / 2.0))
    left = int(round((W - side) / 2.0))
    return img[top:top + side, left:left + side]


def to_tensor(img: np.ndarray) -> np.ndarray:
    """transforms.ToTensor: uint8 HWC -> float32 CHW / 255 (a correctly rounded fp32 division)."""
    return (np.ascontiguousarray(img.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)).astype(np.float32)


# ----------------------------------------------------------------------------- the datasets' sample preparation
KITTI_GPS_SHIFT_LEFT = (1.08, 0.26)        # utils.py:13


def kitti_satellite(sat_u8: np.ndarray, heading: float, gt_shift_x: float, gt_shift_y: float, theta: float,
                    meter_per_pixel: float, shift_range_lat: float = 20.0, shift_range_lon: float = 20.0,
                    rotation_range: float = 10.0, side: int = 512) -> np.ndarray:
    """KITTI_dataset.py:128-157 (train: gt_* drawn by the caller) / :256-288 (test: gt_* from the file list, already
    negated as in :267-268).  `heading` in radians (oxts), returns float32 [3, side, side]."""
    H, W = sat_u8.shape[:2]
    x = affine(sat_u8, rotate_matrix(-heading / np.pi * 180, W, H), NEAREST)
    x = affine(x, translate_matrix(KITTI_GPS_SHIFT_LEFT[0] / meter_per_pixel, KITTI_GPS_SHIFT_LEFT[1] / meter_per_pixel), BILINEAR)
    x = affine(x, translate_matrix(gt_shift_x * (shift_range_lon / meter_per_pixel), -gt_shift_y * (shift_range_lat / meter_per_pixel)),
               BILINEAR)
    x = affine(x, rotate_matrix(theta * rotation_range, W, H), NEAREST)
    x = center_crop(x, side)
    x = resize_bilinear(x, side, side)
    return to_tensor(x)


def ford_satellite(sat_u8: np.ndarray, b_delta_u: float, b_delta_v: float, yaw_deg: float, gt_shift_u: float, gt_shift_v: float,
                   theta: float, shift_range_pixels_lat: float, shift_range_pixels_lon: float, rotation_range: float = 10.0,
                   side: int = 512) -> np.ndarray:
    """Ford_dataset.py:178-209: body-location shift (bilinear), yaw rotation (nearest), random shift (bilinear), random
    rotation (nearest), centre crop, ToTensor."""
    H, W = sat_u8.shape[:2]
    x = affine(sat_u8, translate_matrix(b_delta_u, b_delta_v), BILINEAR)
    x = affine(x, rotate_matrix(yaw_deg, W, H), NEAREST)
    x = affine(x, translate_matrix(gt_shift_u * shift_range_pixels_lat, gt_shift_v * shift_range_pixels_lon), BILINEAR)
    x = affine(x, rotate_matrix(theta * rotation_range, W, H), NEAREST)
    return to_tensor(center_crop(x, side))


def ground_image(grd_u8: np.ndarray, out_h: int = 256, out_w: int = 1024) -> np.ndarray:
    """grdimage_transform (KITTI_dataset.py:299-302, Ford_dataset.py:151-154): Resize([256, 1024]) + ToTensor."""
    return to_tensor(resize_bilinear(grd_u8, out_h, out_w))
