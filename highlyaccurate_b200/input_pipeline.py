"""Sample preparation of the reference's datasets on the GPU.

The reference builds every sample on the CPU with PIL + torchvision inside `Dataset.__getitem__`
(dataLoader/KITTI_dataset.py:128-157 train / :256-288 test; dataLoader/Ford_dataset.py:178-209): rotate the satellite
tile by the vehicle heading, shift it to the camera, apply the ground-truth (or random) shift and rotation, centre-crop,
resize, ToTensor; resize the ground image to 256 x 1024, ToTensor.  Here the decoded uint8 images of a whole batch go
through the same stages in libha_b200.so (csrc/imgproc.cu) and come out as the [B,3,512,512] / [B,3,256,1024] fp32 tensors
`LM_S2GP.forward` takes — identical, bit for bit, to what the reference's loader produces.

The host side only builds the affine coefficients, exactly as PIL's Python layer does (Image.rotate: `round(cos, 15)`,
centre = size / 2, python floats), and ships them in one copy.  Decoding PNGs and reading the file lists stays on the
CPU (`parse_kitti_file_list` reads the reference's `test{1,2}_files.txt` format).  No CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib, engine
from ._lib import check

NEAREST, BILINEAR = 0, 2                    # PIL.Image.Resampling
KITTI_GPS_SHIFT_LEFT = (1.08, 0.26)         # utils.py:13 CameraGPS_shift_left
SAT_SIDE = 512                              # utils.py:11 SatMap_process_sidelength
GRD_H, GRD_W = 256, 1024                    # KITTI_dataset.py:27-28


def rotate_coefficients(angle_deg: float, w: int, h: int) -> List[float]:
    """The affine `data` PIL.Image.Image.rotate(angle) hands to transform() (expand off, centre = image centre).  PIL's exact
    fast paths for 0 / 90 / 180 / 270 degrees are what the fixed-point kernel produces from the same matrix."""
    angle = angle_deg % 360.0
    cx, cy = w / 2, h / 2
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return m


def shift_coefficients(tx: float, ty: float) -> List[float]:
    return [1.0, 0.0, float(tx), 0.0, 1.0, float(ty)]


def _as_u8_batch(img: torch.Tensor) -> Tuple[torch.Tensor, int]:
    engine._require_cuda(img, "image batch")
    if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] not in (3, 4):
        raise _lib.HaError("images must be uint8 [B, H, W, 3] (RGB) or [B, H, W, 4] (RGBX)")
    return img.contiguous(), img.shape[-1]


def affine_chain(img: torch.Tensor, stages: Sequence[Tuple[Sequence[Sequence[float]], int]], crop_side: int) -> torch.Tensor:
    """Runs `stages` = [(per-image coefficient rows [B][6], resample), ...] one after the other (each a
    ha_img_affine_u8 launch over the batch); the last one also centre-crops and converts: -> [B, 3, side, side] fp32."""
    img, bpp = _as_u8_batch(img)
    B, H, W, _ = img.shape
    L = _lib.lib()
    coef = torch.tensor([rows for rows, _ in stages], dtype=torch.float64).reshape(len(stages), B, 6)
    coef = coef.to(img.device, non_blocking=True)
    st = engine._stream_ptr()
    bufs = [torch.empty(B, H, W, 4, dtype=torch.uint8, device=img.device) for _ in range(min(2, len(stages) - 1))]
    out = torch.empty(B, 3, crop_side, crop_side, dtype=torch.float32, device=img.device)
    cur, cur_bpp = img, bpp
    for k, (_, resample) in enumerate(stages):
        last = k == len(stages) - 1
        dst = None if last else bufs[k % 2]
        check(L.ha_img_affine_u8(cur.data_ptr(), cur_bpp, None if last else dst.data_ptr(), out.data_ptr() if last else None,
                                 crop_side if last else 0, B, H, W, coef[k].data_ptr(), resample, st), "ha_img_affine_u8")
        if not last:
            cur, cur_bpp = dst, 4
    return out


def kitti_satellite_batch(sat_u8: torch.Tensor, heading: Sequence[float], gt_shift_x: Sequence[float],
                          gt_shift_y: Sequence[float], theta: Sequence[float], shift_range_lat: float = 20.0,
                          shift_range_lon: float = 20.0, rotation_range: float = 10.0) -> torch.Tensor:
    """KITTI_dataset.py:128-157 / :256-288 for a batch.  sat_u8 [B,H,W,3] uint8 (the decoded satellite tiles), `heading`
    in radians (oxts field 5), gt_shift_x / gt_shift_y / theta in [-1, 1] as the dataset uses them AFTER its sign flip
    (:267-268: gt_shift_x = -float(field)).  Returns [B,3,512,512] fp32 in [0,1]."""
    B, H, W, _ = sat_u8.shape
    mpp = engine.kitti_meter_per_pixel()                                   # utils.get_meter_per_pixel(scale=1)
    px_lat, px_lon = shift_range_lat / mpp, shift_range_lon / mpp          # :60-61
    s1 = [rotate_coefficients(-float(h) / np.pi * 180, W, H) for h in heading]                                   # :128
    s2 = [shift_coefficients(KITTI_GPS_SHIFT_LEFT[0] / mpp, KITTI_GPS_SHIFT_LEFT[1] / mpp)] * B                  # :129-133
    s3 = [shift_coefficients(float(x) * px_lon, -float(y) * px_lat) for x, y in zip(gt_shift_x, gt_shift_y)]     # :141-146
    s4 = [rotate_coefficients(float(t) * rotation_range, W, H) for t in theta]                                   # :149-151
    # :153 center_crop(512); :157 Resize([512, 512]) of a 512 x 512 image is PIL's early-exit copy
    return affine_chain(sat_u8, [(s1, NEAREST), (s2, BILINEAR), (s3, BILINEAR), (s4, NEAREST)], SAT_SIDE)


def ford_satellite_batch(sat_u8: torch.Tensor, b_delta_u: Sequence[float], b_delta_v: Sequence[float], yaw_deg: Sequence[float],
                         gt_shift_u: Sequence[float], gt_shift_v: Sequence[float], theta: Sequence[float],
                         shift_range_pixels_lat: float, shift_range_pixels_lon: float, rotation_range: float = 10.0,
                         side: int = SAT_SIDE) -> torch.Tensor:
    """Ford_dataset.py:178-209 for a batch: body-location shift (bilinear), yaw rotation (nearest), ground-truth shift
    (bilinear), ground-truth rotation (nearest), centre crop to `side`, ToTensor."""
    B, H, W, _ = sat_u8.shape
    s1 = [shift_coefficients(float(u), float(v)) for u, v in zip(b_delta_u, b_delta_v)]                          # :181-184
    s2 = [rotate_coefficients(float(y), W, H) for y in yaw_deg]                                                  # :187
    s3 = [shift_coefficients(float(u) * shift_range_pixels_lat, float(v) * shift_range_pixels_lon)
          for u, v in zip(gt_shift_u, gt_shift_v)]                                                               # :193-198
    s4 = [rotate_coefficients(float(t) * rotation_range, W, H) for t in theta]                                   # :201
    return affine_chain(sat_u8, [(s1, BILINEAR), (s2, NEAREST), (s3, BILINEAR), (s4, NEAREST)], side)


def ground_batch(grd_u8: torch.Tensor, out_h: int = GRD_H, out_w: int = GRD_W) -> torch.Tensor:
    """grdimage_transform (KITTI_dataset.py:299-302, Ford_dataset.py:151-154): Resize([256, 1024]) + ToTensor for a batch
    of decoded ground images [B,H,W,3] uint8 (KITTI 375 x 1242, Ford 860 x 1656) -> [B,3,256,1024] fp32."""
    img, bpp = _as_u8_batch(grd_u8)
    B, H, W, _ = img.shape
    L = _lib.lib()
    need = L.ha_img_resize_workspace_bytes(B, H, W, out_h, out_w)
    ws = torch.empty(max(need, 1), dtype=torch.uint8, device=img.device)
    out = torch.empty(B, 3, out_h, out_w, dtype=torch.float32, device=img.device)
    check(L.ha_img_resize_to_tensor(img.data_ptr(), bpp, B, H, W, out_h, out_w, out.data_ptr(), ws.data_ptr(), need,
                                    engine._stream_ptr()), "ha_img_resize_to_tensor")
    return out


def parse_kitti_file_list(path: str):
    """The on-disk format of dataLoader/test{1,2}_files.txt (KITTI_dataset.py:205-207): one sample per line,
    `<day>/<drive>/<image>.png shift_x shift_y theta`.  Returns (file names, gt_shift_x, gt_shift_y, theta) with the sign
    flip of :267-268 applied, i.e. ready for kitti_satellite_batch; the labels the model is scored against are
    -gt_shift_x, -gt_shift_y, theta (:284-286)."""
    names, sx, sy, th = [], [], [], []
    with open(path, "r") as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            name, x, y, t = line.split(" ")
            names.append(name)
            sx.append(-float(x))
            sy.append(-float(y))
            th.append(float(t))
    return names, sx, sy, th
