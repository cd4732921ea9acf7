"""Input pipeline.  CPU: the numpy oracle (oracle/imgproc.py) is pinned against PIL + torchvision — the
library calls the reference's datasets make (dataLoader/KITTI_dataset.py:128-157, :256-288; Ford_dataset.py:178-209) —
bit for bit on every stage and on the whole sample preparation.  GPU: the CUDA stages (through the C ABI) against the
oracle and against PIL itself at the datasets' full sizes.  Everything is uint8 / integer work: the bar is bit-exact."""
import math
import os

import numpy as np
import pytest
import torch
from PIL import Image
from torchvision import transforms
import torchvision.transforms.functional as TF

os.environ.setdefault("HA_QUIET", "1")
from highlyaccurate_b200 import engine, input_pipeline as P  # noqa: E402
from oracle import imgproc as I  # noqa: E402

MPP = engine.kitti_meter_per_pixel()


def photo(h, w, seed):
    """A smooth random RGB image with a little pixel noise (so that nearest / bilinear / rounding choices all matter)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    a = np.asarray(Image.fromarray(base).resize((w, h), Image.BICUBIC)).copy()
    a ^= rng.integers(0, 8, a.shape, dtype=np.uint8)
    return a


def pil_kitti_satellite(a, heading, gx, gy, theta, lat=20.0, lon=20.0, rot=10.0):
    """The statements of KITTI_dataset.py:256-288 on a PIL image (gt_* after the sign flip of :267-268)."""
    sat_map = Image.fromarray(a)
    sat_rot = sat_map.rotate(-heading / np.pi * 180)
    sat_align_cam = sat_rot.transform(sat_rot.size, Image.AFFINE, (1, 0, 1.08 / MPP, 0, 1, 0.26 / MPP), resample=Image.BILINEAR)
    sat_rand_shift = sat_align_cam.transform(sat_align_cam.size, Image.AFFINE, (1, 0, gx * (lon / MPP), 0, 1, -gy * (lat / MPP)),
                                             resample=Image.BILINEAR)
    out = TF.center_crop(sat_rand_shift.rotate(theta * rot), 512)
    return transforms.Compose([transforms.Resize(size=[512, 512]), transforms.ToTensor()])(out).numpy()


def pil_ford_satellite(a, du, dv, yaw, gu, gv, theta, plat, plon, rot=10.0):
    """Ford_dataset.py:178-209."""
    sat_map = Image.fromarray(a)
    x = sat_map.transform(sat_map.size, Image.AFFINE, (1, 0, du, 0, 1, dv), resample=Image.BILINEAR).rotate(yaw)
    x = x.transform(x.size, Image.AFFINE, (1, 0, gu * plat, 0, 1, gv * plon), resample=Image.BILINEAR).rotate(theta * rot)
    return transforms.ToTensor()(TF.center_crop(x, 512)).numpy()


def pil_ground(a):
    return transforms.Compose([transforms.Resize(size=[256, 1024]), transforms.ToTensor()])(Image.fromarray(a)).numpy()


KITTI_SAMPLES = [(0.31, -0.3455326, -0.9653194, -0.0077375486), (-2.9, 0.37390798, 0.36219868, -0.75242347),
                 (1.5707963267948966, 0.0, 0.0, 0.0), (0.0, 0.933891, 0.42439047, 0.95673674)]


# ------------------------------------------------------------------------------------------------ CPU: oracle vs PIL
@pytest.mark.parametrize("hw", [(96, 128), (128, 128)])
def test_oracle_affine_stages_equal_pil(hw):
    h, w = hw
    a = photo(h, w, 1)
    pil = Image.fromarray(a)
    for ang in (0, 90, 180, 270, 13.7, -123.4, 359.99, 0.01, 45.0):
        np.testing.assert_array_equal(np.asarray(pil.rotate(ang)), I.affine(a, I.rotate_matrix(ang, w, h), I.NEAREST))
        np.testing.assert_array_equal(I.affine_nearest(a, P.rotate_coefficients(ang, w, h)), np.asarray(pil.rotate(ang)),
                                      err_msg="fixed-point path != PIL fast path at %r" % ang)
    for tx, ty in ((5.5151, 1.3277), (-17.25, 30.0), (0.0, 0.0), (3.0, -4.0), (102.13, -99.7), (500.0, 0.0)):
        ref = np.asarray(pil.transform(pil.size, Image.AFFINE, (1, 0, tx, 0, 1, ty), resample=Image.BILINEAR))
        np.testing.assert_array_equal(ref, I.affine(a, I.translate_matrix(tx, ty), I.BILINEAR))


@pytest.mark.parametrize("shape", [(375, 1242, 256, 1024), (96, 128, 64, 100), (215, 414, 64, 256), (100, 100, 100, 100),
                                   (64, 64, 128, 200), (64, 200, 64, 100), (90, 64, 30, 64)])
def test_oracle_resize_equals_pil(shape):
    h, w, oh, ow = shape
    a = photo(h, w, 2)
    ref = np.asarray(transforms.Resize(size=[oh, ow])(Image.fromarray(a)))
    np.testing.assert_array_equal(ref, I.resize_bilinear(a, oh, ow))
    np.testing.assert_array_equal(transforms.ToTensor()(Image.fromarray(ref)).numpy(), I.to_tensor(ref))


def test_oracle_whole_sample_preparation_equals_pil():
    a = photo(512, 512, 3)
    for hd, gx, gy, th in KITTI_SAMPLES[:2]:
        np.testing.assert_array_equal(pil_kitti_satellite(a, hd, gx, gy, th), I.kitti_satellite(a, hd, gx, gy, th, MPP))
    f = photo(640, 600, 4)
    args = (3.7, -11.2, 37.5, 0.4, -0.8, 0.6, 20 / 0.22, 20 / 0.22)
    np.testing.assert_array_equal(pil_ford_satellite(f, *args), I.ford_satellite(f, *args))
    np.testing.assert_array_equal(pil_ground(photo(375, 1242, 5)), I.ground_image(photo(375, 1242, 5)))
    np.testing.assert_array_equal(np.asarray(TF.center_crop(Image.fromarray(f), 512)), I.center_crop(f, 512))


def test_host_coefficients_and_file_list(tmp_path):
    for ang, w, h in ((13.7, 512, 512), (-200.3, 128, 96), (90.0, 96, 128)):
        m = I.rotate_matrix(ang, w, h)
        assert P.rotate_coefficients(ang, w, h) == m
    assert P.shift_coefficients(3, -4.5) == I.translate_matrix(3, -4.5)
    p = tmp_path / "test1_files.txt"
    p.write_text("2011_10_03/2011_10_03_drive_0042_sync/0000000939.png -0.3455326 -0.9653194 -0.0077375486\n"
                 "2011_09_30/2011_09_30_drive_0018_sync/0000002310.png 0.37390798 0.36219868 -0.75242347\n")
    names, sx, sy, th = P.parse_kitti_file_list(str(p))
    assert names[1].endswith("0000002310.png") and sx == [0.3455326, -0.37390798] and sy == [0.9653194, -0.36219868]
    assert th == [-0.0077375486, -0.75242347]
    with pytest.raises(Exception):
        P.ground_batch(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))          # CPU tensor: refused, never emulated


# ------------------------------------------------------------------------------------------------ GPU: CUDA vs oracle / PIL
DEV = "cuda:0"


@pytest.mark.gpu
@pytest.mark.parametrize("hw,bpp", [((96, 128), 3), ((128, 128), 4), ((512, 512), 3)])
def test_gpu_affine_stages_bit_exact(hw, bpp):
    from highlyaccurate_b200 import _lib
    h, w = hw
    imgs = np.stack([photo(h, w, 10 + i) for i in range(3)])
    src = imgs if bpp == 3 else np.concatenate([imgs, np.zeros_like(imgs[..., :1])], axis=-1)
    d_src = torch.from_numpy(src).to(DEV)
    L = _lib.lib()
    cases = [(I.NEAREST, [P.rotate_coefficients(a, w, h) for a in angs]) for angs in ((0, 90, 180), (270, 13.7, -123.4), (359.99, 0.01, 45.0))]
    cases += [(I.BILINEAR, [P.shift_coefficients(*t) for t in ts]) for ts in (((5.5151, 1.3277), (-17.25, 30.0), (0.0, 0.0)),
                                                                              ((3.0, -4.0), (102.13, -99.7), (5000.0, 0.0)))]
    cases += [(I.BILINEAR, [P.rotate_coefficients(a, w, h) for a in (10.0, -77.7, 181.0)])]       # general bilinear affine
    for resample, rows in cases:
        coef = torch.tensor(rows, dtype=torch.float64, device=DEV)
        out = torch.empty(3, h, w, 4, dtype=torch.uint8, device=DEV)
        _lib.check(L.ha_img_affine_u8(d_src.data_ptr(), bpp, out.data_ptr(), None, 0, 3, h, w, coef.data_ptr(), resample,
                                      torch.cuda.current_stream().cuda_stream), "ha_img_affine_u8")
        got = out.cpu().numpy()[..., :3]
        for i in range(3):
            np.testing.assert_array_equal(got[i], I.affine(imgs[i], rows[i], resample))
            pil = Image.fromarray(imgs[i])
            ref = pil.transform(pil.size, Image.AFFINE, rows[i], resample=Image.BILINEAR if resample else Image.NEAREST)
            np.testing.assert_array_equal(got[i], np.asarray(ref))


@pytest.mark.gpu
def test_gpu_kitti_satellite_batch_equals_pil():
    imgs = np.stack([photo(512, 512, 20 + i) for i in range(len(KITTI_SAMPLES))])
    hd, gx, gy, th = zip(*KITTI_SAMPLES)
    got = P.kitti_satellite_batch(torch.from_numpy(imgs).to(DEV), hd, gx, gy, th).cpu().numpy()
    assert got.shape == (len(KITTI_SAMPLES), 3, 512, 512) and got.dtype == np.float32
    for i, s in enumerate(KITTI_SAMPLES):
        np.testing.assert_array_equal(got[i], pil_kitti_satellite(imgs[i], *s))
        np.testing.assert_array_equal(got[i], I.kitti_satellite(imgs[i], *s, MPP))


@pytest.mark.gpu
def test_gpu_ford_satellite_batch_equals_pil():
    imgs = np.stack([photo(1280, 1280, 30 + i) for i in range(2)])
    par = [(3.7, -11.2, 37.5, 0.4, -0.8, 0.6), (-40.25, 18.0, -163.2, -1.0, 1.0, -1.0)]
    plat = plon = 20 / 0.22
    cols = list(zip(*par))
    got = P.ford_satellite_batch(torch.from_numpy(imgs).to(DEV), *cols, plat, plon).cpu().numpy()
    for i, p in enumerate(par):
        np.testing.assert_array_equal(got[i], pil_ford_satellite(imgs[i], *p, plat, plon))


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(375, 1242), (860, 1656), (256, 1024), (256, 1242), (375, 1024), (200, 700)])
def test_gpu_ground_batch_equals_pil(hw):
    h, w = hw
    imgs = np.stack([photo(h, w, 40 + i) for i in range(2)])
    got = P.ground_batch(torch.from_numpy(imgs).to(DEV)).cpu().numpy()
    for i in range(2):
        np.testing.assert_array_equal(got[i], pil_ground(imgs[i]))
        np.testing.assert_array_equal(got[i], I.ground_image(imgs[i]))


@pytest.mark.gpu
def test_gpu_prepared_batch_feeds_the_model():
    """The call chain a user of the reference's loader + model makes: decoded uint8 images -> tensors -> forward(mode='test')."""
    from highlyaccurate_b200.models_kitti import LM_S2GP
    from tests.cases import ref_args
    sat = torch.from_numpy(np.stack([photo(512, 512, 50), photo(512, 512, 51)])).to(DEV)
    grd = torch.from_numpy(np.stack([photo(375, 1242, 52), photo(375, 1242, 53)])).to(DEV)
    s = P.kitti_satellite_batch(sat, [0.3, -1.0], [0.2, -0.4], [0.1, 0.5], [0.3, -0.6])
    g = P.ground_batch(grd)
    net = LM_S2GP(ref_args(N_iters=1)).to(DEV)
    lat, lon, th = net(s, g, mode="test")
    assert lat.shape == (2,) and bool(torch.isfinite(torch.stack([lat, lon, th])).all())
