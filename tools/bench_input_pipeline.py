"""Input pipeline on the GPU vs the reference's PIL / torchvision statements on the host cores.
usage: python tools/bench_input_pipeline.py [B=32] [reps=20] [cpu_samples=8]   -> one JSON line
Algorithmic bytes per KITTI pair: satellite 512 x 512: (3+4) + (4+4) + (4+4) + (4+12) B/px = 10.2 MB;
ground 375 x 1242 -> 256 x 1024: 1.40 MB in, 1.54 MB out + 1.54 MB in, 3.15 MB out = 7.6 MB."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HA_QUIET", "1")
from highlyaccurate_b200 import input_pipeline as P  # noqa: E402
from tests.test_imgproc import photo, pil_ground, pil_kitti_satellite  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    ncpu = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    sat = np.stack([photo(512, 512, i) for i in range(4)])[rng.integers(0, 4, B)]
    grd = np.stack([photo(375, 1242, 10 + i) for i in range(4)])[rng.integers(0, 4, B)]
    hd, gx, gy, th = (rng.uniform(-3, 3, B), rng.uniform(-1, 1, B), rng.uniform(-1, 1, B), rng.uniform(-1, 1, B))
    d_sat, d_grd = torch.from_numpy(sat).to(dev), torch.from_numpy(grd).to(dev)
    hs, hg = torch.from_numpy(sat).pin_memory(), torch.from_numpy(grd).pin_memory()
    for _ in range(3):
        s, g = P.kitti_satellite_batch(d_sat, hd, gx, gy, th), P.ground_batch(d_grd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        s, g = P.kitti_satellite_batch(d_sat, hd, gx, gy, th), P.ground_batch(d_grd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # end to end from pinned host uint8 (what a PNG decoder leaves) incl. the H2D copy
    t0 = time.perf_counter()
    for _ in range(reps):
        s = P.kitti_satellite_batch(hs.to(dev, non_blocking=True), hd, gx, gy, th)
        g = P.ground_batch(hg.to(dev, non_blocking=True))
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / reps
    # the reference's statements (KITTI_dataset.py:256-288, :299-302) on the host, one sample at a time like a DataLoader worker
    t0 = time.perf_counter()
    for i in range(ncpu):
        pil_kitti_satellite(sat[i % B], hd[i % B], gx[i % B], gy[i % B], th[i % B])
        pil_ground(grd[i % B])
    cpu_ms = (time.perf_counter() - t0) * 1e3 / ncpu
    alg = (512 * 512 * (3 + 4 + 4 + 4 + 4 + 4 + 4 + 12) + 375 * 1242 * 3 + 2 * 375 * 1024 * 4 + 256 * 1024 * 12) * B
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    hbm = float(peaks.get("hbm_gbs", 6545.9))
    print(json.dumps({"what": "KITTI sample preparation (satellite 4 affine stages + crop + ToTensor; ground resize + ToTensor)",
                      "batch": B, "gpu_ms_per_batch": ms, "gpu_pairs_per_s": B / ms * 1e3,
                      "gpu_ms_per_batch_from_pinned_uint8": ms_e2e, "launches_per_batch": 4 + 4,
                      "algorithmic_gb_per_s": alg / ms / 1e6, "hbm_peak_gb_per_s": hbm, "frac": alg / ms / 1e6 / hbm,
                      "cpu_ms_per_pair_pil_1core": cpu_ms, "cpu_pairs_per_s_1core": 1e3 / cpu_ms,
                      "h2d_bytes_per_pair_uint8": 512 * 512 * 3 + 375 * 1242 * 3, "h2d_bytes_per_pair_float": (512 * 512 + 256 * 1024) * 12}))


if __name__ == "__main__":
    main()
