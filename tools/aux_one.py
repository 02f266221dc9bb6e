"""Small driver for ncu: flat-field correction, hot-pixel detection and raw-space fuse on one 24 MP frame (launch list / DRAM bytes)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
from pysp_b200 import engine
rng = np.random.default_rng(0)
H, W = 4000, 6000
sensor = engine.to_device(rng.random((H, W), dtype=np.float32))
flat = engine.to_device((0.5 + 0.5 * rng.random((H, W), dtype=np.float32)).astype(np.float32))
for _ in range(2):
    engine.flat_frame_correction(sensor, flat)
    engine.find_hot_pixels_threshold(sensor, 0.025, 5)
    br = [sensor, flat, sensor, flat, sensor]
    engine.fuse_exposures(br, [0.25, 0.5, 1, 2, 4], np.ones((5, 3), np.float32), 4)
torch.cuda.synchronize()
# DNG warp and the remaining point-wise kernels
img = engine.to_device(rng.random((H, W, 3), dtype=np.float32))
k = [(1.0012, -0.0321, 0.0104, -0.0023, 0.0007, -0.0004)] * 3
for _ in range(2):
    engine.warp_rectilinear(img, k, (0.4987, 0.5021))
    engine.bayer_plane_means(flat)
torch.cuda.synchronize()
