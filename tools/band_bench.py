"""Time one row band of the 100 MP frame on ONE GPU (the per-rank work of the 8-GPU row-band workload): rows [b, e) of an
8660 x 11548 frame developed from the band + halo rows, CUDA events, per-kernel times from the library's own event pairs.
    PYSP_B200_LIB=... python tools/band_bench.py [n_ranks] [rank]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pysp_b200 import _capi, engine, parallel, synthetic as syn  # noqa: E402
from pysp_b200.colour import cam_to_rgb_matrix  # noqa: E402
from pysp_b200.wb_cct import CameraWhiteBalance  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 8660, 11548
wbc = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
m, wb = cam_to_rgb_matrix(wbc.get_matrix()), wbc.get_reciprocal_multipliers()
b, e, hb, he = parallel.band_with_halo(H, world, rank, 1)
frame = np.tile(syn.scene(H // 4 + 2, W // 4 + 2, 3), (4, 4))[:H, :W]
held = engine.to_device(np.ascontiguousarray(frame[hb:he]), pad_pitch=True)
out = torch.empty((e - b, W, 3), dtype=torch.float32, device="cuda")
kw = dict(wb=wb, cam_to_srgb=m, stages=1, black=syn.BLACK, white=syn.WHITE, rows=(b, e), frame_height=H, in_row0=hb, out_tensor=out)
lib = _capi.lib()
for _ in range(3):
    engine.develop(held, **kw)
torch.cuda.synchronize()
lib.pysp_timing_enable(1)
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    engine.develop(held, **kw)
c.record()
torch.cuda.synchronize()
tot, n = (ctypes.c_double * 4)(), (ctypes.c_int64 * 4)()
_capi.check(lib.pysp_timing_collect(tot, n))
print(json.dumps({"lib": os.path.basename(_capi.LIB_PATH), "band_rows": e - b, "ms_per_band": a.elapsed_time(c) / 20,
                  "k1_ms": tot[0] / n[0], "k2_ms": tot[1] / n[1]}))
