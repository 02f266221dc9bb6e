"""Small driver for ncu: develop one synthetic 24 MP frame a few times (BASELINE config 2/3)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pysp_b200 import engine, synthetic as syn  # noqa: E402
from pysp_b200.colour import cam_to_rgb_matrix  # noqa: E402
from pysp_b200.wb_cct import CameraWhiteBalance  # noqa: E402

stages = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W = 4000, 6000
wbc = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
m = cam_to_rgb_matrix(wbc.get_matrix())
frame = engine.to_device(syn.scene(H, W, 0))
out = torch.empty((H, W, 3), dtype=torch.float32, device=frame.device)
for _ in range(reps):
    engine.develop(frame, wbc.get_reciprocal_multipliers(), m, stages=stages, black=syn.BLACK, white=syn.WHITE,
                   out_tensor=out)
torch.cuda.synchronize()
print("ok", float(out[::97, ::89].sum()))
