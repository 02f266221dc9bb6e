"""Small driver for compute-sanitizer: every kernel of the library once on small inputs (develop chain incl. edge and partial
tiles, flipped CFA, HDR, Fast, narrow outputs, fuse, flat-field, hot pixels, DNG warp)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pysp_b200 import engine, synthetic as syn  # noqa: E402
from pysp_b200.colour import cam_to_rgb_matrix  # noqa: E402
from pysp_b200.raw_hdr import fusion_constants  # noqa: E402
from pysp_b200.wb_cct import CameraWhiteBalance  # noqa: E402

wbc = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
m, wb = cam_to_rgb_matrix(wbc.get_matrix()), wbc.get_reciprocal_multipliers()
kw = dict(black=syn.BLACK, white=syn.WHITE)
raw = engine.to_device(syn.scene(250, 330, 0), pad_pitch=True)
for pattern, stages, out in (("RGGB", 1, "lin"), ("BGGR", 2, "cam"), ("GBRG", 0, "lin_f16"), ("GRBG", 1, "srgb_u8"), ("RGGB", 1, "srgb_u16")):
    engine.develop(raw, wb, m, stages=stages, pattern=pattern, out=out, **kw)
engine.develop(raw, wb, m, quality="fast", **kw)
dm = torch.zeros((250, 330), dtype=torch.uint8, device="cuda")
engine.develop(raw[20:150].contiguous(), wb, m, stages=1, rows=(30, 130), frame_height=250, in_row0=20, dir_map=dm[30:130], out_row0=30, **kw)
sens = engine.normalize(raw.contiguous(), syn.BLACK, syn.WHITE)
engine.develop(sens * 2.5, wb, m, stages=1, hdr=True)
br, evs = syn.hdr_brackets(120, 160, 5, 5)
tev, offs, bias = fusion_constants(evs, wb)
engine.fuse_exposures([engine.to_device(b) for b in br], offs, bias, int(np.argmax(offs)))
flat = torch.rand((250, 330), device="cuda")[:, :330] + 0.5
engine.flat_frame_correction(sens, flat[:250, :330].contiguous())
engine.bayer_plane_means(sens)
engine.find_hot_pixels_threshold(sens, 0.025, 5)
img = torch.rand((250, 330, 3), device="cuda")
engine.cam_to_rgb(img, m)
engine.srgb_gamma(img)
engine.wb_scale(img, wb, engine.WB_UNDO, normalized=True, max_wb=float(max(wb)))
engine.rgb_to_lab_cv2(img)
k = [(1.0012, -0.0321, 0.0104, -0.0023, 0.0007, -0.0004)] * 3
engine.warp_rectilinear(img, k, (0.49, 0.51))
engine.warp_rectilinear(img, [(2.2, -0.3, 0.1, 0.0, 0.05, -0.04)] * 3, (0.49, 0.51))
engine.remap_lanczos4(img, 1, engine.warp_table(250, 330, k[0], (0.49, 0.51)))
torch.cuda.synchronize()
print("ok", engine.kernel_launches(), "launches")
