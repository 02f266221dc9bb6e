"""Device-resident throughput of every BASELINE.json configuration on one GPU (CUDA events, inputs larger than L2 or L2
flushed between repetitions).  The bench line of record is bench.py (config 2); this records the others.
    python tools/config_bench.py
"""
import json
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
m = cam_to_rgb_matrix(wbc.get_matrix())
wb = wbc.get_reciprocal_multipliers()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


res = {}
kw = dict(black=syn.BLACK, white=syn.WHITE)
f12 = engine.to_device(syn.scene(3000, 4000, 0))
o12 = torch.empty((3000, 4000, 3), dtype=torch.float32, device="cuda")
ms = timed(lambda: engine.develop(f12, wb, m, quality="fast", out_tensor=o12, **kw))
res["config1_fast_12MP"] = {"ms": ms, "mpix_s": 12e6 / ms / 1e3}
f24 = engine.to_device(syn.scene(4000, 6000, 0))
o24 = torch.empty((4000, 6000, 3), dtype=torch.float32, device="cuda")
for name, st in (("config2_ahd_stages1_24MP", 1), ("config3_ahd_stages3_24MP", 3), ("ahd_stages0_24MP", 0)):
    ms = timed(lambda: engine.develop(f24, wb, m, stages=st, out_tensor=o24, **kw))
    res[name] = {"ms": ms, "mpix_s": 24e6 / ms / 1e3}
base = (syn.scene(4000, 6000, 5).astype(np.float32) - 512.0) / 16383.0
evs = [8.0, 9.0, 10.0, 11.0, 12.0]
br = [engine.to_device(np.clip(base * np.float32(2.0 ** (2 - k)), 0, 1).astype(np.float32)) for k in range(5)]
tev, offs, bias = fusion_constants(evs, wb)


def hdr():
    fused, _ = engine.fuse_exposures(br, offs, bias, int(np.argmax(offs)), want_count=False)
    engine.develop(fused, wb, m, stages=1, hdr=True, out_tensor=o24)


ms = timed(hdr)
res["config4_hdr5_fuse_plus_ahd_24MP"] = {"ms": ms, "mpix_s": 24e6 / ms / 1e3, "bracket_mpix_s": 5 * 24e6 / ms / 1e3}
print(json.dumps({"device": torch.cuda.get_device_name(0), "configs": res}))
