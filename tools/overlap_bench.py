"""Do K1 (FMA / issue bound) and K2 (min/max-pipe bound) overlap when kernels of different frames share an SM?

Device-resident 24 MP frames developed round-robin on S streams (S = 1, 2, 3); with small-tile builds (tools/kbench-style
variants named by PYSP_B200_LIB) one CTA of each kernel fits on an SM at the same time.  Prints whole-run Gpix/s per S.
    PYSP_B200_LIB=variants/lib_x.so python tools/overlap_bench.py [stages]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pysp_b200 import _capi, engine, synthetic as syn  # noqa: E402
from pysp_b200.colour import cam_to_rgb_matrix  # noqa: E402
from pysp_b200.wb_cct import CameraWhiteBalance  # noqa: E402

stages = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H, W = 4000, 6000
wbc = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
m, wb = cam_to_rgb_matrix(wbc.get_matrix()), wbc.get_reciprocal_multipliers()
base = syn.scene_base(H, W)
frames = [engine.to_device(syn.scene(H, W, s, base=base)) for s in range(6)]
res = {"lib": os.path.basename(_capi.LIB_PATH), "stages": stages}
for S in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(S)]
    outs = [torch.empty((H, W, 3), dtype=torch.float32, device="cuda") for _ in range(S)]

    def run(n):
        for i in range(n):
            k = i % S
            engine.develop(frames[i % len(frames)], wb, m, stages=stages, black=syn.BLACK, white=syn.WHITE, out_tensor=outs[k], stream=streams[k])

    run(2 * S)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 60
    a.record()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    run(n)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    res["streams_%d_gpix_s" % S] = round(n * H * W / a.elapsed_time(b) / 1e6, 2)
print(json.dumps(res))
