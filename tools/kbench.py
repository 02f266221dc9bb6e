"""A/B kernel timing: develop one synthetic 24 MP frame `reps` times with the library named by PYSP_B200_LIB (or the
product build), print per-kernel mean launch time (CUDA events on the launching stream, via pysp_timing_*) and a
hash of the output so that variants can be checked for bit-identical results.

    python tools/kbench.py [stages] [reps]            # one line of JSON
    python tools/kbench.py --variants a.so b.so ...   # runs each in a subprocess, prints a table
"""
import ctypes
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def one(stages, reps):
    import numpy as np
    import torch
    from pysp_b200 import _capi, engine, synthetic as syn
    from pysp_b200.colour import cam_to_rgb_matrix
    from pysp_b200.wb_cct import CameraWhiteBalance
    H, W = 4000, 6000
    wbc = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    m = cam_to_rgb_matrix(wbc.get_matrix())
    frames = [engine.to_device(syn.scene(H, W, s)) for s in range(4)]          # 192 MB of mosaics > L2
    out = torch.empty((H, W, 3), dtype=torch.float32, device=frames[0].device)
    lib = _capi.lib()

    def run(f):
        engine.develop(f, wbc.get_reciprocal_multipliers(), m, stages=stages, black=syn.BLACK, white=syn.WHITE, out_tensor=out)

    for f in frames:
        run(f)
    torch.cuda.synchronize()
    lib.pysp_timing_enable(1)
    for _ in range(reps):
        for f in frames:
            run(f)
    torch.cuda.synchronize()
    tot = (ctypes.c_double * 4)()
    n = (ctypes.c_int64 * 4)()
    _capi.check(lib.pysp_timing_collect(tot, n))
    lib.pysp_timing_enable(0)
    clk = (ctypes.c_uint64 * 32)()
    lib.pysp_debug_phase_clocks.argtypes = [ctypes.c_void_p]
    lib.pysp_debug_phase_clocks(clk)
    run(frames[0])
    torch.cuda.synchronize()
    h = hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:16]
    res = {"lib": os.path.basename(_capi.LIB_PATH), "stages": stages,
           "k1_ms": tot[0] / max(n[0], 1), "k2_ms": tot[1] / max(n[1], 1), "sha1": h}
    if any(clk):
        for k in range(2):
            t = sum(clk[16 * k:16 * k + 16])
            if t:
                res["k%d_phase_share" % (k + 1)] = [round(clk[16 * k + i] / t, 4) for i in range(16)]
    print(json.dumps(res))


def main():
    if "--variants" in sys.argv:
        i = sys.argv.index("--variants")
        pre = [a for a in sys.argv[1:i]]
        for so in sys.argv[i + 1:]:
            env = dict(os.environ, PYSP_B200_LIB=os.path.abspath(so))
            subprocess.call([sys.executable, os.path.abspath(__file__)] + pre, env=env)
        return
    stages = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    one(stages, reps)


if __name__ == "__main__":
    main()
