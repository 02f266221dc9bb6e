import os, subprocess, sys
code = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import ahd_spec as sp
from pysp_b200 import engine, synthetic as syn
H, W, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
raw = syn.scene(H, W, 1)
m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
out = engine.develop(engine.to_device(raw), syn.wb_multipliers(), m, stages=S, black=syn.BLACK, white=syn.WHITE)
torch.cuda.synchronize()
lin, _ = sp.develop(raw, syn.BLACK, syn.WHITE, syn.wb_multipliers(), syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, S)
print("mismatches", int((out.cpu().numpy().view(np.uint32) != lin.view(np.uint32)).sum()))
'''
for shape in [("8", "8", "0"), ("64", "96", "0"), ("64", "96", "1"), ("200", "320", "2")]:
    for mode in ("3", "1", "2", "0"):
        env = dict(os.environ, PYSP_DISABLE_TMA=mode)
        r = subprocess.run([sys.executable, "-c", code, *shape], env=env, capture_output=True, text=True)
        tail = (r.stdout.strip().splitlines() or [""])[-1] + " | " + (r.stderr.strip().splitlines() or [""])[-1][:150]
        print(shape, "disable_tma=" + mode, "rc", r.returncode, tail, flush=True)
