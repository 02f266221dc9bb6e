"""Pin the float64 accumulation order of the reference's camera -> linear-sRGB matrix product.

Run in the build container only (needs /root/reference):   python tests/golden/make_fma_pins.py [seconds]

`cam_to_rgb_norm` (colorize/transform.py:52-53) evaluates `np.dot(rgb, color_mat.T).astype(float32)`: a float64 dgemm
with K = 3.  Whether the three products are summed with fused multiply-adds changes the float64 result by at most
one ulp, which changes the float32 result only when the sum sits next to a float32 rounding boundary (about one
value in 2^29).  This script searches camera-RGB triples in [0,1] for which
    unfused  f32( (m0*c0 + m1*c1) + m2*c2 )          and
    fused    f32( fma(m2, c2, fma(m1, c1, m0*c0)) )
differ (exact rational arithmetic decides), then runs the UNMODIFIED reference on them and stores inputs and outputs
in tests/golden/dot_fma_pins.npz.  Observed here (numpy + OpenBLAS, x86-64): the reference equals `fused` on every pin.
"""
import os
import sys
import time
from fractions import Fraction

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle import ahd_spec as sp  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from pysp_b200 import synthetic as syn  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def exact_variants(m, c):
    mm = [Fraction(float(x)) for x in m]
    f = [Fraction(float(x)) for x in c]
    p0 = float(mm[0] * f[0])                                   # float(Fraction) is correctly rounded
    unf = float(Fraction(float(Fraction(p0) + Fraction(float(mm[1] * f[1])))) + Fraction(float(mm[2] * f[2])))
    fus = float(Fraction(float(Fraction(p0) + mm[1] * f[1])) + mm[2] * f[2])
    return np.float32(unf), np.float32(fus)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 240.0
    m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    rng = np.random.default_rng(2024)
    pins, unf_v, fus_v, rows = [], [], [], []
    t0 = time.time()
    trials = 0
    n = 8_000_000
    while time.time() - t0 < budget:
        c = rng.random((n, 3), dtype=np.float32)
        cd = c.astype(np.float64)
        trials += n
        for k in range(3):
            u = (m[k, 0] * cd[:, 0] + m[k, 1] * cd[:, 1]) + m[k, 2] * cd[:, 2]
            low = (u.view(np.uint64) & np.uint64((1 << 29) - 1)).astype(np.int64)
            for i in np.nonzero(np.abs(low - (1 << 28)) <= 1)[0]:
                a, b = exact_variants(m[k], c[i])
                if a != b:
                    pins.append(c[i].copy()); unf_v.append(a); fus_v.append(b); rows.append(k)
    x = np.array(pins, dtype=np.float32).reshape(1, -1, 3)
    rh.load()
    from pySP.colorize.transform import cam_to_lin_srgb
    from pySP.wb_cct.helpers_cam_mat import MatXyzToCamera
    y = cam_to_lin_srgb(x.copy(), MatXyzToCamera(np.asarray(syn.MAT_XYZ_TO_CAM), np.asarray(syn.WHITE_XYZ, dtype=np.float64)))
    y = np.asarray(y, dtype=np.float32)
    rows = np.array(rows)
    got = y[0, np.arange(len(rows)), rows]
    print("trials %d, pins %d; reference == fused on %d, == unfused on %d" % (
        trials, len(rows), int((got == np.array(fus_v)).sum()), int((got == np.array(unf_v)).sum())))
    np.savez_compressed(os.path.join(OUT, "dot_fma_pins.npz"), x=x, y=y, row=rows, unfused=np.array(unf_v),
                        fused=np.array(fus_v), m=m, numpy=np.__version__)


if __name__ == "__main__":
    main()
