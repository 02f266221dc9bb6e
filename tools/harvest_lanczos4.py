"""Harvest OpenCV's float32 Lanczos-4 interpolation table (the one cv2.remap(..., INTER_LANCZOS4) uses for float images).

cv2.remap quantises the sampling position to 1/32 px (INTER_BITS = 5) and looks the 8 x 8 tap weights up in a table that
is the outer product of a 32 x 8 one-dimensional table.  The table is DATA of the third-party library the reference
calls (dng_warp_corr/chan_distortion_corr.py:94-97); like the Lab table it is harvested once by probing cv2 with
impulse images, and shipped as pysp_b200/data/lanczos4_tab_f32.npy (1 KB).

    python tools/harvest_lanczos4.py            # needs cv2; writes the table and checks the outer-product structure
"""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pysp_b200", "data", "lanczos4_tab_f32.npy")


def harvest():
    n = 24
    tab = np.zeros((32, 8), dtype=np.float32)
    src = np.zeros((n, n), dtype=np.float32)
    src[12, 12] = 1.0
    for f in range(32):
        # sampling at x = 12 + d + f/32, y = 12: the impulse sits at tap k = 3 - d of the horizontal kernel; the vertical
        # kernel at fraction 0 is the unit impulse, so the sample is the horizontal weight itself
        d = np.arange(-4, 4)
        mx = (12 + d + f / 32.0).astype(np.float32)[None, :]
        my = np.full_like(mx, 12.0)
        out = cv2.remap(src, mx, my, cv2.INTER_LANCZOS4)
        for i, dd in enumerate(d):
            tab[f, 3 - dd] = out[0, i]
    return tab


def check_outer_product(tab):
    rng = np.random.default_rng(0)
    src = np.zeros((24, 24), dtype=np.float32)
    bad = 0
    for _ in range(200):
        fy, fx = rng.integers(0, 32, 2)
        k1, k2 = rng.integers(0, 8, 2)
        src[:] = 0
        src[12 - 3 + k1 + 0, 12 - 3 + k2] = 1.0
        mx = np.array([[12 + fx / 32.0]], dtype=np.float32)
        my = np.array([[12 + fy / 32.0]], dtype=np.float32)
        got = cv2.remap(src, mx, my, cv2.INTER_LANCZOS4)[0, 0]
        want = np.float32(tab[fy, k1]) * np.float32(tab[fx, k2])
        bad += int(got != want)
    return bad


if __name__ == "__main__":
    cv2.setUseOptimized(False)
    t = harvest()
    print("row sums min/max", t.sum(axis=1).min(), t.sum(axis=1).max(), "f=0 row", t[0])
    print("2-D weights that are not the float32 product of the 1-D weights:", check_outer_product(t), "of 200")
    np.save(OUT, t)
    print("wrote", OUT)
