"""ORACLE for the DNG WarpRectilinear step -- test infrastructure only.

CPU restatement of dng_warp_corr/dng_warp_rectilinear_coords.pyx:18-95 (the coordinate table; plain C in
oracle/csrc/warp_table.c because the generated C mixes float and double arithmetic and calls libm's powf) and of the
`cv2.remap(plane, clip(map_x), clip(map_y), cv2.INTER_LANCZOS4)` call at dng_warp_corr/chan_distortion_corr.py:94-97.

Third-party arithmetic: cv2.remap (opencv-python 4.10.0.84 pinned in the reference's requirements.txt; 4.13.0 here) is
restated from its published algorithm: the float maps are quantised to 1/32 px with cvRound(v * 32) (half to even),
the integer part selects the 8 x 8 window starting 3 px up-left, the 5-bit fractions select the weights
tab[fy][k1] * tab[fx][k2] from the float32 Lanczos-4 table (harvested from cv2: tools/harvest_lanczos4.py ->
pysp_b200/data/lanczos4_tab_f32.npy), taps outside the image contribute the border value 0 (BORDER_CONSTANT).  Interior
pixels are summed row by row ( sum += w0*s0 + w1*s1 + ... + w7*s7 ), border pixels tap by tap, in float32.
Pinned by tests/golden/warp_*.npz, made by the unmodified reference (tests/golden/make_golden_warp.py).
"""
import ctypes
import os
import subprocess

import numpy as np

f32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force=False):
    src = os.path.join(_HERE, "csrc", "warp_table.c")
    so = os.path.join(_HERE, "_build", "liboracle_warp.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([gcc, "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_warp_table.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p,
                                          ctypes.c_float, ctypes.c_float, ctypes.c_float]
        lib.oracle_warp_table.restype = None
        _LIB = lib
    return _LIB


def remapping_table(coeffs, width, height, cam_center_norm, scale=1.0, seed=None):
    """compute_remapping_table / compute_offset_remapping_table (pyx:67-95): float32 [H, W, 2] = (x', y')."""
    k = np.asarray(coeffs, dtype=f32)
    assert k.shape == (6,)
    table = np.zeros((height, width, 2), dtype=f32)
    s = None if seed is None else np.ascontiguousarray(seed, dtype=f32)
    _lib().oracle_warp_table(table.ctypes.data, None if s is None else s.ctypes.data, width, height, k.ctypes.data,
                             f32(cam_center_norm[0]), f32(cam_center_norm[1]), f32(scale))
    return table


def lanczos_table():
    return np.load(os.path.join(_HERE, "..", "pysp_b200", "data", "lanczos4_tab_f32.npy")).astype(f32)


def remap_lanczos4(plane, map_x, map_y):
    """cv2.remap(plane, map_x, map_y, cv2.INTER_LANCZOS4) for a float32 plane and float32 maps, default border."""
    plane = np.ascontiguousarray(plane, dtype=f32)
    H, W = plane.shape
    tab = lanczos_table()
    sx = np.rint(np.asarray(map_x, dtype=f32) * f32(32)).astype(np.int64)
    sy = np.rint(np.asarray(map_y, dtype=f32) * f32(32)).astype(np.int64)
    ix, iy, fx, fy = (sx >> 5) - 3, (sy >> 5) - 3, sx & 31, sy & 31
    inside = (ix >= 0) & (ix + 8 <= W) & (iy >= 0) & (iy + 8 <= H)
    acc_int = np.zeros(sx.shape, dtype=f32)
    acc_brd = np.zeros(sx.shape, dtype=f32)
    for k1 in range(8):
        yy = iy + k1
        row = np.zeros(sx.shape, dtype=f32)
        for k2 in range(8):
            xx = ix + k2
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            s = np.where(ok, plane[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0))
            w = (tab[fy, k1] * tab[fx, k2]).astype(f32)
            p = (s * w).astype(f32)
            row = p if k2 == 0 else (row + p).astype(f32)
            acc_brd = np.where(ok, (acc_brd + p).astype(f32), acc_brd)
        acc_int = (acc_int + row).astype(f32)
    return np.where(inside, acc_int, acc_brd)


def apply_warp_rectilinear(image, coeffs, cam_center_norm, scale=1.0, prior=None):
    """opcode_warp_rectilinear (chan_distortion_corr.py:53-98): every plane remapped by its own coefficient set.
    image [H, W, C] float32 -> new array."""
    H, W, C = image.shape
    out = np.empty_like(image)
    for c in range(C):
        seed = None if prior is None else prior[..., c, :]
        t = remapping_table(coeffs[c], W, H, cam_center_norm, scale, seed)
        out[:, :, c] = remap_lanczos4(image[:, :, c], np.clip(t[:, :, 0], 0, W - 1), np.clip(t[:, :, 1], 0, H - 1))
    return out
