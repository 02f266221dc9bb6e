"""Harvest OpenCV's 33^3 RGB->Lab interpolation table (float32 COLOR_RGB2LAB path) into
pysp_b200/data/lab_lut33_i16.npy.

The reference calls cv2.cvtColor(f32, COLOR_RGB2LAB) (debayer/ahd.py:58,62).  For float32 input
OpenCV (pinned by the reference at opencv_python 4.10.0.84, requirements.txt:5; 4.13.0 here) does
not evaluate the analytic Lab formula: it quantises each channel to 1/16384 and interpolates
trilinearly, in integers, in a 33^3 table of int16.  The table is data, and
evaluating cvtColor exactly at the grid nodes (i/32, j/32, k/32) returns the node values
themselves (all interpolation weight on one corner), so the table can be read back losslessly.

Output layout: int16 [33(R)][33(G)][33(B)][3 (L, a, b)], with
  L = v * 100/16384,  a = v * 256/16384 - 128,  b likewise   (v the stored integer).
Run in the container that has cv2; the result is committed (215 622 bytes).
"""
import os
import sys

import cv2
import numpy as np


def harvest():
    g = (np.arange(33, dtype=np.float32) / np.float32(32.0)).astype(np.float32)
    r, gg, b = np.meshgrid(g, g, g, indexing="ij")
    rgb = np.stack([r, gg, b], axis=-1).reshape(1, -1, 3).astype(np.float32)
    lab = cv2.cvtColor(rgb, cv2.COLOR_RGB2LAB).reshape(33, 33, 33, 3).astype(np.float64)
    v = np.empty_like(lab)
    v[..., 0] = lab[..., 0] * 163.84
    v[..., 1] = (lab[..., 1] + 128.0) * 64.0
    v[..., 2] = (lab[..., 2] + 128.0) * 64.0
    vi = np.rint(v)
    assert np.abs(v - vi).max() < 1e-3, np.abs(v - vi).max()
    assert vi.min() >= -32768 and vi.max() <= 32767
    return vi.astype(np.int16)


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pysp_b200", "data",
                       "lab_lut33_i16.npy")
    lut = harvest()
    np.save(out, lut)
    print("wrote", os.path.abspath(out), lut.shape, lut.dtype, int(lut.min()), int(lut.max()),
          "cv2", cv2.__version__)
