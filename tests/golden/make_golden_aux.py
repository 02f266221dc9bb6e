"""Generate the golden fixtures of the steps either side of the develop path by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_aux.py
Call sites exercised (reference file:line): raw_correction.py:25-63 (`flat_frame_correction`),
raw_bad_pixel_corr.py:30-65 (`find_erroneous_pixels_threshold`), raw_hdr.py:7-83 (`fuse_exposures_from_debayer`,
incl. base_types/image_base.py:45-60 wb_undo/wb_apply and colorize/transform.py:21-53).
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from pysp_b200 import synthetic as syn  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def sensor_of(raw):
    return ((raw.astype(f32) - f32(512.0)).clip(0, 16383) / f32(16383.0)).astype(f32)


def flat_field(h, w, seed, zeros=0, negatives=0, dead_plane=False):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(f32)
    r2 = ((y - h / 2) / h) ** 2 + ((x - w / 2) / w) ** 2
    flat = (0.9 - 1.2 * r2 + rng.normal(0, 0.01, size=(h, w))).astype(f32)
    for _ in range(zeros):
        flat[rng.integers(0, h), rng.integers(0, w)] = 0.0
    for _ in range(negatives):
        flat[rng.integers(0, h), rng.integers(0, w)] = -0.2
    if dead_plane:
        flat[1::2, 1::2] = 0.0          # blue plane of the flat is black: quotient all inf -> plane left alone
    return flat


def main():
    rh.load()
    raw_correction = importlib.import_module("pySP.raw_correction")
    bad = importlib.import_module("pySP.raw_bad_pixel_corr")
    raw_hdr = importlib.import_module("pySP.raw_hdr")
    from pySP.image import RawBayerData
    from pySP.base_types.image_base import RawDemosaicData
    meta = dict(numpy=np.__version__)

    def bayer(sensor):
        img = RawBayerData()
        img.sensor_scaled = sensor
        return img

    # ---- flat_frame_correction
    cases = [("aux_flat34x50", 34, 50, dict(), False), ("aux_flat130x70_zeros", 130, 70, dict(zeros=9, negatives=5), False),
             ("aux_flat64x96_clamp_dead", 64, 96, dict(zeros=3, dead_plane=True), True),
             ("aux_flat200x304", 200, 304, dict(zeros=2), False)]
    for name, h, w, kw, clamp in cases:
        sensor = sensor_of(syn.scene(h, w, 31))
        sensor[sensor < 0.05] = f32(0.05)      # keep 0/0 out (a NaN plane maximum is not defined by the reference)
        flat = flat_field(h, w, 32, **kw)
        img = bayer(sensor.copy())
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            raw_correction.flat_frame_correction(img, bayer(flat.copy()), clamp_high=clamp)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), sensor=sensor, flat=flat, clamp=clamp,
                            out=np.asarray(img.sensor_scaled, dtype=f32), **meta)
        print(name, img.sensor_scaled.dtype, float(np.nanmax(img.sensor_scaled)))

    # ---- find_erroneous_pixels_threshold
    for name, h, w, delta, cnt in (("aux_hot34x50", 34, 50, 0.025, 5), ("aux_hot66x130", 66, 130, 0.01, 6), ("aux_hot8x8", 8, 8, 0.025, 3)):
        rng = np.random.default_rng(41)
        sensor = sensor_of(syn.scene(h, w, 33, noise=60.0))
        for _ in range(max(4, h * w // 200)):
            sensor[rng.integers(0, h), rng.integers(0, w)] = f32(rng.uniform(0.6, 1.0))
        masks = bad.find_erroneous_pixels_threshold(bayer(sensor), delta, cnt)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), sensor=sensor, min_delta=delta, min_neighbour_count=cnt,
                            masks=np.stack(masks), **meta)
        print(name, [int(m.sum()) for m in masks])

    # ---- fuse_exposures_from_debayer
    wb = rh.StubWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    wbc = wb.get_reciprocal_multipliers()
    rng = np.random.default_rng(51)
    for name, h, w, evs, norm in (("aux_fusecam24x40", 24, 40, [8.0, 10.0, 11.0, 12.5], False),
                                  ("aux_fusecam16x12_norm", 16, 12, [9.0, 10.0, 10.0], True),
                                  ("aux_fusecam20x28_nondyadic", 20, 28, [9.3, 10.1, 11.45], False)):
        base = rng.uniform(0.0, 1.6, size=(h, w, 3)).astype(f32)
        base[:3, :5] = 0.0                       # black in every exposure: weights sum to zero -> brightest-frame fallback
        base[3:5, :5] = 1e6                      # saturated in every exposure: weight 0 too, fallback with a non-zero value
        images, objs = [], []
        tgt = sum(evs) / len(evs)
        for e in evs:
            x = (base * f32(2.0 ** (tgt - e)) * wbc).astype(f32)
            x = np.minimum(x, (wbc * f32(1.0)).astype(f32)).astype(f32)     # sensor saturation in camera space
            if norm:
                x = (x / max(wbc)).astype(f32)
            images.append(x.copy())
            d = RawDemosaicData(x.copy(), wbc.copy(), wb_norm=norm)
            d.mat_xyz = wb.get_matrix()
            d.current_ev = e
            objs.append(d)
        lin, cnt = raw_hdr.fuse_exposures_from_debayer(objs)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), images=np.stack(images), evs=np.array(evs), wb=wbc, norm=norm,
                            lin=np.asarray(lin, dtype=f32), count=cnt, left=np.stack([o.image for o in objs]), **meta)
        print(name, lin.dtype, cnt.dtype, int((cnt == 0).sum()))


if __name__ == "__main__":
    main()
