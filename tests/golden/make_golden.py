"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference and cv2):
    python tests/golden/make_golden.py
The reference is imported through oracle/ref_harness.py (shims for un-installable third-party
modules, gcc build of its Cython extension into oracle/_ref/).  OpenCV is switched to its generic
code paths (`cv2.setUseOptimized(False)`, IPP off) so the float32 tap orders are specified
(SURVEY.md section 5.7); versions are recorded in each fixture.

Call chain exercised (reference file:line): normalization.py:4-25 -> image.py:191-197 (to_rggb,
demosaic) -> debayer/ahd.py:14-170 -> base_types/image_base.py:62-64 -> colorize/transform.py:76-99;
raw_hdr.py:85-158 for the HDR fuse.  Count maps and the direction map are recovered by wrapping
cv2.blur (debayer/ahd.py:133-134).
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from pysp_b200 import synthetic as syn  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_reference_fast(raw, black, white, pattern="RGGB"):
    """RawBayerData.demosaic(QualityDemosaic.Fast) -> camera RGB and linear sRGB (image.py:171-172)."""
    rh.load()
    from pySP.normalization import bayer_normalize
    from pySP.image import RawBayerData
    from pySP.base_types.image_base import BayerPattern
    from pySP.const import QualityDemosaic
    img = RawBayerData()
    img.sensor_scaled = bayer_normalize(raw, list(black), list(white))
    img.sensor_pattern = {"RGGB": BayerPattern.Rggb, "BGGR": BayerPattern.Bggr, "GRBG": BayerPattern.Grbg,
                          "GBRG": BayerPattern.Gbrg}[pattern]
    img.cam_wb = rh.StubWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    img.current_ev = 10.0
    dem = img.demosaic(QualityDemosaic.Fast)
    cam = np.array(dem.image, dtype=np.float32, copy=True)
    return dict(cam=cam, lin=dem.to_lin_srgb().astype(np.float32))


def run_reference(raw, black, white, stages, pattern="RGGB", mat=syn.MAT_XYZ_TO_CAM, xyz=syn.WHITE_XYZ,
                  sensor_override=None, hdr=False):
    import cv2
    rh.load()
    from pySP.normalization import bayer_normalize
    from pySP.image import RawBayerData, RawRggbBayerData
    from pySP.base_types.image_base import BayerPattern
    from pySP.const import QualityDemosaic
    wb = rh.StubWhiteBalance(mat, xyz)
    sensor = bayer_normalize(raw, list(black), list(white)) if sensor_override is None else sensor_override
    pat = {"RGGB": BayerPattern.Rggb, "BGGR": BayerPattern.Bggr, "GRBG": BayerPattern.Grbg,
           "GBRG": BayerPattern.Gbrg}[pattern]
    captured = []
    real_blur = cv2.blur

    def spy(src, ksize, *a, **k):
        captured.append(np.array(src, copy=True))
        return real_blur(src, ksize, *a, **k)

    cv2.blur = spy
    try:
        if hdr:
            img = RawRggbBayerData(sensor, wb, 10.0, 1.0, pat)
            img.set_hdr(True)
            dem = img.demosaic(QualityDemosaic.Best, stages)
        else:
            img = RawBayerData()
            img.sensor_scaled = sensor
            img.sensor_pattern = pat
            img.cam_wb = wb
            img.current_ev = 10.0
            dem = img.demosaic(QualityDemosaic.Best, stages)
    finally:
        cv2.blur = real_blur
    cam = np.array(dem.image, dtype=np.float32, copy=True)
    lin = dem.to_lin_srgb()
    cnt_h, cnt_v = captured[0], captured[1]
    pick_h = real_blur(cnt_h, (3, 3)) < real_blur(cnt_v, (3, 3))
    return dict(sensor=sensor.astype(np.float32), cam=cam, lin=lin.astype(np.float32),
                cnt_h=cnt_h.astype(np.uint8), cnt_v=cnt_v.astype(np.uint8), pick_h=pick_h)


def make_nonfinite(meta):
    """Float mosaic with +inf, -inf and NaN photosites at R, G1, G2 and B sites (debayer/ahd.py:139-145 blends the two
    candidates multiplicatively, so a non-finite value in the candidate that is not chosen still poisons the pixel).
    stages = 0: cv2.medianBlur's ordering of NaN is implementation-defined, so the median stages are not pinned."""
    sensor = (syn.scene(48, 64, 9).astype(np.float32) / np.float32(16383.0)).astype(np.float32)
    for (y, x), v in (((20, 30), np.inf), ((33, 11), np.nan), ((10, 51), -np.inf), ((41, 40), np.inf), ((8, 9), np.nan),
                      ((0, 0), np.inf), ((47, 63), np.nan), ((24, 63), -np.inf)):
        sensor[y, x] = v
    with np.errstate(invalid="ignore"):
        res = run_reference(None, None, None, 0, "RGGB", sensor_override=sensor)
    np.savez_compressed(os.path.join(OUT, "nonfinite48x64_s0.npz"), stages=0, pattern="RGGB", **res, **meta)
    print("nonfinite", int((~np.isfinite(res["cam"])).sum()), "non-finite values")


def main():
    import cv2
    rh.load()
    rh.pin_numerics(True)
    meta = dict(cv2=cv2.__version__, numpy=np.__version__)
    cases = []
    # name, raw, black, white, stages, pattern
    cases.append(("rand8x8_s0", syn.random_mosaic(8, 8, 1), syn.BLACK, syn.WHITE, 0, "RGGB"))
    cases.append(("rand8x8_s1", syn.random_mosaic(8, 8, 2), syn.BLACK, syn.WHITE, 1, "RGGB"))
    cases.append(("rand10x14_s1", syn.random_mosaic(10, 14, 3), syn.BLACK, syn.WHITE, 1, "RGGB"))
    cases.append(("rand4x6_s2", syn.random_mosaic(4, 6, 4), syn.BLACK, syn.WHITE, 2, "RGGB"))
    for s in (0, 1, 3):
        cases.append(("scene34x50_s%d" % s, syn.scene(34, 50, 5), syn.BLACK, syn.WHITE, s, "RGGB"))
    for pat in ("RGGB", "BGGR", "GRBG", "GBRG"):
        cases.append(("scene64x96_%s" % pat, syn.scene(64, 96, 6), (500, 510, 520, 530),
                      (16383, 16000, 15800, 16100), 1, pat))
    cases.append(("scene130x70_s1", syn.scene(130, 70, 7, noise=60.0), syn.BLACK, syn.WHITE, 1, "RGGB"))
    cases.append(("rand66x130_s1", syn.random_mosaic(66, 130, 8), syn.BLACK, syn.WHITE, 1, "RGGB"))
    for name, raw, black, white, stages, pat in cases:
        res = run_reference(raw, black, white, stages, pat)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), raw=raw, black=np.array(black),
                            white=np.array(white), stages=stages, pattern=pat, **res, **meta)
        print(name, res["lin"].shape, "pick_h frac %.3f" % res["pick_h"].mean())

    # QualityDemosaic.Fast (edge-assisted Gaussian)
    for name, raw, pat in (("fast_rand8x8", syn.random_mosaic(8, 8, 21), "RGGB"), ("fast_scene34x50", syn.scene(34, 50, 22), "RGGB"),
                           ("fast_scene64x96_GBRG", syn.scene(64, 96, 23), "GBRG"), ("fast_flat20x28", np.full((20, 28), 7000, np.uint16), "RGGB"),
                           ("fast_rand66x130", syn.random_mosaic(66, 130, 24), "BGGR")):
        res = run_reference_fast(raw, syn.BLACK, syn.WHITE, pat)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), raw=raw, black=np.array(syn.BLACK), white=np.array(syn.WHITE),
                            pattern=pat, **res, **meta)
        print(name, res["lin"].shape)

    # HDR: f32 mosaic with values up to 3.0, HDR flag set (debayer/ahd.py:52-59)
    rng = np.random.default_rng(9)
    base = syn.scene(48, 64, 9).astype(np.float32) / 16383.0
    gain = np.where(rng.random((48, 64)) < 0.2, 3.0, 1.0).astype(np.float32)
    sensor = (base * gain).astype(np.float32)
    for s in (0, 1):
        res = run_reference(None, None, None, s, "RGGB", sensor_override=sensor, hdr=True)
        np.savez_compressed(os.path.join(OUT, "hdr48x64_s%d.npz" % s), stages=s, pattern="RGGB",
                            hdr=True, **res, **meta)
        print("hdr s%d" % s, res["lin"].shape)

    make_nonfinite(meta)

    # HDR fuse (raw_hdr.py:85-158) + develop of the fused mosaic
    raw_hdr = rh.patch_hdr_ctor()
    from pySP.image import RawRggbBayerData
    from pySP.const import QualityDemosaic
    wb = rh.StubWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    base = (syn.scene(40, 56, 10, noise=0).astype(np.float32) - 512.0) / 16383.0
    brackets, evs = [], []
    for k in range(-2, 3):
        n = np.random.default_rng(k + 2).normal(0, 30.0 / 16383.0, size=base.shape).astype(np.float32)
        x = np.clip(base * np.float32(2.0 ** (-k)) + n, 0, 1).astype(np.float32)
        brackets.append(x)
        evs.append(10.0 + k)
    imgs = [RawRggbBayerData(b, wb, e, 1.0) for b, e in zip(brackets, evs)]
    fused, cnt = raw_hdr.fuse_exposures_to_raw(imgs)
    dem = fused.demosaic(QualityDemosaic.Best, 1)
    np.savez_compressed(os.path.join(OUT, "fuse5_40x56.npz"), brackets=np.stack(brackets),
                        evs=np.array(evs), fused=fused.sensor_scaled.astype(np.float32), count=cnt,
                        lim_sat=fused.lim_sat, target_ev=fused.current_ev, is_hdr=fused.get_hdr(),
                        cam=dem.image.astype(np.float32), lin=dem.to_lin_srgb(), **meta)
    print("fuse", fused.sensor_scaled.dtype, fused.lim_sat)

    # gamma (colorize/transform.py:89-99)
    from pySP.colorize.transform import lin_srgb_to_srgb
    x = np.linspace(-0.1, 1.1, 4096 * 3, dtype=np.float32).reshape(64, 64, 3)
    np.savez_compressed(os.path.join(OUT, "gamma.npz"), x=x, y=lin_srgb_to_srgb(x), **meta)
    y = lin_srgb_to_srgb(x)
    print("gamma dtype", y.dtype)

    # phase kernels (debayer/gaussian.py:19-53) and the 5-tap h (debayer/ahd.py:89-94)
    from pySP.debayer.gaussian import get_rgbg_kernel, CV2_DEFAULT_UNNORM_GAUSSIAN_KERNEL, BayerPatternPosition
    ktl = get_rgbg_kernel(CV2_DEFAULT_UNNORM_GAUSSIAN_KERNEL, BayerPatternPosition.TOP_LEFT)
    kbr = get_rgbg_kernel(CV2_DEFAULT_UNNORM_GAUSSIAN_KERNEL, BayerPatternPosition.BOTTOM_RIGHT)
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), base_tl=np.stack(ktl), base_br=np.stack(kbr),
                        gauss3=cv2.getGaussianKernel(3, 1.0).astype(np.float32), **meta)


if __name__ == "__main__":
    main()
