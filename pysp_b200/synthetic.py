"""Synthetic RGGB mosaics and colour constants for tests and bench."""
import numpy as np

# XYZ -> camera matrix, D65 white, neutral = mat @ white as float32, wb = 1/neutral
MAT_XYZ_TO_CAM = np.array([[0.6722, -0.0635, -0.0963],
                           [-0.4287, 1.2460, 0.2028],
                           [-0.0908, 0.2162, 0.5668]], dtype=np.float32)
WHITE_XYZ = np.array([0.31272 / 0.32903, 1.0, (1.0 - 0.31272 - 0.32903) / 0.32903], dtype=np.float64)
BLACK = (512, 512, 512, 512)
WHITE = (16383, 16383, 16383, 16383)


def neutral():
    return (MAT_XYZ_TO_CAM.astype(np.float64) @ WHITE_XYZ).astype(np.float32)


def wb_multipliers():
    return (1.0 / neutral()).astype(np.float32)


def scene_base(height, width):
    """Noise-free float32 scene in sensor counts (see `scene`)."""
    y, x = np.mgrid[0:height, 0:width].astype(np.float32)
    base = 0.5 + 0.35 * np.sin(x / 37.0) * np.cos(y / 23.0) + 0.1 * np.sin((x + y) / 5.0)
    h2, w2 = height // 2, width // 2
    # zone plate, bottom-right quadrant
    yy = (y[h2:, w2:] - h2) / max(h2, 1)
    xx = (x[h2:, w2:] - w2) / max(w2, 1)
    base[h2:, w2:] = 0.5 + 0.4 * np.cos(0.5 * np.pi * (xx * xx + yy * yy) * max(h2, w2) / 2.0)
    base[h2:, :w2] = 0.42                                   # flat quadrant, bottom-left
    # colour cast per CFA site so R/G/B differ
    cast = np.ones((height, width), dtype=np.float32)
    cast[0::2, 0::2] = 0.55
    cast[1::2, 1::2] = 0.75
    img = 15000.0 * base * cast + 512.0
    ph, pw = max(2, height // 8), max(2, width // 8)
    img[ph:2 * ph, pw:2 * pw] = 20000.0                     # saturated patch
    return img


def scene(height, width, seed=0, noise=30.0, base=None):
    """14-bit RGGB mosaic: smooth sinusoids + diagonal high-frequency term, a zone-plate quadrant, a
    flat quadrant (mass integer ties in the homogeneity vote), a saturated patch (clip path) and
    Gaussian read noise.  `base` = scene_base(height, width) may be passed to reuse it across seeds."""
    rng = np.random.default_rng(seed)
    img = scene_base(height, width) if base is None else base
    h2, w2 = height // 2, width // 2
    if noise > 0:
        img = img + rng.normal(0.0, noise, size=img.shape).astype(np.float32)
        img[h2 + h2 // 2:, :w2 // 2] = 15000.0 * 0.42 + 512.0   # part of the flat quadrant is noise-free
    return np.clip(np.rint(img), 0, 16383).astype(np.uint16)


def random_mosaic(height, width, seed=0, lo=0, hi=16384):
    rng = np.random.default_rng(seed)
    return rng.integers(lo, hi, size=(height, width)).astype(np.uint16)
