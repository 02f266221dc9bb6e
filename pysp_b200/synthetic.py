"""Synthetic RGGB mosaics and colour constants for tests and bench.

Every generator here is bit-reproducible across hosts: only IEEE add / multiply / divide / rint on float32 / float64
arrays and NumPy's integer bit generators are used (no libm or SIMD transcendental, no `Generator.normal`), because the
full-size parity tests compare the GPU result with hashes that the unmodified reference produced on another machine
(tests/golden/fullsize_pins.json) and the input mosaic must be the same there.
"""
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


_TWO_PI_HI = 6.283185307179586            # float64(2 pi)
_TWO_PI_LO = 2.4492935982947064e-16       # 2 pi - float64(2 pi)
_INV_TWO_PI = 0.15915494309189535
_SIN_COEF = []                            # (-1)^k / (2k+1)!, k = 0..10
_f = 1.0
for _k in range(11):
    if _k:
        _f *= (2 * _k) * (2 * _k + 1)
    _SIN_COEF.append((-1.0 if _k & 1 else 1.0) / _f)


def det_sin(x):
    """sin(x) from float64 add / multiply / rint only (each NumPy ufunc rounds once, nothing is fused), so the bits
    are the same on every host.  Two-word range reduction to [-pi, pi], odd Taylor polynomial of degree 21
    (truncation error below 1e-9)."""
    x = np.asarray(x, dtype=np.float64)
    k = np.rint(x * _INV_TWO_PI)
    r = (x - k * _TWO_PI_HI) - k * _TWO_PI_LO
    r2 = r * r
    acc = np.full_like(r, _SIN_COEF[10])
    for c in _SIN_COEF[9::-1]:
        acc = acc * r2 + c
    return r * acc


def det_cos(x):
    return det_sin(np.asarray(x, dtype=np.float64) + 1.5707963267948966)


def det_noise(rng, shape, sigma):
    """Zero-mean noise of standard deviation `sigma` from integer draws only: the sum of four uniform integers on
    [-26, 26] (standard deviation 30.594), scaled.  float32."""
    n = rng.integers(-26, 27, size=(4,) + tuple(shape), dtype=np.int8).astype(np.int16).sum(axis=0, dtype=np.int16)
    return n.astype(np.float32) * np.float32(float(sigma) / 30.594117081556711)


def scene_base(height, width):
    """Noise-free float32 scene in sensor counts (see `scene`)."""
    xs = np.arange(width, dtype=np.float64)
    ys = np.arange(height, dtype=np.float64)
    s1 = det_sin(xs / 37.0).astype(np.float32)[None, :]
    c1 = det_cos(ys / 23.0).astype(np.float32)[:, None]
    diag = det_sin(np.arange(height + width, dtype=np.float64) / 5.0).astype(np.float32)
    iy, ix = np.mgrid[0:height, 0:width]
    base = (np.float32(0.5) + (np.float32(0.35) * s1) * c1) + np.float32(0.1) * diag[iy + ix]
    h2, w2 = height // 2, width // 2
    # zone plate, bottom-right quadrant
    yy = (ys[h2:] - h2) / max(h2, 1)
    xx = (xs[w2:] - w2) / max(w2, 1)
    arg = (0.5 * np.pi * max(h2, w2) / 2.0) * ((xx * xx)[None, :] + (yy * yy)[:, None])
    base[h2:, w2:] = (0.5 + 0.4 * det_cos(arg)).astype(np.float32)
    base[h2:, :w2] = 0.42                                   # flat quadrant, bottom-left
    # colour cast per CFA site so R/G/B differ
    cast = np.ones((height, width), dtype=np.float32)
    cast[0::2, 0::2] = 0.55
    cast[1::2, 1::2] = 0.75
    img = (np.float32(15000.0) * base) * cast + np.float32(512.0)
    ph, pw = max(2, height // 8), max(2, width // 8)
    img[ph:2 * ph, pw:2 * pw] = 20000.0                     # saturated patch
    return img


def scene(height, width, seed=0, noise=30.0, base=None):
    """14-bit RGGB mosaic: smooth sinusoids + diagonal high-frequency term, a zone-plate quadrant, a
    flat quadrant (mass integer ties in the homogeneity vote), a saturated patch (clip path) and
    read noise.  `base` = scene_base(height, width) may be passed to reuse it across seeds."""
    rng = np.random.default_rng(seed)
    img = scene_base(height, width) if base is None else base
    h2, w2 = height // 2, width // 2
    if noise > 0:
        img = img + det_noise(rng, img.shape, noise)
        img[h2 + h2 // 2:, :w2 // 2] = 15000.0 * 0.42 + 512.0   # part of the flat quadrant is noise-free
    return np.clip(np.rint(img), 0, 16383).astype(np.uint16)


def random_mosaic(height, width, seed=0, lo=0, hi=16384):
    rng = np.random.default_rng(seed)
    return rng.integers(lo, hi, size=(height, width)).astype(np.uint16)


def hdr_brackets(height, width, seed=5, n=5):
    """BASELINE config 4: `n` float32 exposures of one scene, one stop apart (brightest first, EV 8, 9, ...), each with its
    own read noise, clipped to [0, 1].  Returns (list of float32 [H, W] mosaics, list of EVs)."""
    base = (scene(height, width, seed, noise=0).astype(np.float32) - np.float32(512.0)) / np.float32(16383.0)
    brackets, evs = [], []
    for k in range(n):
        noise = det_noise(np.random.default_rng(1000 + k), base.shape, 30.0 / 16383.0)
        gain = np.float32(2.0 ** (n // 2 - k))
        brackets.append(np.clip(base * gain + noise, 0, 1).astype(np.float32))
        evs.append(10.0 - n // 2 + k)
    return brackets, evs
