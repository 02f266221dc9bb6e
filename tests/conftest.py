import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = bits(a) != bits(b)
    n = int(bad.sum())
    if n:
        idx = np.argwhere(bad)[:5]
        raise AssertionError("%s: %d of %d values differ; first at %s: %s vs %s" % (
            what, n, bad.size, idx.tolist(), a[tuple(idx[0])], b[tuple(idx[0])]))


def assert_bit_equal_nan(a, b, what=""):
    """Bit equality where both are finite or infinite; NaN must sit at the same places (its payload / sign is not part of
    the contract: x86 produces the negative default NaN for inf * 0, the GPU the positive one)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "%s: NaN at different places (%d vs %d)" % (what, int(na.sum()), int(nb.sum()))
    assert_bit_equal(np.where(na, np.float32(0), a), np.where(nb, np.float32(0), b), what)


def gate2_check(lin, pick_h, lin_generic, what):
    """Parity gate (ii): `lin` / `pick_h` [H,W] against the reference run in its DEFAULT mode
    (OpenCV optimised code paths), reconstructed exactly from tests/golden/gate2_1024x1536_s1.npz and `lin_generic`
    (the generic-mode result, whose hash is checked).  Outside the 4*stages-px neighbourhood of direction flips:
    |delta| <= 1e-4 * max(|ref|, 1e-2) (see the note on the floor below).  Returns (flips of `pick_h`, the reference's own default-vs-generic flips)."""
    import hashlib
    d = golden("gate2_1024x1536_s1")
    H, W, stages = int(d["H"]), int(d["W"]), int(d["stages"])
    lin_generic = np.ascontiguousarray(lin_generic, dtype=np.float32)
    assert hashlib.sha256(lin_generic.tobytes()).hexdigest() == str(d["lin_generic_sha256"]), "generic-mode image is not the pinned one"
    delta = d["delta_i8"].astype(np.int64)
    delta.reshape(-1)[d["big_index"]] = d["big_delta"]
    ref = (lin_generic.view(np.int32).astype(np.int64) + delta).astype(np.int32).view(np.float32)
    assert hashlib.sha256(np.ascontiguousarray(ref).tobytes()).hexdigest() == str(d["lin_default_sha256"])
    pick_def = np.unpackbits(d["pick_default"], axis=1)[:, :W].astype(bool)
    pick_gen = np.unpackbits(d["pick_generic"], axis=1)[:, :W].astype(bool)
    flips = np.asarray(pick_h, dtype=bool) != pick_def
    own = pick_gen != pick_def
    r = 4 * stages                                          # two chained 5x5 medians per stage
    near = np.zeros((H + 2 * r, W + 2 * r), dtype=bool)
    for y, x in np.argwhere(flips):
        near[y:y + 2 * r + 1, x:x + 2 * r + 1] = True
    near = near[r:H + r, r:W + r]
    lin = np.asarray(lin, dtype=np.float32)
    diff = np.abs(lin.astype(np.float64) - ref)
    # SURVEY 8d states the floor as 1e-3.  Linear sRGB values that small are differences of 0.1-sized products (a dark
    # green next to a bright red, say), so a few ulp of the camera values already exceed 1e-4 * 1e-3 there: the reference's
    # OWN two modes break that floor for 11 of 4.7 M values on this frame (all with |ref| < 1.2e-3, absolute error below
    # 2e-7).  Asserted: 1e-4 relative with an absolute floor of 1e-6; the count under the 1e-3 floor is printed.
    strict = int((diff[~near] > 1e-4 * np.maximum(np.abs(ref[~near]), 1e-3)).sum())
    err = diff / np.maximum(np.abs(ref), 1e-2)
    worst = float(err[~near].max())
    print("%s vs default-mode reference: %d direction flips (the reference's own optimised-vs-generic count: %d) of %d; "
          "max relative error outside their %d-px neighbourhood %.3g (floor 1e-2), inside %.3g; values over 1e-4 with the "
          "1e-3 floor: %d" % (what, int(flips.sum()), int(own.sum()), H * W, r, worst,
                              float(err[near].max()) if near.any() else 0.0, strict))
    assert worst <= 1e-4, "%s: relative error %.3g outside the neighbourhood of flips" % (what, worst)
    assert strict <= 11, "%s: %d values break the 1e-3-floor criterion (the reference's own modes: 11)" % (what, strict)
    return int(flips.sum()), int(own.sum())


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_develop_cases():
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".npz") and (f.startswith("rand") or f.startswith("scene")):
            out.append(f[:-4])
    return out


@pytest.fixture(scope="session")
def colour_setup():
    from pysp_b200 import synthetic as syn
    from pysp_b200.wb_cct import CameraWhiteBalance
    return CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
