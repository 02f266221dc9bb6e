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
