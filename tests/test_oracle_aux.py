"""oracle/aux_spec.py (flat-field correction, hot-pixel detection, camera-space HDR fusion) against the fixtures
produced by the unmodified reference (tests/golden/make_golden_aux.py).  Bit-exact."""
import numpy as np
import pytest

from conftest import assert_bit_equal, golden
from oracle import ahd_spec as sp
from oracle import aux_spec as ax
from pysp_b200 import synthetic as syn

FLAT = ["aux_flat34x50", "aux_flat130x70_zeros", "aux_flat64x96_clamp_dead", "aux_flat200x304"]
HOT = ["aux_hot34x50", "aux_hot66x130", "aux_hot8x8"]
FUSE = ["aux_fusecam24x40", "aux_fusecam16x12_norm", "aux_fusecam20x28_nondyadic"]


def assert_same_float_bits(a, b, what):
    """bit equality with NaN == NaN (payloads are not compared)"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    nan = np.isnan(a) & np.isnan(b)
    assert_bit_equal(np.where(nan, 0, a), np.where(nan, 0, b), what)


def test_pairwise_sum_is_numpys():
    rng = np.random.default_rng(5)
    for n in (1, 7, 8, 9, 127, 128, 129, 1000, 30000, 123457):
        a = rng.random(n).astype(np.float32)
        assert ax.pairwise_sum(a) == np.add.reduce(a), n
    m = rng.random((200, 300)).astype(np.float32)
    plane = m[0::2, :].astype(np.float32)[:, 1::2]          # bayer_to_rgbg's strided view (bayer_chan_mixer.py:13-21)
    assert ax.plane_mean(plane) == np.mean(plane)


@pytest.mark.parametrize("name", FLAT)
def test_flat_frame_correction(name):
    d = golden(name)
    assert_same_float_bits(ax.flat_frame_correction(d["sensor"], d["flat"], bool(d["clamp"])), d["out"], name)


@pytest.mark.parametrize("name", HOT)
def test_hot_pixel_threshold(name):
    d = golden(name)
    masks = ax.find_erroneous_pixels_threshold(d["sensor"], float(d["min_delta"]), int(d["min_neighbour_count"]))
    assert np.array_equal(np.stack(masks), d["masks"])


@pytest.mark.parametrize("name", FUSE)
def test_fuse_exposures_from_debayer(name):
    d = golden(name)
    m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    n = len(d["evs"])
    lin, cnt, left = ax.fuse_exposures_from_debayer(list(d["images"]), list(d["evs"]), d["wb"], m, [bool(d["norm"])] * n)
    assert_bit_equal(lin, d["lin"], "fused linear sRGB")
    assert np.array_equal(cnt, d["count"])
    assert_bit_equal(np.stack(left), d["left"], "images after the wb_undo/wb_apply round trip")
