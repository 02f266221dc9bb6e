"""The DNG WarpRectilinear oracle (oracle/warp_spec.py, oracle/csrc/warp_table.c) against fixtures produced by the
unmodified reference (tests/golden/make_golden_warp.py): coordinate tables (dng_warp_rectilinear_coords.pyx:67-95) and the
Lanczos-4 resampling of cv2.remap (chan_distortion_corr.py:94-97), bit for bit in this container."""
import numpy as np
import pytest

from conftest import assert_bit_equal, golden
from oracle import warp_spec as ws

CASES = ["warp_96x128", "warp_70x50_scale"]


@pytest.mark.parametrize("name", CASES)
def test_tables_match_reference(name):
    d = golden(name)
    H, W, _ = d["image"].shape
    for i in range(3):
        t = ws.remapping_table(d["coeffs"][i], W, H, d["centre"], float(d["scale"]))
        assert_bit_equal(t, d["tables"][i], "compute_remapping_table, plane %d" % i)
        tp = ws.remapping_table(d["coeffs"][i], W, H, d["centre"], float(d["scale"]), seed=d["prior"][..., i, :])
        assert_bit_equal(tp, d["tables_prior"][i], "compute_offset_remapping_table, plane %d" % i)


@pytest.mark.parametrize("name", CASES)
def test_warp_matches_reference(name):
    d = golden(name)
    out = ws.apply_warp_rectilinear(d["image"], d["coeffs"], d["centre"], float(d["scale"]))
    assert_bit_equal(out, d["warped"], "apply_opcode_3_warp")
    outp = ws.apply_warp_rectilinear(d["image"], d["coeffs"], d["centre"], float(d["scale"]), prior=d["prior"])
    assert_bit_equal(outp, d["warped_prior"], "apply_opcode_3_warp with a prior")


def test_remap_restatement_against_cv2():
    """the NumPy restatement of cv2.remap(INTER_LANCZOS4) on random maps, incl. positions at and beyond the borders"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    src = rng.random((40, 56), dtype=np.float32)
    mx = rng.uniform(-2, 58, size=(64, 64)).astype(np.float32)
    my = rng.uniform(-2, 42, size=(64, 64)).astype(np.float32)
    cv2.setUseOptimized(False)
    want = cv2.remap(src, mx, my, cv2.INTER_LANCZOS4)
    got = ws.remap_lanczos4(src, mx, my)
    inside = (mx >= 0) & (mx <= 55) & (my >= 0) & (my <= 39)          # apply_opcode_3_warp clips the maps to this range
    assert_bit_equal(got[inside], want[inside], "remap inside the clip range")
