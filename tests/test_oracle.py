"""The oracle (oracle/ahd_spec.py) against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Bit-exact: every operation on the path is an IEEE basic op or an integer
table interpolation."""
import numpy as np
import pytest

from conftest import assert_bit_equal, golden, golden_develop_cases
from oracle import ahd_spec as sp
from pysp_b200 import synthetic as syn

WB = syn.wb_multipliers()


@pytest.mark.parametrize("name", golden_develop_cases())
def test_develop_matches_reference(name):
    d = golden(name)
    lin, cam, ex = sp.develop(d["raw"], d["black"], d["white"], WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ,
                              int(d["stages"]), str(d["pattern"]), keep=True)
    assert_bit_equal(sp.normalize(d["raw"], d["black"], d["white"]), d["sensor"], "bayer_normalize")
    assert np.array_equal(ex["cnt_h"], d["cnt_h"]) and np.array_equal(ex["cnt_v"], d["cnt_v"])
    assert np.array_equal(ex["pick_h"], d["pick_h"])          # AHD direction choice
    assert_bit_equal(cam, d["cam"], "camera RGB")
    assert_bit_equal(lin, d["lin"], "linear sRGB")


@pytest.mark.parametrize("stages", [0, 1])
def test_hdr_branch(stages):
    d = golden("hdr48x64_s%d" % stages)
    m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    cam, ex = sp.ahd_demosaic(d["sensor"], WB, m, stages, hdr=True, keep=True)
    assert np.array_equal(ex["pick_h"], d["pick_h"])
    assert_bit_equal(cam, d["cam"], "camera RGB (HDR)")
    assert_bit_equal(sp.to_lin_srgb(cam, m), d["lin"], "linear sRGB (HDR)")


def test_nonfinite_photosites():
    """+inf / -inf / NaN photosites: the reference blends the candidates multiplicatively (debayer/ahd.py:139-145), so a
    non-finite value in either candidate poisons the pixel; cv2's Lab clamps NaN to 0.  Pinned at stages = 0 (the NaN
    ordering of cv2.medianBlur is implementation-defined)."""
    from conftest import assert_bit_equal_nan
    d = golden("nonfinite48x64_s0")
    m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    with np.errstate(invalid="ignore"):
        cam, ex = sp.ahd_demosaic(d["sensor"], syn.wb_multipliers(), m, 0, keep=True)
        assert np.array_equal(ex["pick_h"], d["pick_h"])
        assert_bit_equal_nan(cam, d["cam"], "camera RGB with non-finite photosites")
        assert_bit_equal_nan(sp.to_lin_srgb(cam, m), d["lin"], "linear sRGB with non-finite photosites")
    assert int(np.isnan(d["cam"]).sum()) > 300


def test_fuse_exposures():
    d = golden("fuse5_40x56")
    fused, cnt, lim, tev = sp.fuse_exposures(list(d["brackets"]), list(d["evs"]), WB)
    assert_bit_equal(fused, d["fused"], "fused mosaic")
    assert np.array_equal(cnt, d["count"])
    assert lim == float(d["lim_sat"]) and tev == float(d["target_ev"])
    m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    cam = sp.ahd_demosaic(fused, WB, m, 1, hdr=True)
    assert_bit_equal(cam, d["cam"], "camera RGB of the fused mosaic")
    assert_bit_equal(sp.to_lin_srgb(cam, m), d["lin"], "linear sRGB of the fused mosaic")


def test_gamma():
    d = golden("gamma")
    assert_bit_equal(sp.lin_srgb_to_srgb(d["x"]), d["y"], "lin_srgb_to_srgb")


def test_constants():
    d = golden("kernels")
    assert np.array_equal(np.stack(sp.phase_kernels(False)), d["base_tl"])
    assert np.array_equal(np.stack(sp.phase_kernels(True)), d["base_br"])
    g = d["gauss3"].ravel()
    assert g[0] == sp.GAUSS_K0 and g[1] == sp.GAUSS_K1 and g[2] == sp.GAUSS_K0
    assert [float(v).hex() for v in sp.H5] == ["-0x1.0533160000000p-2", "0x1.0000000000000p-1",
                                               "0x1.0533160000000p-1", "0x1.0000000000000p-1",
                                               "-0x1.0533160000000p-2"]


def test_lab_table_against_cv2():
    """The harvested 33^3 table + integer trilinear restatement reproduces cv2.cvtColor bit for bit
    (incl. out-of-range inputs and the 1/512 quantisation boundaries)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.2, 1.2, size=(1, 400000, 3)).astype(np.float32)
    k = (rng.integers(0, 16385 * 2, size=(1, 100000, 3)).astype(np.float32) / np.float32(32768.0))
    x = np.concatenate([x, k.astype(np.float32)], axis=1)
    assert_bit_equal(sp.lab_cv(x), cv2.cvtColor(x, cv2.COLOR_RGB2LAB), "Lab")


def test_cv2_backend_is_close():
    """The timing backend (same library calls as the reference, optimised OpenCV) agrees with the pinned
    arithmetic up to the reference's own optimised-vs-generic noise."""
    pytest.importorskip("cv2")
    raw = syn.scene(96, 128, 3)
    a, _ = sp.develop(raw, syn.BLACK, syn.WHITE, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, 0)
    b, _ = sp.develop(raw, syn.BLACK, syn.WHITE, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, 0, backend="cv2")
    close = np.abs(a - b) <= 1e-4 * np.maximum(np.abs(a), 1e-3)
    assert close.mean() > 0.995


def test_gate2_default_mode_reference():
    """Parity gate (ii): the oracle (= the reference in generic mode, bit for bit) against the reference in its default
    mode on a 1.5 MP frame: tolerance outside the neighbourhood of flips, flips equal to the reference's own noise floor."""
    from conftest import gate2_check
    d = golden("gate2_1024x1536_s1")
    raw = syn.scene(int(d["H"]), int(d["W"]), int(d["seed"]))
    lin, cam, ex = sp.develop(raw, syn.BLACK, syn.WHITE, syn.wb_multipliers(), syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ,
                              int(d["stages"]), keep=True)
    flips, own = gate2_check(lin, ex["pick_h"], lin, "oracle")
    assert flips == own


@pytest.mark.parametrize("name", ["fast_rand8x8", "fast_scene34x50", "fast_scene64x96_GBRG", "fast_flat20x28", "fast_rand66x130"])
def test_fast_quality_matches_reference(name):
    d = golden(name)
    lin, cam = sp.develop_fast(d["raw"], d["black"], d["white"], WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, str(d["pattern"]))
    assert_bit_equal(cam, d["cam"], "Fast camera RGB")
    assert_bit_equal(lin, d["lin"], "Fast linear sRGB")


def test_matrix_accumulation_order_pins():
    """colorize/transform.py:52-53: the float64 3x3 is an FMA chain (OpenBLAS dgemm).  The fixture holds inputs on
    which fused and unfused accumulation round to different float32 values, with the unmodified reference's
    outputs (tests/golden/make_fma_pins.py): the oracle must equal the reference, i.e. the fused variant."""
    d = golden("dot_fma_pins")
    m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    assert np.array_equal(m, d["m"])
    assert_bit_equal(sp.to_lin_srgb(d["x"], m), d["y"], "camera -> linear sRGB on the accumulation-order pins")
    n = d["x"].shape[1]
    got = d["y"][0, np.arange(n), d["row"]]
    assert n >= 20 and np.array_equal(got, d["fused"]) and not np.any(got == d["unfused"])
