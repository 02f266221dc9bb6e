"""Host-side logic of the drop-in layer (no GPU): colour matrices, fusion constants, containers, error
behaviour, band planning."""
import numpy as np
import pytest

from oracle import ahd_spec as sp
from pysp_b200 import colour, parallel
from pysp_b200 import synthetic as syn
from pysp_b200.wb_cct import CameraWhiteBalance, MatXyzToCamera


def test_colour_matrix_is_the_reference_formula(colour_setup):
    m = colour.cam_to_rgb_matrix(colour_setup.get_matrix())
    assert m.dtype == np.float64
    assert np.array_equal(m, sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ))
    # detinted: camera neutral maps to equal RGB
    assert np.allclose(m @ np.ones(3), np.ones(3), atol=1e-12)
    # Rec.709 -> XYZ at D65 (Lindbloom): Y row
    assert np.allclose(colour.LinRgbColorspace.REC709.mat_to_xyz()[1], [0.2126, 0.7152, 0.0722], atol=2e-4)
    a = colour.bradford_adapt_matrix(colour.xy_to_XYZ((0.31272, 0.32903)), colour.xy_to_XYZ((0.34567, 0.35850)))
    assert np.allclose(a @ colour.xy_to_XYZ((0.31272, 0.32903)), colour.xy_to_XYZ((0.34567, 0.35850)))


def test_white_balance_accessors(colour_setup):
    wb = colour_setup.get_reciprocal_multipliers()
    assert wb.dtype == np.float32 and np.array_equal(wb, syn.wb_multipliers())
    c = colour_setup.copy()
    assert np.array_equal(c.get_reciprocal_multipliers(), wb)
    mat = colour_setup.get_matrix()
    assert isinstance(mat, MatXyzToCamera) and not mat.mat.flags.writeable
    assert np.allclose(mat.interpolate(mat, 0.3), mat.mat)


def test_fusion_constants_match_oracle():
    from pysp_b200.raw_hdr import fusion_constants
    from conftest import golden
    d = golden("fuse5_40x56")
    tev, offs, bias = fusion_constants(list(d["evs"]), syn.wb_multipliers())
    assert tev == float(d["target_ev"]) and max(offs) == float(d["lim_sat"])
    assert bias.dtype == np.float32 and bias.shape == (5, 3)
    # same float32 values as the reference's full-array expression
    wbm = syn.wb_multipliers()
    for k, off in enumerate(offs):
        full = 1.6 ** (-0.1 * np.abs(off * np.array([wbm[0], wbm[1], wbm[2]], dtype=np.float32)))
        assert np.array_equal(bias[k], full.astype(np.float32))


def test_quality_dispatch_errors(colour_setup):
    from pysp_b200 import QualityDemosaic, RawBayerData, RawRggbBayerData, BayerPattern
    img = RawRggbBayerData(np.zeros((8, 8), dtype=np.float32), colour_setup, 10.0, 1.0)
    with pytest.raises(NotImplementedError):
        img.demosaic("nonsense")
    with pytest.raises(NotImplementedError):
        img.demosaic(QualityDemosaic.Draft)
    raw = RawBayerData()
    raw.sensor_scaled = np.zeros((8, 8), dtype=np.float32)
    raw.sensor_pattern = BayerPattern.Rggb
    raw.cam_wb = colour_setup
    with pytest.raises(NotImplementedError):
        raw.demosaic(None)


def test_readme_aliases():
    import pysp_b200
    assert pysp_b200.RawRgbgDataFromRaw is pysp_b200.RawBayerDataFromRaw
    assert pysp_b200.RawBayerData.debayer is pysp_b200.RawBayerData.demosaic
    with pytest.raises(NotImplementedError):
        pysp_b200.RawBayerDataFromRaw("some_file.dng")


def test_reversible_transform():
    from pysp_b200 import reversible_transform_rggb, BayerPattern
    a = np.arange(24).reshape(4, 6)
    assert np.array_equal(reversible_transform_rggb(a, BayerPattern.Bggr), np.rot90(a, 2))
    assert np.array_equal(reversible_transform_rggb(a, BayerPattern.Gbrg), np.flip(a, 1))
    assert np.array_equal(reversible_transform_rggb(a, BayerPattern.Grbg), np.flip(a, 0))
    with pytest.raises(NotImplementedError):
        reversible_transform_rggb(a, 99)


def test_demosaic_data_state_machine(colour_setup):
    from pysp_b200 import RawDemosaicData
    d = RawDemosaicData(np.ones((2, 2, 3), dtype=np.float32), colour_setup.get_reciprocal_multipliers())
    assert d._wb_applied and not d._wb_normalized and not d.is_valid()
    d.mat_xyz = colour_setup.get_matrix()
    d.current_ev = 9.0
    assert d.is_valid()


@pytest.mark.parametrize("height,world", [(4000, 8), (8660, 8), (64, 3), (10, 4), (6, 8)])
def test_band_planning(height, world):
    prev = 0
    for r in range(world):
        b, e = parallel.band_rows(height, world, r)
        assert b == prev and b % 2 == 0 and e % 2 == 0 and e >= b
        prev = e
        bb, ee, hb, he = parallel.band_with_halo(height, world, r, 1)
        assert hb == max(0, b - 10) and he == min(height, e + 10)
    assert prev == height
    assert parallel.frames_for_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum([parallel.frames_for_rank(256, r, 8) for r in range(8)], [])) == list(range(256))


def test_pySP_alias_package():
    """`pysp_b200.compat.install_as_pySP()` makes the reference's absolute imports resolve against this package
    (README.md:32, base_types/image_base.py:7-10) -- the same module objects, opt-in, never shadowing a real pySP."""
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import pysp_b200.compat as c\n"
        "c.install_as_pySP()\n"
        "from pySP.colorize.transform import cam_to_lin_srgb, lin_srgb_to_srgb\n"
        "from pySP.base_types.image_base import RawDemosaicData, BayerPattern\n"
        "from pySP.image import RawBayerDataFromRaw, RawRggbBayerData\n"
        "from pySP.const import QualityDemosaic\n"
        "from pySP.debayer import debayer_ahd, debayer_eag\n"
        "from pySP.normalization import bayer_normalize\n"
        "from pySP.raw_hdr import fuse_exposures_to_raw\n"
        "from pySP.dng_warp_corr.chan_distortion_corr import apply_opcode_3_warp\n"
        "import pySP, pysp_b200, pysp_b200.colorize.transform as t\n"
        "assert cam_to_lin_srgb is t.cam_to_lin_srgb and sys.modules['pySP.colorize.transform'] is t\n"
        "assert pySP is pysp_b200\n"
        "c.uninstall()\n"
        "assert 'pySP' not in sys.modules\n"
        "print('ok')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
