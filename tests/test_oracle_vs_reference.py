"""Live check of the oracle against the unmodified reference, when it is mounted (build container only;
skipped on the GPU box).  Fresh random inputs each size, cv2 in generic mode."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from oracle import ahd_spec as sp
from oracle import ref_harness as rh
from pysp_b200 import synthetic as syn

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference not mounted")


@pytest.mark.parametrize("shape,stages,seed", [((24, 36), 1, 11), ((52, 40), 2, 12), ((96, 160), 1, 13)])
def test_live_reference(shape, stages, seed):
    rh.load()
    rh.pin_numerics(True)
    try:
        from pySP.normalization import bayer_normalize
        from pySP.const import QualityDemosaic
        raw = syn.scene(shape[0], shape[1], seed, noise=45.0)
        wb = rh.StubWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
        img = rh.make_rggb_container(bayer_normalize(raw, list(syn.BLACK), list(syn.WHITE)), wb)
        dem = img.demosaic(QualityDemosaic.Best, stages)
        lin, cam = sp.develop(raw, syn.BLACK, syn.WHITE, syn.wb_multipliers(), syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, stages)
        assert_bit_equal(cam, dem.image, "camera RGB")
        assert_bit_equal(lin, dem.to_lin_srgb(), "linear sRGB")
    finally:
        rh.pin_numerics(False)
