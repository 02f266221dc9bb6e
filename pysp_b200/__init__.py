"""pysp_b200 -- B200-native implementation of pySP's raw -> linear-sRGB develop path.

Drop-in names (reference module in brackets):
    RawBayerData, RawRggbBayerData, RawBayerDataFromRaw / RawRgbgDataFromRaw   [image.py]
    BayerPattern, RawDemosaicData                                              [base_types/image_base.py]
    QualityDemosaic                                                            [const.py]
    bayer_normalize                                                            [normalization.py]
    debayer_ahd, debayer_eag                                                   [debayer/__init__.py]
    cam_to_lin_srgb, cam_to_rgb_norm, clip_rgb, lin_srgb_to_srgb               [colorize/transform.py]
    fuse_exposures_to_raw, fuse_exposures_from_debayer                         [raw_hdr.py]
    flat_frame_correction, dark_frame_subtraction, bias_frame_subtraction      [raw_correction.py]
    find_erroneous_pixels_threshold                                            [raw_bad_pixel_corr.py]
The compute path is hand-written CUDA (sm_100a) behind a C ABI (include/pysp_b200.h); there is no CPU
fallback -- importing is cheap, but any compute call needs libpysp_b200.so and a CUDA device.
"""
from .const import QualityDemosaic, PatternDemosaic  # noqa: F401


def __getattr__(name):
    # torch is imported lazily so that `import pysp_b200` (e.g. for pysp_b200.build) stays light
    import importlib
    table = {
        "RawBayerData": ".image", "RawRggbBayerData": ".image", "RawBayerDataFromRaw": ".image",
        "RawRgbgDataFromRaw": ".image", "reversible_transform_rggb": ".image",
        "BayerPattern": ".base_types.image_base", "RawDemosaicData": ".base_types.image_base",
        "bayer_normalize": ".normalization", "debayer_ahd": ".debayer", "debayer_eag": ".debayer",
        "cam_to_lin_srgb": ".colorize.transform", "cam_to_rgb_norm": ".colorize.transform",
        "clip_rgb": ".colorize.transform", "lin_srgb_to_srgb": ".colorize.transform",
        "fuse_exposures_to_raw": ".raw_hdr", "fuse_exposures_from_debayer": ".raw_hdr",
        "flat_frame_correction": ".raw_correction", "dark_frame_subtraction": ".raw_correction",
        "bias_frame_subtraction": ".raw_correction", "find_erroneous_pixels_threshold": ".raw_bad_pixel_corr",
        "CameraWhiteBalance": ".wb_cct.cam_wb",
        "MatXyzToCamera": ".wb_cct.helpers_cam_mat",
    }
    if name in table:
        return getattr(importlib.import_module(table[name], __name__), name)
    raise AttributeError("module 'pysp_b200' has no attribute %r" % name)
