"""Image containers of the develop path -- reference: base_types/image_base.py:13-124.

Same attribute names and state machine as the reference (`image`, `_wb_coeff`, `_wb_applied`,
`_wb_normalized`, `mat_xyz`, `current_ev`; `sensor_scaled`, `cam_wb`, `lim_sat`, HDR flag), but `image` /
`sensor_scaled` may be NumPy arrays or CUDA tensors; results keep the caller's kind.
"""
from enum import IntEnum, auto

import numpy as np
import torch

from .. import engine
from .._arrays import as_cuda, give_back, is_numpy
from ..colorize.transform import cam_to_lin_srgb
from ..const import QualityDemosaic  # noqa: F401
from ..wb_cct.helpers_cam_mat import MatXyzToCamera


class BayerPattern(IntEnum):
    Rggb = auto()
    Bggr = auto()
    Grbg = auto()
    Gbrg = auto()


class RawDemosaicData:
    """RGB pixel data after demosaicing: [H, W, 3] float32 camera RGB, white balance applied."""

    def __init__(self, image, wb_coeff, wb_norm=False):
        self.image = image
        self._wb_coeff = wb_coeff
        self._wb_applied = True
        self._wb_normalized = wb_norm
        self.mat_xyz = None
        self.current_ev = np.inf

    def is_valid(self):
        return (self.image is not None and self._wb_coeff is not None
                and type(self.mat_xyz) != type(MatXyzToCamera) and self.current_ev != np.inf)

    def wb_apply(self):
        """Multiply the white-balance coefficients in, if they are not applied already (float32, image_base.py:45-49)."""
        if not self._wb_applied:
            want_np = is_numpy(self.image)
            img = as_cuda(self.image, torch.float32)
            self.image = give_back(engine.wb_scale(img, self._wb_coeff, engine.WB_APPLY), want_np)
            self._wb_applied = True

    def wb_undo(self):
        """Return to pure camera space: float64 division, after removing the normalisation (image_base.py:52-60)."""
        if self._wb_applied:
            want_np = is_numpy(self.image)
            img = as_cuda(self.image, torch.float32)
            # the coefficients keep their dtype (float32 array, float64 array or Python list): NumPy's promotion decides in
            # the reference whether products and quotients are float32 or float64 (engine.wb_dtype_flags)
            self.image = give_back(engine.wb_scale(img, self._wb_coeff, engine.WB_UNDO, normalized=self._wb_normalized,
                                                   max_wb=float(max(self._wb_coeff))), want_np)
            self._wb_applied = False
            self._wb_normalized = False

    def to_lin_srgb(self):
        self.wb_apply()
        return cam_to_lin_srgb(self.image, self.mat_xyz)


class RawCameraData_BaseType:
    def __init__(self):
        self.sensor_scaled = None
        self.cam_wb = None
        self.current_ev = np.inf
        self.lim_sat = 1.0
        self.__is_hdr = False

    def set_hdr(self, is_hdr):
        self.__is_hdr = is_hdr

    def get_hdr(self):
        return self.__is_hdr

    def demosaic(self, quality, postprocess_steps=1):
        return None


class RawBayerData_BaseType(RawCameraData_BaseType):
    def __init__(self):
        super().__init__()
        self.sensor_pattern = None

    def to_rggb(self):
        return None


class RawRggbBayerData_BaseType(RawCameraData_BaseType):
    def __init__(self, sensor_scaled, cam_wb, shot_ev, lim_sat, source_pattern=BayerPattern.Rggb):
        super().__init__()
        self.sensor_scaled = sensor_scaled
        self.cam_wb = cam_wb
        self.current_ev = shot_ev
        self.lim_sat = lim_sat
        self.source_pattern = source_pattern
