"""Raw image containers and the demosaic dispatch -- reference: image.py:143-197.

File ingest (`RawBayerDataFromRaw.__init__`, image.py:199-307: rawpy/exifread/tifftools decode) is out
of scope of the B200 path; `RawBayerDataFromRaw.from_mosaic` builds the same
container from an already decoded mosaic and keeps the 16-bit counts on the device so that
normalisation is fused into the develop kernel.
"""
import numpy as np
import torch

from . import engine
from ._arrays import as_cuda, give_back, is_numpy
from .base_types.image_base import (BayerPattern, RawBayerData_BaseType, RawDemosaicData,
                                    RawRggbBayerData_BaseType)
from .colour import cam_to_rgb_matrix
from .const import QualityDemosaic
from .debayer import debayer_ahd, debayer_eag
from .normalization import bayer_normalize

_PATTERN_NAME = {BayerPattern.Rggb: "RGGB", BayerPattern.Bggr: "BGGR", BayerPattern.Grbg: "GRBG",
                 BayerPattern.Gbrg: "GBRG"}


def reversible_transform_rggb(sensor_data, bayer_pattern):
    """Self-inverse flip taking a mosaic (or an image) between its stored CFA layout and RGGB."""
    if bayer_pattern == BayerPattern.Rggb:
        return sensor_data
    axes = {BayerPattern.Bggr: (0, 1), BayerPattern.Gbrg: (1,), BayerPattern.Grbg: (0,)}.get(bayer_pattern)
    if axes is None:
        raise NotImplementedError(str(bayer_pattern) + " not implemented!")
    if isinstance(sensor_data, np.ndarray):
        return np.flip(sensor_data, axis=axes)
    return torch.flip(sensor_data, dims=axes)


def _demosaic_dispatch(quality):
    """image.py:169-176: Best -> debayer_ahd, Fast -> debayer_eag; Draft (cv2.resize based) is not on the B200 path."""
    if quality == QualityDemosaic.Best:
        return "best"
    if quality == QualityDemosaic.Fast:
        return "fast"
    if quality == QualityDemosaic.Draft:
        raise NotImplementedError("Quality mode not implemented on the B200 path: %s" % str(quality))
    raise NotImplementedError("Quality mode not implemented: %s" % str(quality))


class RawRggbBayerData(RawRggbBayerData_BaseType):
    def demosaic(self, quality, postprocess_steps=1):
        """Demosaic to a new RawDemosaicData; the source is not modified."""
        if _demosaic_dispatch(quality) == "fast":
            debayered = debayer_eag(self)
        else:
            debayered = debayer_ahd(self, postprocess_stages=postprocess_steps)
        debayered.image = reversible_transform_rggb(debayered.image, self.source_pattern)
        return debayered

    debayer = demosaic          # README spelling


class RawBayerData(RawBayerData_BaseType):
    """Bayer mosaic in its stored CFA layout."""

    def __init__(self):
        super().__init__()
        self._counts = None     # optional: un-normalised uint16 counts kept on the device
        self._counts_numpy = False
        self._levels = None

    def to_rggb(self):
        rggb = reversible_transform_rggb(self._sensor(), self.sensor_pattern)
        out = RawRggbBayerData(rggb, self.cam_wb.copy(), self.current_ev, self.lim_sat, self.sensor_pattern)
        return out              # (the reference drops the HDR flag here too, image.py:191-193)

    def _sensor(self):
        if self.sensor_scaled is None and self._counts is not None:
            self.sensor_scaled = bayer_normalize(self._counts, *self._levels)
        return self.sensor_scaled

    def demosaic(self, quality, postprocess_steps=1):
        """to_rggb().demosaic(...), as one kernel chain: the CFA flip is index math on load and store, and
        16-bit counts (from_mosaic) are normalised inside the kernel."""
        q = _demosaic_dispatch(quality)
        wb = self.cam_wb.get_reciprocal_multipliers()
        mat = self.cam_wb.get_matrix()
        stages = max(int(postprocess_steps), 0)
        pattern = _PATTERN_NAME.get(self.sensor_pattern)
        if pattern is None:
            raise NotImplementedError(str(self.sensor_pattern) + " not implemented!")
        if self._counts is not None:
            want_np = self._counts_numpy
            cam = engine.develop(as_cuda(self._counts), wb, cam_to_rgb_matrix(mat), stages=stages, pattern=pattern,
                                 black=self._levels[0], white=self._levels[1], out="cam", quality=q)
        else:
            want_np = is_numpy(self.sensor_scaled)
            cam = engine.develop(as_cuda(self.sensor_scaled, torch.float32), wb, cam_to_rgb_matrix(mat),
                                 stages=stages, pattern=pattern, out="cam", quality=q)
        out = RawDemosaicData(give_back(cam, want_np), wb, wb_norm=False)
        out.mat_xyz = mat
        out.current_ev = self.current_ev
        return out

    debayer = demosaic          # README spelling

    def develop(self, postprocess_steps=1, srgb_gamma=False, half=False, quality=QualityDemosaic.Best):
        """demosaic(QualityDemosaic.Best, n).to_lin_srgb() [-> lin_srgb_to_srgb] as ONE fused chain: the
        clip + float64 camera->linear-sRGB matrix (+ gamma) run in the last kernel's epilogue, so the
        linear-sRGB image is the only thing written to HBM."""
        wb = self.cam_wb.get_reciprocal_multipliers()
        m = cam_to_rgb_matrix(self.cam_wb.get_matrix())
        pattern = _PATTERN_NAME.get(self.sensor_pattern)
        if pattern is None:
            raise NotImplementedError(str(self.sensor_pattern) + " not implemented!")
        kind = "lin_f16" if half else "lin"
        stages = max(int(postprocess_steps), 0)
        q = _demosaic_dispatch(quality)
        if self._counts is not None:
            want_np = self._counts_numpy
            out = engine.develop(as_cuda(self._counts), wb, m, stages=stages, pattern=pattern, black=self._levels[0],
                                 white=self._levels[1], gamma=srgb_gamma, out=kind, quality=q)
        else:
            want_np = is_numpy(self.sensor_scaled)
            out = engine.develop(as_cuda(self.sensor_scaled, torch.float32), wb, m, stages=stages, pattern=pattern,
                                 gamma=srgb_gamma, out=kind, quality=q)
        return give_back(out, want_np)


class RawBayerDataFromRaw(RawBayerData):
    """Bayer data from a raw file (reference: image.py:199-307).

    Decoding raw files needs rawpy / exifread / tifftools, which are outside the B200 develop path; use
    `from_mosaic` with a decoded mosaic and its metadata."""

    def __init__(self, filename_or_data=None):
        super().__init__()
        if filename_or_data is not None:
            raise NotImplementedError(
                "raw-file ingest (rawpy/exifread/tifftools, reference image.py:199-307) is outside the B200 "
                "develop path; decode the file and call RawBayerDataFromRaw.from_mosaic(...)")

    @classmethod
    def from_mosaic(cls, mosaic, black_level_per_channel, white_level_per_channel, pattern, cam_wb, ev=0.0,
                    keep_counts=True):
        """mosaic: uint16 [H,W] (NumPy or CUDA tensor) as rawpy's `raw_image`; black/white in rawpy's
        per-channel order [TL,TR,BR,BL] (image.py:227-229)."""
        self = cls()
        self.sensor_pattern = pattern if isinstance(pattern, BayerPattern) else BayerPattern[str(pattern).capitalize()]
        self.cam_wb = cam_wb
        self.current_ev = ev
        levels = (list(black_level_per_channel), list(white_level_per_channel))
        if keep_counts:
            self._counts_numpy = is_numpy(mosaic)
            self._counts = engine.to_device(mosaic, pad_pitch=True)
            self._levels = levels
        else:
            self.sensor_scaled = bayer_normalize(mosaic, *levels)
        return self


RawRgbgDataFromRaw = RawBayerDataFromRaw     # README spelling (README.md:57-62)
