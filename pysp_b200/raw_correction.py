"""Calibration-frame corrections on the mosaic -- reference: raw_correction.py:7-63.

`flat_frame_correction` runs on the GPU (csrc/prepost.cuh); the plane means are evaluated in NumPy's own float32
pairwise order, so the corrected mosaic is bit-identical to the reference's.  `dark_frame_subtraction` and
`bias_frame_subtraction` are stubs in the reference (they return a copy of their first argument) and stay stubs.
"""
import copy

import torch

from . import engine
from ._arrays import as_cuda, give_back, is_numpy


def dark_frame_subtraction(raw, dark_frame):
    """Not implemented by the reference (raw_correction.py:7-14): returns a copy of `raw`."""
    return copy.copy(raw)


def bias_frame_subtraction(raw, bias_frame):
    """Not implemented by the reference (raw_correction.py:16-23): returns a copy of `raw`."""
    return copy.copy(raw)


def flat_frame_correction(image, flat, clamp_high=False):
    """Apply flat-frame correction in place to `image.sensor_scaled` (raw_correction.py:25-63).

    Per CFA plane: `chan * mean(flat_chan) / flat_chan`; a division by zero takes the largest finite value of the
    plane, negative results are clamped to zero, `clamp_high` also clamps at 1; a plane whose flat is completely
    black is left untouched.
    """
    want_np = is_numpy(image.sensor_scaled)
    sensor = as_cuda(image.sensor_scaled, torch.float32)
    field = as_cuda(flat.sensor_scaled, torch.float32, device=sensor.device)
    image.sensor_scaled = give_back(engine.flat_frame_correction(sensor, field, clamp_high), want_np)
