"""Colour transforms of the develop path -- reference: colorize/transform.py:6-53, 76-99.

The 3x3 is built on the host in float64 with the reference's formula (pysp_b200/colour.py) and applied
on the device in float64 per pixel, rounded once to float32, like `np.dot(f32, f64).astype(f32)`.
"""
import torch

from .. import engine
from .._arrays import as_cuda, give_back, is_numpy
from ..colour import LinRgbColorspace, cam_to_rgb_matrix


def clip_rgb(rgb):
    """Clip an RGB image to [0,1] (transform.py:6-19)."""
    want_np = is_numpy(rgb)
    return give_back(engine.wb_scale(as_cuda(rgb, torch.float32), None, engine.CLIP01), want_np)


def cam_to_rgb_norm(rgb, cam_xyz_matrix, destination_colorspace, clip_highlights=True):
    """Camera RGB -> detinted linear RGB of `destination_colorspace` (transform.py:21-53)."""
    want_np = is_numpy(rgb)
    m = cam_to_rgb_matrix(cam_xyz_matrix, destination_colorspace)
    out = engine.cam_to_rgb(as_cuda(rgb, torch.float32), m, clip=clip_highlights)
    return give_back(out, want_np)


def cam_to_lin_srgb(rgb, cam_xyz_matrix, clip_highlights=True):
    """Camera RGB -> linear sRGB (transform.py:76-87)."""
    return cam_to_rgb_norm(rgb, cam_xyz_matrix, LinRgbColorspace.REC709, clip_highlights)


def lin_srgb_to_srgb(rgb):
    """Apply the sRGB transfer curve in float32 (transform.py:89-99)."""
    want_np = is_numpy(rgb)
    return give_back(engine.srgb_gamma(as_cuda(rgb, torch.float32)), want_np)
