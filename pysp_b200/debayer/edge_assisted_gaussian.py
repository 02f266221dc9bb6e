"""Edge-assisted Gaussian demosaic entry point -- reference: debayer/edge_assisted_gaussian.py:188-201 (`debayer_eag`).

Gradient-weighted bilinear green (l.10-124), green high-pass and the photosite-aware Gaussian upsample of the
colour differences (l.126-158) run as one CUDA kernel (pysp_b200/csrc/eag.cuh) in the same tile pipeline as AHD.
"""
import torch

from .. import engine
from .._arrays import as_cuda, give_back, is_numpy
from ..base_types.image_base import RawDemosaicData
from ..colour import cam_to_rgb_matrix


def debayer(image):
    """Demosaic an RGGB float32 mosaic container (QualityDemosaic.Fast); returns RawDemosaicData (camera RGB)."""
    want_np = is_numpy(image.sensor_scaled)
    sensor = as_cuda(image.sensor_scaled, torch.float32)
    wb = image.cam_wb.get_reciprocal_multipliers()
    mat = image.cam_wb.get_matrix()
    cam = engine.develop(sensor, wb, cam_to_rgb_matrix(mat), out="cam", quality="fast")
    out = RawDemosaicData(give_back(cam, want_np), wb, wb_norm=False)
    out.mat_xyz = mat
    out.current_ev = image.current_ev
    return out
