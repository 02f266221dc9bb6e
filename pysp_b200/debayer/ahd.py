"""AHD demosaic entry point -- reference: debayer/ahd.py:14-170 (`debayer_ahd`).

The whole function body of the reference (white balance, directional greens, Gaussian 4-phase chroma
upsample, cv2-Lab homogeneity maps incl. the Cython `build_map`, vote, select, median post-process) is
one CUDA kernel chain behind `pysp_develop` (pysp_b200/csrc/).
"""
import torch

from .. import engine
from .._arrays import as_cuda, give_back, is_numpy
from ..base_types.image_base import RawDemosaicData
from ..colour import cam_to_rgb_matrix


def debayer(image, postprocess_stages=1):
    """Demosaic an RGGB float32 mosaic container with AHD; returns RawDemosaicData (camera RGB)."""
    want_np = is_numpy(image.sensor_scaled)
    sensor = as_cuda(image.sensor_scaled, torch.float32)
    wb = image.cam_wb.get_reciprocal_multipliers()
    mat = image.cam_wb.get_matrix()
    cam = engine.develop(sensor, wb, cam_to_rgb_matrix(mat), stages=max(int(postprocess_stages), 0),
                         hdr=image.get_hdr(), out="cam")
    out = RawDemosaicData(give_back(cam, want_np), wb, wb_norm=False)
    out.mat_xyz = mat
    out.current_ev = image.current_ev
    return out
