"""DNG WarpRectilinear coordinate tables -- reference: dng_warp_corr/dng_warp_rectilinear_coords.pyx:67-95 (a Cython /
OpenMP extension there; a CUDA kernel behind `pysp_warp_rectilinear_table` here).  Same names and argument order."""
import numpy as np
import torch

from .. import engine
from .._arrays import give_back


def compute_remapping_table(kr0, kr1, kr2, kr3, kt0, kt1, width, height, cam_center_norm_x, cam_center_norm_y, scale,
                            device=None):
    """float32 [height, width, 2] map (x', y') for cv2.remap.  Returns a NumPy array (as the reference) unless `device`
    names a CUDA device, in which case the table stays there."""
    t = engine.warp_table(int(height), int(width), (kr0, kr1, kr2, kr3, kt0, kt1), (cam_center_norm_x, cam_center_norm_y),
                          scale, device=device)
    return give_back(t, device is None)


def compute_offset_remapping_table(seed, kr0, kr1, kr2, kr3, kt0, kt1, width, height, cam_center_norm_x, cam_center_norm_y,
                                   scale):
    """Same, starting from the prior mapping `seed` [height, width, 2] (NumPy or CUDA tensor; the result is of that kind)."""
    want_np = isinstance(seed, np.ndarray)
    s = engine.to_device(np.ascontiguousarray(seed, dtype=np.float32) if want_np else seed, dtype=torch.float32)
    t = engine.warp_table(int(height), int(width), (kr0, kr1, kr2, kr3, kt0, kt1), (cam_center_norm_x, cam_center_norm_y),
                          scale, seed=s)
    return give_back(t, want_np)
