"""Hot-pixel detection on the mosaic -- reference: raw_bad_pixel_corr.py:30-65 (`find_erroneous_pixels_threshold`).

The quantile-based detector, the cross-image consensus and the cv2.inpaint repair (raw_bad_pixel_corr.py:67-150) are
outside the B200 path.
"""
import torch

from . import engine
from ._arrays import as_cuda, give_back, is_numpy


def find_erroneous_pixels_threshold(image, min_delta=0.025, min_neighbour_count=5):
    """Per colour plane (R, G1, B, G2): boolean mask of photosites that exceed more than `min_neighbour_count` of
    their eight same-colour neighbours by more than `min_delta`."""
    want_np = is_numpy(image.sensor_scaled)
    masks = engine.find_hot_pixels_threshold(as_cuda(image.sensor_scaled, torch.float32), min_delta, min_neighbour_count)
    return [give_back(masks[k], want_np) for k in range(4)]
