"""bayer_normalize -- reference: normalization.py:4-25."""
import numpy as np
import torch

from . import engine
from ._arrays import as_cuda, give_back, is_numpy


def bayer_normalize(rgbg, chan_black, chan_sat):
    """Normalise a Bayer mosaic from sensor counts to float32: per CFA site
    clip(x - black, 0, sat) / sat with black/sat ordered [TL, TR, BR, BL] (minimum length 4).

    Accepts a NumPy array (result is NumPy, as in the reference) or a CUDA tensor (result stays on the
    device).  uint16/int16 inputs run the CUDA kernel; the arithmetic is the reference's float32 one.
    """
    if len(chan_black) < 4 or len(chan_sat) < 4:
        raise IndexError("bayer_normalize: black/saturation levels need at least 4 entries")
    want_np = is_numpy(rgbg)
    if want_np and rgbg.dtype not in (np.uint16, np.int16):
        if np.issubdtype(rgbg.dtype, np.integer) and rgbg.min() >= 0 and rgbg.max() <= 65535:
            rgbg = rgbg.astype(np.uint16)
        else:
            raise ValueError("bayer_normalize: integer sensor counts in [0, 65535] expected")
    t = as_cuda(rgbg)
    if t.dtype not in (torch.uint16, torch.int16):
        raise ValueError("bayer_normalize: uint16 mosaic expected, got %s" % t.dtype)
    if t.dim() != 2 or t.shape[0] % 2 or t.shape[1] % 2:
        raise ValueError("bayer_normalize: mosaic must be 2-D with even dimensions")
    return give_back(engine.normalize(t, chan_black, chan_sat), want_np)
