"""Raw-space HDR fusion -- reference: raw_hdr.py:85-158 (`fuse_exposures_to_raw`)."""
import numpy as np
import torch

from . import engine
from ._arrays import as_cuda, give_back, is_numpy
from .image import RawRggbBayerData


def fusion_constants(evs, wb, target_ev=None):
    """Host scalars of the fusion: target EV (mean by default), per-bracket EV offsets 2**(ev-target)
    and the three per-channel bias values 1.6**(-0.1*|offset*wb[c]|) evaluated with NumPy in float32
    exactly as the reference's array expression does (raw_hdr.py:111-136)."""
    evs = [float(e) for e in evs]
    if target_ev is None:
        target_ev = 0
        for e in evs:
            target_ev += e
        target_ev /= len(evs)
    else:
        assert target_ev > 0
    offsets = [2 ** (e - target_ev) for e in evs]
    wbc = np.asarray(wb, dtype=np.float32)[:3]
    bias = np.stack([1.6 ** (-0.1 * np.abs(off * wbc)) for off in offsets]).astype(np.float32)
    return target_ev, offsets, bias


def fuse_exposures_to_raw(in_exposures, target_ev=None):
    """Fuse RGGB exposures into one HDR mosaic, accumulating in list order in float32.
    Returns (HDR RawRggbBayerData with set_hdr(True), int32 contribution-count buffer), or None for an
    empty list."""
    if len(in_exposures) == 0:
        return None
    if len(in_exposures) > 16:
        raise ValueError("fuse_exposures_to_raw: at most 16 brackets")
    want_np = is_numpy(in_exposures[0].sensor_scaled)
    wb = in_exposures[0].cam_wb.get_reciprocal_multipliers()
    target_ev, offsets, bias = fusion_constants([e.current_ev for e in in_exposures], wb, target_ev)
    dev = None
    planes = []
    for e in in_exposures:
        t = as_cuda(e.sensor_scaled, torch.float32, device=dev)
        dev = t.device
        planes.append(t)
    fused, count = engine.fuse_exposures(planes, offsets, bias, int(np.argmax(offsets)))
    hdr = RawRggbBayerData(give_back(fused, want_np), in_exposures[0].cam_wb.copy(), target_ev, max(offsets),
                           in_exposures[0].source_pattern)
    hdr.set_hdr(True)
    return hdr, give_back(count, want_np)
