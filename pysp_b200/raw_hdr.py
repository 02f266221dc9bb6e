"""HDR fusion -- reference: raw_hdr.py:85-158 (`fuse_exposures_to_raw`, raw space) and raw_hdr.py:7-83
(`fuse_exposures_from_debayer`, camera space)."""
import numpy as np
import torch

from . import engine
from ._arrays import as_cuda, give_back, is_numpy
from .colour import cam_to_rgb_matrix
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


def fuse_exposures_from_debayer(in_exposures, target_ev=None):
    """Fuse demosaiced exposures to linear sRGB HDR (raw_hdr.py:7-83).  Returns (linear sRGB image, int32 buffer of
    contributions per pixel and channel), or None when no exposure is valid.  As in the reference, every valid
    exposure goes through wb_undo()/wb_apply() and keeps the (re-rounded) image that leaves."""
    valid = [e for e in in_exposures if e.is_valid()]
    if len(valid) == 0:
        return None
    if len(valid) > 16:
        raise ValueError("fuse_exposures_from_debayer: at most 16 exposures")
    if target_ev is None:
        target_ev = 0
        for e in valid:
            target_ev += e.current_ev
        target_ev /= len(valid)
    else:
        assert target_ev > 0
    offsets = [2 ** (e.current_ev - target_ev) for e in valid]
    off_max = np.max(offsets)
    brightest = max(i for i, o in enumerate(offsets) if o == off_max)
    bias = [np.float32(1.6 ** (-0.1 * o)) for o in offsets]
    want_np = is_numpy(in_exposures[0].image)
    if any(engine.wb_dtype_flags(e._wb_coeff)[0] for e in valid):
        # with float64 coefficients NumPy runs the reference's wb_undo / wb_apply round trip in float64; the fused kernel
        # implements the float32-coefficient case (the EXIF path, helpers_exif.py:79)
        raise ValueError("fuse_exposures_from_debayer: float32 white-balance coefficients expected")
    wb = np.asarray(valid[0]._wb_coeff, dtype=np.float32)
    for e in valid:
        if not np.array_equal(np.asarray(e._wb_coeff, dtype=np.float32), wb):
            raise ValueError("fuse_exposures_from_debayer: exposures must share their white-balance coefficients")
    imgs, dev = [], None
    for e in valid:
        t = as_cuda(e.image, torch.float32, device=dev)
        dev = t.device
        if not e._wb_applied:                       # wb_undo() is a no-op then; wb_apply() still multiplies
            raise ValueError("fuse_exposures_from_debayer: exposures are expected with white balance applied")
        t = t.contiguous()
        if isinstance(e.image, torch.Tensor) and t.data_ptr() == e.image.data_ptr():
            t = t.clone()                           # the reference assigns new arrays; the caller's tensor is not rewritten
        imgs.append(t)
    m = cam_to_rgb_matrix(in_exposures[0].mat_xyz)
    lin, cnt = engine.fuse_exposures_from_debayer(imgs, wb[:3], float(max(wb)), [e._wb_normalized for e in valid],
                                                  [np.float32(o) for o in offsets], bias, brightest, float(off_max), m)
    for e, t in zip(valid, imgs):
        e.image = give_back(t, is_numpy(e.image))
        e._wb_normalized = False
    return give_back(lin, want_np), give_back(cnt, want_np)
