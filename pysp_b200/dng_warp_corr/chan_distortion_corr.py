"""DNG OpcodeList3 WarpRectilinear -- reference: dng_warp_corr/chan_distortion_corr.py.

`apply_opcode_3_warp` parses the opcode block on the host exactly as the reference does (chan_distortion_corr.py:100-128,
53-83) and runs coordinate model + Lanczos-4 resampling of all planes as one CUDA kernel (the reference builds a table
with its Cython extension and calls cv2.remap per plane, chan_distortion_corr.py:85-97)."""
from io import BytesIO
from struct import unpack

import numpy as np
import torch

from .. import engine
from .._arrays import as_cuda, is_numpy


def stack_warp_prior(demosaiced_image, remap_r, remap_g, remap_b):
    """Combine per-channel cv2.remap-style maps (map[y, x] = new_x, new_y) into one prior [H, W, 3, 2]; a missing channel gets
    the identity map (chan_distortion_corr.py:10-41)."""
    h, w = demosaiced_image.shape[0], demosaiced_image.shape[1]
    if remap_r is None or remap_g is None or remap_b is None:
        ident = np.zeros((h, w, 2), dtype=np.float32)
        ident[:, :, 0] = np.arange(w, dtype=np.float32)[None, :]
        ident[:, :, 1] = np.arange(h, dtype=np.float32)[:, None]
        remap_r = ident if remap_r is None else remap_r
        remap_g = ident if remap_g is None else remap_g
        remap_b = ident if remap_b is None else remap_b
    maps = [m.cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m) for m in (remap_r, remap_g, remap_b)]
    return np.stack(maps, axis=2)


def apply_opcode_3_warp(demosaiced_image, ifd_opcode_3_data, scale=1.0, prior=None):
    """Apply the WarpRectilinear operators of an OpcodeList3 block in place and in order; other opcodes are skipped with the
    reference's message.  `demosaiced_image`: float32 [H, W, C], NumPy array (rewritten in place, like the reference) or
    CUDA tensor (rewritten in place on the device)."""
    assert prior is None or tuple(prior.shape) == (demosaiced_image.shape[0], demosaiced_image.shape[1],
                                                   demosaiced_image.shape[2], 2)
    want_np = is_numpy(demosaiced_image)
    state = {"img": None}

    def device_image():
        if state["img"] is None:
            state["img"] = as_cuda(demosaiced_image, torch.float32)
        return state["img"]

    def opcode_warp_rectilinear(data):
        if len(data) < 4:
            return False
        count_planes = int.from_bytes(data[:4], byteorder="big")
        if len(data) != 4 + (6 * 8 * count_planes) + 16 or count_planes != demosaiced_image.shape[2]:
            return False
        coefficients = [unpack(">6d", data[4 + (6 * 8 * i):4 + (6 * 8 * (i + 1))]) for i in range(count_planes)]
        centre = unpack(">2d", data[4 + (6 * 8 * count_planes):4 + (6 * 8 * count_planes) + 16])
        img = device_image()
        p = None if prior is None else as_cuda(prior, torch.float32, device=img.device)
        state["img"] = engine.warp_rectilinear(img, coefficients, centre, scale, p)
        return True

    data = bytes(ifd_opcode_3_data)
    count_opcodes = int.from_bytes(data[:4], byteorder="big")
    offset = 4
    for _ in range(count_opcodes):
        opcode_id = int.from_bytes(data[offset:offset + 4], byteorder="big")
        opcode_var_len = int.from_bytes(data[offset + 12:offset + 16], byteorder="big")
        offset += 16
        if opcode_id == 1:
            opcode_warp_rectilinear(data[offset:offset + opcode_var_len])
        else:
            print("Unimplemented opcode %d" % opcode_id)
        offset += opcode_var_len
    if state["img"] is not None:
        if want_np:
            demosaiced_image[...] = state["img"].cpu().numpy()
        elif state["img"].data_ptr() != demosaiced_image.data_ptr():
            demosaiced_image.copy_(state["img"])


def get_opcode_3_block(filename_or_data):
    """OpcodeList3 data block of a DNG file (chan_distortion_corr.py:130-145).  Needs `tifftools` (file parsing is outside the
    B200 path); returns None when the file cannot be read, as the reference does."""
    try:
        import tifftools
    except ImportError:
        return None
    if type(filename_or_data) == bytes:
        filename_or_data = BytesIO(filename_or_data)
    try:
        info = tifftools.read_tiff(filename_or_data)
        return info["ifds"][0]["tags"][tifftools.Tag.SubIFD.value]["ifds"][0][0]["tags"][51022]["data"]
    except Exception:
        return None
