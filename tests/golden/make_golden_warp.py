"""Golden fixtures of the DNG WarpRectilinear step from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and cv2):   python tests/golden/make_golden_warp.py
The reference's second native component, dng_warp_corr/dng_warp_rectilinear_coords.pyx, is compiled by
oracle/build_ref.py (gcc, -O2 -fopenmp -ffp-contract=off) into oracle/_ref/; `apply_opcode_3_warp`
(dng_warp_corr/chan_distortion_corr.py:27-128) is imported from the reference as is and fed a synthetic OpcodeList3
block (one WarpRectilinear opcode, three planes).  OpenCV runs in its generic code paths.
"""
import importlib.machinery
import importlib.util
import os
import struct
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import build_ref, ref_harness as rh  # noqa: E402
from pysp_b200 import synthetic as syn  # noqa: E402


def load_warp_module():
    rh.load()
    name = "pySP.dng_warp_corr.dng_warp_rectilinear_coords"
    so = build_ref.build(rh.REFERENCE_ROOT, module=("dng_warp_corr", "dng_warp_rectilinear_coords"))
    loader = importlib.machinery.ExtensionFileLoader(name, so)
    spec = importlib.util.spec_from_file_location(name, so, loader=loader)
    ext = importlib.util.module_from_spec(spec)
    loader.exec_module(ext)
    sys.modules[name] = ext
    import importlib as il
    return ext, il.import_module("pySP.dng_warp_corr.chan_distortion_corr")


def opcode_block(coeffs, centre):
    """OpcodeList3 data: count, then (id = 1 WarpRectilinear, version, flags, length, payload) big-endian."""
    payload = struct.pack(">I", len(coeffs)) + b"".join(struct.pack(">6d", *c) for c in coeffs) + struct.pack(">2d", *centre)
    return struct.pack(">I", 1) + struct.pack(">IIII", 1, 0x01030000, 0, len(payload)) + payload


COEFFS = [(1.0012, -0.0321, 0.0104, -0.0023, 0.0007, -0.0004),
          (0.9991, -0.0298, 0.0088, -0.0017, 0.0005, -0.0006),
          (1.0005, -0.0342, 0.0121, -0.0031, 0.0009, -0.0002)]
CENTRE = (0.4987, 0.5021)


def main():
    import cv2
    rh.pin_numerics(True)
    ext, mod = load_warp_module()
    meta = dict(cv2=cv2.__version__, numpy=np.__version__)
    rng = np.random.default_rng(11)
    for name, (H, W), scale in (("warp_96x128", (96, 128), 1.0), ("warp_70x50_scale", (70, 50), 0.6)):
        img = (syn.scene(2 * H, 2 * W, 12).astype(np.float32) / np.float32(16383.0))[:H * 2:2, :W * 2:2]
        img = np.stack([img, np.roll(img, 3, axis=1), rng.random((H, W), dtype=np.float32)], axis=2).astype(np.float32)
        tables = np.stack([ext.compute_remapping_table(*c, W, H, CENTRE[0], CENTRE[1], scale) for c in COEFFS])
        out = np.array(img, copy=True)
        mod.apply_opcode_3_warp(out, opcode_block(COEFFS, CENTRE), scale)
        # a prior mapping (stack_warp_prior with one custom channel map), then the warp on top of it
        shift = np.stack(np.meshgrid(np.arange(W, dtype=np.float32) + 0.37, np.arange(H, dtype=np.float32) - 0.21), axis=2)
        prior = mod.stack_warp_prior(img, shift.astype(np.float32), None, None)
        tables_prior = np.stack([ext.compute_offset_remapping_table(np.ascontiguousarray(prior[..., i, :]), *c, W, H, CENTRE[0], CENTRE[1], scale)
                                 for i, c in enumerate(COEFFS)])
        out_prior = np.array(img, copy=True)
        mod.apply_opcode_3_warp(out_prior, opcode_block(COEFFS, CENTRE), scale, prior)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), image=img, coeffs=np.array(COEFFS), centre=np.array(CENTRE),
                            scale=scale, opcode=np.frombuffer(opcode_block(COEFFS, CENTRE), dtype=np.uint8), tables=tables,
                            warped=out, prior=prior, tables_prior=tables_prior, warped_prior=out_prior, **meta)
        print(name, out.shape, float(np.abs(out - img).mean()))


if __name__ == "__main__":
    main()
