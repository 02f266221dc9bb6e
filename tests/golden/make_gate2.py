"""Parity gate (ii) fixture: the reference in its DEFAULT mode (OpenCV optimised / IPP code paths,
what a pySP user actually runs) next to its generic mode on one 1536x1024 frame, stages = 1.

Run in the build container only (needs /root/reference and cv2):   python tests/golden/make_gate2.py

The default-mode linear-sRGB image is committed relative to the generic-mode one (whose SHA-256 is committed and which
the oracle reproduces bit for bit): per value the difference of the two float32 bit patterns, int8 where it fits
(44 % of the values differ, by a few ulp), a sparse list for the rest (the 4-px neighbourhoods of the direction flips).
Both direction maps are committed packed.  tests reconstruct the default-mode image exactly from these.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)

from oracle import ref_harness as rh  # noqa: E402
from pysp_b200 import synthetic as syn  # noqa: E402
from make_golden import run_reference  # noqa: E402

H, W, SEED, STAGES = 1024, 1536, 31, 1


def main():
    import cv2
    rh.load()
    raw = syn.scene(H, W, SEED)
    rh.pin_numerics(True)
    gen = run_reference(raw, syn.BLACK, syn.WHITE, STAGES)
    rh.pin_numerics(False)
    de = run_reference(raw, syn.BLACK, syn.WHITE, STAGES)
    rh.pin_numerics(True)
    d = de["lin"].view(np.int32).astype(np.int64) - gen["lin"].view(np.int32).astype(np.int64)
    big = np.flatnonzero(np.abs(d) > 127)
    small = d.copy()
    small.reshape(-1)[big] = 0
    np.savez_compressed(
        os.path.join(OUT, "gate2_1024x1536_s1.npz"), H=H, W=W, seed=SEED, stages=STAGES,
        input_sha256=hashlib.sha256(raw.tobytes()).hexdigest(),
        lin_generic_sha256=hashlib.sha256(np.ascontiguousarray(gen["lin"]).tobytes()).hexdigest(),
        lin_default_sha256=hashlib.sha256(np.ascontiguousarray(de["lin"]).tobytes()).hexdigest(),
        delta_i8=small.astype(np.int8), big_index=big.astype(np.int32), big_delta=d.reshape(-1)[big].astype(np.int64),
        pick_generic=np.packbits(gen["pick_h"], axis=1), pick_default=np.packbits(de["pick_h"], axis=1),
        cv2=cv2.__version__, numpy=np.__version__)
    print("direction flips default vs generic:", int((gen["pick_h"] != de["pick_h"]).sum()), "of", H * W,
          "; values differing:", float((d != 0).mean()))


if __name__ == "__main__":
    main()
