"""Pin the develop path at BASELINE.json's FULL sizes to the unmodified reference.

Run in the build container only (needs /root/reference and cv2; minutes of CPU time and tens of GB of RAM per case):
    python tests/golden/make_fullsize_pins.py [case ...]        # default: every case

For each case the UNMODIFIED reference (imported through oracle/ref_harness.py, OpenCV in its generic code paths so
that the float32 tap orders are specified) develops the synthetic frame; what is committed is
  tests/golden/fullsize_pins.json       per case: SHA-256 of the input mosaic, and per 250-row strip the SHA-256 of the
                                        float32 linear-sRGB image, of the camera-RGB image and of the packed AHD
                                        direction map (plus the fused mosaic / contribution count for the HDR case);
  tests/golden/fullsize_dir_<case>.npz  the packed direction map itself (np.packbits of `map_h < map_v`; cfg2 and cfg4).
The `-m gpu` tests (tests/test_gpu_fullsize.py) regenerate the same input (pysp_b200/synthetic.py is bit-reproducible
across hosts), check its hash, run the CUDA path and compare EVERY strip.

Cases (BASELINE.json configs):
  cfg2   6000x4000, seed 0, postprocess_stages = 1            (config 2, the bench frame)
  cfg3   same frame, postprocess_stages = 3                   (config 3)
  cfg4   five 24 MP brackets -> fuse_exposures_to_raw -> HDR develop, stages = 1   (config 4)
  cfg5   one 11548x8660 (100 MP) frame, seed 1, stages = 1    (config 5, the row-band frame)
Reference call chain: normalization.py:4-25 -> image.py:191-197 -> debayer/ahd.py:14-170 -> base_types/image_base.py:62-64
-> colorize/transform.py:76-87; raw_hdr.py:85-158 for cfg4.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from pysp_b200 import synthetic as syn  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
PINS = os.path.join(OUT, "fullsize_pins.json")
STRIP = 250

CASES = {
    "cfg2": dict(H=4000, W=6000, seed=0, stages=1, kind="frame"),
    "cfg3": dict(H=4000, W=6000, seed=0, stages=3, kind="frame"),
    "cfg4": dict(H=4000, W=6000, seed=5, stages=1, kind="hdr5"),
    "cfg5": dict(H=8660, W=11548, seed=1, stages=1, kind="frame"),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def strip_hashes(a, strip=STRIP):
    return [sha(a[y:y + strip]) for y in range(0, a.shape[0], strip)]


def run_case(name):
    import cv2
    sys.path.insert(0, OUT)
    from make_golden import run_reference
    c = CASES[name]
    rh.load()
    rh.pin_numerics(True)
    t0 = time.time()
    rec = dict(c)
    rec.update(strip=STRIP, cv2=cv2.__version__, numpy=np.__version__,
               mode="cv2.setUseOptimized(False), ipp off (generic code paths)")
    if c["kind"] == "frame":
        raw = syn.scene(c["H"], c["W"], c["seed"])
        rec["input_sha256"] = sha(raw)
        res = run_reference(raw, syn.BLACK, syn.WHITE, c["stages"])
    else:
        raw_hdr = rh.patch_hdr_ctor()
        from pySP.image import RawRggbBayerData
        from pySP.const import QualityDemosaic
        wb = rh.StubWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
        brackets, evs = syn.hdr_brackets(c["H"], c["W"], c["seed"], 5)
        rec["input_sha256"] = sha(np.stack(brackets))
        rec["evs"] = evs
        imgs = [RawRggbBayerData(b, wb, e, 1.0) for b, e in zip(brackets, evs)]
        fused, cnt = raw_hdr.fuse_exposures_to_raw(imgs)
        del imgs, brackets
        sensor = np.ascontiguousarray(fused.sensor_scaled, dtype=np.float32)
        rec["fused"] = strip_hashes(sensor)
        rec["count"] = strip_hashes(np.ascontiguousarray(cnt, dtype=np.int32))
        rec["lim_sat"] = float(fused.lim_sat)
        rec["target_ev"] = float(fused.current_ev)
        assert fused.get_hdr()
        res = run_reference(None, None, None, c["stages"], sensor_override=sensor, hdr=True)
    packed = np.packbits(res["pick_h"], axis=1)              # [H, ceil(W/8)]
    rec["lin"] = strip_hashes(res["lin"])
    rec["cam"] = strip_hashes(res["cam"])
    rec["dir"] = strip_hashes(packed)
    rec["pick_h_fraction"] = float(res["pick_h"].mean())
    rec["reference_seconds_generic_mode"] = round(time.time() - t0, 1)
    if name in ("cfg2", "cfg4"):                             # the map itself (about 1 MB each) allows counting mismatches;
        np.savez_compressed(os.path.join(OUT, "fullsize_dir_%s.npz" % name), packed=packed)   # cfg5 (4.3 MB) keeps hashes only
    if name == "cfg3":                                       # cfg3 has the direction map of cfg2 (same frame)
        rec["dir_same_as"] = "cfg2"
    return rec


def main():
    names = sys.argv[1:] or list(CASES)
    for name in names:
        rec = run_case(name)
        pins = json.load(open(PINS)) if os.path.exists(PINS) else {}
        pins[name] = rec
        with open(PINS, "w") as f:
            json.dump(pins, f, indent=1, sort_keys=True)
        print(name, "done in", rec["reference_seconds_generic_mode"], "s; pick_h", rec["pick_h_fraction"], flush=True)


if __name__ == "__main__":
    main()
