"""CPU side of the full-size pins (tests/golden/fullsize_pins.json, made by the unmodified reference on the full
frames): the synthetic inputs regenerate to the pinned hashes on this host, and the oracle reproduces sampled 250-row
strips of the reference's full-frame result when it is run on the strip plus a margin of halo rows (the develop path
has a bounded vertical reach, SURVEY.md section 8a: 6 + 4 * stages rows).  The `-m gpu` tests compare every strip."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ahd_spec as sp
from pysp_b200 import synthetic as syn

PINS = json.load(open(os.path.join(GOLDEN, "fullsize_pins.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def frame24():
    p = PINS["cfg2"]
    raw = syn.scene(p["H"], p["W"], p["seed"])
    assert sha(raw) == p["input_sha256"] == PINS["cfg3"]["input_sha256"]
    return raw


def oracle_strip(raw, pin, i, margin):
    H, strip, stages = pin["H"], pin["strip"], pin["stages"]
    y0, y1 = i * strip, min(H, (i + 1) * strip)
    r0, r1 = max(0, y0 - margin), min(H, y1 + margin)
    lin, cam, ex = sp.develop(raw[r0:r1], syn.BLACK, syn.WHITE, syn.wb_multipliers(), syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ,
                              stages, keep=True)
    s = slice(y0 - r0, y1 - r0)
    return lin[s], cam[s], np.packbits(ex["pick_h"][s], axis=1)


@pytest.mark.parametrize("name,strips", [("cfg2", (0, 7, 15)), ("cfg3", (8,))])
def test_oracle_reproduces_reference_strips(frame24, name, strips):
    pin = PINS[name]
    for i in strips:
        lin, cam, dirs = oracle_strip(frame24, pin, i, 6 + 4 * pin["stages"] + 6)
        assert sha(lin) == pin["lin"][i], "%s: linear sRGB of strip %d" % (name, i)
        assert sha(cam) == pin["cam"][i], "%s: camera RGB of strip %d" % (name, i)
        assert sha(dirs) == pin["dir"][i], "%s: direction map of strip %d" % (name, i)


def test_direction_map_files_match_their_hashes():
    for name in ("cfg2", "cfg4"):
        packed = np.load(os.path.join(GOLDEN, "fullsize_dir_%s.npz" % name))["packed"]
        pin = PINS[name]
        assert [sha(packed[y:y + pin["strip"]]) for y in range(0, pin["H"], pin["strip"])] == pin["dir"]
        frac = np.unpackbits(packed, axis=1)[:, :pin["W"]].mean()
        assert abs(frac - pin["pick_h_fraction"]) < 1e-12
