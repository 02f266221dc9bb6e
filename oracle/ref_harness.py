"""Import the UNMODIFIED reference (bullbin/pySP) in THIS container, for pinning the oracle.

TEST INFRASTRUCTURE ONLY -- never imported by `pysp_b200/`.  `/root/reference` does not exist on
the GPU box, so nothing here runs in `-m gpu` tests or `smoke()`; it is used by
`tests/golden/make_golden.py` (which writes the committed fixtures), by the optional
`tests/test_oracle_vs_reference.py`, which skips when the reference is not mounted, and by the CPU arm
of `bench.py` (`--impl reference`, `cpu_baseline`), which imports the reference from the git-ignored
install `baseline/_ref/pySP` (oracle/install_ref.py) when that travelled with the repo.

What it does (none of it changes reference arithmetic):
  * registers the reference directory as package `pySP` (it uses absolute `pySP.*` imports,
    e.g. base_types/image_base.py:7-10) without copying it;
  * provides in-memory stand-ins for the un-installable third-party modules that the hot path only
    touches at import time: `colour` (only `xy_to_XYZ`, used at colorize/rgb_space.py:3,14,40),
    `rawpy`, `exifread`, `tifftools` (image.py:3,6);
  * loads the gcc-built Cython extension from `oracle/_ref/` (see oracle/build_ref.py) under its
    package name `pySP.debayer.ahd_homogeneity_cython` (debayer/ahd.py:11);
  * offers a duck-typed white-balance controller exposing exactly the three accessors the develop
    path uses (wb_cct/cam_wb.py:236-251 + copy()), because the real controller's solver needs
    colour-science's Ohno-2013 routines.
  * `pin_numerics(True)` switches OpenCV to its generic code paths so float32 tap orders are fully
    specified; `pin_numerics(False)` restores the optimised defaults.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_INSTALLED = os.path.abspath(os.path.join(_HERE, "..", "baseline", "_ref", "pySP"))     # oracle/install_ref.py


def _reference_root():
    env = os.environ.get("PYSP_REFERENCE_ROOT")
    if env:
        return env
    if os.path.exists("/root/reference/debayer/ahd.py"):
        return "/root/reference"
    return _INSTALLED            # the copy that travels to the GPU box (bench.py's CPU arm only)


REFERENCE_ROOT = _reference_root()


def available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, "debayer", "ahd.py"))


def _install_shims():
    if "colour" not in sys.modules:
        colour = types.ModuleType("colour")

        def xy_to_XYZ(xy):
            x, y = float(xy[0]), float(xy[1])
            return np.array([x / y, 1.0, (1.0 - x - y) / y], dtype=np.float64)

        colour.xy_to_XYZ = xy_to_XYZ
        sys.modules["colour"] = colour
    if "rawpy" not in sys.modules:
        rawpy = types.ModuleType("rawpy")

        class LibRawError(Exception):
            pass

        def imread(*a, **k):
            raise LibRawError("rawpy is not installed (oracle shim)")

        rawpy.LibRawError = LibRawError
        rawpy.imread = imread
        for name in ("LibRawFileUnsupportedError", "LibRawIOError", "NotSupportedError"):
            setattr(rawpy, name, LibRawError)
        sys.modules["rawpy"] = rawpy
    if "exifread" not in sys.modules:
        exifread = types.ModuleType("exifread")
        exifread.process_file = lambda *a, **k: {}
        sys.modules["exifread"] = exifread
    if "tifftools" not in sys.modules:
        tifftools = types.ModuleType("tifftools")

        def read_tiff(*a, **k):
            raise IOError("tifftools is not installed (oracle shim)")

        tifftools.read_tiff = read_tiff
        tifftools.Datatype = types.SimpleNamespace(get=lambda *_: None)
        tifftools.Tag = types.SimpleNamespace(SubIFD=types.SimpleNamespace(value=330))
        sys.modules["tifftools"] = tifftools


def load():
    """Return the reference package (module `pySP`)."""
    if "pySP" in sys.modules:
        return sys.modules["pySP"]
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    _install_shims()
    from oracle import build_ref  # noqa: local recipe
    so = build_ref.build(REFERENCE_ROOT)
    name = "pySP.debayer.ahd_homogeneity_cython"
    loader = importlib.machinery.ExtensionFileLoader(name, so)
    ext_spec = importlib.util.spec_from_file_location(name, so, loader=loader)
    ext = importlib.util.module_from_spec(ext_spec)
    loader.exec_module(ext)
    sys.modules[name] = ext                  # found by `from .ahd_homogeneity_cython import ...`
    spec = importlib.util.spec_from_file_location(
        "pySP", os.path.join(REFERENCE_ROOT, "__init__.py"),
        submodule_search_locations=[REFERENCE_ROOT])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["pySP"] = pkg
    old_dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True           # /root/reference is read-only
    try:
        spec.loader.exec_module(pkg)
        for sub in ("pySP.debayer", "pySP.image", "pySP.raw_hdr", "pySP.normalization"):
            importlib.import_module(sub)
    finally:
        sys.dont_write_bytecode = old_dont_write
    return pkg


class StubWhiteBalance:
    """Duck-typed stand-in for CameraWhiteBalanceController (wb_cct/cam_wb.py:236-251)."""

    def __init__(self, mat_xyz_to_cam, white_xyz, neutral=None):
        load()
        from pySP.wb_cct.helpers_cam_mat import MatXyzToCamera
        self._mat = MatXyzToCamera(np.asarray(mat_xyz_to_cam), np.asarray(white_xyz, dtype=np.float64))
        if neutral is None:
            neutral = (np.asarray(mat_xyz_to_cam, dtype=np.float64) @ np.asarray(white_xyz, dtype=np.float64))
        self._neutral = np.asarray(neutral, dtype=np.float32)

    def get_reciprocal_multipliers(self):
        return np.copy(1.0 / self._neutral)      # float32, as on the EXIF path (helpers_exif.py:79)

    def get_matrix(self):
        return self._mat

    def copy(self):
        return StubWhiteBalance(self._mat.mat, self._mat.xyz, self._neutral)


def pin_numerics(pinned=True):
    import cv2
    cv2.setUseOptimized(not pinned)
    try:
        cv2.ipp.setUseIPP(not pinned)
    except Exception:
        pass


def develop(raw_u16, black, white, stages=1, mat=None, xyz=None):
    """The develop path exactly as a pySP user runs it (README.md:54-63): bayer_normalize -> RawBayerData.demosaic(
    QualityDemosaic.Best, stages) -> .to_lin_srgb().  Returns float32 [H, W, 3] linear sRGB.  OpenCV mode is whatever
    `pin_numerics` last set (default: optimised)."""
    load()
    from pySP.normalization import bayer_normalize
    from pySP.image import RawBayerData
    from pySP.base_types.image_base import BayerPattern
    from pySP.const import QualityDemosaic
    from pysp_b200 import synthetic as syn
    img = RawBayerData()
    img.sensor_scaled = bayer_normalize(raw_u16, list(black), list(white))
    img.sensor_pattern = BayerPattern.Rggb
    img.cam_wb = StubWhiteBalance(syn.MAT_XYZ_TO_CAM if mat is None else mat, syn.WHITE_XYZ if xyz is None else xyz)
    img.current_ev = 10.0
    return img.demosaic(QualityDemosaic.Best, stages).to_lin_srgb()


def make_rggb_container(sensor_scaled, wb, ev=10.0, lim_sat=1.0, hdr=False, pattern=None):
    load()
    from pySP.image import RawRggbBayerData
    from pySP.base_types.image_base import BayerPattern
    img = RawRggbBayerData(sensor_scaled, wb, ev, lim_sat,
                           BayerPattern.Rggb if pattern is None else pattern)
    img.set_hdr(hdr)
    return img


def patch_hdr_ctor():
    """raw_hdr.py:150 calls `RawRggbBayerData()` with no arguments (TypeError as shipped).  Replace the
    name inside pySP.raw_hdr by a subclass whose constructor tolerates that; arithmetic lines
    108-148 are untouched."""
    load()
    import pySP.raw_hdr as raw_hdr
    from pySP.image import RawRggbBayerData

    class _Tolerant(RawRggbBayerData):
        def __init__(self, *a, **k):
            if a or k:
                super().__init__(*a, **k)
            else:
                super().__init__(None, None, np.inf, 1.0)

    raw_hdr.RawRggbBayerData = _Tolerant
    return raw_hdr
