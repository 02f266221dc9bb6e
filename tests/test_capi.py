"""The C-ABI library loads and exports every symbol include/pysp_b200.h declares; argument validation
(which happens before any CUDA call) returns the documented codes.  No compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from pysp_b200 import _capi


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_capi.LIB_PATH):
        from pysp_b200 import build
        build.build()
    return _capi.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pysp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pysp_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert "pysp_develop" in names and len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layout_matches_header(lib):
    # compile the header with gcc and compare sizeof/offsetof with the ctypes mirror
    import subprocess
    import tempfile
    if not os.path.exists("/usr/bin/gcc"):
        pytest.skip("gcc not available")
    src = '#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu", sizeof(pysp_develop_args),' \
          'offsetof(pysp_develop_args, cam_to_srgb), offsetof(pysp_develop_args, out), offsetof(pysp_develop_args, lab_lut));}' \
          % os.path.join(ROOT, "include", "pysp_b200.h")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        got = [int(v) for v in subprocess.check_output([os.path.join(d, "t")]).split()]
    A = _capi.DevelopArgs
    assert got == [C.sizeof(A), A.cam_to_srgb.offset, A.out.offset, A.lab_lut.offset]


def _args(**kw):
    buf = np.zeros((16, 16), dtype=np.uint16)
    out = np.zeros((16, 16, 3), dtype=np.float32)
    base = dict(height=16, width=16, pattern="RGGB", in_kind=_capi.IN_U16, in_ptr=buf.ctypes.data, in_pitch=32,
                in_row0=0, in_rows=16, black=[0] * 4, white=[1] * 4, wb=[1, 1, 1], cam_to_srgb=np.eye(3), stages=0,
                is_hdr=False, gamma=False, out_kind=_capi.OUT_LIN_F32, out_ptr=out.ctypes.data, out_pitch=16 * 12,
                out_row0=0, row_begin=0, row_end=16, scratch_ptr=None, scratch_bytes=0, lut_ptr=buf.ctypes.data)
    base.update(kw)
    return _capi.fill_develop_args(**base), (buf, out)


@pytest.mark.parametrize("kw", [dict(height=15), dict(width=2), dict(row_begin=1), dict(row_end=18), dict(in_rows=8),
                                dict(in_pitch=8), dict(out_pitch=16), dict(stages=1), dict(out_kind=7)])
def test_invalid_arguments(lib, kw):
    a, keep = _args(**kw)
    assert lib.pysp_develop(C.byref(a), None) == _capi.ERR_INVALID
    with pytest.raises(ValueError):
        _capi.check(_capi.ERR_INVALID)
    assert len(lib.pysp_last_error()) > 0


def test_unsupported_pattern(lib):
    a, keep = _args(pattern=9)
    assert lib.pysp_develop(C.byref(a), None) == _capi.ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _capi.check(_capi.ERR_UNSUPPORTED)
    with pytest.raises(NotImplementedError):
        _args(pattern="XTRANS")


def test_helpers(lib):
    assert lib.pysp_develop_halo_rows(1) == 10 and lib.pysp_develop_halo_rows(0) == 6 and lib.pysp_develop_halo_rows(3) == 18
    assert lib.pysp_develop_scratch_bytes(100, 50, 0) == 0
    assert lib.pysp_develop_scratch_bytes(100, 50, 1) == (50 + 8) * 100 * 12
    assert lib.pysp_develop_scratch_bytes(100, 50, 3) == 2 * (50 + 24) * 100 * 12
    assert lib.pysp_lab_lut_bytes() == 34 * 34 * 33 * 16
    lut = (np.arange(33 ** 3 * 3, dtype=np.int64).reshape(33, 33, 33, 3) * 7) % 16384
    lut16 = np.ascontiguousarray(lut.astype(np.int16))
    packed = np.zeros(34 * 34 * 33 * 4, dtype=np.uint32)
    assert lib.pysp_lab_lut_pack_host(lut16.ctypes.data, packed.ctypes.data) == 0
    p = packed.reshape(34, 34, 33, 4)
    for ch in range(3):
        assert np.array_equal(p[:33, :33, :, ch] & 0xFFFF, lut[..., ch])
        assert np.array_equal(p[:33, :33, :32, ch] >> 16, lut[:, :, 1:, ch])
        assert np.array_equal(p[:33, :33, 32, ch] >> 16, lut[:, :, 32, ch])
        assert np.array_equal(p[33, :33, :, ch], p[32, :33, :, ch]) and np.array_equal(p[:, 33, :, ch], p[:, 32, :, ch])
    assert not p[..., 3].any()
    assert b"sm_100a" in lib.pysp_version()


def test_no_cpu_fallback(lib):
    """Without a CUDA device a compute call fails loudly (PYSP_ERR_CUDA / RuntimeError)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    a, keep = _args()
    assert lib.pysp_develop(C.byref(a), None) == _capi.ERR_CUDA
    from pysp_b200 import engine
    with pytest.raises(RuntimeError):
        engine.require_cuda()
    from pysp_b200 import bayer_normalize
    with pytest.raises(RuntimeError):
        bayer_normalize(np.zeros((4, 4), dtype=np.uint16), [0] * 4, [1] * 4)
