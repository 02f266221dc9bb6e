"""Parity of the CUDA path (through the C ABI) against the golden fixtures of the reference and against
the oracle on seeded inputs.  Bit-exact for everything up to linear sRGB; the sRGB gamma (powf) to 1e-4
relative, the tolerance BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch

from conftest import assert_bit_equal, golden, golden_develop_cases
from oracle import ahd_spec as sp
from pysp_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

WB = syn.wb_multipliers()
M = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)


@pytest.fixture(scope="module")
def eng():
    from pysp_b200 import engine, _capi
    assert torch.cuda.is_available(), "these tests need the GPU"
    _capi.lib()                       # fails loudly if the CUDA extension is missing
    return engine


def gpu_develop(eng, src, stages, pattern="RGGB", black=syn.BLACK, white=syn.WHITE, out="lin", hdr=False, **kw):
    t = eng.to_device(src)
    r = eng.develop(t, WB, M, stages=stages, pattern=pattern, black=black, white=white, hdr=hdr, out=out, **kw)
    torch.cuda.synchronize()
    return r.cpu().numpy()


@pytest.mark.parametrize("name", golden_develop_cases())
def test_golden_fixtures(eng, name):
    d = golden(name)
    args = (d["raw"], int(d["stages"]), str(d["pattern"]), d["black"], d["white"])
    assert_bit_equal(gpu_develop(eng, *args, out="cam"), d["cam"], "camera RGB")
    assert_bit_equal(gpu_develop(eng, *args, out="lin"), d["lin"], "linear sRGB")


@pytest.mark.parametrize("stages", [0, 1])
def test_golden_hdr(eng, stages):
    d = golden("hdr48x64_s%d" % stages)
    assert_bit_equal(gpu_develop(eng, d["sensor"], stages, hdr=True, out="cam"), d["cam"], "HDR camera RGB")
    assert_bit_equal(gpu_develop(eng, d["sensor"], stages, hdr=True, out="lin"), d["lin"], "HDR linear sRGB")


def test_nonfinite_photosites(eng):
    """+inf / -inf / NaN photosites (HDR / float mosaics): the multiplicative blend of debayer/ahd.py:139-145 poisons every
    pixel whose other candidate is non-finite, np.clip keeps NaN; same NaN set and same finite bits as the reference."""
    from conftest import assert_bit_equal_nan
    d = golden("nonfinite48x64_s0")
    dm = torch.zeros(d["sensor"].shape, dtype=torch.uint8, device="cuda")
    t = eng.to_device(d["sensor"])
    cam = eng.develop(t, WB, M, stages=0, out="cam", dir_map=dm)
    assert np.array_equal(dm.cpu().numpy().astype(bool), d["pick_h"])
    assert_bit_equal_nan(cam.cpu().numpy(), d["cam"], "camera RGB with non-finite photosites")
    assert_bit_equal_nan(gpu_develop(eng, d["sensor"], 0), d["lin"], "linear sRGB with non-finite photosites")


@pytest.mark.parametrize("shape,stages,seed", [((512, 768), 1, 0), ((250, 1002), 0, 1), ((1000, 1500), 3, 2),
                                               ((130, 62), 2, 3)])
def test_against_oracle(eng, shape, stages, seed):
    raw = syn.scene(shape[0], shape[1], seed)
    lin, cam, ex = sp.develop(raw, syn.BLACK, syn.WHITE, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, stages, keep=True)
    got0 = gpu_develop(eng, raw, 0, out="cam")
    # AHD direction choice: the selected image equals the oracle's selection everywhere
    assert_bit_equal(got0, ex["selected"], "selected camera RGB (direction map)")
    assert_bit_equal(gpu_develop(eng, raw, stages, out="cam"), cam, "camera RGB")
    assert_bit_equal(gpu_develop(eng, raw, stages, out="lin"), lin, "linear sRGB")


@pytest.mark.parametrize("pattern", ["BGGR", "GRBG", "GBRG"])
def test_patterns(eng, pattern):
    raw = syn.scene(300, 420, 4)
    lin, _ = sp.develop(raw, (500, 510, 520, 530), (16383, 16000, 15800, 16100), WB, syn.MAT_XYZ_TO_CAM,
                        syn.WHITE_XYZ, 1, pattern)
    assert_bit_equal(gpu_develop(eng, raw, 1, pattern, (500, 510, 520, 530), (16383, 16000, 15800, 16100)), lin, pattern)


@pytest.mark.parametrize("seed", range(10))
def test_ragged_frames_and_bands(eng, seed):
    """Seeded random cases on the GPU: ragged even frame sizes (4x4 upwards, widths that are 2 mod 4 or not 16-byte rows,
    partial tiles on both axes), every CFA pattern, per-site levels, 0-3 stages, a random split into row bands -- the whole
    frame against the oracle, every band (developed from its rows + halo only) against the whole frame."""
    rng = np.random.default_rng(300 + seed)
    H, W = 2 * int(rng.integers(2, 150)), 2 * int(rng.integers(2, 170))
    stages = int(rng.integers(0, 4))
    pattern = ("RGGB", "BGGR", "GRBG", "GBRG")[int(rng.integers(0, 4))]
    black = tuple(int(v) for v in rng.integers(400, 600, 4))
    white = tuple(int(v) for v in rng.integers(15000, 16384, 4))
    raw = syn.random_mosaic(H, W, 400 + seed) if seed % 3 == 0 else syn.scene(H, W, 400 + seed)
    lin, _ = sp.develop(raw, black, white, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, stages, pattern)
    what = "%dx%d %s stages=%d" % (H, W, pattern, stages)
    assert_bit_equal(gpu_develop(eng, raw, stages, pattern, black, white), lin, what)
    padded = eng.develop(eng.to_device(raw, pad_pitch=True), WB, M, stages=stages, pattern=pattern, black=black, white=white)
    assert_bit_equal(padded.cpu().numpy(), lin, what + " (padded pitch)")
    cuts = sorted(set([0, H] + [2 * int(v) for v in rng.integers(1, H // 2, size=min(3, H // 2 - 1))])) if H > 4 else [0, H]
    halo = 6 + 4 * stages
    for rb, re in zip(cuts[:-1], cuts[1:]):
        r0, r1 = max(0, rb - halo), min(H, re + halo)
        t = eng.to_device(np.ascontiguousarray(raw[r0:r1]))
        band = eng.develop(t, WB, M, stages=stages, pattern=pattern, black=black, white=white, rows=(rb, re), frame_height=H,
                           in_row0=r0)
        assert_bit_equal(band.cpu().numpy(), lin[rb:re], what + " band [%d,%d)" % (rb, re))


def test_random_noise_frame(eng):
    raw = syn.random_mosaic(256, 384, 5)      # white noise: every direction vote is contested
    lin, _ = sp.develop(raw, syn.BLACK, syn.WHITE, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, 1)
    assert_bit_equal(gpu_develop(eng, raw, 1), lin, "white-noise frame")


def test_row_bands_equal_whole_frame(eng):
    raw = syn.scene(600, 400, 6)
    for stages in (0, 1, 2):
        whole = gpu_develop(eng, raw, stages)
        halo = 6 + 4 * stages
        parts = []
        for rb, re in ((0, 150), (150, 298), (298, 600)):
            r0, r1 = max(0, rb - halo), min(600, re + halo)
            t = eng.to_device(raw[r0:r1])
            o = eng.develop(t, WB, M, stages=stages, black=syn.BLACK, white=syn.WHITE, rows=(rb, re),
                            frame_height=600, in_row0=r0)
            parts.append(o.cpu().numpy())
        assert_bit_equal(np.concatenate(parts), whole, "bands, stages=%d" % stages)


def test_full_size_windows_against_oracle(eng):
    """24 MP frame (BASELINE config 2): the GPU result inside sampled windows equals the oracle run on the
    window plus margin (size-independent pin of the full-size run), and repeated runs are identical."""
    H, W, stages, margin = 4000, 6000, 1, 16
    raw = syn.scene(H, W, 0)
    t = eng.to_device(raw)
    a = eng.develop(t, WB, M, stages=stages, black=syn.BLACK, white=syn.WHITE)
    b = eng.develop(t, WB, M, stages=stages, black=syn.BLACK, white=syn.WHITE)
    torch.cuda.synchronize()
    assert torch.equal(a.view(torch.int32), b.view(torch.int32))
    for (y, x) in [(0, 0), (H - 256, W - 256), (1990, 2990), (300, 5744), (3744, 0), (1000, 700)]:
        y0, y1, x0, x1 = max(0, y - margin), min(H, y + 256 + margin), max(0, x - margin), min(W, x + 256 + margin)
        lin, _ = sp.develop(raw[y0:y1, x0:x1], syn.BLACK, syn.WHITE, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, stages)
        # compare away from the crop's artificial borders, keep the true frame borders
        cy0 = 0 if y0 == 0 else margin
        cx0 = 0 if x0 == 0 else margin
        cy1 = (y1 - y0) if y1 == H else (y1 - y0 - margin)
        cx1 = (x1 - x0) if x1 == W else (x1 - x0 - margin)
        got = a[y0 + cy0:y0 + cy1, x0 + cx0:x0 + cx1].cpu().numpy()
        assert_bit_equal(got, lin[cy0:cy1, cx0:cx1], "window at (%d,%d)" % (y, x))


def test_fused_outputs(eng):
    raw = syn.scene(200, 300, 8)
    lin, _ = sp.develop(raw, syn.BLACK, syn.WHITE, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, 1)
    half = gpu_develop(eng, raw, 1, out="lin_f16")
    assert half.dtype == np.float16 and np.array_equal(half, lin.astype(np.float16))
    srgb = gpu_develop(eng, raw, 1, gamma=True)
    ref = sp.lin_srgb_to_srgb(lin)
    assert np.all(np.abs(srgb - ref) <= 1e-4 * np.maximum(np.abs(ref), 1e-3))
    # wire formats: the gamma-encoded value rounded to nearest into 8 / 16 bits; the gamma is toleranced (1e-4), so a
    # value within that distance of a rounding boundary may land on the neighbouring code
    for kind, scale, dt in (("srgb_u8", 255.0, np.uint8), ("srgb_u16", 65535.0, np.uint16)):
        q = gpu_develop(eng, raw, 1, out=kind)
        assert q.dtype == dt and q.shape == lin.shape
        exact = ref.astype(np.float64) * scale
        want = np.rint(exact)
        d = np.abs(q.astype(np.float64) - want)
        assert d.max() <= 1
        near = np.abs(exact - np.floor(exact) - 0.5) <= 1e-4 * scale
        assert not (d > 0)[~near].any(), "%s differs away from rounding boundaries" % kind
    for pattern in ("BGGR", "GBRG"):           # flipped stores of the narrow kinds
        q = gpu_develop(eng, raw, 1, pattern=pattern, out="srgb_u8")
        f = gpu_develop(eng, raw, 1, pattern=pattern, gamma=True)
        assert np.abs(q.astype(np.float64) - np.rint(f.astype(np.float64) * 255.0)).max() <= 1


def test_pointwise_entry_points(eng):
    from pysp_b200 import bayer_normalize, cam_to_lin_srgb, lin_srgb_to_srgb, clip_rgb
    from pysp_b200.wb_cct import CameraWhiteBalance
    d = golden("scene64x96_RGGB")
    assert_bit_equal(bayer_normalize(d["raw"], list(d["black"]), list(d["white"])), d["sensor"], "bayer_normalize")
    wb = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    assert_bit_equal(cam_to_lin_srgb(d["cam"], wb.get_matrix()), d["lin"], "cam_to_lin_srgb")
    x = np.random.default_rng(0).uniform(-0.5, 1.5, size=(33, 17, 3)).astype(np.float32)
    assert_bit_equal(cam_to_lin_srgb(x, wb.get_matrix(), clip_highlights=False), sp.mat3_f64(x, M), "unclipped matrix")
    assert_bit_equal(clip_rgb(x), np.clip(x, 0, 1), "clip_rgb")
    pins = golden("dot_fma_pins")       # float64 accumulation order of the reference's np.dot (FMA chain)
    assert_bit_equal(cam_to_lin_srgb(pins["x"], wb.get_matrix()), pins["y"], "cam_to_lin_srgb on the accumulation-order pins")
    g = golden("gamma")
    y = lin_srgb_to_srgb(g["x"])
    assert np.all(np.abs(y - g["y"]) <= 1e-4 * np.maximum(np.abs(g["y"]), 1e-3))
    t = lin_srgb_to_srgb(torch.from_numpy(g["x"]).cuda())
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), y)


def test_drop_in_containers(eng):
    """RawRgbgDataFromRaw(...).debayer(QualityDemosaic.Best).to_lin_srgb() and lin_srgb_to_srgb."""
    import pysp_b200 as P
    from pysp_b200.wb_cct import CameraWhiteBalance
    d = golden("scene64x96_GRBG")
    wb = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    img = P.RawRgbgDataFromRaw.from_mosaic(d["raw"], list(d["black"]), list(d["white"]), P.BayerPattern.Grbg, wb, ev=10.0)
    dem = img.debayer(P.QualityDemosaic.Best)
    assert isinstance(dem.image, np.ndarray) and dem.current_ev == 10.0 and dem.is_valid()
    assert_bit_equal(dem.image, d["cam"], "RawDemosaicData.image")
    assert_bit_equal(dem.to_lin_srgb(), d["lin"], "to_lin_srgb")
    assert_bit_equal(img.develop(1), d["lin"], "fused develop")
    # container path on a float32 sensor (bayer_normalize first), RGGB after to_rggb()
    img2 = P.RawBayerData()
    img2.sensor_scaled = P.bayer_normalize(d["raw"], list(d["black"]), list(d["white"]))
    img2.sensor_pattern = P.BayerPattern.Grbg
    img2.cam_wb = wb
    img2.current_ev = 10.0
    rggb = img2.to_rggb()
    dem2 = rggb.demosaic(P.QualityDemosaic.Best, postprocess_steps=1)
    assert_bit_equal(dem2.image, d["cam"], "RawRggbBayerData.demosaic")
    wbc = np.asarray(dem2._wb_coeff, dtype=np.float32)
    dem2.wb_undo()                                     # base_types/image_base.py:52-60: float64 division, rounded to float32
    undone = (d["cam"].astype(np.float64) / wbc[:3]).astype(np.float32)
    assert_bit_equal(dem2.image, undone, "wb_undo")
    dem2.wb_apply()                                    # image_base.py:45-49: float32 product
    assert_bit_equal(dem2.image, (undone * wbc[:3]).astype(np.float32), "wb_apply")
    norm = P.RawDemosaicData((d["cam"] / max(wbc)).astype(np.float32), wbc, wb_norm=True)
    norm.wb_undo()
    assert_bit_equal(norm.image, (((d["cam"] / max(wbc)).astype(np.float32) * max(wbc)).astype(np.float64) / wbc[:3]).astype(np.float32),
                     "wb_undo of a normalised image")
    # coefficients that are not float32 (a float64 array from a solver, a Python list): the reference's NumPy expressions
    # then run in float64 (image_base.py:45-60 under NEP 50 promotion); same bits here
    for coeff in (wbc.astype(np.float64) * 1.00000013, [float(v) * 0.99999987 for v in wbc]):
        obj = P.RawDemosaicData(np.array(d["cam"], copy=True), coeff, wb_norm=True)
        obj.wb_undo()
        ref = d["cam"] * max(coeff)
        ref = (ref.astype(np.float64) / coeff[:3]).astype(np.float32)
        assert_bit_equal(obj.image, ref, "wb_undo with %s coefficients" % type(coeff).__name__)
        obj.wb_apply()
        assert_bit_equal(obj.image, (ref * coeff[:3]).astype(np.float32), "wb_apply with %s coefficients" % type(coeff).__name__)
    # CUDA tensors in -> CUDA tensors out
    img3 = P.RawRgbgDataFromRaw.from_mosaic(torch.from_numpy(d["raw"].view(np.int16)).cuda(), list(d["black"]),
                                            list(d["white"]), "Grbg", wb, ev=10.0)
    out = img3.develop(1)
    assert out.is_cuda
    assert_bit_equal(out.cpu().numpy(), d["lin"], "device-resident develop")


def test_side_stream_is_safe(eng):
    """Entry points called with a stream that is NOT torch's current stream: their temporaries and outputs are allocated on
    that stream (engine._on), so the caching allocator cannot hand a block that queued kernels still use to tensors of the
    current stream.  Provoked here by churning allocations of the same sizes on the current stream while the side stream
    works; every result must equal the one computed on the current stream."""
    raw = syn.scene(1200, 1800, 13)
    t = eng.to_device(raw)
    kw = dict(stages=2, black=syn.BLACK, white=syn.WHITE)
    want = eng.develop(t, WB, M, **kw)
    flat = torch.rand((1200, 1800), device="cuda") + 0.5
    sens = torch.rand((1200, 1800), device="cuda")
    want_ff = eng.flat_frame_correction(sens, flat)
    torch.cuda.synchronize()
    eng.release_scratch()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    outs, junk = [], []
    for i in range(12):
        outs.append((eng.develop(t, WB, M, stream=side, **kw), eng.flat_frame_correction(sens, flat, stream=side)))
        for _ in range(4):                       # same-sized blocks requested on the current stream while `side` is busy
            junk.append(torch.full((1200, 1800, 3), float(i), device="cuda"))
            junk.append(torch.full((1200, 1800), float(i), device="cuda"))
        junk = junk[-6:]
    side.synchronize()
    torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a.view(torch.int32), want.view(torch.int32))
        assert torch.equal(b.view(torch.int32), want_ff.view(torch.int32))


def test_develop_is_graph_capturable(eng):
    """include/pysp_b200.h: after the first call on a device pysp_develop only launches kernels -- it can be captured into a
    CUDA graph and replayed (no allocation, no synchronising call, no host-side state change on the launch path)."""
    raw = eng.to_device(syn.scene(600, 900, 14))
    raw2 = eng.to_device(syn.scene(600, 900, 15))
    kw = dict(stages=2, black=syn.BLACK, white=syn.WHITE)
    want = eng.develop(raw, WB, M, **kw)
    want2 = eng.develop(raw2, WB, M, **kw)
    src = raw.clone()
    out = torch.empty_like(want)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.develop(src, WB, M, out_tensor=out, **kw)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int32), want.view(torch.int32))
    src.copy_(raw2)                     # new input in the captured buffer, same graph
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.view(torch.int32), want2.view(torch.int32))


def test_frame_pipeline_host_buffers(eng):
    """pysp_b200.pipeline.FramePipeline (the end-to-end batch API of bench.py): pinned host mosaics in, pinned host results
    out, three streams in flight; every frame equals the device-resident develop, for float32 and 8-bit sRGB outputs; the
    copies-only pass used as the bench's ceiling moves the same bytes and leaves the buffers usable."""
    from pysp_b200.pipeline import FramePipeline
    H, W = 480, 720
    raws = [syn.scene(H, W, 30 + i) for i in range(7)]
    pin_in = [torch.from_numpy(r.view(np.int16)).pin_memory() for r in raws]
    for kind in ("lin", "srgb_u8"):
        pipe = FramePipeline(H, W, WB, M, stages=1, black=syn.BLACK, white=syn.WHITE, out=kind)
        pin_out = [pipe.pinned_output() for _ in raws]
        pipe.run(pin_in, pin_out)
        pipe.run_copies_only(pin_in, pin_out)            # overwrites the outputs with whatever the device buffers hold
        pipe.run(pin_in, pin_out)
        assert pipe.h2d_bytes() == H * W * 2 and pipe.d2h_bytes() == H * W * 3 * (4 if kind == "lin" else 1)
        for r, o in zip(raws, pin_out):
            want = eng.develop(eng.to_device(r), WB, M, stages=1, black=syn.BLACK, white=syn.WHITE, out=kind)
            assert torch.equal(o.cuda(), want), kind


def test_hdr_fuse(eng):
    import pysp_b200 as P
    from pysp_b200.wb_cct import CameraWhiteBalance
    d = golden("fuse5_40x56")
    wb = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    imgs = [P.RawRggbBayerData(b, wb, float(e), 1.0) for b, e in zip(d["brackets"], d["evs"])]
    hdr, cnt = P.fuse_exposures_to_raw(imgs)
    assert hdr.get_hdr() and hdr.lim_sat == float(d["lim_sat"]) and hdr.current_ev == float(d["target_ev"])
    assert_bit_equal(hdr.sensor_scaled, d["fused"], "fused mosaic")
    assert np.array_equal(cnt, d["count"])
    dem = hdr.demosaic(P.QualityDemosaic.Best, 1)
    assert_bit_equal(dem.image, d["cam"], "camera RGB of the fused mosaic")
    assert_bit_equal(dem.to_lin_srgb(), d["lin"], "linear sRGB of the fused mosaic")
    assert P.fuse_exposures_to_raw([]) is None


@pytest.mark.parametrize("mode", ["1", "2", "7"])
def test_generic_paths(mode):
    """The non-TMA load/store paths and the IEEE-division path (taken for unaligned tensors / unverified levels)
    give the same bits; forced here through the library's test hook in a fresh process."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r)\n"
        "from oracle import ahd_spec as sp\n"
        "from pysp_b200 import engine, synthetic as syn\n"
        "raw = syn.scene(200, 320, 1)\n"
        "m = sp.cam_to_lin_srgb_matrix(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)\n"
        "for pat, st in (('RGGB', 2), ('BGGR', 1), ('GBRG', 0)):\n"
        "    out = engine.develop(engine.to_device(raw), syn.wb_multipliers(), m, stages=st, pattern=pat, black=syn.BLACK, white=syn.WHITE)\n"
        "    lin, _ = sp.develop(raw, syn.BLACK, syn.WHITE, syn.wb_multipliers(), syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, st, pat)\n"
        "    assert np.array_equal(out.cpu().numpy().view(np.uint32), lin.view(np.uint32)), pat\n"
        "print('ok')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYSP_DISABLE_TMA=mode), capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("name", ["fast_rand8x8", "fast_scene34x50", "fast_scene64x96_GBRG", "fast_flat20x28", "fast_rand66x130"])
def test_fast_quality(eng, name):
    """QualityDemosaic.Fast: RawBayerData.demosaic(Fast).to_lin_srgb() against the reference's golden outputs."""
    import pysp_b200 as P
    from pysp_b200.wb_cct import CameraWhiteBalance
    d = golden(name)
    wb = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    img = P.RawRgbgDataFromRaw.from_mosaic(d["raw"], list(d["black"]), list(d["white"]), str(d["pattern"]).capitalize(), wb, ev=10.0)
    dem = img.demosaic(P.QualityDemosaic.Fast, postprocess_steps=3)
    assert_bit_equal(dem.image, d["cam"], "Fast camera RGB")
    assert_bit_equal(dem.to_lin_srgb(), d["lin"], "Fast linear sRGB")
    assert_bit_equal(img.develop(quality=P.QualityDemosaic.Fast), d["lin"], "Fast fused develop")
    img2 = P.RawBayerData()
    img2.sensor_scaled = P.bayer_normalize(d["raw"], list(d["black"]), list(d["white"]))
    img2.sensor_pattern = img.sensor_pattern
    img2.cam_wb = wb
    img2.current_ev = 10.0
    assert_bit_equal(img2.to_rggb().demosaic(P.QualityDemosaic.Fast).image, d["cam"], "Fast via RawRggbBayerData")


@pytest.mark.parametrize("shape", [(1200, 1808), (3000, 4000)])
def test_fast_quality_against_oracle(eng, shape):
    """QualityDemosaic.Fast against the oracle on a 2 MP frame and at BASELINE config 1's size (4000x3000, 12 MP)."""
    raw = syn.scene(shape[0], shape[1], 3)
    lin, _ = sp.develop_fast(raw, syn.BLACK, syn.WHITE, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    t = eng.to_device(raw)
    out = eng.develop(t, WB, M, black=syn.BLACK, white=syn.WHITE, quality="fast")
    torch.cuda.synchronize()
    assert_bit_equal(out.cpu().numpy(), lin, "Fast %dx%d frame" % (shape[1], shape[0]))


def test_large_frame_smoke(eng):
    """100 MP frame (config 5): runs, finite, deterministic checksum across two runs."""
    H, W = 8660, 11548
    raw = torch.from_numpy(syn.scene(H // 4, W // 4, 1).view(np.int16)).cuda().repeat(4, 4)[:H, :W].contiguous()
    a = eng.develop(raw, WB, M, stages=1, black=syn.BLACK, white=syn.WHITE)
    s1 = a.view(torch.int32).to(torch.int64).sum().item()
    assert torch.isfinite(a).all()
    del a
    b = eng.develop(raw, WB, M, stages=1, black=syn.BLACK, white=syn.WHITE)
    assert b.view(torch.int32).to(torch.int64).sum().item() == s1


def test_lab_exhaustive_against_cv2(eng):
    """cv2's float32 RGB->Lab quantises each channel to cvRound(v*16384) and only uses bits 5..14 of it: 513 distinct
    levels per channel.  Every one of the 513^3 keys goes through the kernel's lab_lookup and through cv2.cvtColor:
    bit-identical Lab (parity gate iii)."""
    cv2 = pytest.importorskip("cv2")
    lv = (np.arange(513, dtype=np.float32) * np.float32(32.0 / 16384.0)).astype(np.float32)
    g, b = np.meshgrid(lv, lv, indexing="ij")
    bad = 0
    for r in lv:
        x = np.stack([np.full_like(g, r), g, b], axis=-1)                 # [513, 513, 3]
        want = cv2.cvtColor(x, cv2.COLOR_RGB2LAB)
        got = eng.rgb_to_lab_cv2(torch.from_numpy(x).cuda()).cpu().numpy()
        bad += int((got.view(np.uint32) != want.view(np.uint32)).sum())
    assert bad == 0, "%d of %d Lab values differ from cv2" % (bad, 3 * 513 ** 3)
    # off-grid and out-of-range inputs
    rng = np.random.default_rng(2)
    x = rng.uniform(-0.3, 1.4, size=(1, 2_000_000, 3)).astype(np.float32)
    assert_bit_equal(eng.rgb_to_lab_cv2(torch.from_numpy(x).cuda()).cpu().numpy(), cv2.cvtColor(x, cv2.COLOR_RGB2LAB), "Lab")


def test_full_size_properties(eng):
    """Size-independent properties at BASELINE's full size (24 MP, configs 2 and 3):
    (1) CFA equivalence -- a BGGR / GRBG / GBRG frame is the mirrored RGGB frame (image.py:143-152), so developing the
        mirrored mosaic under that pattern must give the mirrored RGGB result, bit for bit (exercises the flipped TMA
        boxes, the alternating 8/12-px box margin and the mirrored stores on every tile of a full frame);
    (2) row bands with halo rows equal the whole frame at stages = 3 (halo 18 rows, config 3);
    (3) the fused float32 output equals camera RGB -> pysp_cam_to_lin_srgb applied afterwards."""
    H, W = 4000, 6000
    raw = syn.scene(H, W, 11)
    t = eng.to_device(raw)
    kw = dict(black=syn.BLACK, white=syn.WHITE)
    ref = eng.develop(t, WB, M, stages=1, **kw)
    for pattern, dims in (("BGGR", (0, 1)), ("GRBG", (0,)), ("GBRG", (1,))):
        # level order [TL,TR,BR,BL] is per stored position: equal levels here, so only the geometry is mirrored
        flipped = torch.flip(t, dims=dims).contiguous()
        out = eng.develop(flipped, WB, M, stages=1, pattern=pattern, **kw)
        assert torch.equal(torch.flip(out, dims=dims).view(torch.int32), ref.view(torch.int32)), pattern
        del out, flipped
    whole = eng.develop(t, WB, M, stages=3, **kw)
    halo = 6 + 4 * 3
    for rb, re in ((0, 1334), (1334, 2666), (2666, 4000)):
        r0, r1 = max(0, rb - halo), min(H, re + halo)
        band = eng.develop(t[r0:r1].contiguous(), WB, M, stages=3, rows=(rb, re), frame_height=H, in_row0=r0, **kw)
        assert torch.equal(band.view(torch.int32), whole[rb:re].view(torch.int32)), (rb, re)
        del band
    cam = eng.develop(t, WB, M, stages=1, out="cam", **kw)
    assert torch.equal(eng.cam_to_rgb(cam, M, clip=True).view(torch.int32), ref.view(torch.int32))


def test_hdr_brackets_full_size(eng):
    """BASELINE config 4 at full size: five 24 MP brackets fused in raw space (raw_hdr.py:108-148) and developed with the
    HDR branch of the homogeneity metric (debayer/ahd.py:52-59).  The fuse is point-wise, so its result on sampled windows
    equals the oracle on the cropped brackets exactly; the develop is checked on the same windows with a margin."""
    from pysp_b200.raw_hdr import fusion_constants
    H, W, margin = 4000, 6000, 16
    base = (syn.scene(H, W, 5, noise=0).astype(np.float32) - np.float32(512.0)) / np.float32(16383.0)
    evs = [8.0, 9.0, 10.0, 11.0, 12.0]
    brackets = []
    for k in range(5):
        n = np.random.default_rng(k).normal(0, 30.0 / 16383.0, size=base.shape).astype(np.float32)
        brackets.append(np.clip(base * np.float32(2.0 ** (2 - k)) + n, 0, 1).astype(np.float32))
    tev, offs, bias = fusion_constants(evs, WB)
    fused, cnt = eng.fuse_exposures([eng.to_device(b) for b in brackets], offs, bias, int(np.argmax(offs)))
    out = eng.develop(fused, WB, M, stages=1, hdr=True)
    torch.cuda.synchronize()
    for (y, x) in [(0, 0), (H - 200, W - 200), (1000, 3000), (3000, 200)]:
        y0, y1, x0, x1 = max(0, y - margin), min(H, y + 200 + margin), max(0, x - margin), min(W, x + 200 + margin)
        f_ref, c_ref, _, _ = sp.fuse_exposures([b[y0:y1, x0:x1] for b in brackets], evs, WB)
        assert_bit_equal(fused[y0:y1, x0:x1].cpu().numpy(), f_ref, "fused mosaic window (%d,%d)" % (y, x))
        assert np.array_equal(cnt[y0:y1, x0:x1].cpu().numpy(), c_ref)
        cam = sp.ahd_demosaic(f_ref, WB, M, 1, hdr=True)
        lin = sp.to_lin_srgb(cam, M)
        cy0 = 0 if y0 == 0 else margin
        cx0 = 0 if x0 == 0 else margin
        cy1 = (y1 - y0) if y1 == H else (y1 - y0 - margin)
        cx1 = (x1 - x0) if x1 == W else (x1 - x0 - margin)
        assert_bit_equal(out[y0 + cy0:y0 + cy1, x0 + cx0:x0 + cx1].cpu().numpy(), lin[cy0:cy1, cx0:cx1],
                         "HDR develop window (%d,%d)" % (y, x))


def test_repeatability_under_load(eng):
    """The tile pipeline re-uses shared memory across phases and tiles (TMA store of tile t in flight while tile t+1 is
    computed).  A hazard there would show up as run-to-run differences: 40 back-to-back 24 MP develops, alternating two
    frames so that the persistent CTAs never see the same data twice in a row, must each reproduce their first result."""
    H, W = 4000, 6000
    frames = [eng.to_device(syn.scene(H, W, s)) for s in (21, 22)]
    kw = dict(stages=1, black=syn.BLACK, white=syn.WHITE)
    outs = [torch.empty((H, W, 3), dtype=torch.float32, device="cuda") for _ in range(2)]
    sums = []
    for i in range(40):
        eng.develop(frames[i & 1], WB, M, out_tensor=outs[i & 1], **kw)
        sums.append(outs[i & 1].view(torch.int32).to(torch.int64).sum())
    torch.cuda.synchronize()
    vals = [int(s.item()) for s in sums]
    assert len(set(vals[0::2])) == 1 and len(set(vals[1::2])) == 1, vals


def test_padded_mosaic_through_containers(eng):
    """A mosaic whose rows are not 16-byte multiples (width 2 mod 8) is uploaded into a row-padded buffer by from_mosaic and
    stays padded through RawBayerData.demosaic / develop (TMA path); results equal the oracle."""
    import pysp_b200 as P
    from pysp_b200.wb_cct import CameraWhiteBalance
    raw = syn.scene(122, 202 + 8 * 3, 17)[:, :202]           # 202 x 2 B = 404 B rows
    raw = np.ascontiguousarray(raw)
    wb = CameraWhiteBalance(syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ)
    img = P.RawRgbgDataFromRaw.from_mosaic(raw, list(syn.BLACK), list(syn.WHITE), P.BayerPattern.Rggb, wb, ev=10.0)
    assert img._counts.stride(0) * 2 % 128 == 0 and img._counts.shape == (122, 202)
    lin, cam = sp.develop(raw, syn.BLACK, syn.WHITE, WB, syn.MAT_XYZ_TO_CAM, syn.WHITE_XYZ, 1)
    assert_bit_equal(img.demosaic(P.QualityDemosaic.Best).image, cam, "camera RGB from a padded mosaic")
    assert_bit_equal(img.develop(1), lin, "fused develop from a padded mosaic")
    sensor = P.bayer_normalize(torch.from_numpy(raw.view(np.int16)).cuda(), list(syn.BLACK), list(syn.WHITE))
    assert_bit_equal(sensor.cpu().numpy(), sp.normalize(raw, syn.BLACK, syn.WHITE), "bayer_normalize")
