"""NumPy restatement of pySP's raw -> linear-sRGB develop path (the ORACLE, stage by stage).

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline
leg may import this module; the product (`pysp_b200/`) never does.

Parity status: PINNED against the reference itself.  `tests/golden/make_golden.py` runs the
unmodified reference (imported from /root/reference, OpenCV in generic mode so that every float32
tap order is specified) and stores its outputs under `tests/golden/`; `tests/test_oracle.py` checks
this restatement against those fixtures bit for bit (the reference ships no tests or golden vectors
of its own -- SURVEY.md section 4).

Every float32 operation below is rounded individually (NumPy evaluates one ufunc at a time, no FMA),
which is what the reference does.  Citations are `file:line` in the reference tree.

Third-party arithmetic restated here (source not in the reference tree): OpenCV
(`opencv_python==4.10.0.84`, requirements.txt:5) `copyMakeBorder`, `GaussianBlur`, `filter2D`,
`blur`, `medianBlur`, `cvtColor(COLOR_RGB2LAB)` at the call sites debayer/ahd.py:58,62,64,77-80,
120-121,133-134,151 and debayer/edge_assisted_gaussian.py:141,143; NumPy/OpenBLAS `np.dot`
(colorize/transform.py:52).

`backend="cv2"` swaps the hand-restated primitives for the library calls the reference makes
(same cost profile as the reference; used only for the CPU timing baseline).
"""
import os

import numpy as np

f32 = np.float32
_LUT_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pysp_b200", "data",
                         "lab_lut33_i16.npy")
_LUT = None

# debayer/ahd.py:89-94 : h = normalise(0.125*h_optimal + 0.875*h_fast), evaluated in float32.
_h_opt = np.array([-0.2569, 0.4339, 0.5138, 0.4339, -0.2569], dtype=f32)
_h_fast = np.array([-0.25, 0.5, 0.5, 0.5, -0.25], dtype=f32)
_h = (_h_opt * 0.125) + (_h_fast * (1 - 0.125))
H5 = (_h / _h.sum()).astype(f32)

# cv2.getGaussianKernel(3, 1.0) as float32 (debayer/ahd.py:120)
GAUSS_K0 = f32(0.27406862)
GAUSS_K1 = f32(0.45186275)


def lab_lut():
    global _LUT
    if _LUT is None:
        _LUT = np.load(_LUT_PATH).astype(np.int32)
    return _LUT


# ----------------------------------------------------------------------------------------------
# colour matrices (host side, float64) -- colorize/transform.py:40-49, colorize/rgb_space.py:19-52,
# wb_cct/helpers_cam_mat.py:7-20, wb_cct/standard_ill.py:33
# ----------------------------------------------------------------------------------------------
D65_XY = (0.31272, 0.32903)
REC709_PRIMARIES = ((0.64, 0.33), (0.3, 0.6), (0.15, 0.06))


def xy_to_xyz(xy):
    x, y = float(xy[0]), float(xy[1])
    return np.array([x / y, 1.0, (1.0 - x - y) / y], dtype=np.float64)


def bradford(cur_xyz, tgt_xyz):
    m = np.array([[0.8951000, 0.2664000, -0.1614000],
                  [-0.7502000, 1.7135000, 0.0367000],
                  [0.0389000, -0.0685000, 1.0296000]])
    lc = np.matmul(m, cur_xyz)
    lt = np.matmul(m, tgt_xyz)
    sc = lt / lc
    d = np.array([[sc[0], 0, 0], [0, sc[1], 0], [0, 0, sc[2]]])
    return np.matmul(np.linalg.inv(m), np.matmul(d, m))


def rec709_to_xyz(dest_white_xyz):
    (xr, yr), (xg, yg), (xb, yb) = REC709_PRIMARIES
    m = np.array([[xr / yr, xg / yg, xb / yb],
                  [1, 1, 1],
                  [(1 - xr - yr) / yr, (1 - xg - yg) / yg, (1 - xb - yb) / yb]])
    white = xy_to_xyz(D65_XY)
    s = np.linalg.inv(m) @ white
    m[:, 0] *= s[0]
    m[:, 1] *= s[1]
    m[:, 2] *= s[2]
    return bradford(white, np.array(dest_white_xyz, dtype=np.float64)) @ m


def cam_to_lin_srgb_matrix(mat_xyz_to_cam, white_xyz):
    """3x3 float64 M with lin_srgb = M @ cam_rgb (colorize/transform.py:40-52)."""
    to_xyz = rec709_to_xyz(np.asarray(white_xyz).tolist())
    cm = np.matmul(np.asarray(mat_xyz_to_cam), to_xyz)
    cm = cm / cm.sum(axis=1)[:, np.newaxis]
    return np.linalg.inv(cm)


# ----------------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------------
def split_planes(m):
    """bayer_chan_mixer.py:13-21 -> R(TL), G1(TR), B(BR), G2(BL)."""
    return m[0::2, 0::2], m[0::2, 1::2], m[1::2, 1::2], m[1::2, 0::2]


def join_planes(r, g1, b, g2):
    """bayer_chan_mixer.py:36-41."""
    out = np.zeros((r.shape[0] * 2, r.shape[1] * 2), dtype=r.dtype)
    out[0::2, 0::2] = r
    out[0::2, 1::2] = g1
    out[1::2, 1::2] = b
    out[1::2, 0::2] = g2
    return out


def normalize(raw, black, white):
    """normalization.py:20-25: clip(x - black, 0, white).astype(f32) / white per CFA site
    (divides by white, not white-black).  black/white: 4 numbers ordered R,G1,B,G2."""
    planes = [p.astype(f32) for p in split_planes(raw)]
    out = []
    for p, bl, wh in zip(planes, black, white):
        out.append(np.clip(p - f32(bl), f32(0), f32(wh)).astype(f32) / f32(wh))
    return join_planes(*out)


_FMA_LIB = None


def build_c(force=False):
    """Compile the C part of the oracle (oracle/csrc/dot3_fma.c) with gcc into oracle/_build/."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "csrc", "dot3_fma.c")
    so = os.path.join(here, "_build", "liboracle_dot3.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([gcc, "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so, "-lm"])
    return so


def _fma_lib():
    global _FMA_LIB
    if _FMA_LIB is None:
        import ctypes
        lib = ctypes.CDLL(build_c())
        lib.oracle_dot3_fma.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_dot3_fma.restype = None
        _FMA_LIB = lib
    return _FMA_LIB


def mat3_f64(rgb, m):
    """f32( M . rgb ) accumulated in float64 as the reference's `np.dot(rgb, M.T)` does (colorize/transform.py:52-53):
    OpenBLAS dgemm runs a fused multiply-add chain over k, acc = m0*c0; acc = fma(m1, c1, acc); acc = fma(m2, c2, acc).
    That is the canonical order of the oracle and of the CUDA path; it is pinned by tests/golden/dot_fma_pins.npz
    (inputs where the fused and unfused sums round to different float32 values).  The correctly rounded fma() comes
    from C (oracle/csrc/dot3_fma.c) -- NumPy has no fused multiply-add."""
    c = np.ascontiguousarray(rgb, dtype=f32)
    mm = np.ascontiguousarray(m, dtype=np.float64)
    out = np.empty(c.shape, dtype=f32)
    if c.size:
        _fma_lib().oracle_dot3_fma(c.ctypes.data, c.size // 3, mm.ctypes.data, out.ctypes.data)
    return out


def lab_cv(rgb):
    """cv2.cvtColor(float32, COLOR_RGB2LAB) restated.  Returns float32 Lab.
    Non-finite input: cv2 4.13 (generic and optimised paths alike, probed) clamps with min/max whose NaN operand loses,
    i.e. NaN -> 0, +inf -> 1, -inf -> 0; restated here so that frames with non-finite photosites are defined."""
    lut = lab_lut()
    x = rgb.astype(f32)
    x = np.clip(np.where(np.isnan(x), f32(0), x), f32(0), f32(1))
    c = np.rint(x * f32(16384.0)).astype(np.int32)          # cvRound: half to even
    t = c >> 9
    s = (c >> 5) & 15
    acc = np.zeros(rgb.shape[:-1] + (3,), dtype=np.int32)
    for dr in (0, 1):
        wr = s[..., 0] if dr else 16 - s[..., 0]
        ir = np.minimum(t[..., 0] + dr, 32)
        for dg in (0, 1):
            wg = s[..., 1] if dg else 16 - s[..., 1]
            ig = np.minimum(t[..., 1] + dg, 32)
            for db in (0, 1):
                wb = s[..., 2] if db else 16 - s[..., 2]
                ib = np.minimum(t[..., 2] + db, 32)
                acc += lut[ir, ig, ib] * (wr * wg * wb)[..., None]
    v = ((acc + 2048) >> 12).astype(f32)
    out = np.empty(v.shape, dtype=f32)
    out[..., 0] = v[..., 0] * f32(100.0 / 16384.0)
    out[..., 1] = v[..., 1] * f32(256.0 / 16384.0) - f32(128.0)
    out[..., 2] = v[..., 2] * f32(256.0 / 16384.0) - f32(128.0)
    return out


def gauss3(g):
    """cv2.GaussianBlur(g,(3,3),1.0), generic path: row pass then column pass of
    k1*mid + k0*(left+right), REFLECT_101."""
    p = np.pad(g, ((0, 0), (1, 1)), mode="reflect")
    rowp = GAUSS_K1 * p[:, 1:-1] + GAUSS_K0 * (p[:, :-2] + p[:, 2:])
    p = np.pad(rowp, ((1, 1), (0, 0)), mode="reflect")
    return GAUSS_K1 * p[1:-1, :] + GAUSS_K0 * (p[:-2, :] + p[2:, :])


_K64 = np.array([[1, 6, 1], [6, 36, 6], [1, 6, 1]], dtype=np.float64) / 64.0


def phase_kernels(base_bottom_right):
    """debayer/gaussian.py:19-53 for the 5x5 binomial kernel, as 3x3 correlation kernels on the
    quarter grid, ordered TL, TR, BL, BR (target phase)."""
    k5 = np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1]).astype(np.float64)
    out = []
    for bottom in (False, True):
        for right in (False, True):
            rows = k5[0::2] if bottom == base_bottom_right else k5[1::2]
            sub = rows[:, 0::2] if right == base_bottom_right else rows[:, 1::2]
            if right != base_bottom_right:
                # target is to the right of a left base -> taps at dj = 0,+1 ; left of right base -> -1,0
                sub = np.c_[np.zeros(sub.shape[0]), sub] if right else np.c_[sub, np.zeros(sub.shape[0])]
            if bottom != base_bottom_right:
                sub = np.r_[np.zeros((1, sub.shape[1])), sub] if bottom else np.r_[sub, np.zeros((1, sub.shape[1]))]
            out.append(sub / sub.sum())
    return out  # TL, TR, BL, BR


def correlate3(p, k):
    """cv2.filter2D(p, -1, k) generic path: float32 accumulate over non-zero taps in raster order,
    no FMA, REFLECT_101."""
    pp = np.pad(p, 1, mode="reflect")
    h, w = p.shape
    acc = None
    for a in range(3):
        for b in range(3):
            if k[a, b] == 0:
                continue
            term = f32(k[a, b]) * pp[a:a + h, b:b + w]
            acc = term if acc is None else acc + term
    return acc


def upsample(p, base_bottom_right, corr=correlate3):
    """4-phase photosite-aware Gaussian upsample of a quarter plane
    (debayer/edge_assisted_gaussian.py:140-141)."""
    k_tl, k_tr, k_bl, k_br = phase_kernels(base_bottom_right)
    return join_planes(corr(p, k_tl), corr(p, k_tr), corr(p, k_br), corr(p, k_bl))


def resample_channel(c, g_c, g_hf, base_bottom_right, corr=correlate3):
    """debayer/edge_assisted_gaussian.py:126-143; association as written:
    up(c - g) + (up(g) + hf)."""
    g_up = upsample(g_c, base_bottom_right, corr) + g_hf
    return upsample(c - g_c, base_bottom_right, corr) + g_up


def count_map(lab, vertical):
    """debayer/ahd_homogeneity_cython.pyx:22-58 on an edge-padded Lab image [H+2,W+2,3]."""
    L, A, B = lab[..., 0], lab[..., 1], lab[..., 2]
    c = (slice(1, -1), slice(1, -1))
    if vertical:
        n1 = (slice(0, -2), slice(1, -1))
        n2 = (slice(2, None), slice(1, -1))
    else:
        n1 = (slice(1, -1), slice(0, -2))
        n2 = (slice(1, -1), slice(2, None))

    def d2(s):
        da = A[c] - A[s]
        db = B[c] - B[s]
        return da * da + db * db

    eps_l = np.maximum(np.abs(L[c] - L[n1]), np.abs(L[c] - L[n2]))
    eps_c = np.maximum(d2(n1), d2(n2))
    h, w = L[c].shape
    cnt = np.zeros((h, w), dtype=f32)
    for dy in range(3):
        for dx in range(3):
            s = (slice(dy, dy + h), slice(dx, dx + w))
            da = A[s] - A[c]
            db = B[s] - B[c]
            ok = ((L[s] - L[c]) <= eps_l) & ((da * da + db * db) <= eps_c)
            cnt += ok.astype(f32)
    return cnt


def box3_sum(m):
    """3x3 window sum with REFLECT_101 (cv2.blur((3,3)) = this / 9; counts are small integers so the
    comparison map_h < map_v is the integer comparison of the sums)."""
    p = np.pad(m, 1, mode="reflect")
    h, w = m.shape
    acc = np.zeros((h, w), dtype=f32)
    for dy in range(3):
        for dx in range(3):
            acc += p[dy:dy + h, dx:dx + w]
    return acc


def median5(m):
    """cv2.medianBlur(m, 5): exact 5x5 median, REPLICATE border."""
    p = np.pad(m, 2, mode="edge")
    h, w = m.shape
    out = np.empty((h, w), dtype=m.dtype)
    step = max(1, (1 << 22) // max(w, 1))
    for y0 in range(0, h, step):
        y1 = min(h, y0 + step)
        st = np.stack([p[y0 + dy:y1 + dy, dx:dx + w] for dy in range(5) for dx in range(5)], axis=0)
        st.partition(12, axis=0)
        out[y0:y1] = st[12]
    return out


class _Cv2Backend:
    """Same library calls as the reference (timing baseline only)."""

    def __init__(self):
        import cv2
        self.cv2 = cv2

    def gauss3(self, g):
        return self.cv2.GaussianBlur(g, (3, 3), 1.0)

    def corr(self, p, k):
        return self.cv2.filter2D(p, -1, k)

    def lab(self, rgb):
        return self.cv2.cvtColor(rgb, self.cv2.COLOR_RGB2LAB)

    def median5(self, m):
        return self.cv2.medianBlur(m, 5)

    def mat3(self, rgb, m):
        return np.dot(rgb, m.T).astype(f32)

    def box3(self, m):
        return self.cv2.blur(m, (3, 3))


class _SpecBackend:
    gauss3 = staticmethod(gauss3)
    corr = staticmethod(correlate3)
    lab = staticmethod(lab_cv)
    median5 = staticmethod(median5)
    mat3 = staticmethod(mat3_f64)
    box3 = staticmethod(box3_sum)


# ----------------------------------------------------------------------------------------------
# the develop path
# ----------------------------------------------------------------------------------------------
def ahd_demosaic(sensor, wb, m_cam_to_srgb, stages=1, hdr=False, backend="spec", count_fn=None,
                 keep=False):
    """debayer/ahd.py:14-170 on an RGGB float32 mosaic.  Returns camera-RGB float32 [H,W,3]
    (white balance applied once); with keep=True also a dict of intermediates."""
    be = _Cv2Backend() if backend == "cv2" else _SpecBackend()
    sensor = np.asarray(sensor, dtype=f32)
    wb = np.asarray(wb, dtype=f32)
    H, W = sensor.shape
    assert H % 2 == 0 and W % 2 == 0 and H >= 4 and W >= 4
    r, g1, b, g2 = split_planes(sensor)
    pad = lambda a: np.pad(a, 1, mode="edge")                      # ahd.py:77-80 (BORDER_REFLECT, 1 px)
    r, g1, b, g2 = pad(r) * wb[0], pad(g1) * wb[1], pad(b) * wb[2], pad(g2) * wb[1]
    h = H5
    I = slice(1, -1)
    # ahd.py:97-102
    gh_r = (r[I, :-2] * h[0]) + (g1[I, :-2] * h[1]) + (r[I, I] * h[2]) + (g1[I, I] * h[3]) + (r[I, 2:] * h[4])
    gv_r = (r[:-2, I] * h[0]) + (g2[:-2, I] * h[1]) + (r[I, I] * h[2]) + (g2[I, I] * h[3]) + (r[2:, I] * h[4])
    gh_b = (b[I, :-2] * h[0]) + (g2[I, I] * h[1]) + (b[I, I] * h[2]) + (g2[I, 2:] * h[3]) + (b[I, 2:] * h[4])
    gv_b = (b[:-2, I] * h[0]) + (g1[I, I] * h[1]) + (b[I, I] * h[2]) + (g1[2:, I] * h[3]) + (b[2:, I] * h[4])
    rw, g1w, bw, g2w = r[I, I], g1[I, I], b[I, I], g2[I, I]
    g_h = join_planes(gh_r, g1w, gh_b, g2w)                        # ahd.py:105-106
    g_v = join_planes(gv_r, g1w, gv_b, g2w)
    hf_h = g_h - be.gauss3(g_h)                                    # ahd.py:120-121
    hf_v = g_v - be.gauss3(g_v)
    r_h = resample_channel(rw, gh_r, hf_h, False, be.corr)         # ahd.py:123-127
    r_v = resample_channel(rw, gv_r, hf_v, False, be.corr)
    b_h = resample_channel(bw, gh_b, hf_h, True, be.corr)
    b_v = resample_channel(bw, gv_b, hf_v, True, be.corr)

    def homogeneity(rr, gg, bb, vertical):                         # ahd.py:32-67
        rgb = be.mat3(np.dstack((rr * wb[0], gg * wb[1], bb * wb[2])), m_cam_to_srgb)
        if hdr:
            luma = 0.2126 * rgb[:, :, 0] + 0.7152 * rgb[:, :, 1] + 0.0722 * rgb[:, :, 2]
            lab = be.lab(rgb / (1 + rgb))
            lab[:, :, 0] = luma
        else:
            lab = be.lab(rgb)
        labp = np.pad(lab, ((1, 1), (1, 1), (0, 0)), mode="edge")
        cm = (count_fn or count_map)(labp, vertical)
        return cm, lab

    cnt_h, lab_h = homogeneity(r_h, g_h, b_h, False)
    cnt_v, lab_v = homogeneity(r_v, g_v, b_v, True)
    sum_h = be.box3(cnt_h)                                         # ahd.py:133-134
    sum_v = be.box3(cnt_v)
    pick_h = sum_h < sum_v                                         # ahd.py:136-145 (ties -> V)
    # ahd.py:139-145: a multiplicative blend, not a select -- h*c + v*(1-c) with c in {0, 1}.  For finite candidates it
    # equals the chosen one; a non-finite value in the candidate that is NOT chosen still turns the pixel into NaN
    # (inf * 0), and so does a non-finite native green (both candidates hold it).
    comb = pick_h.astype(f32)[..., None]
    with np.errstate(invalid="ignore"):
        rgb = (np.dstack((r_h, g_h, b_h)).astype(f32) * comb + np.dstack((r_v, g_v, b_v)).astype(f32) * (1 - comb)).astype(f32)
    selected = rgb
    for _ in range(max(int(stages), 0)):                           # ahd.py:148-165
        rr, gg, bb = rgb[..., 0], rgb[..., 1], rgb[..., 2]
        rr = be.median5(rr - gg) + gg
        bb = be.median5(bb - gg) + gg
        gg = (be.median5(gg - rr) + be.median5(gg - bb) + rr + bb) / 2
        rgb = np.dstack((rr, gg, bb)).astype(f32)
    if keep:
        return rgb, dict(g_h=g_h, g_v=g_v, hf_h=hf_h, hf_v=hf_v, r_h=r_h, r_v=r_v, b_h=b_h, b_v=b_v,
                         lab_h=lab_h, lab_v=lab_v, cnt_h=cnt_h, cnt_v=cnt_v, sum_h=sum_h, sum_v=sum_v,
                         pick_h=pick_h, selected=selected)
    return rgb


def to_lin_srgb(cam_rgb, m_cam_to_srgb, backend="spec"):
    """base_types/image_base.py:62-64 -> colorize/transform.py:37-53 (clip to [0,1], float64 3x3)."""
    be = _Cv2Backend() if backend == "cv2" else _SpecBackend()
    return be.mat3(np.clip(cam_rgb, f32(0), f32(1)).astype(f32), m_cam_to_srgb)


def lin_srgb_to_srgb(rgb):
    """colorize/transform.py:89-99, float32."""
    x = np.clip(np.asarray(rgb, dtype=f32), f32(0), f32(1))
    return np.where(x <= 0.0031308, x * 12.92, (1.055 * (x ** (1 / 2.4))) - 0.055).astype(f32)


_FLIPS = {"RGGB": (False, False), "BGGR": (True, True), "GBRG": (False, True), "GRBG": (True, False)}


def to_rggb(a, pattern):
    """image.py:143-152 (self-inverse): RGGB identity, BGGR rot180, GBRG flip x, GRBG flip y."""
    fy, fx = _FLIPS[pattern]
    if fy:
        a = a[::-1]
    if fx:
        a = a[:, ::-1]
    return a


def develop(raw_u16, black, white, wb, mat_xyz_to_cam, white_xyz, stages=1, pattern="RGGB",
            hdr=False, backend="spec", keep=False):
    """u16 mosaic -> linear sRGB f32 [H,W,3]: RawBayerData.demosaic(Best, stages).to_lin_srgb()
    (image.py:191-197, 156-183; base_types/image_base.py:62-64)."""
    m = cam_to_lin_srgb_matrix(mat_xyz_to_cam, white_xyz)
    sensor = normalize(raw_u16, black, white)
    res = ahd_demosaic(to_rggb(sensor, pattern), wb, m, stages, hdr, backend, keep=keep)
    cam, extra = res if keep else (res, None)
    cam = np.ascontiguousarray(to_rggb(cam, pattern))
    lin = to_lin_srgb(cam, m, backend)
    return (lin, cam, extra) if keep else (lin, cam)


def fuse_exposures(brackets, evs, wb, target_ev=None):
    """raw_hdr.py:108-148: EV-aligned, saturation- and WB-weighted per-photosite average of the
    brackets, accumulated in list order in float32; brightest-frame fallback where the weights sum to
    zero.  Returns (hdr_mosaic f32, contribution count int32, max_ev_offset)."""
    brackets = [np.asarray(b, dtype=f32) for b in brackets]
    wb = np.asarray(wb, dtype=f32)
    evs = [float(e) for e in evs]            # Python floats: NumPy weak-scalar promotion keeps float32
    if target_ev is None:
        target_ev = 0
        for e in evs:
            target_ev += e
        target_ev /= len(evs)
    offs = [2 ** (e - target_ev) for e in evs]
    H, W = brackets[0].shape
    ones = np.ones((H // 2, W // 2), dtype=f32)
    nw = join_planes(ones * wb[0], ones * wb[1], ones * wb[2], ones * wb[1])
    sum_p = np.zeros_like(brackets[0])
    sum_w = np.zeros_like(brackets[0])
    cnt = np.zeros((H, W), dtype=np.int32)
    for x, off in zip(brackets, offs):
        bias = 1.6 ** (-0.1 * np.abs(off * nw))
        wgt = (0.5 - np.abs(x - 0.5)) * bias
        sum_w += wgt
        sum_p += x * wgt * off
        cnt[wgt > 0] += 1
    imax = int(np.argmax(offs))
    brightest = np.multiply(brackets[imax], offs[imax])
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.divide(sum_p, sum_w)
    out = np.where(sum_w == 0, brightest, q)
    return out.astype(f32), cnt, max(offs), target_ev


# ----------------------------------------------------------------------------------------------
# QualityDemosaic.Fast: edge-assisted Gaussian (debayer/edge_assisted_gaussian.py:10-201)
# ----------------------------------------------------------------------------------------------
def delta_mix(top, bottom, left, right):
    """edge_assisted_gaussian.py:10-49: bilinear in-fill weighted towards the direction of least change."""
    dy = np.abs(top - bottom)
    dx = np.abs(left - right)
    total = dy + dx
    avg_x = (left + right) / 2
    avg_y = (top + bottom) / 2
    sy = np.divide(dy, total, out=np.ones_like(total) * 0.5, where=total != 0)
    sx = 1 - sy
    return avg_y * sx + avg_x * sy


def eag_demosaic(sensor, wb, backend="spec"):
    """debayer_eag (edge_assisted_gaussian.py:188-201) on an RGGB float32 mosaic -> camera RGB float32."""
    be = _Cv2Backend() if backend == "cv2" else _SpecBackend()
    sensor = np.asarray(sensor, dtype=f32)
    wb = np.asarray(wb, dtype=f32)
    r, g1, b, g2 = split_planes(sensor)
    p1, p2 = np.pad(g1, 1, mode="edge"), np.pad(g2, 1, mode="edge")
    I = slice(1, -1)
    g_at_b = delta_mix(p1[I, I], p1[2:, I], p2[I, I], p2[I, 2:])          # l.95-98
    g_at_r = delta_mix(p2[:-2, I], p2[I, I], p1[I, :-2], p1[I, I])        # l.101-104
    g_up = join_planes(g_at_r, g1, g_at_b, g2) * wb[1]                     # l.124, 193
    rw, bw = r * wb[0], b * wb[2]
    hf = g_up - be.gauss3(g_up)                                            # l.157
    g_r, _, g_b, _ = split_planes(g_up)
    r_up = resample_channel(rw, g_r, hf, False, be.corr)
    b_up = resample_channel(bw, g_b, hf, True, be.corr)
    return np.dstack((r_up, g_up, b_up)).astype(f32)


def develop_fast(raw_u16, black, white, wb, mat_xyz_to_cam, white_xyz, pattern="RGGB", backend="spec"):
    """RawBayerData.demosaic(QualityDemosaic.Fast).to_lin_srgb() (image.py:171-172)."""
    m = cam_to_lin_srgb_matrix(mat_xyz_to_cam, white_xyz)
    cam = eag_demosaic(to_rggb(normalize(raw_u16, black, white), pattern), wb, backend)
    cam = np.ascontiguousarray(to_rggb(cam, pattern))
    return to_lin_srgb(cam, m, backend), cam
