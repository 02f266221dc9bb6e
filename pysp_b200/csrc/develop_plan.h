// Host-side planning of one pysp_develop call: argument validation, row ranges of every kernel in the
// chain (K1 select, then one K2 median launch per stage, each shrinking the band by 4 rows per side),
// scratch ping-pong.  Pure C++ (no CUDA calls) so that tests/host_emu can drive the same plan.
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pysp_b200.h"
#include "pysp_common.cuh"

namespace pysp {

#define PYSP_MAX_STAGES 16

struct DevelopPlan {
    SelectParams select;
    int select_tiles;
    int n_stages;
    MedianParams median[PYSP_MAX_STAGES];
    int median_tiles[PYSP_MAX_STAGES];
};

static inline int plan_fail(char* err, size_t n, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, n, fmt, ap);
    va_end(ap);
    return code;
}

static inline int64_t develop_scratch_bytes(int32_t width, int32_t rows, int32_t stages) {
    if (stages <= 0) return 0;
    int64_t one = (int64_t)(rows + 8 * stages) * width * 3 * (int64_t)sizeof(float);
    return stages >= 2 ? 2 * one : one;
}

static inline int plan_develop(const pysp_develop_args* a, int tw1, int th1, int tw2, int th2, DevelopPlan* plan,
                               char* err, size_t errn) {
    if (!a) return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: null args");
    const int H = a->height, W = a->width;
    if (H < 4 || W < 4 || (H & 1) || (W & 1))
        return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: frame %dx%d must be even and >= 4x4", H, W);
    int flip_y = 0, flip_x = 0;
    switch (a->cfa_pattern) {                 // image.py:143-152
        case PYSP_CFA_RGGB: break;
        case PYSP_CFA_BGGR: flip_y = flip_x = 1; break;
        case PYSP_CFA_GBRG: flip_x = 1; break;
        case PYSP_CFA_GRBG: flip_y = 1; break;
        default:
            return plan_fail(err, errn, PYSP_ERR_UNSUPPORTED, "pysp_develop: CFA pattern %d not implemented", a->cfa_pattern);
    }
    if (a->in_kind != PYSP_IN_U16 && a->in_kind != PYSP_IN_F32)
        return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: bad in_kind %d", a->in_kind);
    if (a->out_kind < PYSP_OUT_CAM_F32 || a->out_kind > PYSP_OUT_LIN_F16)
        return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: bad out_kind %d", a->out_kind);
    if (!a->in || !a->out || !a->lab_lut) return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: null buffer");
    const int rb = a->row_begin, re = a->row_end;
    if (rb < 0 || re > H || rb >= re || (rb & 1) || (re & 1))
        return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: row band [%d,%d) must be even and inside the frame", rb, re);
    if (a->out_row0 > rb)
        return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: out_row0 %d is past row_begin %d", a->out_row0, rb);
    const int S = a->stages > 0 ? a->stages : 0;                 // debayer/ahd.py:163
    if (S > PYSP_MAX_STAGES)
        return plan_fail(err, errn, PYSP_ERR_UNSUPPORTED, "pysp_develop: at most %d postprocess stages", PYSP_MAX_STAGES);
    const int64_t esz = a->in_kind == PYSP_IN_U16 ? 2 : 4;
    if (a->in_pitch_bytes < W * esz || (a->in_pitch_bytes % esz))
        return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: bad in_pitch_bytes");
    const int64_t osz = a->out_kind == PYSP_OUT_LIN_F16 ? 2 : 4;
    if (a->out_pitch_bytes < 3 * W * osz || (a->out_pitch_bytes % osz))
        return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: bad out_pitch_bytes");
    // logical (RGGB-oriented) rows of the band
    const int lb = flip_y ? H - re : rb, le = flip_y ? H - rb : re;
    auto lo = [&](int v) { return v < 0 ? 0 : v; };
    auto hi = [&](int v) { return v > H ? H : v; };
    const int k1b = lo(lb - 4 * S), k1e = hi(le + 4 * S);
    {   // the input buffer must hold the band + halo
        const int nb = lo(k1b - 6), ne = hi(k1e + 6);
        const int sb = flip_y ? H - ne : nb, se = flip_y ? H - nb : ne;
        if (a->in_row0 > sb || a->in_row0 + a->in_rows < se)
            return plan_fail(err, errn, PYSP_ERR_INVALID,
                             "pysp_develop: input holds rows [%d,%d) but rows [%d,%d) are needed (halo %d)", a->in_row0,
                             a->in_row0 + a->in_rows, sb, se, 6 + 4 * S);
    }
    const int64_t need = develop_scratch_bytes(W, re - rb, S);
    if (S > 0 && (!a->scratch || a->scratch_bytes < need))
        return plan_fail(err, errn, PYSP_ERR_INVALID, "pysp_develop: scratch of %lld bytes needed", (long long)need);

    ColorParams c;
    memset(&c, 0, sizeof(c));
    for (int i = 0; i < 3; ++i) c.wb[i] = a->wb[i];
    for (int i = 0; i < 9; ++i) c.m_metric[i] = c.m_out[i] = a->cam_to_srgb[i];
    c.hdr = a->is_hdr ? 1 : 0;
    c.gamma = a->apply_gamma ? 1 : 0;
    FrameGeom g = {H, W, flip_y, flip_x};

    const int64_t scratch_rows = (re - rb) + 8 * S;
    float* sbuf[2] = {(float*)a->scratch, S >= 2 ? (float*)a->scratch + scratch_rows * W * 3 : nullptr};
    const long long spitch = (long long)W * 3 * sizeof(float);

    memset(plan, 0, sizeof(*plan));
    SelectParams& sp = plan->select;
    sp.g = g; sp.c = c;
    sp.in_kind = a->in_kind; sp.in = a->in; sp.in_pitch = a->in_pitch_bytes;
    sp.in_row0 = a->in_row0; sp.in_row1 = a->in_row0 + a->in_rows;
    const int perm[4] = {0, 1, 3, 2};          // [TL,TR,BR,BL] -> index (sy&1)*2+(sx&1)
    for (int i = 0; i < 4; ++i) { sp.black[perm[i]] = a->black[i]; sp.white[perm[i]] = a->white[i]; }
    sp.lut = (const uint2*)a->lab_lut;
    sp.y_begin = k1b; sp.y_end = k1e;
    sp.tiles_x = (W + tw1 - 1) / tw1;
    if (S == 0) {
        sp.out_kind = a->out_kind; sp.out = a->out; sp.out_pitch = a->out_pitch_bytes; sp.out_row0 = a->out_row0;
        sp.store_flip = 1;
    } else {
        sp.out_kind = OUT_CAM_F32; sp.out = sbuf[0]; sp.out_pitch = spitch; sp.out_row0 = k1b; sp.store_flip = 0;
    }
    plan->select_tiles = sp.tiles_x * ((k1e - k1b + th1 - 1) / th1);
    plan->n_stages = S;
    int prev_b = k1b, prev_e = k1e;
    for (int s = 1; s <= S; ++s) {
        MedianParams& mp = plan->median[s - 1];
        mp.g = g; mp.c = c;
        mp.in = sbuf[(s - 1) & 1]; mp.in_pitch = spitch; mp.in_row0 = prev_b; mp.in_row1 = prev_e;
        mp.y_begin = lo(lb - 4 * (S - s)); mp.y_end = hi(le + 4 * (S - s));
        mp.tiles_x = (W + tw2 - 1) / tw2;
        if (s == S) {
            mp.out_kind = a->out_kind; mp.out = a->out; mp.out_pitch = a->out_pitch_bytes; mp.out_row0 = a->out_row0;
            mp.store_flip = 1;
        } else {
            mp.out_kind = OUT_CAM_F32; mp.out = sbuf[s & 1]; mp.out_pitch = spitch; mp.out_row0 = mp.y_begin; mp.store_flip = 0;
        }
        plan->median_tiles[s - 1] = mp.tiles_x * ((mp.y_end - mp.y_begin + th2 - 1) / th2);
        prev_b = mp.y_begin; prev_e = mp.y_end;
    }
    return PYSP_OK;
}

}  // namespace pysp
