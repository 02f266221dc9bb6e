// K2: one chroma-cleanup stage of AHD (debayer/ahd.py:148-161) on camera RGB:
//     r' = med5(r-g)+g ; b' = med5(b-g)+g ; g' = (((med5(g-r') + med5(g-b')) + r') + b') / 2
// cv2.medianBlur(f32, 5): exact 5x5 selection, BORDER_REPLICATE.
//
// One CTA produces a TW x TH tile from the tile + 4 px: phase A forms the two colour-difference planes
// in shared memory, phase B takes their medians on the tile + 2 px and forms the second pair of
// difference planes, phase C takes those medians on the tile and runs the output epilogue (clip, float64
// camera->linear-sRGB matrix, optional gamma) when this is the last stage, so the final image is written
// exactly once.
#pragma once
#include "pysp_common.cuh"

namespace pysp {

#define PYSP_CE(a, b) { float lo_ = fminf(p[a], p[b]); p[b] = fmaxf(p[a], p[b]); p[a] = lo_; }
// Median of 25 by a 99-comparator selection network (exhaustively checked with the 0-1 principle,
// tools/verify_median_network.c).  Only min/max: the result is one of the inputs, as in cv2.
PYSP_HD float median25(float p[25]) {
    PYSP_CE(0, 1) PYSP_CE(3, 4) PYSP_CE(2, 4) PYSP_CE(2, 3) PYSP_CE(6, 7) PYSP_CE(5, 7) PYSP_CE(5, 6)
    PYSP_CE(9, 10) PYSP_CE(8, 10) PYSP_CE(8, 9) PYSP_CE(12, 13) PYSP_CE(11, 13) PYSP_CE(11, 12)
    PYSP_CE(15, 16) PYSP_CE(14, 16) PYSP_CE(14, 15) PYSP_CE(18, 19) PYSP_CE(17, 19) PYSP_CE(17, 18)
    PYSP_CE(21, 22) PYSP_CE(20, 22) PYSP_CE(20, 21) PYSP_CE(23, 24) PYSP_CE(2, 5) PYSP_CE(3, 6)
    PYSP_CE(0, 6) PYSP_CE(0, 3) PYSP_CE(4, 7) PYSP_CE(1, 7) PYSP_CE(1, 4) PYSP_CE(11, 14) PYSP_CE(8, 14)
    PYSP_CE(8, 11) PYSP_CE(12, 15) PYSP_CE(9, 15) PYSP_CE(9, 12) PYSP_CE(13, 16) PYSP_CE(10, 16)
    PYSP_CE(10, 13) PYSP_CE(20, 23) PYSP_CE(17, 23) PYSP_CE(17, 20) PYSP_CE(21, 24) PYSP_CE(18, 24)
    PYSP_CE(18, 21) PYSP_CE(19, 22) PYSP_CE(8, 17) PYSP_CE(9, 18) PYSP_CE(0, 18) PYSP_CE(0, 9)
    PYSP_CE(10, 19) PYSP_CE(1, 19) PYSP_CE(1, 10) PYSP_CE(11, 20) PYSP_CE(2, 20) PYSP_CE(2, 11)
    PYSP_CE(12, 21) PYSP_CE(3, 21) PYSP_CE(3, 12) PYSP_CE(13, 22) PYSP_CE(4, 22) PYSP_CE(4, 13)
    PYSP_CE(14, 23) PYSP_CE(5, 23) PYSP_CE(5, 14) PYSP_CE(15, 24) PYSP_CE(6, 24) PYSP_CE(6, 15)
    PYSP_CE(7, 16) PYSP_CE(7, 19) PYSP_CE(13, 21) PYSP_CE(15, 23) PYSP_CE(7, 13) PYSP_CE(7, 15)
    PYSP_CE(1, 9) PYSP_CE(3, 11) PYSP_CE(5, 17) PYSP_CE(11, 17) PYSP_CE(9, 17) PYSP_CE(4, 10)
    PYSP_CE(6, 12) PYSP_CE(7, 14) PYSP_CE(4, 6) PYSP_CE(4, 7) PYSP_CE(12, 14) PYSP_CE(10, 14)
    PYSP_CE(6, 7) PYSP_CE(10, 12) PYSP_CE(6, 10) PYSP_CE(6, 17) PYSP_CE(12, 17) PYSP_CE(7, 17)
    PYSP_CE(7, 10) PYSP_CE(12, 18) PYSP_CE(7, 12) PYSP_CE(10, 18) PYSP_CE(12, 20) PYSP_CE(10, 20)
    PYSP_CE(10, 12)
    return p[12];
}

template <int TW_, int TH_>
struct MedianTile {
    static constexpr int TW = TW_, TH = TH_;
    static constexpr int AW = TW + 8, AH = TH + 8;     // input region (tile + 4)
    static constexpr int BW = TW + 4, BH = TH + 4;     // first-median region (tile + 2)
    static constexpr int OFF_DR = 0;                   // [AH][AW] r-g
    static constexpr int OFF_DB = OFF_DR + AH * AW;    // [AH][AW] b-g
    static constexpr int OFF_G = OFF_DB + AH * AW;     // [AH][AW] g
    static constexpr int OFF_ER = OFF_G + AH * AW;     // [BH][BW] g-r'
    static constexpr int OFF_EB = OFF_ER + BH * BW;    // [BH][BW] g-b'
    static constexpr int OFF_RP = OFF_EB + BH * BW;    // [TH][TW] r'
    static constexpr int OFF_BP = OFF_RP + TH * TW;    // [TH][TW] b'
    static constexpr int SMEM_BYTES = (OFF_BP + TH * TW) * 4;
};

template <int TW, int TH, bool EDGE>
PYSP_D void median_tile(const MedianParams& p, float* __restrict__ sm, int tile_x, int tile_y) {
    typedef MedianTile<TW, TH> L;
    const int H = p.g.H, W = p.g.W;
    const int x0 = tile_x * TW, y0 = p.y_begin + tile_y * TH;
    float* DR = sm + L::OFF_DR; float* DB = sm + L::OFF_DB; float* G = sm + L::OFF_G;
    float* ER = sm + L::OFF_ER; float* EB = sm + L::OFF_EB;
    float* RP = sm + L::OFF_RP; float* BP = sm + L::OFF_BP;

    // ---- phase A: colour differences on the tile + 4 px (REPLICATE at the frame border) -----------------
    PYSP_ITEMS(it, L::AW * L::AH) {
        int ly = it / L::AW, lx = it - ly * L::AW;
        int y = y0 - 4 + ly, x = x0 - 4 + lx;
        float r = 0.f, g = 0.f, b = 0.f;
        bool ok = true;
        if (EDGE) {
            y = clampi(y, H); x = clampi(x, W);
            ok = y >= p.in_row0 && y < p.in_row1;       // rows this band's buffer does not hold are never consumed
        }
        if (ok) {
            const float* src = (const float*)((const char*)p.in + (long long)(y - p.in_row0) * p.in_pitch) + 3 * (long long)x;
            r = pysp_ldg(src); g = pysp_ldg(src + 1); b = pysp_ldg(src + 2);
        }
        DR[it] = r - g; DB[it] = b - g; G[it] = g;
    }
    PYSP_SYNC();

    // ---- phase B: r', b' and the second difference planes on the tile + 2 px ----------------------------
    PYSP_ITEMS(it, L::BW * L::BH) {
        int ly = it / L::BW, lx = it - ly * L::BW;
        int y = y0 - 2 + ly, x = x0 - 2 + lx;
        if (EDGE) { if (y < 0 || y >= H || x < 0 || x >= W) continue; }
        float wr[25], wb[25];
        // the input region was filled through the clamp, so a window taken around an in-frame pixel is
        // already the REPLICATE window
        int c = (ly + 2) * L::AW + lx + 2;
#pragma unroll
        for (int u = 0; u < 5; ++u)
#pragma unroll
            for (int v = 0; v < 5; ++v) {
                int o = c + (u - 2) * L::AW + (v - 2);
                wr[u * 5 + v] = DR[o]; wb[u * 5 + v] = DB[o];
            }
        float g = G[c];
        float r1 = median25(wr) + g;
        float b1 = median25(wb) + g;
        ER[it] = g - r1; EB[it] = g - b1;
        int ty = ly - 2, tx = lx - 2;
        if (ty >= 0 && ty < TH && tx >= 0 && tx < TW) { RP[ty * TW + tx] = r1; BP[ty * TW + tx] = b1; }
    }
    PYSP_SYNC();

    // ---- phase C: g' on the tile, epilogue, store --------------------------------------------------------
    PYSP_ITEMS(it, TW * TH) {
        int ty = it / TW, tx = it - ty * TW;
        int y = y0 + ty, x = x0 + tx;
        if (y >= p.y_end || x >= W) continue;
        float wr[25], wb[25];
#pragma unroll
        for (int u = 0; u < 5; ++u)
#pragma unroll
            for (int v = 0; v < 5; ++v) {
                int yy = EDGE ? clampi(y + u - 2, H) - (y0 - 2) : ty + u;
                int xx = EDGE ? clampi(x + v - 2, W) - (x0 - 2) : tx + v;
                wr[u * 5 + v] = ER[yy * L::BW + xx]; wb[u * 5 + v] = EB[yy * L::BW + xx];
            }
        float r1 = RP[it], b1 = BP[it];
        Rgb v;
        v.r = r1; v.b = b1;
        v.g = (((median25(wr) + median25(wb)) + r1) + b1) / 2.0f;
        v = finish_pixel(p.c, p.out_kind, v);
        int yy = y, xx = x;
        if (p.store_flip) {
            if (p.g.flip_y) yy = H - 1 - yy;
            if (p.g.flip_x) xx = W - 1 - xx;
        }
        char* row = (char*)p.out + (long long)(yy - p.out_row0) * p.out_pitch;
        if (p.out_kind == OUT_LIN_F16) {
#ifndef PYSP_HOST_EMU
            __half* o16 = (__half*)row + 3 * (long long)xx;
            o16[0] = __float2half_rn(v.r); o16[1] = __float2half_rn(v.g); o16[2] = __float2half_rn(v.b);
#endif
        } else {
            float* o32 = (float*)row + 3 * (long long)xx;
            o32[0] = v.r; o32[1] = v.g; o32[2] = v.b;
        }
    }
}

}  // namespace pysp
