// K2: one chroma-cleanup stage of AHD (debayer/ahd.py:148-161) on camera RGB:
//     r' = med5(r-g)+g ; b' = med5(b-g)+g ; g' = (((med5(g-r') + med5(g-b')) + r') + b') / 2
// cv2.medianBlur(f32, 5): exact 5x5 selection, BORDER_REPLICATE.
//
// The previous kernel of the chain hands over three float planes r-g, b-g, g (see StoreParams), so the
// tile + 4 px box of each plane is fetched by TMA straight into the shared-memory planes the medians read
// (the next tile's boxes are prefetched while the current tile is in phase C).  Phase B takes the medians
// on the tile + 2 px and forms the second pair of difference planes, phase C takes those medians on the tile
// and runs the output epilogue (clip, float64 camera->linear-sRGB matrix, optional gamma) when this is the last
// stage, so the final image is written exactly once.
#pragma once
#include "pysp_common.cuh"
#include "ahd_select.cuh"   // stage_pixel / store_tile_generic / tile_output_box
#include "tma.cuh"
#include "median_block.cuh"
#include "median_block2x4.cuh"

namespace pysp {

template <int TW_, int TH_>
struct MedianTile {
    static constexpr int TW = TW_, TH = TH_;
    static constexpr int AW = TW + 8, AH = TH + 8;     // input box (tile + 4)
    static constexpr int BW = TW + 4, BH = TH + 4;     // first-median region (tile + 2)
    static constexpr int align128(int v) { return (v + 127) / 128 * 128; }
    static constexpr int PLANE_BYTES = align128(AH * AW * 4);
    static constexpr int OFF_BAR = 0;
    static constexpr int OFF_IN = 128;                                  // 3 x [AH][AW] f32: r-g, b-g, g (TMA destination)
    static constexpr int OFF_ER = align128(OFF_IN + 3 * PLANE_BYTES);   // [BH][BW] g-r'
    static constexpr int OFF_EB = OFF_ER + BH * BW * 4;                 // [BH][BW] g-b'
    static constexpr int OFF_RP = OFF_EB + BH * BW * 4;                 // [TH][TW] r'
    static constexpr int OFF_BP = OFF_RP + TH * TW * 4;                 // [TH][TW] b'
    static constexpr int OFF_OUT = align128(OFF_BP + TH * TW * 4);      // [3][TH][TW] output staging tile
    static constexpr int SMEM_BYTES = align128(OFF_OUT + 3 * align128(TH * TW * 4));
    static_assert(TW % 4 == 0, "16-byte rows");
};

template <int TW, int TH>
PYSP_HD bool median_tile_is_edge(const MedianParams& p, int tile_x, int tile_y) {
    const int x0 = tile_x * TW, y0 = p.y_begin + tile_y * TH;
    return x0 < 4 || y0 < 4 || x0 + TW + 4 > p.g.W || y0 + TH + 4 > p.g.H || y0 + TH > p.y_end;
}

// phase A (edge tiles only): the box arrives zero-filled outside the frame; REPLICATE needs the value of the
// clamped in-frame cell, which is always inside the same box
template <int TW, int TH>
PYSP_D void median_fix_border(const MedianParams& p, char* __restrict__ smem, int tile_x, int tile_y) {
    typedef MedianTile<TW, TH> L;
    const int H = p.g.H, W = p.g.W;
    const int bx0 = tile_x * TW - 4, by0 = p.y_begin + tile_y * TH - 4;
    PYSP_ITEMS(it, L::AW * L::AH) {
        int ly = it / L::AW, lx = it - ly * L::AW;
        int y = by0 + ly, x = bx0 + lx;
        if (y >= 0 && y < H && x >= 0 && x < W) continue;
        int src = (clampi(y, H) - by0) * L::AW + (clampi(x, W) - bx0);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float* pl = (float*)(smem + L::OFF_IN + k * L::PLANE_BYTES);
            pl[it] = pl[src];
        }
    }
}

// phase B: r', b' and the second difference planes on the tile + 2 px.  One work item = a 2x2 block of pixels whose
// four 5x5 windows share a 6x6 neighbourhood (median_block.cuh: 67.5 min/max ops per median instead of 180).
template <int TW, int TH, bool EDGE>
PYSP_D void median_phase_b(const MedianParams& p, char* __restrict__ smem, int tile_x, int tile_y) {
    typedef MedianTile<TW, TH> L;
    const int H = p.g.H, W = p.g.W;
    const int x0 = tile_x * TW, y0 = p.y_begin + tile_y * TH;
    const float* DR = (const float*)(smem + L::OFF_IN);
    const float* DB = (const float*)(smem + L::OFF_IN + L::PLANE_BYTES);
    const float* G = (const float*)(smem + L::OFF_IN + 2 * L::PLANE_BYTES);
    float* ER = (float*)(smem + L::OFF_ER); float* EB = (float*)(smem + L::OFF_EB);
    float* RP = (float*)(smem + L::OFF_RP); float* BP = (float*)(smem + L::OFF_BP);
    constexpr int NBX = L::BW / 2, NBY = L::BH / 2;
    PYSP_ITEMS(it, NBX * NBY) {
        int by = it / NBX, bx = it - by * NBX;
        int ly = 2 * by, lx = 2 * bx;                       // region-B coords of the block's top-left pixel
        int y = y0 - 2 + ly, x = x0 - 2 + lx;
        if (EDGE) { if (y < 0 || y >= H || x < 0 || x >= W || y >= p.y_end + 2) continue; }     // even origin, even frame: all in or all out
        // the input planes hold the REPLICATE extension, so the windows of in-frame pixels are final
        float w[6][6], mr[4], mb[4];
        const int c = ly * L::AW + lx;                      // input-plane index of window cell (0,0) = pixel (y-2, x-2)
        // lx and AW are even: a window row is three aligned 8-byte pairs (32-bit loads would have lane stride 2: two-way
        // bank conflicts on every load)
#pragma unroll
        for (int u = 0; u < 6; ++u)
#pragma unroll
            for (int v = 0; v < 3; ++v) { const F2 t = ((const F2*)(DR + c + u * L::AW))[v]; w[u][2 * v] = t.x; w[u][2 * v + 1] = t.y; }
        median25_block2x2(w, mr);
#pragma unroll
        for (int u = 0; u < 6; ++u)
#pragma unroll
            for (int v = 0; v < 3; ++v) { const F2 t = ((const F2*)(DB + c + u * L::AW))[v]; w[u][2 * v] = t.x; w[u][2 * v + 1] = t.y; }
        median25_block2x2(w, mb);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const F2 g2 = *(const F2*)(G + c + (2 + dy) * L::AW + 2);
            F2 r1, b1, er, eb;
            r1.x = mr[2 * dy] + g2.x; r1.y = mr[2 * dy + 1] + g2.y;
            b1.x = mb[2 * dy] + g2.x; b1.y = mb[2 * dy + 1] + g2.y;
            er.x = g2.x - r1.x; er.y = g2.y - r1.y; eb.x = g2.x - b1.x; eb.y = g2.y - b1.y;
            const int o = (ly + dy) * L::BW + lx;
            *(F2*)(ER + o) = er; *(F2*)(EB + o) = eb;
            const int ty = ly + dy - 2, tx = lx - 2;         // even tx: the pair is inside the tile or outside as a whole
            if (ty >= 0 && ty < TH && tx >= 0 && tx < TW) { *(F2*)(RP + ty * TW + tx) = r1; *(F2*)(BP + ty * TW + tx) = b1; }
        }
    }
}

// ---- 2x4-block variants (median_block2x4.cuh: 60.75 min/max per median).  A work item is a 2x4 block of pixels inside its 6x8
// window; rows are 16-byte aligned (block origin and plane widths are multiples of 4), so a window row is two LDS.128.
struct __attribute__((aligned(16))) F4 { float x, y, z, w; };

template <int TW, int TH, bool EDGE>
PYSP_D void median_phase_b4(const MedianParams& p, char* __restrict__ smem, int tile_x, int tile_y) {
    typedef MedianTile<TW, TH> L;
    static_assert(L::BW % 4 == 0 && L::AW % 4 == 0, "2x4 blocks need plane widths that are multiples of 4");
    const int H = p.g.H, W = p.g.W;
    const int x0 = tile_x * TW, y0 = p.y_begin + tile_y * TH;
    const float* DR = (const float*)(smem + L::OFF_IN);
    const float* DB = (const float*)(smem + L::OFF_IN + L::PLANE_BYTES);
    const float* G = (const float*)(smem + L::OFF_IN + 2 * L::PLANE_BYTES);
    float* ER = (float*)(smem + L::OFF_ER); float* EB = (float*)(smem + L::OFF_EB);
    float* RP = (float*)(smem + L::OFF_RP); float* BP = (float*)(smem + L::OFF_BP);
    constexpr int NBX = L::BW / 4, NBY = L::BH / 2;
    PYSP_ITEMS(it, NBX * NBY) {
        int by = it / NBX, bx = it - by * NBX;
        int ly = 2 * by, lx = 4 * bx;                       // region-B coords of the block's top-left pixel
        int y = y0 - 2 + ly, x = x0 - 2 + lx;
        if (EDGE) { if (y < 0 || y >= H || x + 3 < 0 || x >= W || y >= p.y_end + 2) continue; }     // the whole block outside the frame / band
        float w[6][8], mr[8], mb[8];
        const int c = ly * L::AW + lx;                      // input-plane index of window cell (0,0) = pixel (y-2, x-2)
#pragma unroll
        for (int u = 0; u < 6; ++u)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const F4 t = ((const F4*)(DR + c + u * L::AW))[v];
                w[u][4 * v] = t.x; w[u][4 * v + 1] = t.y; w[u][4 * v + 2] = t.z; w[u][4 * v + 3] = t.w;
            }
        median25_block2x4(w, mr);
#pragma unroll
        for (int u = 0; u < 6; ++u)
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const F4 t = ((const F4*)(DB + c + u * L::AW))[v];
                w[u][4 * v] = t.x; w[u][4 * v + 1] = t.y; w[u][4 * v + 2] = t.z; w[u][4 * v + 3] = t.w;
            }
        median25_block2x4(w, mb);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const float* gp = G + c + (2 + dy) * L::AW + 2;           // 8-byte aligned (c is a multiple of 4)
            const F2 ga = ((const F2*)gp)[0], gb = ((const F2*)gp)[1];
            const float g4[4] = {ga.x, ga.y, gb.x, gb.y};
            F4 r1, b1, er, eb;
            float* r1p = &r1.x; float* b1p = &b1.x; float* erp = &er.x; float* ebp = &eb.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r1p[k] = mr[4 * dy + k] + g4[k]; b1p[k] = mb[4 * dy + k] + g4[k];
                erp[k] = g4[k] - r1p[k]; ebp[k] = g4[k] - b1p[k];
            }
            const int o = (ly + dy) * L::BW + lx;
            *(F4*)(ER + o) = er; *(F4*)(EB + o) = eb;
            const int ty = ly + dy - 2, tx = lx - 2;         // tile coords of the block's first pixel: tx = 2 mod 4
            if (ty >= 0 && ty < TH) {
                if (tx >= 0 && tx + 1 < TW) { F2 t; t.x = r1.x; t.y = r1.y; *(F2*)(RP + ty * TW + tx) = t; t.x = b1.x; t.y = b1.y; *(F2*)(BP + ty * TW + tx) = t; }
                if (tx + 2 >= 0 && tx + 3 < TW) { F2 t; t.x = r1.z; t.y = r1.w; *(F2*)(RP + ty * TW + tx + 2) = t; t.x = b1.z; t.y = b1.w; *(F2*)(BP + ty * TW + tx + 2) = t; }
            }
        }
    }
}

template <int TW, int TH, bool EDGE>
PYSP_D void median_phase_c4(const MedianParams& p, char* __restrict__ smem, int tile_x, int tile_y) {
    typedef MedianTile<TW, TH> L;
    static_assert(TW % 4 == 0, "2x4 blocks");
    const int H = p.g.H, W = p.g.W;
    const int x0 = tile_x * TW, y0 = p.y_begin + tile_y * TH;
    const float* ER = (const float*)(smem + L::OFF_ER); const float* EB = (const float*)(smem + L::OFF_EB);
    const float* RP = (const float*)(smem + L::OFF_RP); const float* BP = (const float*)(smem + L::OFF_BP);
    float* out = (float*)(smem + L::OFF_OUT);
    constexpr int NBX = TW / 4, NBY = TH / 2;
    PYSP_ITEMS(it, NBX * NBY) {
        int by = it / NBX, bx = it - by * NBX;
        int ty = 2 * by, tx = 4 * bx;
        int y = y0 + ty, x = x0 + tx;
        if (EDGE) { if (y >= p.y_end || x >= W) continue; }
        int ry[6], rx[8];                                   // region-B rows/cols of the 6x8 window (REPLICATE at the frame border)
#pragma unroll
        for (int k = 0; k < 6; ++k) ry[k] = EDGE ? clampi(y + k - 2, H) - (y0 - 2) : ty + k;
#pragma unroll
        for (int k = 0; k < 8; ++k) rx[k] = EDGE ? clampi(x + k - 2, W) - (x0 - 2) : tx + k;
        float w[6][8], mr[8], mb[8];
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            if (EDGE) {
#pragma unroll
                for (int v = 0; v < 8; ++v) w[u][v] = ER[ry[u] * L::BW + rx[v]];
            } else {
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const F4 t = ((const F4*)(ER + ry[u] * L::BW + tx))[v];
                    w[u][4 * v] = t.x; w[u][4 * v + 1] = t.y; w[u][4 * v + 2] = t.z; w[u][4 * v + 3] = t.w;
                }
            }
        }
        median25_block2x4(w, mr);
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            if (EDGE) {
#pragma unroll
                for (int v = 0; v < 8; ++v) w[u][v] = EB[ry[u] * L::BW + rx[v]];
            } else {
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const F4 t = ((const F4*)(EB + ry[u] * L::BW + tx))[v];
                    w[u][4 * v] = t.x; w[u][4 * v + 1] = t.y; w[u][4 * v + 2] = t.z; w[u][4 * v + 3] = t.w;
                }
            }
        }
        median25_block2x4(w, mb);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int dy = k >> 2, dx = k & 3;
            if (EDGE) { if (x + dx >= W) continue; }          // W is even, not necessarily a multiple of 4
            int o = (ty + dy) * TW + tx + dx;
            float r1 = RP[o], b1 = BP[o];
            Rgb v;
            v.r = r1; v.b = b1;
            v.g = (((mr[k] + mb[k]) + r1) + b1) / 2.0f;
            stage_pixel<TW, TH>(out, p.st, p.g, p.c, ty + dy, tx + dx, v);
        }
    }
}

// phase C: g' on the tile, epilogue -> output staging tile
template <int TW, int TH, bool EDGE>
PYSP_D void median_phase_c(const MedianParams& p, char* __restrict__ smem, int tile_x, int tile_y) {
    typedef MedianTile<TW, TH> L;
    const int H = p.g.H, W = p.g.W;
    const int x0 = tile_x * TW, y0 = p.y_begin + tile_y * TH;
    const float* ER = (const float*)(smem + L::OFF_ER); const float* EB = (const float*)(smem + L::OFF_EB);
    const float* RP = (const float*)(smem + L::OFF_RP); const float* BP = (const float*)(smem + L::OFF_BP);
    float* out = (float*)(smem + L::OFF_OUT);
    constexpr int NBX = TW / 2, NBY = TH / 2;
    PYSP_ITEMS(it, NBX * NBY) {
        int by = it / NBX, bx = it - by * NBX;
        int ty = 2 * by, tx = 2 * bx;
        int y = y0 + ty, x = x0 + tx;
        if (EDGE) { if (y >= p.y_end || x >= W) continue; }
        int ry[6], rx[6];                                   // region-B rows/cols of the 6x6 window (REPLICATE at the frame border)
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            ry[k] = EDGE ? clampi(y + k - 2, H) - (y0 - 2) : ty + k;
            rx[k] = EDGE ? clampi(x + k - 2, W) - (x0 - 2) : tx + k;
        }
        float w[6][6], mr[4], mb[4];
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            if (EDGE) {
#pragma unroll
                for (int v = 0; v < 6; ++v) w[u][v] = ER[ry[u] * L::BW + rx[v]];
            } else {                                        // tx and BW are even: three aligned pairs per window row
#pragma unroll
                for (int v = 0; v < 3; ++v) { const F2 t = ((const F2*)(ER + ry[u] * L::BW + tx))[v]; w[u][2 * v] = t.x; w[u][2 * v + 1] = t.y; }
            }
        }
        median25_block2x2(w, mr);
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            if (EDGE) {
#pragma unroll
                for (int v = 0; v < 6; ++v) w[u][v] = EB[ry[u] * L::BW + rx[v]];
            } else {
#pragma unroll
                for (int v = 0; v < 3; ++v) { const F2 t = ((const F2*)(EB + ry[u] * L::BW + tx))[v]; w[u][2 * v] = t.x; w[u][2 * v + 1] = t.y; }
            }
        }
        median25_block2x2(w, mb);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dy = k >> 1, dx = k & 1;
            int o = (ty + dy) * TW + tx + dx;
            float r1 = RP[o], b1 = BP[o];
            Rgb v;
            v.r = r1; v.b = b1;
            v.g = (((mr[k] + mb[k]) + r1) + b1) / 2.0f;
            stage_pixel<TW, TH>(out, p.st, p.g, p.c, ty + dy, tx + dx, v);
        }
    }
}

}  // namespace pysp
