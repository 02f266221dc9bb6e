// DNG WarpRectilinear (post-demosaic lens correction):
//   coordinate table  dng_warp_corr/dng_warp_rectilinear_coords.pyx:18-40 (compute_table), :44-65 (offset_table)
//   resampling        cv2.remap(plane, clip(map_x), clip(map_y), cv2.INTER_LANCZOS4), dng_warp_corr/chan_distortion_corr.py:94-97
//
// The table follows the arithmetic types of the C that Cython generates from the .pyx: float32 throughout, except the two
// tangential terms whose Python literal `2` is a C double (evaluated in double, rounded once), and `r = sqrt(..)` which is
// libc's double sqrt rounded back (identical to the correctly rounded float sqrt).  `x ** n` is libm's powf there:
// dx**2 / r**2 are taken as the float product (correctly rounded; powf(x, 2) agrees), r**4 and r**6 through float64
// (correctly rounded but for ~2^-28 of the inputs, where glibc's powf -- itself not correctly rounded, and CPU-dependent --
// may differ by one unit in the last place).  The parity contract for the table is therefore a tolerance (2 ulp).
//
// cv2.remap quantises the sampling position to 1/32 px (cvRound(v * 32), half to even); the 8x8 window starts 3 px up-left
// of its integer part; the tap weights are tab[fy][k1] * tab[fx][k2] from OpenCV's float32 Lanczos-4 table (data harvested
// from cv2, pysp_b200/data/lanczos4_tab_f32.npy); taps outside the image contribute 0 (BORDER_CONSTANT).  Interior windows
// are summed row by row, border windows tap by tap, in float32 -- OpenCV's own order, so that the result is bit-identical
// when the table is.
#pragma once
#include "pysp_common.cuh"

namespace pysp {

struct WarpGeom {
    int H, W;
    float cx, cy, m, scale;     // optical centre (pixels), normalisation radius: host scalars of pyx:73-77
};

struct WarpTableParams {
    WarpGeom g;
    float k[6];                 // kr0..kr3, kt0, kt1
    const float* seed;          // optional [H][W][2] prior mapping (offset_table), else the pixel grid
    long long seed_pitch;
    float* table;               // [H][W][2]
    long long table_pitch;
};

#define PYSP_WARP_MAX_PLANES 4
struct WarpApplyParams {
    WarpGeom g;
    int planes;
    float k[PYSP_WARP_MAX_PLANES][6];
    const float* prior;         // optional [H][W][planes][2]
    const float* src;           // [H][W][planes] interleaved
    float* dst;
    const float* tab;           // [32][8] Lanczos-4 weights (device)
};

struct RemapParams {
    int H, W;
    const float* src; long long src_pitch; int src_step;     // plane of an interleaved image: element (y, x) at src[y*pitch/4 + x*step]
    float* dst; long long dst_pitch; int dst_step;
    const float* map; long long map_pitch;                   // [H][W][2]
    const float* tab;
};

// one coordinate pair (pyx:25-40 / 51-65)
PYSP_HD void warp_coord(const WarpGeom& g, const float* k, float sx, float sy, float* ox, float* oy) {
#ifdef __CUDA_ARCH__
    const float dx = __fdiv_rn(sx - g.cx, g.m), dy = __fdiv_rn(sy - g.cy, g.m);
    const float dx2 = dx * dx, dy2 = dy * dy;
    const float r = __fsqrt_rn(dx2 + dy2);
    const double rd = (double)r, r2d = __dmul_rn(rd, rd), r4d = __dmul_rn(r2d, r2d);
    const float r2 = __double2float_rn(r2d), r4 = __double2float_rn(r4d), r6 = __double2float_rn(__dmul_rn(r4d, r2d));
    const float f = ((k[0] + (k[1] * r2)) + (k[2] * r4)) + (k[3] * r6);
    const float dxr = f * dx, dyr = f * dy;
    const double xy2 = __dmul_rn(__dmul_rn(2.0, (double)dx), (double)dy);
    const float dxt = __double2float_rn(__dadd_rn(__dmul_rn((double)k[4], xy2),
                                                  __dmul_rn((double)k[5], __dadd_rn((double)r2, __dmul_rn(2.0, (double)dx2)))));
    const float dyt = __double2float_rn(__dadd_rn(__dmul_rn((double)k[5], xy2),
                                                  __dmul_rn((double)k[4], __dadd_rn((double)r2, __dmul_rn(2.0, (double)dy2)))));
    const float xp = g.cx + (g.m * (dxr + dxt));
    const float yp = g.cy + (g.m * (dyr + dyt));
    *ox = sx + ((xp - sx) * g.scale);
    *oy = sy + ((yp - sy) * g.scale);
#else
    (void)g; (void)k; *ox = sx; *oy = sy;
#endif
}

#ifdef __CUDACC__
// cv2.remap, INTER_LANCZOS4, float32, BORDER_CONSTANT(0): one sample of one plane at the (already clipped) position (mx, my)
__device__ __forceinline__ float lanczos4_sample(const float* __restrict__ src, long long pitch_f, int step, int H, int W,
                                                 const float* __restrict__ tab, float mx, float my) {
    const int sx = __float2int_rn(mx * 32.0f), sy = __float2int_rn(my * 32.0f);     // cvRound: half to even
    const int ix = (sx >> 5) - 3, iy = (sy >> 5) - 3;
    // the shared-memory copy of the table is transposed ([tap][fraction]): lanes with different fractions hit different banks
    const float* wx = tab + (sx & 31);
    const float* wy = tab + (sy & 31);
    float hx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hx[i] = wx[i * 32];
    float sum = 0.0f;
    if (ix >= 0 && ix + 8 <= W && iy >= 0 && iy + 8 <= H) {
        const float* p = src + (long long)iy * pitch_f + (long long)ix * step;
#pragma unroll
        for (int r = 0; r < 8; ++r, p += pitch_f) {
            const float vy = wy[r * 32];
            float row = __ldg(p) * (vy * hx[0]);
#pragma unroll
            for (int c = 1; c < 8; ++c) row = row + __ldg(p + c * step) * (vy * hx[c]);
            sum = sum + row;
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
            const int yy = iy + r;
            if (yy < 0 || yy >= H) continue;
            const float vy = wy[r * 32];
            const float* p = src + (long long)yy * pitch_f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int xx = ix + c;
                if (xx >= 0 && xx < W) sum = sum + __ldg(p + (long long)xx * step) * (vy * hx[c]);
            }
        }
    }
    return sum;
}

__device__ __forceinline__ float clip_coord(float v, int n) { return fminf(fmaxf(v, 0.0f), (float)(n - 1)); }   // np.clip

// compute_remapping_table / compute_offset_remapping_table: 8 (16 with a seed) B/px of HBM traffic
__global__ void __launch_bounds__(256) warp_table_kernel(const WarpTableParams p) {
    const long long n = (long long)p.g.H * p.g.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.g.W), x = (int)(i - (long long)y * p.g.W);
        float sx = (float)x, sy = (float)y;
        if (p.seed) {
            const float2 s = *(const float2*)((const char*)p.seed + (long long)y * p.seed_pitch + (long long)x * 8);
            sx = s.x; sy = s.y;
        }
        float2 o;
        warp_coord(p.g, p.k, sx, sy, &o.x, &o.y);
        *(float2*)((char*)p.table + (long long)y * p.table_pitch + (long long)x * 8) = o;
    }
}

// cv2.remap of one plane through a table held in HBM (the reference's two-step form)
__global__ void __launch_bounds__(256) remap_lanczos4_kernel(const RemapParams p) {
    __shared__ float tab[256];
    tab[(threadIdx.x & 7) * 32 + (threadIdx.x >> 3)] = p.tab[threadIdx.x];      // [32][8] -> [8][32]
    __syncthreads();
    const long long n = (long long)p.H * p.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
        const float2 m = *(const float2*)((const char*)p.map + (long long)y * p.map_pitch + (long long)x * 8);
        const float v = lanczos4_sample(p.src, p.src_pitch / 4, p.src_step, p.H, p.W, tab, clip_coord(m.x, p.W), clip_coord(m.y, p.H));
        *(float*)((char*)p.dst + (long long)y * p.dst_pitch + (long long)x * p.dst_step * 4) = v;
    }
}

// the same sample taken from a patch of the source plane staged in shared memory: `patch` holds rows [py0, py0 + PH) and
// columns [px0, px0 + PW) of the plane, zero outside the image (= BORDER_CONSTANT), and the caller guarantees that the 8x8
// window lies inside the patch.  Interior windows are summed row by row, border windows tap by tap (OpenCV's two orders); a
// zero-filled tap adds +0.0 * w, which leaves a float32 sum unchanged, so the border order over the padded patch gives the
// bits of the tap-skipping loop.
template <int PW>
__device__ __forceinline__ float lanczos4_sample_smem(const float* __restrict__ patch, int px0, int py0, int H, int W,
                                                      const float* __restrict__ tab, int sx, int sy) {
    const int ix = (sx >> 5) - 3, iy = (sy >> 5) - 3;
    // the shared-memory copy of the table is transposed ([tap][fraction]): lanes with different fractions hit different banks
    const float* wx = tab + (sx & 31);
    const float* wy = tab + (sy & 31);
    float hx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hx[i] = wx[i * 32];
    const float* p = patch + (iy - py0) * PW + (ix - px0);
    float sum = 0.0f;
    if (ix >= 0 && ix + 8 <= W && iy >= 0 && iy + 8 <= H) {
#pragma unroll
        for (int r = 0; r < 8; ++r, p += PW) {
            const float vy = wy[r * 32];
            float row = p[0] * (vy * hx[0]);
#pragma unroll
            for (int c = 1; c < 8; ++c) row = row + p[c] * (vy * hx[c]);
            sum = sum + row;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r, p += PW) {
            const float vy = wy[r * 32];
#pragma unroll
            for (int c = 0; c < 8; ++c) sum = sum + p[c] * (vy * hx[c]);
        }
    }
    return sum;
}

// opcode_warp_rectilinear for every plane of an interleaved image in ONE kernel: the coordinates are computed in registers
// and never written to HBM (24 B/px of algorithmic traffic for three planes instead of 72 with tables).  A block owns a
// 32 x 8 pixel patch of the output; lens distortion is smooth, so the source windows of its 256 pixels cover a patch only a
// little larger.  Per plane the block finds that bounding box (warp + block min/max of the window origins), stages it in
// shared memory with coalesced row loads, and every thread takes its 64 taps from there (one shared-memory wavefront per
// tap instead of three to four L1 wavefronts for a strided global gather).  A block whose box does not fit the staging
// buffer (extreme coefficients) gathers from global memory instead: same arithmetic, same bits.
#define PYSP_WARP_PW 56
#define PYSP_WARP_PH 28
#ifndef PYSP_WARP_SMEM
#define PYSP_WARP_SMEM 1
#endif
__global__ void __launch_bounds__(256) warp_apply_kernel(const WarpApplyParams p) {
    __shared__ float tab[256];
    __shared__ float patch[PYSP_WARP_MAX_PLANES * PYSP_WARP_PH * PYSP_WARP_PW];
    __shared__ int box[2][4];
    tab[(threadIdx.x & 7) * 32 + (threadIdx.x >> 3)] = p.tab[threadIdx.x];      // [32][8] -> [8][32]
    if (threadIdx.x < 8) box[threadIdx.x >> 2][threadIdx.x & 3] = (threadIdx.x & 2) ? -0x7fffffff : 0x7fffffff;
    __syncthreads();
    const int tiles_x = (p.g.W + 31) / 32, tiles_y = (p.g.H + 7) / 8;
    const long long pitch_f = (long long)p.g.W * p.planes;
    int flip = 0;
    for (int t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x, flip ^= 1) {
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        const int x = tx * 32 + (threadIdx.x & 31), y = ty * 8 + (threadIdx.x >> 5);
        const bool live = x < p.g.W && y < p.g.H;
        int sxq[PYSP_WARP_MAX_PLANES], syq[PYSP_WARP_MAX_PLANES];
        int lox = 0x7fffffff, loy = 0x7fffffff, hix = -0x7fffffff, hiy = -0x7fffffff;
#pragma unroll
        for (int c = 0; c < PYSP_WARP_MAX_PLANES; ++c) {
            sxq[c] = syq[c] = 0;
            if (c < p.planes && live) {
                float sx = (float)x, sy = (float)y;
                if (p.prior) {
                    const float2 s = *(const float2*)(p.prior + (((long long)y * p.g.W + x) * p.planes + c) * 2);
                    sx = s.x; sy = s.y;
                }
                float mx, my;
                warp_coord(p.g, p.k[c], sx, sy, &mx, &my);
                sxq[c] = __float2int_rn(clip_coord(mx, p.g.W) * 32.0f);      // cvRound: half to even
                syq[c] = __float2int_rn(clip_coord(my, p.g.H) * 32.0f);
                lox = min(lox, (sxq[c] >> 5) - 3); hix = max(hix, (sxq[c] >> 5) - 3);
                loy = min(loy, (syq[c] >> 5) - 3); hiy = max(hiy, (syq[c] >> 5) - 3);
            }
        }
        bool fits = false;
        int px0 = 0, py0 = 0;
        if (PYSP_WARP_SMEM) {
            // bounding box of the window origins of every live pixel and plane of the block (two alternating boxes: the one
            // of the next tile is re-armed while this one is in use)
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                lox = min(lox, __shfl_xor_sync(0xffffffffu, lox, o)); loy = min(loy, __shfl_xor_sync(0xffffffffu, loy, o));
                hix = max(hix, __shfl_xor_sync(0xffffffffu, hix, o)); hiy = max(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
            }
            if ((threadIdx.x & 31) == 0) {
                atomicMin(&box[flip][0], lox); atomicMin(&box[flip][1], loy); atomicMax(&box[flip][2], hix); atomicMax(&box[flip][3], hiy);
            }
            __syncthreads();                                  // (also: every thread is done with the previous tile's patch)
            px0 = box[flip][0]; py0 = box[flip][1];
            const int pw = box[flip][2] + 8 - px0, ph = box[flip][3] + 8 - py0;
            fits = pw <= PYSP_WARP_PW && ph <= PYSP_WARP_PH;
            if (threadIdx.x < 4) box[flip ^ 1][threadIdx.x] = (threadIdx.x & 2) ? -0x7fffffff : 0x7fffffff;
            if (fits) {
                for (int i = threadIdx.x; i < ph * PYSP_WARP_PW; i += 256) {
                    const int r = i / PYSP_WARP_PW, cc = i - r * PYSP_WARP_PW;
                    const int yy = py0 + r, xx = px0 + cc;
                    const bool in = cc < pw && yy >= 0 && yy < p.g.H && xx >= 0 && xx < p.g.W;
                    const float* q = p.src + (long long)yy * pitch_f + (long long)xx * p.planes;
                    for (int c = 0; c < p.planes; ++c) patch[c * (PYSP_WARP_PH * PYSP_WARP_PW) + i] = in ? __ldg(q + c) : 0.0f;
                }
            }
            __syncthreads();                                  // patch filled; box[flip] read by everyone
        }
        if (live) {
#pragma unroll                                   // static plane index: sxq / syq stay in registers
            for (int c = 0; c < PYSP_WARP_MAX_PLANES; ++c) {
                if (c >= p.planes) break;
                float v;
                if (fits) v = lanczos4_sample_smem<PYSP_WARP_PW>(patch + c * (PYSP_WARP_PH * PYSP_WARP_PW), px0, py0, p.g.H, p.g.W, tab, sxq[c], syq[c]);
                else v = lanczos4_sample(p.src + c, pitch_f, p.planes, p.g.H, p.g.W, tab, (float)sxq[c] * 0.03125f, (float)syq[c] * 0.03125f);
                p.dst[((long long)y * p.g.W + x) * p.planes + c] = v;
            }
        }
    }
}
#endif

}  // namespace pysp
