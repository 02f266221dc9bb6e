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
    const float* wx = tab + (sx & 31) * 8;
    const float* wy = tab + (sy & 31) * 8;
    float hx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) hx[i] = wx[i];
    float sum = 0.0f;
    if (ix >= 0 && ix + 8 <= W && iy >= 0 && iy + 8 <= H) {
        const float* p = src + (long long)iy * pitch_f + (long long)ix * step;
#pragma unroll
        for (int r = 0; r < 8; ++r, p += pitch_f) {
            const float vy = wy[r];
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
            const float vy = wy[r];
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
    tab[threadIdx.x] = p.tab[threadIdx.x];
    __syncthreads();
    const long long n = (long long)p.H * p.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
        const float2 m = *(const float2*)((const char*)p.map + (long long)y * p.map_pitch + (long long)x * 8);
        const float v = lanczos4_sample(p.src, p.src_pitch / 4, p.src_step, p.H, p.W, tab, clip_coord(m.x, p.W), clip_coord(m.y, p.H));
        *(float*)((char*)p.dst + (long long)y * p.dst_pitch + (long long)x * p.dst_step * 4) = v;
    }
}

// opcode_warp_rectilinear for every plane of an interleaved image in ONE kernel: the coordinates are computed in registers
// and never written to HBM (24 B/px of algorithmic traffic for three planes instead of 72 with tables).  A block owns a
// 32 x 8 pixel patch so that the 8x8 gather windows of neighbouring threads overlap in L1.
__global__ void __launch_bounds__(256) warp_apply_kernel(const WarpApplyParams p) {
    __shared__ float tab[256];
    tab[threadIdx.x] = p.tab[threadIdx.x];
    __syncthreads();
    const int tiles_x = (p.g.W + 31) / 32, tiles_y = (p.g.H + 7) / 8;
    const long long pitch_f = (long long)p.g.W * p.planes;
    for (int t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        const int x = tx * 32 + (threadIdx.x & 31), y = ty * 8 + (threadIdx.x >> 5);
        if (x >= p.g.W || y >= p.g.H) continue;
        for (int c = 0; c < p.planes; ++c) {
            float sx = (float)x, sy = (float)y;
            if (p.prior) {
                const float2 s = *(const float2*)(p.prior + (((long long)y * p.g.W + x) * p.planes + c) * 2);
                sx = s.x; sy = s.y;
            }
            float mx, my;
            warp_coord(p.g, p.k[c], sx, sy, &mx, &my);
            p.dst[((long long)y * p.g.W + x) * p.planes + c] =
                lanczos4_sample(p.src + c, pitch_f, p.planes, p.g.H, p.g.W, tab, clip_coord(mx, p.g.W), clip_coord(my, p.g.H));
        }
    }
}
#endif

}  // namespace pysp
