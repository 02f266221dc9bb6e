// Steps either side of the develop path: flat-field correction and hot-pixel detection on the
// mosaic before it, camera-space HDR fusion of demosaiced exposures after it.  All HBM-bound point/stencil work.
//
// flat_frame_correction (raw_correction.py:25-63) needs np.mean of each CFA plane of the flat field in float32, and
// NumPy's float32 sum is a *pairwise* sum with a fixed shape (loops_utils.h.src, @TYPE@_pairwise_sum): blocks of at
// most 128 elements are summed with eight strided accumulators, blocks are combined by recursive halving (left half
// rounded down to a multiple of 8).  The tree only depends on the element count, so the host lays it out once
// (PlaneSumPlan) and the device evaluates it in exactly that order: leaf_sum_kernel, then tree_sum_kernel level by
// level.  The result is bit-identical to np.mean.
#pragma once
#include "pysp_common.cuh"

namespace pysp {

struct PlaneSumTables {         // device pointers into the caller's workspace
    const int* leaf_off;        // [n_leaves] first plane element of the leaf
    const int* leaf_len;        // [n_leaves] 1..128
    const int* node_l;          // [n_nodes] value index of the left child (leaves are 0..n_leaves-1, nodes follow)
    const int* node_r;
    const int* group_start;     // [n_groups + 1] nodes of one group are independent (same height)
    int n_leaves, n_nodes, n_groups;
    float* val;                 // [4][n_leaves + n_nodes]
};

struct FlatParams {             // raw_correction.py:25-63
    const float* sensor; long long sensor_pitch;
    const float* flat; long long flat_pitch;
    float* out; long long out_pitch;
    int H, W;
    int clamp_high;
    float n_f32;                // float32(h * w), the divisor of np.mean
    PlaneSumTables t;
    float* mean;                // [4] device
    int* stat;                  // [4][2]: ordered-int maximum of the finite quotients, number of quotients that are not +-inf
};

struct HotParams {              // raw_bad_pixel_corr.py:30-65
    const float* sensor; long long pitch;
    int H, W;
    float min_delta;
    int min_count;
    uint8_t* masks;             // [4][h][w], planes in the reference's order R, G1, B, G2
};

#define PYSP_MAX_EXPOSURES 16
struct FuseCamParams {          // raw_hdr.py:7-83
    float* img[PYSP_MAX_EXPOSURES];     // [n_px][3] camera RGB, white balance applied; rewritten when write_back
    int n;
    long long n_px;
    float wb[3], max_wb;
    int normalized[PYSP_MAX_EXPOSURES];
    float ev_off[PYSP_MAX_EXPOSURES];   // float32(2**(ev_i - target))
    float bias[PYSP_MAX_EXPOSURES];     // float32(1.6**(-0.1*ev_off_i))
    int brightest;                      // last exposure whose offset equals the maximum
    double off_max;                     // the maximum offset as float64 (np.max returns a strong float64 scalar)
    double m[9];
    float* out; int32_t* count;
    int write_back;
};

// order-preserving float <-> int map (for atomicMax on floats of either sign)
PYSP_HD int float_key(float f) { int b = (int)pysp_as_uint(f); return b >= 0 ? b : b ^ 0x7FFFFFFF; }
PYSP_HD float key_float(int k) { return pysp_as_float((uint32_t)(k >= 0 ? k : k ^ 0x7FFFFFFF)); }
#define PYSP_KEY_NONE ((int)0x80000000)     // below every finite float's key: "no finite value seen"

#ifndef PYSP_HOST_EMU
// plane p in the reference's order R(TL), G1(TR), B(BR), G2(BL): row / column parity of its sites
__device__ __forceinline__ int plane_py(int p) { return p >> 1; }
__device__ __forceinline__ int plane_px(int p) { return (p == 1 || p == 2) ? 1 : 0; }

__device__ __forceinline__ float plane_elem(const float* base, long long pitch, int w, int py, int px, int e) {
    int i = e / w, j = e - i * w;
    return *((const float*)((const char*)base + (long long)(2 * i + py) * pitch) + 2 * j + px);
}

// one thread = one leaf of one plane, evaluated exactly as NumPy's unrolled block sum
__global__ void __launch_bounds__(128) leaf_sum_kernel(const float* mosaic, long long pitch, int W, PlaneSumTables t) {
    const int w = W >> 1;
    const int total = 4 * t.n_leaves;
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < total; id += gridDim.x * blockDim.x) {
        const int p = id / t.n_leaves, leaf = id - p * t.n_leaves;
        const int py = plane_py(p), px = plane_px(p);
        const int off = t.leaf_off[leaf], n = t.leaf_len[leaf];
        float res;
        if (n < 8) {
            res = 0.0f;
            for (int i = 0; i < n; ++i) res = res + plane_elem(mosaic, pitch, w, py, px, off + i);
        } else {
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = plane_elem(mosaic, pitch, w, py, px, off + j);
            int i = 8;
            for (; i < n - (n % 8); i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = r[j] + plane_elem(mosaic, pitch, w, py, px, off + i + j);
            }
            res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            for (; i < n; ++i) res = res + plane_elem(mosaic, pitch, w, py, px, off + i);
        }
        t.val[(long long)p * (t.n_leaves + t.n_nodes) + leaf] = res;
    }
}

// one block = one plane: combine the leaves in the order of NumPy's recursion, then mean = sum / float32(n)
__global__ void __launch_bounds__(1024) tree_sum_kernel(PlaneSumTables t, float n_f32, float* mean) {
    float* val = t.val + (long long)blockIdx.x * (t.n_leaves + t.n_nodes);
    for (int g = 0; g < t.n_groups; ++g) {
        const int a = t.group_start[g], b = t.group_start[g + 1];
        for (int k = a + threadIdx.x; k < b; k += blockDim.x) val[t.n_leaves + k] = val[t.node_l[k]] + val[t.node_r[k]];
        __syncthreads();
    }
    if (threadIdx.x == 0) mean[blockIdx.x] = val[t.n_leaves + t.n_nodes - 1] / n_f32;
}

__device__ __forceinline__ float flat_quotient(const FlatParams& p, const float* mean, int y, int x, float* chan) {
    const int pos = ((y & 1) << 1) | (x & 1);                       // TL, TR, BL, BR
    const int plane = pos == 0 ? 0 : (pos == 1 ? 1 : (pos == 3 ? 2 : 3));
    const float c = *((const float*)((const char*)p.sensor + (long long)y * p.sensor_pitch) + x);
    const float f = *((const float*)((const char*)p.flat + (long long)y * p.flat_pitch) + x);
    *chan = c;
    return (c * mean[plane]) / f;                                   // raw_correction.py:46
}

__global__ void __launch_bounds__(256) flat_stats_kernel(FlatParams p) {
    __shared__ int s_max[4], s_cnt[4];
    if (threadIdx.x < 4) { s_max[threadIdx.x] = PYSP_KEY_NONE; s_cnt[threadIdx.x] = 0; }
    __syncthreads();
    float mean[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) mean[k] = p.mean[k];
    int lmax[4] = {PYSP_KEY_NONE, PYSP_KEY_NONE, PYSP_KEY_NONE, PYSP_KEY_NONE}, lcnt[4] = {0, 0, 0, 0};
    const long long n = (long long)p.H * p.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
        float c;
        const float q = flat_quotient(p, mean, y, x, &c);
        const int pos = ((y & 1) << 1) | (x & 1);
        const int plane = pos == 0 ? 0 : (pos == 1 ? 1 : (pos == 3 ? 2 : 3));
        const bool inf = isinf(q), fin = !inf && !isnan(q);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (plane == k) {
                if (fin) lmax[k] = max(lmax[k], float_key(q));
                lcnt[k] += inf ? 0 : 1;
            }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (lmax[k] != PYSP_KEY_NONE) atomicMax(&s_max[k], lmax[k]);
        if (lcnt[k]) atomicAdd(&s_cnt[k], lcnt[k]);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        if (s_max[threadIdx.x] != PYSP_KEY_NONE) atomicMax(&p.stat[2 * threadIdx.x], s_max[threadIdx.x]);
        if (s_cnt[threadIdx.x]) atomicAdd(&p.stat[2 * threadIdx.x + 1], s_cnt[threadIdx.x]);
    }
}

__global__ void __launch_bounds__(256) flat_apply_kernel(FlatParams p) {
    float mean[4], vmax[4];
    bool all_inf[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mean[k] = p.mean[k];
        const int key = p.stat[2 * k];
        vmax[k] = key == PYSP_KEY_NONE ? pysp_as_float(0x7FC00000u) : key_float(key);
        all_inf[k] = p.stat[2 * k + 1] == 0;                        // np.isinf(output).all(): leave the plane alone
    }
    const long long n = (long long)p.H * p.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
        float c;
        float q = flat_quotient(p, mean, y, x, &c);
        const int pos = ((y & 1) << 1) | (x & 1);
        const int plane = pos == 0 ? 0 : (pos == 1 ? 1 : (pos == 3 ? 2 : 3));
        float m = vmax[0]; bool ai = all_inf[0];
#pragma unroll
        for (int k = 1; k < 4; ++k) if (plane == k) { m = vmax[k]; ai = all_inf[k]; }
        if (ai) {
            q = c;
        } else {
            if (q == pysp_as_float(0x7F800000u)) q = m;             // +inf -> largest finite value of the plane
            if (q < 0.0f) q = 0.0f;
            if (p.clamp_high && q > 1.0f) q = 1.0f;
        }
        *((float*)((char*)p.out + (long long)y * p.out_pitch) + x) = q;
    }
}

// one thread = one photosite; its eight same-colour neighbours are two mosaic pixels away (np.pad mode="reflect" on the plane)
__global__ void __launch_bounds__(256) hot_pixel_kernel(HotParams p) {
    const int h = p.H >> 1, w = p.W >> 1;
    const long long n = (long long)p.H * p.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
        const int py = y & 1, px = x & 1, qy = y >> 1, qx = x >> 1;
        const int plane = py == 0 ? px : (px ? 2 : 3);
        const float ref = *((const float*)((const char*)p.sensor + (long long)y * p.pitch) + x) - p.min_delta;
        int cnt = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (dy == 0 && dx == 0) continue;
                const int ny = reflect101(qy + dy, h), nx = reflect101(qx + dx, w);
                const float v = *((const float*)((const char*)p.sensor + (long long)(2 * ny + py) * p.pitch) + 2 * nx + px);
                cnt += ref > v ? 1 : 0;
            }
        p.masks[(long long)plane * h * w + (long long)qy * w + qx] = cnt > p.min_count ? 1 : 0;
    }
}

// one thread = one pixel (three channels), exposures strictly in list order
__global__ void __launch_bounds__(256) fuse_cam_kernel(FuseCamParams p) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n_px; i += (long long)gridDim.x * blockDim.x) {
        float fused[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float sum_w = 0.0f, sum_p = 0.0f, bright = 0.0f;
            int cnt = 0;
            for (int k = 0; k < p.n; ++k) {
                float v = p.img[k][3 * i + c];
                if (p.normalized[k]) v = v * p.max_wb;                                   // image_base.py:56-57
                const float u = __double2float_rn(__ddiv_rn((double)v, (double)p.wb[c])); // wb_undo (float64 division)
                const float wgt = (0.5f - fabsf(u - 0.5f)) * p.bias[k];                  // raw_hdr.py:59-62
                sum_w = sum_w + wgt;
                const float a = u * p.wb[c];                                             // wb_apply
                if (p.write_back) p.img[k][3 * i + c] = a;
                if (k == p.brightest) bright = a;
                sum_p = sum_p + ((a * wgt) * p.ev_off[k]);                               // raw_hdr.py:71
                cnt += wgt > 0.0f ? 1 : 0;
            }
            const float q = sum_p / sum_w;
            fused[c] = sum_w == 0.0f ? __double2float_rn(__dmul_rn((double)bright, p.off_max)) : q;   // raw_hdr.py:75-80
            if (p.count) p.count[3 * i + c] = cnt;
        }
        float* o = p.out + 3 * i;                                                         // clip_highlights=False (raw_hdr.py:81)
        o[0] = dot3_f64(p.m + 0, fused[0], fused[1], fused[2]);
        o[1] = dot3_f64(p.m + 3, fused[0], fused[1], fused[2]);
        o[2] = dot3_f64(p.m + 6, fused[0], fused[1], fused[2]);
    }
}
#endif

}  // namespace pysp
