// Point-wise kernels of the develop path: normalisation, camera->linear-sRGB matrix, sRGB gamma, raw-space
// HDR fusion.  All HBM-bound; each thread handles a few contiguous elements with vector accesses.
#pragma once
#include "pysp_common.cuh"

namespace pysp {

struct NormalizeParams {      // normalization.py:4-25 (stored orientation, no flips)
    const uint16_t* in; long long in_pitch;
    float* out; long long out_pitch;
    int H, W;
    float black[4], white[4]; // by position TL,TR,BL,BR
};

struct MatrixParams {         // colorize/transform.py:21-53 on [n] RGB pixels
    const float* in; void* out; long long n;
    double m[9];
    int clip, gamma, out_f16;
};

struct GammaParams { const float* in; float* out; long long n; };   // colorize/transform.py:89-99

#define PYSP_MAX_BRACKETS 16
struct FuseParams {           // raw_hdr.py:108-148
    const float* in[PYSP_MAX_BRACKETS]; long long in_pitch;
    int n, H, W;
    float ev_off[PYSP_MAX_BRACKETS];        // float32(2**(ev_i - target))
    float bias[PYSP_MAX_BRACKETS][3];       // 1.6**(-0.1*|ev_off*wb[c]|), evaluated by the host in float32
    int brightest;                          // argmax(ev_off)
    float* out; long long out_pitch;
    int32_t* count; long long count_pitch;  // may be null
};

#ifndef PYSP_HOST_EMU
// Row-wise work split shared by the point-wise kernels: a work item is `VEC` consecutive elements of one row, so that
// every access is a 16-byte vector and the only division is a 32-bit one per item.
struct RowItems {
    int per_row; long long total;
    __device__ RowItems(int rows, int cols, int vec) : per_row((cols + vec - 1) / vec), total((long long)rows * ((cols + vec - 1) / vec)) {}
};
#define PYSP_ROW_ITEMS(ri, y, c)                                                                              \
    for (long long it_ = blockIdx.x * (long long)blockDim.x + threadIdx.x, y = 0, c = 0;                      \
         it_ < (ri).total && ((y = it_ / (ri).per_row), (c = it_ - y * (ri).per_row), true);                  \
         it_ += (long long)gridDim.x * blockDim.x)

__device__ __forceinline__ float norm_site(float raw, float black, float white) {
    return fminf(fmaxf(raw - black, 0.0f), white) / white;            // normalization.py:20-23
}

__global__ void __launch_bounds__(256) normalize_kernel(NormalizeParams p) {
    // one item = eight photosites of a row: one 16-byte load, two 16-byte stores
    const RowItems ri(p.H, p.W, 8);
    const bool vec = (p.W % 8 == 0) && (p.in_pitch % 16 == 0) && (p.out_pitch % 16 == 0) &&
                     (((size_t)p.in | (size_t)p.out) % 16 == 0);
    PYSP_ROW_ITEMS(ri, y, c) {
        const int x0 = (int)c * 8;
        const uint16_t* src = (const uint16_t*)((const char*)p.in + y * p.in_pitch) + x0;
        float* dst = (float*)((char*)p.out + y * p.out_pitch) + x0;
        const int pos = ((int)y & 1) << 1;
        const float b0 = p.black[pos], b1 = p.black[pos + 1], w0 = p.white[pos], w1 = p.white[pos + 1];
        if (vec) {
            const uint4 v = *(const uint4*)src;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            float o[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[2 * k] = norm_site((float)(w[k] & 0xFFFFu), b0, w0);
                o[2 * k + 1] = norm_site((float)(w[k] >> 16), b1, w1);
            }
            *(float4*)dst = make_float4(o[0], o[1], o[2], o[3]);
            *(float4*)(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
            for (int k = 0; k < 8 && x0 + k < p.W; ++k) dst[k] = norm_site((float)src[k], (k & 1) ? b1 : b0, (k & 1) ? w1 : w0);
        }
    }
}

__global__ void __launch_bounds__(256) matrix_kernel(MatrixParams p) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
        const float* s = p.in + 3 * i;
        float c0 = s[0], c1 = s[1], c2 = s[2];
        if (p.clip) { c0 = clip01(c0); c1 = clip01(c1); c2 = clip01(c2); }
        float r = dot3_f64(p.m + 0, c0, c1, c2), g = dot3_f64(p.m + 3, c0, c1, c2), b = dot3_f64(p.m + 6, c0, c1, c2);
        if (p.gamma) { r = srgb_gamma(r); g = srgb_gamma(g); b = srgb_gamma(b); }
        if (p.out_f16) {
            __half* o = (__half*)p.out + 3 * i;
            o[0] = __float2half_rn(r); o[1] = __float2half_rn(g); o[2] = __float2half_rn(b);
        } else {
            float* o = (float*)p.out + 3 * i;
            o[0] = r; o[1] = g; o[2] = b;
        }
    }
}

// RawDemosaicData.wb_apply / wb_undo (base_types/image_base.py:45-60) and clip_rgb (colorize/transform.py:6-19) on
// [n][3] float32 pixels.  mode 0: x * wb[c].  mode 1: float32(float64(x [* max_wb]) / wb[c]).  mode 2: clip to [0,1].
// The coefficients keep the dtype the caller's object has, as NumPy's promotion does in the reference: float32 coefficients
// multiply in float32; float64 ones (wb_f64) make the product a float64 that is rounded to float32 once; the normalisation
// factor max(wb) multiplies in float32 unless it is a NumPy float64 scalar (max_f64), which promotes the image to float64.
struct WbParams { const float* in; float* out; long long n; double wb[3]; double max_wb; int mode; int normalized; int wb_f64, max_f64; };
__global__ void __launch_bounds__(256) wb_kernel(WbParams p) {
    const long long total = 3 * p.n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % 3);
        float v = p.in[i];
        if (p.mode == 0) {
            v = p.wb_f64 ? __double2float_rn(__dmul_rn((double)v, p.wb[c])) : v * (float)p.wb[c];
        } else if (p.mode == 1) {
            double d = (double)v;
            if (p.normalized) d = p.max_f64 ? __dmul_rn(d, p.max_wb) : (double)(v * (float)p.max_wb);
            v = __double2float_rn(__ddiv_rn(d, p.wb[c]));
        } else {
            v = clip01(v);
        }
        p.out[i] = v;
    }
}

// cv2.cvtColor(float32 RGB, COLOR_RGB2LAB) alone (debayer/ahd.py:58,62), for stage tests: the same lab_lookup the fused
// kernel calls, unpacked to float Lab exactly as cv2 returns it (L = v*100/16384, a/b = v/64 - 128)
struct LabParams { const float* in; float* out; long long n; const uint4* lut; };
__global__ void __launch_bounds__(256) lab_kernel(LabParams p) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x) {
        const LabQ q = lab_lookup(p.lut, p.in[3 * i], p.in[3 * i + 1], p.in[3 * i + 2]);
        p.out[3 * i] = q.L;
        p.out[3 * i + 1] = ((float)(q.ab & 0xFFFFu) * 0.015625f) - 128.0f;
        p.out[3 * i + 2] = ((float)(q.ab >> 16) * 0.015625f) - 128.0f;
    }
}

__global__ void __launch_bounds__(256) gamma_kernel(GammaParams p) {
    const long long n4 = (((size_t)p.in | (size_t)p.out) % 16 == 0) ? p.n / 4 : 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = ((const float4*)p.in)[i];
        ((float4*)p.out)[i] = make_float4(srgb_gamma(v.x), srgb_gamma(v.y), srgb_gamma(v.z), srgb_gamma(v.w));
    }
    for (long long i = 4 * n4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x)
        p.out[i] = srgb_gamma(p.in[i]);
}

// raw_hdr.py:108-148.  VEC: an item is four photosites of a row (one 16-byte access per bracket and output; the host checks
// alignment), else one photosite.  Brackets are accumulated strictly in list order; the next bracket's load is issued before
// the current one is accumulated, the occupancy of a 32-register kernel does the rest of the latency hiding.
template <bool VEC>
__global__ void __launch_bounds__(256) fuse_kernel(const FuseParams p) {
    constexpr int V = VEC ? 4 : 1;
    const RowItems ri(p.H, p.W, V);
    PYSP_ROW_ITEMS(ri, y, c) {
        const int x0 = (int)c * V;
        const int ch0 = ((int)y & 1) + (VEC ? 0 : (x0 & 1));           // channel of element 0: R G / G B rows; odd elements: + 1
        const long long off = y * p.in_pitch + (long long)x0 * 4;
        float v[V], nx[V], sum_w[V], sum_p[V], bright[V];
        int cnt[V];
#pragma unroll
        for (int j = 0; j < V; ++j) { sum_w[j] = 0.0f; sum_p[j] = 0.0f; bright[j] = 0.0f; cnt[j] = 0; }
        if (VEC) { const float4 t = *(const float4*)((const char*)p.in[0] + off); v[0] = t.x; v[V > 1 ? 1 : 0] = t.y; v[V > 2 ? 2 : 0] = t.z; v[V > 3 ? 3 : 0] = t.w; }
        else v[0] = *(const float*)((const char*)p.in[0] + off);
        for (int k = 0; k < p.n; ++k) {
#pragma unroll
            for (int j = 0; j < V; ++j) nx[j] = v[j];
            if (k + 1 < p.n) {
                if (VEC) { const float4 t = *(const float4*)((const char*)p.in[k + 1] + off); nx[0] = t.x; nx[V > 1 ? 1 : 0] = t.y; nx[V > 2 ? 2 : 0] = t.z; nx[V > 3 ? 3 : 0] = t.w; }
                else nx[0] = *(const float*)((const char*)p.in[k + 1] + off);
            }
            const float bias0 = p.bias[k][ch0], bias1 = p.bias[k][VEC ? ch0 + 1 : ch0], ev = p.ev_off[k];
            const bool is_bright = k == p.brightest;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const float w = (0.5f - fabsf(v[j] - 0.5f)) * ((j & 1) ? bias1 : bias0);     // raw_hdr.py:135-141
                sum_w[j] = sum_w[j] + w;
                sum_p[j] = sum_p[j] + ((v[j] * w) * ev);
                cnt[j] += w > 0.0f ? 1 : 0;
                if (is_bright) bright[j] = v[j] * ev;
                v[j] = nx[j];
            }
        }
        float o[V];
#pragma unroll
        for (int j = 0; j < V; ++j) { const float q = sum_p[j] / sum_w[j]; o[j] = (sum_w[j] == 0.0f) ? bright[j] : q; }
        float* dst = (float*)((char*)p.out + y * p.out_pitch) + x0;
        int32_t* dc = p.count ? (int32_t*)((char*)p.count + y * p.count_pitch) + x0 : nullptr;
        if (VEC) {
            *(float4*)dst = make_float4(o[0], o[V > 1 ? 1 : 0], o[V > 2 ? 2 : 0], o[V > 3 ? 3 : 0]);
            if (dc) *(int4*)dc = make_int4(cnt[0], cnt[V > 1 ? 1 : 0], cnt[V > 2 ? 2 : 0], cnt[V > 3 ? 3 : 0]);
        } else {
            dst[0] = o[0];
            if (dc) dc[0] = cnt[0];
        }
    }
}
#endif

}  // namespace pysp
