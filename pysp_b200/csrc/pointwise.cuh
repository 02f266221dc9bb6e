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
__global__ void __launch_bounds__(256) normalize_kernel(NormalizeParams p) {
    // one thread = two horizontally adjacent photosites
    long long n = (long long)p.H * (p.W >> 1);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int y = (int)(i / (p.W >> 1)), x = (int)(i - (long long)y * (p.W >> 1)) * 2;
        const uint16_t* src = (const uint16_t*)((const char*)p.in + y * p.in_pitch) + x;
        float* dst = (float*)((char*)p.out + y * p.out_pitch) + x;
        int pos = (y & 1) << 1;
        float a = (float)src[0], b = (float)src[1];
        a = fminf(fmaxf(a - p.black[pos], 0.0f), p.white[pos]) / p.white[pos];
        b = fminf(fmaxf(b - p.black[pos + 1], 0.0f), p.white[pos + 1]) / p.white[pos + 1];
        dst[0] = a; dst[1] = b;
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

__global__ void __launch_bounds__(256) gamma_kernel(GammaParams p) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < p.n; i += (long long)gridDim.x * blockDim.x)
        p.out[i] = srgb_gamma(p.in[i]);
}

__global__ void __launch_bounds__(256) fuse_kernel(FuseParams p) {
    long long n = (long long)p.H * p.W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
        int ch = (y & 1) + (x & 1);
        float sum_w = 0.0f, sum_p = 0.0f, bright = 0.0f;
        int cnt = 0;
        for (int k = 0; k < p.n; ++k) {          // strictly in list order (raw_hdr.py:135-141)
            float v = *((const float*)((const char*)p.in[k] + y * p.in_pitch) + x);
            float w = (0.5f - fabsf(v - 0.5f)) * p.bias[k][ch];
            sum_w = sum_w + w;
            sum_p = sum_p + ((v * w) * p.ev_off[k]);
            cnt += w > 0.0f ? 1 : 0;
            if (k == p.brightest) bright = v * p.ev_off[k];
        }
        float q = sum_p / sum_w;
        *((float*)((char*)p.out + y * p.out_pitch) + x) = (sum_w == 0.0f) ? bright : q;
        if (p.count) *((int32_t*)((char*)p.count + y * p.count_pitch) + x) = cnt;
    }
}
#endif

}  // namespace pysp
