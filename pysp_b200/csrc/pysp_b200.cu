// libpysp_b200.so: kernels + C ABI (include/pysp_b200.h).  sm_100a only, no CPU path.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/pysp_b200.h"
#include "ahd_select.cuh"
#include "median_stage.cuh"
#include "pointwise.cuh"
#include "develop_plan.h"

namespace pysp {

constexpr int K1_TW = 60, K1_TH = 28, K1_THREADS = 256;
constexpr int K2_TW = 60, K2_TH = 28, K2_THREADS = 256;

__global__ void __launch_bounds__(K1_THREADS, 2) ahd_select_kernel(const __grid_constant__ SelectParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x - tile_y * p.tiles_x;
    const int x0 = tile_x * K1_TW, y0 = p.y_begin + tile_y * K1_TH;
    const bool edge = x0 < 6 || y0 < 6 || x0 + K1_TW + 6 > p.g.W || y0 + K1_TH + 6 > p.g.H || y0 + K1_TH > p.y_end;
    if (edge) select_tile<K1_TW, K1_TH, true>(p, smem, tile_x, tile_y);
    else select_tile<K1_TW, K1_TH, false>(p, smem, tile_x, tile_y);
}

__global__ void __launch_bounds__(K2_THREADS, 2) median_stage_kernel(const __grid_constant__ MedianParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x - tile_y * p.tiles_x;
    const int x0 = tile_x * K2_TW, y0 = p.y_begin + tile_y * K2_TH;
    const bool edge = x0 < 4 || y0 < 4 || x0 + K2_TW + 4 > p.g.W || y0 + K2_TH + 4 > p.g.H || y0 + K2_TH > p.y_end;
    if (edge) median_tile<K2_TW, K2_TH, true>(p, smem, tile_x, tile_y);
    else median_tile<K2_TW, K2_TH, false>(p, smem, tile_x, tile_y);
}

}  // namespace pysp

using namespace pysp;

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PYSP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return PYSP_OK;
}

// ---- optional per-kernel timing (bench only): CUDA events recorded on the launching stream ---------------
#define PYSP_TIMING_SLOTS 4
#define PYSP_TIMING_MAX 2048
static std::mutex g_tmutex;
static bool g_timing = false;
static struct { cudaEvent_t a, b; int slot; } g_tev[PYSP_TIMING_MAX];
static int g_tn = 0;

struct TimedLaunch {     // records an event pair around one launch when timing is enabled
    cudaStream_t s; int idx;
    TimedLaunch(int slot, cudaStream_t stream) : s(stream), idx(-1) {
        if (!g_timing) return;
        std::lock_guard<std::mutex> lk(g_tmutex);
        if (g_tn >= PYSP_TIMING_MAX) return;
        idx = g_tn++;
        g_tev[idx].slot = slot;
        cudaEventCreate(&g_tev[idx].a); cudaEventCreate(&g_tev[idx].b);
        cudaEventRecord(g_tev[idx].a, s);
    }
    ~TimedLaunch() { if (idx >= 0) cudaEventRecord(g_tev[idx].b, s); }
};

static int ensure_device() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(PYSP_ERR_CUDA, "no CUDA device: %s (pysp_b200 has no CPU path)", cudaGetErrorString(e));
    return PYSP_OK;
}

static int grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    const long long cap = 148LL * 16;          // a few waves of the 148 SMs; kernels grid-stride
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" {

const char* pysp_last_error(void) { return g_err; }
const char* pysp_version(void) { return "pysp_b200 0.1 (sm_100a)"; }
int64_t pysp_kernel_launches(void) { return g_launches.load(); }

void pysp_timing_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(g_tmutex);
    for (int i = 0; i < g_tn; ++i) { cudaEventDestroy(g_tev[i].a); cudaEventDestroy(g_tev[i].b); }
    g_tn = 0;
    g_timing = on != 0;
}

int pysp_timing_collect(double* total_ms, int64_t* launches) {
    std::lock_guard<std::mutex> lk(g_tmutex);
    for (int k = 0; k < PYSP_TIMING_SLOTS; ++k) { total_ms[k] = 0.0; launches[k] = 0; }
    for (int i = 0; i < g_tn; ++i) {
        cudaError_t e = cudaEventSynchronize(g_tev[i].b);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, g_tev[i].a, g_tev[i].b);
        if (e != cudaSuccess) return fail(PYSP_ERR_CUDA, "pysp_timing_collect: %s", cudaGetErrorString(e));
        total_ms[g_tev[i].slot] += ms; launches[g_tev[i].slot] += 1;
        cudaEventDestroy(g_tev[i].a); cudaEventDestroy(g_tev[i].b);
    }
    g_tn = 0;
    return PYSP_OK;
}

int32_t pysp_develop_halo_rows(int32_t stages) { return 6 + 4 * (stages > 0 ? stages : 0); }

int64_t pysp_develop_scratch_bytes(int32_t width, int32_t rows, int32_t stages) {
    return develop_scratch_bytes(width, rows, stages);
}

int64_t pysp_lab_lut_bytes(void) { return 33LL * 33 * 33 * 8; }

int pysp_lab_lut_pack_host(const int16_t* lut, void* packed) {
    if (!lut || !packed) return fail(PYSP_ERR_INVALID, "pysp_lab_lut_pack_host: null pointer");
    uint32_t* o = (uint32_t*)packed;
    for (int i = 0; i < 33 * 33 * 33; ++i) {
        uint32_t L = (uint16_t)lut[3 * i], a = (uint16_t)lut[3 * i + 1], b = (uint16_t)lut[3 * i + 2];
        o[2 * i] = L | (a << 16);
        o[2 * i + 1] = b;
    }
    return PYSP_OK;
}

int pysp_develop(const pysp_develop_args* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    DevelopPlan plan;
    int rc = plan_develop(a, K1_TW, K1_TH, K2_TW, K2_TH, &plan, g_err, sizeof(g_err));
    if (rc) return rc;
    rc = ensure_device();
    if (rc) return rc;
    {
        cudaError_t e1 = cudaFuncSetAttribute(ahd_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              SelectTile<K1_TW, K1_TH>::SMEM_BYTES);
        cudaError_t e2 = cudaFuncSetAttribute(median_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              MedianTile<K2_TW, K2_TH>::SMEM_BYTES);
        if (e1 != cudaSuccess || e2 != cudaSuccess)
            return fail(PYSP_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    }
    {
        TimedLaunch t(0, stream);
        ahd_select_kernel<<<plan.select_tiles, K1_THREADS, SelectTile<K1_TW, K1_TH>::SMEM_BYTES, stream>>>(plan.select);
    }
    rc = check_launch("ahd_select_kernel");
    if (rc) return rc;
    for (int s = 0; s < plan.n_stages; ++s) {
        {
            TimedLaunch t(1, stream);
            median_stage_kernel<<<plan.median_tiles[s], K2_THREADS, MedianTile<K2_TW, K2_TH>::SMEM_BYTES, stream>>>(plan.median[s]);
        }
        rc = check_launch("median_stage_kernel");
        if (rc) return rc;
    }
    return PYSP_OK;
}

int pysp_normalize_u16(const uint16_t* in, int64_t in_pitch, float* out, int64_t out_pitch, int32_t H, int32_t W,
                       const float black[4], const float white[4], void* stream) {
    if (!in || !out || !black || !white) return fail(PYSP_ERR_INVALID, "pysp_normalize_u16: null pointer");
    if (H < 2 || W < 2 || (H & 1) || (W & 1)) return fail(PYSP_ERR_INVALID, "pysp_normalize_u16: dims must be even");
    if (in_pitch < 2LL * W || out_pitch < 4LL * W) return fail(PYSP_ERR_INVALID, "pysp_normalize_u16: bad pitch");
    int rc = ensure_device();
    if (rc) return rc;
    NormalizeParams p;
    p.in = in; p.in_pitch = in_pitch; p.out = out; p.out_pitch = out_pitch; p.H = H; p.W = W;
    const int perm[4] = {0, 1, 3, 2};
    for (int i = 0; i < 4; ++i) { p.black[perm[i]] = black[i]; p.white[perm[i]] = white[i]; }
    normalize_kernel<<<grid_for((long long)H * (W / 2), 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("normalize_kernel");
}

int pysp_cam_to_lin_srgb(const float* in, void* out, int64_t n, const double m[9], int32_t clip, int32_t gamma,
                         int32_t out_f16, void* stream) {
    if (!in || !out || !m || n < 0) return fail(PYSP_ERR_INVALID, "pysp_cam_to_lin_srgb: bad argument");
    if (n == 0) return PYSP_OK;
    int rc = ensure_device();
    if (rc) return rc;
    MatrixParams p;
    p.in = in; p.out = out; p.n = n; p.clip = clip; p.gamma = gamma; p.out_f16 = out_f16;
    for (int i = 0; i < 9; ++i) p.m[i] = m[i];
    matrix_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("matrix_kernel");
}

int pysp_lin_srgb_to_srgb(const float* in, float* out, int64_t n, void* stream) {
    if (!in || !out || n < 0) return fail(PYSP_ERR_INVALID, "pysp_lin_srgb_to_srgb: bad argument");
    if (n == 0) return PYSP_OK;
    int rc = ensure_device();
    if (rc) return rc;
    GammaParams p = {in, out, n};
    gamma_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("gamma_kernel");
}

int pysp_fuse_exposures(const float* const* brackets, int32_t n, int64_t in_pitch, int32_t H, int32_t W,
                        const float* ev_offset, const float* bias, int32_t brightest, float* out, int64_t out_pitch,
                        int32_t* count, int64_t count_pitch, void* stream) {
    if (!brackets || !ev_offset || !bias || !out) return fail(PYSP_ERR_INVALID, "pysp_fuse_exposures: null pointer");
    if (n < 1 || n > PYSP_MAX_BRACKETS) return fail(PYSP_ERR_INVALID, "pysp_fuse_exposures: 1..%d brackets", PYSP_MAX_BRACKETS);
    if (H < 2 || W < 2 || (H & 1) || (W & 1)) return fail(PYSP_ERR_INVALID, "pysp_fuse_exposures: dims must be even");
    if (brightest < 0 || brightest >= n) return fail(PYSP_ERR_INVALID, "pysp_fuse_exposures: bad brightest index");
    int rc = ensure_device();
    if (rc) return rc;
    FuseParams p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < n; ++i) {
        if (!brackets[i]) return fail(PYSP_ERR_INVALID, "pysp_fuse_exposures: null bracket %d", i);
        p.in[i] = brackets[i]; p.ev_off[i] = ev_offset[i];
        for (int c = 0; c < 3; ++c) p.bias[i][c] = bias[3 * i + c];
    }
    p.in_pitch = in_pitch; p.n = n; p.H = H; p.W = W; p.brightest = brightest;
    p.out = out; p.out_pitch = out_pitch; p.count = count; p.count_pitch = count_pitch;
    fuse_kernel<<<grid_for((long long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("fuse_kernel");
}

}  // extern "C"
