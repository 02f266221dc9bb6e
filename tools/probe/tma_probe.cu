// Standalone probe of the TMA load path (same helpers as pysp_b200/csrc/tma.cuh).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../pysp_b200/csrc/tma.cuh"
using namespace pysp;

template <typename T>
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int box_w, int box_h, T* out, int variant) {
    extern __shared__ __align__(128) char smem[];
    uint64_t* bar = (uint64_t*)smem;
    T* stage = (T*)(smem + 128);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (variant & 1) fence_async_smem();
        mbar_expect_tx(bar, box_w * box_h * sizeof(T));
        tma_load_2d(stage, &map, x, y, bar);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < box_w * box_h; i += blockDim.x) out[i] = stage[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
int run(int rows, int cols, int box_w, int box_h, int x, int y, int variant, CUtensorMapDataType dt) {
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    std::vector<T> h(rows * cols);
    for (int i = 0; i < rows * cols; ++i) h[i] = (T)(i % 1000 + 1);
    T *d, *o;
    cudaMalloc(&d, h.size() * sizeof(T)); cudaMalloc(&o, box_w * box_h * sizeof(T));
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(T)};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d ", (int)r);
    int smem = 128 + box_w * box_h * sizeof(T) + 128;
    cudaFuncSetAttribute(probe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<T><<<1, 128, smem>>>(m, x, y, box_w, box_h, o, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<T> res(box_w * box_h);
        cudaMemcpy(res.data(), o, res.size() * sizeof(T), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < box_h; ++r2) for (int c = 0; c < box_w; ++c) {
            int gy = y + r2, gx = x + c;
            T want = (gy >= 0 && gy < rows && gx >= 0 && gx < cols) ? h[gy * cols + gx] : (T)0;
            bad += res[r2 * box_w + c] != want;
        }
        printf("bad=%d", bad);
    }
    printf("\n");
    return 0;
}

int main(int argc, char** argv) {
    int v = atoi(argv[1]);
    switch (v) {
        case 0: return run<float>(64, 64, 16, 16, 8, 8, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        case 1: return run<float>(64, 64, 16, 16, 8, 8, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        case 2: return run<float>(64, 64, 16, 16, -6, -6, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        case 3: return run<uint16_t>(64, 96, 72, 40, -6, -6, 1, CU_TENSOR_MAP_DATA_TYPE_UINT16);
        case 4: return run<uint16_t>(8, 8, 72, 40, -6, -6, 1, CU_TENSOR_MAP_DATA_TYPE_UINT16);
        case 5: return run<float>(100, 96, 68, 36, 10, 10, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        case 7: return run<float>(64, 64, 16, 16, 9, 8, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        case 8: return run<float>(64, 64, 16, 16, -8, -8, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        case 9: return run<float>(64, 64, 16, 16, 12, 3, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        case 10: return run<float>(100, 96, 68, 36, 12, 10, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
        case 11: return run<uint16_t>(4000, 6000, 72, 40, 592, 274, 1, CU_TENSOR_MAP_DATA_TYPE_UINT16);
        case 12: return run<uint16_t>(4000, 6000, 72, 40, -8, -6, 1, CU_TENSOR_MAP_DATA_TYPE_UINT16);
        case 6: return run<uint16_t>(4000, 6000, 72, 40, 594, 274, 1, CU_TENSOR_MAP_DATA_TYPE_UINT16);
    }
    return 0;
}
