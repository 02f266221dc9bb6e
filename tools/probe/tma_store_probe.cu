// Standalone probe of the TMA store path: which box origins are legal?
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../pysp_b200/csrc/tma.cuh"
using namespace pysp;

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int box_w, int box_h) {
    extern __shared__ __align__(128) char smem[];
    float* tile = (float*)smem;
    for (int i = threadIdx.x; i < box_w * box_h; i += blockDim.x) tile[i] = 1000.0f + i;
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) { tma_store_2d(tile, &map, x, y); tma_store_commit(); tma_store_wait_read(); }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    int rows = 64, cols = 288, box_w = 180, box_h = 28, x = atoi(argv[1]), y = atoi(argv[2]);
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    float* d; cudaMalloc(&d, rows * cols * 4); cudaMemset(d, 0, rows * cols * 4);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    probe<<<1, 128, box_w * box_h * 4>>>(m, x, y, box_w, box_h);
    cudaError_t e = cudaDeviceSynchronize();
    printf("x=%d y=%d encode=%d kernel: %s ", x, y, (int)r, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> h(rows * cols); cudaMemcpy(h.data(), d, rows * cols * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int gy = 0; gy < rows; ++gy) for (int gx = 0; gx < cols; ++gx) {
            int r2 = gy - y, c = gx - x;
            float want = (r2 >= 0 && r2 < box_h && c >= 0 && c < box_w) ? 1000.0f + r2 * box_w + c : 0.0f;
            bad += h[gy * cols + gx] != want;
        }
        printf("bad=%d", bad);
    }
    printf("\n");
    return 0;
}
