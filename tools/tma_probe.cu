// Standalone probe of the TMA tile load used by the tile kernels (tma.cuh): one CTA loads a 66 x 68 float box at
// (x, y) = (-1, -1) of a small raster and prints a few cells.  Variants are chosen on the command line:
//   probe <dst_space 0=cluster 1=cta> <oob 0=zero 1=nan> <l2 0=none 1=128B> <rows> <cols> <box_cols> <box_rows> <x> <y>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int SPACE>
__global__ void k_probe(const __grid_constant__ CUtensorMap map, float *out, int x, int y, unsigned bytes) {
    __shared__ __align__(128) float sz[66 * 68];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        if (SPACE == 0)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(sz)), "l"((unsigned long long)(uintptr_t)&map), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(sz)), "l"((unsigned long long)(uintptr_t)&map), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int k = threadIdx.x; k < 66 * 68; k += blockDim.x) out[k] = sz[k];
}

int main(int argc, char **argv) {
    int space = argc > 1 ? atoi(argv[1]) : 0, oob = argc > 2 ? atoi(argv[2]) : 1, l2 = argc > 3 ? atoi(argv[3]) : 1;
    int rows = argc > 4 ? atoi(argv[4]) : 512, cols = argc > 5 ? atoi(argv[5]) : 512;
    float *d, *o, *h = (float *)malloc((size_t)rows * cols * 4);
    for (int i = 0; i < rows * cols; i++) h[i] = (float)i;
    cudaMalloc(&d, (size_t)rows * cols * 4);
    cudaMalloc(&o, 66 * 68 * 4);
    cudaMemcpy(d, h, (size_t)rows * cols * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * 4};
    int bc = argc > 6 ? atoi(argv[6]) : 68, br = argc > 7 ? atoi(argv[7]) : 66, X = argc > 8 ? atoi(argv[8]) : -1, Y = argc > 9 ? atoi(argv[9]) : -1;
    cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br}, es[2] = {1, 1};
    CUresult r = ((Fn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          oob ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    unsigned bytes = (unsigned)(bc * br * 4);
    if (space == 0) k_probe<0><<<1, 256>>>(map, o, X, Y, bytes); else k_probe<1><<<1, 256>>>(map, o, X, Y, bytes);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        float res[66 * 68];
        cudaMemcpy(res, o, sizeof(res), cudaMemcpyDeviceToHost);
        printf("cells: [0][0]=%g [0][1]=%g [1][0]=%g [1][1]=%g [1][2]=%g [2][1]=%g (want oob oob oob 0 1 %d)\n", res[0], res[1], res[68],
               res[69], res[70], res[2 * 68 + 1], cols);
    }
    return 0;
}
