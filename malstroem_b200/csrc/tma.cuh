// tma.cuh — 2-D tile loads with the Tensor Memory Accelerator (cp.async.bulk.tensor.2d, completion on an mbarrier).
//
// The tile kernels of the path stage a 64 x 64 tile plus its apron in shared memory.  Done with per-thread LDG -> STS,
// every one of the ~4.5 k elements costs an index division, a bounds test and two instructions in the issue slots the
// stencil needs (the tile kernels are 75-82 % issue-bound, profiles/r02a); with TMA one thread issues one instruction
// per raster, the copy engine walks the box, and cells outside the raster arrive as NaN (float rasters) or 0 (integer
// rasters) without a branch: out-of-bounds fill replaces the apron tests.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ms {

// Host: tensor map of a row-major rows x cols raster of 4-byte elements with a box of box_rows x box_cols elements
// (box_cols * 4 must be a multiple of 16 bytes).  nan_fill: float raster, cells outside arrive as NaN (else 0).
// Returns false when the raster cannot be described (base not 16-byte aligned, cols not a multiple of 4, driver entry
// point missing): the caller keeps its LDG -> STS path.
bool tma_map_2d(CUtensorMap *out, const void *base, int64_t rows, int64_t cols, int box_rows, int box_cols, bool is_float,
                bool nan_fill);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    // make the initialised barrier visible to the async proxy (the copy engine arrives on it)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// box with its first element at raster (row y, column x) -> dst (128-byte aligned shared memory)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((unsigned long long)(uintptr_t)map), "r"(smem_u32(bar)), "r"(x), "r"(y)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

}  // namespace ms
