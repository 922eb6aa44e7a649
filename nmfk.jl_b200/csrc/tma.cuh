// Tensor-map TMA helpers shared by the tcgen05 kernels (fro_gemm.cu, kl_tiled_tc2.cu): cuTensorMapEncodeTiled through the
// runtime's driver entry point (no link-time dependency on libcuda), and the 2-D tile load (SASS: UTMALDG).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

namespace nmfk {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// FP32 matrix with `inner` contiguous elements per row, `outer` rows `ld` floats apart -> boxes of box_inner x box_outer;
// out-of-range elements read as zero
inline bool tma_make_map_f32(CUtensorMap* map, const void* base, long long inner, long long outer, long long ld, int box_inner,
                             int box_outer, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = tma_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// c0 = coordinate along the contiguous dimension, c1 = row; completion (full box bytes) counted on `bar`
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     tc::smem_u32(dst)),
                 "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_map(const CUtensorMap* map) { asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory"); }

}  // namespace nmfk
