// tma2d.cuh -- 2-D tiled TMA loads of raw GGUF rows: W viewed as [N rows][row_bytes] with a box of
// (rows x bytes).  The tensor map is encoded on the host per call (a few hundred ns, captured by value in
// CUDA graphs) through the runtime's driver-entry-point query, so the library does not link libcuda.
#pragma once

#include <cuda.h>          // CUtensorMap (types only)
#include <string.h>

#include "ggb_common.cuh"

typedef CUresult (*GgbEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline GgbEncodeTiledFn ggb_encode_tiled_fn() {
    static GgbEncodeTiledFn fn = []() -> GgbEncodeTiledFn {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<GgbEncodeTiledFn>(ptr);
    }();
    return fn;
}

// W [nrows][row_bytes] (row_bytes % 16 == 0, W 16-byte aligned), box = box_rows x box_bytes (box_bytes % 16 == 0,
// <= 1024).  Elements are uint32, so the inner coordinate of a load is byte_offset / 4 (byte_offset % 16 == 0).
// Rows past nrows and bytes past the row end are zero-filled.  Returns GGB_OK or a negative status.
inline int ggb_make_row_tensor_map(CUtensorMap *tmap, const void *W, int64_t nrows, int64_t row_bytes, int box_rows,
                                   int box_bytes, const char *who) {
    memset(tmap, 0, sizeof *tmap);
    GgbEncodeTiledFn enc = ggb_encode_tiled_fn();
    if (!enc) {
        ggb_set_error("%s: cuTensorMapEncodeTiled is not available from this driver", who);
        return GGB_E_CUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)(row_bytes / 4), (cuuint64_t)nrows};
    const cuuint64_t gstride[1] = {(cuuint64_t)row_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)(box_bytes / 4), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(W), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ggb_set_error("%s: cuTensorMapEncodeTiled failed (%d)", who, (int)r);
        return GGB_E_CUDA;
    }
    return GGB_OK;
}

#ifdef __CUDACC__
__device__ __forceinline__ void tma_load_2d_hint(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
#endif
