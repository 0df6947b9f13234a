// mmvq_sb_q4k.cu -- the Q4_K instantiations of the fast MMVQ path (see mmvq_sb_body.cuh, mmvq_sb.cu)
#include "mmvq_sb_body.cuh"

int ggb_mmvq_sb_run_q4k(const void *W, const void *X, void *Y, int xy_dtype, int64_t B, int64_t K, int64_t N, void *ws, int64_t ws_bytes,
    cudaStream_t stream) {
    return dispatch_dtype<T_Q4_K>(W, X, Y, xy_dtype, B, K, N, ws, ws_bytes, stream);
}
int64_t ggb_mmvq_sb_run_q4k_image_bytes(int64_t B, int64_t K) { return sb_image_bytes<T_Q4_K>(B, K); }

// the batch-1 activation image for mmvq_rg.cu (same layout as this path's: [8][ipr] quants | [2][ipr] scale table)
int ggb_mmvq_sb_prequant_q4k_b1(const void *X, int xy_dtype, int64_t K, void *image, cudaStream_t stream) {
    switch (xy_dtype) {
        case GGB_F32:  return launch_prequant<T_Q4_K, 1, float>(X, K, image, stream);
        case GGB_F16:  return launch_prequant<T_Q4_K, 1, __half>(X, K, image, stream);
        case GGB_BF16: return launch_prequant<T_Q4_K, 1, __nv_bfloat16>(X, K, image, stream);
        default: return 1;
    }
}
