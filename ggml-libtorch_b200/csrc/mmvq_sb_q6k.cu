// mmvq_sb_q6k.cu -- the Q6_K instantiations of the fast MMVQ path (see mmvq_sb_body.cuh, mmvq_sb.cu)
#include "mmvq_sb_body.cuh"

int ggb_mmvq_sb_run_q6k(const void *W, const void *X, void *Y, int xy_dtype, int64_t B, int64_t K, int64_t N, void *ws, int64_t ws_bytes,
    cudaStream_t stream) {
    return dispatch_dtype<T_Q6_K>(W, X, Y, xy_dtype, B, K, N, ws, ws_bytes, stream);
}
int64_t ggb_mmvq_sb_run_q6k_image_bytes(int64_t B, int64_t K) { return sb_image_bytes<T_Q6_K>(B, K); }
