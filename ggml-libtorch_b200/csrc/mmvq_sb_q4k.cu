// mmvq_sb_q4k.cu -- the Q4_K instantiations of the fast MMVQ path (see mmvq_sb_body.cuh, mmvq_sb.cu)
#include "mmvq_sb_body.cuh"

int ggb_mmvq_sb_run_q4k(const void *W, const void *X, void *Y, int xy_dtype, int64_t B, int64_t K, int64_t N, cudaStream_t stream) {
    return dispatch_dtype<T_Q4_K>(W, X, Y, xy_dtype, B, K, N, stream);
}
