// placeholder replaced by the real kernel
#include "ggb_common.cuh"
int64_t ggb_mmvq_workspace_bytes_impl(int, int64_t, int64_t, int64_t) { return 0; }
int ggb_mmvq_impl(const void *, const void *, void *, int, int, int64_t, int64_t, int64_t, void *, int64_t, cudaStream_t) {
    ggb_set_error("ggb_mmvq: not built"); return GGB_E_TYPE;
}
