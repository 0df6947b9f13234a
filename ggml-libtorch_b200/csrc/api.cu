// api.cu -- the C ABI of libggml_b200.so (include/ggml_b200.h): argument validation, error plumbing and
// per-device caches.  No torch types, no allocation, no synchronisation: every entry point validates on
// the host, enqueues on the caller's stream and returns a status code.
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "ggb_common.cuh"

// kernels (one translation unit each)
int ggb_dequantize_impl(const void *W, int type, int64_t nelem, void *out, int out_dtype, cudaStream_t stream);
int ggb_quantize_q8_1_impl(const void *X, int x_dtype, int64_t rows, int64_t K, void *out, int layout, int wtype,
                           cudaStream_t stream);
int ggb_mmvq_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                  void *ws, int64_t ws_bytes, cudaStream_t stream);
int64_t ggb_mmvq_workspace_bytes_impl(int type, int64_t B, int64_t K, int64_t N);
int ggb_mmq_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t M, int64_t K, int64_t N,
                 int64_t ldy, void *ws, int64_t ws_bytes, cudaStream_t stream);
int64_t ggb_mmq_workspace_bytes_impl(int type, int64_t M, int64_t K, int64_t N);
void ggb_mmvq_sb_set_static_weights(int on);
void ggb_mmvq_sb_tune(int warps, int smem_cap_kb);
void ggb_mmq_tune(int nb);
void ggb_mmvq_tune_tiled(int max_box_bytes);
void ggb_mmvq_sb_set_l2_prefetch(int mb);
void ggb_mmvq_mma_tune(int cs, int s_log2_plus1, int min_batch);
void ggb_mmvq_rg_set_mode(int mode);
int ggb_mmvq_mma_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                      void *ws, int64_t ws_bytes, cudaStream_t stream);
int ggb_mmvq_sb_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                     void *ws, int64_t ws_bytes, cudaStream_t stream);
void ggb_mmvq_sb_set_prequant_min_k(int k);
int ggb_moe_mmvq_impl(const void *W, const void *X, const int32_t *ids, void *Y, int type, int xy_dtype, int64_t tokens,
                      int64_t top_k, int64_t n_experts, int64_t K, int64_t N, cudaStream_t stream);

// ---------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

void ggb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------- per-device cache
const DeviceInfo &ggb_device_info() {
    static DeviceInfo info[64];
    static std::once_flag once[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    std::call_once(once[dev], [dev]() {
        int v = 0;
        info[dev].dev = dev;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        info[dev].sm_count = v > 0 ? v : 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        info[dev].max_smem_optin = v > 0 ? v : 232448;
    });
    return info[dev];
}

// ---------------------------------------------------------------- type table
static int block_elems(int type) {
    switch (type) {
        case T_Q4_0: case T_Q4_1: case T_Q5_0: case T_Q5_1: case T_Q8_0: return 32;
        case T_Q2_K: case T_Q3_K: case T_Q4_K: case T_Q5_K: case T_Q6_K: return 256;
        case T_IQ4_NL: return 32;
        case T_IQ4_XS: return 256;
        case T_IQ2_XXS: case T_IQ2_XS: case T_IQ2_S: case T_IQ3_XXS: case T_IQ3_S: case T_IQ1_S: case T_IQ1_M: return 256;
        default: return 0;
    }
}
static int block_bytes(int type) {
    switch (type) {
        case T_Q4_0: return 18;  case T_Q4_1: return 20;  case T_Q5_0: return 22;
        case T_Q5_1: return 24;  case T_Q8_0: return 34;
        case T_Q2_K: return 84;  case T_Q3_K: return 110; case T_Q4_K: return 144;
        case T_Q5_K: return 176; case T_Q6_K: return 210;
        case T_IQ4_NL: return 18; case T_IQ4_XS: return 136;
        case T_IQ2_XXS: return 66; case T_IQ2_XS: return 74; case T_IQ2_S: return 82;
        case T_IQ3_XXS: return 98; case T_IQ3_S: return 110; case T_IQ1_S: return 50; case T_IQ1_M: return 56;
        default: return 0;
    }
}

static bool mmq_has_no_case(int type) {      // the IQ formats: dequantize + MMVQ only (the reference's mmq.cu:222-252 has no case either)
    switch (type) {
        case T_IQ4_NL: case T_IQ4_XS: case T_IQ2_XXS: case T_IQ2_XS: case T_IQ2_S: case T_IQ3_XXS: case T_IQ3_S: case T_IQ1_S:
        case T_IQ1_M: return true;
        default: return false;
    }
}

static bool check_type(const char *fn, int type) {
    if (block_elems(type) == 0) {
        ggb_set_error("%s: unsupported ggml type %d", fn, type);
        return false;
    }
    return true;
}
static bool check_dtype(const char *fn, int dt) {
    if (dt != GGB_F32 && dt != GGB_F16 && dt != GGB_BF16) {
        ggb_set_error("%s: unknown dtype code %d", fn, dt);
        return false;
    }
    return true;
}

// ---------------------------------------------------------------- gather target (see GgbYOut, ggb_common.cuh)
static thread_local const GgbYOut *g_yout = nullptr;
static thread_local const GgbXIn *g_xin = nullptr;
const GgbYOut *ggb_current_yout() { return g_yout; }
const GgbXIn *ggb_current_xin() { return g_xin; }

extern "C" {

int ggb_version(void) { return GGB_VERSION; }
const char *ggb_last_error(void) { return g_err; }
int ggb_block_elems(int type) { return block_elems(type); }
int ggb_block_bytes(int type) { return block_bytes(type); }

int ggb_set_option(int option, int value) {
    g_err[0] = 0;
    switch (option) {
        case GGB_OPT_STATIC_WEIGHTS:
            ggb_mmvq_sb_set_static_weights(value);
            return GGB_OK;
        case GGB_OPT_MMVQ_L2_PREFETCH:
            ggb_mmvq_sb_set_l2_prefetch(value);
            return GGB_OK;
        case GGB_OPT_MMVQ_TILED:
            ggb_mmvq_tune_tiled(value);
            return GGB_OK;
        case GGB_OPT_MMVQ_MMA:
            ggb_mmvq_mma_tune(value & 0xff, (value >> 8) & 0xff, (value >> 16) & 0xff);
            return GGB_OK;
        case GGB_OPT_MMQ_TILES:
            ggb_mmq_tune(value);
            return GGB_OK;
        case GGB_OPT_MMVQ_WARPS:
            ggb_mmvq_sb_tune(value & 0xff, value >> 8);
            return GGB_OK;
        case GGB_OPT_MMVQ_RG:
            ggb_mmvq_rg_set_mode(value);
            return GGB_OK;
        case GGB_OPT_MMVQ_PREQUANT:
            ggb_mmvq_sb_set_prequant_min_k(value);
            return GGB_OK;
        default:
            ggb_set_error("ggb_set_option: unknown option %d", option);
            return GGB_E_ARG;
    }
}

int ggb_dequantize(const void *W, int type, int64_t m, int64_t n, void *out, int out_dtype, void *stream) {
    g_err[0] = 0;
    if (!check_type("ggb_dequantize", type)) return GGB_E_TYPE;
    if (!check_dtype("ggb_dequantize", out_dtype)) return GGB_E_DTYPE;
    if (m < 0 || n < 0 || n % block_elems(type) != 0) {
        ggb_set_error("ggb_dequantize: shape [%lld, %lld] is not a whole number of %d-value blocks per row",
                      (long long)m, (long long)n, block_elems(type));
        return GGB_E_SHAPE;
    }
    if (m == 0 || n == 0) return GGB_OK;
    if (!W || !out) {
        ggb_set_error("ggb_dequantize: null pointer");
        return GGB_E_ARG;
    }
    if (reinterpret_cast<uintptr_t>(out) & 15) {
        ggb_set_error("ggb_dequantize: out must be 16-byte aligned");
        return GGB_E_ARG;
    }
    if (reinterpret_cast<uintptr_t>(W) & 1) {
        ggb_set_error("ggb_dequantize: W must be 2-byte aligned");
        return GGB_E_ARG;
    }
    return ggb_dequantize_impl(W, type, m * n, out, out_dtype, static_cast<cudaStream_t>(stream));
}

int64_t ggb_q8_1_bytes(int64_t rows, int64_t K, int layout) {
    if (rows < 0 || K < 0) return 0;
    const int64_t Kp = (layout == GGB_Q8_1) ? (K + 511) / 512 * 512 : (K - K % 512 + 512);
    return rows * (Kp / 32) * 36;
}

int ggb_quantize_q8_1(const void *X, int x_dtype, int64_t rows, int64_t K, void *out, int layout, int wtype,
                      void *stream) {
    g_err[0] = 0;
    if (!check_dtype("ggb_quantize_q8_1", x_dtype)) return GGB_E_DTYPE;
    if (layout != GGB_Q8_1 && layout != GGB_Q8_1_MMQ) {
        ggb_set_error("ggb_quantize_q8_1: unknown layout %d", layout);
        return GGB_E_ARG;
    }
    if (rows < 0 || K < 0) {
        ggb_set_error("ggb_quantize_q8_1: negative shape");
        return GGB_E_SHAPE;
    }
    if (rows == 0) return GGB_OK;
    if (!out || (K > 0 && !X)) {
        ggb_set_error("ggb_quantize_q8_1: null pointer");
        return GGB_E_ARG;
    }
    if (reinterpret_cast<uintptr_t>(out) & 15) {
        ggb_set_error("ggb_quantize_q8_1: out must be 16-byte aligned");
        return GGB_E_ARG;
    }
    return ggb_quantize_q8_1_impl(X, x_dtype, rows, K, out, layout, wtype, static_cast<cudaStream_t>(stream));
}

static int check_matmul(const char *fn, const void *W, const void *X, const void *Y, int type, int dt,
                        int64_t rows, int64_t K, int64_t N) {
    if (!check_type(fn, type)) return GGB_E_TYPE;
    if (!check_dtype(fn, dt)) return GGB_E_DTYPE;
    if (rows < 0 || N < 0 || K <= 0 || K % block_elems(type) != 0) {
        ggb_set_error("%s: K=%lld must be a positive multiple of the %d-value block", fn, (long long)K,
                      block_elems(type));
        return GGB_E_SHAPE;
    }
    if (rows == 0 || N == 0) return 1;          // nothing to do
    if (!W || !X || !Y) {
        ggb_set_error("%s: null pointer", fn);
        return GGB_E_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(W) & 1) || (reinterpret_cast<uintptr_t>(X) & 3) ||
        (reinterpret_cast<uintptr_t>(Y) & 3)) {
        ggb_set_error("%s: W must be 2-byte aligned, X and Y 4-byte aligned", fn);
        return GGB_E_ARG;
    }
    return GGB_OK;
}

int ggb_mmvq_fused(const void *W, const void *X, const void *X2, int64_t ldx, int type, int xy_dtype, int64_t B, int64_t K,
                   int64_t N, void *const *y_peers, int n_peers, int64_t ldy, int64_t col_offset, int64_t split_row,
                   int64_t col_offset2, void *ws, int64_t ws_bytes, void *stream) {
    g_err[0] = 0;
    const bool split = split_row > 0 && split_row < N;
    if (!y_peers || n_peers < 1 || n_peers > 8 || col_offset < 0 || ldx < K ||
        (split ? (col_offset + split_row > ldy || col_offset2 < 0 || col_offset2 + (N - split_row) > ldy) : col_offset + N > ldy)) {
        ggb_set_error("ggb_mmvq_fused: need 1..8 output buffers, ldx >= K and every output column inside ldy");
        return GGB_E_ARG;
    }
    for (int i = 0; i < n_peers; i++)
        if (!y_peers[i] || (reinterpret_cast<uintptr_t>(y_peers[i]) & 3)) {
            ggb_set_error("ggb_mmvq_fused: output buffer %d is null or not 4-byte aligned", i);
            return GGB_E_ARG;
        }
    if (X2 && (reinterpret_cast<uintptr_t>(X2) & 3)) {
        ggb_set_error("ggb_mmvq_fused: X2 must be 4-byte aligned");
        return GGB_E_ARG;
    }
    const int rc0 = check_matmul("ggb_mmvq_fused", W, X, y_peers[0], type, xy_dtype, B, K, N);
    if (rc0 < 0) return rc0;
    if (rc0 > 0) return GGB_OK;
    if (B > 8) {
        ggb_set_error("ggb_mmvq_fused: batch %lld > 8", (long long)B);
        return GGB_E_SHAPE;
    }
    if (ggb_mmvq_workspace_bytes(type, B, K, N) > ws_bytes || (ws_bytes > 0 && !ws)) {
        ggb_set_error("ggb_mmvq_fused: workspace too small (%lld < %lld)", (long long)ws_bytes,
                      (long long)ggb_mmvq_workspace_bytes(type, B, K, N));
        return GGB_E_ARG;
    }
    if (!ggb_mmvq_fused_supported(type, B, K)) {        // the same verdict as the up-front query, whatever kernels could do it
        ggb_set_error("ggb_mmvq_fused: type %d with B=%lld, K=%lld is not on the fused path (use ggb_mmvq + an all-gather)",
                      type, (long long)B, (long long)K);
        return GGB_E_SHAPE;
    }
    GgbYOut y = {};
    for (int i = 0; i < n_peers; i++) y.base[i] = y_peers[i];
    y.n = n_peers;
    y.ld = ldy;
    y.col0 = col_offset;
    y.split = split ? split_row : 0x7fffffffffffffffLL;
    y.col1 = split ? col_offset2 : 0;
    GgbXIn x = {};
    x.x2 = X2;
    x.ldx = ldx;
    g_yout = &y;
    g_xin = &x;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = ggb_mmvq_mma_impl(W, X, nullptr, type, xy_dtype, B, K, N, ws, ws_bytes, st);
    if (rc == 1) rc = ggb_mmvq_sb_impl(W, X, nullptr, type, xy_dtype, B, K, N, ws, ws_bytes, st);
    g_yout = nullptr;
    g_xin = nullptr;
    if (rc == 1) {
        ggb_set_error("ggb_mmvq_fused: type %d with B=%lld, K=%lld is not on the fused path (use ggb_mmvq + an all-gather)",
                      type, (long long)B, (long long)K);
        return GGB_E_SHAPE;
    }
    return rc;
}

int ggb_mmvq_gather(const void *W, const void *X, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                    void *const *y_peers, int n_peers, int64_t ldy, int64_t col_offset, void *ws, int64_t ws_bytes,
                    void *stream) {
    return ggb_mmvq_fused(W, X, nullptr, K, type, xy_dtype, B, K, N, y_peers, n_peers, ldy, col_offset, 0, 0, ws, ws_bytes,
                          stream);
}

int ggb_mmvq_fused_supported(int type, int64_t B, int64_t K) {
    if (B < 1 || B > 8 || K <= 0 || K % 256 != 0) return 0;
    switch (type) {
        case T_Q4_K: case T_Q5_K: return 1;
        case T_IQ4_XS: return K % 512 == 0 ? 1 : 0;         // 136-byte super-blocks: rows are 16-byte aligned every 2 of them
        case T_Q6_K: return K % 2048 == 0 ? 1 : 0;          // 210-byte super-blocks: rows are 16-byte aligned every 8 of them
        default: return 0;
    }
}

int64_t ggb_mmvq_workspace_bytes(int type, int64_t B, int64_t K, int64_t N) {
    if (block_elems(type) == 0 || B <= 0 || K <= 0 || N <= 0) return 0;
    return ggb_mmvq_workspace_bytes_impl(type, B, K, N);
}

int ggb_mmvq(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
             void *ws, int64_t ws_bytes, void *stream) {
    g_err[0] = 0;
    const int rc = check_matmul("ggb_mmvq", W, X, Y, type, xy_dtype, B, K, N);
    if (rc < 0) return rc;
    if (rc > 0) return GGB_OK;
    if (B > 8) {
        ggb_set_error("ggb_mmvq: batch %lld > 8 (use ggb_mmq)", (long long)B);
        return GGB_E_SHAPE;
    }
    if (ggb_mmvq_workspace_bytes(type, B, K, N) > ws_bytes || (ws_bytes > 0 && !ws)) {
        ggb_set_error("ggb_mmvq: workspace too small (%lld < %lld)", (long long)ws_bytes,
                      (long long)ggb_mmvq_workspace_bytes(type, B, K, N));
        return GGB_E_ARG;
    }
    return ggb_mmvq_impl(W, X, Y, type, xy_dtype, B, K, N, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int ggb_moe_mmvq(const void *W, const void *X, const int32_t *topk_ids, void *Y, int type, int xy_dtype, int64_t tokens,
                 int64_t top_k, int64_t n_experts, int64_t K, int64_t N, void *stream) {
    g_err[0] = 0;
    if (top_k < 1 || n_experts < 1 || tokens < 0 || tokens > ((int64_t)1 << 40) / top_k) {
        ggb_set_error("ggb_moe_mmvq: need top_k >= 1, n_experts >= 1, tokens >= 0 (got %lld, %lld, %lld)", (long long)top_k,
                      (long long)n_experts, (long long)tokens);
        return GGB_E_SHAPE;
    }
    if (top_k > 65535 || n_experts > 0x7fffffff) {
        ggb_set_error("ggb_moe_mmvq: top_k %lld / n_experts %lld out of range", (long long)top_k, (long long)n_experts);
        return GGB_E_SHAPE;
    }
    const int rc = check_matmul("ggb_moe_mmvq", W, X, Y, type, xy_dtype, tokens, K, N);
    if (rc < 0) return rc;
    if (rc > 0) return GGB_OK;
    if (!topk_ids || (reinterpret_cast<uintptr_t>(topk_ids) & 3)) {
        ggb_set_error("ggb_moe_mmvq: topk_ids must be a 4-byte aligned device pointer to int32");
        return GGB_E_ARG;
    }
    return ggb_moe_mmvq_impl(W, X, topk_ids, Y, type, xy_dtype, tokens, top_k, n_experts, K, N,
                             static_cast<cudaStream_t>(stream));
}

int64_t ggb_mmq_workspace_bytes(int type, int64_t M, int64_t K, int64_t N) {
    if (block_elems(type) == 0 || mmq_has_no_case(type) || M <= 0 || K <= 0 || N <= 0) return 0;
    return ggb_mmq_workspace_bytes_impl(type, M, K, N);
}

int ggb_mmq(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t M, int64_t K, int64_t N,
            int64_t ldy, void *ws, int64_t ws_bytes, void *stream) {
    g_err[0] = 0;
    if (mmq_has_no_case(type)) {
        // the reference's ggml_mul_mat_a8 has no case for the IQ types (mmq.cu:222-252 leaves Y uninitialised)
        ggb_set_error("ggb_mmq: ggml type %d has dequantize and MMVQ only (no MMQ in the reference either)", type);
        return GGB_E_TYPE;
    }
    const int rc = check_matmul("ggb_mmq", W, X, Y, type, xy_dtype, M, K, N);
    if (rc < 0) return rc;
    if (rc > 0) return GGB_OK;
    if (ldy < N) {
        ggb_set_error("ggb_mmq: ldy=%lld < N=%lld", (long long)ldy, (long long)N);
        return GGB_E_SHAPE;
    }
    const int64_t need = ggb_mmq_workspace_bytes(type, M, K, N);
    if (need > ws_bytes || !ws || (reinterpret_cast<uintptr_t>(ws) & 255)) {
        ggb_set_error("ggb_mmq: workspace must be 256-byte aligned and >= %lld bytes (got %lld)", (long long)need,
                      (long long)ws_bytes);
        return GGB_E_ARG;
    }
    return ggb_mmq_impl(W, X, Y, type, xy_dtype, M, K, N, ldy, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
