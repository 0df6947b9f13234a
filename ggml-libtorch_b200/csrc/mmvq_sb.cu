// mmvq_sb.cu -- dispatch and process-wide tuning state of the fast MMVQ path (mmvq_sb_body.cuh).  The kernels are
// instantiated in one translation unit per weight type (mmvq_sb_q4k.cu, mmvq_sb_q5k.cu, mmvq_sb_q6k.cu,
// mmvq_sb_iq4xs.cu) so that they compile in parallel.
#include "ggb_common.cuh"

namespace ggb_sb {
int g_pdl_prefetch = 0;      // ggb_set_option(GGB_OPT_STATIC_WEIGHTS, 1)
int g_force_warps = 0;       // tuning: 8 or 16 warps (0 = automatic)
int g_smem_cap_kb = 0;       // tuning: dynamic shared memory budget per CTA in KB (0 = automatic)
int g_staged = 1;            // tuning: 0 = quantize the whole activation row(s) up front
int g_defer_reduce = 1;      // tuning: 0 = reduce every slot right after computing it
int g_prequant_min_k = 1 << 30; // B <= 2: every fused-SwiGLU activation with K >= 4096 (and plain rows at least this long) gets its q8_1
                             // activation image from a pre-launch instead of quantizing in all 148 CTAs (0 = never); needs the workspace.
                             // Measured: SwiGLU down projection K = 28672 40.9 -> 33.2 us (Q4_K), 44.0 -> 39.4 (Q6_K); PLAIN activations
                             // lose (K = 14336: Q4_K 10.0 -> 11.3 us, Q6_K 17.1 -> 22.5: two dependency resolutions instead of one)
int g_l2_prefetch_mb = 0;    // L2 prefetch budget per launch in MB (0 = off: measured neutral on the Llama-8B shapes, +1.5 %
                             // on the 70B ones, and every thread's prefetch is a serialised uniform-datapath instruction)
}  // namespace ggb_sb
using namespace ggb_sb;

namespace {
unsigned long long *g_trace = nullptr;   // debug: device buffer of 8 * 256 stamps per launch (ring kernel: 4 per CTA, rg kernel: 8)
int g_trace_launch = 0;
}  // namespace

unsigned long long *ggb_mmvq_trace_slot() { return g_trace ? g_trace + (size_t)(g_trace_launch++) * 8 * 256 : nullptr; }

void ggb_mmvq_sb_set_static_weights(int on) { g_pdl_prefetch = on ? 1 : 0; }
int ggb_mmvq_static_weights() { return g_pdl_prefetch; }
int ggb_static_weights() { return g_pdl_prefetch; }
int ggb_mmvq_l2_prefetch_mb() { return g_l2_prefetch_mb; }
extern "C" void ggb_debug_mmvq_trace(void *device_buffer) {      // debug only (declared in ggml_b200.h): 8 * 256 u64 per launch
    g_trace = static_cast<unsigned long long *>(device_buffer);
    g_trace_launch = 0;
}
void ggb_mmvq_sb_set_l2_prefetch(int mb) { g_l2_prefetch_mb = mb < 0 ? 0 : mb; }
void ggb_mmvq_sb_tune(int warps, int smem_cap_kb) {
    g_defer_reduce = (warps & 0x40) ? 0 : 1;
    g_staged = (warps & 0x80) ? 0 : 1;
    warps &= ~0x80;
    warps &= 0x3f;
    g_force_warps = warps;
    g_smem_cap_kb = smem_cap_kb;
}

// one per translation unit mmvq_sb_<type>.cu
#define GGB_SB_RUN(name) \
    int name(const void *W, const void *X, void *Y, int xy_dtype, int64_t B, int64_t K, int64_t N, void *ws, int64_t ws_bytes, \
             cudaStream_t stream); \
    int64_t name##_image_bytes(int64_t B, int64_t K)
GGB_SB_RUN(ggb_mmvq_sb_run_q4k);
GGB_SB_RUN(ggb_mmvq_sb_run_q5k);
GGB_SB_RUN(ggb_mmvq_sb_run_q6k);
GGB_SB_RUN(ggb_mmvq_sb_run_iq4xs);

// mmvq_rg.cu: batch 1 with register-resident activations (returns 1 when it does not apply; 2 = applies, launch it)
int ggb_mmvq_rg_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                     const void *image, int probe_only, cudaStream_t stream);
int ggb_mmvq_rgx_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                      cudaStream_t stream);
namespace ggb_sb { extern int g_rg_mode; }
int ggb_mmvq_sb_prequant_q4k_b1(const void *X, int xy_dtype, int64_t K, void *image, cudaStream_t stream);
int ggb_mmvq_rg_wants_image(int type, int64_t B, int64_t K, const void *x2);
int64_t ggb_mmvq_rg_image_bytes(int64_t K);
int ggb_mmvq_rg_prequant(const void *X, int type, int xy_dtype, int64_t K, void *image, cudaStream_t stream);

// bytes of the pre-launched activation image ggb_mmvq_sb_impl would use for (type, B, K); 0 = it quantizes in-kernel
int64_t ggb_mmvq_sb_workspace_bytes(int type, int64_t B, int64_t K) {
    if (B > 2 || K % 256 != 0 || g_prequant_min_k <= 0 || K < 4096) return 0;
    switch (type) {
        case T_Q4_K: return ggb_mmvq_sb_run_q4k_image_bytes(B, K);
        case T_Q5_K: return ggb_mmvq_sb_run_q5k_image_bytes(B, K);
        case T_Q6_K: return ggb_mmvq_sb_run_q6k_image_bytes(B, K);
        case T_IQ4_XS: return ggb_mmvq_sb_run_iq4xs_image_bytes(B, K);
        default: return 0;
    }
}
void ggb_mmvq_sb_set_prequant_min_k(int k) { g_prequant_min_k = k < 0 ? 0 : k; }

// returns GGB_OK / a negative status, or 1 when this path does not apply (the caller uses mmvq.cu)
int ggb_mmvq_sb_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                     void *ws, int64_t ws_bytes, cudaStream_t stream) {
    if (K % 256 != 0 || K > (1 << 20) || (reinterpret_cast<uintptr_t>(W) & 15)) return 1;
    const GgbXIn xin = ggb_make_xin(K);
    // Q6_K, K = 28672, fused SwiGLU: the register kernel behind its own activation pre-launch (mmvq_rg.cu)
    if (ggb_mmvq_rg_wants_image(type, B, K, xin.x2) && ws != nullptr && (reinterpret_cast<uintptr_t>(ws) & 15) == 0 &&
        ws_bytes >= ggb_mmvq_rg_image_bytes(K) && ggb_mmvq_rg_impl(W, X, Y, type, xy_dtype, B, K, N, ws, 1, stream) == 2) {
        const int rc = ggb_mmvq_rg_prequant(X, type, xy_dtype, K, ws, stream);
        if (rc != GGB_OK) return rc;
        return ggb_mmvq_rg_impl(W, X, Y, type, xy_dtype, B, K, N, ws, 0, stream);
    }
    if (ggb_mmvq_rg_impl(W, X, Y, type, xy_dtype, B, K, N, nullptr, 1, stream) == 2) {
        const void *image = nullptr;
        // (the register kernel shares its activation quantization -- SwiGLU included -- across a CTA pair; a pre-launch in front of
        // it measured slower in the 70B block, 81.0 vs 73.2 us: its CTAs keep 14 SMs from taking their CTA of the GEMV early.  Only
        // an explicit row-length threshold asks for it.)
        if (type == T_Q4_K && ws != nullptr && (reinterpret_cast<uintptr_t>(ws) & 15) == 0 &&
            ws_bytes >= ggb_mmvq_sb_run_q4k_image_bytes(1, K) && g_prequant_min_k > 0 && K >= g_prequant_min_k) {
            const int rc = ggb_mmvq_sb_prequant_q4k_b1(X, xy_dtype, K, ws, stream);
            if (rc != GGB_OK) return rc;
            image = ws;
        }
        return ggb_mmvq_rg_impl(W, X, Y, type, xy_dtype, B, K, N, image, 0, stream);
    }
    if ((g_rg_mode & 2) && B == 1) {             // debug: the generic register-resident kernel for Q5_K / Q6_K / IQ4_XS too
        const int rc = ggb_mmvq_rgx_impl(W, X, Y, type, xy_dtype, B, K, N, stream);
        if (rc != 1) return rc;
    }
    switch (type) {
        case T_Q4_K: return ggb_mmvq_sb_run_q4k(W, X, Y, xy_dtype, B, K, N, ws, ws_bytes, stream);
        case T_Q5_K: return ggb_mmvq_sb_run_q5k(W, X, Y, xy_dtype, B, K, N, ws, ws_bytes, stream);
        case T_Q6_K: return ggb_mmvq_sb_run_q6k(W, X, Y, xy_dtype, B, K, N, ws, ws_bytes, stream);
        case T_IQ4_XS: return ggb_mmvq_sb_run_iq4xs(W, X, Y, xy_dtype, B, K, N, ws, ws_bytes, stream);
        default: return 1;
    }
}
