/*
 * ggml_b200.h -- C ABI of libggml_b200.so: the GGUF quantized-weight hot path (dequantize, q8_1
 * activation quantization, MMVQ, MMQ) as hand-written sm_100a CUDA.
 *
 * This is the drop-in boundary.  Each entry point replaces one host launcher of the reference
 * (paths relative to /root/reference/hf-kernels/ggml-kernels unless noted); the reference-side
 * binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer borrowed for the call;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); work is enqueued
 *     asynchronously, nothing synchronises, nothing allocates -> CUDA-graph capturable;
 *   - the caller has made the right device current (the reference uses OptionalCUDAGuard,
 *     ggml/ggml_kernel.cu:70,85);
 *   - return value 0 = enqueued; negative = GGB_E_* and ggb_last_error() (thread-local) says why.
 *     Nothing throws or exits across the ABI (the reference printf+exit()s, ggml/cuda_utils.h:19-27).
 *   - `type` is a ggml_type id (ggml/ggml-common.h:1128-1161): Q4_0=2 Q4_1=3 Q5_0=6 Q5_1=7 Q8_0=8
 *     Q2_K=10 Q3_K=11 Q4_K=12 Q5_K=13 Q6_K=14; the IQ family -- IQ2_XXS=16 IQ2_XS=17 IQ3_XXS=18 IQ1_S=19 IQ4_NL=20
 *     IQ3_S=21 IQ2_S=22 IQ4_XS=23 IQ1_M=29 -- for ggb_dequantize and ggb_mmvq only (the coverage the reference has
 *     for them: ggml/dequantize.cuh:256-433, ggml/vecdotq.cuh:607-888; ggb_mmq returns GGB_E_TYPE where the
 *     reference's switch, ggml/mmq.cu:222-252, falls through and leaves Y uninitialised).  IQ3_S uses the
 *     reference's codebook (ggml-common.h:685, top level 62), not the later ggml one.
 *   - dtype codes: GGB_F32=0, GGB_F16=1, GGB_BF16=2.
 */
#ifndef GGML_B200_H
#define GGML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGB_VERSION 100

enum ggb_dtype { GGB_F32 = 0, GGB_F16 = 1, GGB_BF16 = 2 };

enum ggb_status {
    GGB_OK = 0,
    GGB_E_TYPE = -1,      /* unsupported ggml_type */
    GGB_E_SHAPE = -2,     /* shape not divisible by the block size / out of range */
    GGB_E_DTYPE = -3,     /* unsupported dtype code */
    GGB_E_ARG = -4,       /* null pointer, misalignment, workspace too small */
    GGB_E_CUDA = -5       /* a CUDA runtime call or launch failed */
};

enum ggb_q8_layout {
    GGB_Q8_1 = 0,         /* rows of block_q8_1 (36 B), K padded to ceil(K/512)*512  -- MMVQ operand */
    GGB_Q8_1_MMQ = 1      /* block_q8_1_mmq (144 B), chunk-major, K padded to K-K%512+512 -- MMQ operand */
};

int         ggb_version(void);
/* Thread-local description of the last failure on this thread ("" if none). */
const char *ggb_last_error(void);

/*
 * Process-wide options (no reference counterpart; the reference has no tunables).
 * GGB_OPT_STATIC_WEIGHTS = 1 declares that weight tensors are never written by work enqueued on the
 * same stream shortly before a matmul (true for inference weights).  MMVQ launches are programmatic
 * dependent launches: with this option the weight stream starts while the PREVIOUS kernel on the stream
 * is still draining (activations are still read, and outputs written, only after it has completed).
 * Default 0: everything waits for the previous kernel.
 */
enum ggb_option {
    GGB_OPT_STATIC_WEIGHTS = 1,   /* see above; honoured by every kernel that streams weights: dequantize, all MMVQ kernels, MMQ */
    GGB_OPT_MMVQ_WARPS = 2,    /* tuning/debug: value = warps (0 auto, 8, 16) | shared-memory cap in KB << 8 */
    GGB_OPT_MMQ_TILES = 3,     /* tuning/debug bits: 1 = one 256-token tile per weight stage, 2 = no clusters, 4 = no tail launch, 8 = no CTA-pair MMA (0 = automatic) */
    GGB_OPT_MMVQ_MMA = 5,      /* tuning/debug: super-blocks per chunk | (log2 K-split + 1) << 8 | minimum batch << 16 (0 fields = automatic) */
    GGB_OPT_MMVQ_TILED = 6,    /* tuning/debug: generic MMVQ kernel, largest TMA box in bytes for the tiled weight stages (0 = per-row bulk copies, -1 = default: per type, measured) */
    GGB_OPT_MMVQ_PREQUANT = 8, /* MMVQ batches 1-2 on the ring kernels: the fused-SwiGLU activation of ggb_mmvq_fused (K >= 4096), and plain rows of at least this many values (default: none), take their q8_1 activation image from a pre-launch into the workspace instead of quantizing in every CTA; 0 = never.  (The batch-1 register kernel decides for itself: rows of K >= 12288 get their own 128-thread pre-launch, shorter ones quantize in-kernel, shared by a CTA pair.) */
    GGB_OPT_MMVQ_RG = 7,       /* tuning/debug: bit 0 = batch-1 kernels with register-resident activations (default 1); bit 1 = also for the types where the generic kernel measured faster; bits 8.. debug */
    GGB_OPT_MMVQ_L2_PREFETCH = 4 /* MB of weights per MMVQ launch pulled into L2 while the previous kernel drains (default 0 = off) */
};
int         ggb_set_option(int option, int value);
/* Options and tuning knobs are process-wide plain integers shared by all devices and threads: set them once, before
 * any concurrent use of the library (the matmul / dequantize entry points themselves are re-entrant and stateless). */

/* Debug only: per-CTA timestamps of the following MMVQ launches are written to `device_buffer` (8 * 256 uint64 per
 * launch, scripts/mmvq_trace.py); NULL switches it off.  Not part of the drop-in surface. */
void        ggb_debug_mmvq_trace(void *device_buffer);

/* 1 when (type, B, K) is on the fused-epilogue path of ggb_mmvq_gather / ggb_mmvq_fused, else 0: a pure query, the
 * same on every rank, so a tensor-parallel caller decides ONCE whether to use the fused exchange. */
int         ggb_mmvq_fused_supported(int type, int64_t B, int64_t K);

/* Block geometry of a ggml_type (0 if unsupported).  ggml/ggml-common.h:1163-1240 (qk) / :17-108. */
int         ggb_block_elems(int type);
int         ggb_block_bytes(int type);

/*
 * Dequantize W (raw GGUF bytes, uint8 [m, n/QK*block_bytes], row-major contiguous) into out [m, n].
 * Replaces ggml_dequantize (ggml/ggml_kernel.cu:68-78) + ggml_get_to_fp16_cuda (ggml/dequantize.cuh:525-568)
 * and, for fp32 output, the CPU op /root/reference/ggml-cpu/custom_ops.cpp:11-36.
 * Arithmetic is the CPU op's fp32 order (separately rounded mul then add/sub) followed by ONE
 * round-to-nearest conversion to out_dtype, so fp32 output is bit-identical to the CPU op and
 * fp16/bf16 output equals torch's .to(dtype) of it.  (The reference CUDA op computes in half.)
 */
int         ggb_dequantize(const void *W, int type, int64_t m, int64_t n,
                           void *out, int out_dtype, void *stream);

/*
 * q8_1 activation quantization of X [rows, K] (x_dtype, contiguous).
 * layout GGB_Q8_1     replaces quantize_row_q8_1_cuda   (ggml/ggml_kernel.cu:13-66)
 * layout GGB_Q8_1_MMQ replaces quantize_mmq_q8_1_cuda   (ggml/mmq.cu:109-177); `wtype` selects
 *                     need_sum (ds = half2(d,sum)) vs fp32 d as mmq_need_sum does (ggml/mmq.cu:84-106).
 * `out` must hold ggb_q8_1_bytes(rows, K, layout) bytes.
 */
int64_t     ggb_q8_1_bytes(int64_t rows, int64_t K, int layout);
int         ggb_quantize_q8_1(const void *X, int x_dtype, int64_t rows, int64_t K,
                              void *out, int layout, int wtype, void *stream);

/*
 * MMVQ: Y[B, N] = q8_1(X[B, K]) . W[N, K]^T, 1 <= B <= 8 (the reference allows B == 1 only).
 * Replaces ggml_mul_mat_vec_a8 (ggml/ggml_kernel.cu:80-193): quantize_row_q8_1_cuda + mul_mat_vec_q
 * (ggml/mmvq.cuh:2-38).  X and Y share `xy_dtype`.
 * Launches: batches 1-2 (and every batch of the types without a tensor-core path: the codebook IQ formats) are ONE launch --
 * the activations are quantized to q8_1 inside the GEMV kernel -- except batch 1 of the K-quants Q4_K / Q5_K / Q6_K on rows
 * of K >= 12288, whose q8_1 activation image is written by a small pre-launch; batches 3-8 of Q2_K .. Q6_K, the five legacy
 * types, IQ4_NL and IQ4_XS (2-8 for K > 8192) are TWO programmatic-dependent launches: the q8_1 activation image is written
 * once into `ws`, then the tensor-core GEMV reads it (in-kernel it was repeated by all 148 CTAs).  `ws` must hold
 * ggb_mmvq_workspace_bytes(type, B, K, N) bytes (0 for the one-launch cases with K < 4096; ~1.3 bytes per activation value
 * otherwise); it is scratch: nothing survives the call.
 */
int64_t     ggb_mmvq_workspace_bytes(int type, int64_t B, int64_t K, int64_t N);
int         ggb_mmvq(const void *W, const void *X, void *Y, int type, int xy_dtype,
                     int64_t B, int64_t K, int64_t N, void *ws, int64_t ws_bytes, void *stream);

/*
 * MMVQ with the all-gather of a column-parallel layer fused into its epilogue (no reference counterpart: the reference is
 * single-GPU).  Same computation as ggb_mmvq on this rank's row slice W [N, K/QK*block_bytes], but
 * every result is stored to y_peers[i][tok * ldy + col_offset + row] for i < n_peers: the caller passes the base pointers
 * of one symmetric buffer [B, ldy] as mapped on every rank of the tensor-parallel group (NVLink peer mappings; its own
 * copy included), ldy = N_total and col_offset = rank * N, so each rank writes its slice into every rank's copy of the
 * full output.  The caller orders the exchange with a cross-rank barrier on the stream before the result is read.
 * Types Q4_K / Q5_K / Q6_K / IQ4_XS with K % 256 == 0 (Q6_K: K % 2048 == 0, IQ4_XS: K % 512 == 0), 1 <= B <= 8 (exactly
 * ggb_mmvq_fused_supported); anything else returns GGB_E_SHAPE and
 * the caller uses ggb_mmvq + an all-gather.  Workspace: ggb_mmvq_workspace_bytes.
 */
int         ggb_mmvq_gather(const void *W, const void *X, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                            void *const *y_peers, int n_peers, int64_t ldy, int64_t col_offset,
                            void *ws, int64_t ws_bytes, void *stream);

/*
 * The general form (the Llama MLP block on quantized weights as two launches).
 *   activation:  X [B, K] with row stride ldx (elements).  X2 != NULL: the activation is silu(X) * X2, computed on the fly
 *                with torch eager's roundings (silu and the product each rounded to the tensor dtype), so a down
 *                projection fed by the gate and up outputs equals ggb_mmvq(W, F.silu(g) * u) bit for bit.
 *   output:      as ggb_mmvq_gather; with 0 < split_row < N, rows >= split_row go to column col_offset2 + (row -
 *                split_row): W = [gate_shard; up_shard] stacked writes the gate slice and the up slice of one
 *                [B, 2 * N_total] buffer in ONE launch.
 * Same types / shapes as ggb_mmvq_gather.
 */
int         ggb_mmvq_fused(const void *W, const void *X, const void *X2, int64_t ldx, int type, int xy_dtype,
                           int64_t B, int64_t K, int64_t N, void *const *y_peers, int n_peers, int64_t ldy,
                           int64_t col_offset, int64_t split_row, int64_t col_offset2,
                           void *ws, int64_t ws_bytes, void *stream);

/*
 * Cross-rank barrier of a tensor-parallel group on the stream, as one programmatic-dependent launch (the barrier the fused
 * exchange above needs before the gathered buffer is read; no reference counterpart).  flag_peers[i] = rank i's copy of a
 * symmetric block of 16 uint32, zero-initialised once, as mapped on this rank; every rank calls it the same number of
 * times.  Word 9 of the rank's own block is set when a peer never arrived (the wait is bounded).  Graph-capturable.
 */
int         ggb_tp_barrier(void *const *flag_peers, int n_peers, int rank, void *stream);

/*
 * Grouped MMVQ for mixture-of-experts decode.  No counterpart in this reference; the op is
 * the ggml_moe_a8_vec of the reference's upstream (vLLM csrc/quantization/gguf: quantize_row_q8_1 + moe_vec_q*), whose
 * argument meaning it keeps:
 *   W         raw GGUF bytes of n_experts stacked matrices, uint8 [n_experts, N, K/QK*block_bytes], contiguous;
 *   X         [tokens, K] activations (xy_dtype);      topk_ids  int32 [tokens * top_k], row-major [tokens, top_k];
 *   Y         [tokens * top_k, N]:  Y[i, :] = q8_1(X[i / top_k, :]) . W[topk_ids[i]]^T   (the caller weights and sums
 *             the top_k rows of a token).  With top_k = 1 and tokens = the number of pairs this is the second (down)
 *             projection of an expert MLP: one activation row per pair.
 * An id outside [0, n_experts) yields a zero output row (the upstream kernel reads out of bounds).  Every ggml_type of
 * ggb_mmvq, any number of tokens, ONE launch per 65535 pairs (grid = row ranges x pairs, each CTA quantizes its pair's
 * activation row in shared memory); no workspace.  Numerics: exactly ggb_mmvq's generic kernel at batch 1 -- integer dots
 * exact, fp32 sums in a (type, K)-dependent order.
 */
int         ggb_moe_mmvq(const void *W, const void *X, const int32_t *topk_ids, void *Y, int type, int xy_dtype,
                         int64_t tokens, int64_t top_k, int64_t n_experts, int64_t K, int64_t N, void *stream);

/*
 * MMQ: Y[M, N] (row stride ldy elements) = q8_1_mmq(X[M, K]) . W[N, K]^T.
 * Replaces ggml_mul_mat_a8 (ggml/mmq.cu:180-255): quantize_mmq_q8_1_cuda + mul_mat_q
 * (ggml/mmq.cuh:1917-2031).  `ws` must hold ggb_mmq_workspace_bytes(...) bytes (the quantized
 * activations; the reference allocates them with torch::empty on every call, ggml/mmq.cu:207-208).
 * Numerics: the tensor cores multiply fp16 operands, A = fp16(dequantize(W)), B = fp16(d8 * q8) with the q8_1
 * quantization of the reference, fp32 accumulation (one extra 2^-11 relative rounding per operand).
 * Range: activations of ANY fp32 / bf16 magnitude -- a token whose largest |x| is outside [2^-8, 2^15] is stored
 * divided by a power of two and its outputs are multiplied back (exact); weights must dequantize to |w| <= 65504
 * (fp16), which every real GGUF tensor does by orders of magnitude.  Launches: the activation image kernel, the GEMM,
 * and for some (M, N) a second GEMM launch over the last partial round of tiles -- all programmatic dependent launches.
 */
int64_t     ggb_mmq_workspace_bytes(int type, int64_t M, int64_t K, int64_t N);
int         ggb_mmq(const void *W, const void *X, void *Y, int type, int xy_dtype,
                    int64_t M, int64_t K, int64_t N, int64_t ldy,
                    void *ws, int64_t ws_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GGML_B200_H */
