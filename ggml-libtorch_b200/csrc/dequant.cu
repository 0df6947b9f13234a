// dequant.cu -- bandwidth-bound GGUF block dequantizers for sm_100a.
//
// Replaces ggml_dequantize / ggml_get_to_fp16_cuda / dequantize_block* of the reference
// (hf-kernels/ggml-kernels/ggml/ggml_kernel.cu:68-78, ggml/dequantize.cuh:3-254, 435-568) and the CPU op
// ggml-cpu/custom_ops.cpp:11-36.  Not a port: the reference launches one 32/64-thread CTA per
// super-block, reads bytes one at a time, computes in half and stores 2-byte values at stride 16/32.
//
// Design (one CTA = one tile of 8192 consecutive values = 256 legacy blocks or 32 super-blocks):
//   1. the raw tile bytes (<= 8704 B, contiguous in the row-major weight stream) are staged into
//      shared memory by ONE bulk asynchronous copy (cp.async.bulk -> TMA engine, mbarrier
//      completion); 7-8 CTAs are resident per SM so ~50-60 KB of loads are in flight per SM
//      without any per-thread load instructions or register staging;
//   2. legacy types: each thread unpacks items of 8 consecutive output values; K-quants: the tile's
//      (d*sc, dmin*m) pairs are decoded once into a table, then each thread takes "column units" = 8
//      consecutive bytes of a block and all 16 / 32 values packed in them (unpack_col.cuh).  The results
//      are the CPU op's fp32 arithmetic bit for bit (separately rounded product and sum, never a
//      contracted multiply-add of the two), then ONE round-to-nearest conversion to the output dtype;
//   3. every 8 values are written with one 128-bit store (two for fp32); a warp's stores cover whole
//      64-byte segments.
// Algorithmic bytes per value: block_bytes/QK read + sizeof(out) written.
#include "unpack_col.cuh"
#include "unpack_iq.cuh"

namespace {

constexpr int kTileValues = 8192;
constexpr int kThreads = 256;

// ---- the tile kernel --------------------------------------------------------------------------
template <int TYPE, typename OutT>
__global__ void __launch_bounds__(kThreads)
dequant_tile_kernel(const uint8_t *__restrict__ W, OutT *__restrict__ out, int64_t nblocks, int w_aligned16, int static_w) {
    constexpr int QK = BlockTraits<TYPE>::QK, BS = BlockTraits<TYPE>::BS;
    constexpr int TILE_BLOCKS = kTileValues / QK;
    constexpr int TILE_BYTES = TILE_BLOCKS * BS;           // multiple of 16 for every type
    constexpr int ITEMS_PER_BLOCK = QK / 8;
    static_assert(TILE_BYTES % 16 == 0, "bulk copy size must be a multiple of 16");

    __shared__ __align__(128) uint8_t tile[TILE_BYTES];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * TILE_BLOCKS;
    const int nb = (int)min((int64_t)TILE_BLOCKS, nblocks - b0);
    const uint8_t *src = W + b0 * BS;

    // programmatic dependent launch: the first wave of this grid becomes resident under the tail of the previous
    // kernel; with static weights its tile loads start right away, everything else waits for that kernel
    ggb_pdl_launch_dependents();
    if (nb == TILE_BLOCKS && w_aligned16) {
        // TMA path: one elected thread arms the mbarrier and issues one bulk copy for the tile.
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_expect_tx(&bar, TILE_BYTES);
            if (!static_w) ggb_pdl_wait();
            bulk_g2s_hint(tile, src, TILE_BYTES, &bar, l2_policy_evict_first());
        }
        ggb_pdl_wait();
        __syncthreads();                                   // barrier object is initialised for everyone
        mbar_wait(&bar, 0);
    } else {
        ggb_pdl_wait();
        // ragged last tile or a weight pointer that is not 16-byte aligned: cooperative 2-byte copy
        const int nbytes = nb * BS;
        for (int i = 2 * tid; i < nbytes; i += 2 * kThreads)
            *reinterpret_cast<uint16_t *>(tile + i) = *reinterpret_cast<const uint16_t *>(src + i);
        __syncthreads();
    }

    OutT *dst = out + b0 * QK;
    if constexpr (QK == 256 && TYPE != T_IQ4_XS && !kIsIqGrid<TYPE>) {
        // K-quants: decode every (d*sc, dmin*m) pair of the tile ONCE into a table, then one "column unit" = 8
        // consecutive bytes of a block and all the 16 / 32 values packed in them (unpack_col.cuh).
        using CU = ColUnpack<TYPE>;
        __shared__ float4 tab[TILE_BLOCKS * CU::NSUB];
        bool bad = false;
        for (int e = tid; e < nb * CU::NSUB; e += kThreads) tab[e] = CU::entry(tile + (e / CU::NSUB) * BS, e % CU::NSUB, bad);
        if (!__syncthreads_or(bad)) {
            // consecutive threads take consecutive units: a warp's stores cover whole 64-byte segments
#pragma unroll
            for (int k = 0; k < TILE_BLOCKS * CU::UNITS / kThreads; k++) {
                const int ug = tid + k * kThreads, blk = ug / CU::UNITS;
                if (blk < nb) CU::run(tile + blk * BS, ug % CU::UNITS, tab + blk * CU::NSUB, dst + (int64_t)blk * QK);
            }
            return;
        }
        // a negative / non-finite super-block scale somewhere in the tile: the conversion-based unpacker
        constexpr int BLK_STEP = kThreads / ITEMS_PER_BLOCK;
        const int p = tid % ITEMS_PER_BLOCK, blk0 = tid / ITEMS_PER_BLOCK;
        for (int blk = blk0; blk < nb; blk += BLK_STEP) {
            float y[8];
            Unpack8<TYPE>::run(tile + blk * BS, p, y);
            store8(dst + ((int64_t)blk * ITEMS_PER_BLOCK + p) * 8, y);
        }
    } else {
        // legacy types, IQ4_NL, IQ4_XS and the codebook IQ formats (a warp per super-block).  item = 8 consecutive values.  Thread t always handles octet p = t % ITEMS_PER_BLOCK of blocks
        // t / ITEMS_PER_BLOCK + k * BLK_STEP, so everything derived from p (shifts, byte offsets) is loop-invariant;
        // consecutive threads write consecutive 16-byte chunks (a warp store covers 512 contiguous bytes).
        constexpr int BLK_STEP = kThreads / ITEMS_PER_BLOCK;
        static_assert(kThreads % ITEMS_PER_BLOCK == 0 && TILE_BLOCKS % BLK_STEP == 0, "tile/thread geometry");
        const int p = tid % ITEMS_PER_BLOCK, blk0 = tid / ITEMS_PER_BLOCK;
#pragma unroll
        for (int k = 0; k < TILE_BLOCKS / BLK_STEP; k++) {
            const int blk = blk0 + k * BLK_STEP;
            if (blk < nb) {
                float y[8];
                Unpack8<TYPE>::run(tile + blk * BS, p, y);
                store8(dst + ((int64_t)blk * ITEMS_PER_BLOCK + p) * 8, y);
            }
        }
    }
}

template <int TYPE, typename OutT>
int launch_dequant(const void *W, void *out, int64_t nelem, cudaStream_t stream) {
    constexpr int QK = BlockTraits<TYPE>::QK;
    const int64_t nblocks = nelem / QK;
    const int64_t ntiles = (nblocks + kTileValues / QK - 1) / (kTileValues / QK);
    if (ntiles == 0) return GGB_OK;
    if (ntiles > 0x7fffffffLL) {
        ggb_set_error("ggb_dequantize: %lld values exceed the grid limit", (long long)nelem);
        return GGB_E_SHAPE;
    }
    const int aligned = (reinterpret_cast<uintptr_t>(W) & 15) == 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ntiles);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<dequant_tile_kernel<TYPE, OutT>>(ggb_device_info().dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, dequant_tile_kernel<TYPE, OutT>, static_cast<const uint8_t *>(W),
                                    static_cast<OutT *>(out), nblocks, aligned, ggb_static_weights()));
    return GGB_OK;
}

template <int TYPE>
int dispatch_dtype(const void *W, void *out, int64_t nelem, int out_dtype, cudaStream_t stream) {
    switch (out_dtype) {
        case GGB_F32:  return launch_dequant<TYPE, float>(W, out, nelem, stream);
        case GGB_F16:  return launch_dequant<TYPE, __half>(W, out, nelem, stream);
        case GGB_BF16: return launch_dequant<TYPE, __nv_bfloat16>(W, out, nelem, stream);
        default:
            ggb_set_error("ggb_dequantize: unknown out_dtype %d", out_dtype);
            return GGB_E_DTYPE;
    }
}

}  // namespace

int ggb_dequantize_impl(const void *W, int type, int64_t nelem, void *out, int out_dtype, cudaStream_t stream) {
    switch (type) {
        case T_Q4_0: return dispatch_dtype<T_Q4_0>(W, out, nelem, out_dtype, stream);
        case T_Q4_1: return dispatch_dtype<T_Q4_1>(W, out, nelem, out_dtype, stream);
        case T_Q5_0: return dispatch_dtype<T_Q5_0>(W, out, nelem, out_dtype, stream);
        case T_Q5_1: return dispatch_dtype<T_Q5_1>(W, out, nelem, out_dtype, stream);
        case T_Q8_0: return dispatch_dtype<T_Q8_0>(W, out, nelem, out_dtype, stream);
        case T_Q2_K: return dispatch_dtype<T_Q2_K>(W, out, nelem, out_dtype, stream);
        case T_Q3_K: return dispatch_dtype<T_Q3_K>(W, out, nelem, out_dtype, stream);
        case T_Q4_K: return dispatch_dtype<T_Q4_K>(W, out, nelem, out_dtype, stream);
        case T_Q5_K: return dispatch_dtype<T_Q5_K>(W, out, nelem, out_dtype, stream);
        case T_Q6_K: return dispatch_dtype<T_Q6_K>(W, out, nelem, out_dtype, stream);
        case T_IQ4_NL: return dispatch_dtype<T_IQ4_NL>(W, out, nelem, out_dtype, stream);
        case T_IQ4_XS: return dispatch_dtype<T_IQ4_XS>(W, out, nelem, out_dtype, stream);
        case T_IQ2_XXS: return dispatch_dtype<T_IQ2_XXS>(W, out, nelem, out_dtype, stream);
        case T_IQ2_XS: return dispatch_dtype<T_IQ2_XS>(W, out, nelem, out_dtype, stream);
        case T_IQ2_S: return dispatch_dtype<T_IQ2_S>(W, out, nelem, out_dtype, stream);
        case T_IQ3_XXS: return dispatch_dtype<T_IQ3_XXS>(W, out, nelem, out_dtype, stream);
        case T_IQ3_S: return dispatch_dtype<T_IQ3_S>(W, out, nelem, out_dtype, stream);
        case T_IQ1_S: return dispatch_dtype<T_IQ1_S>(W, out, nelem, out_dtype, stream);
        case T_IQ1_M: return dispatch_dtype<T_IQ1_M>(W, out, nelem, out_dtype, stream);
        default:
            ggb_set_error("ggb_dequantize: unsupported ggml type %d", type);
            return GGB_E_TYPE;
    }
}
