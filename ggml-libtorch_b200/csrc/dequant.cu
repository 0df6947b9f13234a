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
//   2. each thread unpacks items of 8 consecutive output values: super-block scales / minima are
//      decoded in registers, arithmetic is the CPU op's fp32 order with separately rounded
//      __fmul_rn / __fadd_rn / __fsub_rn (never contracted into FMA), then ONE round-to-nearest
//      conversion to the output dtype;
//   3. every item is written with one 128-bit store (two for fp32); consecutive threads write
//      consecutive 16-byte chunks, so a warp store covers 512 contiguous bytes.
// Algorithmic bytes per value: block_bytes/QK read + sizeof(out) written.
#include "ggb_common.cuh"

namespace {

constexpr int kTileValues = 8192;
constexpr int kThreads = 256;

// ---- per-type scale helpers -------------------------------------------------------------------
// Q4_K / Q5_K packed 6-bit (scale, min) pairs, 12 bytes (reference layout: dequantize.cuh:154-161)
__device__ __forceinline__ void k4_scale_min(int j, const uint8_t *s, int &sc, int &mn) {
    if (j < 4) {
        sc = s[j] & 63;
        mn = s[j + 4] & 63;
    } else {
        sc = (s[j + 4] & 0x0f) | ((s[j - 4] >> 6) << 4);
        mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4);
    }
}
// Q3_K 6-bit signed scales, 12 bytes (gguf-py quants.py:442-454)
__device__ __forceinline__ int q3k_scale(const uint8_t *s, int is) {
    const int lo = (is < 8) ? (s[is] & 0x0f) : (s[is - 8] >> 4);
    const int hi = (s[8 + (is & 3)] >> (2 * (is >> 2))) & 3;
    return (lo | (hi << 4)) - 32;
}

// ---- unpack 8 consecutive values [8p, 8p+8) of one block ---------------------------------------
// `b` points at the block (shared memory); alignment is whatever BS guarantees (>= 2 bytes).
template <int TYPE> struct Unpack8;

template <> struct Unpack8<T_Q4_0> {       // y = (q - 8) * d                    ggml-quants.hpp:11-20
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const float d = ggb_h2f(ld_u16(b));
        const uint8_t *src = b + 2 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        const uint32_t w0 = ld_u32_a2(src) >> sh, w1 = ld_u32_a2(src + 4) >> sh;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            y[i] = __fmul_rn((float)((int)((w0 >> (8 * i)) & 0xf) - 8), d);
            y[i + 4] = __fmul_rn((float)((int)((w1 >> (8 * i)) & 0xf) - 8), d);
        }
    }
};

template <> struct Unpack8<T_Q4_1> {       // y = q * d + m  (two roundings)     ggml-quants.hpp:31-41
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const float d = ggb_h2f(ld_u16(b)), m = ggb_h2f(ld_u16(b + 2));
        const uint8_t *src = b + 4 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        const uint32_t w0 = ld_u32_a4(src) >> sh, w1 = ld_u32_a4(src + 4) >> sh;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            y[i] = __fadd_rn(__fmul_rn((float)((w0 >> (8 * i)) & 0xf), d), m);
            y[i + 4] = __fadd_rn(__fmul_rn((float)((w1 >> (8 * i)) & 0xf), d), m);
        }
    }
};

template <> struct Unpack8<T_Q5_0> {       // y = (q5 - 16) * d                  ggml-quants.hpp:52-67
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const float d = ggb_h2f(ld_u16(b));
        const uint32_t hb = (ld_u32_a2(b + 2) >> (8 * p)) & 0xff;     // 5th bit of element e is bit e of qh
        const uint8_t *src = b + 6 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        const uint32_t w0 = ld_u32_a2(src) >> sh, w1 = ld_u32_a2(src + 4) >> sh;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q0 = (int)(((w0 >> (8 * i)) & 0xf) | (((hb >> i) & 1) << 4)) - 16;
            const int q1 = (int)(((w1 >> (8 * i)) & 0xf) | (((hb >> (i + 4)) & 1) << 4)) - 16;
            y[i] = __fmul_rn((float)q0, d);
            y[i + 4] = __fmul_rn((float)q1, d);
        }
    }
};

template <> struct Unpack8<T_Q5_1> {       // y = q5 * d + m                     ggml-quants.hpp:78-94
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const float d = ggb_h2f(ld_u16(b)), m = ggb_h2f(ld_u16(b + 2));
        const uint32_t hb = (ld_u32_a4(b + 4) >> (8 * p)) & 0xff;
        const uint8_t *src = b + 8 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        const uint32_t w0 = ld_u32_a4(src) >> sh, w1 = ld_u32_a4(src + 4) >> sh;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t q0 = ((w0 >> (8 * i)) & 0xf) | (((hb >> i) & 1) << 4);
            const uint32_t q1 = ((w1 >> (8 * i)) & 0xf) | (((hb >> (i + 4)) & 1) << 4);
            y[i] = __fadd_rn(__fmul_rn((float)q0, d), m);
            y[i + 4] = __fadd_rn(__fmul_rn((float)q1, d), m);
        }
    }
};

template <> struct Unpack8<T_Q8_0> {       // y = q * d                          ggml-quants.hpp:105-110
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const float d = ggb_h2f(ld_u16(b));
        const uint8_t *src = b + 2 + 8 * p;
        const uint32_t w0 = ld_u32_a2(src), w1 = ld_u32_a2(src + 4);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            y[i] = __fmul_rn((float)(int)(int8_t)(w0 >> (8 * i)), d);
            y[i + 4] = __fmul_rn((float)(int)(int8_t)(w1 >> (8 * i)), d);
        }
    }
};

template <> struct Unpack8<T_Q2_K> {       // y = (d*sc)*q - dmin*m              gguf-py quants.py:406-428
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const int n = p >> 4, s = (p >> 2) & 3, l0 = (p & 3) * 8;
        const int is = 8 * n + 2 * s + (l0 >> 4);
        const float d = ggb_h2f(ld_u16(b + 80)), dmin = ggb_h2f(ld_u16(b + 82));
        const uint32_t scb = b[is];
        const float dl = __fmul_rn(d, (float)(scb & 0xf));
        const float ml = __fmul_rn(dmin, (float)(scb >> 4));
        const uint8_t *src = b + 16 + 32 * n + l0;
        const uint32_t w0 = ld_u32_a4(src) >> (2 * s), w1 = ld_u32_a4(src + 4) >> (2 * s);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            y[i] = __fsub_rn(__fmul_rn(dl, (float)((w0 >> (8 * i)) & 3)), ml);
            y[i + 4] = __fsub_rn(__fmul_rn(dl, (float)((w1 >> (8 * i)) & 3)), ml);
        }
    }
};

template <> struct Unpack8<T_Q3_K> {       // y = (d*sc6) * (q2 - (hbit ? 0 : 4)) gguf-py quants.py:433-472
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const int n = p >> 4, j = (p >> 2) & 3, l0 = (p & 3) * 8;
        const int is = 8 * n + 2 * j + (l0 >> 4);
        const float d = ggb_h2f(ld_u16(b + 108));
        const float dl = __fmul_rn(d, (float)q3k_scale(b + 96, is));
        const uint8_t *src = b + 32 + 32 * n + l0;
        const uint32_t w0 = ld_u32_a2(src) >> (2 * j), w1 = ld_u32_a2(src + 4) >> (2 * j);
        const int hs = 4 * n + j;
        const uint32_t h0 = ld_u32_a2(b + l0) >> hs, h1 = ld_u32_a2(b + l0 + 4) >> hs;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q0 = (int)((w0 >> (8 * i)) & 3) - (((h0 >> (8 * i)) & 1) ? 0 : 4);
            const int q1 = (int)((w1 >> (8 * i)) & 3) - (((h1 >> (8 * i)) & 1) ? 0 : 4);
            y[i] = __fmul_rn(dl, (float)q0);
            y[i + 4] = __fmul_rn(dl, (float)q1);
        }
    }
};

template <> struct Unpack8<T_Q4_K> {       // y = (d*sc)*q - dmin*m              gguf-py quants.py:504-522
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const int il = p >> 3, hi = (p >> 2) & 1, l0 = (p & 3) * 8;
        int sc, mn;
        k4_scale_min(2 * il + hi, b + 4, sc, mn);
        const float d = ggb_h2f(ld_u16(b)), dmin = ggb_h2f(ld_u16(b + 2));
        const float dl = __fmul_rn(d, (float)sc);
        const float ml = __fmul_rn(dmin, (float)mn);
        const uint2 w = *reinterpret_cast<const uint2 *>(b + 16 + 32 * il + l0);      // 8-byte aligned
        const uint32_t w0 = w.x >> (4 * hi), w1 = w.y >> (4 * hi);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            y[i] = __fsub_rn(__fmul_rn(dl, (float)((w0 >> (8 * i)) & 0xf)), ml);
            y[i + 4] = __fsub_rn(__fmul_rn(dl, (float)((w1 >> (8 * i)) & 0xf)), ml);
        }
    }
};

template <> struct Unpack8<T_Q5_K> {       // as Q4_K with q = ql | hbit << 4    gguf-py quants.py:527-549
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const int il = p >> 3, hi = (p >> 2) & 1, l0 = (p & 3) * 8;
        const int j = 2 * il + hi;
        int sc, mn;
        k4_scale_min(j, b + 4, sc, mn);
        const float d = ggb_h2f(ld_u16(b)), dmin = ggb_h2f(ld_u16(b + 2));
        const float dl = __fmul_rn(d, (float)sc);
        const float ml = __fmul_rn(dmin, (float)mn);
        const uint2 w = *reinterpret_cast<const uint2 *>(b + 48 + 32 * il + l0);
        const uint2 h = *reinterpret_cast<const uint2 *>(b + 16 + l0);
        const uint32_t w0 = w.x >> (4 * hi), w1 = w.y >> (4 * hi);
        const uint32_t h0 = h.x >> j, h1 = h.y >> j;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t q0 = ((w0 >> (8 * i)) & 0xf) | (((h0 >> (8 * i)) & 1) << 4);
            const uint32_t q1 = ((w1 >> (8 * i)) & 0xf) | (((h1 >> (8 * i)) & 1) << 4);
            y[i] = __fsub_rn(__fmul_rn(dl, (float)q0), ml);
            y[i + 4] = __fsub_rn(__fmul_rn(dl, (float)q1), ml);
        }
    }
};

template <> struct Unpack8<T_Q6_K> {       // y = (d*sc_i8) * (q6 - 32)          gguf-py quants.py:554-572
    __device__ static void run(const uint8_t *b, int p, float (&y)[8]) {
        const int ip = p >> 4, r = (p >> 2) & 3, l0 = (p & 3) * 8;
        const int is = 8 * ip + 2 * r + (l0 >> 4);
        const float d = ggb_h2f(ld_u16(b + 208));
        const float dl = __fmul_rn(d, (float)(int)(int8_t)b[192 + is]);
        const uint8_t *ql = b + 64 * ip + 32 * (r & 1) + l0;
        const uint8_t *qh = b + 128 + 32 * ip + l0;
        const int sl = 4 * (r >> 1), shh = 2 * r;
        const uint32_t w0 = ld_u32_a2(ql) >> sl, w1 = ld_u32_a2(ql + 4) >> sl;
        const uint32_t h0 = ld_u32_a2(qh) >> shh, h1 = ld_u32_a2(qh + 4) >> shh;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q0 = (int)(((w0 >> (8 * i)) & 0xf) | (((h0 >> (8 * i)) & 3) << 4)) - 32;
            const int q1 = (int)(((w1 >> (8 * i)) & 0xf) | (((h1 >> (8 * i)) & 3) << 4)) - 32;
            y[i] = __fmul_rn(dl, (float)q0);
            y[i + 4] = __fmul_rn(dl, (float)q1);
        }
    }
};

// ---- the tile kernel --------------------------------------------------------------------------
template <int TYPE, typename OutT>
__global__ void __launch_bounds__(kThreads)
dequant_tile_kernel(const uint8_t *__restrict__ W, OutT *__restrict__ out, int64_t nblocks, int w_aligned16) {
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

    if (nb == TILE_BLOCKS && w_aligned16) {
        // TMA path: one elected thread arms the mbarrier and issues one bulk copy for the tile.
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_expect_tx(&bar, TILE_BYTES);
            bulk_g2s_hint(tile, src, TILE_BYTES, &bar, l2_policy_evict_first());
        }
        __syncthreads();                                   // barrier object is initialised for everyone
        mbar_wait(&bar, 0);
    } else {
        // ragged last tile or a weight pointer that is not 16-byte aligned: cooperative 2-byte copy
        const int nbytes = nb * BS;
        for (int i = 2 * tid; i < nbytes; i += 2 * kThreads)
            *reinterpret_cast<uint16_t *>(tile + i) = *reinterpret_cast<const uint16_t *>(src + i);
        __syncthreads();
    }

    OutT *dst = out + b0 * QK;
    const int nitems = nb * ITEMS_PER_BLOCK;
#pragma unroll
    for (int it = tid; it < TILE_BLOCKS * ITEMS_PER_BLOCK; it += kThreads) {
        if (it < nitems) {
            const int blk = it / ITEMS_PER_BLOCK, p = it % ITEMS_PER_BLOCK;
            float y[8];
            Unpack8<TYPE>::run(tile + blk * BS, p, y);
            store8(dst + (int64_t)it * 8, y);
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
    dequant_tile_kernel<TYPE, OutT><<<(unsigned)ntiles, kThreads, 0, stream>>>(
        static_cast<const uint8_t *>(W), static_cast<OutT *>(out), nblocks, aligned);
    GGB_CUDA_TRY(cudaGetLastError());
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
        default:
            ggb_set_error("ggb_dequantize: unsupported ggml type %d", type);
            return GGB_E_TYPE;
    }
}
