// unpack.cuh -- per-type GGUF block unpackers shared by the dequantizers (dequant.cu) and the MMQ
// weight-tile producer (mmq.cu).
//
// For every type:
//   Unpack8<TYPE>::NSUB            (scale, min) pairs per block (sub-blocks; 1 for the legacy types)
//   Unpack8<TYPE>::sub_of(p)       which pair the 8 values [8p, 8p+8) of a block use
//   Unpack8<TYPE>::sub_scale(b, j) pair j as fp32 (dl, ml) = (d*sc_j, dmin*m_j)   [legacy: (d, m)]
//   Unpack8<TYPE>::run_sc(b, p, s, y)   the 8 values as fp32 given their pair s
//   Unpack8<TYPE>::run(b, p, y)         = run_sc(b, p, sub_scale(b, sub_of(p)), y)
// Arithmetic is the reference CPU op's (ggml-cpu/ggml-quants.hpp:4-112 for the legacy types, gguf-py
// quants.py:404-572 for the K-quants, its IQ4_NL / IQ4_XS classes for those two): every product and sum is a separately
// rounded fp32 operation (__fmul_rn / __fadd_rn / __fsub_rn are never contracted into FMA).
// Integer quants become fp32 WITHOUT the quarter-rate I2F pipe: PRMT builds the float 2^23 + byte and
// one exact FADD subtracts 2^23 + offset (both are integers below 2^24, so the difference is exact).
// `b` may live in shared or global memory; its alignment is what the block size guarantees.
#pragma once

#include "ggb_common.cuh"

// ---- exact byte -> fp32 -------------------------------------------------------------------------
// q[i] = (float)(byte i of w) - OFF, exactly
template <int OFF>
__device__ __forceinline__ void bytes_to_f32(uint32_t w, float *q) {
    constexpr float kMagic = 8388608.0f + (float)OFF;
    q[0] = __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540)), kMagic);
    q[1] = __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7541)), kMagic);
    q[2] = __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7542)), kMagic);
    q[3] = __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7543)), kMagic);
}
__device__ __forceinline__ uint32_t spread4_bit4(uint32_t x) {       // bit i of x (i < 4) -> bit 4 of byte i
    return ((x * 0x00204081u) & 0x01010101u) << 4;
}

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

template <int TYPE> struct Unpack8;

#define GGB_UNPACK_RUN                                                                  \
    __device__ __forceinline__ static void run(const uint8_t *b, int p, float (&y)[8]) { \
        run_sc(b, p, sub_scale(b, sub_of(p)), y);                                       \
    }

template <> struct Unpack8<T_Q4_0> {       // y = (q - 8) * d                    ggml-quants.hpp:11-20
    static constexpr int NSUB = 1;
    __device__ __forceinline__ static int sub_of(int) { return 0; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int) { return make_float2(ggb_h2f(ld_u16(b)), 0.0f); }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const uint8_t *src = b + 2 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        float q[8];
        bytes_to_f32<8>((ld_u32_a2(src) >> sh) & 0x0f0f0f0fu, q);
        bytes_to_f32<8>((ld_u32_a2(src + 4) >> sh) & 0x0f0f0f0fu, q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fmul_rn(q[i], s.x);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q4_1> {       // y = q * d + m  (two roundings)     ggml-quants.hpp:31-41
    static constexpr int NSUB = 1;
    __device__ __forceinline__ static int sub_of(int) { return 0; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int) {
        return make_float2(ggb_h2f(ld_u16(b)), ggb_h2f(ld_u16(b + 2)));
    }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const uint8_t *src = b + 4 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        float q[8];
        bytes_to_f32<0>((ld_u32_a4(src) >> sh) & 0x0f0f0f0fu, q);
        bytes_to_f32<0>((ld_u32_a4(src + 4) >> sh) & 0x0f0f0f0fu, q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fadd_rn(__fmul_rn(q[i], s.x), s.y);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q5_0> {       // y = (q5 - 16) * d                  ggml-quants.hpp:52-67
    static constexpr int NSUB = 1;
    __device__ __forceinline__ static int sub_of(int) { return 0; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int) { return make_float2(ggb_h2f(ld_u16(b)), 0.0f); }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const uint32_t hb = (ld_u32_a2(b + 2) >> (8 * p)) & 0xff;     // 5th bit of element e is bit e of qh
        const uint8_t *src = b + 6 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        float q[8];
        bytes_to_f32<16>(((ld_u32_a2(src) >> sh) & 0x0f0f0f0fu) | spread4_bit4(hb & 0xf), q);
        bytes_to_f32<16>(((ld_u32_a2(src + 4) >> sh) & 0x0f0f0f0fu) | spread4_bit4(hb >> 4), q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fmul_rn(q[i], s.x);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q5_1> {       // y = q5 * d + m                     ggml-quants.hpp:78-94
    static constexpr int NSUB = 1;
    __device__ __forceinline__ static int sub_of(int) { return 0; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int) {
        return make_float2(ggb_h2f(ld_u16(b)), ggb_h2f(ld_u16(b + 2)));
    }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const uint32_t hb = (ld_u32_a4(b + 4) >> (8 * p)) & 0xff;
        const uint8_t *src = b + 8 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        float q[8];
        bytes_to_f32<0>(((ld_u32_a4(src) >> sh) & 0x0f0f0f0fu) | spread4_bit4(hb & 0xf), q);
        bytes_to_f32<0>(((ld_u32_a4(src + 4) >> sh) & 0x0f0f0f0fu) | spread4_bit4(hb >> 4), q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fadd_rn(__fmul_rn(q[i], s.x), s.y);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q8_0> {       // y = q * d                          ggml-quants.hpp:105-110
    static constexpr int NSUB = 1;
    __device__ __forceinline__ static int sub_of(int) { return 0; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int) { return make_float2(ggb_h2f(ld_u16(b)), 0.0f); }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const uint8_t *src = b + 2 + 8 * p;
        float q[8];
        bytes_to_f32<128>(ld_u32_a2(src) ^ 0x80808080u, q);           // int8 + 128 as an unsigned byte
        bytes_to_f32<128>(ld_u32_a2(src + 4) ^ 0x80808080u, q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fmul_rn(q[i], s.x);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q2_K> {       // y = (d*sc)*q - dmin*m              gguf-py quants.py:406-428
    static constexpr int NSUB = 16;
    __device__ __forceinline__ static int sub_of(int p) { return p >> 1; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int is) {
        const uint32_t scb = b[is];
        return make_float2(__fmul_rn(ggb_h2f(ld_u16(b + 80)), (float)(scb & 0xf)),
                           __fmul_rn(ggb_h2f(ld_u16(b + 82)), (float)(scb >> 4)));
    }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const int n = p >> 4, sft = (p >> 2) & 3, l0 = (p & 3) * 8;
        const uint8_t *src = b + 16 + 32 * n + l0;
        float q[8];
        bytes_to_f32<0>((ld_u32_a4(src) >> (2 * sft)) & 0x03030303u, q);
        bytes_to_f32<0>((ld_u32_a4(src + 4) >> (2 * sft)) & 0x03030303u, q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fsub_rn(__fmul_rn(s.x, q[i]), s.y);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q3_K> {       // y = (d*sc6) * (q2 - (hbit ? 0 : 4)) gguf-py quants.py:433-472
    static constexpr int NSUB = 16;
    __device__ __forceinline__ static int sub_of(int p) { return p >> 1; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int is) {
        return make_float2(__fmul_rn(ggb_h2f(ld_u16(b + 108)), (float)q3k_scale(b + 96, is)), 0.0f);
    }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const int n = p >> 4, j = (p >> 2) & 3, l0 = (p & 3) * 8;
        const uint8_t *src = b + 32 + 32 * n + l0;
        const int hs = 4 * n + j;
        const uint32_t lo0 = (ld_u32_a2(src) >> (2 * j)) & 0x03030303u, lo1 = (ld_u32_a2(src + 4) >> (2 * j)) & 0x03030303u;
        const uint32_t h0 = (ld_u32_a2(b + l0) >> hs) & 0x01010101u, h1 = (ld_u32_a2(b + l0 + 4) >> hs) & 0x01010101u;
        float q[8];
        bytes_to_f32<4>(lo0 | (h0 << 2), q);                          // raw = q2 + 4*hbit, value = raw - 4
        bytes_to_f32<4>(lo1 | (h1 << 2), q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fmul_rn(s.x, q[i]);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q4_K> {       // y = (d*sc)*q - dmin*m              gguf-py quants.py:504-522
    static constexpr int NSUB = 8;
    __device__ __forceinline__ static int sub_of(int p) { return p >> 2; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int j) {
        int sc, mn;
        k4_scale_min(j, b + 4, sc, mn);
        return make_float2(__fmul_rn(ggb_h2f(ld_u16(b)), (float)sc), __fmul_rn(ggb_h2f(ld_u16(b + 2)), (float)mn));
    }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const int il = p >> 3, hi = (p >> 2) & 1, l0 = (p & 3) * 8;
        const uint2 w = *reinterpret_cast<const uint2 *>(b + 16 + 32 * il + l0);      // 8-byte aligned
        float q[8];
        bytes_to_f32<0>((w.x >> (4 * hi)) & 0x0f0f0f0fu, q);
        bytes_to_f32<0>((w.y >> (4 * hi)) & 0x0f0f0f0fu, q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fsub_rn(__fmul_rn(s.x, q[i]), s.y);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q5_K> {       // as Q4_K with q = ql | hbit << 4    gguf-py quants.py:527-549
    static constexpr int NSUB = 8;
    __device__ __forceinline__ static int sub_of(int p) { return p >> 2; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int j) { return Unpack8<T_Q4_K>::sub_scale(b, j); }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const int il = p >> 3, hi = (p >> 2) & 1, l0 = (p & 3) * 8;
        const int j = 2 * il + hi;
        const uint2 w = *reinterpret_cast<const uint2 *>(b + 48 + 32 * il + l0);
        const uint2 h = *reinterpret_cast<const uint2 *>(b + 16 + l0);
        float q[8];
        bytes_to_f32<0>(((w.x >> (4 * hi)) & 0x0f0f0f0fu) | (((h.x >> j) & 0x01010101u) << 4), q);
        bytes_to_f32<0>(((w.y >> (4 * hi)) & 0x0f0f0f0fu) | (((h.y >> j) & 0x01010101u) << 4), q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fsub_rn(__fmul_rn(s.x, q[i]), s.y);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_Q6_K> {       // y = (d*sc_i8) * (q6 - 32)          gguf-py quants.py:554-572
    static constexpr int NSUB = 16;
    __device__ __forceinline__ static int sub_of(int p) { return p >> 1; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int is) {
        return make_float2(__fmul_rn(ggb_h2f(ld_u16(b + 208)), (float)(int)(int8_t)b[192 + is]), 0.0f);
    }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const int ip = p >> 4, r = (p >> 2) & 3, l0 = (p & 3) * 8;
        const uint8_t *ql = b + 64 * ip + 32 * (r & 1) + l0;
        const uint8_t *qh = b + 128 + 32 * ip + l0;
        const int sl = 4 * (r >> 1), shh = 2 * r;
        const uint32_t lo0 = (ld_u32_a2(ql) >> sl) & 0x0f0f0f0fu, lo1 = (ld_u32_a2(ql + 4) >> sl) & 0x0f0f0f0fu;
        const uint32_t h0 = (ld_u32_a2(qh) >> shh) & 0x03030303u, h1 = (ld_u32_a2(qh + 4) >> shh) & 0x03030303u;
        float q[8];
        bytes_to_f32<32>(lo0 | (h0 << 4), q);
        bytes_to_f32<32>(lo1 | (h1 << 4), q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fmul_rn(s.x, q[i]);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_IQ4_NL> {     // y = d * kvalues[q]                 gguf-py quants.py IQ4_NL; dequantize.cuh:400-416
    static constexpr int NSUB = 1;
    __device__ __forceinline__ static int sub_of(int) { return 0; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int) { return make_float2(ggb_h2f(ld_u16(b)), 0.0f); }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const uint8_t *src = b + 2 + 8 * (p & 1);
        const int sh = 4 * (p >> 1);
        float q[8];
        bytes_to_f32<128>(iq4nl_lut4((ld_u32_a2(src) >> sh) & 0x0f0f0f0fu) ^ 0x80808080u, q);
        bytes_to_f32<128>(iq4nl_lut4((ld_u32_a2(src + 4) >> sh) & 0x0f0f0f0fu) ^ 0x80808080u, q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fmul_rn(s.x, q[i]);
    }
    GGB_UNPACK_RUN
};

template <> struct Unpack8<T_IQ4_XS> {     // y = (d*(ls-32)) * kvalues[q]       gguf-py quants.py IQ4_XS; dequantize.cuh:419-433
    static constexpr int NSUB = 8;
    __device__ __forceinline__ static int sub_of(int p) { return p >> 2; }
    __device__ __forceinline__ static float2 sub_scale(const uint8_t *b, int ib) {
        return make_float2(__fmul_rn(ggb_h2f(ld_u16(b)), (float)iq4xs_scale(b, ib)), 0.0f);
    }
    __device__ __forceinline__ static void run_sc(const uint8_t *b, int p, float2 s, float (&y)[8]) {
        const uint8_t *src = b + 8 + 16 * (p >> 2) + 8 * (p & 1);
        const int sh = 4 * ((p >> 1) & 1);
        float q[8];
        bytes_to_f32<128>(iq4nl_lut4((ld_u32_a4(src) >> sh) & 0x0f0f0f0fu) ^ 0x80808080u, q);
        bytes_to_f32<128>(iq4nl_lut4((ld_u32_a4(src + 4) >> sh) & 0x0f0f0f0fu) ^ 0x80808080u, q + 4);
#pragma unroll
        for (int i = 0; i < 8; i++) y[i] = __fmul_rn(s.x, q[i]);
    }
    GGB_UNPACK_RUN
};

#undef GGB_UNPACK_RUN
