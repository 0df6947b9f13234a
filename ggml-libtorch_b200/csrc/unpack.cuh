// unpack.cuh -- per-type GGUF block unpackers shared by the dequantizers (dequant.cu) and the MMQ
// weight-tile producer (mmq.cu).
//
// Unpack8<TYPE>::run(block, p, y) writes the 8 consecutive values [8p, 8p+8) of one block as fp32,
// in the reference CPU op's arithmetic (ggml-cpu/ggml-quants.hpp:4-112 for the legacy types, gguf-py
// quants.py:404-572 for the K-quants): every product and sum is a separately
// rounded fp32 operation (__fmul_rn / __fadd_rn / __fsub_rn are never contracted into FMA).
// `block` may live in shared or global memory; its alignment is what the block size guarantees.
#pragma once

#include "ggb_common.cuh"

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

