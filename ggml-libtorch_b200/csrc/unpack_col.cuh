// unpack_col.cuh -- "column unit" dequantizers for the K-quants (dequant.cu).
//
// Unpack8 (unpack.cuh) produces 8 CONSECUTIVE outputs per call, so every packed byte is fetched once per
// bit-field it holds (twice for Q4_K/Q5_K, four times for Q2_K/Q3_K/Q6_K) and every value costs a shift, a
// mask, a PRMT, an FADD, an FMUL and an FSUB.  Here a unit is 8 consecutive BYTES of a block and ALL the values
// packed in them: 16 outputs (Q4_K, Q5_K) or 32 (Q2_K, Q3_K, Q6_K), written as two / four 128-bit stores.
//
// Arithmetic is still the reference's, bit for bit (gguf-py quants.py:404-572):
//     y = fl( fl(dl * q) - ml ),   dl = fl(d * sc), ml = fl(dmin * m)
// but fl(dl * q) comes out of ONE FFMA with no int -> float conversion and no shift:
//     x = as_float(0x4B000000 | field) = 2^23 + field,   field = q << s  (the bit-field where it sits in the byte)
//     fl(dl * q) = fma(a, x, c),   a = dl / 2^s,  c = -(a * 2^23)          (both exact: powers of two)
// a*x + c is exactly dl*q in real arithmetic, so the single rounding of the FMA is the rounding of the product.
// The one case that differs is the SIGN of an exact zero (fma gives +0 where fl(-|dl| * 0) is -0), so a tile
// whose table holds a negative, -0 or non-finite dl takes the Unpack8 path instead (no real model has one:
// Q2_K/Q4_K/Q5_K super-block scales are max/63-style positive numbers).
// Q3_K and Q6_K have signed quants and no minima: x - (2^23 + offset) is exact, then one FMUL (as Unpack8 does).
#pragma once

#include "unpack.cuh"

template <int J>
__device__ __forceinline__ float byte_magic(uint32_t w) {              // 2^23 + (byte J of w)
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + J));
}
// 4 outputs from the 4 byte fields of w: fl(fl(a * field) - ml) with e = {a, c, ml, -}
__device__ __forceinline__ void col_word_fma(uint32_t w, const float4 &e, float *y) {
    y[0] = __fsub_rn(__fmaf_rn(e.x, byte_magic<0>(w), e.y), e.z);
    y[1] = __fsub_rn(__fmaf_rn(e.x, byte_magic<1>(w), e.y), e.z);
    y[2] = __fsub_rn(__fmaf_rn(e.x, byte_magic<2>(w), e.y), e.z);
    y[3] = __fsub_rn(__fmaf_rn(e.x, byte_magic<3>(w), e.y), e.z);
}
// 4 outputs fl(dl * (byte - OFF))
template <int OFF>
__device__ __forceinline__ void col_word_mul(uint32_t w, float dl, float *y) {
    constexpr float kMagic = 8388608.0f + (float)OFF;
    y[0] = __fmul_rn(dl, __fsub_rn(byte_magic<0>(w), kMagic));
    y[1] = __fmul_rn(dl, __fsub_rn(byte_magic<1>(w), kMagic));
    y[2] = __fmul_rn(dl, __fsub_rn(byte_magic<2>(w), kMagic));
    y[3] = __fmul_rn(dl, __fsub_rn(byte_magic<3>(w), kMagic));
}
// two consecutive 32-bit words at a 2-byte-aligned shared address: aligned loads + funnel shifts
__device__ __forceinline__ uint2 lds_u64_a2(const uint8_t *a) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(a);
    const uint32_t *p = reinterpret_cast<const uint32_t *>(u & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(u & 2) * 8;
    const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

// ColUnpack<TYPE>:
//   NSUB      table entries per block (= Unpack8<TYPE>::NSUB)
//   UNITS     units per block
//   entry(b, j, bad)   table entry of sub-block j; sets bad when the tile must take the Unpack8 path
//   run(b, u, tab, dst)  unit u of block b: tab = the block's NSUB entries, dst = the block's 256 outputs
template <int TYPE> struct ColUnpack;

__device__ __forceinline__ float4 col_entry_fma(float2 s, float scale, bool &bad) {
    const float a = s.x * scale;                                          // exact (power of two)
    bad = bad || (__float_as_uint(s.x) >> 31) || !isfinite(s.x) || !isfinite(s.y);
    return make_float4(a, -(a * 8388608.0f), s.y, 0.0f);
}

template <> struct ColUnpack<T_Q4_K> {
    static constexpr int NSUB = 8, UNITS = 16;
    __device__ __forceinline__ static float4 entry(const uint8_t *b, int j, bool &bad) {
        return col_entry_fma(Unpack8<T_Q4_K>::sub_scale(b, j), (j & 1) ? 0.0625f : 1.0f, bad);   // odd groups: high nibbles = 16 q
    }
    template <typename OutT>
    __device__ __forceinline__ static void run(const uint8_t *b, int u, const float4 *tab, OutT *dst) {
        const int il = u >> 2, o = u & 3;
        const uint2 w = *reinterpret_cast<const uint2 *>(b + 16 + 32 * il + 8 * o);
        const float4 e0 = tab[2 * il], e1 = tab[2 * il + 1];
        float y[8];
        col_word_fma(w.x & 0x0f0f0f0fu, e0, y);
        col_word_fma(w.y & 0x0f0f0f0fu, e0, y + 4);
        store8(dst + 64 * il + 8 * o, y);
        col_word_fma(w.x & 0xf0f0f0f0u, e1, y);
        col_word_fma(w.y & 0xf0f0f0f0u, e1, y + 4);
        store8(dst + 64 * il + 32 + 8 * o, y);
    }
};

template <> struct ColUnpack<T_Q5_K> {
    static constexpr int NSUB = 8, UNITS = 16;
    __device__ __forceinline__ static float4 entry(const uint8_t *b, int j, bool &bad) {
        return col_entry_fma(Unpack8<T_Q5_K>::sub_scale(b, j), 1.0f, bad);
    }
    template <typename OutT>
    __device__ __forceinline__ static void run(const uint8_t *b, int u, const float4 *tab, OutT *dst) {
        const int il = u >> 2, o = u & 3;
        const uint2 w = *reinterpret_cast<const uint2 *>(b + 48 + 32 * il + 8 * o);
        const uint2 h = *reinterpret_cast<const uint2 *>(b + 16 + 8 * o);
        const float4 e0 = tab[2 * il], e1 = tab[2 * il + 1];
        const uint32_t hx = h.x >> (2 * il), hy = h.y >> (2 * il);         // bit 0: group 2il, bit 1: group 2il+1
        float y[8];
        col_word_fma((w.x & 0x0f0f0f0fu) | ((hx << 4) & 0x10101010u), e0, y);
        col_word_fma((w.y & 0x0f0f0f0fu) | ((hy << 4) & 0x10101010u), e0, y + 4);
        store8(dst + 64 * il + 8 * o, y);
        col_word_fma(((w.x >> 4) & 0x0f0f0f0fu) | ((hx << 3) & 0x10101010u), e1, y);
        col_word_fma(((w.y >> 4) & 0x0f0f0f0fu) | ((hy << 3) & 0x10101010u), e1, y + 4);
        store8(dst + 64 * il + 32 + 8 * o, y);
    }
};

template <> struct ColUnpack<T_Q2_K> {       // scales[16] @0, qs[64] @16, d @80, dmin @82
    static constexpr int NSUB = 16, UNITS = 8;
    __device__ __forceinline__ static float4 entry(const uint8_t *b, int is, bool &bad) {
        const int j = (is >> 1) & 3;                                       // the field sits at bits 2j..2j+1: 4^j * q
        const float scale = j == 0 ? 1.0f : j == 1 ? 0.25f : j == 2 ? 0.0625f : 0.015625f;
        return col_entry_fma(Unpack8<T_Q2_K>::sub_scale(b, is), scale, bad);
    }
    template <typename OutT>
    __device__ __forceinline__ static void run(const uint8_t *b, int u, const float4 *tab, OutT *dst) {
        const int n = u >> 2, o = u & 3;
        const uint8_t *src = b + 16 + 32 * n + 8 * o;                      // 4-byte aligned (84 = 4 * 21)
        const uint32_t w0 = ld_u32_a4(src), w1 = ld_u32_a4(src + 4);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float4 e = tab[8 * n + 2 * j + (o >> 1)];
            const uint32_t mask = 0x03030303u << (2 * j);
            float y[8];
            col_word_fma(w0 & mask, e, y);
            col_word_fma(w1 & mask, e, y + 4);
            store8(dst + 128 * n + 32 * j + 8 * o, y);
        }
    }
};

template <> struct ColUnpack<T_Q3_K> {       // hmask[32] @0, qs[64] @32, scales[12] @96, d @108
    static constexpr int NSUB = 16, UNITS = 8;
    __device__ __forceinline__ static float4 entry(const uint8_t *b, int is, bool &) {
        return make_float4(Unpack8<T_Q3_K>::sub_scale(b, is).x, 0.0f, 0.0f, 0.0f);
    }
    template <typename OutT>
    __device__ __forceinline__ static void run(const uint8_t *b, int u, const float4 *tab, OutT *dst) {
        const int n = u >> 2, o = u & 3;
        const uint2 qs = lds_u64_a2(b + 32 + 32 * n + 8 * o);
        uint2 hm = lds_u64_a2(b + 8 * o);
        hm.x >>= 4 * n;                                                    // bit j of every byte: group 4n + j
        hm.y >>= 4 * n;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float dl = tab[8 * n + 2 * j + (o >> 1)].x;
            const uint32_t hx = (j < 2) ? (hm.x << (2 - j)) : (hm.x >> (j - 2));
            const uint32_t hy = (j < 2) ? (hm.y << (2 - j)) : (hm.y >> (j - 2));
            float y[8];                                                    // raw = q2 + 4 * hbit, value = raw - 4
            col_word_mul<4>(((qs.x >> (2 * j)) & 0x03030303u) | (hx & 0x04040404u), dl, y);
            col_word_mul<4>(((qs.y >> (2 * j)) & 0x03030303u) | (hy & 0x04040404u), dl, y + 4);
            store8(dst + 128 * n + 32 * j + 8 * o, y);
        }
    }
};

template <> struct ColUnpack<T_Q6_K> {       // ql[128] @0, qh[64] @128, scales[16] @192, d @208
    static constexpr int NSUB = 16, UNITS = 8;
    __device__ __forceinline__ static float4 entry(const uint8_t *b, int is, bool &) {
        return make_float4(Unpack8<T_Q6_K>::sub_scale(b, is).x, 0.0f, 0.0f, 0.0f);
    }
    template <typename OutT>
    __device__ __forceinline__ static void run(const uint8_t *b, int u, const float4 *tab, OutT *dst) {
        const int ip = u >> 2, o = u & 3;
        const uint2 qa = lds_u64_a2(b + 64 * ip + 8 * o), qb = lds_u64_a2(b + 64 * ip + 32 + 8 * o);
        const uint2 qh = lds_u64_a2(b + 128 + 32 * ip + 8 * o);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float dl = tab[8 * ip + 2 * r + (o >> 1)].x;
            const uint2 ql = (r & 1) ? qb : qa;
            const uint32_t lx = (r < 2) ? (ql.x & 0x0f0f0f0fu) : ((ql.x >> 4) & 0x0f0f0f0fu);
            const uint32_t ly = (r < 2) ? (ql.y & 0x0f0f0f0fu) : ((ql.y >> 4) & 0x0f0f0f0fu);
            const uint32_t hx = (r == 0) ? (qh.x << 4) : (r == 1) ? (qh.x << 2) : (r == 2) ? qh.x : (qh.x >> 2);
            const uint32_t hy = (r == 0) ? (qh.y << 4) : (r == 1) ? (qh.y << 2) : (r == 2) ? qh.y : (qh.y >> 2);
            float y[8];                                                    // raw = value + 32
            col_word_mul<32>(lx | (hx & 0x30303030u), dl, y);
            col_word_mul<32>(ly | (hy & 0x30303030u), dl, y + 4);
            store8(dst + 128 * ip + 32 * r + 8 * o, y);
        }
    }
};
