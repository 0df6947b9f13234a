// mmvq_group_ops.cuh -- per-type unpack of one 32-value group from raw GGUF bytes in shared memory and the reference's
// per-type combine algebra; shared by the generic MMVQ kernel (mmvq.cu) and the register-resident batch-1 kernel
// (mmvq_rg.cu).  Restates vec_dot_*_q8_1 of hf-kernels/ggml-kernels/ggml/vecdotq.cuh (line ranges at `combine`).
#pragma once

#include "q8_1.cuh"
#include "iq_grid.cuh"

namespace ggb_group {

template <int TYPE> struct VqTraits;
//   NEED_SQ : aux = (int16 sq1 << 16 | int16 sq0), the sums of the q8 quants over each 16-value half
//   NEED_S8 : aux = bits of the fp16-rounded float sum of the raw inputs (block_q8_1.s)
template <> struct VqTraits<T_Q4_0> { static constexpr bool NEED_SQ = false, NEED_S8 = true;  };
template <> struct VqTraits<T_Q4_1> { static constexpr bool NEED_SQ = false, NEED_S8 = true;  };
template <> struct VqTraits<T_Q5_0> { static constexpr bool NEED_SQ = false, NEED_S8 = true;  };
template <> struct VqTraits<T_Q5_1> { static constexpr bool NEED_SQ = false, NEED_S8 = true;  };
template <> struct VqTraits<T_Q8_0> { static constexpr bool NEED_SQ = false, NEED_S8 = false; };
template <> struct VqTraits<T_Q2_K> { static constexpr bool NEED_SQ = true,  NEED_S8 = false; };
template <> struct VqTraits<T_Q3_K> { static constexpr bool NEED_SQ = true,  NEED_S8 = false; };
template <> struct VqTraits<T_Q4_K> { static constexpr bool NEED_SQ = true,  NEED_S8 = false; };
template <> struct VqTraits<T_Q5_K> { static constexpr bool NEED_SQ = true,  NEED_S8 = false; };
template <> struct VqTraits<T_Q6_K> { static constexpr bool NEED_SQ = true,  NEED_S8 = false; };
template <> struct VqTraits<T_IQ4_NL> { static constexpr bool NEED_SQ = false, NEED_S8 = false; };
template <> struct VqTraits<T_IQ4_XS> { static constexpr bool NEED_SQ = false, NEED_S8 = false; };
template <> struct VqTraits<T_IQ2_XXS> { static constexpr bool NEED_SQ = false, NEED_S8 = false; };
template <> struct VqTraits<T_IQ2_XS>  { static constexpr bool NEED_SQ = false, NEED_S8 = false; };
template <> struct VqTraits<T_IQ2_S>   { static constexpr bool NEED_SQ = false, NEED_S8 = false; };
template <> struct VqTraits<T_IQ3_XXS> { static constexpr bool NEED_SQ = false, NEED_S8 = false; };
template <> struct VqTraits<T_IQ3_S>   { static constexpr bool NEED_SQ = false, NEED_S8 = false; };
template <> struct VqTraits<T_IQ1_S>   { static constexpr bool NEED_SQ = false, NEED_S8 = true;  };   // the shift goes through q8_1's s field
template <> struct VqTraits<T_IQ1_M>   { static constexpr bool NEED_SQ = false, NEED_S8 = false; };   // the shift is folded into the int8 weights

struct GroupW {
    uint32_t w[8];      // the 32 raw quants of the group, 4 per word, element order
    float a0, a1;       // scale of elements 0..15 / 16..31 (super-block d already applied)
    float m0, m1;       // minima (dmin applied) or, for Q4_1/Q5_1, m0 = block minimum
};

__device__ __forceinline__ uint32_t spread4(uint32_t x) {       // bit i of x -> bit 4 of byte i
    return ((x * 0x00204081u) & 0x01010101u) << 4;
}
__device__ __forceinline__ void k4_scale_min(int j, const uint8_t *s, int &sc, int &mn) {   // dequantize.cuh:154-161
    if (j < 4) {
        sc = s[j] & 63;
        mn = s[j + 4] & 63;
    } else {
        sc = (s[j + 4] & 0x0f) | ((s[j - 4] >> 6) << 4);
        mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4);
    }
}
__device__ __forceinline__ int q3k_scale(const uint8_t *s, int is) {
    const int lo = (is < 8) ? (s[is] & 0x0f) : (s[is - 8] >> 4);
    const int hi = (s[8 + (is & 3)] >> (2 * (is >> 2))) & 3;
    return (lo | (hi << 4)) - 32;
}

// NWORDS consecutive 32-bit words from a 2-byte-aligned shared address: NWORDS + 1 aligned loads and funnel shifts
// instead of 2 * NWORDS 16-bit loads (the 16-bit loads of the 18/22/34-byte blocks kept the LSU queue full: 45 % of all
// stall samples of the Q8_0 kernel were on them).  May read up to 4 bytes past the last word: still inside the stage ring.
template <int NWORDS>
__device__ __forceinline__ void lds_words_a2(const uint8_t *a, uint32_t (&out)[NWORDS]) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(a);
    const uint32_t *p = reinterpret_cast<const uint32_t *>(u & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(u & 2) * 8;
    uint32_t w[NWORDS + 1];
#pragma unroll
    for (int i = 0; i <= NWORDS; i++) w[i] = p[i];
#pragma unroll
    for (int i = 0; i < NWORDS; i++) out[i] = __funnelshift_r(w[i], w[i + 1], sh);
}

// Unpack group `g` (32 values) of a staged row segment that starts on a block boundary.
template <int TYPE>
__device__ __forceinline__ void load_group(const uint8_t *seg, int g, GroupW &o) {
    constexpr int BS = BlockTraits<TYPE>::BS;
    o.m0 = o.m1 = 0.0f;
    if constexpr (TYPE == T_Q4_0) {
        const uint8_t *b = seg + g * BS;
        o.a0 = o.a1 = ggb_h2f(ld_u16(b));
        uint32_t qs[4];
        lds_words_a2<4>(b + 2, qs);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            o.w[i] = qs[i] & 0x0f0f0f0fu;
            o.w[4 + i] = (qs[i] >> 4) & 0x0f0f0f0fu;
        }
    } else if constexpr (TYPE == T_Q4_1) {
        const uint8_t *b = seg + g * BS;
        o.a0 = o.a1 = ggb_h2f(ld_u16(b));
        o.m0 = ggb_h2f(ld_u16(b + 2));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t q = ld_u32_a4(b + 4 + 4 * i);
            o.w[i] = q & 0x0f0f0f0fu;
            o.w[4 + i] = (q >> 4) & 0x0f0f0f0fu;
        }
    } else if constexpr (TYPE == T_Q5_0) {
        const uint8_t *b = seg + g * BS;
        o.a0 = o.a1 = ggb_h2f(ld_u16(b));
        uint32_t hq[5];                                      // qh, then the 16 bytes of nibbles
        lds_words_a2<5>(b + 2, hq);
        const uint32_t qh = hq[0];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t q = hq[1 + i];
            o.w[i] = (q & 0x0f0f0f0fu) | spread4((qh >> (4 * i)) & 0xf);
            o.w[4 + i] = ((q >> 4) & 0x0f0f0f0fu) | spread4((qh >> (16 + 4 * i)) & 0xf);
        }
    } else if constexpr (TYPE == T_Q5_1) {
        const uint8_t *b = seg + g * BS;
        o.a0 = o.a1 = ggb_h2f(ld_u16(b));
        o.m0 = ggb_h2f(ld_u16(b + 2));
        const uint32_t qh = ld_u32_a4(b + 4);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t q = ld_u32_a4(b + 8 + 4 * i);
            o.w[i] = (q & 0x0f0f0f0fu) | spread4((qh >> (4 * i)) & 0xf);
            o.w[4 + i] = ((q >> 4) & 0x0f0f0f0fu) | spread4((qh >> (16 + 4 * i)) & 0xf);
        }
    } else if constexpr (TYPE == T_Q8_0) {
        const uint8_t *b = seg + g * BS;
        o.a0 = o.a1 = ggb_h2f(ld_u16(b));
        lds_words_a2<8>(b + 2, o.w);
    } else if constexpr (TYPE == T_Q2_K) {
        const uint8_t *b = seg + (g >> 3) * BS;
        const int j = g & 7, n = j >> 2, s = j & 3;
        const float d = ggb_h2f(ld_u16(b + 80)), dmin = ggb_h2f(ld_u16(b + 82));
        const uint32_t s0 = b[8 * n + 2 * s], s1 = b[8 * n + 2 * s + 1];
        o.a0 = d * (float)(s0 & 0xf);
        o.m0 = dmin * (float)(s0 >> 4);
        o.a1 = d * (float)(s1 & 0xf);
        o.m1 = dmin * (float)(s1 >> 4);
#pragma unroll
        for (int i = 0; i < 8; i++) o.w[i] = (ld_u32_a4(b + 16 + 32 * n + 4 * i) >> (2 * s)) & 0x03030303u;
    } else if constexpr (TYPE == T_Q3_K) {
        const uint8_t *b = seg + (g >> 3) * BS;
        const int j = g & 7, n = j >> 2, jj = j & 3;
        const float d = ggb_h2f(ld_u16(b + 108));
        o.a0 = d * (float)q3k_scale(b + 96, 8 * n + 2 * jj);
        o.a1 = d * (float)q3k_scale(b + 96, 8 * n + 2 * jj + 1);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t lo = (ld_u32_a2(b + 32 + 32 * n + 4 * i) >> (2 * jj)) & 0x03030303u;
            const uint32_t hb = (ld_u32_a2(b + 4 * i) >> (4 * n + jj)) & 0x01010101u;
            o.w[i] = lo | (hb << 2);                         // raw = value + 4
        }
    } else if constexpr (TYPE == T_Q4_K) {
        const uint8_t *b = seg + (g >> 3) * BS;
        const int j = g & 7, il = j >> 1, hi = j & 1;
        int sc, mn;
        k4_scale_min(j, b + 4, sc, mn);
        o.a0 = o.a1 = ggb_h2f(ld_u16(b)) * (float)sc;
        o.m0 = o.m1 = ggb_h2f(ld_u16(b + 2)) * (float)mn;
#pragma unroll
        for (int i = 0; i < 8; i++) o.w[i] = (ld_u32_a4(b + 16 + 32 * il + 4 * i) >> (4 * hi)) & 0x0f0f0f0fu;
    } else if constexpr (TYPE == T_Q5_K) {
        const uint8_t *b = seg + (g >> 3) * BS;
        const int j = g & 7, il = j >> 1, hi = j & 1;
        int sc, mn;
        k4_scale_min(j, b + 4, sc, mn);
        o.a0 = o.a1 = ggb_h2f(ld_u16(b)) * (float)sc;
        o.m0 = o.m1 = ggb_h2f(ld_u16(b + 2)) * (float)mn;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t lo = (ld_u32_a4(b + 48 + 32 * il + 4 * i) >> (4 * hi)) & 0x0f0f0f0fu;
            const uint32_t hb = (ld_u32_a4(b + 16 + 4 * i) >> j) & 0x01010101u;
            o.w[i] = lo | (hb << 4);
        }
    } else if constexpr (TYPE == T_IQ4_NL) {
        const uint8_t *b = seg + g * BS;
        o.a0 = o.a1 = ggb_h2f(ld_u16(b));
        uint32_t qs[4];
        lds_words_a2<4>(b + 2, qs);
#pragma unroll
        for (int i = 0; i < 4; i++) iq4nl_lut8(qs[i], o.w[i], o.w[4 + i]);      // codebook values as int8 bytes
    } else if constexpr (TYPE == T_IQ4_XS) {
        const uint8_t *b = seg + (g >> 3) * BS;
        const int ib = g & 7;
        o.a0 = o.a1 = ggb_h2f(ld_u16(b)) * (float)iq4xs_scale(b, ib);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            iq4nl_lut8(ld_u32_a4(b + 8 + 16 * ib + 4 * i), o.w[i], o.w[4 + i]);
        }
    } else if constexpr (kIsIqGrid<TYPE>) {
        // codebook formats (iq_grid.cuh; vecdotq.cuh:607-841): four octets of grid values, signs applied, as int8 words.
        //   IQ2_XXS / IQ3_*: a0 = a1 = (d * (0.5 + ls)) * step          -> combine: (a0 * d8) * dot
        //   IQ2_XS / IQ2_S:  a0 = 0.5 + ls1, a1 = 0.5 + ls2, m0 = d / 4 -> (m0 * d8) * (a0 * dot0 + a1 * dot1)
        //   IQ1_S:           weights grid + 1, a0 = d * (2 s + 1), m0 = the shift -1 +- 1/8  -> a0 * (d8 * dot + s8 * m0)
        //   IQ1_M:           weights 8 * grid +- 1 (exact integers: the reference's sumi + delta * sumy, times 8),
        //                    a0 = 2 s1 + 1, a1 = 2 s2 + 1, m0 = d / 8   -> (m0 * d8) * (a0 * dot0 + a1 * dot1)
        const uint8_t *b = seg + (g >> 3) * BS;
        const int ib = g & 7;
        const float d = ggb_iq::block_d<TYPE>(b);
        int ls[4];
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const ggb_iq::Octet oc = ggb_iq::octet<TYPE>(b, ib, l);
            ls[l] = oc.ls;
            if constexpr (TYPE == T_IQ1_S) {
                o.w[2 * l] = oc.lo;
                o.w[2 * l + 1] = oc.hi;
                o.m0 = oc.neg_shift ? -1.125f : -0.875f;
            } else if constexpr (TYPE == T_IQ1_M) {
                // 8 * (v - 1) +- 1 for v in {0, 1, 2}: bytes 8 v - 9 (shift -1/8) or 8 v - 7 (+1/8), two's complement per byte;
                // 8 v <= 16 and the subtrahend never borrow across bytes once 0x80 is lent to every byte and taken back
                const uint32_t sub = oc.neg_shift ? 0x09090909u : 0x07070707u;
                o.w[2 * l] = (((oc.lo << 3) | 0x80808080u) - sub) ^ 0x80808080u;
                o.w[2 * l + 1] = (((oc.hi << 3) | 0x80808080u) - sub) ^ 0x80808080u;
            } else {
                o.w[2 * l] = ggb_iq::apply_signs4(oc.lo, oc.sign8);
                o.w[2 * l + 1] = ggb_iq::apply_signs4(oc.hi, oc.sign8 >> 4);
            }
        }
        if constexpr (TYPE == T_IQ2_XXS) {
            o.a0 = o.a1 = (d * (0.5f + (float)ls[0])) * 0.25f;
        } else if constexpr (TYPE == T_IQ3_XXS || TYPE == T_IQ3_S) {
            o.a0 = o.a1 = (d * (0.5f + (float)ls[0])) * 0.5f;
        } else if constexpr (TYPE == T_IQ2_XS || TYPE == T_IQ2_S) {
            o.a0 = 0.5f + (float)ls[0];
            o.a1 = 0.5f + (float)ls[2];
            o.m0 = d * 0.25f;
        } else if constexpr (TYPE == T_IQ1_S) {
            o.a0 = o.a1 = d * (float)ls[0];
        } else {   // T_IQ1_M
            o.a0 = (float)ls[0];
            o.a1 = (float)ls[2];
            o.m0 = d * 0.125f;
        }
    } else {   // T_Q6_K
        const uint8_t *b = seg + (g >> 3) * BS;
        const int j = g & 7, ip = j >> 2, r = j & 3;
        const float d = ggb_h2f(ld_u16(b + 208));
        o.a0 = d * (float)(int)(int8_t)b[192 + 8 * ip + 2 * r];
        o.a1 = d * (float)(int)(int8_t)b[192 + 8 * ip + 2 * r + 1];
        const uint8_t *ql = b + 64 * ip + 32 * (r & 1);
        const uint8_t *qh = b + 128 + 32 * ip;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t lo = (ld_u32_a2(ql + 4 * i) >> (4 * (r >> 1))) & 0x0f0f0f0fu;
            const uint32_t hb = (ld_u32_a2(qh + 4 * i) >> (2 * r)) & 0x03030303u;
            o.w[i] = lo | (hb << 4);                         // raw = value + 32
        }
    }
}

// fp32 contribution of one (row, batch, group): the reference's per-type algebra
// (vecdotq.cuh:45-65 q4_0, :69-91 q4_1, :93-125 q5_0, :130-162 q5_1, :166-178 q8_0,
//  :203-231 q2_K, :235-268 q3_K, :272-327 q4_K/q5_K, :331-349 q6_K).
template <int TYPE>
__device__ __forceinline__ float combine(const GroupW &g, int dot0, int dot1, float d8, uint32_t aux) {
    if constexpr (TYPE == T_Q4_0 || TYPE == T_Q5_0) {
        const float s8 = __uint_as_float(aux);
        const float off = (TYPE == T_Q4_0) ? 8.0f : 16.0f;
        return g.a0 * ((float)(dot0 + dot1) * d8 - off * s8);
    } else if constexpr (TYPE == T_Q4_1 || TYPE == T_Q5_1) {
        const float s8 = __uint_as_float(aux);
        const float d4d8 = __half2float(__float2half_rn(g.a0 * d8));      // __hmul2(dm, ds8) in half
        const float m4s8 = __half2float(__float2half_rn(g.m0 * s8));
        return (float)(dot0 + dot1) * d4d8 + m4s8;
    } else if constexpr (TYPE == T_Q8_0 || TYPE == T_IQ4_NL || TYPE == T_IQ4_XS || TYPE == T_IQ2_XXS || TYPE == T_IQ3_XXS ||
                         TYPE == T_IQ3_S) {                                           // vecdotq.cuh:842-888 for the two IQ4
        return (g.a0 * d8) * (float)(dot0 + dot1);
    } else if constexpr (TYPE == T_IQ2_XS || TYPE == T_IQ2_S || TYPE == T_IQ1_M) {    // vecdotq.cuh:631-704, 797-841
        return (g.m0 * d8) * (g.a0 * (float)dot0 + g.a1 * (float)dot1);
    } else if constexpr (TYPE == T_IQ1_S) {                                          // vecdotq.cuh:762-795
        return g.a0 * (d8 * (float)(dot0 + dot1) + __uint_as_float(aux) * g.m0);
    } else {
        const int sq0 = (int)(int16_t)(aux & 0xffffu), sq1 = (int)(int16_t)(aux >> 16);
        if constexpr (TYPE == T_Q2_K) {
            return d8 * (g.a0 * (float)dot0 + g.a1 * (float)dot1) - d8 * (g.m0 * (float)sq0 + g.m1 * (float)sq1);
        } else if constexpr (TYPE == T_Q3_K) {
            return d8 * (g.a0 * (float)(dot0 - 4 * sq0) + g.a1 * (float)(dot1 - 4 * sq1));
        } else if constexpr (TYPE == T_Q6_K) {
            return d8 * (g.a0 * (float)(dot0 - 32 * sq0) + g.a1 * (float)(dot1 - 32 * sq1));
        } else {   // Q4_K, Q5_K
            return d8 * (g.a0 * (float)(dot0 + dot1)) - d8 * (g.m0 * (float)(sq0 + sq1));
        }
    }
}


}  // namespace ggb_group
