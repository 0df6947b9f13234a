// iq_grid.cuh -- the codebook ("grid") formats IQ2_XXS, IQ2_XS, IQ2_S, IQ3_XXS, IQ3_S, IQ1_S, IQ1_M: decode of one
// 8-value octet from raw GGUF bytes, shared by the dequantizers (unpack_iq.cuh) and the MMVQ group unpack (mmvq_group_ops.cuh).
//
// Replaces dequantize_block_iq* (hf-kernels/ggml-kernels/ggml/dequantize.cuh:256-399) and the weight side of
// vec_dot_iq*_q8_1 (ggml/vecdotq.cuh:607-841); block structs ggml/ggml-common.h:110-168.  The codebooks come from
// iq_tables.cuh, GENERATED from gguf-py (scripts/gen_iq_tables.py, which also compares them with the reference's tables);
// they are read through the read-only data path (__ldg: 2-16 KB per table, L1-resident).  Byte j of an entry = value j.
//
// A super-block is 8 groups `ib` of 32 values, a group 4 octets `l` of 8 values.  Per type, relative to the block start b:
//   IQ2_XXS (66 B):  d @0 | per ib 8 bytes @2+8ib: 4 grid indices, then a word {4 x 7-bit sign index, 4-bit scale}
//   IQ2_XS  (74 B):  d @0 | 32 x u16 @2: {9-bit grid index, 7-bit sign index} | 8 scale bytes @66 (nibble per 16 values)
//   IQ2_S   (82 B):  d @0 | 32 grid-index bytes @2 | 32 sign bytes @34 | 8 x qh @66 (2 high index bits per octet) | 8 scale bytes @74
//   IQ3_XXS (98 B):  d @0 | 64 grid-index bytes @2 (4 values each) | per ib a word @66+4ib {4 x 7-bit sign index, 4-bit scale}
//   IQ3_S  (110 B):  d @0 | 64 grid-index bytes @2 | 8 x qh @66 (ninth index bits) | 32 sign bytes @74 | 4 scale bytes @106
//   IQ1_S   (50 B):  d @0 | 32 grid-index bytes @2 | 8 x u16 qh @34: {4 x 3 high index bits, 3-bit scale, shift sign}
//   IQ1_M   (56 B):  32 grid-index bytes @0 | 16 x qh @32 (per octet: 3 high index bits + shift sign) | 4 x u16 @48: 4 x 3-bit
//                    scales each + one nibble of the fp16 super-block scale
// Every block size is even and nothing else: all multi-byte fields are read with 2-byte-aligned loads.
#pragma once

#include "ggb_common.cuh"
#include "iq_tables.cuh"

namespace ggb_iq {

// the 7-bit sign index plus the eighth bit that makes the parity even (the reference tabulates this as ksigns_iq2xs)
__device__ __forceinline__ uint32_t ksign8(uint32_t i7) { return i7 | ((uint32_t)(__popc(i7) & 1) << 7); }

// bit i of x (i < 4) -> 0x01 in byte i
__device__ __forceinline__ uint32_t spread4_bit0(uint32_t x) { return (x * 0x00204081u) & 0x01010101u; }

// four unsigned grid bytes (all >= 1) negated where the matching bit of sg4 is set, as int8 bytes: ~g + 1 per byte never
// carries because no grid value of the signed formats is 0
__device__ __forceinline__ uint32_t apply_signs4(uint32_t g, uint32_t sg4) {
    const uint32_t one = spread4_bit0(sg4 & 0xfu);
    return (g ^ (one * 0xffu)) + one;
}

struct Octet {
    uint32_t lo, hi;      // values 0..3 / 4..7, one byte each: unsigned magnitudes (IQ2 / IQ3), or grid + 1 in {0, 1, 2} (IQ1)
    uint32_t sign8;       // IQ2 / IQ3: bit j = value j is negative
    int ls;               // the octet's integer sub-scale as the reference uses it (IQ2 / IQ3: 0..15; IQ1: 2 s + 1)
    bool neg_shift;       // IQ1: the shift is -1 - 1/8 instead of -1 + 1/8
};

template <int TYPE>
__device__ __forceinline__ Octet octet(const uint8_t *b, int ib, int l) {
    Octet o;
    o.sign8 = 0;
    o.neg_shift = false;
    if constexpr (TYPE == T_IQ2_XXS) {
        const uint8_t *q2 = b + 2 + 8 * ib;
        const uint32_t aux = ld_u32_a2(q2 + 4);
        const uint64_t g = __ldg(&iq2xxs_grid[q2[l]]);
        o.lo = (uint32_t)g; o.hi = (uint32_t)(g >> 32);
        o.sign8 = ksign8((aux >> (7 * l)) & 127u);
        o.ls = (int)(aux >> 28);
    } else if constexpr (TYPE == T_IQ2_XS) {
        const uint32_t q = ld_u16(b + 2 + 2 * (4 * ib + l));
        const uint64_t g = __ldg(&iq2xs_grid[q & 511u]);
        o.lo = (uint32_t)g; o.hi = (uint32_t)(g >> 32);
        o.sign8 = ksign8(q >> 9);
        o.ls = (b[66 + ib] >> (4 * (l >> 1))) & 0xf;
    } else if constexpr (TYPE == T_IQ2_S) {
        const uint32_t idx = b[2 + 4 * ib + l] | (((uint32_t)b[66 + ib] << (8 - 2 * l)) & 0x300u);
        const uint64_t g = __ldg(&iq2s_grid[idx]);
        o.lo = (uint32_t)g; o.hi = (uint32_t)(g >> 32);
        o.sign8 = b[34 + 4 * ib + l];
        o.ls = (b[74 + ib] >> (4 * (l >> 1))) & 0xf;
    } else if constexpr (TYPE == T_IQ3_XXS) {
        const uint8_t *q3 = b + 2 + 8 * ib + 2 * l;
        const uint32_t aux = ld_u32_a2(b + 66 + 4 * ib);
        o.lo = __ldg(&iq3xxs_grid[q3[0]]);
        o.hi = __ldg(&iq3xxs_grid[q3[1]]);
        o.sign8 = ksign8((aux >> (7 * l)) & 127u);
        o.ls = (int)(aux >> 28);
    } else if constexpr (TYPE == T_IQ3_S) {
        const uint8_t *qs = b + 2 + 8 * ib + 2 * l;
        const uint32_t qh = b[66 + ib];
        o.lo = __ldg(&iq3s_grid[qs[0] | ((qh << (8 - 2 * l)) & 256u)]);
        o.hi = __ldg(&iq3s_grid[qs[1] | ((qh << (7 - 2 * l)) & 256u)]);
        o.sign8 = b[74 + 4 * ib + l];
        o.ls = (b[106 + (ib >> 1)] >> (4 * (ib & 1))) & 0xf;
    } else if constexpr (TYPE == T_IQ1_S) {
        const uint32_t qh = ld_u16(b + 34 + 2 * ib);
        const uint64_t g = __ldg(&iq1s_grid[b[2 + 4 * ib + l] | (((qh >> (3 * l)) & 7u) << 8)]);
        o.lo = (uint32_t)g; o.hi = (uint32_t)(g >> 32);
        o.ls = 2 * (int)((qh >> 12) & 7u) + 1;
        o.neg_shift = (qh & 0x8000u) != 0;
    } else {   // T_IQ1_M
        const uint32_t qhl = (uint32_t)b[32 + 2 * ib + (l >> 1)] >> (4 * (l & 1));
        const uint64_t g = __ldg(&iq1s_grid[b[4 * ib + l] | ((qhl & 7u) << 8)]);
        o.lo = (uint32_t)g; o.hi = (uint32_t)(g >> 32);
        const uint32_t sc = ld_u16(b + 48 + 2 * (ib >> 1)) >> (6 * (ib & 1) + 3 * (l >> 1));
        o.ls = 2 * (int)(sc & 7u) + 1;
        o.neg_shift = (qhl & 8u) != 0;
    }
    return o;
}

// the super-block's fp16 scale as fp32 (IQ1_M: gathered from the top nibbles of its four 16-bit scale words)
template <int TYPE>
__device__ __forceinline__ float block_d(const uint8_t *b) {
    if constexpr (TYPE == T_IQ1_M) {
        const uint32_t s0 = ld_u16(b + 48), s1 = ld_u16(b + 50), s2 = ld_u16(b + 52), s3 = ld_u16(b + 54);
        return ggb_h2f((uint16_t)((s0 >> 12) | ((s1 >> 8) & 0x00f0u) | ((s2 >> 4) & 0x0f00u) | (s3 & 0xf000u)));
    } else {
        return ggb_h2f((uint16_t)ld_u16(b));
    }
}

}  // namespace ggb_iq
