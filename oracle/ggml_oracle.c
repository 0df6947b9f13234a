/*
 * ggml_oracle.c -- CPU restatement of the reference's GGUF quantized-weight hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (ggml-libtorch_b200/) never links, imports or calls it and has no CPU fallback.
 *
 * What is restated (reference paths relative to /root/reference):
 *   - legacy dequantize Q4_0/Q4_1/Q5_0/Q5_1/Q8_0   ggml-cpu/ggml-quants.hpp:4-112
 *   - K-quant dequantize Q2_K..Q6_K: element layout hf-kernels/ggml-kernels/ggml/dequantize.cuh:101-254,
 *     block structs ggml/ggml-common.h:67-108, fp32 operation order of gguf-py 0.19.0
 *     gguf/quants.py:404-572 (third-party `gguf`, pinned >=0.10.0 by requirements-dev.txt:3;
 *     it is the expected-value source of every reference test)
 *   - IQ4_NL / IQ4_XS dequantize: block structs ggml/ggml-common.h (block_iq4_nl,
 *     block_iq4_xs), element layout ggml/dequantize.cuh:400-433, the 16-entry codebook kvalues_iq4nl
 *     (ggml-common.h:1060), fp32 operation order of gguf-py quants.py (classes IQ4_NL, IQ4_XS); MMVQ math
 *     ggml/vecdotq.cuh:842-888.  The reference has no MMQ for these two types and neither has orc_mmq.
 *   - IQ2_XXS / IQ2_XS / IQ2_S / IQ3_XXS / IQ3_S / IQ1_S / IQ1_M (rest of SURVEY.md section 8f rank 3): block structs
 *     ggml/ggml-common.h:110-168, dequantize ggml/dequantize.cuh:256-399 (fp32 operation order, identical to gguf-py's
 *     classes of the same names), MMVQ math ggml/vecdotq.cuh:607-841.  The codebooks (iq_tables.h) are GENERATED from
 *     gguf-py by scripts/gen_iq_tables.py and were compared there with the reference's tables (ggml-common.h:193-1057).
 *     IQ3_S follows the REFERENCE (grid levels 4..52, 62 with d * (0.5 + ls) * 0.5), which differs from today's gguf-py
 *     (levels 1..13, 15 with d * (1 + 2 ls)) on the top level only: 15.5 vs 15.
 *   - q8_1 activation quantization               ggml/ggml_kernel.cu:13-50
 *   - q8_1_mmq activation quantization           ggml/mmq.cu:84-154, ggml/mmq.cuh:176-181
 *   - MMVQ math (integer dots + scale algebra)   ggml/vecdotq.cuh:45-345, 347-605; ggml/mmvq.cuh:2-38
 *   - MMQ math                                   ggml/mmq.cuh (vec_dot_*_q8_1_{dp4a,mma}), ggml/mmq.cu:180-255
 *
 * Parity pins: dequantize is pinned bit-for-bit against oracle/_ref (the reference's own
 * ggml-cpu op, 5 legacy types) and against gguf-py (all 10 types); see tests/test_oracle.py
 * and tests/golden/.  q8_1 bytes and MMVQ/MMQ values: PARITY UNPINNED by the reference (it has
 * no CPU implementation and no test that inspects them); the restatement follows the CUDA
 * source line by line and is cross-checked against x @ dequant(W).T within the reference's
 * own tolerances.
 *
 * Every fp32 product/sum below is a separately rounded operation: build with -ffp-contract=off.
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ggml_oracle.h"
#include "iq_tables.h"

/* ------------------------------------------------------------------ fp16 helpers */

static inline float h2f(uint16_t h) {
    /* exact widening of IEEE binary16 */
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp  = (h >> 10) & 0x1fu;
    uint32_t man  = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do { man <<= 1; e++; } while ((man & 0x400u) == 0);
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static inline uint16_t f2h(float f) {
    /* round-to-nearest-even narrowing, == CUDA __float2half / __float2half_rn */
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) {                       /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? (0x200u | ((ax >> 13) & 0x3ffu)) : 0));
    }
    if (ax >= 0x477ff000u) {                       /* >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x33000001u) {                        /* < 2^-25 (or == 2^-25 ties to 0) */
        return (uint16_t)sign;
    }
    int32_t e = (int32_t)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) {                                 /* subnormal result */
        shift = 13 + (-14 - e);
        base = 0;
    } else {
        shift = 13;
        base = (uint32_t)(e + 15) << 10;
        m &= 0x7fffffu;
    }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | (base + q));          /* carry into exponent is correct by construction */
}

static inline uint16_t ld16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

/* ------------------------------------------------------------------ type table */

int orc_block_elems(int type) {
    switch (type) {
        case ORC_Q4_0: case ORC_Q4_1: case ORC_Q5_0: case ORC_Q5_1: case ORC_Q8_0: case ORC_Q8_1: return 32;
        case ORC_Q2_K: case ORC_Q3_K: case ORC_Q4_K: case ORC_Q5_K: case ORC_Q6_K: return 256;
        case ORC_IQ4_NL: return 32;
        case ORC_IQ4_XS: return 256;
        case ORC_IQ2_XXS: case ORC_IQ2_XS: case ORC_IQ2_S: case ORC_IQ3_XXS: case ORC_IQ3_S: case ORC_IQ1_S: case ORC_IQ1_M: return 256;
        default: return 0;
    }
}

int orc_block_bytes(int type) {
    switch (type) {
        case ORC_Q4_0: return 18;  case ORC_Q4_1: return 20;  case ORC_Q5_0: return 22;
        case ORC_Q5_1: return 24;  case ORC_Q8_0: return 34;  case ORC_Q8_1: return 36;
        case ORC_Q2_K: return 84;  case ORC_Q3_K: return 110; case ORC_Q4_K: return 144;
        case ORC_Q5_K: return 176; case ORC_Q6_K: return 210;
        case ORC_IQ4_NL: return 18; case ORC_IQ4_XS: return 136;
        case ORC_IQ2_XXS: return 66; case ORC_IQ2_XS: return 74; case ORC_IQ2_S: return 82;
        case ORC_IQ3_XXS: return 98; case ORC_IQ3_S: return 110; case ORC_IQ1_S: return 50; case ORC_IQ1_M: return 56;
        default: return 0;
    }
}

/* Q4_K / Q5_K 6-bit (scale,min) pairs: ggml/dequantize.cuh:154-161 */
static inline void k4_scale_min(int j, const uint8_t *s, int *sc, int *mn) {
    if (j < 4) {
        *sc = s[j] & 63;
        *mn = s[j + 4] & 63;
    } else {
        *sc = (s[j + 4] & 0x0f) | ((s[j - 4] >> 6) << 4);
        *mn = (s[j + 4] >> 4)   | ((s[j]     >> 6) << 4);
    }
}

/* Q3_K 6-bit signed scale `is` of 16: gguf/quants.py:442-454, ggml/dequantize.cuh:140-145 */
static inline int q3k_scale(const uint8_t *s, int is) {
    int lo = (is < 8) ? (s[is] & 0x0f) : (s[is - 8] >> 4);
    int hi = (s[8 + (is & 3)] >> (2 * (is >> 2))) & 3;
    return (lo | (hi << 4)) - 32;
}

/* ------------------------------------------------------------------ integer unpack (shared by dequant and dots) */

/* Write the 32 or 256 integer quants of one block, in element order, before any scaling.
 * For types with an additive offset (q-8, q-16, q-32, q3 -4) the offset is already applied. */
/* the non-linear 4-bit codebook of IQ4_NL / IQ4_XS (a format constant: ggml-common.h:1060, gguf-py IQ4_NL.kvalues) */
static const int8_t kvalues_iq4nl[16] = {-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113};

/* IQ4_XS 6-bit sub-block scale ib (0..7), offset -32 applied: low 4 bits from scales_l (bytes 4..7), high 2 from scales_h
 * (bytes 2..3)   (dequantize.cuh:426-427, vecdotq.cuh:877) */
static inline int iq4xs_scale(const uint8_t *b, int ib) {
    const int lo = (b[4 + ib / 2] >> (4 * (ib % 2))) & 0x0f;
    const int hi = (ld16(b + 2) >> (2 * ib)) & 3;
    return (lo | (hi << 4)) - 32;
}

static void unpack_block_ints(int type, const uint8_t *b, int *q) {
    switch (type) {
    case ORC_Q4_0:                                   /* ggml-quants.hpp:11-20 */
        for (int j = 0; j < 16; j++) { q[j] = (b[2 + j] & 0x0f) - 8; q[j + 16] = (b[2 + j] >> 4) - 8; }
        break;
    case ORC_Q4_1:                                   /* ggml-quants.hpp:31-41 */
        for (int j = 0; j < 16; j++) { q[j] = b[4 + j] & 0x0f; q[j + 16] = b[4 + j] >> 4; }
        break;
    case ORC_Q5_0: {                                 /* ggml-quants.hpp:52-67 */
        uint32_t qh; memcpy(&qh, b + 2, 4);
        for (int j = 0; j < 16; j++) {
            int h0 = ((qh >> j) << 4) & 0x10, h1 = (qh >> (j + 12)) & 0x10;
            q[j] = ((b[6 + j] & 0x0f) | h0) - 16; q[j + 16] = ((b[6 + j] >> 4) | h1) - 16;
        }
        break; }
    case ORC_Q5_1: {                                 /* ggml-quants.hpp:78-94 */
        uint32_t qh; memcpy(&qh, b + 4, 4);
        for (int j = 0; j < 16; j++) {
            int h0 = ((qh >> j) << 4) & 0x10, h1 = (qh >> (j + 12)) & 0x10;
            q[j] = (b[8 + j] & 0x0f) | h0; q[j + 16] = (b[8 + j] >> 4) | h1;
        }
        break; }
    case ORC_Q8_0:                                   /* ggml-quants.hpp:105-110 */
        for (int j = 0; j < 32; j++) q[j] = (int8_t)b[2 + j];
        break;
    case ORC_Q2_K:                                   /* dequantize.cuh:107-120 */
        for (int n = 0; n < 2; n++) for (int s = 0; s < 4; s++) for (int l = 0; l < 32; l++)
            q[128 * n + 32 * s + l] = (b[16 + 32 * n + l] >> (2 * s)) & 3;
        break;
    case ORC_Q3_K:                                   /* dequantize.cuh:129-151 */
        for (int n = 0; n < 2; n++) for (int j = 0; j < 4; j++) for (int l = 0; l < 32; l++) {
            int lo = (b[32 + 32 * n + l] >> (2 * j)) & 3;
            int hb = (b[l] >> (4 * n + j)) & 1;
            q[128 * n + 32 * j + l] = lo - (hb ? 0 : 4);
        }
        break;
    case ORC_Q4_K:                                   /* dequantize.cuh:171-193 */
        for (int il = 0; il < 4; il++) for (int l = 0; l < 32; l++) {
            q[64 * il + l]      = b[16 + 32 * il + l] & 0x0f;
            q[64 * il + 32 + l] = b[16 + 32 * il + l] >> 4;
        }
        break;
    case ORC_Q5_K:                                   /* dequantize.cuh:204-227 */
        for (int il = 0; il < 4; il++) for (int l = 0; l < 32; l++) {
            int h = b[16 + l];
            q[64 * il + l]      = (b[48 + 32 * il + l] & 0x0f) | (((h >> (2 * il)) & 1) << 4);
            q[64 * il + 32 + l] = (b[48 + 32 * il + l] >> 4)   | (((h >> (2 * il + 1)) & 1) << 4);
        }
        break;
    case ORC_Q6_K:                                   /* dequantize.cuh:237-253 */
        for (int ip = 0; ip < 2; ip++) for (int l = 0; l < 32; l++) {
            const uint8_t *ql = b + 64 * ip; int h = b[128 + 32 * ip + l];
            q[128 * ip + l]      = ((ql[l] & 0x0f)      | (((h >> 0) & 3) << 4)) - 32;
            q[128 * ip + 32 + l] = ((ql[l + 32] & 0x0f) | (((h >> 2) & 3) << 4)) - 32;
            q[128 * ip + 64 + l] = ((ql[l] >> 4)        | (((h >> 4) & 3) << 4)) - 32;
            q[128 * ip + 96 + l] = ((ql[l + 32] >> 4)   | (((h >> 6) & 3) << 4)) - 32;
        }
        break;
    case ORC_IQ4_NL:                                 /* dequantize.cuh:400-416: codebook values */
        for (int j = 0; j < 16; j++) { q[j] = kvalues_iq4nl[b[2 + j] & 0x0f]; q[j + 16] = kvalues_iq4nl[b[2 + j] >> 4]; }
        break;
    case ORC_IQ4_XS:                                 /* dequantize.cuh:419-433 */
        for (int ib = 0; ib < 8; ib++) for (int j = 0; j < 16; j++) {
            q[32 * ib + j]      = kvalues_iq4nl[b[8 + 16 * ib + j] & 0x0f];
            q[32 * ib + 16 + j] = kvalues_iq4nl[b[8 + 16 * ib + j] >> 4];
        }
        break;
    default: break;
    }
}


/* ------------------------------------------------------------------ IQ1 / IQ2 / IQ3 codebook formats */

static inline int is_iq_grid_type(int type) {
    return type == ORC_IQ2_XXS || type == ORC_IQ2_XS || type == ORC_IQ2_S || type == ORC_IQ3_XXS || type == ORC_IQ3_S ||
           type == ORC_IQ1_S || type == ORC_IQ1_M;
}
static inline uint32_t ld32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
/* ksigns_iq2xs (ggml-common.h:1013): the 7-bit index plus an eighth bit that makes the parity even */
static inline uint32_t ksign8(uint32_t i7) {
    uint32_t p = i7 ^ (i7 >> 4); p ^= p >> 2; p ^= p >> 1;
    return i7 | ((p & 1u) << 7);
}

/* One 32-value group ib (0..7) of a super-block: the integers the reference's vec_dot multiplies with the q8 quants
 * (IQ2 / IQ3: signed grid values; IQ1: grid + 1 in {0, 1, 2}) and the group's scale parts:
 *   sc[0], sc[1]  integer sub-scales of elements 0..15 / 16..31 as the reference uses them (ls; IQ1: 2 * s + 1)
 *   dlt[l]        IQ1: the additive shift of the 8 values l (-1 +- 0.125), else 0 */
static void iq_group(int type, const uint8_t *b, int ib, int *q, int *sc, float *dlt) {
    for (int l = 0; l < 4; l++) dlt[l] = 0.0f;
    switch (type) {
    case ORC_IQ2_XXS: {                              /* dequantize.cuh:256-273, vecdotq.cuh:607-629 */
        const uint8_t *q2 = b + 2 + 8 * ib;
        const uint32_t aux32 = ld32(q2 + 4);
        sc[0] = sc[1] = (int)(aux32 >> 28);
        for (int l = 0; l < 4; l++) {
            const uint64_t g = iq2xxs_grid[q2[l]];
            const uint32_t sg = ksign8((aux32 >> (7 * l)) & 127);
            for (int j = 0; j < 8; j++) { int v = (int)((g >> (8 * j)) & 255); q[8 * l + j] = (sg >> j) & 1 ? -v : v; }
        }
        break; }
    case ORC_IQ2_XS: {                               /* dequantize.cuh:275-291, vecdotq.cuh:631-663 */
        sc[0] = b[66 + ib] & 0x0f; sc[1] = b[66 + ib] >> 4;
        for (int l = 0; l < 4; l++) {
            const uint32_t qv = ld16(b + 2 + 2 * (4 * ib + l));
            const uint64_t g = iq2xs_grid[qv & 511];
            const uint32_t sg = ksign8(qv >> 9);
            for (int j = 0; j < 8; j++) { int v = (int)((g >> (8 * j)) & 255); q[8 * l + j] = (sg >> j) & 1 ? -v : v; }
        }
        break; }
    case ORC_IQ2_S: {                                /* dequantize.cuh:293-307, vecdotq.cuh:665-704 */
        sc[0] = b[74 + ib] & 0x0f; sc[1] = b[74 + ib] >> 4;
        for (int l = 0; l < 4; l++) {
            const uint32_t idx = b[2 + 4 * ib + l] | (((uint32_t)b[66 + ib] << (8 - 2 * l)) & 0x300);
            const uint64_t g = iq2s_grid[idx];
            const uint32_t sg = b[34 + 4 * ib + l];
            for (int j = 0; j < 8; j++) { int v = (int)((g >> (8 * j)) & 255); q[8 * l + j] = (sg >> j) & 1 ? -v : v; }
        }
        break; }
    case ORC_IQ3_XXS: {                              /* dequantize.cuh:309-331, vecdotq.cuh:706-732 */
        const uint8_t *q3 = b + 2 + 8 * ib;
        const uint32_t aux32 = ld32(b + 66 + 4 * ib);
        sc[0] = sc[1] = (int)(aux32 >> 28);
        for (int l = 0; l < 4; l++) {
            const uint32_t g1 = iq3xxs_grid[q3[2 * l]], g2 = iq3xxs_grid[q3[2 * l + 1]];
            const uint32_t sg = ksign8((aux32 >> (7 * l)) & 127);
            for (int j = 0; j < 4; j++) {
                int v1 = (int)((g1 >> (8 * j)) & 255), v2 = (int)((g2 >> (8 * j)) & 255);
                q[8 * l + j] = (sg >> j) & 1 ? -v1 : v1;
                q[8 * l + 4 + j] = (sg >> (4 + j)) & 1 ? -v2 : v2;
            }
        }
        break; }
    case ORC_IQ3_S: {                                /* dequantize.cuh:333-354, vecdotq.cuh:734-760 */
        const uint8_t *qs = b + 2 + 8 * ib;
        const uint32_t qh = b[66 + ib];
        sc[0] = sc[1] = (b[106 + ib / 2] >> (4 * (ib % 2))) & 0x0f;
        for (int l = 0; l < 4; l++) {
            const uint32_t g1 = iq3s_grid[qs[2 * l] | ((qh << (8 - 2 * l)) & 256)];
            const uint32_t g2 = iq3s_grid[qs[2 * l + 1] | ((qh << (7 - 2 * l)) & 256)];
            const uint32_t sg = b[74 + 4 * ib + l];
            for (int j = 0; j < 4; j++) {
                int v1 = (int)((g1 >> (8 * j)) & 255), v2 = (int)((g2 >> (8 * j)) & 255);
                q[8 * l + j] = (sg >> j) & 1 ? -v1 : v1;
                q[8 * l + 4 + j] = (sg >> (4 + j)) & 1 ? -v2 : v2;
            }
        }
        break; }
    case ORC_IQ1_S: {                                /* dequantize.cuh:356-376, vecdotq.cuh:762-795 */
        const uint32_t qh = ld16(b + 34 + 2 * ib);
        sc[0] = sc[1] = 2 * (int)((qh >> 12) & 7) + 1;
        for (int l = 0; l < 4; l++) {
            const uint64_t g = iq1s_grid[b[2 + 4 * ib + l] | (((qh >> (3 * l)) & 7) << 8)];
            dlt[l] = (qh & 0x8000) ? -1.0f - 0.125f : -1.0f + 0.125f;
            for (int j = 0; j < 8; j++) q[8 * l + j] = (int)((g >> (8 * j)) & 255);
        }
        break; }
    case ORC_IQ1_M: {                                /* dequantize.cuh:378-399, vecdotq.cuh:797-841 */
        const uint32_t tmp = ld16(b + 48 + 2 * (ib / 2)) >> (6 * (ib % 2));
        sc[0] = 2 * (int)(tmp & 7) + 1; sc[1] = 2 * (int)((tmp >> 3) & 7) + 1;
        for (int l = 0; l < 4; l++) {
            const uint32_t qhl = (uint32_t)b[32 + 2 * ib + l / 2] >> (4 * (l % 2));
            const uint64_t g = iq1s_grid[b[4 * ib + l] | ((qhl & 7) << 8)];
            dlt[l] = (qhl & 8) ? -1.0f - 0.125f : -1.0f + 0.125f;
            for (int j = 0; j < 8; j++) q[8 * l + j] = (int)((g >> (8 * j)) & 255);
        }
        break; }
    default: break;
    }
}
/* the super-block's fp16 scale (IQ1_M: spread over the top nibbles of its four 16-bit scale words) */
static inline float iq_block_d(int type, const uint8_t *b) {
    if (type != ORC_IQ1_M) return h2f(ld16(b));
    const uint32_t s0 = ld16(b + 48), s1 = ld16(b + 50), s2 = ld16(b + 52), s3 = ld16(b + 54);
    return h2f((uint16_t)((s0 >> 12) | ((s1 >> 8) & 0x00f0) | ((s2 >> 4) & 0x0f00) | (s3 & 0xf000)));
}

static void iq_dequant_block(int type, const uint8_t *b, float *y) {
    const float d = iq_block_d(type, b);
    for (int ib = 0; ib < 8; ib++) {
        int q[32], sc[2]; float dlt[4];
        iq_group(type, b, ib, q, sc, dlt);
        for (int l = 0; l < 4; l++) {
            const int s = sc[l / 2];
            float db;
            if (type == ORC_IQ1_S || type == ORC_IQ1_M) {
                db = d * (float)s;                                       /* d * (2 s + 1) */
                for (int j = 0; j < 8; j++) { const float t = (float)q[8 * l + j] + dlt[l]; y[32 * ib + 8 * l + j] = db * t; }
            } else {
                db = d * (0.5f + (float)s);
                db = db * ((type == ORC_IQ3_XXS || type == ORC_IQ3_S) ? 0.5f : 0.25f);
                /* d * grid * sign: the sign is an exact factor */
                for (int j = 0; j < 8; j++) {
                    const int v = q[8 * l + j];
                    const float t = db * (float)(v < 0 ? -v : v);
                    y[32 * ib + 8 * l + j] = v < 0 ? -t : t;
                }
            }
        }
    }
}

/* one super-block against 256 q8_1 activations (vecdotq.cuh:607-841), double accumulation */
static double iq_block_dot(int type, const uint8_t *b, const int8_t *a, const float *dd, const float *ss) {
    const double d = iq_block_d(type, b);
    double acc = 0.0;
    for (int ib = 0; ib < 8; ib++) {
        int q[32], sc[2]; float dlt[4];
        iq_group(type, b, ib, q, sc, dlt);
        const int8_t *u = a + 32 * ib;
        int s0 = 0, s1 = 0;
        for (int j = 0; j < 16; j++) { s0 += q[j] * u[j]; s1 += q[16 + j] * u[16 + j]; }
        switch (type) {
        case ORC_IQ2_XXS: acc += d * (0.5 + sc[0]) * (double)dd[ib] * 0.25 * (s0 + s1); break;
        case ORC_IQ2_XS: case ORC_IQ2_S:
            acc += d * (double)dd[ib] * 0.25 * ((0.5 + sc[0]) * s0 + (0.5 + sc[1]) * s1); break;
        case ORC_IQ3_XXS: case ORC_IQ3_S: acc += d * (0.5 + sc[0]) * (double)dd[ib] * 0.5 * (s0 + s1); break;
        case ORC_IQ1_S:                              /* d1q * (d8 * sumi + s8 * delta): the shift goes through q8_1's s field */
            acc += d * sc[0] * ((double)dd[ib] * (s0 + s1) + (double)ss[ib] * (double)dlt[0]); break;
        case ORC_IQ1_M: {                            /* the shift goes through the integer sums of the quants */
            double f[2] = {0.0, 0.0};
            for (int l = 0; l < 4; l++) {
                int sumy = 0;
                for (int j = 0; j < 8; j++) sumy += u[8 * l + j];
                f[l / 2] += (double)dlt[l] * sumy;
            }
            acc += d * (double)dd[ib] * ((s0 + f[0]) * sc[0] + (s1 + f[1]) * sc[1]); break; }
        default: break;
        }
    }
    return acc;
}

/* ------------------------------------------------------------------ dequantize */

static void dequant_block(int type, const uint8_t *b, float *y) {
    int q[256];
    if (is_iq_grid_type(type)) { iq_dequant_block(type, b, y); return; }
    unpack_block_ints(type, b, q);
    switch (type) {
    case ORC_Q4_0: case ORC_Q5_0: {                  /* y = (q - off) * d */
        const float d = h2f(ld16(b));
        for (int j = 0; j < 32; j++) y[j] = (float)q[j] * d;
        break; }
    case ORC_Q8_0: {
        const float d = h2f(ld16(b));
        for (int j = 0; j < 32; j++) y[j] = (float)q[j] * d;
        break; }
    case ORC_Q4_1: case ORC_Q5_1: {                  /* y = q*d + m  (two roundings) */
        const float d = h2f(ld16(b)), m = h2f(ld16(b + 2));
        for (int j = 0; j < 32; j++) { float t = (float)q[j] * d; y[j] = t + m; }
        break; }
    case ORC_Q2_K: {                                 /* quants.py:406-428 */
        const float d = h2f(ld16(b + 80)), dmin = h2f(ld16(b + 82));
        for (int is = 0; is < 16; is++) {
            const float dl = d * (float)(b[is] & 0x0f);
            const float ml = dmin * (float)(b[is] >> 4);
            for (int l = 0; l < 16; l++) { float t = dl * (float)q[16 * is + l]; y[16 * is + l] = t - ml; }
        }
        break; }
    case ORC_Q3_K: {                                 /* quants.py:433-472 */
        const float d = h2f(ld16(b + 108));
        for (int is = 0; is < 16; is++) {
            const float dl = d * (float)q3k_scale(b + 96, is);
            for (int l = 0; l < 16; l++) y[16 * is + l] = dl * (float)q[16 * is + l];
        }
        break; }
    case ORC_Q4_K: case ORC_Q5_K: {                  /* quants.py:504-549 */
        const float d = h2f(ld16(b)), dmin = h2f(ld16(b + 2));
        for (int j = 0; j < 8; j++) {
            int sc, mn; k4_scale_min(j, b + 4, &sc, &mn);
            const float dl = d * (float)sc;
            const float ml = dmin * (float)mn;
            for (int l = 0; l < 32; l++) { float t = dl * (float)q[32 * j + l]; y[32 * j + l] = t - ml; }
        }
        break; }
    case ORC_Q6_K: {                                 /* quants.py:554-572 */
        const float d = h2f(ld16(b + 208));
        for (int is = 0; is < 16; is++) {
            const float dl = d * (float)(int8_t)b[192 + is];
            for (int l = 0; l < 16; l++) y[16 * is + l] = dl * (float)q[16 * is + l];
        }
        break; }
    case ORC_IQ4_NL: {                               /* gguf-py IQ4_NL: y = d * kvalues[q] */
        const float d = h2f(ld16(b));
        for (int j = 0; j < 32; j++) y[j] = d * (float)q[j];
        break; }
    case ORC_IQ4_XS: {                               /* gguf-py IQ4_XS: dl = d * (ls - 32); y = dl * kvalues[q] */
        const float d = h2f(ld16(b));
        for (int ib = 0; ib < 8; ib++) {
            const float dl = d * (float)iq4xs_scale(b, ib);
            for (int l = 0; l < 32; l++) y[32 * ib + l] = dl * (float)q[32 * ib + l];
        }
        break; }
    default: break;
    }
}

int orc_dequantize(const void *w, int type, int64_t nelem, float *out) {
    const int qk = orc_block_elems(type), bs = orc_block_bytes(type);
    if (qk == 0 || type == ORC_Q8_1) return -1;
    if (nelem % qk) return -2;
    const uint8_t *p = (const uint8_t *)w;
    const int64_t nb = nelem / qk;
    for (int64_t i = 0; i < nb; i++) dequant_block(type, p + i * bs, out + i * qk);
    return 0;
}

int orc_dequantize_mt(const void *w, int type, int64_t rows, int64_t cols, float *out) {
    const int qk = orc_block_elems(type), bs = orc_block_bytes(type);
    if (qk == 0 || type == ORC_Q8_1) return -1;
    if (cols % qk) return -2;
    const uint8_t *p = (const uint8_t *)w;
    const int64_t row_bytes = cols / qk * bs;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++)
        for (int64_t i = 0; i < cols / qk; i++) dequant_block(type, p + r * row_bytes + i * bs, out + r * cols + i * qk);
    return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* explicit thread count for the bench legs: torchrun exports OMP_NUM_THREADS=1 to its workers */
int orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------ q8_1 activation quantization */

/* One 32-wide group.  amax/sum follow the 5-step xor butterfly of ggml_kernel.cu:34-37:
 * after the stage with mask m, lane i holds s[i] + s[i^m]; every lane ends with the same value. */
static void q8_1_group(const float *x, int n_valid, int8_t *qs, float *d_out, float *sum_out) {
    float v[32], a[32], s[32], t[32];
    for (int i = 0; i < 32; i++) { v[i] = (i < n_valid) ? x[i] : 0.0f; a[i] = fabsf(v[i]); s[i] = v[i]; }
    for (int mask = 16; mask > 0; mask >>= 1) {
        for (int i = 0; i < 32; i++) t[i] = fmaxf(a[i], a[i ^ mask]);
        memcpy(a, t, sizeof a);
        for (int i = 0; i < 32; i++) t[i] = s[i] + s[i ^ mask];
        memcpy(s, t, sizeof s);
    }
    const float amax = a[0];
    const float d = amax / 127;
    for (int i = 0; i < 32; i++) qs[i] = (amax == 0.0f) ? 0 : (int8_t)roundf(v[i] / d);
    *d_out = d;
    *sum_out = s[0];
}

int64_t orc_q8_1_row_bytes(int64_t K) { return (K + 511) / 512 * 512 / 32 * 36; }

/* ggml_kernel.cu:13-66: rows of block_q8_1 {half d; half s; int8 qs[32]}, K padded to 512 with zeros */
int orc_quantize_q8_1(const float *x, int64_t rows, int64_t K, void *out) {
    const int64_t Kp = (K + 511) / 512 * 512;
    uint8_t *o = (uint8_t *)out;
    for (int64_t r = 0; r < rows; r++) {
        for (int64_t g = 0; g < Kp / 32; g++) {
            uint8_t *blk = o + (r * (Kp / 32) + g) * 36;
            float d, s;
            int64_t k0 = g * 32;
            int nv = (int)((K - k0) > 32 ? 32 : ((K - k0) < 0 ? 0 : (K - k0)));
            q8_1_group(x + r * K + k0, nv, (int8_t *)(blk + 4), &d, &s);
            uint16_t hd = f2h(d), hs = f2h(s);
            memcpy(blk, &hd, 2); memcpy(blk + 2, &hs, 2);
        }
    }
    return 0;
}

int64_t orc_q8_1_mmq_bytes(int64_t M, int64_t K) { return M * ((K - K % 512 + 512) / 32) * 36; }

/* mmq.cu:109-177: block_q8_1_mmq {half2 ds[4] | float d[4]; int8 qs[128]} (144 B),
 * block index (k/128)*M + row, K padded to K - K%512 + 512. */
int orc_quantize_q8_1_mmq(const float *x, int64_t M, int64_t K, int need_sum, void *out) {
    const int64_t Kp = K - K % 512 + 512;
    uint8_t *o = (uint8_t *)out;
    for (int64_t r = 0; r < M; r++) {
        for (int64_t c = 0; c < Kp / 128; c++) {
            uint8_t *blk = o + (c * M + r) * 144;
            for (int g = 0; g < 4; g++) {
                float d, s;
                int64_t k0 = c * 128 + g * 32;
                int nv = (int)((K - k0) > 32 ? 32 : ((K - k0) < 0 ? 0 : (K - k0)));
                q8_1_group(x + r * K + k0, nv, (int8_t *)(blk + 16 + 32 * g), &d, &s);
                if (need_sum) {
                    uint16_t hd = f2h(d), hs = f2h(s);
                    memcpy(blk + 4 * g, &hd, 2); memcpy(blk + 4 * g + 2, &hs, 2);
                } else {
                    memcpy(blk + 4 * g, &d, 4);
                }
            }
        }
    }
    return 0;
}

int orc_mmq_need_sum(int type) {                     /* mmq.cu:84-106 */
    return type == ORC_Q4_0 || type == ORC_Q4_1 || type == ORC_Q5_1 || type == ORC_Q4_K || type == ORC_Q5_K;
}


/* ------------------------------------------------------------------ MMQ A operand (product-defined, not reference) */

/* round-to-nearest-even double -> binary16 with ONE rounding (used where the exact value is not an fp32) */
static uint16_t d2h(double x) {
    if (x != x) return 0x7e00;
    uint16_t sign = (uint16_t)(signbit(x) ? 0x8000u : 0);
    double a = fabs(x);
    if (a == 0.0) return sign;
    if (a >= 65520.0) return (uint16_t)(sign | 0x7c00u);
    int e;
    (void)frexp(a, &e);                       /* a = m * 2^e, m in [0.5, 1)  ->  a in [2^(e-1), 2^e) */
    int ex = e - 1;
    if (ex < -14) ex = -14;                    /* subnormal quantum 2^-24 */
    const double quantum = ldexp(1.0, ex - 10);
    const double r = nearbyint(a / quantum);   /* default rounding mode: nearest-even */
    const double v = r * quantum;
    /* v is exactly representable in binary16 (or rounds up to the next binade, still representable) */
    return (uint16_t)(sign | f2h((float)v));
}

/* The fp16 A operand the MMQ kernel feeds the tensor cores (ggml-libtorch_b200/csrc/mmq_unpack.cuh):
 *   Q4_K, Q5_K: a = fp16( q * fp16(d*sc) - fp16(dmin*m) )   (one fp16 FMA: exact value, single rounding)
 *   Q6_K      : a = fp16( (q-32) * fp16(d*sc) )
 *   others    : a = fp16( dequantize(w) )                   (the fp32 dequantizer value, one RN) */
int orc_mmq_a_operand(const void *w, int type, int64_t nelem, uint16_t *out) {
    const int qk = orc_block_elems(type), bs = orc_block_bytes(type);
    if (qk == 0 || type == ORC_Q8_1) return -1;
    if (nelem % qk) return -2;
    const uint8_t *p = (const uint8_t *)w;
    const int64_t nb = nelem / qk;
    for (int64_t i = 0; i < nb; i++) {
        const uint8_t *b = p + i * bs;
        uint16_t *y = out + i * qk;
        if (type == ORC_Q4_K || type == ORC_Q5_K) {
            int q[256];
            unpack_block_ints(type, b, q);
            const float d = h2f(ld16(b)), dmin = h2f(ld16(b + 2));
            for (int j = 0; j < 8; j++) {
                int sc, mn; k4_scale_min(j, b + 4, &sc, &mn);
                const double dl = h2f(f2h(d * (float)sc)), ml = h2f(f2h(dmin * (float)mn));
                for (int l = 0; l < 32; l++) y[32 * j + l] = d2h((double)q[32 * j + l] * dl - ml);
            }
        } else if (type == ORC_Q6_K) {
            int q[256];
            unpack_block_ints(type, b, q);
            const float d = h2f(ld16(b + 208));
            for (int is = 0; is < 16; is++) {
                const double dl = h2f(f2h(d * (float)(int8_t)b[192 + is]));
                for (int l = 0; l < 16; l++) y[16 * is + l] = d2h((double)q[16 * is + l] * dl);
            }
        } else {
            float f[256];
            dequant_block(type, b, f);
            for (int l = 0; l < qk; l++) y[l] = f2h(f[l]);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ quantized dot products */

/* One weight row (K elements of `type`) against one activation row given as per-32 groups:
 * q8[k], d8[g] (already the value the kernel would read: fp16-rounded for q8_1 / need_sum
 * layouts, fp32 for the !need_sum mmq layout) and s8[g] (fp16-rounded sum of raw inputs).
 * `mmq_mode` selects the MMQ (mma path, the one an sm_100 build of the reference takes) algebra:
 * K-quant minima as dmin*m*s8 (mmq.cuh:1351-1361) instead of MMVQ's d8*sum(q8)
 * (vecdotq.cuh:280-283), and Q4_0/Q5_0 offsets applied to the quants instead of through s8.
 * Accumulation in double: this is the tolerance oracle. */
static double row_dot(int type, const uint8_t *w, int64_t K, const int8_t *q8, const float *d8,
                      const float *s8, int mmq_mode) {
    const int qk = orc_block_elems(type), bs = orc_block_bytes(type);
    double acc = 0.0;
    int q[256];
    for (int64_t ib = 0; ib < K / qk; ib++) {
        const uint8_t *b = w + ib * bs;
        const int8_t *a = q8 + ib * qk;
        const float *dd = d8 + ib * (qk / 32);
        const float *ss = s8 + ib * (qk / 32);
        if (is_iq_grid_type(type)) { acc += iq_block_dot(type, b, a, dd, ss); continue; }
        /* raw (un-offset) quants for the types whose offset is folded through s8 */
        unpack_block_ints(type, b, q);
        switch (type) {
        case ORC_Q4_0: case ORC_Q5_0: {
            const int off = (type == ORC_Q4_0) ? 8 : 16;
            const float d = h2f(ld16(b));
            int sumi = 0;
            if (mmq_mode) {                          /* mmq.cuh:333-391, 535-687: offset applied to the quants */
                for (int j = 0; j < 32; j++) sumi += q[j] * a[j];
                acc += (double)d * (double)dd[0] * sumi;
            } else {                                 /* vecdotq.cuh:45-65, 93-125: d*(sumi*d8 - off*s8) */
                for (int j = 0; j < 32; j++) sumi += (q[j] + off) * a[j];
                acc += (double)d * ((double)sumi * dd[0] - (double)off * ss[0]);
            }
            break; }
        case ORC_Q4_1: case ORC_Q5_1: {              /* vecdotq.cuh:69-91, 130-162: half products */
            int sumi = 0;
            for (int j = 0; j < 32; j++) sumi += q[j] * a[j];
            const float d = h2f(ld16(b)), m = h2f(ld16(b + 2));
            const float d4d8 = h2f(f2h(d * dd[0]));   /* __hmul2(dm, ds8): fp16 RN of the exact product */
            const float m4s8 = h2f(f2h(m * ss[0]));
            acc += (double)sumi * d4d8 + (double)m4s8;
            break; }
        case ORC_Q8_0: {                             /* vecdotq.cuh:166-178 */
            int sumi = 0;
            for (int j = 0; j < 32; j++) sumi += q[j] * a[j];
            const float d = h2f(ld16(b));
            acc += (double)(d * dd[0]) * sumi;
            break; }
        case ORC_Q2_K: {                             /* vecdotq.cuh:203-231 */
            const float d = h2f(ld16(b + 80)), dmin = h2f(ld16(b + 82));
            double sd = 0.0, sm = 0.0;
            for (int is = 0; is < 16; is++) {
                int dot = 0, sq = 0;
                for (int l = 0; l < 16; l++) { dot += q[16 * is + l] * a[16 * is + l]; sq += a[16 * is + l]; }
                sd += (double)dd[is / 2] * (dot * (b[is] & 0x0f));
                sm += (double)dd[is / 2] * (sq * (b[is] >> 4));
            }
            acc += (double)d * sd - (double)dmin * sm;
            break; }
        case ORC_Q3_K: {                             /* vecdotq.cuh:235-268 */
            const float d = h2f(ld16(b + 108));
            double sf = 0.0;
            for (int is = 0; is < 16; is++) {
                int dot = 0;
                for (int l = 0; l < 16; l++) dot += q[16 * is + l] * a[16 * is + l];
                sf += (double)dd[is / 2] * (dot * q3k_scale(b + 96, is));
            }
            acc += (double)d * sf;
            break; }
        case ORC_Q4_K: case ORC_Q5_K: {              /* vecdotq.cuh:272-327; MMQ minima via s8: mmq.cuh */
            const float d = h2f(ld16(b)), dmin = h2f(ld16(b + 2));
            double sd = 0.0, sm = 0.0;
            for (int j = 0; j < 8; j++) {
                int sc, mn; k4_scale_min(j, b + 4, &sc, &mn);
                int dot = 0, sq = 0;
                for (int l = 0; l < 32; l++) { dot += q[32 * j + l] * a[32 * j + l]; sq += a[32 * j + l]; }
                sd += (double)dd[j] * (dot * sc);
                sm += mmq_mode ? (double)ss[j] * mn : (double)dd[j] * (sq * mn);
            }
            acc += (double)d * sd - (double)dmin * sm;
            break; }
        case ORC_Q6_K: {                             /* vecdotq.cuh:331-349 */
            const float d = h2f(ld16(b + 208));
            double sf = 0.0;
            for (int is = 0; is < 16; is++) {
                int dot = 0;
                for (int l = 0; l < 16; l++) dot += q[16 * is + l] * a[16 * is + l];
                sf += (double)dd[is / 2] * (dot * (int8_t)b[192 + is]);
            }
            acc += (double)d * sf;
            break; }
        case ORC_IQ4_NL: {                           /* vecdotq.cuh:842-864 */
            int sumi = 0;
            for (int j = 0; j < 32; j++) sumi += q[j] * a[j];
            const float d = h2f(ld16(b));
            acc += (double)(d * dd[0]) * sumi;
            break; }
        case ORC_IQ4_XS: {                           /* vecdotq.cuh:867-888 */
            const float d = h2f(ld16(b));
            for (int ib = 0; ib < 8; ib++) {
                int sumi = 0;
                for (int l = 0; l < 32; l++) sumi += q[32 * ib + l] * a[32 * ib + l];
                acc += (double)d * (double)iq4xs_scale(b, ib) * (double)dd[ib] * sumi;
            }
            break; }
        default: break;
        }
    }
    return acc;
}

/* MMVQ: W [N rows of `type`, K cols] x q8_1 rows (from orc_quantize_q8_1) -> Y[B][N] fp32.
 * ggml_kernel.cu:80-193 + mmvq.cuh:2-38 (the reference allows B == 1 only). */
int orc_mmvq(const void *W, int type, const void *q8_1, int64_t B, int64_t K, int64_t N, float *Y) {
    const int qk = orc_block_elems(type), bs = orc_block_bytes(type);
    if (qk == 0 || type == ORC_Q8_1) return -1;
    if (K % qk) return -2;
    const int64_t Kp = (K + 511) / 512 * 512, ng = Kp / 32;
    int8_t *q8 = (int8_t *)malloc((size_t)Kp);
    float *d8 = (float *)malloc(sizeof(float) * (size_t)ng), *s8 = (float *)malloc(sizeof(float) * (size_t)ng);
    if (!q8 || !d8 || !s8) { free(q8); free(d8); free(s8); return -3; }
    const uint8_t *w = (const uint8_t *)W;
    for (int64_t bi = 0; bi < B; bi++) {
        const uint8_t *row = (const uint8_t *)q8_1 + bi * ng * 36;
        for (int64_t g = 0; g < ng; g++) {
            d8[g] = h2f(ld16(row + g * 36)); s8[g] = h2f(ld16(row + g * 36 + 2));
            memcpy(q8 + g * 32, row + g * 36 + 4, 32);
        }
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < N; n++)
            Y[bi * N + n] = (float)row_dot(type, w + n * (K / qk) * bs, K, q8, d8, s8, 0);
    }
    free(q8); free(d8); free(s8);
    return 0;
}

/* MMQ: W [N, K] x q8_1_mmq blocks (from orc_quantize_q8_1_mmq with need_sum = orc_mmq_need_sum(type))
 * -> Y[M][N] fp32, dst[j*N+i] as mmq.cuh:1764-1790. */
int orc_mmq(const void *W, int type, const void *q8_1_mmq, int64_t M, int64_t K, int64_t N, float *Y) {
    const int qk = orc_block_elems(type), bs = orc_block_bytes(type);
    if (qk == 0 || type == ORC_Q8_1 || type == ORC_IQ4_NL || type == ORC_IQ4_XS || is_iq_grid_type(type)) return -1;   /* mmq.cu:222-252: no IQ case */
    if (K % qk) return -2;
    const int need_sum = orc_mmq_need_sum(type);
    const int64_t Kp = K - K % 512 + 512, ng = Kp / 32;
    int8_t *q8 = (int8_t *)malloc((size_t)Kp);
    float *d8 = (float *)malloc(sizeof(float) * (size_t)ng), *s8 = (float *)calloc((size_t)ng, sizeof(float));
    if (!q8 || !d8 || !s8) { free(q8); free(d8); free(s8); return -3; }
    const uint8_t *w = (const uint8_t *)W;
    for (int64_t j = 0; j < M; j++) {
        for (int64_t c = 0; c < Kp / 128; c++) {
            const uint8_t *blk = (const uint8_t *)q8_1_mmq + (c * M + j) * 144;
            for (int g = 0; g < 4; g++) {
                if (need_sum) { d8[4 * c + g] = h2f(ld16(blk + 4 * g)); s8[4 * c + g] = h2f(ld16(blk + 4 * g + 2)); }
                else          { memcpy(&d8[4 * c + g], blk + 4 * g, 4); }
            }
            memcpy(q8 + c * 128, blk + 16, 128);
        }
#pragma omp parallel for schedule(static)
        for (int64_t n = 0; n < N; n++)
            Y[j * N + n] = (float)row_dot(type, w + n * (K / qk) * bs, K, q8, d8, s8, 1);
    }
    free(q8); free(d8); free(s8);
    return 0;
}
