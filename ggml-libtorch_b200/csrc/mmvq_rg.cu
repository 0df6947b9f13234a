// mmvq_rg.cu -- batch-1 MMVQ with the q8_1 activations RESIDENT IN REGISTERS (Q4_K: K % 256 == 0, K <= 16384, 24576, 28672;
// Q5_K: the same; Q6_K: K % 2048 == 0, the same row lengths; Q3_K: K % 2048 == 0, K <= 24576; Q2_K: K % 1024 == 0, K <= 24576).
//
// Same contract as the other MMVQ kernels (ggml_mul_mat_vec_a8 = quantize_row_q8_1_cuda + mul_mat_vec_q + vec_dot_q4_K_q8_1,
// hf-kernels/ggml-kernels/ggml/ggml_kernel.cu:13-66,80-193; ggml/mmvq.cuh:2-38; ggml/vecdotq.cuh:266-327).
//
// Why a second kernel next to mmvq_sb_body.cuh.  Round-1 measurements (profiles/r01_*) showed the half-SM ring
// kernel CONSUMER-bound, not HBM-bound: its DRAM traffic equals the algorithmic bytes, but eight warps per SM at ~1
// instruction per weight value (of which a quarter are shared-memory loads of the activations) run the stream phase at
// just the HBM rate, so every microsecond the launch spends waiting for its dependency and quantizing X (2-3 us of an
// 8.5 us period, per-CTA globaltimer traces) is lost.  Here the consumer is made ~2x faster than HBM so that the rings
// pre-filled during that serial part are drained at shared-memory speed and the remainder streams at HBM speed:
//   * a lane owns ONE 128-value item position of a row chunk (32 items = 4096 values) and keeps that item's q8_1
//     activations (32 words of quants + 4 d8 + 4 d8*sum(q8)) in REGISTERS for the whole launch: no activation loads in the
//     main loop (5 LDS.128 per 128 weights instead of 15), ~0.6 instructions per weight value;
//   * 12 warps in a half-SM CTA (<= 80 registers, <= 113 KB): three warps per scheduler, and still two CTAs per SM so that
//     the NEXT launch (programmatic dependent launch) is resident and pre-fills its rings while this one runs;
//   * a slot is one row chunk (<= 2304 B): finer work granularity (shorter tail), per-warp self-refilling rings as before;
//   * rows longer than 4096 values are split into chunks owned by different warps (warp w: chunk w % C); the chunk
//     partials of a row meet in shared memory and are added in chunk order at the end, so the summation order depends on
//     K only (N-sharded results stay bit-identical);
//   * four rows are reduced at once (6 shuffles per 4 rows instead of 20).
// The activation quantization itself is unchanged (q8_1.cuh: bit-identical to quantize_q8_1).
#include <type_traits>

#include "q8_1.cuh"

namespace ggb_sb {
extern int g_pdl_prefetch;
int g_rg_mode = 1;           // ggb_set_option(GGB_OPT_MMVQ_RG): bit 0 = use this kernel; bits 8.. = debug (1 = skip quantization); bits 12..15 stages, 16..19 early stages, 20..23 cluster size, 24..25 cluster policy, 26 plain launches, 27..31 one type each back on its previous kernel
}
unsigned long long *ggb_mmvq_trace_slot();

namespace {

// warps per CTA: 12 (rows of 1, 2, 3, 4 or 6 chunks), or 14 for the seven chunks of K = 28672 (the 70B down projection; 72 registers,
// no spills)
constexpr int kDefaultCluster = 2;       // CTAs sharing one activation quantization (measured)
constexpr int kChunkItems = 32;          // items (128 values) per row chunk = one per lane

struct RgParams {
    const uint8_t *W;
    const void *X;
    GgbXIn xin;
    GgbYOut yo;
    int64_t N;
    int K, nsb, ipr, row_bytes;
    int C;               // chunks per row
    int RW;              // row lanes per CTA = NW / C
    int off_scales, off_part, off_bar, off_slots;
    int spc;             // super-blocks per chunk
    int early;           // rows per warp requested before the dependency wait (static weights), 1..kStages
    long long wstep;     // bytes between two consecutive rows of a warp = RW * row_bytes
    const uint8_t *image; // != nullptr: the q8_1 activation image written by the previous launch (mmvq_sb_prequant_kernel,
                          // [8][ipr] quants at 0, [2][ipr] scale table at off_scales): every lane loads ITS item from it
    int part_rows;       // rows the partial-sum area holds (C > 1)
    int x_vec;
    int pdl_prefetch;
    int dbg;
    int share;           // the CTAs of a cluster share the activation quantization (else a cluster launch only places the CTAs)
    unsigned long long *trace;
};

__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c) { return __dp4a((int)a, (int)b, c); }
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// exact (float)(a * b) for |a * b| < 2^22
__device__ __forceinline__ float imul_to_f32(int a, int b) { return __int_as_float(a * b + 0x4B400000) - 12582912.0f; }
template <int J>
__device__ __forceinline__ float byte_to_f32(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + J)) - 8388608.0f; }
template <int J>
__device__ __forceinline__ int byte_of(uint32_t w) { return (int)__byte_perm(w, 0, 0x4440 + J); }
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// ---- thread-block cluster helpers: the CTAs of a cluster share the activation quantization through distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// asynchronous remote stores that count their bytes on an mbarrier of the DESTINATION CTA: the reader waits on its own
// barrier for the image's byte count.  (A barrier.cluster with release semantics instead also waits for the stores of the
// ~100 KB of cp.async weight requests in flight: measured 2400-3900 cycles per launch.)
__device__ __forceinline__ void st_async_v2(uint32_t addr, uint32_t x, uint32_t y, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                 ::"r"(addr), "r"(x), "r"(y), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_f32(uint32_t addr, float x, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(addr), "r"(__float_as_uint(x)), "r"(mbar) : "memory");
}

template <int TYPE> struct RgOps;

// ---- Q4_K: one 128-value item (half h of a 144-byte super-block) against the lane's resident activations
template <> struct RgOps<T_Q4_K> {
    static constexpr int BS = 144;
    // second activation word per 32-value group: d8 * sum(q8) (the minima's factor)
    __device__ __forceinline__ static uint32_t aux_word(float d8, int s_lo16, int s_hi16) { return __float_as_uint(d8 * (float)(s_lo16 + s_hi16)); }
    // hdp: the lane's super-block header in the ring stage (the chunk's bytes as they lie in memory); qsp: the 64 bytes of
    // nibbles of its half h (32-bit shared-memory addresses)
    __device__ __forceinline__ static float item(uint32_t hdp, uint32_t qsp, const uint32_t (&a)[32], const float4 d8,
                                                 const float4 ds, int h) {
        const uint4 hd = lds128(hdp);
        const uint4 q0 = lds128(qsp), q1 = lds128(qsp + 16), q2 = lds128(qsp + 32), q3 = lds128(qsp + 48);
        const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&hd.x));
        // (scale, min) of the item's four 32-value groups, one per byte (dequantize.cuh:154-161)
        uint32_t sc, mn;
        if (h) {
            sc = (hd.w & 0x0f0f0f0fu) | ((hd.y >> 2) & 0x30303030u);
            mn = ((hd.w >> 4) & 0x0f0f0f0fu) | ((hd.z >> 2) & 0x30303030u);
        } else {
            sc = hd.y & 0x3f3f3f3fu;
            mn = hd.z & 0x3f3f3f3fu;
        }
        // bytes 0..31 of the item: low nibbles = group 0 (a[0..7]), high nibbles = group 1 (a[8..15]); bytes 32..63: groups 2, 3
        const uint32_t qa[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        const uint32_t qb[8] = {q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
        int g0 = 0, g1 = 0, g2 = 0, g3 = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            g0 = dp4a_ss(qa[i] & 0x0f0f0f0fu, a[i], g0);
            g1 = dp4a_us(qa[i] & 0xf0f0f0f0u, a[8 + i], g1);               // 16 x the high nibbles
            g2 = dp4a_ss(qb[i] & 0x0f0f0f0fu, a[16 + i], g2);
            g3 = dp4a_us(qb[i] & 0xf0f0f0f0u, a[24 + i], g3);
        }
        float fd = d8.x * imul_to_f32(g0, byte_of<0>(sc));
        fd = fmaf(d8.y, imul_to_f32(g1 >> 4, byte_of<1>(sc)), fd);
        fd = fmaf(d8.z, imul_to_f32(g2, byte_of<2>(sc)), fd);
        fd = fmaf(d8.w, imul_to_f32(g3 >> 4, byte_of<3>(sc)), fd);
        // minima: dmin * sum_g m_g * (d8_g * sum q8_g)    (vecdotq.cuh:280-288)
        float t = byte_to_f32<0>(mn) * ds.x;
        t = fmaf(byte_to_f32<1>(mn), ds.y, t);
        t = fmaf(byte_to_f32<2>(mn), ds.z, t);
        t = fmaf(byte_to_f32<3>(mn), ds.w, t);
        return dm.x * fd - dm.y * t;
    }
};

// ---- Q5_K: {d, dmin, 12 scale bytes} @0 (the Q4_K header), 32 bytes of fifth bits @16 (bit j of byte l = value l of group j),
// 128 bytes of nibbles @48 (dequantize.cuh:204-227, vecdotq.cuh:299-327).  176-byte super-blocks are 16-byte aligned.  The item =
// half h = groups 4h .. 4h + 3.  The fifth bits never meet the nibbles: group jj's bit plane is taken out of the (pre-shifted) qh
// word with the mask 0x01010101 << jj and dotted separately -- dp4a gives 2^jj * sum(bit * q8), one shift per group restores 16 x.
template <> struct RgOps<T_Q5_K> {
    static constexpr int BS = 176;
    __device__ __forceinline__ static uint32_t aux_word(float d8, int s_lo16, int s_hi16) { return __float_as_uint(d8 * (float)(s_lo16 + s_hi16)); }
    __device__ __forceinline__ static float item(uint32_t hp, const uint32_t (&a)[32], const float4 d8, const float4 ds, int h) {
        const uint4 hd = lds128(hp);
        const uint4 h0 = lds128(hp + 16), h1 = lds128(hp + 32);
        const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&hd.x));
        uint32_t sc, mn;
        if (h) {
            sc = (hd.w & 0x0f0f0f0fu) | ((hd.y >> 2) & 0x30303030u);
            mn = ((hd.w >> 4) & 0x0f0f0f0fu) | ((hd.z >> 2) & 0x30303030u);
        } else {
            sc = hd.y & 0x3f3f3f3fu;
            mn = hd.z & 0x3f3f3f3fu;
        }
        const uint32_t sh = 4u * (uint32_t)h;
        const uint32_t qh[8] = {h0.x >> sh, h0.y >> sh, h0.z >> sh, h0.w >> sh, h1.x >> sh, h1.y >> sh, h1.z >> sh, h1.w >> sh};
        int g0 = 0, g1 = 0, g2 = 0, g3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        const uint32_t qsp = hp + 48 + 64 * h;
        {
            const uint4 q0 = lds128(qsp), q1 = lds128(qsp + 16);
            const uint32_t qa[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                g0 = dp4a_ss(qa[i] & 0x0f0f0f0fu, a[i], g0);
                g1 = dp4a_us(qa[i] & 0xf0f0f0f0u, a[8 + i], g1);               // 16 x the high nibbles
                b0 = dp4a_ss(qh[i] & 0x01010101u, a[i], b0);
                b1 = dp4a_ss(qh[i] & 0x02020202u, a[8 + i], b1);              // 2 x the bits
            }
        }
        {
            const uint4 q2 = lds128(qsp + 32), q3 = lds128(qsp + 48);
            const uint32_t qb[8] = {q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                g2 = dp4a_ss(qb[i] & 0x0f0f0f0fu, a[16 + i], g2);
                g3 = dp4a_us(qb[i] & 0xf0f0f0f0u, a[24 + i], g3);
                b2 = dp4a_ss(qh[i] & 0x04040404u, a[16 + i], b2);             // 4 x
                b3 = dp4a_ss(qh[i] & 0x08080808u, a[24 + i], b3);             // 8 x
            }
        }
        // |dot * sc| reaches 32 * 31 * 127 * 63 = 7.9 M: past the 2^22 of the mantissa trick, so the conversion instruction
        float fd = d8.x * (float)((g0 + (b0 << 4)) * byte_of<0>(sc));
        fd = fmaf(d8.y, (float)(((g1 >> 4) + (b1 << 3)) * byte_of<1>(sc)), fd);
        fd = fmaf(d8.z, (float)((g2 + (b2 << 2)) * byte_of<2>(sc)), fd);
        fd = fmaf(d8.w, (float)(((g3 >> 4) + (b3 << 1)) * byte_of<3>(sc)), fd);
        float t = byte_to_f32<0>(mn) * ds.x;
        t = fmaf(byte_to_f32<1>(mn), ds.y, t);
        t = fmaf(byte_to_f32<2>(mn), ds.z, t);
        t = fmaf(byte_to_f32<3>(mn), ds.w, t);
        return dm.x * fd - dm.y * t;
    }
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
// N consecutive 32-bit words from a 2-byte-aligned shared address: N + 1 aligned loads + funnel shifts (sh = 0 or 16)
template <int N>
__device__ __forceinline__ void lds_words(uint32_t addr_al, uint32_t sh, uint32_t (&out)[N]) {
    uint32_t w[N + 1];
#pragma unroll
    for (int i = 0; i <= N; i++) w[i] = lds32(addr_al + 4 * i);
#pragma unroll
    for (int i = 0; i < N; i++) out[i] = __funnelshift_r(w[i], w[i + 1], sh);
}

// ---- Q6_K: ql[128] @0, qh[64] @128, 16 int8 scales @192 (one per 16 values), d @208; 210-byte super-blocks are 2-byte aligned
// (in the stage: 4-byte aligned for even super-blocks, 2 bytes off for odd ones).  Value = d * sc * (q6 - 32)
// (dequantize.cuh:230-254, vecdotq.cuh:331-349).  The item = half h: values 128 h .. 128 h + 127 = four groups j of 32:
//   j = 0: ql[64h + l] & 15 | (qh[32h + l] & 3) << 4        j = 1: ql[64h + 32 + l] & 15 | (qh >> 2 & 3) << 4
//   j = 2: ql[64h + l] >> 4 | (qh >> 4 & 3) << 4            j = 3: ql[64h + 32 + l] >> 4 | (qh >> 6 & 3) << 4      (l = 0..31)
// with scale byte 8h + 2j + l / 16.  The weights go to dp4a as the unsigned q6; the -32 comes back through the sums of the q8
// quants over each 16-value half (second activation word: two int16), all in integers: t = sum_u sc_u * (dot_u - 32 * sq_u).
template <> struct RgOps<T_Q6_K> {
    static constexpr int BS = 210;
    __device__ __forceinline__ static uint32_t aux_word(float, int s_lo16, int s_hi16) { return ((uint32_t)s_lo16 & 0xffffu) | ((uint32_t)s_hi16 << 16); }
    __device__ __forceinline__ static float item(uint32_t hp, const uint32_t (&a)[32], const float4 d8, const float4 ds, int h) {
        const uint32_t sh = (hp & 2u) * 8u, al = hp & ~3u;
        uint32_t qh[8], scw[2];
        lds_words<8>(al + 128 + 32 * h, sh, qh);
        lds_words<2>(al + 192 + 8 * h, sh, scw);
        const float d = ggb_h2f((uint16_t)lds16(hp + 208));
        int dot[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        {
            uint32_t ql[8];
            lds_words<8>(al + 64 * h, sh, ql);
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t q0 = (ql[w] & 0x0f0f0f0fu) | ((qh[w] << 4) & 0x30303030u);
                const uint32_t q2 = ((ql[w] >> 4) & 0x0f0f0f0fu) | (qh[w] & 0x30303030u);
                dot[0][w >> 2] = dp4a_us(q0, a[w], dot[0][w >> 2]);
                dot[2][w >> 2] = dp4a_us(q2, a[16 + w], dot[2][w >> 2]);
            }
        }
        {
            uint32_t ql[8];
            lds_words<8>(al + 64 * h + 32, sh, ql);
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t q1 = (ql[w] & 0x0f0f0f0fu) | ((qh[w] << 2) & 0x30303030u);
                const uint32_t q3 = ((ql[w] >> 4) & 0x0f0f0f0fu) | ((qh[w] >> 2) & 0x30303030u);
                dot[1][w >> 2] = dp4a_us(q1, a[8 + w], dot[1][w >> 2]);
                dot[3][w >> 2] = dp4a_us(q3, a[24 + w], dot[3][w >> 2]);
            }
        }
        const float d8v[4] = {d8.x, d8.y, d8.z, d8.w};
        const uint32_t sqv[4] = {__float_as_uint(ds.x), __float_as_uint(ds.y), __float_as_uint(ds.z), __float_as_uint(ds.w)};
        float f = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t sw = scw[j >> 1] >> (16 * (j & 1));
            const int s0 = (int)(int8_t)(sw & 0xffu), s1 = (int)(int8_t)((sw >> 8) & 0xffu);
            const int sq0 = (int)(int16_t)(sqv[j] & 0xffffu), sq1 = (int)(int16_t)(sqv[j] >> 16);
            const int t = (dot[j][0] - 32 * sq0) * s0 + (dot[j][1] - 32 * sq1) * s1;
            f = fmaf(d8v[j], (float)t, f);
        }
        return d * f;
    }
};

// ---- Q3_K: hmask[32] @0, qs[64] @32 (2-bit), 12 scale bytes @96 (sixteen 6-bit scales, one per 16 values), d @108; 110-byte
// super-blocks, 2-byte aligned like Q6_K.  Value = d * (sc - 32) * (q2 + 4 * hbit - 4)   (dequantize.cuh:129-151,
// vecdotq.cuh:235-268).  The item = half n: qs bytes 32 n .. 32 n + 31 hold four groups j at bit 2 j, hmask bit 4 n + j is the
// group's high bit.  Neither is ever shifted into place: the group's plane is masked out where it lies (0x03030303 << 2 j of
// the qs word, 0x01010101 << j of the pre-shifted hmask word), dp4a returns 4^j * sum(q2 * q8) and 2^j * sum(hbit * q8), and
// one shift per 16-value sub-block undoes the factors; the -4 goes through the integer half-sums of the q8 quants.
template <> struct RgOps<T_Q3_K> {
    static constexpr int BS = 110;
    __device__ __forceinline__ static uint32_t aux_word(float, int s_lo16, int s_hi16) { return ((uint32_t)s_lo16 & 0xffffu) | ((uint32_t)s_hi16 << 16); }
    __device__ __forceinline__ static float item(uint32_t hp, const uint32_t (&a)[32], const float4 d8, const float4 ds, int h) {
        const uint32_t sh = (hp & 2u) * 8u, al = hp & ~3u;
        uint32_t S[3];
        lds_words<3>(al + 96, sh, S);
        const float d = ggb_h2f((uint16_t)lds16(hp + 108));
        // the item's eight 6-bit scales, one per byte (before the -32): k = 0..3 in scA, 4..7 in scB (gguf-py quants.py:442-454)
        const uint32_t s4 = 4u * (uint32_t)h;
        const uint32_t scA = ((S[0] >> s4) & 0x0f0f0f0fu) | (((S[2] >> s4) & 0x03030303u) << 4);
        const uint32_t scB = ((S[1] >> s4) & 0x0f0f0f0fu) | (((S[2] >> (s4 + 2)) & 0x03030303u) << 4);
        const uint32_t sqv[4] = {__float_as_uint(ds.x), __float_as_uint(ds.y), __float_as_uint(ds.z), __float_as_uint(ds.w)};
        int t[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 2; u++) {                               // the 16-value halves of the four groups: words 4u .. 4u + 3
            uint32_t ql[4], hm[4];
            lds_words<4>(al + 32 + 32 * h + 16 * u, sh, ql);
            lds_words<4>(al + 16 * u, sh, hm);
            int gq[4] = {0, 0, 0, 0}, gb[4] = {0, 0, 0, 0};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint32_t hw = hm[w] >> s4;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    gq[j] = dp4a_us(ql[w] & (0x03030303u << (2 * j)), a[8 * j + 4 * u + w], gq[j]);
                    gb[j] = dp4a_ss(hw & (0x01010101u << j), a[8 * j + 4 * u + w], gb[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = 2 * j + u;                            // sub-block inside the item
                const int sc = (int)(((k < 4 ? scA : scB) >> (8 * (k & 3))) & 0xffu) - 32;
                const int sq = u ? (int)(int16_t)(sqv[j] >> 16) : (int)(int16_t)(sqv[j] & 0xffffu);
                const int dot = (gq[j] >> (2 * j)) + ((gb[j] >> j) << 2) - 4 * sq;
                t[j] += sc * dot;
            }
        }
        float f = d8.x * (float)t[0];
        f = fmaf(d8.y, (float)t[1], f);
        f = fmaf(d8.z, (float)t[2], f);
        f = fmaf(d8.w, (float)t[3], f);
        return d * f;
    }
};

// ---- Q2_K: 16 scale bytes @0 (4-bit scale | 4-bit minimum, one per 16 values), qs[64] @16 (2-bit), d @80, dmin @82; 84-byte
// super-blocks are 4-byte aligned.  Value = d * sc * q2 - dmin * m   (dequantize.cuh:107-120, vecdotq.cuh:203-231).  The item = half n:
// qs bytes 32 n .. 32 n + 31, group j at bit 2 j (masked in place: dp4a returns 4^j * sum(q2 * q8)), scale byte 8 n + 2 j + u for
// the 16-value half u; the minima go through the integer half-sums of the q8 quants.
template <> struct RgOps<T_Q2_K> {
    static constexpr int BS = 84;
    __device__ __forceinline__ static uint32_t aux_word(float, int s_lo16, int s_hi16) { return ((uint32_t)s_lo16 & 0xffffu) | ((uint32_t)s_hi16 << 16); }
    __device__ __forceinline__ static float item(uint32_t hp, const uint32_t (&a)[32], const float4 d8, const float4 ds, int h) {
        const uint32_t S0 = lds32(hp + 8 * h), S1 = lds32(hp + 8 * h + 4);
        const uint32_t dmw = lds32(hp + 80);
        const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&dmw));
        const uint32_t sqv[4] = {__float_as_uint(ds.x), __float_as_uint(ds.y), __float_as_uint(ds.z), __float_as_uint(ds.w)};
        int ts[4] = {0, 0, 0, 0}, tm[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 2; u++) {
            uint32_t ql[4];
#pragma unroll
            for (int w = 0; w < 4; w++) ql[w] = lds32(hp + 16 + 32 * h + 16 * u + 4 * w);
            int gq[4] = {0, 0, 0, 0};
#pragma unroll
            for (int w = 0; w < 4; w++)
#pragma unroll
                for (int j = 0; j < 4; j++) gq[j] = dp4a_us(ql[w] & (0x03030303u << (2 * j)), a[8 * j + 4 * u + w], gq[j]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int k = 2 * j + u;
                const uint32_t sb = ((k < 4 ? S0 : S1) >> (8 * (k & 3))) & 0xffu;
                const int sq = u ? (int)(int16_t)(sqv[j] >> 16) : (int)(int16_t)(sqv[j] & 0xffffu);
                ts[j] += (int)(sb & 0xfu) * (gq[j] >> (2 * j));
                tm[j] += (int)(sb >> 4) * sq;
            }
        }
        float fs = d8.x * (float)ts[0], fm = d8.x * (float)tm[0];
        fs = fmaf(d8.y, (float)ts[1], fs); fm = fmaf(d8.y, (float)tm[1], fm);
        fs = fmaf(d8.z, (float)ts[2], fs); fm = fmaf(d8.z, (float)tm[2], fm);
        fs = fmaf(d8.w, (float)ts[3], fs); fm = fmaf(d8.w, (float)tm[3], fm);
        return dm.x * fs - dm.y * fm;
    }
};

// 16-byte asynchronous copy global -> shared (LDGSTS), L2 only, evict-first
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void *gsrc, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "l"(policy) : "memory");
}
// the same, predicated on (have > need): a branch around an LDGSTS costs ptxas 12.9 three padding instructions and a BSSY/BSYNC pair
__device__ __forceinline__ void cp_async16_if(int have, int need, uint32_t smem_dst, const void *gsrc, uint64_t policy) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %3, %4;\n\t"
                 "@p cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n\t}"
                 ::"r"(smem_dst), "l"(gsrc), "l"(policy), "r"(have), "r"(need) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ long long clk64() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}

// Weight pipeline.  A warp owns chunk c (<= 16 super-blocks = <= 2304 contiguous bytes) of rows rl, rl + RW, ... and keeps
// S of those row chunks in flight in a private ring.  The copy is COALESCED: the warp moves the chunk as consecutive
// 16-byte pieces (lane l: pieces l, l + 32, ...; cp.async.cg = LDGSTS, no registers, no L1), the image in shared memory is
// the chunk's bytes as they are, and every lane then reads ITS item (header + 64 bytes of nibbles, five conflict-free
// LDS.128: super-blocks are 144 = 128 + 16 bytes apart).  cp.async groups are per thread: wait_group + __syncwarp make
// the whole chunk visible to the warp; there is no CTA-wide synchronisation in the main loop.
// (Measured first, scripts/probes/stream_probe.cu: a cp.async.bulk per 2304-byte chunk caps the stream at ~5.1 TB/s -- the
// TMA unit takes ~115 cycles per request whatever its size -- while coalesced 16-byte-per-lane loads reach 7.1 TB/s; and
// lane-private pieces (every lane copying its own 5 x 16 bytes) touch every 32-byte sector twice: 12.4 us instead of 8.)
// S = row chunks in flight per warp (ring stages): 4 where the ring fits beside the activations in half an SM's shared
// memory (K = 4096: 12 warps x 4 x 2304 B + 5 KB = exactly the 113 KB two resident CTAs may have each), else 3.
// FULL = every chunk has 16 super-blocks (K % 4096 == 0): all lanes own an item and the copy needs no bounds.
// The activation image as its own small launch, in THIS kernel's layout ([8][ipr] quants | [2][ipr] {d8 x 4 | aux x 4}): used where
// the in-kernel pass is too slow for the warps the GEMV has (Q6_K, K = 28672 with the fused SwiGLU: 7 x 2 warps, 28672 exponentials).
// 128 threads = one warp per register bank at <= 40 registers: a CTA of it fits beside the GEMV's CTA and its neighbour's on every
// SM (see "Registers" below), so it does not keep the GEMV's CTAs from becoming resident early.
template <int TYPE, typename T>
__global__ void __launch_bounds__(128)
mmvq_rg_prequant_kernel(const T *__restrict__ X, const GgbXIn xin, const int K, const int x_vec, uint8_t *__restrict__ image, const int off_scales) {
    using Ops = RgOps<TYPE>;
    ggb_pdl_launch_dependents();
    ggb_pdl_wait();
    const int ipr = K >> 7, ngroups = K >> 5;
    const int t4 = threadIdx.x & 3;
    const int g = blockIdx.x * 32 + (threadIdx.x >> 2);
    float v[8];
    if (g < ngroups) {
        load_act8<T>(X, static_cast<const T *>(xin.x2), xin.ldx, 0, g * 32 + t4 * 8, x_vec != 0, v);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 0.0f;
    }
    const Q8Group q = q8_1_quantize_group<false>(v);
    int mine = __dp4a((int)q.q_lo, 0x01010101, __dp4a((int)q.q_hi, 0x01010101, 0));
    mine += __shfl_xor_sync(0xffffffffu, mine, 1);
    const int other = __shfl_xor_sync(0xffffffffu, mine, 2);
    if (g < ngroups) {
        const int it = g >> 2, j = g & 3, i = 2 * j + (t4 >> 1);
        const float d8 = __half2float(__float2half_rn(q.d));
        uint4 *act_q = reinterpret_cast<uint4 *>(image);
        uint4 *act_s = reinterpret_cast<uint4 *>(image + off_scales);
        *reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(act_q + i * ipr + it) + 8 * (t4 & 1)) = make_uint2(q.q_lo, q.q_hi);
        if (t4 == 0) {
            reinterpret_cast<float *>(act_s + it)[j] = d8;
            reinterpret_cast<uint32_t *>(act_s + ipr + it)[j] = Ops::aux_word(d8, mine, other);
        }
    }
}

// Registers.  The register file is split over the four SM sub-partitions (16 K registers each) and the warps of a CTA are dealt
// round-robin: a 12-warp CTA (3 warps per sub-partition x 80 registers = 7680) leaves room for 8704, a 14-warp CTA puts 4 warps on
// two of them -- at 72 registers (9216) the two kinds could NOT share an SM, although two 12-warp or two 14-warp CTAs can: in the
// 70B block (gate|up: 12 warps, down: 14) every SM then held two CTAs of one launch or none, the gate|up launch ran on 74 SMs and
// took 49 us instead of 41 (%smid stamps, profiles/r02_block_trace.log).  The 14-warp variant is therefore compiled for 64
// registers (bounds of 512 threads x 2; 24 bytes of spill stores outside the dot products).
template <int TYPE, typename T, int S, bool FULL, int NW>
__global__ void __launch_bounds__(NW == 14 ? 512 : NW * 32, 2)
mmvq_rg_kernel(const RgParams p) {
    using Ops = RgOps<TYPE>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint4 *act_q = reinterpret_cast<uint4 *>(smem);                            // [8][ipr]: 16-value chunk i of item it at i * ipr + it
    float4 *act_s = reinterpret_cast<float4 *>(smem + p.off_scales);           // [2][ipr]: d8 of the item's 4 groups | d8 * sum(q8)
    float *part_sm = reinterpret_cast<float *>(smem + p.off_part);             // [part_rows][C] chunk partials (C > 1)
    uint8_t *ring = smem + p.off_slots;                                        // [warp][stage][stage_bytes]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ipr = p.ipr, C = p.C, RW = p.RW;

    const int64_t r_begin = p.N * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = p.N * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int nrows = (int)(r_end - r_begin);

    unsigned long long *tr = p.trace ? p.trace + blockIdx.x * 8 : nullptr;
    if (tr && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[7] = (gtime() & ~0xffULL) | smid;                          // (the SM in the low byte: which CTAs share one)
        tr[0] = clk64();
    }
    ggb_pdl_launch_dependents();
    // Cluster of CL CTAs (launch attribute; CL = 1: none): each CTA quantizes 1 / CL of the activation row and writes its part
    // into every CTA's shared memory (DSMEM), so the row is read from L2 and quantized CL times less often per SM.  The
    // arrive here / wait before the first remote store proves that every CTA of the cluster has started.
    const uint32_t CL = p.share ? cluster_nctarank() : 1u, crank = CL > 1 ? cluster_ctarank() : 0u;
    uint64_t *act_bar = reinterpret_cast<uint64_t *>(smem + p.off_bar);        // counts the bytes of the image as they land (CL > 1)
    if (CL > 1) {
        if (tid == 0) {
            mbar_init(act_bar, 1);
            fence_mbar_init();
            mbar_expect_tx(act_bar, (uint32_t)(p.K + p.K / 4));                // quants + (d8, d8 * sum q8) per 32-value group
        }
        __syncthreads();
        cluster_arrive();          // (before any weight request is in flight: nothing for the release to wait for)
    }

    // warp -> (chunk c, row lane rl): rows rl, rl + RW, ... of this CTA's range, chunk c of each
    const int c = warp % C, rl = warp / C;
    const int my_rows = (rl < RW && nrows > rl) ? (nrows - rl + RW - 1) / RW : 0;
    const int sb0 = c * p.spc;                                   // first super-block of the chunk
    const int nsb_c = FULL ? 16 : min(p.spc, p.nsb - sb0);       // super-blocks in it
    const int n16 = nsb_c * Ops::BS / 16;                        // 16-byte pieces per row chunk (the launcher keeps chunks whole)
    const int sbl = lane >> 1, h = lane & 1;                     // this lane's item: half h of super-block sb0 + sbl
    const bool live = FULL || sbl < nsb_c;
    const int item = 2 * (sb0 + sbl) + h;
    // a stage is always 16 super-blocks wide, whatever the chunk holds: every shared-memory offset is an immediate.  (With a
    // run-time stage size ptxas 12.9 encoded the LDGSTS whose shared address is register + uniform register with an
    // unset uniform register and descriptor: illegal instruction at run time.)
    constexpr int stage_bytes = 16 * Ops::BS;
    uint8_t *my_ring = ring + (size_t)warp * S * stage_bytes;
    const uint32_t my_ring_s = smem_u32(my_ring) + lane * 16;    // (32-bit shared address computed once: with the generic
                                                                 // -> shared conversion inside the copy loop ptxas 12.9 emitted an
                                                                 // LDGSTS with a bad descriptor register: illegal instruction)
    const uint64_t pol = l2_policy_evict_first();

    // this lane's pieces of a row chunk: lane, lane + 32, ... (npc of them).  The source advances by one row per request.
    const int npc = (n16 - lane + 31) >> 5;
    const uint8_t *src_next = p.W + (r_begin + rl) * (int64_t)p.row_bytes + (int64_t)sb0 * Ops::BS + lane * 16;
    int to_issue = my_rows;
    constexpr int NP16 = stage_bytes / 16;                       // 16-byte pieces of a full row chunk (Q4_K: 144, Q6_K: 210)
    constexpr int NPQ = (NP16 + 31) / 32;                        // ... per lane, rounded up
    constexpr int LAST_LANES = NP16 - 32 * (NPQ - 1);            // lanes that own a piece in the last round (16 / 18)
    const int lo_last = lane < LAST_LANES;
    auto issue_row = [&](uint32_t dst) {          // this warp's next row -> the stage at dst (an empty group past the end)
        const int go = to_issue > 0;
        if (FULL) {
#pragma unroll
            for (int q = 0; q < NPQ - 1; q++) cp_async16_if(go, 0, dst + 512 * q, src_next + 512 * q, pol);
            cp_async16_if(go & lo_last, 0, dst + 512 * (NPQ - 1), src_next + 512 * (NPQ - 1), pol);
        } else {
            const int have = go ? npc : 0;
#pragma unroll
            for (int q = 0; q < NPQ; q++) cp_async16_if(have, q, dst + 512 * q, src_next + 512 * q, pol);
        }
        cp_async_commit();
        src_next += p.wstep;
        to_issue--;
    };
    // static weights: the first rows are requested while the previous kernel of the stream still runs (only p.early of the
    // S stages: the other CTA of this SM -- the PREVIOUS launch -- is about to read its activations through the same
    // load/store unit, and its requests queue behind these)
    const int early = p.pdl_prefetch ? p.early : 0;
#pragma unroll
    for (int s = 0; s < S; s++)
        if (s < early) issue_row(my_ring_s + s * stage_bytes);
    ggb_pdl_wait();
    if (tr && tid == 0) tr[1] = clk64();

    // ===================== q8_1-quantize X into shared memory (bit-identical to quantize_q8_1) =====================
    // 4 lanes per 32-value group, U groups per thread and pass (U = 4 on long rows: the pass is a latency chain of loads,
    // shuffles and a division, and every CTA repeats the whole row)
    const T *X = static_cast<const T *>(p.X);
    const int ngroups_all = p.K >> 5;
    // this CTA's share of the groups: [g_first, ngroups_row)
    const int g_share = (ngroups_all + (int)CL - 1) / (int)CL;
    const int g_first = (int)crank * g_share, ngroups_row = min(ngroups_all, g_first + g_share);
    const uint32_t act_q_s = smem_u32(act_q), act_s_s = smem_u32(act_s);
    constexpr int GPI = (NW * 32) / 4;
    auto quantize_pass = [&](auto u_tag, int base, bool first) {
        constexpr int U = decltype(u_tag)::value;
        const int t4 = lane & 3;
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int g = base + u * GPI + (tid >> 2);
            if (g < ngroups_row && !(p.dbg & 1)) {
                load_act8<T>(X, static_cast<const T *>(p.xin.x2), p.xin.ldx, 0, g * 32 + t4 * 8, p.x_vec != 0, v[u]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) v[u][i] = 0.0f;
            }
        }
        // the rest of the weight pipeline goes out BEHIND the first activation loads
        if (first) {
#pragma unroll
            for (int s = 0; s < S; s++)
                if (s >= early) issue_row(my_ring_s + s * stage_bytes);
            if (tr && tid == 0) tr[2] = clk64() + (long long)(v[0][0] == 123.456f);      // after the first X loads have landed
            if (CL > 1) cluster_wait();                           // every CTA of the cluster runs: its shared memory may be written
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (base + u * GPI >= ngroups_row) break;             // warp-uniform
            const int g = base + u * GPI + (tid >> 2);
            const Q8Group q = q8_1_quantize_group<false>(v[u]);
            int mine = __dp4a((int)q.q_lo, 0x01010101, __dp4a((int)q.q_hi, 0x01010101, 0));
            mine += __shfl_xor_sync(0xffffffffu, mine, 1);
            const int other = __shfl_xor_sync(0xffffffffu, mine, 2);
            if (g < ngroups_row) {
                const int it = g >> 2, j = g & 3;
                const int i = 2 * j + (t4 >> 1);
                const float d8 = __half2float(__float2half_rn(q.d));
                if (CL == 1) {
                    uint8_t *dst = reinterpret_cast<uint8_t *>(act_q + i * ipr + it) + 8 * (t4 & 1);
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(q.q_lo, q.q_hi);
                    if (t4 == 0) {
                        reinterpret_cast<float *>(act_s + it)[j] = d8;
                        reinterpret_cast<uint32_t *>(act_s + ipr + it)[j] = Ops::aux_word(d8, mine, other);
                    }
                } else {
                    const uint32_t qa = act_q_s + (uint32_t)(i * ipr + it) * 16u + 8u * (t4 & 1);
                    const uint32_t sa = act_s_s + (uint32_t)it * 16u + 4u * j;
                    const uint32_t bar_s = smem_u32(act_bar);
                    for (uint32_t r = 0; r < CL; r++) {           // the same bytes into every CTA of the cluster (its own included)
                        const uint32_t mb = mapa(bar_s, r);
                        st_async_v2(mapa(qa, r), q.q_lo, q.q_hi, mb);
                        if (t4 == 0) {
                            st_async_f32(mapa(sa, r), d8, mb);
                            st_async_f32(mapa(sa + (uint32_t)ipr * 16u, r), __uint_as_float(Ops::aux_word(d8, mine, other)), mb);
                        }
                    }
                }
            }
        }
    };
    uint32_t a[32];
    float4 d8 = make_float4(0.f, 0.f, 0.f, 0.f), ds = d8;
    if (p.image != nullptr) {
        // pre-quantized: no shared-memory staging, no CTA barrier -- 10 x 16 bytes per lane straight from L2
#pragma unroll
        for (int s = 0; s < S; s++)
            if (s >= early) issue_row(my_ring_s + s * stage_bytes);
        const uint4 *iq = reinterpret_cast<const uint4 *>(p.image);
        const float4 *is = reinterpret_cast<const float4 *>(p.image + p.off_scales);
        if (live) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint4 v = __ldg(iq + i * ipr + item);
                a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
            }
            d8 = __ldg(is + item);
            ds = __ldg(is + ipr + item);
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) a[i] = 0u;
        }
        if (tr && tid == 0) { tr[2] = clk64(); tr[3] = tr[2]; }
    } else {
    if (ngroups_row - g_first > 2 * GPI) {
        for (int base = g_first; base < ngroups_row || base == g_first; base += 4 * GPI) quantize_pass(std::integral_constant<int, 4>{}, base, base == g_first);
    } else {
        for (int base = g_first; base < ngroups_row || base == g_first; base += 2 * GPI) quantize_pass(std::integral_constant<int, 2>{}, base, base == g_first);
    }
    if (tr && tid == 0) tr[3] = clk64();
    if (CL > 1) {
        mbar_wait(act_bar, 0);     // every CTA's part has landed in this CTA's shared memory
    } else {
        __syncthreads();
    }

    // this lane's item of every row chunk it will see: activations -> registers
    if (live) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint4 v = act_q[i * ipr + item];
            a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
        }
        d8 = act_s[item];
        ds = act_s[ipr + item];
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++) a[i] = 0u;
    }
    }
    if (tr && tid == 0) tr[4] = clk64();

    // ===================== main loop: one step = chunk c of this warp's next row =====================
    // Rows go four at a time (one reduction per four).  With S == 4 the stage of a row is its position in the group: all
    // shared-memory offsets are constants.  Lanes without an item (partial chunks) read super-block 0 and their result is
    // dropped: no branch around the dot products.
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    const uint32_t hd0 = smem_u32(my_ring) + (live ? sbl : 0) * Ops::BS;  // the lane's super-block in stage 0
    const int qoff = 16 + 64 * h;
    // (opaque to the compiler: under the 80-register cap it would otherwise recompute these addresses from the thread
    // index for every row, ~14 instructions)
    uint32_t hd0_k = hd0, ring_k = my_ring_s;
    asm volatile("" : "+r"(hd0_k), "+r"(ring_k));
    uint32_t hdp = hd0_k;                                                // (S != 4: stage pointers advance and wrap)
    uint32_t dstp = ring_k;
    int stage = 0;
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j0 = 0; j0 < my_rows; j0 += 4) {
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            part[jj] = 0.0f;
            if (j0 + jj < my_rows) {                                   // warp-uniform
                const uint32_t hp = (S == 4) ? hd0_k + jj * stage_bytes : hdp;
                const uint32_t dp = (S == 4) ? ring_k + jj * stage_bytes : dstp;
                cp_async_wait<S - 1>();                                // this thread's pieces of the oldest row have landed
                __syncwarp();                                          // ... and so have the other lanes'
                float v;
                if constexpr (TYPE == T_Q4_K) v = Ops::item(hp, hp + qoff, a, d8, ds, h);
                else v = Ops::item(hp, a, d8, ds, h);
                part[jj] = (FULL || live) ? v : 0.0f;
                __syncwarp();                                          // every lane has read the stage
                issue_row(dp);                                         // refill it
                if (S != 4) {
                    hdp += stage_bytes;
                    dstp += stage_bytes;
                    if (++stage == S) { stage = 0; hdp = hd0_k; dstp = ring_k; }
                }
            }
        }
        if (tr && tid == 0 && j0 == 0) tr[5] = clk64();
        // four rows reduced together: xor 16 and xor 8 fold the four values onto the four lane octets, then 4, 2, 1.
        // Every row's sum is the same fixed tree whatever its position in the group (fp32 add is commutative).
        float x0 = (up16 ? part[2] : part[0]) + __shfl_xor_sync(0xffffffffu, up16 ? part[0] : part[2], 16);
        float x1 = (up16 ? part[3] : part[1]) + __shfl_xor_sync(0xffffffffu, up16 ? part[1] : part[3], 16);
        float v = (up8 ? x1 : x0) + __shfl_xor_sync(0xffffffffu, up8 ? x0 : x1, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        const int jr = j0 + (up16 ? 2 : 0) + (up8 ? 1 : 0);            // the row this lane octet holds
        if ((lane & 7) == 0 && jr < my_rows) {
            const int lr = rl + jr * RW;                               // row inside the CTA's range
            if (C == 1) ggb_yout_store<T>(p.yo, 0, r_begin + lr, v);
            else part_sm[lr * C + c] = v;
        }
    }
    cp_async_wait<0>();
    if (C > 1) {
        __syncthreads();
        for (int lr = tid; lr < nrows; lr += NW * 32) {
            float s = part_sm[lr * C];
            for (int cc = 1; cc < C; cc++) s += part_sm[lr * C + cc];
            ggb_yout_store<T>(p.yo, 0, r_begin + lr, s);
        }
    }
    if (tr && tid == 0) tr[6] = clk64();
}

template <int TYPE, typename T, int S, bool FULL, int NW>
int launch_cfg(RgParams &p, int grid, int smem, int cluster, cudaStream_t stream) {
    const DeviceInfo &di = ggb_device_info();
    static bool attr_set[64] = {};               // per instantiation and device; idempotent, so a race is harmless
    if (!attr_set[di.dev]) {
        GGB_CUDA_TRY(cudaFuncSetAttribute(mmvq_rg_kernel<TYPE, T, S, FULL, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        attr_set[di.dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[3];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cluster;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    // the two CTAs of a cluster on DIFFERENT SMs: with room for two CTAs per SM the default policy may put a pair on one SM,
    // which leaves another SM without a CTA of this launch (measured in the 70B block, gate|up -> down -> gate|up: the CTAs of
    // one launch ended up to 13 us apart; profiles/r02_block_trace.log)
    attr[2].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
    const int pol = (ggb_sb::g_rg_mode >> 24) & 3;
    attr[2].val.clusterSchedulingPolicyPreference = pol == 1 ? cudaClusterSchedulingPolicyDefault
                                                   : pol == 2 ? cudaClusterSchedulingPolicyLoadBalancing : cudaClusterSchedulingPolicySpread;
    cfg.attrs = attr;
    cfg.numAttrs = cluster > 1 ? 3 : 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmvq_rg_kernel<TYPE, T, S, FULL, NW>>(di.dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmvq_rg_kernel<TYPE, T, S, FULL, NW>, p));
    return GGB_OK;
}

template <int TYPE, typename T>
int launch(const void *W, const void *X, void *Y, int64_t K, int64_t N, const void *image, int probe_only, cudaStream_t stream) {
    using Ops = RgOps<TYPE>;
    const DeviceInfo &di = ggb_device_info();
    RgParams p;
    p.image = static_cast<const uint8_t *>(image);
    p.W = static_cast<const uint8_t *>(W);
    p.X = X;
    p.yo = ggb_make_yout(Y, N);
    p.xin = ggb_make_xin(K);
    p.N = N;
    p.K = (int)K;
    p.nsb = (int)(K / 256);
    p.ipr = 2 * p.nsb;
    p.row_bytes = p.nsb * Ops::BS;
    p.C = (p.nsb + 15) / 16;
    constexpr bool Q6 = TYPE == T_Q6_K;
    // warps: 12; seven chunks (K = 28672) take 14 (Q4_K) or 7 (Q6_K: its 3360-byte stages leave room for one row lane only)
    // (Q4_K with a pre-launched image also takes 7 warps: beside a 14-warp CTA and the 12-warp neighbour no register bank has room
    // for a CTA of the pre-launch, beside a 7-warp one it has -- see "Registers")
    constexpr bool Q5 = TYPE == T_Q5_K;
    constexpr bool Q3 = TYPE == T_Q3_K || TYPE == T_Q2_K;      // (the two 2-bit-plane types share the launch table)
    if (Q3 && p.C == 7) return 1;                       // (not instantiated for 7 chunks: the generic kernel keeps K = 28672)
    const int nw = (p.C == 7) ? ((Q6 || Q5 || image != nullptr) ? 7 : 14) : 12;
    const bool full = p.nsb % 16 == 0;
    if (p.C > 7 || nw % p.C != 0 || p.row_bytes % 16 != 0 || (p.C == 7 && !full)) return 1;
    // Q6_K, K = 28672 with the fused SwiGLU activation: 7 warps x 2 CTAs would spend 4.6 us on the 28672 exponentials (70B block
    // 91.3 us): that case takes its activations from the pre-launch above (ggb_mmvq_rg_wants_image / ggb_mmvq_rg_prequant)
    if (Q6 && p.C == 7 && p.xin.x2 != nullptr && image == nullptr) return 1;
    // balanced chunks when every chunk then starts and ends on a 16-byte boundary (Q6_K: multiples of 8 super-blocks), else
    // 16-super-block chunks and a shorter last one
    p.spc = (p.nsb + p.C - 1) / p.C;
    if ((p.spc * Ops::BS) % 16 != 0 || ((p.nsb - (p.C - 1) * p.spc) * Ops::BS) % 16 != 0) {
        p.spc = 16;
        if (((p.nsb - (p.C - 1) * 16) * Ops::BS) % 16 != 0) return 1;
    }
    p.RW = nw / p.C;
    p.wstep = (long long)p.RW * p.row_bytes;
    p.x_vec = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.xin.x2) & 15) == 0 &&
              (p.xin.ldx * (int64_t)sizeof(T)) % 16 == 0;
    p.pdl_prefetch = ggb_sb::g_pdl_prefetch;
    p.dbg = (ggb_sb::g_rg_mode >> 8) & 0xff;
    p.early = (ggb_sb::g_rg_mode >> 16) & 0xf;
    const int64_t max_ctas = (N + p.RW - 1) / p.RW;
    int grid = (int)(max_ctas < di.sm_count ? max_ctas : di.sm_count);
    // CTAs per cluster sharing the activation quantization (debug bits 20..23 of the mode: 1, 2 or 4; default below)
    int cluster = (ggb_sb::g_rg_mode >> 20) & 0xf;
    if (cluster != 1 && cluster != 2 && cluster != 4) cluster = K >= 8192 ? kDefaultCluster : 1;   // short rows: the pass is a latency chain, splitting it gains nothing
    if (image != nullptr || grid < 2 * cluster) cluster = 1;
    p.share = cluster > 1;
    // Placement.  Two CTAs fit on an SM (this launch and the next are meant to share it), and a plain launch onto idle SMs was
    // seen to put TWO of its CTAs on every second SM (per-CTA %smid stamps, profiles/r02_block_trace.log: 72 SMs with two CTAs
    // of a launch, 72 with none -- and the launch after it inherits the other slots, so the pattern persists).  A cluster launch
    // with the spread policy puts the CTAs of a pair on different SMs; it is used for placement even when nothing is shared.
    const int place = (ggb_sb::g_rg_mode >> 26) & 1 ? 1 : 2;               // debug bit 26: plain launches
    if (cluster == 1 && place == 2 && grid >= 4) cluster = 2;
    grid = (grid + cluster - 1) / cluster * cluster;       // (CTAs past the last row quantize their share and exit)
    if (grid > di.sm_count) grid -= cluster;
    p.part_rows = p.C > 1 ? (int)((N + grid - 1) / grid + 1) : 0;
    if ((int64_t)p.part_rows * p.C * 4 > 16384) return 1;
    p.off_scales = ((int)K + 15) & ~15;                     // (also the scale table's offset inside a pre-launched image)
    // a pre-launched image is read straight into registers: no activation area in shared memory, the ring gets the room
    p.off_part = image != nullptr ? 0 : p.off_scales + 2 * p.ipr * 16;
    p.off_bar = (p.off_part + p.part_rows * p.C * 4 + 7) & ~7;                // (only a cluster launch uses it)
    p.off_slots = (p.off_bar + (p.share ? 8 : 0) + 127) & ~127;
    const int half_sm = (di.max_smem_optin + 1024) / 2 - 1024;               // two CTAs per SM
    const int force_s = (ggb_sb::g_rg_mode >> 12) & 0xf;                     // debug: 3 or 4 stages
    constexpr int stage_bytes = 16 * Ops::BS;
    int S = 4;
    while (S > 2 && p.off_slots + nw * S * stage_bytes > half_sm) S--;
    if (force_s == 3 && S > 3) S = 3;
    // instantiated -- Q4_K: 12 warps x {3, 4} stages, 14 warps x {2, 3}, 7 warps x {3, 4}; Q5_K: 12 warps x {2, 3}, 7 warps x {3, 4};
    // Q6_K: 12 warps x 2, 7 warps x {2, 3}
    if ((nw == 14 || (nw == 7 && Q6) || (nw == 12 && Q5)) && S > 3) S = 3;
    if (nw == 7 && !Q6 && S < 3) return 1;
    if (Q6 && nw == 12) S = 2;
    if (!Q6 && !Q5 && nw == 12 && S < 3) return 1;
    const int smem = p.off_slots + nw * S * stage_bytes;
    if (smem > half_sm) return 1;
    if (probe_only) return 2;
    if (p.early == 0 || p.early > S) p.early = 1;
    p.trace = ggb_mmvq_trace_slot();
    if constexpr (Q3) {
        if (S >= 4) return full ? launch_cfg<TYPE, T, 4, true, 12>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 4, false, 12>(p, grid, smem, cluster, stream);
        return full ? launch_cfg<TYPE, T, 3, true, 12>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 3, false, 12>(p, grid, smem, cluster, stream);
    } else if constexpr (Q5) {
        if (nw == 7)
            return S == 4 ? launch_cfg<TYPE, T, 4, true, 7>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 3, true, 7>(p, grid, smem, cluster, stream);
        if (S == 3) return full ? launch_cfg<TYPE, T, 3, true, 12>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 3, false, 12>(p, grid, smem, cluster, stream);
        return full ? launch_cfg<TYPE, T, 2, true, 12>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 2, false, 12>(p, grid, smem, cluster, stream);
    } else if constexpr (Q6) {
        if (nw == 7)
            return S == 3 ? launch_cfg<TYPE, T, 3, true, 7>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 2, true, 7>(p, grid, smem, cluster, stream);
        return full ? launch_cfg<TYPE, T, 2, true, 12>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 2, false, 12>(p, grid, smem, cluster, stream);
    } else {
        if (nw == 14)
            return S == 3 ? launch_cfg<TYPE, T, 3, true, 14>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 2, true, 14>(p, grid, smem, cluster, stream);
        if (nw == 7)
            return S == 4 ? launch_cfg<TYPE, T, 4, true, 7>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 3, true, 7>(p, grid, smem, cluster, stream);
        if (S == 4) return full ? launch_cfg<TYPE, T, 4, true, 12>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 4, false, 12>(p, grid, smem, cluster, stream);
        return full ? launch_cfg<TYPE, T, 3, true, 12>(p, grid, smem, cluster, stream) : launch_cfg<TYPE, T, 3, false, 12>(p, grid, smem, cluster, stream);
    }
}

template <int TYPE>
int dispatch_dtype(const void *W, const void *X, void *Y, int dt, int64_t K, int64_t N, const void *image, int probe_only,
                   cudaStream_t stream) {
    switch (dt) {
        case GGB_F32:  return launch<TYPE, float>(W, X, Y, K, N, image, probe_only, stream);
        case GGB_F16:  return launch<TYPE, __half>(W, X, Y, K, N, image, probe_only, stream);
        case GGB_BF16: return launch<TYPE, __nv_bfloat16>(W, X, Y, K, N, image, probe_only, stream);
        default: return 1;
    }
}

}  // namespace

void ggb_mmvq_rg_set_mode(int mode) { ggb_sb::g_rg_mode = mode; }

// returns GGB_OK / a negative status, or 1 when this kernel does not apply (the caller goes on to mmvq_sb / mmvq);
// probe_only: 2 = it applies (nothing launched)
int ggb_mmvq_rg_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                     const void *image, int probe_only, cudaStream_t stream) {
    if (!(ggb_sb::g_rg_mode & 1) || B != 1 || K % 256 != 0 || K > 28672 || (reinterpret_cast<uintptr_t>(W) & 15)) return 1;
    switch (type) {
        case T_Q4_K: return dispatch_dtype<T_Q4_K>(W, X, Y, xy_dtype, K, N, image, probe_only, stream);
        case T_Q2_K: return (K % 1024 == 0 && (ggb_sb::g_rg_mode & (1u << 31)) == 0)       // (84-byte super-blocks: rows and chunks 16-byte aligned every 4; debug bit 31: off)
                                ? dispatch_dtype<T_Q2_K>(W, X, Y, xy_dtype, K, N, image, probe_only, stream) : 1;
        case T_Q3_K: return (K % 2048 == 0 && (ggb_sb::g_rg_mode & (1 << 30)) == 0)        // (debug bit 30: generic kernel)
                                ? dispatch_dtype<T_Q3_K>(W, X, Y, xy_dtype, K, N, image, probe_only, stream) : 1;
        case T_Q5_K: return (ggb_sb::g_rg_mode & (1 << 29)) == 0 ? dispatch_dtype<T_Q5_K>(W, X, Y, xy_dtype, K, N, image, probe_only, stream) : 1;   // (debug bit 29: ring kernel)
        case T_Q6_K: return (K % 2048 == 0 && (ggb_sb::g_rg_mode & (1 << 27)) == 0)        // (debug bit 27: Q6_K stays on the ring kernel)
                                ? dispatch_dtype<T_Q6_K>(W, X, Y, xy_dtype, K, N, image, probe_only, stream) : 1;
        default: return 1;
    }
}

// The one case that wants its activation image pre-launched in this kernel's layout (see mmvq_rg_prequant_kernel)
int ggb_mmvq_rg_wants_image(int type, int64_t B, int64_t K, const void *x2) {
    // (plain rows too: measured on the 70B down projection, 7 warps behind the pre-launch vs the in-kernel pair quantization:
    //  Q4_K 24.0 vs 26.0 us (14 warps), Q6_K 34.5 vs 37.6 us (7 warps))
    (void)x2;
    // rows of K >= 12288 (measured with / without: K = 14336 Q4_K 9.03 / 9.72 us, Q6_K 13.7 / 14.1; K = 28672 see above); debug bit 9:
    // from K = 8192
    const int64_t min_k = ((ggb_sb::g_rg_mode >> 8) & 2) ? 8192 : 12288;
    if (!(ggb_sb::g_rg_mode & 1) || B != 1 || K < min_k || K > 28672) return 0;
    if (type == T_Q5_K) return (ggb_sb::g_rg_mode & (1 << 29)) == 0;
    return type == T_Q4_K ? ((ggb_sb::g_rg_mode >> 28) & 1) == 0 : (type == T_Q6_K && !(ggb_sb::g_rg_mode & (1 << 27)));   // (debug bits 27 / 28 / 29: off)
}
int64_t ggb_mmvq_rg_image_bytes(int64_t K) { return ((K + 15) & ~(int64_t)15) + 2 * (K / 128) * 16; }

namespace {
template <int TYPE, typename T>
int launch_prequant(const void *X, int64_t K, void *image, cudaStream_t stream) {
    const GgbXIn xin = ggb_make_xin(K);
    const int x_vec = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(xin.x2) & 15) == 0 &&
                      (xin.ldx * (int64_t)sizeof(T)) % 16 == 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((K / 32 + 31) / 32));
    cfg.blockDim = dim3(128);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmvq_rg_prequant_kernel<TYPE, T>>(ggb_device_info().dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmvq_rg_prequant_kernel<TYPE, T>, static_cast<const T *>(X), xin, (int)K, x_vec,
                                    static_cast<uint8_t *>(image), (int)((K + 15) & ~(int64_t)15)));
    return GGB_OK;
}
}  // namespace

int ggb_mmvq_rg_prequant(const void *X, int type, int xy_dtype, int64_t K, void *image, cudaStream_t stream) {
    if (type != T_Q6_K && type != T_Q4_K && type != T_Q5_K) return 1;
    const bool q6 = type == T_Q6_K;              // (Q4_K and Q5_K share the image: the second word is d8 * sum(q8))
    switch (xy_dtype) {
        case GGB_F32:  return q6 ? launch_prequant<T_Q6_K, float>(X, K, image, stream) : launch_prequant<T_Q4_K, float>(X, K, image, stream);
        case GGB_F16:  return q6 ? launch_prequant<T_Q6_K, __half>(X, K, image, stream) : launch_prequant<T_Q4_K, __half>(X, K, image, stream);
        case GGB_BF16: return q6 ? launch_prequant<T_Q6_K, __nv_bfloat16>(X, K, image, stream) : launch_prequant<T_Q4_K, __nv_bfloat16>(X, K, image, stream);
        default: return 1;
    }
}
