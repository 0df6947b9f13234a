// mmvq_sb_body.cuh -- the kernel and host launcher of the fast MMVQ path, included by one translation unit per weight
// type (mmvq_sb_q4k.cu, ...: the 200 kernel instantiations compile in parallel) -- see the description below.
#pragma once
// The fast MMVQ path: one lane owns one 128-value half super-block of two weight rows.
//
// Same contract as mmvq.cu (ggml_mul_mat_vec_a8 = quantize_row_q8_1_cuda + mul_mat_vec_q + vec_dot_*_q8_1,
// hf-kernels/ggml-kernels/ggml/ggml_kernel.cu:13-66,80-193; ggml/mmvq.cuh:2-209; ggml/vecdotq.cuh:45-605),
// taken when K % 256 == 0 and rows are 16-byte aligned (every Llama shape); mmvq.cu stays as the path for
// the remaining shapes and types.  The kernel is HBM-bound by construction: the instruction budget per
// super-block is several times below the issue rate needed to keep up with 6.5 TB/s, so what matters is
// keeping ~50-100 KB per SM in flight at all times and hiding the serial start of every launch.
//
//   * One CTA per SM per launch, but a HALF-SM CTA (7 consumer warps + 1 producer warp, <= 113 KB of
//     shared memory, <= 128 registers): launches are programmatic dependent launches, so the CTAs of the
//     NEXT launch become resident next to the current ones.  With GGB_OPT_STATIC_WEIGHTS their producer
//     starts streaming weights immediately, while their consumers wait for the dependency
//     (griddepcontrol.wait) before touching X or Y: HBM keeps streaming through the serial part of
//     every launch (dependency resolution, activation quantization, tail).
//   * The CTA owns a contiguous range of weight rows.  The producer thread streams it through a ring of
//     small slots with cp.async.bulk (TMA engine), one mbarrier pair per slot, L2 evict-first.  A slot is
//     exactly ONE consumer warp's next piece of work (a row pair x <= 32 items of 128 values), so warps
//     never wait for each other.
//   * Prologue: the consumers quantize the B x K activations to q8_1 in shared memory, bit-identical to
//     quantize_q8_1 (q8_1.cuh), in an item-transposed layout [16-byte chunk][item] so that lanes owning
//     consecutive items read consecutive 16-byte words (conflict-free LDS.128), plus per-group fp32 tables
//     {d8, d8 * sum(q8)}.
//   * Per slot a lane unpacks ITS half super-block of rows 2p and 2p+1 (128-bit shared loads, scales
//     decoded in packed bytes), takes the integer dots with dp4a against activation words that both rows
//     reuse, applies the reference's scale/min algebra in fp32; the lanes of a row pair reduce with
//     xor-shuffles.  Int -> float conversions use the exact magic-number form (IMAD/PRMT + FADD) instead
//     of the quarter-rate I2F pipe.
#include "q8_1.cuh"

namespace {

constexpr int kMaxSlots = 64;
constexpr int kBatchBlock = 4;          // batch rows processed per pass over an item (register budget)

struct SbParams {
    const uint8_t *W;
    const void *X;
    GgbXIn xin;
    GgbYOut yo;
    int64_t N;
    int K;
    int nsb;             // super-blocks per row
    int ipr;             // items (128 values) per row = 2 * nsb
    int row_bytes;
    int lpp;             // lanes per row pair (power of two <= 32)
    int rows_per_slot;   // 2 * (32 / lpp)
    int nchunks;         // ceil(ipr / 32) when lpp == 32, else 1
    int slot_stride;     // shared-memory bytes per ring slot (multiple of 128)
    int chunk_stride;    // bytes between the two rows of a slot when nchunks > 1 (multiple of 16)
    int warps_used;      // consumer warps that take slots (min(NW, nslots))
    int nslots;          // ring depth (>= warps_used)
    int off_scales, off_slots, off_bars;
    int x_vec;
    int pdl_prefetch;    // 1: weights are static, the producer does not wait for the previous kernel
    int l2_prefetch;     // bytes of this CTA's weight range to pull into L2 while waiting for the previous kernel
    int staged;          // long rows, B <= 2: quantize the activations K-chunk by K-chunk around the first row group's slots
    const uint8_t *image; // != nullptr: the q8_1 activation image (this kernel's shared-memory bytes [0, image_bytes), written by
    int image_bytes;      //             mmvq_sb_prequant_kernel, the previous launch of the stream) is copied in instead of quantizing
    unsigned long long *trace;   // debug: per-CTA globaltimer stamps {start, after dependency wait, first slot, end}
};

__device__ __forceinline__ uint4 lds128(const void *p) { return *reinterpret_cast<const uint4 *>(p); }

__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c) { return __dp4a((int)a, (int)b, c); }
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {      // unsigned bytes x signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// exact (float)(a * b) for |a * b| < 2^22: the product lands in the mantissa of 1.5 * 2^23
__device__ __forceinline__ float imul_to_f32(int a, int b) {
    return __int_as_float(a * b + 0x4B400000) - 12582912.0f;
}
__device__ __forceinline__ float int_to_f32(int a) {          // |a| < 2^22
    return __int_as_float(a + 0x4B400000) - 12582912.0f;
}
// exact (float)(byte j of w)
template <int J>
__device__ __forceinline__ float byte_to_f32(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540 + J)) - 8388608.0f;
}
template <int J>
__device__ __forceinline__ int byte_of(uint32_t w) { return (int)__byte_perm(w, 0, 0x4440 + J); }

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Per-type item operations.  An item is 128 consecutive values of a row = half `h` of super-block `sb`.
//   BS        bytes per 256 values
//   SCF4      float4 entries per (batch row, item) in the activation scale table
//   NEED_S8   the type uses block_q8_1.s (fp16-rounded float sum of the inputs) instead of integer sums
//   store_scales(tab, stride, j, d8, sq0, sq1, s8)   called by one lane per 32-value group j (0..3) of an
//             item: tab = &table[b][0][item] (float4 units, entry e at tab + e * stride)
//   run<BB>(w0, w1, h, aq, as, cs, acc)  w0/w1 = the super-block of the two rows in shared memory;
//             aq = &act_q[b0][0][item] (uint4 units: 16-value chunk i of batch b at aq[(b * 8 + i) * cs]),
//             as = &table[b0][0][item] (batch b, entry e at as[(b * SCF4 + e) * cs])
// ------------------------------------------------------------------------------------------------
template <int TYPE> struct SbOps;

// ---- Q4_K: d, dmin, 12 bytes of 6-bit (scale, min), 128 bytes of nibbles (vecdotq.cuh:272-327 algebra)
template <> struct SbOps<T_Q4_K> {
    static constexpr int BS = 144, SCF4 = 2;
    static constexpr bool NEED_S8 = false;

    __device__ __forceinline__ static void store_scales(float4 *tab, int stride, int j, float d8, int sq0, int sq1, float) {
        reinterpret_cast<float *>(tab)[j] = d8;
        reinterpret_cast<float *>(tab + stride)[j] = d8 * (float)(sq0 + sq1);
    }

    template <int BB>
    __device__ __forceinline__ static void run(const uint8_t *w0, const uint8_t *w1, int h, const uint4 *aq,
                                               const float4 *as, int cs, float (&acc)[2][BB]) {
        const uint8_t *wr[2] = {w0, w1};
        uint32_t sc[2], mn[2];
        float d[2], dmin[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint4 hd = lds128(wr[r]);
            const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&hd.x));
            d[r] = dm.x;
            dmin[r] = dm.y;
            // (scale, min) of the item's four 32-value groups, one per byte (dequantize.cuh:154-161)
            const uint32_t scA = hd.y & 0x3f3f3f3fu, mA = hd.z & 0x3f3f3f3fu;
            const uint32_t scB = (hd.w & 0x0f0f0f0fu) | ((hd.y >> 2) & 0x30303030u);
            const uint32_t mB = ((hd.w >> 4) & 0x0f0f0f0fu) | ((hd.z >> 2) & 0x30303030u);
            sc[r] = h ? scB : scA;
            mn[r] = h ? mB : mA;
        }
        float4 d8[BB];
#pragma unroll
        for (int b = 0; b < BB; b++) d8[b] = as[(b * SCF4) * cs];
        float fd[2][BB];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < BB; b++) fd[r][b] = 0.0f;

#pragma unroll
        for (int il = 0; il < 2; il++) {
            uint32_t lo[2][8], hi[2][8];
            int sc0[2], sc1[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint8_t *qs = wr[r] + 16 + 64 * h + 32 * il;
                const uint4 qa = lds128(qs), qb = lds128(qs + 16);
                const uint32_t q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    lo[r][i] = q[i] & 0x0f0f0f0fu;
                    hi[r][i] = q[i] & 0xf0f0f0f0u;                              // 16 x the high nibbles
                }
                if (il) { sc0[r] = byte_of<2>(sc[r]); sc1[r] = byte_of<3>(sc[r]); }
                else    { sc0[r] = byte_of<0>(sc[r]); sc1[r] = byte_of<1>(sc[r]); }
            }
#pragma unroll
            for (int b = 0; b < BB; b++) {
                const uint4 a0 = aq[(b * 8 + 4 * il + 0) * cs], a1 = aq[(b * 8 + 4 * il + 1) * cs];
                const uint4 a2 = aq[(b * 8 + 4 * il + 2) * cs], a3 = aq[(b * 8 + 4 * il + 3) * cs];
                const float d8lo = il ? d8[b].z : d8[b].x, d8hi = il ? d8[b].w : d8[b].y;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    int dl = dp4a_ss(lo[r][0], a0.x, 0), dl2 = dp4a_ss(lo[r][4], a1.x, 0);
                    int dh = dp4a_us(hi[r][0], a2.x, 0), dh2 = dp4a_us(hi[r][4], a3.x, 0);
                    dl = dp4a_ss(lo[r][1], a0.y, dl);  dl2 = dp4a_ss(lo[r][5], a1.y, dl2);
                    dh = dp4a_us(hi[r][1], a2.y, dh);  dh2 = dp4a_us(hi[r][5], a3.y, dh2);
                    dl = dp4a_ss(lo[r][2], a0.z, dl);  dl2 = dp4a_ss(lo[r][6], a1.z, dl2);
                    dh = dp4a_us(hi[r][2], a2.z, dh);  dh2 = dp4a_us(hi[r][6], a3.z, dh2);
                    dl = dp4a_ss(lo[r][3], a0.w, dl);  dl2 = dp4a_ss(lo[r][7], a1.w, dl2);
                    dh = dp4a_us(hi[r][3], a2.w, dh);  dh2 = dp4a_us(hi[r][7], a3.w, dh2);
                    fd[r][b] = fmaf(d8lo, imul_to_f32(dl + dl2, sc0[r]), fd[r][b]);
                    fd[r][b] = fmaf(d8hi, imul_to_f32((dh + dh2) >> 4, sc1[r]), fd[r][b]);
                }
            }
        }
        // minima: dmin * sum_j m_j * (d8_j * sum q8_j)
        float fm[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            fm[r][0] = byte_to_f32<0>(mn[r]); fm[r][1] = byte_to_f32<1>(mn[r]);
            fm[r][2] = byte_to_f32<2>(mn[r]); fm[r][3] = byte_to_f32<3>(mn[r]);
        }
#pragma unroll
        for (int b = 0; b < BB; b++) {
            const float4 s = as[(b * SCF4 + 1) * cs];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                float t = fm[r][0] * s.x;
                t = fmaf(fm[r][1], s.y, t); t = fmaf(fm[r][2], s.z, t); t = fmaf(fm[r][3], s.w, t);
                acc[r][b] += d[r] * fd[r][b] - dmin[r] * t;
            }
        }
    }
};

// ---- Q5_K: Q4_K plus a fifth bit per value in qh[32] (vecdotq.cuh:272-327 algebra)
template <> struct SbOps<T_Q5_K> {
    static constexpr int BS = 176, SCF4 = 2;
    static constexpr bool NEED_S8 = false;

    __device__ __forceinline__ static void store_scales(float4 *tab, int stride, int j, float d8, int sq0, int sq1, float) {
        reinterpret_cast<float *>(tab)[j] = d8;
        reinterpret_cast<float *>(tab + stride)[j] = d8 * (float)(sq0 + sq1);
    }

    template <int BB>
    __device__ __forceinline__ static void run(const uint8_t *w0, const uint8_t *w1, int h, const uint4 *aq,
                                               const float4 *as, int cs, float (&acc)[2][BB]) {
        const uint8_t *wr[2] = {w0, w1};
        uint32_t sc[2], mn[2];
        float d[2], dmin[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint4 hd = lds128(wr[r]);
            const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&hd.x));
            d[r] = dm.x;
            dmin[r] = dm.y;
            const uint32_t scA = hd.y & 0x3f3f3f3fu, mA = hd.z & 0x3f3f3f3fu;
            const uint32_t scB = (hd.w & 0x0f0f0f0fu) | ((hd.y >> 2) & 0x30303030u);
            const uint32_t mB = ((hd.w >> 4) & 0x0f0f0f0fu) | ((hd.z >> 2) & 0x30303030u);
            sc[r] = h ? scB : scA;
            mn[r] = h ? mB : mA;
        }
        float4 d8[BB];
#pragma unroll
        for (int b = 0; b < BB; b++) d8[b] = as[(b * SCF4) * cs];
        float fd[2][BB];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < BB; b++) fd[r][b] = 0.0f;

#pragma unroll
        for (int il = 0; il < 2; il++) {
            uint32_t lo[2][8], hi[2][8];
            float sc0[2], sc1[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint8_t *qs = wr[r] + 48 + 64 * h + 32 * il;
                const uint4 qa = lds128(qs), qb = lds128(qs + 16);
                const uint4 ha = lds128(wr[r] + 16), hb = lds128(wr[r] + 32);
                const uint32_t q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                const uint32_t qh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                const int s_lo = 4 * h + 2 * il;                                // fifth bit of group 2*il / 2*il+1
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t t = qh[i] >> s_lo;
                    lo[r][i] = (q[i] & 0x0f0f0f0fu) | ((t << 4) & 0x10101010u);
                    hi[r][i] = ((q[i] >> 4) & 0x0f0f0f0fu) | ((t << 3) & 0x10101010u);
                }
                if (il) { sc0[r] = byte_to_f32<2>(sc[r]); sc1[r] = byte_to_f32<3>(sc[r]); }
                else    { sc0[r] = byte_to_f32<0>(sc[r]); sc1[r] = byte_to_f32<1>(sc[r]); }
            }
#pragma unroll
            for (int b = 0; b < BB; b++) {
                const uint4 a0 = aq[(b * 8 + 4 * il + 0) * cs], a1 = aq[(b * 8 + 4 * il + 1) * cs];
                const uint4 a2 = aq[(b * 8 + 4 * il + 2) * cs], a3 = aq[(b * 8 + 4 * il + 3) * cs];
                const float d8lo = il ? d8[b].z : d8[b].x, d8hi = il ? d8[b].w : d8[b].y;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    int dl = dp4a_ss(lo[r][0], a0.x, 0), dl2 = dp4a_ss(lo[r][4], a1.x, 0);
                    int dh = dp4a_ss(hi[r][0], a2.x, 0), dh2 = dp4a_ss(hi[r][4], a3.x, 0);
                    dl = dp4a_ss(lo[r][1], a0.y, dl);  dl2 = dp4a_ss(lo[r][5], a1.y, dl2);
                    dh = dp4a_ss(hi[r][1], a2.y, dh);  dh2 = dp4a_ss(hi[r][5], a3.y, dh2);
                    dl = dp4a_ss(lo[r][2], a0.z, dl);  dl2 = dp4a_ss(lo[r][6], a1.z, dl2);
                    dh = dp4a_ss(hi[r][2], a2.z, dh);  dh2 = dp4a_ss(hi[r][6], a3.z, dh2);
                    dl = dp4a_ss(lo[r][3], a0.w, dl);  dl2 = dp4a_ss(lo[r][7], a1.w, dl2);
                    dh = dp4a_ss(hi[r][3], a2.w, dh);  dh2 = dp4a_ss(hi[r][7], a3.w, dh2);
                    fd[r][b] = fmaf(d8lo, int_to_f32(dl + dl2) * sc0[r], fd[r][b]);      // products < 2^24: exact
                    fd[r][b] = fmaf(d8hi, int_to_f32(dh + dh2) * sc1[r], fd[r][b]);
                }
            }
        }
        float fm[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            fm[r][0] = byte_to_f32<0>(mn[r]); fm[r][1] = byte_to_f32<1>(mn[r]);
            fm[r][2] = byte_to_f32<2>(mn[r]); fm[r][3] = byte_to_f32<3>(mn[r]);
        }
#pragma unroll
        for (int b = 0; b < BB; b++) {
            const float4 s = as[(b * SCF4 + 1) * cs];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                float t = fm[r][0] * s.x;
                t = fmaf(fm[r][1], s.y, t); t = fmaf(fm[r][2], s.z, t); t = fmaf(fm[r][3], s.w, t);
                acc[r][b] += d[r] * fd[r][b] - dmin[r] * t;
            }
        }
    }
};

// NW consecutive 32-bit words starting at a 2-byte-aligned shared address: aligned loads + funnel shifts
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

// ---- Q6_K: ql[128], qh[64], 16 int8 scales (one per 16 values), d; 210-byte blocks are only 2-byte aligned
// (vecdotq.cuh:331-349 algebra; the -32 offset goes through per-16 sums of the activation quants)
template <> struct SbOps<T_Q6_K> {
    static constexpr int BS = 210, SCF4 = 3;
    static constexpr bool NEED_S8 = false;

    // entry 0: d8 of the item's four groups; entries 1-2: 32 * d8 * sum(q8) of its eight 16-value chunks
    __device__ __forceinline__ static void store_scales(float4 *tab, int stride, int j, float d8, int sq0, int sq1, float) {
        reinterpret_cast<float *>(tab)[j] = d8;
        float *f = reinterpret_cast<float *>(tab + (1 + (j >> 1)) * stride) + 2 * (j & 1);
        f[0] = d8 * (float)(32 * sq0);
        f[1] = d8 * (float)(32 * sq1);
    }

    template <int BB>
    __device__ __forceinline__ static void run(const uint8_t *w0, const uint8_t *w1, int h, const uint4 *aq,
                                               const float4 *as, int cs, float (&acc)[2][BB]) {
        const uint8_t *wr[2] = {w0, w1};
        uint32_t ql[2][16], qh[2][8];
        float fsc[2][8], d[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            lds_words_a2<16>(wr[r] + 64 * h, ql[r]);
            lds_words_a2<8>(wr[r] + 128 + 32 * h, qh[r]);
            uint32_t sc[2];
            lds_words_a2<2>(wr[r] + 192 + 8 * h, sc);
            d[r] = ggb_h2f(ld_u16(wr[r] + 208));
            const uint32_t s0 = sc[0] ^ 0x80808080u, s1 = sc[1] ^ 0x80808080u;   // int8 -> biased byte
            fsc[r][0] = byte_to_f32<0>(s0) - 128.0f; fsc[r][1] = byte_to_f32<1>(s0) - 128.0f;
            fsc[r][2] = byte_to_f32<2>(s0) - 128.0f; fsc[r][3] = byte_to_f32<3>(s0) - 128.0f;
            fsc[r][4] = byte_to_f32<0>(s1) - 128.0f; fsc[r][5] = byte_to_f32<1>(s1) - 128.0f;
            fsc[r][6] = byte_to_f32<2>(s1) - 128.0f; fsc[r][7] = byte_to_f32<3>(s1) - 128.0f;
        }
        float fd[2][BB];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < BB; b++) fd[r][b] = 0.0f;

        // group g (32 values) of the item: g = 0,1 low nibbles of ql[0..31] / ql[32..63], g = 2,3 high nibbles;
        // its two 16-value chunks 2g, 2g+1 carry scales 2g, 2g+1 (dequantize.cuh:237-253)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t v[2][8];
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t l = ql[r][8 * (g & 1) + i];
                    const uint32_t lo = (g < 2) ? (l & 0x0f0f0f0fu) : ((l >> 4) & 0x0f0f0f0fu);
                    const uint32_t hb = (g == 0) ? (qh[r][i] << 4) : (g == 1) ? (qh[r][i] << 2)
                                      : (g == 2) ? qh[r][i] : (qh[r][i] >> 2);
                    v[r][i] = lo | (hb & 0x30303030u);                           // raw = value + 32
                }
#pragma unroll
            for (int b = 0; b < BB; b++) {
                const uint4 a0 = aq[(b * 8 + 2 * g) * cs], a1 = aq[(b * 8 + 2 * g + 1) * cs];
                const float4 d8 = as[(b * SCF4) * cs];
                const float4 s16 = as[(b * SCF4 + 1 + (g >> 1)) * cs];
                const float d8g = (g == 0) ? d8.x : (g == 1) ? d8.y : (g == 2) ? d8.z : d8.w;
                const float s0 = (g & 1) ? s16.z : s16.x, s1 = (g & 1) ? s16.w : s16.y;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    int d0 = dp4a_ss(v[r][0], a0.x, 0), d1 = dp4a_ss(v[r][4], a1.x, 0);
                    d0 = dp4a_ss(v[r][1], a0.y, d0);  d1 = dp4a_ss(v[r][5], a1.y, d1);
                    d0 = dp4a_ss(v[r][2], a0.z, d0);  d1 = dp4a_ss(v[r][6], a1.z, d1);
                    d0 = dp4a_ss(v[r][3], a0.w, d0);  d1 = dp4a_ss(v[r][7], a1.w, d1);
                    fd[r][b] = fmaf(fsc[r][2 * g], fmaf(d8g, int_to_f32(d0), -s0), fd[r][b]);
                    fd[r][b] = fmaf(fsc[r][2 * g + 1], fmaf(d8g, int_to_f32(d1), -s1), fd[r][b]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < BB; b++) acc[r][b] += d[r] * fd[r][b];
    }
};

// ---- IQ4_XS: d, scales_h (2 bits x 8), scales_l (4 bits x 8), 128 bytes of 4-bit codes into the non-linear codebook
// (vecdotq.cuh:867-888 algebra: sum_ib d * (ls_ib - 32) * d8_ib * dot_ib).  136-byte super-blocks: 8-byte aligned loads.
template <> struct SbOps<T_IQ4_XS> {
    static constexpr int BS = 136, SCF4 = 1;
    static constexpr bool NEED_S8 = false;

    __device__ __forceinline__ static void store_scales(float4 *tab, int, int j, float d8, int, int, float) {
        reinterpret_cast<float *>(tab)[j] = d8;
    }

    template <int BB>
    __device__ __forceinline__ static void run(const uint8_t *w0, const uint8_t *w1, int h, const uint4 *aq,
                                               const float4 *as, int cs, float (&acc)[2][BB]) {
        const uint8_t *wr[2] = {w0, w1};
        float fsc[2][4], d[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint2 hd = *reinterpret_cast<const uint2 *>(wr[r]);
            d[r] = ggb_h2f(hd.x & 0xffffu);
            const uint32_t sl = hd.y >> (16 * h), sh = hd.x >> (16 + 8 * h);      // the item's four sub-blocks
#pragma unroll
            for (int g = 0; g < 4; g++)
                fsc[r][g] = int_to_f32((int)(((sl >> (4 * g)) & 0xfu) | (((sh >> (2 * g)) & 3u) << 4)) - 32);
        }
        float fd[2][BB];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < BB; b++) fd[r][b] = 0.0f;

        // sub-block g (32 values): 16 bytes of codes, low nibbles = values 0..15 (chunk 2g), high = 16..31 (chunk 2g+1)
#pragma unroll
        for (int g = 0; g < 4; g++) {
            uint32_t vlo[2][4], vhi[2][4];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint8_t *qs = wr[r] + 8 + 64 * h + 16 * g;
                const uint2 qa = *reinterpret_cast<const uint2 *>(qs), qb = *reinterpret_cast<const uint2 *>(qs + 8);
                iq4nl_lut8(qa.x, vlo[r][0], vhi[r][0]);
                iq4nl_lut8(qa.y, vlo[r][1], vhi[r][1]);
                iq4nl_lut8(qb.x, vlo[r][2], vhi[r][2]);
                iq4nl_lut8(qb.y, vlo[r][3], vhi[r][3]);
            }
#pragma unroll
            for (int b = 0; b < BB; b++) {
                const uint4 a0 = aq[(b * 8 + 2 * g) * cs], a1 = aq[(b * 8 + 2 * g + 1) * cs];
                const float4 d8 = as[(b * SCF4) * cs];
                const float d8g = (g == 0) ? d8.x : (g == 1) ? d8.y : (g == 2) ? d8.z : d8.w;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    int d0 = dp4a_ss(vlo[r][0], a0.x, 0), d1 = dp4a_ss(vhi[r][0], a1.x, 0);
                    d0 = dp4a_ss(vlo[r][1], a0.y, d0);  d1 = dp4a_ss(vhi[r][1], a1.y, d1);
                    d0 = dp4a_ss(vlo[r][2], a0.z, d0);  d1 = dp4a_ss(vhi[r][2], a1.z, d1);
                    d0 = dp4a_ss(vlo[r][3], a0.w, d0);  d1 = dp4a_ss(vhi[r][3], a1.w, d1);
                    fd[r][b] = fmaf(fsc[r][g] * d8g, int_to_f32(d0 + d1), fd[r][b]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < BB; b++) acc[r][b] += d[r] * fd[r][b];
    }
};

// all B batch rows of one item pair, kBatchBlock at a time (the weights are re-read from shared memory for
// every block of batch rows: cheap next to the B-proportional dp4a work)
template <int TYPE, int B, int B0>
__device__ __forceinline__ void run_batches(const uint8_t *w0, const uint8_t *w1, int h, const uint4 *aq,
                                            const float4 *as, int cs, float (&acc)[2][B]) {
    if constexpr (B0 < B) {
        constexpr int BB = (B - B0 < kBatchBlock) ? (B - B0) : kBatchBlock;
        float part[2][BB];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < BB; b++) part[r][b] = 0.0f;
        SbOps<TYPE>::template run<BB>(w0, w1, h, aq + (B0 * 8) * cs, as + (B0 * SbOps<TYPE>::SCF4) * cs, cs, part);
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < BB; b++) acc[r][B0 + b] += part[r][b];
        run_batches<TYPE, B, B0 + BB>(w0, w1, h, aq, as, cs, acc);
    }
}

// quantize + store 32-value group `gi` of the range [g_lo, g_lo + span) x B: x = this lane's 8 values (lane t4 of the quad)
template <int TYPE, int B>
__device__ __forceinline__ void sb_q_store1(uint4 *act_q, float4 *act_s, int ipr, float (&x)[8], int gi, int g_lo, int span, int t4) {
    using Ops = SbOps<TYPE>;
    const bool live = gi < B * span;
    const int b = (live && B > 1) ? gi / span : 0;
    const int g = live ? g_lo + gi - b * span : 0;
    const Q8Group q = q8_1_quantize_group<Ops::NEED_S8>(x);
    int mine = __dp4a((int)q.q_lo, 0x01010101, __dp4a((int)q.q_hi, 0x01010101, 0));
    mine += __shfl_xor_sync(0xffffffffu, mine, 1);                 // sum of this 16-value half
    const int other = __shfl_xor_sync(0xffffffffu, mine, 2);
    if (live) {
        const int it = g >> 2, j = g & 3;                          // item, group inside the item
        const int i = 2 * j + (t4 >> 1);                           // 16-value chunk inside the item
        uint8_t *dst = reinterpret_cast<uint8_t *>(act_q + (b * 8 + i) * ipr + it) + 8 * (t4 & 1);
        *reinterpret_cast<uint2 *>(dst) = make_uint2(q.q_lo, q.q_hi);
        if (t4 == 0)
            Ops::store_scales(act_s + (b * Ops::SCF4) * ipr + it, ipr, j, __half2float(__float2half_rn(q.d)), mine, other,
                              __half2float(__float2half_rn(q.sum)));
    }
}

// ------------------------------------------------------------------------------------------------
// The activation image as its own (tiny) launch: long rows, and the fused SwiGLU down projection, whose activation
// quantization would otherwise be repeated by all 148 CTAs of the GEMV (K = 28672 with SwiGLU: ~25 us of a 67 us kernel).
// One quad per 32-value group, 64 groups per CTA; the image is byte for byte what mmvq_sb_kernel / mmvq_rg_kernel keep in
// shared memory ([B][8][ipr] quants | [B][SCF4][ipr] scale table), so the GEMV's prologue becomes one coalesced copy.
// Programmatic dependent launch on both sides: the GEMV's CTAs are resident (and stream weights) while this runs.
// ------------------------------------------------------------------------------------------------
template <int TYPE, int B, typename T>
__global__ void __launch_bounds__(256)
mmvq_sb_prequant_kernel(const void *Xv, const GgbXIn xin, const int K, const int x_vec, uint8_t *image, const int off_scales) {
    pdl_launch_dependents();
    pdl_wait();
    uint4 *act_q = reinterpret_cast<uint4 *>(image);
    float4 *act_s = reinterpret_cast<float4 *>(image + off_scales);
    const T *X = static_cast<const T *>(Xv);
    const int ipr = K >> 7, span = K >> 5, total = B * span;
    const int t4 = threadIdx.x & 3;
    const int gi = blockIdx.x * 64 + (threadIdx.x >> 2);
    float v[8];
    if (gi < total) {
        const int b = (B == 1) ? 0 : gi / span, g = gi - b * span;
        load_act8<T>(X, static_cast<const T *>(xin.x2), xin.ldx, b, g * 32 + t4 * 8, x_vec != 0, v);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 0.0f;
    }
    sb_q_store1<TYPE, B>(act_q, act_s, ipr, v, gi, 0, span, t4);
}

// ------------------------------------------------------------------------------------------------
// the kernel: NW warps, all consumers; every warp streams its own slots (no producer warp, no round trip
// through another thread when a slot is recycled).  NW = 8 with <= 113 KB of shared memory lets TWO CTAs
// share an SM (this launch and the next one); NW = 16 takes the whole SM (large B x K activations).
// ------------------------------------------------------------------------------------------------
template <int TYPE, int B, typename T, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1)
mmvq_sb_kernel(const SbParams p, const int g_defer) {
    using Ops = SbOps<TYPE>;
    extern __shared__ __align__(128) uint8_t smem[];
    uint4 *act_q = reinterpret_cast<uint4 *>(smem);                            // [B][8][ipr]
    float4 *act_s = reinterpret_cast<float4 *>(smem + p.off_scales);           // [B][SCF4][ipr]
    uint8_t *slots = smem + p.off_slots;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.off_bars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nsb = p.nsb, ipr = p.ipr;

    // this CTA's contiguous row range
    const int64_t r_begin = p.N * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = p.N * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int nrows = (int)(r_end - r_begin);
    const int ngroups = (nrows + p.rows_per_slot - 1) / p.rows_per_slot;      // row groups

    if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 0] = gtime();
    if (tid < p.nslots) {
        mbar_init(&full[tid], 1);
        fence_mbar_init();
    }
    __syncthreads();
    pdl_launch_dependents();

    // Warp w owns row groups w, w + warps_used, ... and a private ring of `depth` slots (the first
    // nslots % warps_used warps get one more).  Its k-th slot (row group k / nchunks, chunk k % nchunks)
    // lives in ring slot base + k % depth, phase k / depth: the warp itself refills a slot as soon as it
    // has consumed it, so a barrier is only ever waited on by the thread that armed it, one phase ahead.
    const int wu = p.warps_used;
    const int extra = p.nslots % wu;
    const int depth = p.nslots / wu + (warp < extra ? 1 : 0);
    const int ring_base = warp * (p.nslots / wu) + min(warp, extra);
    const int my_groups = warp < wu ? (ngroups - warp + wu - 1) / wu : 0;
    const int my_slots = my_groups * p.nchunks;
    const uint8_t *wbase = p.W + r_begin * p.row_bytes;
    const uint64_t pol = l2_policy_evict_first();

    // issue state (lane 0): the next slot to stream, kept incrementally -- no divisions on the per-slot path
    int is_k = 0, is_ring = ring_base, is_c = 0, is_row0 = warp * p.rows_per_slot;
    const int row_step = wu * p.rows_per_slot;
    auto issue_next = [&]() {                      // lane 0 only
        const int nvalid = min(p.rows_per_slot, nrows - is_row0);
        uint8_t *dst = slots + is_ring * p.slot_stride;
        if (p.nchunks == 1) {
            const uint32_t bytes = (uint32_t)(nvalid * p.row_bytes);
            mbar_expect_tx(&full[is_ring], bytes);
            bulk_g2s_hint(dst, wbase + (int64_t)is_row0 * p.row_bytes, bytes, &full[is_ring], pol);
            is_row0 += row_step;
        } else {
            const int sb0 = is_c * 16;
            const uint32_t bytes = (uint32_t)(min(16, nsb - sb0) * Ops::BS);
            mbar_expect_tx(&full[is_ring], bytes * (uint32_t)nvalid);
            const uint8_t *src = wbase + (int64_t)is_row0 * p.row_bytes + sb0 * Ops::BS;
            bulk_g2s_hint(dst, src, bytes, &full[is_ring], pol);
            if (nvalid > 1) bulk_g2s_hint(dst + p.chunk_stride, src + p.row_bytes, bytes, &full[is_ring], pol);
            if (++is_c == p.nchunks) { is_c = 0; is_row0 += row_step; }
        }
        is_k++;
        if (++is_ring == ring_base + depth) is_ring = ring_base;
    };

    // While the previous kernel of the stream is still running (programmatic dependent launch made this
    // CTA resident early), pull the head of this CTA's weight range into L2: a pure hint, safe whatever
    // produced W, and it keeps HBM streaming through dependency resolution and the prologue below.
    if (p.l2_prefetch > 0) {
        const int64_t total = min((int64_t)nrows * p.row_bytes, (int64_t)p.l2_prefetch);
        constexpr int kChunk = 2048;
        for (int64_t off = (int64_t)tid * kChunk; off < total; off += (int64_t)NW * 32 * kChunk)
            bulk_prefetch_l2(wbase + off, (uint32_t)min((int64_t)kChunk, total - off));
    }
    // static weights: start streaming before the dependency on the previous kernel resolves
    if (p.pdl_prefetch && lane == 0)
        while (is_k < min(depth, my_slots)) issue_next();
    pdl_wait();
    if (!p.pdl_prefetch && lane == 0)
        while (is_k < min(depth, my_slots)) issue_next();
    if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 1] = gtime();

    // ===================== q8_1-quantize the activations into shared memory =====================
    const T *X = static_cast<const T *>(p.X);
    const int ngroups_row = p.K >> 5;
    const int t4 = lane & 3;
    constexpr int GPI = (NW * 32) / 4;                        // groups per pass of the CTA
    constexpr int U = (B <= 2) ? 4 : 2;                       // passes in flight (independent loads, interleaved math)
    auto q_store1 = [&](float (&x)[8], int gi, int g_lo, int span) { sb_q_store1<TYPE, B>(act_q, act_s, ipr, x, gi, g_lo, span, t4); };
    auto quantize_range = [&](int g_lo, int g_hi) {
        const int span = g_hi - g_lo, total = B * span;
        for (int base = 0; base < total; base += U * GPI) {
            float v[U][8];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int gi = base + u * GPI + (tid >> 2);
                if (gi < total) {
                    const int b = (B == 1) ? 0 : gi / span, g = g_lo + gi - b * span;
                    load_act8<T>(X, static_cast<const T *>(p.xin.x2), p.xin.ldx, b, g * 32 + t4 * 8, p.x_vec != 0, v[u]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[u][i] = 0.0f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (base + u * GPI >= total) break;               // warp-uniform
                q_store1(v[u], base + u * GPI + (tid >> 2), g_lo, span);
            }
        }
    };
    // Long rows, small batches: only the first K-chunk (32 items = 128 groups) is quantized up front; chunk c + 1 is
    // loaded before and quantized after this warp's slot of chunk c of its FIRST row group (below), so 0.9 us instead
    // of 3.3 us of quantization (K = 14336) sit in front of the weight stream.
    constexpr int kChunkGroups = 128;
    const bool staged = p.staged != 0;
    if (p.image != nullptr) {
        // the image was written by the previous launch (mmvq_sb_prequant_kernel): one coalesced copy, L2 hits
        const uint4 *src = reinterpret_cast<const uint4 *>(p.image);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (int i = tid; i < (p.image_bytes >> 4); i += NW * 32) dst[i] = src[i];
    } else {
        quantize_range(0, staged ? min(kChunkGroups, ngroups_row) : ngroups_row);
    }
    __syncthreads();
    if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 2] = gtime();

    // ===================== one slot = this warp's next row group x chunk =====================
    const int lpp = p.lpp;
    const int pair = lane / lpp, sbl = lane - pair * lpp;          // row pair inside the slot, lane inside the pair
    int ring = ring_base;
    uint32_t phase = 0;
    if (p.nchunks == 1 && ipr == 32 && g_defer) {
        // Short rows (K = 4096: one slot = one row pair x all 32 items).  The lane partials of slot g are reduced and
        // stored while slot g + 1 is being computed: the 5-step shuffle chain (~150 cycles of pure latency per slot on
        // a warp that shares its scheduler with one other) overlaps the next slot's shared-memory loads and dp4a chains.
        float prev[2][B];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < B; b++) prev[r][b] = 0.0f;
        for (int g = 0; g <= my_groups; g++) {
            float acc[2][B];
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int b = 0; b < B; b++) acc[r][b] = 0.0f;
            if (g < my_groups) {
                mbar_wait(&full[ring], phase);
                const uint8_t *w0 = slots + ring * p.slot_stride + (lane >> 1) * Ops::BS;
                run_batches<TYPE, B, 0>(w0, w0 + p.row_bytes, lane & 1, act_q + lane, act_s + lane, ipr, acc);
            }
            // reduction + store of the PREVIOUS slot (g - 1): independent of everything above
            const int lr0 = (warp + (g - 1) * wu) * 2;
#pragma unroll
            for (int b = 0; b < B; b++) {
                const bool up = (lane & 16) != 0;
                float v = (up ? prev[1][b] : prev[0][b]) + __shfl_xor_sync(0xffffffffu, up ? prev[0][b] : prev[1][b], 16);
#pragma unroll
                for (int m = 8; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                const int lr = lr0 + (up ? 1 : 0);
                if (g > 0 && (lane & 15) == 0 && lr < nrows) ggb_yout_store<T>(p.yo, b, r_begin + lr, v);
            }
            if (g < my_groups) {
                __syncwarp();
                if (lane == 0 && is_k < my_slots) {    // recycle the slot: every lane has finished reading it
                    fence_proxy_async();
                    issue_next();
                }
                if (++ring == ring_base + depth) { ring = ring_base; phase ^= 1u; }
            }
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int b = 0; b < B; b++) prev[r][b] = acc[r][b];
        }
        if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 3] = gtime();
        return;
    }
    // one slot: chunk c of the current row group
    auto consume_slot = [&](int c, float (&acc)[2][B]) {
        mbar_wait(&full[ring], phase);
        const int it = c * 32 + sbl;
        if (it < ipr) {
            const uint8_t *w0, *w1;
            if (p.nchunks == 1) {
                w0 = slots + ring * p.slot_stride + (2 * pair) * p.row_bytes + (it >> 1) * Ops::BS;
                w1 = w0 + p.row_bytes;
            } else {
                w0 = slots + ring * p.slot_stride + (sbl >> 1) * Ops::BS;
                w1 = w0 + p.chunk_stride;
            }
            run_batches<TYPE, B, 0>(w0, w1, it & 1, act_q + it, act_s + it, ipr, acc);
        }
        __syncwarp();
        if (lane == 0 && is_k < my_slots) {        // recycle the slot: every lane has finished reading it
            fence_proxy_async();
            issue_next();
        }
        if (++ring == ring_base + depth) { ring = ring_base; phase ^= 1u; }
    };
    // reduce over the lanes of each row pair; the first lane of the pair stores
    auto finish_group = [&](int g, float (&acc)[2][B]) {
        const int lr0 = (warp + g * wu) * p.rows_per_slot + 2 * pair;
        if (lpp == 32) {
            // both rows at once: lanes 0-15 end up with row 0, lanes 16-31 with row 1
#pragma unroll
            for (int b = 0; b < B; b++) {
                const bool up = (lane & 16) != 0;
                float v = (up ? acc[1][b] : acc[0][b]) + __shfl_xor_sync(0xffffffffu, up ? acc[0][b] : acc[1][b], 16);
#pragma unroll
                for (int m = 8; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                const int lr = lr0 + (up ? 1 : 0);
                if ((lane & 15) == 0 && lr < nrows) ggb_yout_store<T>(p.yo, b, r_begin + lr, v);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; r++) {
#pragma unroll
                for (int b = 0; b < B; b++) {
                    float v = acc[r][b];
                    for (int m = lpp >> 1; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                    if (sbl == 0 && lr0 + r < nrows) ggb_yout_store<T>(p.yo, b, r_begin + lr0 + r, v);
                }
            }
        }
    };

    int g_first = 0;
    if (staged) {
        // every warp walks the chunks (the quantization and its barriers are CTA-wide); warps that own a row group
        // consume its slots in between
        constexpr int UC = (B * kChunkGroups + GPI - 1) / GPI;    // passes of the CTA that cover one chunk of every batch row
        static_assert(B > 2 || UC <= 4, "a K-chunk of the activations is a few passes");
        float acc[2][B];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < B; b++) acc[r][b] = 0.0f;
        for (int c = 0; c < p.nchunks; c++) {
            const int g_lo = (c + 1) * kChunkGroups, span = min(kChunkGroups, ngroups_row - g_lo);
            const bool more = c + 1 < p.nchunks;
            // chunk c + 1 of X, kept in its storage type while it waits (4 registers per 8 fp16 / bf16 values)
            uint4 raw[UC][sizeof(T) == 4 ? 2 : 1];
            if (more) {
#pragma unroll
                for (int u = 0; u < UC; u++) {
                    const int gi = u * GPI + (tid >> 2);
#pragma unroll
                    for (int h = 0; h < (sizeof(T) == 4 ? 2 : 1); h++) raw[u][h] = make_uint4(0, 0, 0, 0);
                    if (gi < B * span) {
                        const int b = (B == 1) ? 0 : gi / span, g = g_lo + gi - b * span;
                        const T *src = X + (int64_t)b * p.xin.ldx + g * 32 + t4 * 8;
                        raw[u][0] = *reinterpret_cast<const uint4 *>(src);                 // staged mode requires x_vec
                        if (sizeof(T) == 4) raw[u][sizeof(T) == 4 ? 1 : 0] = *reinterpret_cast<const uint4 *>(src + 4);
                    }
                }
            }
            if (my_groups > 0) consume_slot(c, acc);
            if (more) {
#pragma unroll
                for (int u = 0; u < UC; u++) {
                    if (u * GPI >= B * span) break;               // warp-uniform
                    float x[8];
                    if constexpr (sizeof(T) == 4) {
                        const float *f = reinterpret_cast<const float *>(&raw[u][0]);
#pragma unroll
                        for (int i = 0; i < 8; i++) x[i] = f[i];
                    } else {
                        const T *h = reinterpret_cast<const T *>(&raw[u][0]);
#pragma unroll
                        for (int i = 0; i < 8; i++) x[i] = to_f32<T>(h[i]);
                    }
                    q_store1(x, u * GPI + (tid >> 2), g_lo, span);
                }
                __syncthreads();
            }
        }
        if (my_groups > 0) finish_group(0, acc);
        g_first = 1;
    }
    for (int g = g_first; g < my_groups; g++) {
        float acc[2][B];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int b = 0; b < B; b++) acc[r][b] = 0.0f;
        for (int c = 0; c < p.nchunks; c++) consume_slot(c, acc);
        finish_group(g, acc);
    }
    if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 3] = gtime();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
}  // namespace

// tuning state shared by the per-type translation units (defined once in mmvq_sb.cu)
namespace ggb_sb {
extern int g_pdl_prefetch, g_force_warps, g_smem_cap_kb, g_staged, g_defer_reduce, g_l2_prefetch_mb, g_prequant_min_k;
}
unsigned long long *ggb_mmvq_trace_slot();

namespace {
using namespace ggb_sb;

// bytes of the activation image of (TYPE, B, K) = the kernels' shared-memory prefix
template <int TYPE>
int64_t sb_image_bytes(int64_t B, int64_t K) {
    return ((B * K + 15) & ~(int64_t)15) + B * SbOps<TYPE>::SCF4 * (K / 128) * 16;
}

template <int TYPE, int B, typename T>
int launch_prequant(const void *X, int64_t K, void *image, cudaStream_t stream) {
    const GgbXIn xin = ggb_make_xin(K);
    const int x_vec = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(xin.x2) & 15) == 0 &&
                      (xin.ldx * (int64_t)sizeof(T)) % 16 == 0;
    const int groups = (int)(B * (K / 32));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((groups + 63) / 64);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmvq_sb_prequant_kernel<TYPE, B, T>>(ggb_device_info().dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmvq_sb_prequant_kernel<TYPE, B, T>, X, xin, (int)K, x_vec, static_cast<uint8_t *>(image),
                                    (int)((B * K + 15) & ~(int64_t)15)));
    return GGB_OK;
}

template <int TYPE, int B, typename T, int NW>
int launch(const void *W, const void *X, void *Y, int64_t K, int64_t N, const void *image, cudaStream_t stream) {
    using Ops = SbOps<TYPE>;
    constexpr int kThreads = NW * 32;
    const DeviceInfo &di = ggb_device_info();
    SbParams p;
    p.W = static_cast<const uint8_t *>(W);
    p.X = X;
    p.yo = ggb_make_yout(Y, N);
    p.N = N;
    p.K = (int)K;
    p.nsb = (int)(K / 256);
    p.ipr = 2 * p.nsb;
    p.row_bytes = p.nsb * Ops::BS;
    if (p.row_bytes % 16 != 0) return 1;         // bulk copies need 16-byte rows (Q6_K: K % 2048 == 0)
    const bool pow2 = (p.ipr & (p.ipr - 1)) == 0;
    p.lpp = (pow2 && p.ipr < 32) ? p.ipr : 32;
    p.rows_per_slot = 2 * (32 / p.lpp);
    p.nchunks = (p.ipr + 31) / 32;
    if (p.nchunks == 1) {
        p.chunk_stride = 0;
        p.slot_stride = (p.rows_per_slot * p.row_bytes + 127) & ~127;
    } else {
        p.chunk_stride = 16 * Ops::BS;                          // multiple of 16 for every type (BS is even)
        p.slot_stride = (2 * p.chunk_stride + 127) & ~127;
    }
    p.xin = ggb_make_xin(K);
    p.x_vec = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.xin.x2) & 15) == 0 &&
              (p.xin.ldx * (int64_t)sizeof(T)) % 16 == 0;
    // measured (same box, vs quantizing everything up front): Q4_K +3.4 % at K = 14336 and 28672, Q6_K +2.3 % at 14336 but
    // -7.7 % at 28672 (six CTA-wide barriers around heavier slots)
    p.image = static_cast<const uint8_t *>(image);
    p.image_bytes = (int)sb_image_bytes<TYPE>(B, K);
    p.staged = (image == nullptr && p.nchunks > 1 && (TYPE != T_Q6_K || p.nchunks <= 4) && B <= 2 && g_staged && p.x_vec &&
                p.xin.x2 == nullptr) ? 1 : 0;
    p.pdl_prefetch = g_pdl_prefetch;
    p.trace = ggb_mmvq_trace_slot();
    const int act_q = (int)(B * K);
    p.off_scales = (act_q + 15) & ~15;
    p.off_slots = (p.off_scales + B * Ops::SCF4 * p.ipr * 16 + 127) & ~127;
    const int bar_bytes = kMaxSlots * 8;
    int smem_cap = (NW == 8) ? (di.max_smem_optin + 1024) / 2 - 1024 : di.max_smem_optin;   // two CTAs per SM for NW == 8
    if (g_smem_cap_kb > 0) smem_cap = g_smem_cap_kb * 1024 < di.max_smem_optin ? g_smem_cap_kb * 1024 : di.max_smem_optin;
    const int avail = smem_cap - p.off_slots - bar_bytes - 128;
    int nst = avail / p.slot_stride;
    if (nst < (NW == 8 ? 12 : 3)) return 1;      // ring too shallow: next configuration / the caller falls back
    if (nst > kMaxSlots) nst = kMaxSlots;
    p.warps_used = nst < NW ? nst : NW;
    p.nslots = nst;
    p.off_bars = (p.off_slots + nst * p.slot_stride + 127) & ~127;
    const int smem = p.off_bars + bar_bytes;

    static bool attr_set[64] = {};               // per instantiation and device; idempotent, so a race is harmless
    if (!attr_set[di.dev]) {
        GGB_CUDA_TRY(cudaFuncSetAttribute(mmvq_sb_kernel<TYPE, B, T, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          di.max_smem_optin));
        attr_set[di.dev] = true;
    }
    const int64_t max_ctas = (N + 1) / 2;
    const int grid = (int)(max_ctas < di.sm_count ? max_ctas : di.sm_count);
    p.l2_prefetch = (int)(((int64_t)g_l2_prefetch_mb << 20) / grid) & ~15;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmvq_sb_kernel<TYPE, B, T, NW>>(di.dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmvq_sb_kernel<TYPE, B, T, NW>, p, g_defer_reduce));
    return GGB_OK;
}

template <int TYPE, int B, typename T>
int launch_nw(const void *W, const void *X, void *Y, int64_t K, int64_t N, void *ws, int64_t ws_bytes, cudaStream_t stream) {
    // B <= 2: the activation image as its own launch when the caller's workspace holds it and the row is long or the
    // activation is a fused SwiGLU (ggb_mmvq_workspace_bytes asks for it); the GEMV then copies it instead of quantizing
    const void *image = nullptr;
    if constexpr (B <= 2) {
        const GgbXIn xin = ggb_make_xin(K);
        if (ws != nullptr && ws_bytes >= sb_image_bytes<TYPE>(B, K) && (reinterpret_cast<uintptr_t>(ws) & 15) == 0 &&
            g_prequant_min_k > 0 && (K >= g_prequant_min_k || (xin.x2 != nullptr && K >= 4096))) {
            const int rc = launch_prequant<TYPE, B, T>(X, K, ws, stream);
            if (rc != GGB_OK) return rc;
            image = ws;
        }
    }
    // half-SM CTAs (two launches resident per SM) whenever the activations leave room for a useful ring
    if (g_force_warps != 16) {
        const int rc = launch<TYPE, B, T, 8>(W, X, Y, K, N, image, stream);
        if (rc != 1 || g_force_warps == 8) return rc;
    }
    return launch<TYPE, B, T, 16>(W, X, Y, K, N, image, stream);
}

template <int TYPE, typename T>
int dispatch_batch(const void *W, const void *X, void *Y, int64_t B, int64_t K, int64_t N, void *ws, int64_t ws_bytes, cudaStream_t stream) {
    switch (B) {
        case 1: return launch_nw<TYPE, 1, T>(W, X, Y, K, N, ws, ws_bytes, stream);
        case 2: return launch_nw<TYPE, 2, T>(W, X, Y, K, N, ws, ws_bytes, stream);
        case 3: return launch_nw<TYPE, 3, T>(W, X, Y, K, N, ws, ws_bytes, stream);
        case 4: return launch_nw<TYPE, 4, T>(W, X, Y, K, N, ws, ws_bytes, stream);
        default: break;
    }
    // IQ4_XS has no tensor-core path behind it and its codebook lookups fill the register file: measured, batches 5-8
    // are faster on the generic kernel (36 vs 43 us at 14336 x 4096, B = 8), so they are not instantiated here
    if constexpr (TYPE != T_IQ4_XS) {
        switch (B) {
            case 5: return launch_nw<TYPE, 5, T>(W, X, Y, K, N, ws, ws_bytes, stream);
            case 6: return launch_nw<TYPE, 6, T>(W, X, Y, K, N, ws, ws_bytes, stream);
            case 7: return launch_nw<TYPE, 7, T>(W, X, Y, K, N, ws, ws_bytes, stream);
            case 8: return launch_nw<TYPE, 8, T>(W, X, Y, K, N, ws, ws_bytes, stream);
            default: break;
        }
    }
    return 1;
}

template <int TYPE>
int dispatch_dtype(const void *W, const void *X, void *Y, int dt, int64_t B, int64_t K, int64_t N, void *ws, int64_t ws_bytes,
                   cudaStream_t stream) {
    switch (dt) {
        case GGB_F32:  return dispatch_batch<TYPE, float>(W, X, Y, B, K, N, ws, ws_bytes, stream);
        case GGB_F16:  return dispatch_batch<TYPE, __half>(W, X, Y, B, K, N, ws, ws_bytes, stream);
        case GGB_BF16: return dispatch_batch<TYPE, __nv_bfloat16>(W, X, Y, B, K, N, ws, ws_bytes, stream);
        default: return 1;
    }
}

}  // namespace
