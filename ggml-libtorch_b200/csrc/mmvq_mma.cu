// mmvq_mma.cu -- MMVQ for decode batches 3..8: the integer dots run on the tensor cores (mma.sync s8).
//
// Same contract as mmvq.cu / mmvq_sb.cu (ggml_mul_mat_vec_a8 = quantize_row_q8_1_cuda + mul_mat_vec_q +
// vec_dot_*_q8_1; hf-kernels/ggml-kernels/ggml/ggml_kernel.cu:13-66,80-193, ggml/mmvq.cuh:2-209,
// ggml/vecdotq.cuh:272-349).  The dp4a kernels cost ~4.4 us per batch row on a Llama-8B matrix: above two
// rows they are instruction-bound, not HBM-bound.  Here one m16n8k32 MMA takes the 32-value integer dots of
// SIXTEEN weight rows against EIGHT tokens at once, so the cost is the same for every batch size <= 8.
//
// Two launches chained by programmatic dependent launch:
//   (1) mmvq_act_image_kernel: q8_1-quantizes the B x K activations ONCE (bit-identical to quantize_q8_1,
//       q8_1.cuh) into the caller's workspace, laid out exactly as the GEMV kernel wants it in shared memory.
//       (In-kernel quantization, as mmvq_sb.cu does for batches 1-2, repeats the work in all 148 CTAs: at
//       B = 8 that was as many instructions as the GEMV itself.)
//   (2) mmvq_mma_kernel, one CTA per SM:
//   * A operand = weights (rows q = lane / 4 and q + 8), B operand = q8_1 activations (token q), both read
//     straight from shared memory as 64-bit words: the dot product does not care about the order of k, so
//     lane t = lane % 4 simply owns bytes 8t..8t+7 of each 32-value group on both sides (no ldmatrix, no
//     transposes).  C gives lane (q, t) the dots of rows q, q + 8 against tokens 2t, 2t + 1.
//   * Scales stay integers as long as possible: float(sc * dot) comes out of ONE IMAD (the product lands in
//     the mantissa of 1.5 * 2^23) + one FADD, then one FFMA with the token's d8; d is applied once per
//     super-block.  The 6-bit scales of a super-block are decoded once per warp (lane (q, t) decodes groups
//     t, t + 4 of rows q, q + 8 and parks them in a per-warp scratch).
//   * Minima: sum_g m[row, g] * (d8 * sum q8)[token, g] is a 16 x 8 x 8 GEMM per super-block: two tf32
//     m16n8k8 MMAs (the fp32 activation factor split in tf32 hi + lo, m is exact), whose A fragment is
//     exactly the (row, group) pairs the lane decoded itself.
//   * A CTA owns a contiguous range of rows, cut into row groups of 16.  S = 2^k warps (a team) share a row
//     group and split K in chunks of `cs` super-blocks (warp s takes chunks s, s + S, ...); the team's
//     partial sums meet in shared memory and are added in warp order, so the summation order depends on K,
//     the batch and the type only -- never on N or the grid: N-sharded results stay bit-identical.
//   * Weights stream through per-warp rings of (16 rows x chunk) slots, ONE 2-D tiled TMA load per slot (a
//     bulk copy per row costs ~35 instructions each: UBLKCP takes uniform operands), L2 evict-first; the
//     activation image arrives by bulk copies; programmatic dependent launch and the L2 prefetch of the
//     weight range are those of mmvq_sb.cu.
#include "q8_1.cuh"
#include "tma2d.cuh"

namespace {

constexpr int kMaxSlots = 128;
constexpr int kRowsPerGroup = 16;
constexpr int kScrWords = 16 * 8;       // per-warp scratch buffer: 16 rows x {sc[t], sc[t+4]} x 4 lanes (int32)

// Activation image (workspace and shared memory, same layout):
//   [0, B * act_stride)             int8 quants, token-major, act_stride = K + 32 (conflict-free 64-bit reads)
//   [off_d, ...)   float2 per (token pair p, group G): {d8[2p][G], d8[2p+1][G]}, pair stride d_stride2 float2
//   [off_s, ...)   float2 per (token b, group G): tf32 split {hi, lo} of d8 * sum(q8); token stride s_stride2.
//                  Q6_K (s16): per 16-value sub-block instead, of 32 * d8 * sum(q8 of the sub-block)
struct ActLayout {
    int act_stride, off_d, d_stride2, off_s, s_stride2, total;
};
__host__ __device__ inline ActLayout act_layout(int B, int K, bool s16) {
    ActLayout L;
    const int ng = K / 32, npairs = (B + 1) / 2;
    L.act_stride = K + 32;
    L.off_d = (B * L.act_stride + 15) & ~15;
    L.d_stride2 = ng + 2;                                   // + 16 B: token pairs land in different banks
    L.off_s = (L.off_d + npairs * L.d_stride2 * 8 + 15) & ~15;
    L.s_stride2 = (s16 ? 2 * ng : ng) + 2;
    L.total = (L.off_s + B * L.s_stride2 * 8 + 15) & ~15;
    return L;
}

struct MmaParams {
    const uint8_t *act_img;  // workspace: the activation image written by mmvq_act_image_kernel
    GgbYOut yo;
    int64_t N;
    int K, B;
    int nsb;               // super-blocks per row
    int cs;                // super-blocks per chunk
    int nchunks;
    int s_log2;            // log2 of the warps per team
    int slot_row_stride;   // bytes between rows of a slot = chunk bytes (the TMA box is dense)
    int slot_stride;
    int depth;             // ring slots per warp
    int accy_rgs;          // row groups per team in the phase accumulator (P > 1)
    int phases;            // K phases: the shared-memory image holds K / phases columns at a time (1: the whole row)
    int off_scr, off_part, off_accy, off_slots, off_bars;
    int row_bytes;
    unsigned long long *trace;   // debug: per-CTA globaltimer stamps {start, dependency resolved, activations in, end}
    int pdl_prefetch;
    // inline_q != 0: no image launch -- the two CTAs of a cluster quantize half of the B x K activations each and write their
    // groups into BOTH CTAs' shared-memory image (st.async, bytes counted on the destination's mbarrier)
    int inline_q;
    const void *X;
    GgbXIn xin;
    int x_vec;
};

__device__ __forceinline__ uint4 lds128(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ uint2 lds64(const void *p) { return *reinterpret_cast<const uint2 *>(p); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// thread-block cluster helpers (the pair that shares one activation quantization; same scheme as mmvq_rg.cu)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_v2(uint32_t addr, uint32_t x, uint32_t y, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                 ::"r"(addr), "r"(x), "r"(y), "r"(mbar) : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t addr, uint32_t x, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(addr), "r"(x), "r"(mbar) : "memory");
}

constexpr int kMagicBits = 0x4B400000;              // 1.5 * 2^23
constexpr float kMagic = 12582912.0f;
// exact (float)(a * b): for |a * b| < 2^22 (Q4_K: 32 * 15 * 127 * 63 = 3.84 M) the product lands in the mantissa of
// 1.5 * 2^23 and one FADD extracts it; Q5_K (up to 7.9 M) takes the conversion instruction
template <int TYPE>
__device__ __forceinline__ float imul_to_f32(int a, int b) {
    if constexpr (TYPE == T_Q4_K) return __int_as_float(a * b + kMagicBits) - kMagic;
    else return __int2float_rn(a * b);
}

// int8 dots of 16 weight rows (A: a0/a2 = row q, a1/a3 = row q + 8) against 8 tokens (B) over 32 values
__device__ __forceinline__ void mma_s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
        : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}
// tf32 16 x 8 x 8, accumulating
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------
// launch 1: the activation image
// ------------------------------------------------------------------------------------------------
// SMODE: what the per-group table at off_s holds -- 0: tf32 {hi, lo} of d8 * sum(q8); 1: the same per 16-value sub-block, times
// 32 (Q6_K); 2: {s, 0} with s = the fp16-rounded sum of the raw inputs, block_q8_1.s (the legacy types' offsets and minima go
// through it: vecdotq.cuh:45-162)
template <typename T, int SMODE>
__global__ void __launch_bounds__(256)
mmvq_act_image_kernel(const T *__restrict__ X, const GgbXIn xin, uint8_t *__restrict__ img, int B, int K, int x_vec) {
    pdl_launch_dependents();
    pdl_wait();                                             // X comes from the previous kernel of the stream
    constexpr bool S16 = SMODE == 1;
    const ActLayout L = act_layout(B, K, S16);
    const int ng = K >> 5;
    const int t4 = threadIdx.x & 3;
    const int gi = blockIdx.x * 64 + (threadIdx.x >> 2);    // group index over 2 * npairs tokens (warp-uniform validity per 8 groups)
    const int ntok = ((B + 1) / 2) * 2;
    if (gi >= ntok * ng) return;                            // ng is a multiple of 8: whole warps leave together
    const int b = gi / ng, g = gi - b * ng;
    float v[8];
    if (b < B) {
        load_act8<T>(X, static_cast<const T *>(xin.x2), xin.ldx, b, g * 32 + t4 * 8, x_vec != 0, v);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = 0.0f;            // the missing token of an odd batch: zero scales
    }
    const Q8Group qg = q8_1_quantize_group<SMODE == 2>(v);
    int sum = __dp4a((int)qg.q_lo, 0x01010101, __dp4a((int)qg.q_hi, 0x01010101, 0));
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    const int sum16 = sum;                                   // this lane pair's 16-value half of the group
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (b < B) *reinterpret_cast<uint2 *>(img + b * L.act_stride + g * 32 + t4 * 8) = make_uint2(qg.q_lo, qg.q_hi);
    const float d8 = __half2float(__float2half_rn(qg.d));
    if (t4 == 0) reinterpret_cast<float *>(img + L.off_d)[((b >> 1) * L.d_stride2 + g) * 2 + (b & 1)] = d8;
    if (b < B) {
        if (S16) {
            if ((t4 & 1) == 0) {
                const float s8 = d8 * (float)(32 * sum16);
                const float hi = __uint_as_float(__float_as_uint(s8) & 0xffffe000u);
                reinterpret_cast<float2 *>(img + L.off_s)[b * L.s_stride2 + 2 * g + (t4 >> 1)] = make_float2(hi, s8 - hi);
            }
        } else if (SMODE == 2) {
            if (t4 == 0)
                reinterpret_cast<float2 *>(img + L.off_s)[b * L.s_stride2 + g] = make_float2(__half2float(__float2half_rn(qg.sum)), 0.0f);
        } else if (t4 == 0) {
            const float s8 = d8 * (float)sum;
            const float hi = __uint_as_float(__float_as_uint(s8) & 0xffffe000u);
            reinterpret_cast<float2 *>(img + L.off_s)[b * L.s_stride2 + g] = make_float2(hi, s8 - hi);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-type super-block step.  r0 / r1 = the super-block of weight rows q / q + 8 in the slot; ap = int8
// activations of token q for this super-block (+ 8t); dp = {d8, d8} pairs of tokens 2t, 2t+1; sp = {hi, lo}
// of token q; scr = this warp's scratch buffer for this step.
// ------------------------------------------------------------------------------------------------
template <int TYPE> struct MmaOps;

// Q4_K / Q5_K header: {d, dmin, 12 bytes of 6-bit (scale, min)} (dequantize.cuh:154-161).  Lane t decodes
// groups t and t + 4.
struct K4Head {
    float d, dmin;
    int sc_a, sc_b;        // scales of groups t, t + 4
    uint32_t m_a, m_b;     // minima of groups t, t + 4 as tf32 (exact)
};
__device__ __forceinline__ K4Head k4_head(const uint8_t *row, int t) {
    const uint4 hd = lds128(row);
    const float2 dm = __half22float2(*reinterpret_cast<const __half2 *>(&hd.x));
    const uint32_t sh = 8u * t;
    const uint32_t s0 = (hd.y >> sh) & 0xffu, s1 = (hd.z >> sh) & 0xffu, s2 = (hd.w >> sh) & 0xffu;
    K4Head h;
    h.d = dm.x;
    h.dmin = dm.y;
    h.sc_a = (int)(s0 & 63u);
    h.sc_b = (int)((s2 & 15u) | ((s0 >> 6) << 4));
    h.m_a = __float_as_uint((float)(s1 & 63u));
    h.m_b = __float_as_uint((float)((s2 >> 4) | ((s1 >> 6) << 4)));
    return h;
}

template <int TYPE>
__device__ __forceinline__ void k45_step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                         const float2 *sp, int *scr, int q, int t, float (&acc)[4]) {
    constexpr int QS = (TYPE == T_Q4_K) ? 16 : 48;          // offset of the nibbles
    ap += 8 * t;
    const K4Head h0 = k4_head(r0, t), h1 = k4_head(r1, t);
    // scales: row-major scratch, 8 ints per row ordered {sc[t], sc[t+4]} by t
    *reinterpret_cast<int2 *>(scr + q * 8 + 2 * t) = make_int2(h0.sc_a, h0.sc_b);
    *reinterpret_cast<int2 *>(scr + (q + 8) * 8 + 2 * t) = make_int2(h1.sc_a, h1.sc_b);
    // minima: C[row, token] += m[row, g] * s8[token, g] over the 8 groups (tf32 hi + lo)
    float cm[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    {
        const float2 sa = sp[t], sb = sp[t + 4];
        mma_tf32(cm, h0.m_a, h1.m_a, h0.m_b, h1.m_b, __float_as_uint(sa.x), __float_as_uint(sb.x));
        mma_tf32(cm, h0.m_a, h1.m_a, h0.m_b, h1.m_b, __float_as_uint(sa.y), __float_as_uint(sb.y));
    }
    __syncwarp();
    int sc0[8], sc1[8];                                      // sc0[2j] = group j, sc0[2j+1] = group j + 4 (row q); sc1: row q + 8
    {
        const int4 a = *reinterpret_cast<const int4 *>(scr + q * 8), b = *reinterpret_cast<const int4 *>(scr + q * 8 + 4);
        const int4 c = *reinterpret_cast<const int4 *>(scr + (q + 8) * 8), d = *reinterpret_cast<const int4 *>(scr + (q + 8) * 8 + 4);
        sc0[0] = a.x; sc0[1] = a.y; sc0[2] = a.z; sc0[3] = a.w; sc0[4] = b.x; sc0[5] = b.y; sc0[6] = b.z; sc0[7] = b.w;
        sc1[0] = c.x; sc1[1] = c.y; sc1[2] = c.z; sc1[3] = c.w; sc1[4] = d.x; sc1[5] = d.y; sc1[6] = d.z; sc1[7] = d.w;
    }
    uint2 qh0, qh1;
    if constexpr (TYPE == T_Q5_K) {                         // fifth bits: bit g of byte j belongs to group g
        qh0 = lds64(r0 + 16 + 8 * t);
        qh1 = lds64(r1 + 16 + 8 * t);
    }
    float fs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < 4; i++) {                           // groups 2i (low nibbles) and 2i + 1 (high nibbles)
        const uint2 w0 = lds64(r0 + QS + 32 * i + 8 * t), w1 = lds64(r1 + QS + 32 * i + 8 * t);
        const uint2 b0 = lds64(ap + 64 * i), b1 = lds64(ap + 64 * i + 32);
        const float4 d8 = *reinterpret_cast<const float4 *>(dp + 2 * i);      // {d8a[2i], d8b[2i], d8a[2i+1], d8b[2i+1]}
        uint32_t l0x = w0.x & 0x0f0f0f0fu, l1x = w1.x & 0x0f0f0f0fu, l0y = w0.y & 0x0f0f0f0fu, l1y = w1.y & 0x0f0f0f0fu;
        uint32_t h0x = (w0.x >> 4) & 0x0f0f0f0fu, h1x = (w1.x >> 4) & 0x0f0f0f0fu;
        uint32_t h0y = (w0.y >> 4) & 0x0f0f0f0fu, h1y = (w1.y >> 4) & 0x0f0f0f0fu;
        if constexpr (TYPE == T_Q5_K) {
            l0x |= ((qh0.x >> (2 * i)) & 0x01010101u) << 4;      l1x |= ((qh1.x >> (2 * i)) & 0x01010101u) << 4;
            l0y |= ((qh0.y >> (2 * i)) & 0x01010101u) << 4;      l1y |= ((qh1.y >> (2 * i)) & 0x01010101u) << 4;
            h0x |= ((qh0.x >> (2 * i + 1)) & 0x01010101u) << 4;  h1x |= ((qh1.x >> (2 * i + 1)) & 0x01010101u) << 4;
            h0y |= ((qh0.y >> (2 * i + 1)) & 0x01010101u) << 4;  h1y |= ((qh1.y >> (2 * i + 1)) & 0x01010101u) << 4;
        }
        int cl[4], ch[4];
        mma_s8(cl, l0x, l1x, l0y, l1y, b0.x, b0.y);
        mma_s8(ch, h0x, h1x, h0y, h1y, b1.x, b1.y);
        // group g lives at scratch index 2 * (g & 3) + (g >> 2)
        const int g0 = 2 * i, g1 = 2 * i + 1;
        const int i0 = 2 * (g0 & 3) + (g0 >> 2), i1 = 2 * (g1 & 3) + (g1 >> 2);
        fs[0] = fmaf(d8.x, imul_to_f32<TYPE>(cl[0], sc0[i0]), fs[0]);
        fs[1] = fmaf(d8.y, imul_to_f32<TYPE>(cl[1], sc0[i0]), fs[1]);
        fs[2] = fmaf(d8.x, imul_to_f32<TYPE>(cl[2], sc1[i0]), fs[2]);
        fs[3] = fmaf(d8.y, imul_to_f32<TYPE>(cl[3], sc1[i0]), fs[3]);
        fs[0] = fmaf(d8.z, imul_to_f32<TYPE>(ch[0], sc0[i1]), fs[0]);
        fs[1] = fmaf(d8.w, imul_to_f32<TYPE>(ch[1], sc0[i1]), fs[1]);
        fs[2] = fmaf(d8.z, imul_to_f32<TYPE>(ch[2], sc1[i1]), fs[2]);
        fs[3] = fmaf(d8.w, imul_to_f32<TYPE>(ch[3], sc1[i1]), fs[3]);
    }
    acc[0] += h0.d * fs[0] - h0.dmin * cm[0];
    acc[1] += h0.d * fs[1] - h0.dmin * cm[1];
    acc[2] += h1.d * fs[2] - h1.dmin * cm[2];
    acc[3] += h1.d * fs[3] - h1.dmin * cm[3];
}

template <> struct MmaOps<T_Q4_K> {
    static constexpr int BS = 144;
    static constexpr int SMODE = 0;
    __device__ __forceinline__ static void step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                const float2 *sp, int *scr, int q, int t, float (&acc)[4]) {
        k45_step<T_Q4_K>(r0, r1, ap, dp, sp, scr, q, t, acc);
    }
};
template <> struct MmaOps<T_Q5_K> {
    static constexpr int BS = 176;
    static constexpr int SMODE = 0;
    __device__ __forceinline__ static void step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                const float2 *sp, int *scr, int q, int t, float (&acc)[4]) {
        k45_step<T_Q5_K>(r0, r1, ap, dp, sp, scr, q, t, acc);
    }
};

// ---- IQ4_XS: d @0, scales_h @2 (2 bits x 8), scales_l @4 (4 bits x 8), 128 bytes of 4-bit codebook indices @8; 136-byte
// blocks are 8-byte aligned.  Value = d * (ls - 32) * kvalues_iq4nl[code] (dequantize.cuh:419-433, vecdotq.cuh:867-888).
// The codebook values ARE int8, so after the register lookup (iq4nl_lut8: eight codes per raw word) they go straight into
// the s8 MMA: lane t owns bytes 4t..4t+3 of the sub-block's 16 code bytes = values 4t..4t+3 (low nibbles, a0/a1) and
// 16+4t..16+4t+3 (high nibbles, a2/a3), and reads the matching activation bytes for B.  No minima, no offset.
template <> struct MmaOps<T_IQ4_XS> {
    static constexpr int BS = 136;
    static constexpr int SMODE = 0;
    __device__ __forceinline__ static float scale_of(uint2 hd, int ib) {
        const int ls = (int)(((hd.y >> (4 * ib)) & 0xfu) | (((hd.x >> (16 + 2 * ib)) & 3u) << 4)) - 32;
        return __int_as_float(ls + kMagicBits) - kMagic;     // exact
    }
    __device__ __forceinline__ static void step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                const float2 *, int *scr, int q, int t, float (&acc)[4]) {
        const uint2 h0 = lds64(r0), h1 = lds64(r1);
        // lane t decodes sub-blocks t, t + 4 of rows q, q + 8 into the warp's scratch (row-major, {s[t], s[t+4]} by t)
        float *fscr = reinterpret_cast<float *>(scr);
        *reinterpret_cast<float2 *>(fscr + q * 8 + 2 * t) = make_float2(scale_of(h0, t), scale_of(h0, t + 4));
        *reinterpret_cast<float2 *>(fscr + (q + 8) * 8 + 2 * t) = make_float2(scale_of(h1, t), scale_of(h1, t + 4));
        __syncwarp();
        float s0[8], s1[8];                                      // s0[2j] = sub-block j, s0[2j+1] = sub-block j + 4 (row q); s1: row q + 8
        {
            const float4 a = *reinterpret_cast<const float4 *>(fscr + q * 8), b = *reinterpret_cast<const float4 *>(fscr + q * 8 + 4);
            const float4 c = *reinterpret_cast<const float4 *>(fscr + (q + 8) * 8), d = *reinterpret_cast<const float4 *>(fscr + (q + 8) * 8 + 4);
            s0[0] = a.x; s0[1] = a.y; s0[2] = a.z; s0[3] = a.w; s0[4] = b.x; s0[5] = b.y; s0[6] = b.z; s0[7] = b.w;
            s1[0] = c.x; s1[1] = c.y; s1[2] = c.z; s1[3] = c.w; s1[4] = d.x; s1[5] = d.y; s1[6] = d.z; s1[7] = d.w;
        }
        float fs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int ib = 0; ib < 8; ib++) {
            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(r0 + 8 + 16 * ib + 4 * t);
            const uint32_t w1 = *reinterpret_cast<const uint32_t *>(r1 + 8 + 16 * ib + 4 * t);
            const uint32_t b0 = *reinterpret_cast<const uint32_t *>(ap + 32 * ib + 4 * t);
            const uint32_t b1 = *reinterpret_cast<const uint32_t *>(ap + 32 * ib + 16 + 4 * t);
            uint32_t a0, a1, a2, a3;
            iq4nl_lut8(w0, a0, a2);
            iq4nl_lut8(w1, a1, a3);
            int c[4];
            mma_s8(c, a0, a1, a2, a3, b0, b1);
            const float2 d8 = dp[ib];
            const int si = 2 * (ib & 3) + (ib >> 2);
            // |dot| <= 32 * 127 * 127 < 2^22: exact int -> float through the mantissa of 1.5 * 2^23
            fs[0] = fmaf(__int_as_float(c[0] + kMagicBits) - kMagic, d8.x * s0[si], fs[0]);
            fs[1] = fmaf(__int_as_float(c[1] + kMagicBits) - kMagic, d8.y * s0[si], fs[1]);
            fs[2] = fmaf(__int_as_float(c[2] + kMagicBits) - kMagic, d8.x * s1[si], fs[2]);
            fs[3] = fmaf(__int_as_float(c[3] + kMagicBits) - kMagic, d8.y * s1[si], fs[3]);
        }
        const float d0 = ggb_h2f(h0.x & 0xffffu), d1 = ggb_h2f(h1.x & 0xffffu);
        acc[0] += d0 * fs[0];
        acc[1] += d0 * fs[1];
        acc[2] += d1 * fs[2];
        acc[3] += d1 * fs[3];
    }
};

// ---- Q6_K: ql[128] @0, qh[64] @128, 16 int8 scales @192 (one per 16 values), d @208; 210-byte blocks are only
// 2-byte aligned.  Value = d * sc * (q6 - 32): the weights go to the MMA as the unsigned q6 (one m16n8k16 per 16-value
// sub-block), and the -32 becomes  - d * sum_sub sc[row, sub] * (32 * d8 * sum q8)[token, sub]: a 16 x 8 x 16 GEMM per
// super-block on the tf32 MMA, like the minima of Q4_K.   (dequantize.cuh:230-254, vecdotq.cuh:331-349)
__device__ __forceinline__ void mma_u8s8_k16(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k16.row.col.s32.u8.s8.s32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %7, %7, %7};"
        : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3])
        : "r"(a0), "r"(a1), "r"(b0), "r"(0));
}
template <bool AL>
__device__ __forceinline__ uint32_t ldw_a2(const uint8_t *p) {          // 32-bit word at a 2-byte-aligned address
    if (AL) return *reinterpret_cast<const uint32_t *>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - 2);
    return __funnelshift_r(w[0], w[1], 16);
}
template <int J>
__device__ __forceinline__ int sbyte_of(uint32_t w) {                   // sign-extended byte J
    return (int)(int8_t)(w >> (8 * J));
}

template <bool AL>
__device__ __forceinline__ void q6_step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                        const float2 *sp, int t, float (&acc)[4]) {
    const uint8_t *rr[2] = {r0, r1};
    const uint32_t sh_t = 8u * t;
    float cm[4] = {0.0f, 0.0f, 0.0f, 0.0f}, fs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int ip = 0; ip < 2; ip++) {                         // the two 128-value halves: sub-blocks 8ip .. 8ip + 7
        int sc[2][8];
        uint32_t a_lo[2], a_hi[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const uint32_t w0 = ldw_a2<AL>(rr[r] + 192 + 8 * ip), w1 = ldw_a2<AL>(rr[r] + 196 + 8 * ip);
            sc[r][0] = sbyte_of<0>(w0); sc[r][1] = sbyte_of<1>(w0); sc[r][2] = sbyte_of<2>(w0); sc[r][3] = sbyte_of<3>(w0);
            sc[r][4] = sbyte_of<0>(w1); sc[r][5] = sbyte_of<1>(w1); sc[r][6] = sbyte_of<2>(w1); sc[r][7] = sbyte_of<3>(w1);
            // tf32 A fragment of this k8 step: columns t and t + 4 = sub-blocks 8ip + t, 8ip + t + 4 (exact: |sc| <= 128)
            a_lo[r] = __float_as_uint((float)(int)(int8_t)(w0 >> sh_t));
            a_hi[r] = __float_as_uint((float)(int)(int8_t)(w1 >> sh_t));
        }
        {
            const float2 sa = sp[8 * ip + t], sb = sp[8 * ip + t + 4];
            mma_tf32(cm, a_lo[0], a_lo[1], a_hi[0], a_hi[1], __float_as_uint(sa.x), __float_as_uint(sb.x));
            mma_tf32(cm, a_lo[0], a_lo[1], a_hi[0], a_hi[1], __float_as_uint(sa.y), __float_as_uint(sb.y));
        }
        // w[row][u][r]: this lane's 4 values (bytes 16u + 4t ..) of sub-block 8ip + 2r + u
        uint32_t w[2][2][4];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint32_t l0 = ldw_a2<AL>(rr[r] + 64 * ip + 16 * u + 4 * t);
                const uint32_t l1 = ldw_a2<AL>(rr[r] + 64 * ip + 32 + 16 * u + 4 * t);
                const uint32_t h = ldw_a2<AL>(rr[r] + 128 + 32 * ip + 16 * u + 4 * t);
                w[r][u][0] = (l0 & 0x0f0f0f0fu) | ((h << 4) & 0x30303030u);
                w[r][u][1] = (l1 & 0x0f0f0f0fu) | ((h << 2) & 0x30303030u);
                w[r][u][2] = ((l0 >> 4) & 0x0f0f0f0fu) | (h & 0x30303030u);
                w[r][u][3] = ((l1 >> 4) & 0x0f0f0f0fu) | ((h >> 2) & 0x30303030u);
            }
#pragma unroll
        for (int r4 = 0; r4 < 4; r4++) {                     // 32-value group 4ip + r4 = sub-blocks 2 r4, 2 r4 + 1 of this half
            const uint8_t *a = ap + 16 * (8 * ip + 2 * r4) + 4 * t;
            const uint32_t ba = *reinterpret_cast<const uint32_t *>(a), bb = *reinterpret_cast<const uint32_t *>(a + 16);
            int ca[4], cb[4];
            mma_u8s8_k16(ca, w[0][0][r4], w[1][0][r4], ba);
            mma_u8s8_k16(cb, w[0][1][r4], w[1][1][r4], bb);
            const float2 d8 = dp[4 * ip + r4];
            fs[0] = fmaf(d8.x, __int2float_rn(ca[0] * sc[0][2 * r4] + cb[0] * sc[0][2 * r4 + 1]), fs[0]);
            fs[1] = fmaf(d8.y, __int2float_rn(ca[1] * sc[0][2 * r4] + cb[1] * sc[0][2 * r4 + 1]), fs[1]);
            fs[2] = fmaf(d8.x, __int2float_rn(ca[2] * sc[1][2 * r4] + cb[2] * sc[1][2 * r4 + 1]), fs[2]);
            fs[3] = fmaf(d8.y, __int2float_rn(ca[3] * sc[1][2 * r4] + cb[3] * sc[1][2 * r4 + 1]), fs[3]);
        }
    }
    const float d0 = ggb_h2f(ld_u16(r0 + 208)), d1 = ggb_h2f(ld_u16(r1 + 208));
    acc[0] += d0 * (fs[0] - cm[0]);
    acc[1] += d0 * (fs[1] - cm[1]);
    acc[2] += d1 * (fs[2] - cm[2]);
    acc[3] += d1 * (fs[3] - cm[3]);
}

// ---- IQ4_NL: 18-byte blocks {d, 16 bytes of 4-bit codebook indices} (dequantize.cuh:400-416, vecdotq.cuh:842-864).  Eight
// blocks = 144 bytes are treated as the kernel's 256-value unit; a block's code words are 2-byte aligned (odd blocks
// 4-byte aligned), its fp16 d is the sub-block scale.  Otherwise IQ4_XS's step.
template <> struct MmaOps<T_IQ4_NL> {
    static constexpr int BS = 144;
    static constexpr int SMODE = 0;
    __device__ __forceinline__ static void step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                const float2 *, int *scr, int q, int t, float (&acc)[4]) {
        float *fscr = reinterpret_cast<float *>(scr);
        *reinterpret_cast<float2 *>(fscr + q * 8 + 2 * t) = make_float2(ggb_h2f(ld_u16(r0 + 18 * t)), ggb_h2f(ld_u16(r0 + 18 * (t + 4))));
        *reinterpret_cast<float2 *>(fscr + (q + 8) * 8 + 2 * t) = make_float2(ggb_h2f(ld_u16(r1 + 18 * t)), ggb_h2f(ld_u16(r1 + 18 * (t + 4))));
        __syncwarp();
        float s0[8], s1[8];                                      // s0[2j] = block j, s0[2j+1] = block j + 4 (row q); s1: row q + 8
        {
            const float4 a = *reinterpret_cast<const float4 *>(fscr + q * 8), b = *reinterpret_cast<const float4 *>(fscr + q * 8 + 4);
            const float4 c = *reinterpret_cast<const float4 *>(fscr + (q + 8) * 8), d = *reinterpret_cast<const float4 *>(fscr + (q + 8) * 8 + 4);
            s0[0] = a.x; s0[1] = a.y; s0[2] = a.z; s0[3] = a.w; s0[4] = b.x; s0[5] = b.y; s0[6] = b.z; s0[7] = b.w;
            s1[0] = c.x; s1[1] = c.y; s1[2] = c.z; s1[3] = c.w; s1[4] = d.x; s1[5] = d.y; s1[6] = d.z; s1[7] = d.w;
        }
        float fs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int ib = 0; ib < 8; ib++) {
            const uint32_t w0 = (ib & 1) ? ldw_a2<true>(r0 + 18 * ib + 2 + 4 * t) : ldw_a2<false>(r0 + 18 * ib + 2 + 4 * t);
            const uint32_t w1 = (ib & 1) ? ldw_a2<true>(r1 + 18 * ib + 2 + 4 * t) : ldw_a2<false>(r1 + 18 * ib + 2 + 4 * t);
            const uint32_t b0 = *reinterpret_cast<const uint32_t *>(ap + 32 * ib + 4 * t);
            const uint32_t b1 = *reinterpret_cast<const uint32_t *>(ap + 32 * ib + 16 + 4 * t);
            uint32_t a0, a1, a2, a3;
            iq4nl_lut8(w0, a0, a2);
            iq4nl_lut8(w1, a1, a3);
            int c[4];
            mma_s8(c, a0, a1, a2, a3, b0, b1);
            const float2 d8 = dp[ib];
            const int si = 2 * (ib & 3) + (ib >> 2);
            fs[0] = fmaf(__int_as_float(c[0] + kMagicBits) - kMagic, d8.x * s0[si], fs[0]);
            fs[1] = fmaf(__int_as_float(c[1] + kMagicBits) - kMagic, d8.y * s0[si], fs[1]);
            fs[2] = fmaf(__int_as_float(c[2] + kMagicBits) - kMagic, d8.x * s1[si], fs[2]);
            fs[3] = fmaf(__int_as_float(c[3] + kMagicBits) - kMagic, d8.y * s1[si], fs[3]);
        }
        acc[0] += fs[0];
        acc[1] += fs[1];
        acc[2] += fs[2];
        acc[3] += fs[3];
    }
};

template <> struct MmaOps<T_Q6_K> {
    static constexpr int BS = 210;
    static constexpr int SMODE = 1;
    __device__ __forceinline__ static void step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                const float2 *sp, int *, int, int t, float (&acc)[4]) {
        if ((reinterpret_cast<uintptr_t>(r0) & 2) == 0) q6_step<true>(r0, r1, ap, dp, sp, t, acc);     // warp-uniform
        else q6_step<false>(r0, r1, ap, dp, sp, t, acc);
    }
};


// ---- Q2_K and Q3_K (round 2): the 2-bit formats on the Q6_K scheme -- one m16n8k16 u8 x s8 MMA per 16-value sub-block, the
// per-sub-block factors as integers, minima / offsets through the tf32 MMA on the S16 table (32 * d8 * sum q8 per sub-block).
//   Q2_K: scales[16] @0 (4-bit scale | 4-bit minimum), qs[64] @16, d @80, dmin @82 (84-byte super-blocks, 4-byte aligned):
//         d * sum_k sc_k * dot_k  -  dmin * sum_k m_k * (d8 * sum q8)_k              (dequantize.cuh:107-120, vecdotq.cuh:203-231)
//   Q3_K: hmask[32] @0, qs[64] @32, 12 scale bytes @96, d @108 (110-byte super-blocks, 2-byte aligned): weights q2 | hbit << 2 in
//         0..7, d * sum_k (sc_k - 32) * (dot_k - 4 * sum q8_k)                      (dequantize.cuh:129-151, vecdotq.cuh:235-268)
// Lane t owns bytes 4t..4t+3 of every 16-byte run; half ip = values 128 ip .., group j at bit 2 j of the qs byte.
template <int TYPE, bool AL>
__device__ __forceinline__ void q23_step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                         const float2 *sp, int t, float (&acc)[4]) {
    constexpr bool Q2 = TYPE == T_Q2_K;
    constexpr int QS = Q2 ? 16 : 32;
    const uint8_t *rr[2] = {r0, r1};
    const uint32_t sh_t = 8u * t;
    float cm[4] = {0.0f, 0.0f, 0.0f, 0.0f}, fs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int ip = 0; ip < 2; ip++) {
        int sc[2][8];
        uint32_t a_lo[2], a_hi[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            uint32_t w0, w1;                                   // the half's eight sub-block bytes (Q2_K: scale | min << 4; Q3_K: 6-bit scale)
            if constexpr (Q2) {
                w0 = ldw_a2<true>(rr[r] + 8 * ip);
                w1 = ldw_a2<true>(rr[r] + 8 * ip + 4);
            } else {
                const uint32_t S0 = ldw_a2<AL>(rr[r] + 96), S1 = ldw_a2<AL>(rr[r] + 100), S2 = ldw_a2<AL>(rr[r] + 104);
                w0 = ((S0 >> (4 * ip)) & 0x0f0f0f0fu) | (((S2 >> (4 * ip)) & 0x03030303u) << 4);
                w1 = ((S1 >> (4 * ip)) & 0x0f0f0f0fu) | (((S2 >> (4 * ip + 2)) & 0x03030303u) << 4);
            }
            int mb_lo, mb_hi;                                  // the tf32 factor of sub-blocks t and t + 4 (Q2_K: minimum; Q3_K: scale - 32)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int b = (int)(((k < 4 ? w0 : w1) >> (8 * (k & 3))) & 0xffu);
                sc[r][k] = Q2 ? (b & 0xf) : b - 32;
            }
            {
                const int blo = (int)((w0 >> sh_t) & 0xffu), bhi = (int)((w1 >> sh_t) & 0xffu);
                mb_lo = Q2 ? (blo >> 4) : blo - 32;
                mb_hi = Q2 ? (bhi >> 4) : bhi - 32;
            }
            a_lo[r] = __float_as_uint((float)mb_lo);           // exact tf32 operands (|value| <= 32)
            a_hi[r] = __float_as_uint((float)mb_hi);
        }
        {
            const float2 sa = sp[8 * ip + t], sb = sp[8 * ip + t + 4];
            mma_tf32(cm, a_lo[0], a_lo[1], a_hi[0], a_hi[1], __float_as_uint(sa.x), __float_as_uint(sb.x));
            mma_tf32(cm, a_lo[0], a_lo[1], a_hi[0], a_hi[1], __float_as_uint(sa.y), __float_as_uint(sb.y));
        }
        uint32_t wq[2][2], wh[2][2];                           // [row][16-value half u]: this lane's 4 bytes of qs (and of hmask)
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int u = 0; u < 2; u++) {
                wq[r][u] = ldw_a2<Q2 ? true : AL>(rr[r] + QS + 32 * ip + 16 * u + 4 * t);
                if constexpr (!Q2) wh[r][u] = ldw_a2<AL>(rr[r] + 16 * u + 4 * t) >> (4 * ip);
            }
#pragma unroll
        for (int j = 0; j < 4; j++) {                          // 32-value group 4 ip + j = sub-blocks 2 j, 2 j + 1 of this half
            uint32_t a[2][2];
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    a[r][u] = (wq[r][u] >> (2 * j)) & 0x03030303u;
                    if constexpr (!Q2) a[r][u] |= ((wh[r][u] >> j) & 0x01010101u) << 2;
                }
            const uint8_t *ab = ap + 16 * (8 * ip + 2 * j) + 4 * t;
            const uint32_t ba = *reinterpret_cast<const uint32_t *>(ab), bb = *reinterpret_cast<const uint32_t *>(ab + 16);
            int ca[4], cb[4];
            mma_u8s8_k16(ca, a[0][0], a[1][0], ba);
            mma_u8s8_k16(cb, a[0][1], a[1][1], bb);
            const float2 d8 = dp[4 * ip + j];
            fs[0] = fmaf(d8.x, __int2float_rn(ca[0] * sc[0][2 * j] + cb[0] * sc[0][2 * j + 1]), fs[0]);
            fs[1] = fmaf(d8.y, __int2float_rn(ca[1] * sc[0][2 * j] + cb[1] * sc[0][2 * j + 1]), fs[1]);
            fs[2] = fmaf(d8.x, __int2float_rn(ca[2] * sc[1][2 * j] + cb[2] * sc[1][2 * j + 1]), fs[2]);
            fs[3] = fmaf(d8.y, __int2float_rn(ca[3] * sc[1][2 * j] + cb[3] * sc[1][2 * j + 1]), fs[3]);
        }
    }
    if constexpr (Q2) {
        // cm = sum_k m_k * 32 * (d8 * sum q8)_k
        const uint32_t dm0 = ldw_a2<true>(r0 + 80), dm1 = ldw_a2<true>(r1 + 80);
        const float d0 = ggb_h2f(dm0 & 0xffffu), d1 = ggb_h2f(dm1 & 0xffffu);
        const float n0 = ggb_h2f(dm0 >> 16) * 0.03125f, n1 = ggb_h2f(dm1 >> 16) * 0.03125f;
        acc[0] += d0 * fs[0] - n0 * cm[0];
        acc[1] += d0 * fs[1] - n0 * cm[1];
        acc[2] += d1 * fs[2] - n1 * cm[2];
        acc[3] += d1 * fs[3] - n1 * cm[3];
    } else {
        // cm = sum_k (sc_k - 32) * 32 * (d8 * sum q8)_k; the offset is -4 per value: -4 / 32 = -1/8
        const float d0 = ggb_h2f(ld_u16(r0 + 108)), d1 = ggb_h2f(ld_u16(r1 + 108));
        acc[0] += d0 * (fs[0] - 0.125f * cm[0]);
        acc[1] += d0 * (fs[1] - 0.125f * cm[1]);
        acc[2] += d1 * (fs[2] - 0.125f * cm[2]);
        acc[3] += d1 * (fs[3] - 0.125f * cm[3]);
    }
}
template <> struct MmaOps<T_Q2_K> {
    static constexpr int BS = 84;
    static constexpr int SMODE = 1;
    __device__ __forceinline__ static void step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                const float2 *sp, int *, int, int t, float (&acc)[4]) {
        q23_step<T_Q2_K, true>(r0, r1, ap, dp, sp, t, acc);
    }
};
template <> struct MmaOps<T_Q3_K> {
    static constexpr int BS = 110;
    static constexpr int SMODE = 1;
    __device__ __forceinline__ static void step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                const float2 *sp, int *, int, int t, float (&acc)[4]) {
        if ((reinterpret_cast<uintptr_t>(r0) & 2) == 0) q23_step<T_Q3_K, true>(r0, r1, ap, dp, sp, t, acc);     // warp-uniform
        else q23_step<T_Q3_K, false>(r0, r1, ap, dp, sp, t, acc);
    }
};

// ---- the legacy 32-value block types (round 2): eight blocks = one 256-value unit of the kernel, like IQ4_NL.  Lane t owns
// values 4t..4t+3 (a0/a1) and 16+4t..16+4t+3 (a2/a3) of every block and the matching activation bytes.  Offsets and minima
// go through block_q8_1.s = the fp16 sum of the raw inputs, as in the reference (vecdotq.cuh:45-162): SMODE 2.
//   Q4_0 / Q5_0:  d * (d8 * sumi - off * s)  ->  sum_g (d d8) sumi  -  off * sum_g d[row, g] s[token, g]; the second sum is a
//                 16 x 8 x 8 GEMM on the tf32 MMA (d and s are fp16 values: exact tf32 operands), like the minima of Q4_K
//   Q8_0:         (d d8) sumi
//   Q4_1 / Q5_1:  sumi * h(d d8) + h(m s) with BOTH products rounded to fp16 -- one HMUL2 of the block's (d, m) pair with the
//                 token's (d8, s) pair, exactly the reference's __hmul2(dm, ds8) (vecdotq.cuh:69-91, 130-162)
template <int TYPE>
struct LegacyMma {
    static constexpr int BB = BlockTraits<TYPE>::BS;                // bytes per 32-value block
    static constexpr int BS = 8 * BB;
    static constexpr int SMODE = 2;
    static constexpr bool HAS_M = (TYPE == T_Q4_1 || TYPE == T_Q5_1);
    static constexpr bool FIVE = (TYPE == T_Q5_0 || TYPE == T_Q5_1);
    static constexpr int QH = HAS_M ? 4 : 2;                        // offset of the fifth bits (Q5_*)
    static constexpr int QS = (TYPE == T_Q4_0 || TYPE == T_Q8_0) ? 2 : TYPE == T_Q4_1 ? 4 : TYPE == T_Q5_0 ? 6 : 8;

    // word at r + OFF: the block sizes make it 4-byte aligned for every block (Q4_1, Q5_1) or for the odd blocks only
    template <int IB, int OFF>
    __device__ __forceinline__ static uint32_t word(const uint8_t *r, int t) {
        constexpr bool AL = ((IB * BB + OFF) % 4) == 0;
        return ldw_a2<AL>(r + IB * BB + OFF + 4 * t);
    }
    template <int IB>
    __device__ __forceinline__ static void frag(const uint8_t *r, int t, uint32_t &lo, uint32_t &hi) {
        if constexpr (TYPE == T_Q8_0) {
            lo = word<IB, QS>(r, t);
            hi = word<IB, QS + 16>(r, t);
        } else {
            const uint32_t w = word<IB, QS>(r, t);
            lo = w & 0x0f0f0f0fu;
            hi = (w >> 4) & 0x0f0f0f0fu;
            if constexpr (FIVE) {                                   // bit j of qh = fifth bit of value j
                constexpr bool AL = ((IB * BB + QH) % 4) == 0;
                const uint32_t qh = ldw_a2<AL>(r + IB * BB + QH);
                lo |= ((((qh >> (4 * t)) & 0xfu) * 0x00204081u) & 0x01010101u) << 4;
                hi |= ((((qh >> (16 + 4 * t)) & 0xfu) * 0x00204081u) & 0x01010101u) << 4;
            }
        }
    }
    template <int IB>
    __device__ __forceinline__ static void group(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                 const float2 *s_p0, const float2 *s_p1, const float (&s0)[8], const float (&s1)[8],
                                                 int t, float (&fs)[4]) {
        uint32_t a0, a1, a2, a3;
        frag<IB>(r0, t, a0, a2);
        frag<IB>(r1, t, a1, a3);
        const uint32_t b0 = *reinterpret_cast<const uint32_t *>(ap + 32 * IB + 4 * t);
        const uint32_t b1 = *reinterpret_cast<const uint32_t *>(ap + 32 * IB + 16 + 4 * t);
        int c[4];
        mma_s8(c, a0, a1, a2, a3, b0, b1);
        float cf[4];
#pragma unroll
        for (int i = 0; i < 4; i++) cf[i] = __int_as_float(c[i] + kMagicBits) - kMagic;       // |dot| <= 32 * 127 * 128 < 2^22: exact
        const float2 d8 = dp[IB];
        if constexpr (HAS_M) {
            const __half2 dm0 = *reinterpret_cast<const __half2 *>(r0 + IB * BB), dm1 = *reinterpret_cast<const __half2 *>(r1 + IB * BB);
            const __half2 ta = __floats2half2_rn(d8.x, s_p0[IB].x), tb = __floats2half2_rn(d8.y, s_p1[IB].x);   // (d8, s): fp16 values, exact
            const float2 p00 = __half22float2(__hmul2(dm0, ta)), p01 = __half22float2(__hmul2(dm0, tb));
            const float2 p10 = __half22float2(__hmul2(dm1, ta)), p11 = __half22float2(__hmul2(dm1, tb));
            fs[0] += fmaf(cf[0], p00.x, p00.y);
            fs[1] += fmaf(cf[1], p01.x, p01.y);
            fs[2] += fmaf(cf[2], p10.x, p10.y);
            fs[3] += fmaf(cf[3], p11.x, p11.y);
        } else {
            constexpr int si = 2 * (IB & 3) + (IB >> 2);
            fs[0] = fmaf(cf[0], d8.x * s0[si], fs[0]);
            fs[1] = fmaf(cf[1], d8.y * s0[si], fs[1]);
            fs[2] = fmaf(cf[2], d8.x * s1[si], fs[2]);
            fs[3] = fmaf(cf[3], d8.y * s1[si], fs[3]);
        }
    }
    __device__ __forceinline__ static void step(const uint8_t *r0, const uint8_t *r1, const uint8_t *ap, const float2 *dp,
                                                const float2 *sp, const float2 *s_p0, const float2 *s_p1, int *scr, int q, int t,
                                                float (&acc)[4]) {
        float s0[8], s1[8];
        float cm[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if constexpr (!HAS_M) {
            // lane t fetches d of blocks t, t + 4 of rows q, q + 8: the warp's scratch gives every lane all eight per row, and the
            // four values are the tf32 A fragment of the offset GEMM
            const float d0a = ggb_h2f(ld_u16(r0 + BB * t)), d0b = ggb_h2f(ld_u16(r0 + BB * (t + 4)));
            const float d1a = ggb_h2f(ld_u16(r1 + BB * t)), d1b = ggb_h2f(ld_u16(r1 + BB * (t + 4)));
            float *fscr = reinterpret_cast<float *>(scr);
            *reinterpret_cast<float2 *>(fscr + q * 8 + 2 * t) = make_float2(d0a, d0b);
            *reinterpret_cast<float2 *>(fscr + (q + 8) * 8 + 2 * t) = make_float2(d1a, d1b);
            if constexpr (TYPE != T_Q8_0)
                mma_tf32(cm, __float_as_uint(d0a), __float_as_uint(d1a), __float_as_uint(d0b), __float_as_uint(d1b),
                         __float_as_uint(sp[t].x), __float_as_uint(sp[t + 4].x));
            __syncwarp();
            const float4 a = *reinterpret_cast<const float4 *>(fscr + q * 8), b = *reinterpret_cast<const float4 *>(fscr + q * 8 + 4);
            const float4 c = *reinterpret_cast<const float4 *>(fscr + (q + 8) * 8), d = *reinterpret_cast<const float4 *>(fscr + (q + 8) * 8 + 4);
            s0[0] = a.x; s0[1] = a.y; s0[2] = a.z; s0[3] = a.w; s0[4] = b.x; s0[5] = b.y; s0[6] = b.z; s0[7] = b.w;
            s1[0] = c.x; s1[1] = c.y; s1[2] = c.z; s1[3] = c.w; s1[4] = d.x; s1[5] = d.y; s1[6] = d.z; s1[7] = d.w;
        }
        float fs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        group<0>(r0, r1, ap, dp, s_p0, s_p1, s0, s1, t, fs);
        group<1>(r0, r1, ap, dp, s_p0, s_p1, s0, s1, t, fs);
        group<2>(r0, r1, ap, dp, s_p0, s_p1, s0, s1, t, fs);
        group<3>(r0, r1, ap, dp, s_p0, s_p1, s0, s1, t, fs);
        group<4>(r0, r1, ap, dp, s_p0, s_p1, s0, s1, t, fs);
        group<5>(r0, r1, ap, dp, s_p0, s_p1, s0, s1, t, fs);
        group<6>(r0, r1, ap, dp, s_p0, s_p1, s0, s1, t, fs);
        group<7>(r0, r1, ap, dp, s_p0, s_p1, s0, s1, t, fs);
        constexpr float off = (TYPE == T_Q4_0) ? 8.0f : (TYPE == T_Q5_0) ? 16.0f : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] += fs[i] - off * cm[i];
    }
};
template <> struct MmaOps<T_Q4_0> : LegacyMma<T_Q4_0> {};
template <> struct MmaOps<T_Q4_1> : LegacyMma<T_Q4_1> {};
template <> struct MmaOps<T_Q5_0> : LegacyMma<T_Q5_0> {};
template <> struct MmaOps<T_Q5_1> : LegacyMma<T_Q5_1> {};
template <> struct MmaOps<T_Q8_0> : LegacyMma<T_Q8_0> {};

// ------------------------------------------------------------------------------------------------
// launch 2: the GEMV
// ------------------------------------------------------------------------------------------------
// PHASED = false: the whole activation image is resident (P == 1 at compile time: the common case keeps its code)
template <int TYPE, typename T, int NW, bool PHASED>
__global__ void __launch_bounds__(NW * 32, 1)
mmvq_mma_kernel(const MmaParams p, const __grid_constant__ CUtensorMap tmapW) {
    using Ops = MmaOps<TYPE>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int P = PHASED ? p.phases : 1, pc = PHASED ? p.nchunks / P : p.nchunks;   // chunks per phase (P divides nchunks)
    const ActLayout Lg = act_layout(p.B, p.K, (Ops::SMODE == 1));                        // the image in the workspace
    const ActLayout L = act_layout(p.B, p.K / P, (Ops::SMODE == 1));                     // one phase of it in shared memory
    int *scr_all = reinterpret_cast<int *>(smem + p.off_scr);                   // [NW][2][kScrWords]
    float4 *accy = reinterpret_cast<float4 *>(smem + p.off_accy);               // [team][row group][32] (P > 1)
    float4 *part = reinterpret_cast<float4 *>(smem + p.off_part);               // [2][NW][32]
    uint8_t *slots = smem + p.off_slots;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.off_bars);
    uint64_t *act_full = full + kMaxSlots;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = lane >> 2, t = lane & 3;
    const int nsb = p.nsb;

    // this CTA's contiguous range of row groups
    const int64_t ngr_total = (p.N + kRowsPerGroup - 1) / kRowsPerGroup;
    const int64_t g_begin = ngr_total * (int64_t)blockIdx.x / gridDim.x;
    const int64_t g_end = ngr_total * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int64_t r_begin = g_begin * kRowsPerGroup;
    const int64_t r_end = min(g_end * kRowsPerGroup, p.N);
    const int nrows = (int)(r_end - r_begin);
    const int nrg = (int)(g_end - g_begin);

    const int nslots = NW * p.depth;
    const bool inl = !PHASED && p.inline_q != 0;                   // (host: only with P == 1 and an even grid of cluster pairs)
    if (tid < nslots) mbar_init(&full[tid], 1);
    if (tid == 0) {
        mbar_init(act_full, 1);
        if (inl) {
            fence_mbar_init();
            // every byte both CTAs will store into THIS CTA's image: quants of the B tokens, d8 of the padded token count, s table
            const int ng = p.K >> 5, ntok = ((p.B + 1) >> 1) << 1;
            mbar_expect_tx(act_full, (uint32_t)(p.B * p.K + ntok * ng * 4 + p.B * ng * (Ops::SMODE == 1 ? 16 : 8)));
        }
    }
    fence_mbar_init();
    __syncthreads();
    if (inl) cluster_arrive();         // (nothing in flight yet for the release to wait for; the wait sits before the first remote store)
    pdl_launch_dependents();
    if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 0] = gtime();

    const int S = 1 << p.s_log2, NT = NW >> p.s_log2;
    const int team = warp >> p.s_log2, s = warp & (S - 1);
    const int my_rgs = (nrg - team + NT - 1) / NT;                 // may be <= 0
    const int my_chunks = (pc - s + S - 1) >> p.s_log2;             // per phase; >= 1 (S <= pc)
    const int my_units = my_rgs > 0 ? my_rgs * my_chunks * P : 0;
    const int ring_base = warp * p.depth;
    const uint64_t pol = l2_policy_evict_first();

    // issue state, warp-uniform: the next unit (phase, row group, chunk) of this warp to stream
    int is_k = 0, is_ring = ring_base, is_rg = team, is_c = s, is_ph = 0;
    auto issue_next = [&]() {                                      // whole warp (lane 0 issues)
        if (lane == 0) {
            // rows past N and bytes past the row end are zero-filled by the TMA unit and count as transferred
            // (blocks that are not multiples of 16 bytes: the box starts at the 16-byte boundary below the chunk)
            mbar_expect_tx(&full[is_ring], (uint32_t)(kRowsPerGroup * p.slot_row_stride));
            tma_load_2d_hint(slots + is_ring * p.slot_stride, &tmapW, (((is_ph * pc + is_c) * p.cs * Ops::BS) & ~15) >> 2,
                             (int)r_begin + is_rg * kRowsPerGroup, &full[is_ring], pol);
        }
        is_c += S;
        if (is_c >= pc) {
            is_c = s;
            is_rg += NT;
            if (is_rg >= nrg) { is_rg = team; is_ph++; }
        }
        is_k++;
        if (++is_ring == ring_base + p.depth) is_ring = ring_base;
    };

    if (p.pdl_prefetch)
        while (is_k < min(p.depth, my_units)) issue_next();
    pdl_wait();                                                    // the activation image is complete
    if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 1] = gtime();
    // phase `ph` of the activation image: the whole image in one piece when P == 1, else the K-slice of every token's
    // quants and of both scale tables
    auto load_image = [&](int ph) {                                // one thread
        if (P == 1) {
            mbar_expect_tx(act_full, (uint32_t)L.total);
            constexpr int kPiece = 32768;
            for (int off = 0; off < L.total; off += kPiece)
                bulk_g2s(smem + off, p.act_img + off, (uint32_t)min(kPiece, L.total - off), act_full);
            return;
        }
        const int Kp = p.K / P, k0 = ph * Kp, g0 = k0 >> 5, ngp = Kp >> 5, npairs = (p.B + 1) >> 1;
        const int sm = (Ops::SMODE == 1) ? 2 : 1;
        mbar_expect_tx(act_full, (uint32_t)(p.B * Kp + npairs * ngp * 8 + p.B * ngp * sm * 8));
        for (int b = 0; b < p.B; b++)
            bulk_g2s(smem + b * L.act_stride, p.act_img + b * Lg.act_stride + k0, (uint32_t)Kp, act_full);
        for (int pr = 0; pr < npairs; pr++)
            bulk_g2s(smem + L.off_d + pr * L.d_stride2 * 8, p.act_img + Lg.off_d + (pr * Lg.d_stride2 + g0) * 8,
                     (uint32_t)(ngp * 8), act_full);
        for (int b = 0; b < p.B; b++)
            bulk_g2s(smem + L.off_s + b * L.s_stride2 * 8, p.act_img + Lg.off_s + (b * Lg.s_stride2 + g0 * sm) * 8,
                     (uint32_t)(ngp * sm * 8), act_full);
    };
    if (inl) {
        // ---- q8_1-quantize this CTA's half of the (token, group) pairs (bit-identical to quantize_q8_1 / the image kernel) and
        // store every group into both CTAs of the pair.  Four lanes per group, two groups per thread and pass.
        const T *X = static_cast<const T *>(p.X);
        const int ng = p.K >> 5, ntok = ((p.B + 1) >> 1) << 1;
        const int half = (ntok * ng) >> 1;                          // (ng is a multiple of 8: whole warps stay together)
        const uint32_t crank = cluster_ctarank();
        const int g_begin = (int)crank * half, g_end = g_begin + half;
        const int t4 = lane & 3;
        constexpr int GPP = NW * 8;                                 // groups per pass
        const uint32_t img_s = smem_u32(smem), bar_s = smem_u32(act_full);
        const uint32_t img0 = mapa(img_s, 0), img1 = mapa(img_s, 1), bar0 = mapa(bar_s, 0), bar1 = mapa(bar_s, 1);
        for (int base = g_begin; base < g_end; base += 2 * GPP) {
            float v[2][8];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int gi = base + u * GPP + (tid >> 2);
                const int b = gi / ng, g = gi - b * ng;
                if (gi < g_end && b < p.B) {
                    load_act8<T>(X, static_cast<const T *>(p.xin.x2), p.xin.ldx, b, g * 32 + t4 * 8, p.x_vec != 0, v[u]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[u][i] = 0.0f;
                }
            }
            if (base == g_begin) {
                if (!p.pdl_prefetch)
                    while (is_k < min(p.depth, my_units)) issue_next();
                cluster_wait();                                     // the peer runs: its shared memory and barrier may be written
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (base + u * GPP >= g_end) break;                 // block-uniform
                const int gi = base + u * GPP + (tid >> 2);
                if (gi >= g_end) continue;                          // warp-uniform (8 groups per warp, half % 8 == 0)
                const int b = gi / ng, g = gi - b * ng;
                const Q8Group qg = q8_1_quantize_group<Ops::SMODE == 2>(v[u]);
                int sum = __dp4a((int)qg.q_lo, 0x01010101, __dp4a((int)qg.q_hi, 0x01010101, 0));
                sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                const int sum16 = sum;
                sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                const float d8 = __half2float(__float2half_rn(qg.d));
                const uint32_t qoff = (uint32_t)(b * L.act_stride + g * 32 + t4 * 8);
                const uint32_t doff = (uint32_t)(L.off_d + (((b >> 1) * L.d_stride2 + g) * 2 + (b & 1)) * 4);
                uint32_t soff = 0, sx = 0, sy = 0;
                bool s_on = false;
                if (b < p.B) {
                    if (Ops::SMODE == 1) {
                        s_on = (t4 & 1) == 0;
                        const float s8 = d8 * (float)(32 * sum16);
                        const float hi = __uint_as_float(__float_as_uint(s8) & 0xffffe000u);
                        sx = __float_as_uint(hi); sy = __float_as_uint(s8 - hi);
                        soff = (uint32_t)(L.off_s + (b * L.s_stride2 + 2 * g + (t4 >> 1)) * 8);
                    } else {
                        s_on = t4 == 0;
                        soff = (uint32_t)(L.off_s + (b * L.s_stride2 + g) * 8);
                        if (Ops::SMODE == 2) {
                            sx = __float_as_uint(__half2float(__float2half_rn(qg.sum)));
                        } else {
                            const float s8 = d8 * (float)sum;
                            const float hi = __uint_as_float(__float_as_uint(s8) & 0xffffe000u);
                            sx = __float_as_uint(hi); sy = __float_as_uint(s8 - hi);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const uint32_t ib = r ? img1 : img0, mb = r ? bar1 : bar0;
                    if (b < p.B) st_async_v2(ib + qoff, qg.q_lo, qg.q_hi, mb);
                    if (t4 == 0) st_async_b32(ib + doff, __float_as_uint(d8), mb);
                    if (s_on) st_async_v2(ib + soff, sx, sy, mb);
                }
            }
        }
    } else {
        if (warp == NW - 1 && lane == 0) load_image(0);            // (the last warp has the fewest units to prime)
        if (!p.pdl_prefetch)
            while (is_k < min(p.depth, my_units)) issue_next();
    }
    mbar_wait(act_full, 0);
    if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 2] = gtime();

    // ===================== row groups of 16, K split over the team =====================
    const int qc = min(q, p.B - 1);                                // lanes of absent tokens shadow the last one
    const int tc = min(t, (p.B - 1) >> 1);                         // ... and absent token pairs the last pair
    const uint8_t *act_tok = smem + qc * L.act_stride;
    const float2 *d_pair = reinterpret_cast<const float2 *>(smem + L.off_d) + tc * L.d_stride2;
    const float2 *s_tok = reinterpret_cast<const float2 *>(smem + L.off_s) + qc * L.s_stride2;
    // (SMODE 2) the s tables of the two tokens whose outputs this lane holds
    const float2 *s_pair0 = reinterpret_cast<const float2 *>(smem + L.off_s) + min(2 * tc, p.B - 1) * L.s_stride2;
    const float2 *s_pair1 = reinterpret_cast<const float2 *>(smem + L.off_s) + min(2 * tc + 1, p.B - 1) * L.s_stride2;
    int *scr = scr_all + warp * (2 * kScrWords);
    int ring = ring_base;
    uint32_t phase = 0;
    int step_no = 0;
    for (int ph = 0; ph < P; ph++) {
    if (ph > 0) {
        // every warp is done with the previous K-slice of the image: fetch the next one over it
        __syncthreads();
        if (warp == NW - 1 && lane == 0) {
            fence_proxy_async();
            load_image(ph);
        }
        mbar_wait(act_full, ph & 1);
    }
    for (int rgi = 0; rgi < my_rgs; rgi++) {
        const int rg = team + rgi * NT;
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int cl = s; cl < pc; cl += S) {
            const int sbl0 = cl * p.cs;                            // first super-block of the chunk inside this phase
            const int sb0 = ph * pc * p.cs + sbl0;                 // ... and inside the row
            const int nsbc = min(p.cs, nsb - sb0);
            mbar_wait(&full[ring], phase);
            const uint8_t *r0 = slots + ring * p.slot_stride + q * p.slot_row_stride + ((sb0 * Ops::BS) & 15);
            const uint8_t *r1 = r0 + 8 * p.slot_row_stride;
#pragma unroll 1
            for (int i = 0; i < nsbc; i++) {
                const int sb = sbl0 + i;
                if constexpr (Ops::SMODE == 2)
                    Ops::step(r0 + i * Ops::BS, r1 + i * Ops::BS, act_tok + sb * 256, d_pair + sb * 8, s_tok + sb * 8, s_pair0 + sb * 8,
                              s_pair1 + sb * 8, scr + (step_no & 1) * kScrWords, q, t, acc);
                else
                    Ops::step(r0 + i * Ops::BS, r1 + i * Ops::BS, act_tok + sb * 256, d_pair + sb * 8, s_tok + sb * ((Ops::SMODE == 1) ? 16 : 8),
                              scr + (step_no & 1) * kScrWords, q, t, acc);
                step_no++;
            }
            __syncwarp();
            if (is_k < my_units) {                                 // recycle the slot: every lane has finished reading it
                if (lane == 0) fence_proxy_async();
                issue_next();
            }
            if (++ring == ring_base + p.depth) { ring = ring_base; phase ^= 1u; }
        }
        // partial sums of the team's warps meet in shared memory and are added in warp order
        if (S > 1) {
            float4 *pb = part + ((ph * my_rgs + rgi) & 1) * (NW * 32);
            pb[warp * 32 + lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            named_bar_sync(1 + team, S * 32);
            if (s == 0) {
                const float4 *src = pb + (team << p.s_log2) * 32 + lane;
                float4 v = src[0];
                for (int j = 1; j < S; j++) {
                    const float4 w = src[j * 32];
                    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                }
                acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w;
            }
        }
        if (s == 0 && P > 1) {
            // the row group's sum over the phases so far lives in shared memory (only this warp touches its entry)
            float4 *a = accy + (team * p.accy_rgs + rgi) * 32 + lane;
            if (ph > 0) {
                const float4 o = *a;
                acc[0] += o.x; acc[1] += o.y; acc[2] += o.z; acc[3] += o.w;
            }
            if (ph + 1 < P) *a = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        if (s == 0 && ph + 1 == P) {
            const int lr = rg * kRowsPerGroup + q;
            const int tok = 2 * t;
            if (lr < nrows) {
                if (tok < p.B) ggb_yout_store<T>(p.yo, tok, r_begin + lr, acc[0]);
                if (tok + 1 < p.B) ggb_yout_store<T>(p.yo, tok + 1, r_begin + lr, acc[1]);
            }
            if (lr + 8 < nrows) {
                if (tok < p.B) ggb_yout_store<T>(p.yo, tok, r_begin + lr + 8, acc[2]);
                if (tok + 1 < p.B) ggb_yout_store<T>(p.yo, tok + 1, r_begin + lr + 8, acc[3]);
            }
        }
    }
    }
    if (p.trace && tid == 0) p.trace[blockIdx.x * 4 + 3] = gtime();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int g_cs = 0, g_s_log2 = -1;          // tuning overrides (0 / -1 = automatic)
int g_min_batch = 3;                  // batches below this stay on the dp4a kernels
int g_inline_max_groups = 256;        // in-kernel pair quantization when a CTA's half of the (token, group) pairs is at most this (0 = never).
                                      // Measured (profiles/r02_mmvq_mma_inline.log, Q4_K): 14336 x 4096 batch 3 / 4 12.7 -> 12.1 / 12.2 us, but batch 8
                                      // 13.1 -> 14.0 and every longer row loses (4096 x 14336 batch 8: 13.8 -> 34 us): the pass is a chain of
                                      // L2 loads, shuffles and remote stores per group, and 512 threads get through 256 groups at best

}  // namespace

int ggb_mmvq_static_weights();        // mmvq_sb.cu
unsigned long long *ggb_mmvq_trace_slot();

namespace {

// fills p and the shared-memory size for NW warps and `cs` super-blocks per chunk; false if the ring would be too shallow
template <int TYPE>
bool plan(MmaParams &p, int &smem, int NW, int cs, int P, int B, int K, int64_t N, int grid, int max_smem) {
    using Ops = MmaOps<TYPE>;
    p.cs = cs < p.nsb ? cs : p.nsb;
    // TMA box: the chunk's bytes, from the 16-byte boundary at or below its start (Q6_K: 210-byte blocks)
    const int rs = (Ops::BS % 16 == 0) ? p.cs * Ops::BS : ((p.cs * Ops::BS + 14 + 15) & ~15);
    if (rs > 1024) return false;                               // TMA box <= 256 uint32 elements
    p.nchunks = (p.nsb + p.cs - 1) / p.cs;
    if (P > 1 && (p.nsb % p.cs != 0 || p.nchunks % P != 0 || (K / P) % 512 != 0)) return false;
    p.phases = P;
    const ActLayout L = act_layout(B, K / P, (Ops::SMODE == 1));
    const int pc = p.nchunks / P;
    int sl = 0;
    // S = warps per team, a power of two <= min(chunks per phase, NW), chosen from K alone (the summation order must not
    // depend on N or the grid).  Measured, Q4_K batch 4 / 8 (profiles/r02_mmvq_mma_teams.log): teams of 2 win on K <= 8192
    // (14336 x 4096: 14.5 / 14.8 -> 12.7 / 13.4 us, 28672 x 8192: 36.0 / 40.8 -> 31.3 / 35.5) -- fewer partial-sum barriers,
    // more row groups in flight --, teams of 4 on longer rows (4096 x 14336: 14.4 -> 13.7 at batch 8, 8192 x 28672: 47.2 -> 44.2)
    const int sl_cap = K <= 8192 ? 1 : 2;
    while (sl < sl_cap && (2 << sl) <= pc && (2 << sl) <= NW) sl++;
    if (g_s_log2 >= 0 && (1 << g_s_log2) <= pc && (1 << g_s_log2) <= NW) sl = g_s_log2;
    p.s_log2 = sl;
    p.slot_row_stride = rs;
    p.slot_stride = (kRowsPerGroup * rs + 127) & ~127;
    p.off_scr = (L.total + 127) & ~127;
    p.off_part = p.off_scr + NW * 2 * kScrWords * 4;
    p.off_accy = p.off_part + 2 * NW * 32 * 16;
    const int NT = NW >> sl;
    const int64_t ngr = (N + kRowsPerGroup - 1) / kRowsPerGroup;
    const int nrg_max = (int)((ngr + grid - 1) / grid) + 1;
    p.accy_rgs = (nrg_max + NT - 1) / NT;
    // the carried sums of a phased walk get a FIXED reserve (32 row groups per CTA + rounding: up to 75 776 rows on 148 SMs), so
    // which (P, warps, chunk, team) plan fits -- and with it the fp32 summation order -- depends on (type, batch, K) only, never
    // on N or the grid: an N-shard and the full matrix take the same plan (round-1 advisor finding)
    constexpr int kAccyReserveRgs = 32;
    if (P > 1 && NT * p.accy_rgs > kAccyReserveRgs + NT) return false;
    p.off_slots = (p.off_accy + (P > 1 ? (kAccyReserveRgs + NT) * 32 * 16 : 0) + 127) & ~127;
    const int bar_bytes = (kMaxSlots + 1) * 8;
    const int avail = max_smem - p.off_slots - bar_bytes - 128;
    if (avail <= 0) return false;
    int depth = avail / (p.slot_stride * NW);
    if (depth < 2) return false;
    if (depth > kMaxSlots / NW) depth = kMaxSlots / NW;
    if (depth > 6) depth = 6;
    p.depth = depth;
    p.off_bars = (p.off_slots + NW * depth * p.slot_stride + 127) & ~127;
    smem = p.off_bars + bar_bytes;
    return true;
}

template <int TYPE, typename T, int NW, bool PHASED>
int launch_gemv(MmaParams &p, int smem, const CUtensorMap &tmap, cudaStream_t stream) {
    const DeviceInfo &di = ggb_device_info();
    static bool attr_set[64] = {};
    if (!attr_set[di.dev]) {
        GGB_CUDA_TRY(cudaFuncSetAttribute(mmvq_mma_kernel<TYPE, T, NW, PHASED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          di.max_smem_optin));
        attr_set[di.dev] = true;
    }
    const int64_t ngr = (p.N + kRowsPerGroup - 1) / kRowsPerGroup;
    const int grid = (int)(ngr < di.sm_count ? ngr : di.sm_count);
    p.trace = ggb_mmvq_trace_slot();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.inline_q ? 2 : 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmvq_mma_kernel<TYPE, T, NW, PHASED>>(di.dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmvq_mma_kernel<TYPE, T, NW, PHASED>, p, tmap));
    return GGB_OK;
}

template <int TYPE, typename T>
int launch(const void *W, const void *X, void *Y, int64_t B, int64_t K, int64_t N, void *ws, cudaStream_t stream) {
    using Ops = MmaOps<TYPE>;
    const DeviceInfo &di = ggb_device_info();
    MmaParams p;
    p.act_img = static_cast<const uint8_t *>(ws);
    p.yo = ggb_make_yout(Y, N);
    p.N = N;
    p.K = (int)K;
    p.B = (int)B;
    p.nsb = (int)(K / 256);
    p.row_bytes = p.nsb * Ops::BS;
    p.pdl_prefetch = ggb_mmvq_static_weights();
    if (p.row_bytes % 16 != 0 || N > 0x7fffffff) return 1;
    // preferred chunk: 2 super-blocks for short rows (more K-parallelism), 4 for long ones; smaller if the ring needs it.
    // If the whole activation image does not fit beside a useful ring (long rows x large batch), the row is walked in
    // K phases with one slice of the image resident at a time.
    const int pref = g_cs > 0 ? g_cs : (p.nsb <= 16 ? 2 : 4);
    const int64_t ngr_all = (N + kRowsPerGroup - 1) / kRowsPerGroup;
    const int grid_est = (int)(ngr_all < di.sm_count ? ngr_all : di.sm_count);
    int smem = 0, nw = 0;
    for (int P : {1, 2, 4, 7, 8, 14, 16}) {
        for (int NW : {16, 8}) {
            for (int cs = pref; cs >= 1 && !nw; cs >>= 1)
                if (plan<TYPE>(p, smem, NW, cs, P, (int)B, (int)K, N, grid_est, di.max_smem_optin)) nw = NW;
            if (nw) break;
        }
        if (nw) break;
    }
    if (!nw) return 1;
    alignas(64) CUtensorMap tmap;
    {
        const int rc = ggb_make_row_tensor_map(&tmap, W, N, p.row_bytes, kRowsPerGroup, p.slot_row_stride, "ggb_mmvq");
        if (rc != GGB_OK) return rc;
    }
    const GgbXIn xin0 = ggb_make_xin(K);
    p.X = X;
    p.xin = xin0;
    p.x_vec = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(xin0.x2) & 15) == 0 &&
              (xin0.ldx * (int64_t)sizeof(T)) % 16 == 0;
    // Short activation tiles: no image launch.  The CTA pairs of the GEMV quantize half of the tile each and exchange the halves
    // through distributed shared memory -- one dependency resolution and one trip through L2 less (measured per-CTA timeline,
    // profiles/r02_mmvq_mma_trace.log: 4.4 us from the end of the previous GEMV to this one's dependency through the image launch).
    const int64_t half_groups = ((B + 1) / 2) * 2 * (K / 32) / 2;
    p.inline_q = (p.phases == 1 && grid_est >= 2 && grid_est % 2 == 0 && g_inline_max_groups > 0 && half_groups <= g_inline_max_groups) ? 1 : 0;
    // launch 1: the activation image
    if (!p.inline_q) {
        const int ntok = (int)((B + 1) / 2) * 2;
        const int ngroups = ntok * (int)(K / 32);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((ngroups + 63) / 64);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const GgbXIn xin = ggb_make_xin(K);
        const int x_vec = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(xin.x2) & 15) == 0 &&
                          (xin.ldx * (int64_t)sizeof(T)) % 16 == 0;
        GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmvq_act_image_kernel<T, Ops::SMODE>>(di.dev)));
        GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmvq_act_image_kernel<T, Ops::SMODE>, static_cast<const T *>(X), xin, static_cast<uint8_t *>(ws),
                                        (int)B, (int)K, x_vec));
    }
    if (p.phases == 1)
        return nw == 16 ? launch_gemv<TYPE, T, 16, false>(p, smem, tmap, stream) : launch_gemv<TYPE, T, 8, false>(p, smem, tmap, stream);
    return nw == 16 ? launch_gemv<TYPE, T, 16, true>(p, smem, tmap, stream) : launch_gemv<TYPE, T, 8, true>(p, smem, tmap, stream);
}

template <int TYPE>
int dispatch_dtype(const void *W, const void *X, void *Y, int dt, int64_t B, int64_t K, int64_t N, void *ws, cudaStream_t stream) {
    switch (dt) {
        case GGB_F32:  return launch<TYPE, float>(W, X, Y, B, K, N, ws, stream);
        case GGB_F16:  return launch<TYPE, __half>(W, X, Y, B, K, N, ws, stream);
        case GGB_BF16: return launch<TYPE, __nv_bfloat16>(W, X, Y, B, K, N, ws, stream);
        default: return 1;
    }
}

bool mma_applies(int type, int64_t B, int64_t K) {
    // measured crossover with the dp4a kernel (Llama-8B shapes): batch 3 for K = 4096, batch 2 for the long rows, where
    // the dp4a kernel's per-CTA activation quantization (3.3 us at K = 14336) costs more than the image kernel's launch
    const int64_t min_b = (g_min_batch == 3 && K > 8192) ? 2 : g_min_batch;
    const bool legacy = type == T_Q4_0 || type == T_Q4_1 || type == T_Q5_0 || type == T_Q5_1 || type == T_Q8_0;
    return (type == T_Q4_K || type == T_Q5_K || type == T_Q6_K || type == T_Q2_K || type == T_Q3_K || type == T_IQ4_XS || type == T_IQ4_NL || legacy) &&
           B >= min_b && B <= 8 &&
           K % 256 == 0 && K <= (1 << 16);
}

}  // namespace

void ggb_mmvq_mma_tune(int cs, int s_log2_plus1, int min_batch) {
    g_cs = cs;
    g_s_log2 = s_log2_plus1 - 1;
    if ((min_batch & 0x7f) > 0) g_min_batch = min_batch & 0x7f;
    g_inline_max_groups = (min_batch & 0x80) ? 0 : 256;           // bit 7: always the image launch
}

// workspace for the activation image (0 when this path does not apply)
int64_t ggb_mmvq_mma_workspace_bytes(int type, int64_t B, int64_t K) {
    if (!mma_applies(type, B, K)) return 0;
    return act_layout((int)B, (int)K, type == T_Q6_K || type == T_Q2_K || type == T_Q3_K).total;
}

// returns GGB_OK / a negative status, or 1 when this path does not apply (the caller tries mmvq_sb.cu, then mmvq.cu)
int ggb_mmvq_mma_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                      void *ws, int64_t ws_bytes, cudaStream_t stream) {
    if (!mma_applies(type, B, K) || (reinterpret_cast<uintptr_t>(W) & 15)) return 1;
    if (!ws || (reinterpret_cast<uintptr_t>(ws) & 15) || ws_bytes < act_layout((int)B, (int)K, type == T_Q6_K || type == T_Q2_K || type == T_Q3_K).total) return 1;
    int rc;
    switch (type) {
        case T_Q4_K: rc = dispatch_dtype<T_Q4_K>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q5_K: rc = dispatch_dtype<T_Q5_K>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q6_K: rc = dispatch_dtype<T_Q6_K>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_IQ4_XS: rc = dispatch_dtype<T_IQ4_XS>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_IQ4_NL: rc = dispatch_dtype<T_IQ4_NL>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q2_K: rc = dispatch_dtype<T_Q2_K>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q3_K: rc = dispatch_dtype<T_Q3_K>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q4_0: rc = dispatch_dtype<T_Q4_0>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q4_1: rc = dispatch_dtype<T_Q4_1>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q5_0: rc = dispatch_dtype<T_Q5_0>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q5_1: rc = dispatch_dtype<T_Q5_1>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        case T_Q8_0: rc = dispatch_dtype<T_Q8_0>(W, X, Y, xy_dtype, B, K, N, ws, stream); break;
        default: return 1;
    }
    if (rc != 1 || B < 6 || ggb_current_yout() || ggb_current_xin()) return rc;
    // the B x K activation image does not fit in shared memory next to a useful ring (very long rows): two passes
    // over the weights with half the batch each still beat the dp4a kernels, which would split the batch as well.
    // The second pass's image kernel waits for the first pass's GEMV (stream order), so both can share `ws`.
    const int64_t b1 = B / 2;
    const size_t esz = xy_dtype == GGB_F32 ? 4 : 2;
    rc = ggb_mmvq_mma_impl(W, X, Y, type, xy_dtype, b1, K, N, ws, ws_bytes, stream);
    if (rc != GGB_OK) return rc;
    return ggb_mmvq_mma_impl(W, static_cast<const uint8_t *>(X) + b1 * K * esz, static_cast<uint8_t *>(Y) + b1 * N * esz, type,
                             xy_dtype, B - b1, K, N, ws, ws_bytes, stream);
}
