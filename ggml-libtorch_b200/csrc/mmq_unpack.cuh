// mmq_unpack.cuh -- half-precision weight-tile unpackers for the MMQ A operand (headline K-quants).
//
// The tensor cores consume fp16, so for Q4_K / Q5_K / Q6_K the A tile is produced directly in half2
// arithmetic instead of fp32 + convert (3x fewer instructions per weight):
//     Q4_K, Q5_K :  a = fma_h( q, fp16(d*sc), -fp16(dmin*m) )      one fp16 FMA, single rounding
//     Q6_K       :  a = (q - 32) * fp16(d*sc)                      one fp16 multiply
// q is the exact integer quant (0..15, 0..31, 0..63) built with PRMT as the fp16 number 1024 + q and an
// exact HSUB2.  The (d*sc, dmin*m) pairs are computed in fp32 exactly as the dequantizer does and rounded
// to fp16 once per sub-block (a per-stage table in shared memory).  oracle/ggml_oracle.c::orc_mmq_a_operand
// restates this operand definition for the operand-exact parity test (tests/test_gpu_mmq.py).
// Every other type goes through the fp32 Unpack8 path and one cvt.rn (bit-identical to ggml_dequantize fp16).
#pragma once

#include "unpack.cuh"

__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }

// bytes (e0,e1,e2,e3) of t -> half2 {1024+e0, 1024+e1} and {1024+e2, 1024+e3}
__device__ __forceinline__ __half2 magic_h2_lo(uint32_t t) { return u32_as_h2(__byte_perm(t, 0x00006464u, 0x5140)); }
__device__ __forceinline__ __half2 magic_h2_hi(uint32_t t) { return u32_as_h2(__byte_perm(t, 0x00006464u, 0x5342)); }

// 8 bytes from a 2-byte-aligned shared-memory address (the misalignment is warp-uniform in the MMQ tile)
__device__ __forceinline__ uint2 ld_u64_a2(const uint8_t *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    if (a & 2) {
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
        return make_uint2(__funnelshift_r(w0, w1, 16), __funnelshift_r(w1, w2, 16));
    }
    return make_uint2(q[0], q[1]);
}

// FAST32: the type also provides run32(), which turns one 128-bit load (32 values of one 64-column k-step:
// 16 consecutive columns + the 16 columns 32 further on) into four 16-byte chunks of the A tile.
template <int TYPE> struct HalfUnpack { static constexpr bool HAS = false, FAST32 = false; static constexpr int NSUB = 1; };

template <> struct HalfUnpack<T_Q4_K> {
    static constexpr bool HAS = true, FAST32 = true;
    static constexpr int NSUB = 8;
    // k-step q (64 columns = sub-blocks 2q, 2q+1) of a row, bytes 16*half .. 16*half+15 of its 32 nibble bytes:
    // low nibbles -> columns 16*half + 0..15 (chunks o[0], o[1]), high nibbles -> columns 32 + 16*half + 0..15 (o[2], o[3])
    __device__ __forceinline__ static void run32(const uint8_t *b, int q, int half, uint2 sc, uint4 (&o)[4]) {
        const uint4 w = *reinterpret_cast<const uint4 *>(b + 16 + 32 * q + 16 * half);
        const __half2 dl0 = u32_as_h2(__byte_perm(sc.x, sc.x, 0x1010)), nml0 = u32_as_h2(__byte_perm(sc.x, sc.x, 0x3232));
        const __half2 dl1 = u32_as_h2(__byte_perm(sc.y, sc.y, 0x1010)), nml1 = u32_as_h2(__byte_perm(sc.y, sc.y, 0x3232));
        const __half2 k1024 = u32_as_h2(0x64006400u);
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t t0 = ww[i] & 0x0f0f0f0fu, t1 = (ww[i] >> 4) & 0x0f0f0f0fu;
            lo[2 * i] = h2_as_u32(__hfma2(__hsub2(magic_h2_lo(t0), k1024), dl0, nml0));
            lo[2 * i + 1] = h2_as_u32(__hfma2(__hsub2(magic_h2_hi(t0), k1024), dl0, nml0));
            hi[2 * i] = h2_as_u32(__hfma2(__hsub2(magic_h2_lo(t1), k1024), dl1, nml1));
            hi[2 * i + 1] = h2_as_u32(__hfma2(__hsub2(magic_h2_hi(t1), k1024), dl1, nml1));
        }
        o[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        o[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        o[2] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        o[3] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    }
    __device__ __forceinline__ static int sub_of(int p) { return p >> 2; }
    __device__ __forceinline__ static uint32_t sub_scale(const uint8_t *b, int j) {     // {fp16(dl), -fp16(ml)}
        const float2 s = Unpack8<T_Q4_K>::sub_scale(b, j);
        return h2_as_u32(__floats2half2_rn(s.x, -s.y));
    }
    __device__ __forceinline__ static uint4 run(const uint8_t *b, int p, uint32_t sc) {
        const int il = p >> 3, hi = (p >> 2) & 1, l0 = (p & 3) * 8;
        const uint2 w = *reinterpret_cast<const uint2 *>(b + 16 + 32 * il + l0);
        const uint32_t t0 = (w.x >> (4 * hi)) & 0x0f0f0f0fu, t1 = (w.y >> (4 * hi)) & 0x0f0f0f0fu;
        const __half2 dl = u32_as_h2(__byte_perm(sc, sc, 0x1010)), nml = u32_as_h2(__byte_perm(sc, sc, 0x3232));
        const __half2 k1024 = u32_as_h2(0x64006400u);
        uint4 o;
        o.x = h2_as_u32(__hfma2(__hsub2(magic_h2_lo(t0), k1024), dl, nml));
        o.y = h2_as_u32(__hfma2(__hsub2(magic_h2_hi(t0), k1024), dl, nml));
        o.z = h2_as_u32(__hfma2(__hsub2(magic_h2_lo(t1), k1024), dl, nml));
        o.w = h2_as_u32(__hfma2(__hsub2(magic_h2_hi(t1), k1024), dl, nml));
        return o;
    }
};

template <> struct HalfUnpack<T_Q5_K> {
    static constexpr bool HAS = true, FAST32 = false;
    static constexpr int NSUB = 8;
    __device__ __forceinline__ static int sub_of(int p) { return p >> 2; }
    __device__ __forceinline__ static uint32_t sub_scale(const uint8_t *b, int j) { return HalfUnpack<T_Q4_K>::sub_scale(b, j); }
    __device__ __forceinline__ static uint4 run(const uint8_t *b, int p, uint32_t sc) {
        const int il = p >> 3, hi = (p >> 2) & 1, l0 = (p & 3) * 8, j = p >> 2;
        const uint2 w = *reinterpret_cast<const uint2 *>(b + 48 + 32 * il + l0);
        const uint2 h = *reinterpret_cast<const uint2 *>(b + 16 + l0);
        const uint32_t t0 = ((w.x >> (4 * hi)) & 0x0f0f0f0fu) | (((h.x >> j) & 0x01010101u) << 4);
        const uint32_t t1 = ((w.y >> (4 * hi)) & 0x0f0f0f0fu) | (((h.y >> j) & 0x01010101u) << 4);
        const __half2 dl = u32_as_h2(__byte_perm(sc, sc, 0x1010)), nml = u32_as_h2(__byte_perm(sc, sc, 0x3232));
        const __half2 k1024 = u32_as_h2(0x64006400u);
        uint4 o;
        o.x = h2_as_u32(__hfma2(__hsub2(magic_h2_lo(t0), k1024), dl, nml));
        o.y = h2_as_u32(__hfma2(__hsub2(magic_h2_hi(t0), k1024), dl, nml));
        o.z = h2_as_u32(__hfma2(__hsub2(magic_h2_lo(t1), k1024), dl, nml));
        o.w = h2_as_u32(__hfma2(__hsub2(magic_h2_hi(t1), k1024), dl, nml));
        return o;
    }
};

// 16 bytes from a 2-byte-aligned shared-memory address as four words (the misalignment is warp-uniform in the MMQ tile)
__device__ __forceinline__ uint4 ld_u128_a2(const uint8_t *p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    if (a & 2) {
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
        return make_uint4(__funnelshift_r(w0, w1, 16), __funnelshift_r(w1, w2, 16), __funnelshift_r(w2, w3, 16),
                          __funnelshift_r(w3, w4, 16));
    }
    return make_uint4(q[0], q[1], q[2], q[3]);
}

template <> struct HalfUnpack<T_Q6_K> {
    static constexpr bool HAS = true, FAST32 = true;
    static constexpr int NSUB = 16;
    __device__ __forceinline__ static int sub_of(int p) { return p >> 1; }
    __device__ __forceinline__ static uint32_t sub_scale(const uint8_t *b, int is) {     // {fp16(dl), 0}
        const float2 s = Unpack8<T_Q6_K>::sub_scale(b, is);
        return h2_as_u32(__floats2half2_rn(s.x, 0.0f));
    }
    // k-step q (64 columns) of a row = the 32-value groups r = 2 (q & 1), 2 (q & 1) + 1 of half ip = q >> 1; this thread
    // takes l = 16 half .. 16 half + 15 of both: ql[64 ip + l] / ql[64 ip + 32 + l] (low nibbles for q even, high for q odd)
    // and qh[32 ip + l] bits 2r..2r+1.  sc = the row's scale-table words {2q, 2q+1 | 2q+2, 2q+3}: entry 4 (q & 1) ... the
    // caller passes the two table words of sub-blocks 8 ip + 2 r + half (first group) and 8 ip + 2 r + 2 + half (second).
    __device__ __forceinline__ static void run32(const uint8_t *b, int q, int half, uint2 sc, uint4 (&o)[4]) {
        const int ip = q >> 1, hi = q & 1;
        const uint4 la = ld_u128_a2(b + 64 * ip + 16 * half), lb = ld_u128_a2(b + 64 * ip + 32 + 16 * half);
        const uint4 qh = ld_u128_a2(b + 128 + 32 * ip + 16 * half);
        const __half2 dl0 = u32_as_h2(__byte_perm(sc.x, sc.x, 0x1010)), dl1 = u32_as_h2(__byte_perm(sc.y, sc.y, 0x1010));
        const __half2 k1056 = u32_as_h2(0x64206420u);                 // 1024 + 32: raw - 32 is the signed quant
        const uint32_t a[4] = {la.x, la.y, la.z, la.w}, c[4] = {lb.x, lb.y, lb.z, lb.w}, h[4] = {qh.x, qh.y, qh.z, qh.w};
        uint32_t lo[8], up[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t t0, t1;
            if (hi == 0) {
                t0 = (a[i] & 0x0f0f0f0fu) | ((h[i] << 4) & 0x30303030u);
                t1 = (c[i] & 0x0f0f0f0fu) | ((h[i] << 2) & 0x30303030u);
            } else {
                t0 = ((a[i] >> 4) & 0x0f0f0f0fu) | (h[i] & 0x30303030u);
                t1 = ((c[i] >> 4) & 0x0f0f0f0fu) | ((h[i] >> 2) & 0x30303030u);
            }
            lo[2 * i] = h2_as_u32(__hmul2(__hsub2(magic_h2_lo(t0), k1056), dl0));
            lo[2 * i + 1] = h2_as_u32(__hmul2(__hsub2(magic_h2_hi(t0), k1056), dl0));
            up[2 * i] = h2_as_u32(__hmul2(__hsub2(magic_h2_lo(t1), k1056), dl1));
            up[2 * i + 1] = h2_as_u32(__hmul2(__hsub2(magic_h2_hi(t1), k1056), dl1));
        }
        o[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        o[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        o[2] = make_uint4(up[0], up[1], up[2], up[3]);
        o[3] = make_uint4(up[4], up[5], up[6], up[7]);
    }
    __device__ __forceinline__ static uint4 run(const uint8_t *b, int p, uint32_t sc) {
        const int ip = p >> 4, r = (p >> 2) & 3, l0 = (p & 3) * 8;
        const uint2 ql = ld_u64_a2(b + 64 * ip + 32 * (r & 1) + l0);
        const uint2 qh = ld_u64_a2(b + 128 + 32 * ip + l0);
        const int sl = 4 * (r >> 1), shh = 2 * r;
        const uint32_t t0 = ((ql.x >> sl) & 0x0f0f0f0fu) | (((qh.x >> shh) & 0x03030303u) << 4);
        const uint32_t t1 = ((ql.y >> sl) & 0x0f0f0f0fu) | (((qh.y >> shh) & 0x03030303u) << 4);
        const __half2 dl = u32_as_h2(__byte_perm(sc, sc, 0x1010));
        const __half2 k1056 = u32_as_h2(0x64206420u);                 // 1024 + 32: raw - 32 is the signed quant
        uint4 o;
        o.x = h2_as_u32(__hmul2(__hsub2(magic_h2_lo(t0), k1056), dl));
        o.y = h2_as_u32(__hmul2(__hsub2(magic_h2_hi(t0), k1056), dl));
        o.z = h2_as_u32(__hmul2(__hsub2(magic_h2_lo(t1), k1056), dl));
        o.w = h2_as_u32(__hmul2(__hsub2(magic_h2_hi(t1), k1056), dl));
        return o;
    }
};
