// ggb_common.cuh -- shared device/host helpers for libggml_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/ggml_b200.h"

// ---------------------------------------------------------------- ggml type ids (ggml-common.h:1128-1161)
enum : int {
    T_Q4_0 = 2, T_Q4_1 = 3, T_Q5_0 = 6, T_Q5_1 = 7, T_Q8_0 = 8,
    T_Q2_K = 10, T_Q3_K = 11, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14,
    T_IQ4_NL = 20, T_IQ4_XS = 23,         // dequantize + MMVQ only, like the reference (no MMQ case: mmq.cu:222-252)
    // the codebook formats (iq_grid.cuh): dequantize + MMVQ only as well
    T_IQ2_XXS = 16, T_IQ2_XS = 17, T_IQ3_XXS = 18, T_IQ1_S = 19, T_IQ3_S = 21, T_IQ2_S = 22, T_IQ1_M = 29
};

template <int TYPE> struct BlockTraits;
template <> struct BlockTraits<T_Q4_0> { static constexpr int QK = 32,  BS = 18;  };
template <> struct BlockTraits<T_Q4_1> { static constexpr int QK = 32,  BS = 20;  };
template <> struct BlockTraits<T_Q5_0> { static constexpr int QK = 32,  BS = 22;  };
template <> struct BlockTraits<T_Q5_1> { static constexpr int QK = 32,  BS = 24;  };
template <> struct BlockTraits<T_Q8_0> { static constexpr int QK = 32,  BS = 34;  };
template <> struct BlockTraits<T_Q2_K> { static constexpr int QK = 256, BS = 84;  };
template <> struct BlockTraits<T_Q3_K> { static constexpr int QK = 256, BS = 110; };
template <> struct BlockTraits<T_Q4_K> { static constexpr int QK = 256, BS = 144; };
template <> struct BlockTraits<T_Q5_K> { static constexpr int QK = 256, BS = 176; };
template <> struct BlockTraits<T_Q6_K> { static constexpr int QK = 256, BS = 210; };
template <> struct BlockTraits<T_IQ4_NL> { static constexpr int QK = 32,  BS = 18;  };
template <> struct BlockTraits<T_IQ4_XS> { static constexpr int QK = 256, BS = 136; };
template <> struct BlockTraits<T_IQ2_XXS> { static constexpr int QK = 256, BS = 66;  };
template <> struct BlockTraits<T_IQ2_XS>  { static constexpr int QK = 256, BS = 74;  };
template <> struct BlockTraits<T_IQ2_S>   { static constexpr int QK = 256, BS = 82;  };
template <> struct BlockTraits<T_IQ3_XXS> { static constexpr int QK = 256, BS = 98;  };
template <> struct BlockTraits<T_IQ3_S>   { static constexpr int QK = 256, BS = 110; };
template <> struct BlockTraits<T_IQ1_S>   { static constexpr int QK = 256, BS = 50;  };
template <> struct BlockTraits<T_IQ1_M>   { static constexpr int QK = 256, BS = 56;  };
template <int TYPE> constexpr bool kIsIqGrid = (TYPE == T_IQ2_XXS || TYPE == T_IQ2_XS || TYPE == T_IQ2_S || TYPE == T_IQ3_XXS ||
                                                TYPE == T_IQ3_S || TYPE == T_IQ1_S || TYPE == T_IQ1_M);

// ---------------------------------------------------------------- host-side error plumbing (api.cu owns the storage)
void ggb_set_error(const char *fmt, ...);

#define GGB_CUDA_TRY(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ggb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return GGB_E_CUDA;                                                          \
        }                                                                               \
    } while (0)

struct DeviceInfo {
    int dev;             // CUDA device ordinal (0..63), index for per-device one-time function attributes
    int sm_count;
    int max_smem_optin;
};
// cached per device after the first call (the reference re-queries on every launch, mmq.cu:17-54)
const DeviceInfo &ggb_device_info();
// ggb_set_option(GGB_OPT_STATIC_WEIGHTS): weight tensors are never written by work enqueued shortly before
int ggb_static_weights();

// Every kernel of the library asks for the LARGEST shared-memory carve-out.  An SM changes its L1 / shared-memory split only while
// it is idle, so a kernel whose preferred split differs from the running kernel's cannot become resident beside it: under
// programmatic dependent launch the next kernel then starts only when the SM has drained -- and its CTAs land two per SM on the
// SMs that drain first (measured with %smid stamps in the 70B block, profiles/r02_block_trace.log: 72 SMs with two CTAs of a
// launch, 72 with none, and the pattern persists because every launch inherits the slots of the one before the last).
// Once per kernel instantiation and device; idempotent, so a race between threads is harmless.
template <auto Kern>                         // (the kernel itself is the template argument: one flag per kernel, not per signature)
inline cudaError_t ggb_prefer_max_smem_once(int dev) {
    static bool done[64] = {};
    if (dev < 0 || dev >= 64 || done[dev]) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess) done[dev] = true;
    return e;
}

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ float ggb_h2f(uint16_t bits) {
    return __half2float(__ushort_as_half(bits));          // exact widening
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 2-byte-aligned little-endian reads from shared or global memory
__device__ __forceinline__ uint32_t ld_u16(const uint8_t *p) {
    return *reinterpret_cast<const uint16_t *>(p);
}
__device__ __forceinline__ uint32_t ld_u32_a2(const uint8_t *p) {   // 2-byte aligned
    return ld_u16(p) | (ld_u16(p + 2) << 16);
}
__device__ __forceinline__ uint32_t ld_u32_a4(const uint8_t *p) {   // 4-byte aligned
    return *reinterpret_cast<const uint32_t *>(p);
}

// IQ4_NL / IQ4_XS: the 16-entry non-linear codebook (format constant kvalues_iq4nl, ggml-common.h:1060 =
// gguf-py IQ4_NL.kvalues) as four little-endian words of int8 bytes.  `w` holds four 4-bit codes, one per byte; the
// result holds their codebook values as int8 bytes: two PRMTs look the low 3 bits up in each table half, bit 3 selects.
__device__ __forceinline__ uint32_t iq4nl_lut4(uint32_t w) {
    const uint32_t s = w & 0x07070707u, t = s | (s >> 4);
    const uint32_t sel = __byte_perm(t, 0u, 0x4420);                      // code i -> selector nibble i
    const uint32_t lo = __byte_perm(0xBFAD9881u, 0xF6EADDCFu, sel);       // -127 -104 -83 -65 | -49 -35 -22 -10
    const uint32_t hi = __byte_perm(0x26190D01u, 0x71594535u, sel);       //    1   13  25  38 |  53  69  89 113
    const uint32_t m = ((w >> 3) & 0x01010101u) * 0xffu;
    return (hi & m) | (lo & ~m);
}
// The same lookup for all eight codes of a raw qs word at once (MMVQ fast path): the 3-bit nibbles of q & 0x77777777 ARE
// PRMT selectors, so four PRMTs look bytes [l0 h0 l1 h1] / [l2 h2 l3 h3] up in both table halves, four more split them into
// low-nibble and high-nibble words, and bit 3 of every code picks the half through a sign-replicating PRMT (selector
// nibbles 8..B; inline PTX because __byte_perm only honours three selector bits).  15 instructions per 8 values.
__device__ __forceinline__ uint32_t ggb_prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ void iq4nl_lut8(uint32_t q, uint32_t &vlo, uint32_t &vhi) {
    const uint32_t s = q & 0x77777777u, s2 = s >> 16;
    const uint32_t al = ggb_prmt(0xBFAD9881u, 0xF6EADDCFu, s), bl = ggb_prmt(0xBFAD9881u, 0xF6EADDCFu, s2);
    const uint32_t ah = ggb_prmt(0x26190D01u, 0x71594535u, s), bh = ggb_prmt(0x26190D01u, 0x71594535u, s2);
    const uint32_t mlo = ggb_prmt(q << 4, 0u, 0xBA98u), mhi = ggb_prmt(q, 0u, 0xBA98u);
    const uint32_t lo_l = ggb_prmt(al, bl, 0x6420u), lo_h = ggb_prmt(ah, bh, 0x6420u);
    const uint32_t hi_l = ggb_prmt(al, bl, 0x7531u), hi_h = ggb_prmt(ah, bh, 0x7531u);
    vlo = (lo_h & mlo) | (lo_l & ~mlo);
    vhi = (hi_h & mhi) | (hi_l & ~mhi);
}
// IQ4_XS 6-bit sub-block scale ib (0..7) with the -32 offset applied (dequantize.cuh:426-427)
__device__ __forceinline__ int iq4xs_scale(const uint8_t *b, int ib) {
    const int lo = (b[4 + (ib >> 1)] >> (4 * (ib & 1))) & 0x0f;
    const int hi = (ld_u16(b + 2) >> (2 * ib)) & 3;
    return (lo | (hi << 4)) - 32;
}

// ---------------------------------------------------------------- output conversion: ONE round-to-nearest from fp32
template <typename T> struct OutCvt;
template <> struct OutCvt<float> {
    static constexpr int DT = GGB_F32;
};
template <> struct OutCvt<__half> {
    static constexpr int DT = GGB_F16;
};
template <> struct OutCvt<__nv_bfloat16> {
    static constexpr int DT = GGB_BF16;
};

__device__ __forceinline__ uint32_t pack2(float a, float b, __half *) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t pack2(float a, float b, __nv_bfloat16 *) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// store 8 consecutive outputs (16-byte aligned destination)
__device__ __forceinline__ void store8(float *dst, const float (&v)[8]) {
    reinterpret_cast<float4 *>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4 *>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <typename T>
__device__ __forceinline__ void store8(T *dst, const float (&v)[8]) {
    uint4 u;
    u.x = pack2(v[0], v[1], (T *)nullptr);
    u.y = pack2(v[2], v[3], (T *)nullptr);
    u.z = pack2(v[4], v[5], (T *)nullptr);
    u.w = pack2(v[6], v[7], (T *)nullptr);
    *reinterpret_cast<uint4 *>(dst) = u;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---------------------------------------------------------------- output target of the MMVQ kernels
// Y[tok][col0 + row] with leading dimension ld, replicated into `n` buffers: n = 1 is the plain op; n = tp with peer
// pointers (NVLink P2P mappings of a symmetric buffer) makes the kernel's epilogue the all-gather of a column-parallel
// layer (ggb_mmvq_gather): every rank writes its slice straight into every rank's copy of the full output.
struct GgbYOut {
    void *base[8];
    int n;
    long long ld, col0;
    long long split, col1;     // rows >= split go to column col1 + (row - split): two stacked projections (gate | up)
};
// Activation source of the MMVQ kernels: X [B, K] with row stride ldx; with x2 != nullptr the activation is
// SwiGLU(X, X2) = silu(X) * X2 computed on the fly with torch's eager roundings (silu and the product each rounded to
// the tensor dtype), so a fused down-projection equals  mmvq(W, F.silu(g) * u)  bit for bit.
struct GgbXIn {
    const void *x2;
    long long ldx;
};
// thread-local targets installed by ggb_mmvq_fused for the duration of its call (nullptr: the plain arguments)
const GgbYOut *ggb_current_yout();
const GgbXIn *ggb_current_xin();
inline GgbYOut ggb_make_yout(void *Y, long long N) {
    if (const GgbYOut *g = ggb_current_yout()) return *g;
    GgbYOut y = {};
    y.base[0] = Y;
    y.n = 1;
    y.ld = N;
    y.col0 = 0;
    y.split = 0x7fffffffffffffffLL;
    y.col1 = 0;
    return y;
}
inline GgbXIn ggb_make_xin(long long K) {
    if (const GgbXIn *x = ggb_current_xin()) return *x;
    GgbXIn x = {};
    x.x2 = nullptr;
    x.ldx = K;
    return x;
}
#ifdef __CUDACC__
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <typename T>
__device__ __forceinline__ void ggb_yout_store(const GgbYOut &y, long long tok, long long row, float v) {
    const T h = from_f32<T>(v);
    const long long idx = tok * y.ld + (row < y.split ? y.col0 + row : y.col1 + (row - y.split));
    static_cast<T *>(y.base[0])[idx] = h;
    // (the plain op must not pay for the gather: a predicated 8-way loop here cost the batch-2 dp4a kernel 19 %)
    if (y.n > 1) {
#pragma unroll 1
        for (int i = 1; i < y.n; i++) static_cast<T *>(y.base[i])[idx] = h;
    }
}
// silu(g) * u as torch eager computes it for tensors of dtype T: x / (1 + exp(-x)) in fp32, rounded to T, times u in
// fp32, rounded to T  (aten/native/cuda/ActivationSiluKernel.cu, BinaryMulKernel.cu; identity roundings for fp32)
template <typename T>
__device__ __forceinline__ float ggb_swiglu(float g, float u) {
    const float s = to_f32<T>(from_f32<T>(g / (1.0f + expf(-g))));
    return to_f32<T>(from_f32<T>(s * u));
}
#endif

// ---------------------------------------------------------------- programmatic dependent launch
// Kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may become resident
// while the previous kernel of the stream drains.  ggb_pdl_wait() returns once that kernel has completed and its
// writes are visible: nothing it may have produced is read, and nothing it may still read is written, before it.
__device__ __forceinline__ void ggb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void ggb_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier + bulk-copy (TMA 1-D) PTX
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// global -> shared::cta bulk copy, completion counted on `bar`.  16-byte aligned src/dst, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// same with an L2 evict-first policy: weights/blocks that are streamed exactly once
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// pull `bytes` (multiple of 16, 16-byte aligned source) into L2 without a destination: a pure hint
__device__ __forceinline__ void bulk_prefetch_l2(const void *gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
