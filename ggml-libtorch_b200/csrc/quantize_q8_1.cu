// quantize_q8_1.cu -- standalone q8_1 / q8_1_mmq activation quantizers (bit-exact restatement of
// quantize_q8_1, hf-kernels/ggml-kernels/ggml/ggml_kernel.cu:13-66, and quantize_mmq_q8_1, ggml/mmq.cu:109-177).
//
// The reference maps one thread to one value (2-byte loads, 1-byte stores).  Here 4 lanes own one
// 32-value group, each lane loading 8 consecutive values with one 128-bit load and storing its
// 8 quants with one 64-bit store (2 x 32-bit for the 36-byte q8_1 blocks, which are only 4-byte
// aligned).  HBM-bound: algorithmic bytes per value = sizeof(x) read + 1.125 written.
#include "q8_1.cuh"

namespace {

constexpr int kThreads = 256;

// layout GGB_Q8_1: out = rows x (Kp/32) blocks of {half d, half s, int8 qs[32]}, Kp = ceil(K/512)*512
template <typename T>
__global__ void __launch_bounds__(kThreads)
quantize_q8_1_kernel(const T *__restrict__ x, uint8_t *__restrict__ out, int64_t rows, int K, int Kp, bool vec_ok) {
    const int groups_per_row = Kp / 32;
    const int64_t gid = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 2;   // group index
    const int t = threadIdx.x & 3;
    const int64_t total = rows * groups_per_row;
    const bool live = gid < total;                      // whole warps stay converged for the shuffles
    const int64_t row = live ? gid / groups_per_row : 0;
    const int g = live ? (int)(gid % groups_per_row) : 0;
    const int k0 = g * 32 + t * 8;

    float v[8];
    if (live && k0 + 8 <= K) {
        load8_f32<T>(x + row * K + k0, vec_ok, v);
    } else {                                            // zero padding past K (ggml_kernel.cu:29)
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (live && k0 + i < K) ? to_f32<T>(x[row * K + k0 + i]) : 0.0f;
    }
    const Q8Group r = q8_1_quantize_group<true>(v);
    if (!live) return;
    uint8_t *blk = out + (row * groups_per_row + g) * 36;
    uint32_t *qs = reinterpret_cast<uint32_t *>(blk + 4 + 8 * t);
    qs[0] = r.q_lo;
    qs[1] = r.q_hi;
    if (t == 0) *reinterpret_cast<__half2 *>(blk) = __floats2half2_rn(r.d, r.sum);
}

// layout GGB_Q8_1_MMQ: block (k/128)*M + row of {half2 ds[4] | float d[4]; int8 qs[128]}, Kp = K - K%512 + 512
template <typename T, bool NEED_SUM>
__global__ void __launch_bounds__(kThreads)
quantize_q8_1_mmq_kernel(const T *__restrict__ x, uint8_t *__restrict__ out, int64_t M, int K, int Kp, bool vec_ok) {
    const int groups_per_row = Kp / 32;
    const int64_t gid = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 2;
    const int t = threadIdx.x & 3;
    const int64_t total = M * groups_per_row;
    const bool live = gid < total;
    const int64_t row = live ? gid / groups_per_row : 0;
    const int g = live ? (int)(gid % groups_per_row) : 0;
    const int k0 = g * 32 + t * 8;

    float v[8];
    if (live && k0 + 8 <= K) {
        load8_f32<T>(x + row * K + k0, vec_ok, v);
    } else {                                            // zero padding past K (ggml_kernel.cu:29)
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (live && k0 + i < K) ? to_f32<T>(x[row * K + k0 + i]) : 0.0f;
    }
    const Q8Group r = q8_1_quantize_group<NEED_SUM>(v);
    if (!live) return;
    uint8_t *blk = out + ((int64_t)(g >> 2) * M + row) * 144;
    const int gi = g & 3;
    *reinterpret_cast<uint2 *>(blk + 16 + 32 * gi + 8 * t) = make_uint2(r.q_lo, r.q_hi);
    if (t == 0) {
        if (NEED_SUM) *reinterpret_cast<__half2 *>(blk + 4 * gi) = __floats2half2_rn(r.d, r.sum);
        else          *reinterpret_cast<float *>(blk + 4 * gi) = r.d;
    }
}

template <typename T>
int launch(const void *X, int64_t rows, int64_t K, void *out, int layout, bool need_sum, cudaStream_t stream) {
    const T *x = static_cast<const T *>(X);
    const bool vec_ok = (reinterpret_cast<uintptr_t>(X) % (8 * sizeof(T)) == 0) && (K % 8 == 0);
    const int64_t Kp = (layout == GGB_Q8_1) ? (K + 511) / 512 * 512 : (K - K % 512 + 512);
    const int64_t nthreads = rows * (Kp / 32) * 4;
    const int64_t nblk = (nthreads + kThreads - 1) / kThreads;
    if (nblk == 0) return GGB_OK;
    if (nblk > 0x7fffffffLL || Kp > 0x7fffffffLL) {
        ggb_set_error("ggb_quantize_q8_1: problem too large (%lld x %lld)", (long long)rows, (long long)K);
        return GGB_E_SHAPE;
    }
    uint8_t *o = static_cast<uint8_t *>(out);
    if (layout == GGB_Q8_1) {
        quantize_q8_1_kernel<T><<<(unsigned)nblk, kThreads, 0, stream>>>(x, o, rows, (int)K, (int)Kp, vec_ok);
    } else if (need_sum) {
        quantize_q8_1_mmq_kernel<T, true><<<(unsigned)nblk, kThreads, 0, stream>>>(x, o, rows, (int)K, (int)Kp, vec_ok);
    } else {
        quantize_q8_1_mmq_kernel<T, false><<<(unsigned)nblk, kThreads, 0, stream>>>(x, o, rows, (int)K, (int)Kp, vec_ok);
    }
    GGB_CUDA_TRY(cudaGetLastError());
    return GGB_OK;
}

}  // namespace

bool ggb_mmq_need_sum(int type) {      // mmq_need_sum, ggml/mmq.cu:84-106
    return type == T_Q4_0 || type == T_Q4_1 || type == T_Q5_1 || type == T_Q4_K || type == T_Q5_K;
}

int ggb_quantize_q8_1_impl(const void *X, int x_dtype, int64_t rows, int64_t K, void *out, int layout,
                           int wtype, cudaStream_t stream) {
    const bool need_sum = ggb_mmq_need_sum(wtype);
    switch (x_dtype) {
        case GGB_F32:  return launch<float>(X, rows, K, out, layout, need_sum, stream);
        case GGB_F16:  return launch<__half>(X, rows, K, out, layout, need_sum, stream);
        case GGB_BF16: return launch<__nv_bfloat16>(X, rows, K, out, layout, need_sum, stream);
        default:
            ggb_set_error("ggb_quantize_q8_1: unknown x_dtype %d", x_dtype);
            return GGB_E_DTYPE;
    }
}
