// mmvq_rgx.cu -- the register-resident batch-1 MMVQ design of mmvq_rg.cu for the types without a hand-tuned item routine:
// Q4_0, Q4_1, Q5_0, Q8_0, Q2_K, IQ4_NL by default (Q5_1, Q3_K, Q5_K, Q6_K, IQ4_XS on request: GGB_OPT_MMVQ_RG bit 1).
// Measured, batch 1, 14336 x 4096, vs the generic kernel (profiles/r02_mmvq_rgx.log): Q4_0 12.7 -> 7.8 us (0.65 of the HBM peak),
// Q4_1 11.7 -> 8.5 (0.66), Q5_0 14.9 -> 10.1 (0.61), Q8_0 15.0 -> 11.2 (0.85), Q2_K 13.1 -> 10.6, IQ4_NL 16.8 -> 12.1.
//
// Same contract as every MMVQ kernel (ggml_mul_mat_vec_a8 = quantize_row_q8_1_cuda + mul_mat_vec_q + vec_dot_*_q8_1,
// hf-kernels/ggml-kernels/ggml/ggml_kernel.cu:13-66,80-193; ggml/mmvq.cuh:2-128; ggml/vecdotq.cuh:45-349).
//
// Round 1's generic kernel (mmvq.cu: producer warp + CTA-wide stage ring, one lane per 32-value group, activations read
// from shared memory for every group) sat at 0.19-0.61 of the HBM peak at batch 1.  What made the Q4_K kernel fast
// carries over unchanged: a warp owns chunk c (16 units of 256 values) of its rows and a lane ONE
// 128-value item of it, whose q8_1 activations stay in registers for the whole launch; row chunks arrive through a private
// ring of coalesced 16-byte cp.async copies; this launch and the next are co-resident (programmatic dependent launch).
// Per item the lane unpacks its four 32-value groups with the same per-type routines as the generic kernel
// (mmvq_group_ops.cuh: load_group / combine restate the reference's vec_dot algebra) -- fewer instructions than a tuned
// routine would need is left on the table, the structure is what counts.
#include <type_traits>

#include "mmvq_group_ops.cuh"

namespace ggb_sb {
extern int g_pdl_prefetch;
extern int g_rg_mode;
}

namespace {
using namespace ggb_group;

constexpr int kNW = 12;                  // warps per CTA

struct RgxParams {
    const uint8_t *W;
    const void *X;
    GgbXIn xin;
    GgbYOut yo;
    int64_t N;
    int K, nsb, ipr, row_bytes;          // nsb = units of 256 values per row, ipr = items (128 values) per row
    int C, RW, spc;                      // chunks per row, row lanes per CTA = kNW / C, units per chunk
    int off_scales, off_part, off_slots;
    int part_rows;
    int x_vec;
    int pdl_prefetch;
    long long wstep;                     // bytes between two consecutive rows of a warp
};

__device__ __forceinline__ void cp_async16_if(int have, int need, uint32_t smem_dst, const void *gsrc, uint64_t policy) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.s32 p, %3, %4;\n\t"
                 "@p cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n\t}"
                 ::"r"(smem_dst), "l"(gsrc), "l"(policy), "r"(have), "r"(need) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bytes per 256 values
template <int TYPE> constexpr int unit_bytes() { return 256 / BlockTraits<TYPE>::QK * BlockTraits<TYPE>::BS; }

// the item's contribution: groups 4 * it .. 4 * it + 3 of the chunk in `stage`
template <int TYPE>
__device__ __forceinline__ float item(const uint8_t *stage, int it, const uint32_t (&a)[32], const float (&d8)[4], const uint32_t (&aux)[4]) {
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        GroupW w;
        load_group<TYPE>(stage, 4 * it + j, w);
        int dot0 = 0, dot1 = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            dot0 = __dp4a((int)w.w[i], (int)a[8 * j + i], dot0);
            dot1 = __dp4a((int)w.w[4 + i], (int)a[8 * j + 4 + i], dot1);
        }
        sum += combine<TYPE>(w, dot0, dot1, d8[j], aux[j]);
    }
    return sum;
}

template <int TYPE, typename T, int S>
__global__ void __launch_bounds__(kNW * 32, 2)
mmvq_rgx_kernel(const RgxParams p) {
    constexpr int UB = unit_bytes<TYPE>();
    constexpr int stage_bytes = 16 * UB;                       // always 16 units wide: every offset is an immediate
    constexpr int NP = (stage_bytes / 16 + 31) / 32;           // 16-byte pieces per lane and row chunk
    extern __shared__ __align__(128) uint8_t smem[];
    uint4 *act_q = reinterpret_cast<uint4 *>(smem);                            // [8][ipr]: 16-value chunk i of item it at i * ipr + it
    uint4 *act_s = reinterpret_cast<uint4 *>(smem + p.off_scales);             // [2][ipr]: d8 of the item's 4 groups | their aux words
    float *part_sm = reinterpret_cast<float *>(smem + p.off_part);
    uint8_t *ring = smem + p.off_slots;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ipr = p.ipr, C = p.C, RW = p.RW;
    const int64_t r_begin = p.N * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = p.N * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int nrows = (int)(r_end - r_begin);
    ggb_pdl_launch_dependents();

    const int c = warp % C, rl = warp / C;
    const int my_rows = (rl < RW && nrows > rl) ? (nrows - rl + RW - 1) / RW : 0;
    const int sb0 = c * p.spc;
    const int nsb_c = min(p.spc, p.nsb - sb0);
    const int n16 = nsb_c * UB / 16;                            // (the launcher guarantees nsb_c * UB % 16 == 0)
    const int il = lane;                                       // item inside the chunk: half (lane & 1) of unit (lane >> 1)
    const bool live = (lane >> 1) < nsb_c;
    const int item_row = 2 * sb0 + il;                          // item inside the row
    uint8_t *my_ring = ring + (size_t)warp * S * stage_bytes;
    const uint32_t my_ring_s = smem_u32(my_ring) + lane * 16;
    const uint64_t pol = l2_policy_evict_first();
    const uint8_t *src_next = p.W + (r_begin + rl) * (int64_t)p.row_bytes + (int64_t)sb0 * UB + lane * 16;
    int to_issue = my_rows;
    auto issue_row = [&](uint32_t dst) {
        const int have = to_issue > 0 ? (n16 - lane + 31) >> 5 : 0;
#pragma unroll
        for (int q = 0; q < NP; q++) cp_async16_if(have, q, dst + 512 * q, src_next + 512 * q, pol);
        cp_async_commit();
        src_next += p.wstep;
        to_issue--;
    };
    if (p.pdl_prefetch) issue_row(my_ring_s);
    ggb_pdl_wait();

    // ===================== q8_1-quantize X into shared memory (bit-identical to quantize_q8_1) =====================
    const T *X = static_cast<const T *>(p.X);
    const int ngroups_row = p.K >> 5;
    constexpr int GPI = (kNW * 32) / 4;
    const int t4 = lane & 3;
    for (int base = 0; base < ngroups_row; base += 2 * GPI) {
        float v[2][8];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int g = base + u * GPI + (tid >> 2);
            if (g < ngroups_row) {
                load_act8<T>(X, static_cast<const T *>(p.xin.x2), p.xin.ldx, 0, g * 32 + t4 * 8, p.x_vec != 0, v[u]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) v[u][i] = 0.0f;
            }
        }
        if (base == 0) {
#pragma unroll
            for (int s = 0; s < S; s++)
                if (s >= (p.pdl_prefetch ? 1 : 0)) issue_row(my_ring_s + s * stage_bytes);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (base + u * GPI >= ngroups_row) break;             // warp-uniform
            const int g = base + u * GPI + (tid >> 2);
            const Q8Group q = q8_1_quantize_group<VqTraits<TYPE>::NEED_S8>(v[u]);
            uint32_t aux = 0;
            if constexpr (VqTraits<TYPE>::NEED_SQ) {
                int mine = __dp4a((int)q.q_lo, 0x01010101, __dp4a((int)q.q_hi, 0x01010101, 0));
                mine += __shfl_xor_sync(0xffffffffu, mine, 1);                 // 16-value half sums
                const int other = __shfl_xor_sync(0xffffffffu, mine, 2);
                aux = ((uint32_t)mine & 0xffffu) | ((uint32_t)other << 16);   // valid on t4 == 0
            } else if constexpr (VqTraits<TYPE>::NEED_S8) {
                aux = __float_as_uint(__half2float(__float2half_rn(q.sum)));
            }
            if (g < ngroups_row) {
                const int it = g >> 2, j = g & 3;
                const int i = 2 * j + (t4 >> 1);
                uint8_t *dst = reinterpret_cast<uint8_t *>(act_q + i * ipr + it) + 8 * (t4 & 1);
                *reinterpret_cast<uint2 *>(dst) = make_uint2(q.q_lo, q.q_hi);
                if (t4 == 0) {
                    reinterpret_cast<float *>(act_s + it)[j] = __half2float(__float2half_rn(q.d));
                    reinterpret_cast<uint32_t *>(act_s + ipr + it)[j] = aux;
                }
            }
        }
    }
    __syncthreads();

    uint32_t a[32];
    float d8[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t aux[4] = {0u, 0u, 0u, 0u};
    if (live) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint4 v = act_q[i * ipr + item_row];
            a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
        }
        const uint4 dv = act_s[item_row], av = act_s[ipr + item_row];
        d8[0] = __uint_as_float(dv.x); d8[1] = __uint_as_float(dv.y); d8[2] = __uint_as_float(dv.z); d8[3] = __uint_as_float(dv.w);
        aux[0] = av.x; aux[1] = av.y; aux[2] = av.z; aux[3] = av.w;
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++) a[i] = 0u;
    }

    // ===================== main loop: one step = chunk c of this warp's next row =====================
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    const int itl = live ? il : 0;                                  // lanes without an item read item 0; their result is dropped
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    int stage = 0;
    for (int j0 = 0; j0 < my_rows; j0 += 4) {
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            part[jj] = 0.0f;
            if (j0 + jj < my_rows) {                                   // warp-uniform
                cp_async_wait<S - 1>();
                __syncwarp();
                const float v = item<TYPE>(my_ring + stage * stage_bytes, itl, a, d8, aux);
                part[jj] = live ? v : 0.0f;
                __syncwarp();
                issue_row(my_ring_s + stage * stage_bytes);
                if (++stage == S) stage = 0;
            }
        }
        float x0 = (up16 ? part[2] : part[0]) + __shfl_xor_sync(0xffffffffu, up16 ? part[0] : part[2], 16);
        float x1 = (up16 ? part[3] : part[1]) + __shfl_xor_sync(0xffffffffu, up16 ? part[1] : part[3], 16);
        float v = (up8 ? x1 : x0) + __shfl_xor_sync(0xffffffffu, up8 ? x0 : x1, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        const int jr = j0 + (up16 ? 2 : 0) + (up8 ? 1 : 0);
        if ((lane & 7) == 0 && jr < my_rows) {
            const int lr = rl + jr * RW;
            if (C == 1) ggb_yout_store<T>(p.yo, 0, r_begin + lr, v);
            else part_sm[lr * C + c] = v;
        }
    }
    cp_async_wait<0>();
    if (C > 1) {
        __syncthreads();
        for (int lr = tid; lr < nrows; lr += kNW * 32) {
            float s = part_sm[lr * C];
            for (int cc = 1; cc < C; cc++) s += part_sm[lr * C + cc];
            ggb_yout_store<T>(p.yo, 0, r_begin + lr, s);
        }
    }
}

template <int TYPE, typename T, int S>
int launch_s(RgxParams &p, int grid, int smem, cudaStream_t stream) {
    const DeviceInfo &di = ggb_device_info();
    static bool attr_set[64] = {};               // per instantiation and device; idempotent, so a race is harmless
    if (!attr_set[di.dev]) {
        GGB_CUDA_TRY(cudaFuncSetAttribute(mmvq_rgx_kernel<TYPE, T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        attr_set[di.dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmvq_rgx_kernel<TYPE, T, S>>(di.dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmvq_rgx_kernel<TYPE, T, S>, p));
    return GGB_OK;
}

template <int TYPE, typename T>
int launch(const void *W, const void *X, void *Y, int64_t K, int64_t N, cudaStream_t stream) {
    constexpr int UB = unit_bytes<TYPE>();
    constexpr int stage_bytes = 16 * UB;
    const DeviceInfo &di = ggb_device_info();
    if (K % 256 != 0 || K > 16384) return 1;
    RgxParams p;
    p.W = static_cast<const uint8_t *>(W);
    p.X = X;
    p.yo = ggb_make_yout(Y, N);
    p.xin = ggb_make_xin(K);
    p.N = N;
    p.K = (int)K;
    p.nsb = (int)(K / 256);
    p.ipr = 2 * p.nsb;
    p.row_bytes = p.nsb * UB;
    p.C = (p.nsb + 15) / 16;
    if (p.C > 4 || kNW % p.C != 0 || p.row_bytes % 16 != 0) return 1;
    // balanced chunks when every chunk then starts and ends on a 16-byte boundary, else 16-unit chunks + a shorter last one
    p.spc = (p.nsb + p.C - 1) / p.C;
    if ((p.spc * UB) % 16 != 0 || ((p.nsb - (p.C - 1) * p.spc) * UB) % 16 != 0) {
        p.spc = 16;
        if (((p.nsb - (p.C - 1) * 16) * UB) % 16 != 0) return 1;
    }
    p.RW = kNW / p.C;
    p.wstep = (long long)p.RW * p.row_bytes;
    p.x_vec = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.xin.x2) & 15) == 0 &&
              (p.xin.ldx * (int64_t)sizeof(T)) % 16 == 0;
    p.pdl_prefetch = ggb_sb::g_pdl_prefetch;
    const int64_t max_ctas = (N + p.RW - 1) / p.RW;
    const int grid = (int)(max_ctas < di.sm_count ? max_ctas : di.sm_count);
    p.part_rows = p.C > 1 ? (int)((N + grid - 1) / grid + 1) : 0;
    if ((int64_t)p.part_rows * p.C * 4 > 16384) return 1;
    p.off_scales = ((int)K + 15) & ~15;
    p.off_part = p.off_scales + 2 * p.ipr * 16;
    p.off_slots = (p.off_part + p.part_rows * p.C * 4 + 127) & ~127;
    const int half_sm = (di.max_smem_optin + 1024) / 2 - 1024;               // two CTAs per SM
    int S = 4;
    while (S > 2 && p.off_slots + kNW * S * stage_bytes + 16 > half_sm) S--;
    const int smem = p.off_slots + kNW * S * stage_bytes + 16;       // (+16: the 2-byte-aligned word reads may look 4 bytes past a stage)
    if (smem > half_sm) return 1;
    switch (S) {
        case 4: return launch_s<TYPE, T, 4>(p, grid, smem, stream);
        case 3: return launch_s<TYPE, T, 3>(p, grid, smem, stream);
        default: return launch_s<TYPE, T, 2>(p, grid, smem, stream);
    }
}

template <int TYPE>
int dispatch_dtype(const void *W, const void *X, void *Y, int dt, int64_t K, int64_t N, cudaStream_t stream) {
    switch (dt) {
        case GGB_F32:  return launch<TYPE, float>(W, X, Y, K, N, stream);
        case GGB_F16:  return launch<TYPE, __half>(W, X, Y, K, N, stream);
        case GGB_BF16: return launch<TYPE, __nv_bfloat16>(W, X, Y, K, N, stream);
        default: return 1;
    }
}

}  // namespace

// returns GGB_OK / a negative status, or 1 when this kernel does not apply (the caller goes on to mmvq.cu)
int ggb_mmvq_rgx_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                      cudaStream_t stream) {
    if (!(ggb_sb::g_rg_mode & 1) || B != 1 || (reinterpret_cast<uintptr_t>(W) & 15)) return 1;
    const bool all = (ggb_sb::g_rg_mode & 2) != 0;            // debug: also the types the ring kernel serves
    switch (type) {
        case T_Q4_0: return dispatch_dtype<T_Q4_0>(W, X, Y, xy_dtype, K, N, stream);
        case T_Q4_1: return dispatch_dtype<T_Q4_1>(W, X, Y, xy_dtype, K, N, stream);
        case T_Q5_0: return dispatch_dtype<T_Q5_0>(W, X, Y, xy_dtype, K, N, stream);
        case T_Q5_1: return all ? dispatch_dtype<T_Q5_1>(W, X, Y, xy_dtype, K, N, stream) : 1;   // measured: equal on the 8B shapes, slower on 28672 x 8192 (two stages only)
        case T_Q8_0: return dispatch_dtype<T_Q8_0>(W, X, Y, xy_dtype, K, N, stream);
        case T_Q2_K: return dispatch_dtype<T_Q2_K>(W, X, Y, xy_dtype, K, N, stream);
        case T_Q3_K: return all ? dispatch_dtype<T_Q3_K>(W, X, Y, xy_dtype, K, N, stream) : 1;   // measured slower (22.7 vs 19.0 us): the unpack spills at 80 registers
        case T_IQ4_NL: return dispatch_dtype<T_IQ4_NL>(W, X, Y, xy_dtype, K, N, stream);
        case T_Q5_K: return all ? dispatch_dtype<T_Q5_K>(W, X, Y, xy_dtype, K, N, stream) : 1;
        case T_Q6_K: return all ? dispatch_dtype<T_Q6_K>(W, X, Y, xy_dtype, K, N, stream) : 1;
        case T_IQ4_XS: return all ? dispatch_dtype<T_IQ4_XS>(W, X, Y, xy_dtype, K, N, stream) : 1;
        default: return 1;
    }
}
