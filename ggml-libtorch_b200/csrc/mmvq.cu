// mmvq.cu -- quantized GEMV for decode batches 1..8 (MMVQ) on sm_100a.
//
// Replaces ggml_mul_mat_vec_a8 = quantize_row_q8_1_cuda + mul_mat_vec_q + the per-type vec_dot_*_q8_1
// (hf-kernels/ggml-kernels/ggml/ggml_kernel.cu:13-66, 80-193; ggml/mmvq.cuh:2-209; ggml/vecdotq.cuh:45-605).
// The reference runs two kernels (quantize, then one 32-thread CTA per weight row with 2/4-byte loads)
// and is limited to batch 1.  This is ONE launch:
//
//   * prologue: the CTA quantizes the B x K activations to q8_1 in shared memory (int8 quants, the
//     fp16-rounded scale d and, per weight type, the fp16-rounded sum s or the integer half-sums of
//     the quants) -- bit-identical to quantize_q8_1 (q8_1.cuh);
//   * weight streaming: one CTA per SM owns a contiguous, byte-balanced range of weight rows.  A
//     producer warp moves (32 rows x 2048-column) stages of raw GGUF bytes into a shared-memory ring
//     with cp.async.bulk (TMA engine, mbarrier completion, L2 evict-first): no per-thread load
//     instructions, 100-200 KB in flight per SM;
//   * compute: 16 consumer warps, two rows per warp (each activation word fetched from shared memory
//     is used for both rows), one lane per 32-value group: raw quants unpacked in registers, __dp4a
//     integer dots, fp32 scale/min algebra in the reference's per-type form, warp-shuffle reduction.
//
// Algorithmic bytes per launch: N*K/QK*block_bytes (weights, streamed once).  Tolerance contract
// (tests/test_gpu_mmvq.py): integer dots are exact; only the fp32 summation order differs from the
// reference, so results agree with the oracle to ~1e-6 relative of the output scale.
#include "q8_1.cuh"
#include "tma2d.cuh"

namespace {

constexpr int kConsumerWarps = 16;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kRowsPerWarp = 2;
constexpr int kStageRows = kConsumerWarps * kRowsPerWarp;       // 32
constexpr int kChunkElems = 2048;                               // columns per stage
constexpr int kChunkGroups = kChunkElems / 32;
constexpr int kMaxStages = 8;

struct Params {
    const uint8_t *W;
    const void *X;
    void *Y;
    int64_t N;
    int K;
    int row_bytes;       // K / QK * BS
    int chunk_bytes;     // kChunkElems / QK * BS
    int nchunks;
    int seg_stride;      // shared-memory bytes per staged row segment (multiple of 16)
    int nstages;
    int fast;            // rows and segments are 16-byte aligned: TMA bulk copies
    int nbox;            // > 0: a stage arrives as nbox 2-D tiled TMA boxes of (32 rows x box_bytes), one instruction each
    int box_bytes;       //      (a bulk copy per row costs ~35 serialised instructions: UBLKCP takes uniform operands);
    int gpb;             //      a box holds gpb = 64 / nbox whole 32-value groups of every row: [box][row][box_bytes]
    int x_vec;           // X is 16-byte aligned
    int off_scales;      // byte offsets into dynamic shared memory
    int off_stages;
    int off_bars;
};

// Grouped (mixture-of-experts) launches only: blockIdx.y is one (token, expert) pair; see ggb_moe_mmvq_impl below.
struct MoeParams {
    const int32_t *ids;  // [pairs] expert of pair i (token i / top_k)
    int64_t expert_stride;   // bytes between the [N, row_bytes] matrices of consecutive experts
    int64_t ldy;         // elements between the output rows of consecutive pairs
    int n_experts;
    int top_k;
};

// ---------------------------------------------------------------- per-type traits
}  // namespace
#include "mmvq_group_ops.cuh"
namespace {
using namespace ggb_group;

// ---------------------------------------------------------------- the kernel
template <int TYPE, int B, typename T, bool MOE>
__device__ __forceinline__ void mmvq_body(const Params &p, const CUtensorMap &tmapW, const MoeParams &moe) {
    extern __shared__ __align__(128) uint8_t smem[];
    int8_t *act_q = reinterpret_cast<int8_t *>(smem);                          // [B][K]
    float2 *act_s = reinterpret_cast<float2 *>(smem + p.off_scales);           // [B][K/32] {d8, aux}
    uint8_t *stages = smem + p.off_stages;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.off_bars);
    uint64_t *empty = full + kMaxStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, ngroups_row = K >> 5;
    const int stage_bytes = kStageRows * p.seg_stride;

    // this CTA's contiguous row range (byte-balanced: rows differ by at most one across CTAs)
    const int64_t r_begin = p.N * (int64_t)blockIdx.x / gridDim.x;
    const int64_t r_end = p.N * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int nrows = (int)(r_end - r_begin);
    const int nrgroups = (nrows + kStageRows - 1) / kStageRows;

    // grouped launch: this CTA's pair selects the expert's matrix, the token's activation row and the output row
    const uint8_t *Wbase = p.W;
    const T *Xrow = static_cast<const T *>(p.X);
    T *Y = static_cast<T *>(p.Y);
    if constexpr (MOE) {
        const int pair = blockIdx.y;
        const int e = moe.ids[pair];
        Y += (int64_t)pair * moe.ldy;
        if (e < 0 || e >= moe.n_experts) {
            // an id outside [0, E) (padding / an expert held by another rank): the pair's output row is zero
            for (int i = tid; i < nrows; i += kThreads) Y[r_begin + i] = from_f32<T>(0.0f);
            return;
        }
        Wbase += (int64_t)e * moe.expert_stride;
        Xrow += (int64_t)(pair / moe.top_k) * K;
    }

    if (tid == 0) {
        for (int s = 0; s < p.nstages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ===================== producer warp: stream (row group, chunk) stages =====================
        const uint64_t pol = l2_policy_evict_first();
        int it = 0;
        for (int rg = 0; rg < nrgroups; rg++) {
            const int nvalid = min(kStageRows, nrows - rg * kStageRows);
            const uint8_t *wrow = Wbase + (r_begin + (int64_t)rg * kStageRows) * p.row_bytes;
            for (int c = 0; c < p.nchunks; c++, it++) {
                const int s = it % p.nstages;
                const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                const int seg_bytes = min(p.chunk_bytes, p.row_bytes - c * p.chunk_bytes);
                uint8_t *dst = stages + s * stage_bytes;
                const uint8_t *src = wrow + (int64_t)c * p.chunk_bytes;
                if (p.nbox) {
                    // rows past N and bytes past the row end are zero-filled and still counted: the stage is always whole
                    if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                    __syncwarp();
                    if (lane < p.nbox)
                        tma_load_2d_hint(dst + lane * (kStageRows * p.box_bytes), &tmapW, (c * p.chunk_bytes + lane * p.box_bytes) >> 2,
                                         (int)(r_begin + (int64_t)rg * kStageRows), &full[s], pol);
                } else if (p.fast) {
                    if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(nvalid * seg_bytes));
                    __syncwarp();
                    if (lane < nvalid)
                        bulk_g2s_hint(dst + lane * p.seg_stride, src + (int64_t)lane * p.row_bytes, (uint32_t)seg_bytes,
                                      &full[s], pol);
                } else {
                    // rows that are not 16-byte aligned (K not a multiple of 256 for some types): 2-byte copies
                    for (int r = 0; r < nvalid; r++) {
                        const uint16_t *s16 = reinterpret_cast<const uint16_t *>(src + (int64_t)r * p.row_bytes);
                        uint16_t *d16 = reinterpret_cast<uint16_t *>(dst + r * p.seg_stride);
                        for (int i = lane; i < seg_bytes / 2; i += 32) d16[i] = s16[i];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                }
            }
        }
        return;
    }

    // ===================== consumers: quantize the activations to q8_1 in shared memory =====================
    {
        const T *X = Xrow;
        const int t4 = lane & 3;
        const int total = B * ngroups_row;                       // groups; 4 lanes each
        for (int base = 0; base < total; base += (kConsumerWarps * 32) / 4) {
            const int gi = base + (tid >> 2);
            const bool live = gi < total;
            const int b = live ? gi / ngroups_row : 0;
            const int g = live ? gi - b * ngroups_row : 0;
            float v[8];
            if (live) {
                load8_f32<T>(X + (int64_t)b * K + g * 32 + t4 * 8, p.x_vec != 0, v);
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = 0.0f;
            }
            const Q8Group q = q8_1_quantize_group<VqTraits<TYPE>::NEED_S8>(v);
            uint32_t aux = 0;
            if constexpr (VqTraits<TYPE>::NEED_SQ) {
                int mine = __dp4a((int)q.q_lo, 0x01010101, __dp4a((int)q.q_hi, 0x01010101, 0));
                mine += __shfl_xor_sync(0xffffffffu, mine, 1);                 // 16-value half sums
                const int other = __shfl_xor_sync(0xffffffffu, mine, 2);
                aux = ((uint32_t)mine & 0xffffu) | ((uint32_t)other << 16);   // valid on t4 == 0
            } else if constexpr (VqTraits<TYPE>::NEED_S8) {
                aux = __float_as_uint(__half2float(__float2half_rn(q.sum)));
            }
            if (live) {
                *reinterpret_cast<uint2 *>(act_q + (int64_t)b * K + g * 32 + t4 * 8) = make_uint2(q.q_lo, q.q_hi);
                if (t4 == 0)
                    act_s[b * ngroups_row + g] = make_float2(__half2float(__float2half_rn(q.d)), __uint_as_float(aux));
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kConsumerWarps * 32) : "memory");
    }

    // ===================== consumers: stream stages, two rows per warp =====================
    int it = 0;
    for (int rg = 0; rg < nrgroups; rg++) {
        float acc[kRowsPerWarp][B];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; r++)
#pragma unroll
            for (int b = 0; b < B; b++) acc[r][b] = 0.0f;

        for (int c = 0; c < p.nchunks; c++, it++) {
            const int s = it % p.nstages;
            const uint32_t ph = (uint32_t)(it / p.nstages) & 1u;
            mbar_wait(&full[s], ph);
            // row pitch inside the stage: the chunk's bytes, or one box's bytes in the tiled layout
            const int pitch = p.nbox ? p.box_bytes : p.seg_stride;
            const uint8_t *seg0 = stages + s * stage_bytes + (warp * kRowsPerWarp) * pitch;
            const int ng = min(kChunkGroups, ngroups_row - c * kChunkGroups);
            for (int g = lane; g < ng; g += 32) {
                GroupW w0, w1;
                int gl = g;
                const uint8_t *sg = seg0;
                if (p.nbox) {
                    const int bx = g / p.gpb;
                    gl = g - bx * p.gpb;
                    sg = seg0 + bx * (kStageRows * p.box_bytes);
                }
                load_group<TYPE>(sg, gl, w0);
                load_group<TYPE>(sg + pitch, gl, w1);
                const int gk = c * kChunkGroups + g;
#pragma unroll
                for (int b = 0; b < B; b++) {
                    const uint4 ua = *reinterpret_cast<const uint4 *>(act_q + (int64_t)b * K + gk * 32);
                    const uint4 ub = *reinterpret_cast<const uint4 *>(act_q + (int64_t)b * K + gk * 32 + 16);
                    const float2 sc = act_s[b * ngroups_row + gk];
                    const uint32_t aux = __float_as_uint(sc.y);
                    int d00 = __dp4a((int)w0.w[0], (int)ua.x, 0), d01 = __dp4a((int)w0.w[4], (int)ub.x, 0);
                    int d10 = __dp4a((int)w1.w[0], (int)ua.x, 0), d11 = __dp4a((int)w1.w[4], (int)ub.x, 0);
                    d00 = __dp4a((int)w0.w[1], (int)ua.y, d00); d01 = __dp4a((int)w0.w[5], (int)ub.y, d01);
                    d10 = __dp4a((int)w1.w[1], (int)ua.y, d10); d11 = __dp4a((int)w1.w[5], (int)ub.y, d11);
                    d00 = __dp4a((int)w0.w[2], (int)ua.z, d00); d01 = __dp4a((int)w0.w[6], (int)ub.z, d01);
                    d10 = __dp4a((int)w1.w[2], (int)ua.z, d10); d11 = __dp4a((int)w1.w[6], (int)ub.z, d11);
                    d00 = __dp4a((int)w0.w[3], (int)ua.w, d00); d01 = __dp4a((int)w0.w[7], (int)ub.w, d01);
                    d10 = __dp4a((int)w1.w[3], (int)ua.w, d10); d11 = __dp4a((int)w1.w[7], (int)ub.w, d11);
                    acc[0][b] += combine<TYPE>(w0, d00, d01, sc.x, aux);
                    acc[1][b] += combine<TYPE>(w1, d10, d11, sc.x, aux);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }

        // warp-shuffle reduction, lane 0 stores (rows past the CTA's range computed on stale bytes: dropped)
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; r++) {
            const int lr = rg * kStageRows + warp * kRowsPerWarp + r;
#pragma unroll
            for (int b = 0; b < B; b++) {
                float v = acc[r][b];
#pragma unroll
                for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
                if (lane == 0 && lr < nrows) Y[(int64_t)b * p.N + r_begin + lr] = from_f32<T>(v);
            }
        }
    }
}

template <int TYPE, int B, typename T>
__global__ void __launch_bounds__(kThreads, 1)
mmvq_kernel(const Params p, const __grid_constant__ CUtensorMap tmapW) {
    mmvq_body<TYPE, B, T, false>(p, tmapW, MoeParams{});
}

// One launch for every (token, expert) pair of a mixture-of-experts layer: grid = (row ranges, pairs), batch 1 per CTA.
template <int TYPE, typename T>
__global__ void __launch_bounds__(kThreads, 1)
moe_mmvq_kernel(const Params p, const __grid_constant__ CUtensorMap tmapW, const MoeParams moe) {
    mmvq_body<TYPE, 1, T, true>(p, tmapW, moe);
}

// ---------------------------------------------------------------- host side
int g_max_box = -1;          // GGB_OPT_MMVQ_TILED: largest TMA box (bytes) of the tiled stages; 0 = per-row bulk copies; -1 = per type

// Measured on one box (14336 x 4096 / 28672 x 8192, batch 1, tiled vs per-row copies): Q4_0 -6 / -17 %, Q4_1 -8 / -24 %,
// Q2_K -9 / -18 % (the producer warp's ~1100 serialised instructions per stage were the bound); Q5_0 +3 / -4 %, Q3_K and
// IQ4_NL +4 % (consumer-bound: nothing to gain, and a 32-row box is fetched row after row), Q8_0 +27 / +19 % (four boxes
// of 544-byte runs); Q5_1 -2 / -12 %.  At batch 8 the consumers are the bound and tiled costs 3 %.  So: tiled where it
// won (batch 1-2), per-row bulk copies elsewhere.
template <int TYPE> constexpr bool kTiledByDefault = (TYPE == T_Q4_0 || TYPE == T_Q4_1 || TYPE == T_Q5_1 || TYPE == T_Q2_K);

// Stage geometry and shared-memory plan of one launch (shared by the plain and the grouped kernel).  Returns GGB_OK, an
// error, or 1 when the B x K activations do not fit beside a 2-stage ring.
template <int TYPE, int B>
int plan_launch(Params &p, CUtensorMap &tmap, int &smem, const void *W, const void *X, void *Y, int64_t K, int64_t N,
                bool allow_tiled, const DeviceInfo &di) {
    constexpr int QK = BlockTraits<TYPE>::QK, BS = BlockTraits<TYPE>::BS;
    p.W = static_cast<const uint8_t *>(W);
    p.X = X;
    p.Y = Y;
    p.N = N;
    p.K = (int)K;
    p.row_bytes = (int)(K / QK * BS);
    p.chunk_bytes = kChunkElems / QK * BS;
    p.nchunks = (int)((K + kChunkElems - 1) / kChunkElems);
    const int seg = (int)(K < kChunkElems ? p.row_bytes : p.chunk_bytes);
    p.seg_stride = (seg + 15) & ~15;
    p.fast = ((reinterpret_cast<uintptr_t>(W) & 15) == 0) && (p.row_bytes % 16 == 0);
    // tiled stage: split the chunk into nbox equal boxes of whole groups (super-blocks for the 256-value types), each a
    // multiple of 16 bytes and at most 1024 (the TMA box limit of 256 uint32 elements); rows shorter than a chunk and
    // block sizes that do not split that way (Q6_K: 210 bytes) keep the per-row bulk copies
    p.nbox = p.box_bytes = p.gpb = 0;
    const int max_box = !allow_tiled ? 0 : g_max_box >= 0 ? g_max_box : ((kTiledByDefault<TYPE> && B <= 2) ? 1024 : 0);
    if (p.fast && K >= kChunkElems && N <= 0x7fffffff) {
        for (int nb = 1; nb <= 8; nb *= 2) {
            const int bb = p.chunk_bytes / nb, gpb = kChunkGroups / nb;
            if (bb <= max_box && bb % 16 == 0 && (gpb * 32) % QK == 0 && bb == gpb * 32 / QK * BS) {
                p.nbox = nb;
                p.box_bytes = bb;
                p.gpb = gpb;
                break;
            }
        }
    }
    memset(&tmap, 0, sizeof tmap);
    if (p.nbox) {
        const int rc = ggb_make_row_tensor_map(&tmap, W, N, p.row_bytes, kStageRows, p.box_bytes, "ggb_mmvq");
        if (rc != GGB_OK) return rc;
    }
    p.x_vec = (reinterpret_cast<uintptr_t>(X) & 15) == 0;
    const int act_q = (int)(B * K);
    p.off_scales = (act_q + 15) & ~15;
    p.off_stages = (p.off_scales + (int)(B * (K / 32) * 8) + 127) & ~127;
    const int stage_bytes = kStageRows * p.seg_stride;
    const int bar_bytes = 2 * kMaxStages * 8;
    const int avail = di.max_smem_optin - p.off_stages - bar_bytes - 128;
    int nst = avail / stage_bytes;
    if (nst < 2) return 1;                       // activations do not fit beside a 2-stage ring: caller splits B
    if (nst > kMaxStages) nst = kMaxStages;
    p.nstages = nst;
    p.off_bars = (p.off_stages + nst * stage_bytes + 127) & ~127;
    smem = p.off_bars + bar_bytes;
    return GGB_OK;
}

template <int TYPE, int B, typename T>
int launch(const void *W, const void *X, void *Y, int64_t K, int64_t N, cudaStream_t stream) {
    const DeviceInfo &di = ggb_device_info();
    Params p;
    alignas(64) CUtensorMap tmap;
    int smem = 0;
    const int prc = plan_launch<TYPE, B>(p, tmap, smem, W, X, Y, K, N, true, di);
    if (prc != GGB_OK) return prc;

    static bool attr_set[64] = {};               // per instantiation and device; idempotent, so a race is harmless
    if (!attr_set[di.dev]) {
        GGB_CUDA_TRY(cudaFuncSetAttribute(mmvq_kernel<TYPE, B, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          di.max_smem_optin));
        attr_set[di.dev] = true;
    }
    const int grid = (int)(N < di.sm_count ? N : di.sm_count);
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmvq_kernel<TYPE, B, T>>(di.dev)));
    mmvq_kernel<TYPE, B, T><<<grid, kThreads, smem, stream>>>(p, tmap);
    GGB_CUDA_TRY(cudaGetLastError());
    return GGB_OK;
}

// Grouped MMVQ for mixture-of-experts decode (no counterpart in the reference -- the op is
// vLLM's ggml_moe_a8_vec): Y[i, :] = q8_1(X[i / top_k, :]) . W[ids[i]]^T for every pair i < pairs, ONE launch.  The
// grid is (row ranges, pairs); a CTA quantizes its pair's activation row and streams its rows of the expert's matrix
// through the same ring as the plain kernel (per-row bulk copies: a tiled tensor map would have to span the experts).
// Pairs that share an expert stream its rows once each (they meet in L2 when they run together).
template <int TYPE, typename T>
int launch_moe(const void *W, const void *X, const int32_t *ids, void *Y, int64_t pairs, int64_t top_k, int64_t n_experts,
               int64_t K, int64_t N, cudaStream_t stream) {
    const DeviceInfo &di = ggb_device_info();
    Params p;
    alignas(64) CUtensorMap tmap;
    int smem = 0;
    const int prc = plan_launch<TYPE, 1>(p, tmap, smem, W, X, Y, K, N, false, di);
    if (prc == 1) {
        ggb_set_error("ggb_moe_mmvq: K=%lld is too large for one activation row in shared memory", (long long)K);
        return GGB_E_SHAPE;
    }
    if (prc != GGB_OK) return prc;
    MoeParams moe;
    moe.expert_stride = N * (int64_t)p.row_bytes;
    moe.ldy = N;
    moe.n_experts = (int)n_experts;
    moe.top_k = (int)top_k;
    p.fast = p.fast && (moe.expert_stride % 16 == 0);

    static bool attr_set[64] = {};
    if (!attr_set[di.dev]) {
        GGB_CUDA_TRY(cudaFuncSetAttribute(moe_mmvq_kernel<TYPE, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          di.max_smem_optin));
        attr_set[di.dev] = true;
    }
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<moe_mmvq_kernel<TYPE, T>>(di.dev)));
    // gridDim.y is limited to 65535: whole tokens per launch
    const int64_t per_launch = (65535 / top_k) * top_k;
    for (int64_t p0 = 0; p0 < pairs; p0 += per_launch) {
        const int64_t np = pairs - p0 < per_launch ? pairs - p0 : per_launch;
        // row ranges per pair: the split that minimises (waves of CTAs) x (rows per CTA + the prologue's worth of rows)
        int gx = 1;
        {
            const int64_t gx_max = N < 4 * di.sm_count ? N : 4 * di.sm_count;
            double best = 1e300;
            for (int64_t g = 1; g <= gx_max; g++) {
                const int64_t waves = (g * np + di.sm_count - 1) / di.sm_count;
                const double cost = (double)waves * ((double)((N + g - 1) / g) + 48.0);
                if (cost < best) {
                    best = cost;
                    gx = (int)g;
                }
            }
        }
        p.X = static_cast<const T *>(X) + (p0 / top_k) * K;
        p.Y = static_cast<T *>(Y) + p0 * N;
        moe.ids = ids + p0;
        moe_mmvq_kernel<TYPE, T><<<dim3(gx, (unsigned)np), kThreads, smem, stream>>>(p, tmap, moe);
        GGB_CUDA_TRY(cudaGetLastError());
    }
    return GGB_OK;
}

template <int TYPE>
int dispatch_moe_dtype(const void *W, const void *X, const int32_t *ids, void *Y, int dt, int64_t pairs, int64_t top_k,
                       int64_t n_experts, int64_t K, int64_t N, cudaStream_t stream) {
    switch (dt) {
        case GGB_F32:  return launch_moe<TYPE, float>(W, X, ids, Y, pairs, top_k, n_experts, K, N, stream);
        case GGB_F16:  return launch_moe<TYPE, __half>(W, X, ids, Y, pairs, top_k, n_experts, K, N, stream);
        case GGB_BF16: return launch_moe<TYPE, __nv_bfloat16>(W, X, ids, Y, pairs, top_k, n_experts, K, N, stream);
        default:
            ggb_set_error("ggb_moe_mmvq: unknown dtype %d", dt);
            return GGB_E_DTYPE;
    }
}

template <int TYPE, typename T>
int dispatch_batch(const void *W, const void *X, void *Y, int64_t B, int64_t K, int64_t N, cudaStream_t stream) {
    int rc;
    switch (B) {
        case 1: rc = launch<TYPE, 1, T>(W, X, Y, K, N, stream); break;
        case 2: rc = launch<TYPE, 2, T>(W, X, Y, K, N, stream); break;
        case 3: rc = launch<TYPE, 3, T>(W, X, Y, K, N, stream); break;
        case 4: rc = launch<TYPE, 4, T>(W, X, Y, K, N, stream); break;
        case 5: rc = launch<TYPE, 5, T>(W, X, Y, K, N, stream); break;
        case 6: rc = launch<TYPE, 6, T>(W, X, Y, K, N, stream); break;
        case 7: rc = launch<TYPE, 7, T>(W, X, Y, K, N, stream); break;
        case 8: rc = launch<TYPE, 8, T>(W, X, Y, K, N, stream); break;
        default:
            ggb_set_error("ggb_mmvq: batch %lld outside 1..8", (long long)B);
            return GGB_E_SHAPE;
    }
    if (rc != 1) return rc;
    if (B == 1) {
        ggb_set_error("ggb_mmvq: K=%lld is too large for one activation row in shared memory", (long long)K);
        return GGB_E_SHAPE;
    }
    // the B x K activation tile does not fit in shared memory: two launches over half the batch each
    const int64_t b1 = B / 2;
    rc = dispatch_batch<TYPE, T>(W, X, Y, b1, K, N, stream);
    if (rc != GGB_OK) return rc;
    return dispatch_batch<TYPE, T>(W, static_cast<const T *>(X) + b1 * K, static_cast<T *>(Y) + b1 * N, B - b1, K, N,
                                   stream);
}

template <int TYPE>
int dispatch_dtype(const void *W, const void *X, void *Y, int dt, int64_t B, int64_t K, int64_t N, cudaStream_t stream) {
    switch (dt) {
        case GGB_F32:  return dispatch_batch<TYPE, float>(W, X, Y, B, K, N, stream);
        case GGB_F16:  return dispatch_batch<TYPE, __half>(W, X, Y, B, K, N, stream);
        case GGB_BF16: return dispatch_batch<TYPE, __nv_bfloat16>(W, X, Y, B, K, N, stream);
        default:
            ggb_set_error("ggb_mmvq: unknown dtype %d", dt);
            return GGB_E_DTYPE;
    }
}

}  // namespace

void ggb_mmvq_tune_tiled(int max_box_bytes) { g_max_box = max_box_bytes < 0 ? -1 : (max_box_bytes > 1024 ? 1024 : max_box_bytes); }

// mmvq_mma.cu: tensor-core integer dots for batches >= 3 (returns 1 when it does not apply)
int ggb_mmvq_mma_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                      void *ws, int64_t ws_bytes, cudaStream_t stream);
int64_t ggb_mmvq_mma_workspace_bytes(int type, int64_t B, int64_t K);

int64_t ggb_mmvq_sb_workspace_bytes(int type, int64_t B, int64_t K);
int64_t ggb_mmvq_workspace_bytes_impl(int type, int64_t B, int64_t K, int64_t) {
    const int64_t a = ggb_mmvq_mma_workspace_bytes(type, B, K), b = ggb_mmvq_sb_workspace_bytes(type, B, K);
    return a > b ? a : b;
}

// mmvq_sb.cu: the super-block-per-lane fast path (returns 1 when it does not apply)
int ggb_mmvq_sb_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                     void *ws, int64_t ws_bytes, cudaStream_t stream);

int ggb_mmvq_rgx_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                      cudaStream_t stream);

int ggb_mmvq_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t B, int64_t K, int64_t N,
                  void *ws, int64_t ws_bytes, cudaStream_t stream) {
    if (K > (1 << 24)) {
        ggb_set_error("ggb_mmvq: K=%lld too large", (long long)K);
        return GGB_E_SHAPE;
    }
    {
        const int rc = ggb_mmvq_mma_impl(W, X, Y, type, xy_dtype, B, K, N, ws, ws_bytes, stream);
        if (rc != 1) return rc;
    }
    {
        const int rc = ggb_mmvq_sb_impl(W, X, Y, type, xy_dtype, B, K, N, ws, ws_bytes, stream);
        if (rc != 1) return rc;
    }
    {
        const int rc = ggb_mmvq_rgx_impl(W, X, Y, type, xy_dtype, B, K, N, stream);      // batch 1, the other types
        if (rc != 1) return rc;
    }
    switch (type) {
        case T_Q4_0: return dispatch_dtype<T_Q4_0>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q4_1: return dispatch_dtype<T_Q4_1>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q5_0: return dispatch_dtype<T_Q5_0>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q5_1: return dispatch_dtype<T_Q5_1>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q8_0: return dispatch_dtype<T_Q8_0>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q2_K: return dispatch_dtype<T_Q2_K>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q3_K: return dispatch_dtype<T_Q3_K>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q4_K: return dispatch_dtype<T_Q4_K>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q5_K: return dispatch_dtype<T_Q5_K>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_Q6_K: return dispatch_dtype<T_Q6_K>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ4_NL: return dispatch_dtype<T_IQ4_NL>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ4_XS: return dispatch_dtype<T_IQ4_XS>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ2_XXS: return dispatch_dtype<T_IQ2_XXS>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ2_XS: return dispatch_dtype<T_IQ2_XS>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ2_S: return dispatch_dtype<T_IQ2_S>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ3_XXS: return dispatch_dtype<T_IQ3_XXS>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ3_S: return dispatch_dtype<T_IQ3_S>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ1_S: return dispatch_dtype<T_IQ1_S>(W, X, Y, xy_dtype, B, K, N, stream);
        case T_IQ1_M: return dispatch_dtype<T_IQ1_M>(W, X, Y, xy_dtype, B, K, N, stream);
        default:
            ggb_set_error("ggb_mmvq: unsupported ggml type %d", type);
            return GGB_E_TYPE;
    }
}

int ggb_moe_mmvq_impl(const void *W, const void *X, const int32_t *ids, void *Y, int type, int xy_dtype, int64_t tokens,
                      int64_t top_k, int64_t n_experts, int64_t K, int64_t N, cudaStream_t stream) {
    if (K > (1 << 24)) {
        ggb_set_error("ggb_moe_mmvq: K=%lld too large", (long long)K);
        return GGB_E_SHAPE;
    }
    const int64_t pairs = tokens * top_k;
#define GGB_MOE_CASE(TT) case TT: return dispatch_moe_dtype<TT>(W, X, ids, Y, xy_dtype, pairs, top_k, n_experts, K, N, stream)
    switch (type) {
        GGB_MOE_CASE(T_Q4_0); GGB_MOE_CASE(T_Q4_1); GGB_MOE_CASE(T_Q5_0); GGB_MOE_CASE(T_Q5_1); GGB_MOE_CASE(T_Q8_0);
        GGB_MOE_CASE(T_Q2_K); GGB_MOE_CASE(T_Q3_K); GGB_MOE_CASE(T_Q4_K); GGB_MOE_CASE(T_Q5_K); GGB_MOE_CASE(T_Q6_K);
        GGB_MOE_CASE(T_IQ4_NL); GGB_MOE_CASE(T_IQ4_XS); GGB_MOE_CASE(T_IQ2_XXS); GGB_MOE_CASE(T_IQ2_XS);
        GGB_MOE_CASE(T_IQ2_S); GGB_MOE_CASE(T_IQ3_XXS); GGB_MOE_CASE(T_IQ3_S); GGB_MOE_CASE(T_IQ1_S); GGB_MOE_CASE(T_IQ1_M);
        default:
            ggb_set_error("ggb_moe_mmvq: unsupported ggml type %d", type);
            return GGB_E_TYPE;
    }
#undef GGB_MOE_CASE
}
