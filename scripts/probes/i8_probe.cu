// i8_probe.cu -- measurement probe (not part of the library): what can tcgen05.mma.kind::i8 deliver for MMQ with EXACT q8_1
// semantics on a B200?  (SURVEY App. D levers 1-2)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o scripts/probes/i8_probe scripts/probes/i8_probe.cu
//   scripts/probes/i8_probe            (on the GPU box; prints one JSON object)
//
// Four measurements, one persistent CTA per SM each, operands resident in shared memory (no loads, no unpacking: every
// number is an UPPER bound for a real kernel of that structure):
//   check      one 128 x 256 x 128 int8 tile (4 MMAs of K = 32) against a CPU int32 reference -> the kind::i8 instruction and
//              SWIZZLE_128B descriptors are right;
//   peak_i8    back-to-back tcgen05.mma.kind::i8 M=128 N=256 K=32, accumulating in TMEM           -> the int8 tensor peak
//   peak_f16   the same loop with kind::f16 (K = 16)                                              -> the fp16 tensor peak
//   drain      the exact-q8_1 structure: ggml carries a float scale per 32 K-values on BOTH operands, so every K = 32 MMA
//              group needs its own s32 accumulator, drained from TMEM and rescaled on CUDA cores:
//                  acc[i][j] += float(C_g[i][j]) * sA[i][g] * sB[j][g]
//              1 MMA-issuing thread + 8 drain warps (two per TMEM lane quadrant, 128 columns each, 128 fp32 accumulators per
//              thread), the two 256-column accumulators of TMEM double-buffered so the drain of group g overlaps the MMA of
//              group g + 1.  `ops` = 3 (cvt, fmul, ffma) is the reference's own arithmetic (mmq.cuh:1351-1361); `ops` = 2 folds
//              sA * sB into one pre-multiplied scale (what two-plane sc*q folding would buy; its doubled MMA count is NOT
//              charged here, so it is optimistic).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {      // K-major SWIZZLE_128B, 8-row atoms 1024 B apart
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND>     // 0 = i8, 1 = f16
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int BM = 128, BN = 256;
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // s8 x s8 -> s32
constexpr uint32_t IDESC_F16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);                             // f16 x f16 -> f32

// byte offset of element (row r, byte column c < 128) in a SWIZZLE_128B K-major tile
__host__ __device__ inline int sw128(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 4) ^ (r & 7))) << 4) + (c & 15); }

struct Smem {
    uint8_t A[BM * 128];     // 16 KB
    uint8_t B[BN * 128];     // 32 KB
    float sA[8][BM];         // per-group weight-side scales  (d * sc)
    float sB[8][BN];         // per-group activation-side scales (d8)
    uint64_t bar[8];
    uint32_t tmem;
};

__device__ void common_setup(Smem *s, const uint8_t *gA, const uint8_t *gB, int warp, int tid, int nthreads) {
    for (int i = tid; i < BM * 128 / 16; i += nthreads) reinterpret_cast<uint4 *>(s->A)[i] = reinterpret_cast<const uint4 *>(gA)[i];
    for (int i = tid; i < BN * 128 / 16; i += nthreads) reinterpret_cast<uint4 *>(s->B)[i] = reinterpret_cast<const uint4 *>(gB)[i];
    for (int i = tid; i < 8 * BM; i += nthreads) (&s->sA[0][0])[i] = 1.0f + (float)(i % 63) * 0.015625f;
    for (int i = tid; i < 8 * BN; i += nthreads) (&s->sB[0][0])[i] = 0.5f + (float)(i % 17) * 0.03125f;
    if (tid == 0) {
        for (int i = 0; i < 8; i++) mbar_init(&s->bar[i], i >= 2 && i < 4 ? 8 : 1);     // bar[2..3] = accumulator-empty (8 drain warps)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s->tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}
__device__ void common_teardown(Smem *s, int warp) {
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s->tmem) : "memory");
    }
}

// ---------------------------------------------------------------- check: C[128][256] = A . B^T over K = 128
__global__ void __launch_bounds__(128) check_kernel(const uint8_t *gA, const uint8_t *gB, int32_t *C) {
    extern __shared__ __align__(1024) uint8_t raw[];
    Smem *s = reinterpret_cast<Smem *>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    common_setup(s, gA, gB, warp, tid, 128);
    const uint32_t tmem = s->tmem;
    if (tid == 0) {
        for (int kk = 0; kk < 4; kk++)
            umma<0>(tmem, desc_sw128(smem_u32(s->A) + kk * 32), desc_sw128(smem_u32(s->B) + kk * 32), IDESC_I8, kk > 0);
        umma_commit(&s->bar[0]);
    }
    mbar_wait(&s->bar[0], 0);
    tc_fence_after();
    for (int ch = 0; ch < BN / 32; ch++) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++) C[(warp * 32 + lane) * BN + ch * 32 + j] = (int32_t)v[j];
    }
    common_teardown(s, warp);
}

// ---------------------------------------------------------------- peak: back-to-back MMAs from one thread
template <int KIND>
__global__ void __launch_bounds__(128) peak_kernel(const uint8_t *gA, const uint8_t *gB, int iters, int32_t *sink) {
    extern __shared__ __align__(1024) uint8_t raw[];
    Smem *s = reinterpret_cast<Smem *>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5;
    common_setup(s, gA, gB, warp, tid, 128);
    const uint32_t tmem = s->tmem;
    if (tid == 0) {
        const uint32_t a = smem_u32(s->A), b = smem_u32(s->B);
        for (int it = 0; it < iters; it++) {
            // 4 MMAs walk the 128-byte K extent of the stage; two accumulators alternate like a real tile loop
            const uint32_t d = tmem + (uint32_t)((it & 1) * BN);
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
                umma<KIND>(d, desc_sw128(a + kk * 32), desc_sw128(b + kk * 32), KIND == 0 ? IDESC_I8 : IDESC_F16, (it > 1 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&s->bar[0]);
    }
    mbar_wait(&s->bar[0], 0);
    tc_fence_after();
    if (warp == 0 && sink) {
        uint32_t v[32];
        tmem_ld32(tmem, v);
        tmem_ld_wait();
        if (v[0] == 0x12345678u) sink[0] = (int32_t)v[1];
    }
    common_teardown(s, warp);
}

// ---------------------------------------------------------------- drain: per-32 groups, double-buffered accumulators
// warp 0: MMA issuer; warps 1..8: drain (quadrant = warp & 3, column half = (warp - 1) >> 2)
template <int OPS>
__global__ void __launch_bounds__(288, 1) drain_kernel(const uint8_t *gA, const uint8_t *gB, int groups, int tiles, float *out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    Smem *s = reinterpret_cast<Smem *>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    common_setup(s, gA, gB, warp, tid, 288);
    const uint32_t tmem = s->tmem;
    uint64_t *full = &s->bar[0], *empty = &s->bar[2];
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t a = smem_u32(s->A), b = smem_u32(s->B);
            int it = 0;
            for (int t = 0; t < tiles; t++)
                for (int g = 0; g < groups; g++, it++) {
                    const int buf = it & 1;
                    mbar_wait(&empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    umma<0>(tmem + (uint32_t)(buf * BN), desc_sw128(a + (g & 3) * 32), desc_sw128(b + (g & 3) * 32), IDESC_I8, 0u);
                    umma_commit(&full[buf]);
                }
        }
        __syncwarp();
    } else {
        const int quad = warp & 3, half = (warp - 1) >> 2;
        const int row = quad * 32 + lane;
        int it = 0;
        for (int t = 0; t < tiles; t++) {
            float acc[128];
#pragma unroll
            for (int j = 0; j < 128; j++) acc[j] = 0.0f;
            for (int g = 0; g < groups; g++, it++) {
                const int buf = it & 1;
                const float sa = s->sA[g & 7][row];
                const float4 *sb4 = reinterpret_cast<const float4 *>(&s->sB[g & 7][half * 128]);
                mbar_wait(&full[buf], (uint32_t)(it >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + half * 128);
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    uint32_t v[32];
                    tmem_ld32(taddr + ch * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 sb = sb4[ch * 8 + q];
                        const float sbv[4] = {sb.x, sb.y, sb.z, sb.w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int j = ch * 32 + q * 4 + e;
                            const float c = (float)(int32_t)v[q * 4 + e];
                            if (OPS == 3) acc[j] = fmaf(c * sa, sbv[e], acc[j]);
                            else acc[j] = fmaf(c, sbv[e], acc[j]);          // pre-multiplied scale: cvt + ffma
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[buf]);
            }
            // tile epilogue: keep the accumulators alive (one store per 128 values; a real kernel writes them all)
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 128; j++) sum += acc[j];
            out[((size_t)blockIdx.x * tiles + t) * 256 + (warp - 1) * 32 + lane] = sum;
        }
    }
    common_teardown(s, warp);
}

int main() {
    int dev = 0, sms = 0, clk = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    // operands: small signed values, laid out K-major with the 128-byte swizzle
    std::vector<int8_t> A(BM * 128), B(BN * 128);
    std::vector<uint8_t> As(BM * 128), Bs(BN * 128);
    srand(1);
    for (auto &v : A) v = (int8_t)(rand() % 31 - 15);
    for (auto &v : B) v = (int8_t)(rand() % 255 - 127);
    for (int r = 0; r < BM; r++) for (int c = 0; c < 128; c++) As[sw128(r, c)] = (uint8_t)A[r * 128 + c];
    for (int r = 0; r < BN; r++) for (int c = 0; c < 128; c++) Bs[sw128(r, c)] = (uint8_t)B[r * 128 + c];
    uint8_t *dA, *dB;
    int32_t *dC;
    float *dOut;
    CK(cudaMalloc(&dA, As.size()));
    CK(cudaMalloc(&dB, Bs.size()));
    CK(cudaMalloc(&dC, BM * BN * 4));
    CK(cudaMalloc(&dOut, (size_t)sms * 64 * 256 * 4));
    CK(cudaMemcpy(dA, As.data(), As.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bs.data(), Bs.size(), cudaMemcpyHostToDevice));
    const int smem = (int)sizeof(Smem) + 1024;
    CK(cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(peak_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(peak_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(drain_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(drain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));

    // ---- check
    check_kernel<<<1, 128, smem>>>(dA, dB, dC);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> C(BM * BN);
    CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int i = 0; i < BM; i++)
        for (int j = 0; j < BN; j++) {
            int32_t ref = 0;
            for (int k = 0; k < 128; k++) ref += (int32_t)A[i * 128 + k] * (int32_t)B[j * 128 + k];
            if (ref != C[i * BN + j]) bad++;
        }

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto time_ms = [&](auto launch) {
        launch();
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        return best;
    };
    // ---- peaks: iters x 4 MMAs of 128 x 256 x (K = 32 int8 | 16 fp16) per CTA
    const int iters = 20000;
    const float ms_i8 = time_ms([&] { peak_kernel<0><<<sms, 128, smem>>>(dA, dB, iters, nullptr); });
    const float ms_f16 = time_ms([&] { peak_kernel<1><<<sms, 128, smem>>>(dA, dB, iters, nullptr); });
    const double ops_i8 = 2.0 * BM * BN * 32 * 4.0 * iters * sms, ops_f16 = 2.0 * BM * BN * 16 * 4.0 * iters * sms;
    // ---- drain
    const int groups = 128, tiles = 16;                    // K = 4096 per tile
    const float ms_d3 = time_ms([&] { drain_kernel<3><<<sms, 288, smem>>>(dA, dB, groups, tiles, dOut); });
    const float ms_d2 = time_ms([&] { drain_kernel<2><<<sms, 288, smem>>>(dA, dB, groups, tiles, dOut); });
    const double ops_d = 2.0 * BM * BN * 32 * (double)groups * tiles * sms;
    const double cyc_d3 = ms_d3 * 1e-3 * clk * 1e3 / ((double)groups * tiles), cyc_d2 = ms_d2 * 1e-3 * clk * 1e3 / ((double)groups * tiles);
    printf("{\"sms\": %d, \"sm_clock_khz\": %d, \"check_bad\": %ld, \"check_elems\": %d,\n"
           " \"peak_i8_tops\": %.1f, \"peak_i8_ms\": %.4f, \"peak_f16_tflops\": %.1f, \"peak_f16_ms\": %.4f,\n"
           " \"drain_exact_q8_1\": {\"tile\": \"128x256, K=32 groups, 8 drain warps, double-buffered TMEM\", \"groups_per_tile\": %d, \"tiles_per_cta\": %d,\n"
           "   \"ops3_cvt_fmul_ffma\": {\"ms\": %.4f, \"tops\": %.1f, \"cycles_per_group_at_max_clock\": %.0f},\n"
           "   \"ops2_cvt_ffma\": {\"ms\": %.4f, \"tops\": %.1f, \"cycles_per_group_at_max_clock\": %.0f}}}\n",
           sms, clk, bad, BM * BN, ops_i8 / (ms_i8 * 1e-3) / 1e12, ms_i8, ops_f16 / (ms_f16 * 1e-3) / 1e12, ms_f16, groups, tiles,
           ms_d3, ops_d / (ms_d3 * 1e-3) / 1e12, cyc_d3, ms_d2, ops_d / (ms_d2 * 1e-3) / 1e12, cyc_d2);
    return bad ? 2 : 0;
}
