// mmq.cu -- quantized GEMM for prefill (MMQ) on sm_100a tensor cores: tcgen05.mma + TMEM + TMA bulk copies.
//
// Replaces ggml_mul_mat_a8 = quantize_mmq_q8_1_cuda + mul_mat_q (hf-kernels/ggml-kernels/ggml/mmq.cu:109-255,
// ggml/mmq.cuh:1917-2031, tile picker kernel_instances/mmq_kernel.cuh:11-87).  The reference unpacks
// weights to int32 lanes in shared memory and issues Ampere mma.sync s8 tiles of at most 64 x 128 with a
// float rescale after every K=32 step on CUDA cores (mmq.cuh:1351-1361); Q2_K/Q3_K never touch tensor cores.
//
// Design (DESIGN.md "MMQ"): ggml's formats carry a float scale per 32 K-values on BOTH operands, so an
// integer accumulator has to be drained and rescaled on CUDA cores every 32-deep step -- on tcgen05 that
// drain costs ~8x the MMA itself.  Instead both operands are brought to fp16 with their scales applied
// and the whole K extent accumulates in fp32 inside TMEM:
//     A[i,k] = fp16( dequantize(W)[i,k] )        -- bit-identical to ggml_dequantize(..., fp16)
//     B[j,k] = fp16( d8[j,k/32] * q8[j,k] )      -- the q8_1 quantization of X (bit-identical d8/q8)
//     Y[j,i] = sum_k A[i,k] * B[j,k]             -- tcgen05.mma.kind::f16, fp32 accumulate in TMEM
// Launches: (1) quantize X to q8_1 and write the fp16 B operand as a pre-swizzled tile image into the workspace (one CTA
// per token; a token of out-of-range magnitude is stored divided by a power of two, undone in the epilogue); (2) the
// persistent warp-specialised GEMM; (3) for some (M, N) the last partial round of tiles as a second GEMM launch of
// 128 x 64 tiles.  The GEMM for M > 256 runs as CTA PAIRS (cluster of 2): the two CTAs own two adjacent 128-row blocks and
// the leader issues ONE tcgen05.mma.cta_group::2 (M = 256 weight rows, N = 256 tokens, K = 16) per step for both; each
// CTA supplies its own unpacked A tile and half of the B tile.  Per CTA (14 warps):
//     warp 0      producer: 2-D TMA of raw GGUF row segments (128 rows x 256 columns per raw stage) and bulk copies of
//                 this CTA's part of each pre-swizzled B tile, mbarrier tx-count
//     warps 6-13  unpack: raw bytes -> fp16 A tile in the 128-byte-swizzle K-major layout UMMA reads (half2 arithmetic)
//     warp 1      one elected thread issues the MMAs and tcgen05.commit's the stage / the accumulator (pair: the peer's
//                 MMA warp relays "my A and my half of B are in place" with a remote mbarrier arrive)
//     warps 2-5   epilogue: tcgen05.ld the fp32 accumulator, convert, transpose through an 8 KB staging buffer, store
//                 256-byte token rows of Y in 16-byte vectors
// Roofline: tensor pipe (dense fp16 -> fp32 accumulate); ops = 2*M*N*K.
#include <cuda.h>          // CUtensorMap (types only: the encode entry point is fetched at run time)

#include "q8_1.cuh"
#include "mmq_unpack.cuh"

bool ggb_mmq_need_sum(int type);     // quantize_q8_1.cu

namespace {

constexpr int BM = 128;              // weight rows per tile (UMMA M)
constexpr int KSTEP = 64;            // K elements per A/B stage: 128 B of fp16 = one SWIZZLE_128B row
constexpr int RAWK = 256;            // K elements per raw-weight stage
constexpr int NRAW = 2;
// NB = token tiles (of BN = 256 tokens) that share one unpacked A stage: 1 -> three A/B stages and a double-buffered
// accumulator (the epilogue overlaps the next tile); 2 -> every unpacked weight feeds 512 tokens (half the unpack work
// per MMA), two A/B stages, both 256-column accumulators live in TMEM at once
__host__ __device__ constexpr int nstage_of(int nb) { return nb == 1 ? 3 : 2; }
constexpr int kUnpackWarps = 8;
constexpr int kFirstEpiWarp = 2, kFirstUnpackWarp = 6;
constexpr int kThreads = (kFirstUnpackWarp + kUnpackWarps) * 32;     // 448
constexpr int A_STAGE_BYTES = BM * 128;

struct Params {
    const uint8_t *W;
    const uint8_t *Ximg;     // pre-swizzled fp16 B operand image (workspace)
    const float *row_scale;  // per token: the power of two its image row was divided by (1 for ordinary magnitudes)
    void *Y;
    int64_t N, M, ldy;
    int K;
    int row_bytes;           // bytes per weight row
    int nsb;                 // raw stages per tile = ceil(K / 256)
    int Mp;                  // tokens padded to a multiple of BN
    int n_row_blocks, n_tok_tiles, n_tok_groups;
    int raw_slot;            // bytes per row in a raw stage
    int w_fast;              // W 16-byte aligned and total bytes % 16 == 0: bulk copies
    int w_lo;                // low 4 bits of W (0 in the fast path)
    int img_plane;           // token atoms (8 tokens) per K-plane of the B image (the full activation matrix's)
    int ntiles_run;          // cluster-level tiles this launch processes (the rest belongs to a tail launch)
    int off_b, off_raw, off_tab, off_epi, off_bar;
    int pdl;                 // 0: the launch has no programmatic dependency to wait for (tail launch: its inputs were
                             //    complete before the main launch released it); 1: wait before B loads and Y stores;
                             //    2: as 1, and weights are static (raw weight tiles may be fetched before the wait)
    int y_vec;               // Y rows can take 16-byte stores (Y 16-byte aligned, ldy * sizeof(T) % 16 == 0)
};

// ---------------------------------------------------------------- tcgen05 / descriptor helpers
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row atoms 1024 B apart (SBO), LBO unused (=1),
    // descriptor version 1 (sm_100), layout_type 2 = SWIZZLE_128B.   (cute/arch/mma_sm100_desc.hpp)
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// multicast variants for a thread-block cluster: the arrive / the copied bytes + complete_tx land at the same
// CTA-relative offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// cta_group::2 (CTA pair) variants: ONE instruction issued by the leader CTA drives the tensor cores of both SMs on a
// 256-row tile; each CTA supplies its own 128 rows of A and HALF of B, so the shared-memory traffic per MMA drops from
// 12 KB to 8 KB per SM and the B stage from 64 KB to 32 KB
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {      // sees arrivals from the peer CTA
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
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

__device__ __forceinline__ void bulk_g2s_plain(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 2-D tiled TMA load (box = 128 rows x BOX_INNER bytes of raw GGUF data), completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- launch 1: q8_1 quantize -> fp16 B image
// Image layout: [K/64 stages][Mp/8 token atoms][8 rows x 128 B, 16-byte chunks XOR-swizzled by row&7] --
// exactly the bytes a SWIZZLE_128B TMA load would put in shared memory, so a BN x 64 B tile is ONE
// contiguous bulk copy.  d8 is rounded to fp16 for the need_sum weight types (it is stored in a half2
// there, mmq.cu:149-153) and kept in fp32 otherwise.
// Range: B = fp16(d8 * q8) reaches the token's amax, so a token whose largest activation is outside fp16's comfortable
// range (the reference keeps int8 quants and fp32 scales: it takes |x| up to ~65504 * 127 and arbitrarily small rows) is
// stored divided by a power of two 2^e (exact) that brings its amax to [2^8, 2^9); the GEMM's epilogue multiplies the
// token's outputs by 2^e again (row_scale).  Ordinary rows (2^-8 <= amax <= 2^15) have e = 0 and take no extra work.
template <typename T, bool D_HALF>
__global__ void __launch_bounds__(256)
quantize_b_image_kernel(const T *__restrict__ x, uint8_t *__restrict__ img, float *__restrict__ row_scale, int64_t M, int K, int Kp,
                        int Mp, bool vec_ok) {
    ggb_pdl_launch_dependents();        // the GEMM's CTAs may become resident (barriers, TMEM, raw weight tiles) meanwhile
    ggb_pdl_wait();                     // x comes from, and the image may still be read by, the previous kernels
    __shared__ float s_amax[8];
    const int64_t row = blockIdx.x;     // one CTA per token
    const int groups_per_row = Kp / 32;
    const int t = threadIdx.x & 3;
    const T *xr = x + row * K;
    auto load_group = [&](int g, float (&v)[8]) {
        const int k0 = g * 32 + t * 8;
        if (k0 + 8 <= K) {
            load8_f32<T>(xr + k0, vec_ok, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = (k0 + i < K) ? to_f32<T>(xr[k0 + i]) : 0.0f;
        }
    };
    // pass 1: the token's amax
    float amax = 0.0f;
    for (int g = threadIdx.x >> 2; g < groups_per_row; g += 64) {
        float v[8];
        load_group(g, v);
#pragma unroll
        for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, m));
    if ((threadIdx.x & 31) == 0) s_amax[threadIdx.x >> 5] = amax;
    __syncthreads();
    amax = s_amax[0];
#pragma unroll
    for (int w = 1; w < 8; w++) amax = fmaxf(amax, s_amax[w]);
    int e = 0;
    if (amax > 32768.0f || (amax < 0.00390625f && amax > 0.0f)) e = ilogbf(amax) - 8;      // NaN / inf rows: e = 0, they propagate
    if (!(amax <= 3.0e38f)) e = 0;
    const float inv = ldexpf(1.0f, -e);          // exact: |e| <= 149 + 8
    if (threadIdx.x == 0) row_scale[row] = ldexpf(1.0f, e);
    // pass 2 (the row is L1 / L2 resident): q8_1 groups, bit-identical to quantize_mmq_q8_1, folded into fp16
    const int64_t atom = row >> 3;
    const int r8 = (int)(row & 7);
    for (int g0 = 0; g0 < groups_per_row; g0 += 64) {
        const int g = g0 + (threadIdx.x >> 2);
        const bool live = g < groups_per_row;
        float v[8];
        if (live) {
            load_group(g, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = 0.0f;
        }
        const Q8Group r = q8_1_quantize_group<false>(v);
        if (!live) continue;
        const float d = (D_HALF ? __half2float(__float2half_rn(r.d)) : r.d);
        float y[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            y[i] = __fmul_rn(__fmul_rn(d, (float)(int)(int8_t)(r.q_lo >> (8 * i))), inv);
            y[i + 4] = __fmul_rn(__fmul_rn(d, (float)(int)(int8_t)(r.q_hi >> (8 * i))), inv);
        }
        const int k0 = g * 32 + t * 8;
        const int ks = k0 >> 6, c = (k0 & 63) >> 3;
        uint8_t *dst = img + ((int64_t)ks * (Mp >> 3) + atom) * 1024 + r8 * 128 + ((c ^ r8) << 4);
        store8(reinterpret_cast<__half *>(dst), y);
    }
}

// ---------------------------------------------------------------- launch 2: the GEMM
// CL = thread-block cluster size (1 or 2).  The CTAs of a cluster work on CL adjacent 128-row blocks of the SAME
// token group in lock step: each CTA fetches 1/CL of every B stage and multicasts it to all of them, so the
// L2 -> SM traffic of the B operand (the limiter at 128 rows per tile) drops by CL.
// PAIR (CL == 2 only) = the two CTAs of the cluster issue ONE cta_group::2 MMA per step (see umma_f16_pair) instead of
// two independent ones with a multicast B stage.
template <int TYPE, int BN, int NB, int CL, typename T, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1)
mmq_kernel(const Params p, const __grid_constant__ CUtensorMap tmapW) {
    static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
    constexpr int NSTAGE = PAIR ? 3 : nstage_of(NB);
    constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1u);
    constexpr int NACC = (NB == 1) ? 2 : 1;                   // accumulator sets in TMEM
    constexpr int QK = BlockTraits<TYPE>::QK, BS = BlockTraits<TYPE>::BS;
    constexpr int SB_BYTES = RAWK / QK * BS;                  // raw bytes per row per raw stage
    // TMA box width in bytes (also the raw row slot).  The box must START on a 16-byte boundary, so for block sizes
    // that are not multiples of 16 (Q2_K 84, Q3_K 110, Q6_K 210) it starts up to 14 bytes early.
    constexpr int BOX_INNER = (SB_BYTES + ((SB_BYTES % 16) ? 14 : 0) + 15) & ~15;
    constexpr int B_TILE_BYTES = BN * 128;                    // one token tile of a stage in the image
    constexpr int B_TILE_LOCAL = PAIR ? B_TILE_BYTES / 2 : B_TILE_BYTES;   // ... and the part of it this CTA holds
    constexpr int B_STAGE_BYTES = NB * B_TILE_LOCAL;
    constexpr int ACC_COLS = (BN < 32) ? 32 : BN;             // TMEM columns per accumulator buffer
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);   // f16 x f16 -> f32, K-major

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned tiles; the launch reserves the slack
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sA = smem;
    uint8_t *sB = smem + p.off_b;
    uint8_t *sRaw = smem + p.off_raw;
    uint32_t *sTab = reinterpret_cast<uint32_t *>(smem + p.off_tab);     // half2 (d*sc, -dmin*m) per sub-block, per raw stage
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *raw_full = bars, *raw_empty = bars + NRAW;
    uint64_t *a_full = bars + 2 * NRAW, *b_full = a_full + NSTAGE, *ab_empty = b_full + NSTAGE;
    uint64_t *acc_full = ab_empty + NSTAGE, *acc_empty = acc_full + 2;
    uint64_t *peer_full = acc_empty + 2;                       // PAIR, leader only: the peer CTA's A and B of a stage are in place
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(peer_full + NSTAGE);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_rbg = (p.n_row_blocks + CL - 1) / CL;          // row-block groups (CL row blocks each; a phantom one is all zeros)
    const int ntiles = min(n_rbg * p.n_tok_groups, p.ntiles_run);     // cluster-level tiles
    const int ksteps_total = p.nsb * (RAWK / KSTEP);
    const int crank = (CL == 1) ? 0 : (int)(blockIdx.x % CL), cid = blockIdx.x / CL, ncl = gridDim.x / CL;

    if (tid == 0) {
        for (int s = 0; s < NRAW; s++) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kUnpackWarps); }
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(&a_full[s], kUnpackWarps);
            mbar_init(&b_full[s], 1);
            mbar_init(&ab_empty[s], PAIR ? 1 : CL);            // PAIR: one multicast commit by the leader
            mbar_init(&peer_full[s], 1);
        }
        for (int s = 0; s < 2; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], PAIR ? 8 : 4); }   // NB == 2 uses slot 0 only
        fence_mbar_init();
    }
    if (warp == 1) {   // TMEM: all 512 columns (two accumulator buffers); one CTA per SM so no contention
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (CL == 1) __syncthreads();
    else cluster_sync_all();                                   // peers' barriers are initialised before anything targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================ producer ============================
        // Programmatic dependent launch: this CTA may be resident while the activation-image kernel still runs.
        // Dependents are released only AFTER this kernel's own dependency has resolved, so a dependent that needs
        // nothing newer than this kernel's inputs (the tail launch) may skip its wait and fill the SMs that the last
        // round of this launch leaves idle.
        bool waited = (p.pdl == 0);
        if (p.pdl == 1) { ggb_pdl_wait(); waited = true; if (lane == 0) ggb_pdl_launch_dependents(); }
        int it = 0, rit = 0;
        for (int tile = cid; tile < ntiles; tile += ncl) {
            const int rb = (tile % n_rbg) * CL + crank, tg = tile / n_rbg;
            const int tt = tg * NB, nb_valid = min(NB, p.n_tok_tiles - tt);
            const int64_t row0 = (int64_t)rb * BM;
            for (int sb = 0; sb < p.nsb; sb++, rit++) {
                const int rs = rit % NRAW;
                mbar_wait(&raw_empty[rs], ((uint32_t)(rit / NRAW) & 1u) ^ 1u);
                uint8_t *dst = sRaw + rs * (BM * p.raw_slot);
                const int seg_len = min(SB_BYTES, p.row_bytes - sb * SB_BYTES);
                if (p.w_fast) {
                    // ONE tiled TMA load: 128 rows x BOX_INNER bytes starting at this raw stage's column; rows past
                    // N and bytes past the row end are zero-filled by the TMA unit.
                    if (lane == 0) {
                        mbar_expect_tx(&raw_full[rs], (uint32_t)(BM * BOX_INNER));
                        tma_load_2d(dst, &tmapW, ((sb * SB_BYTES) & ~15) / 2, (int)row0, &raw_full[rs]);
                    }
                    __syncwarp();
                } else {
                    for (int r = 0; r < BM; r++) {
                        if (row0 + r >= p.N) break;
                        const uint16_t *s16 = reinterpret_cast<const uint16_t *>(p.W + (row0 + r) * p.row_bytes + (int64_t)sb * SB_BYTES);
                        uint16_t *d16 = reinterpret_cast<uint16_t *>(dst + r * p.raw_slot);
                        for (int i = lane; i < seg_len / 2; i += 32) d16[i] = s16[i];
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&raw_full[rs]);
                }
                if (!waited) { ggb_pdl_wait(); waited = true; if (lane == 0) ggb_pdl_launch_dependents(); }
                for (int q = 0; q < RAWK / KSTEP; q++, it++) {
                    const int s = it % NSTAGE;
                    mbar_wait(&ab_empty[s], ((uint32_t)(it / NSTAGE) & 1u) ^ 1u);
                    if (lane == 0) {
                        const int ks = sb * (RAWK / KSTEP) + q;
                        // the nb_valid token tiles of a group are adjacent in the image: one contiguous copy
                        const uint8_t *src = p.Ximg + ((int64_t)ks * p.img_plane + (int64_t)tt * (BN >> 3)) * 1024;
                        if (PAIR) {
                            // this CTA's half (tokens crank * BN/2 ..) of each token tile of the stage
                            mbar_expect_tx(&b_full[s], (uint32_t)(nb_valid * B_TILE_LOCAL));
                            for (int tb = 0; tb < nb_valid; tb++)
                                bulk_g2s_plain(sB + s * B_STAGE_BYTES + tb * B_TILE_LOCAL, src + tb * B_TILE_BYTES + crank * B_TILE_LOCAL,
                                               (uint32_t)B_TILE_LOCAL, &b_full[s]);
                        } else if (CL == 1) {
                            mbar_expect_tx(&b_full[s], (uint32_t)(nb_valid * B_TILE_BYTES));
                            bulk_g2s_plain(sB + s * B_STAGE_BYTES, src, (uint32_t)(nb_valid * B_TILE_BYTES), &b_full[s]);
                        } else {
                            mbar_expect_tx(&b_full[s], (uint32_t)(nb_valid * B_TILE_BYTES));
                            const uint32_t part = (uint32_t)(nb_valid * B_TILE_BYTES) / CL;
                            bulk_g2s_mc(sB + s * B_STAGE_BYTES + crank * part, src + crank * part, part, &b_full[s], CL_MASK);
                        }
                    }
                    __syncwarp();
                }
            }
        }
        if (!waited) { ggb_pdl_wait(); if (lane == 0) ggb_pdl_launch_dependents(); }
        else if (p.pdl == 0 && lane == 0) ggb_pdl_launch_dependents();
    } else if (warp == 1) {
        // ============================ MMA issuer (one thread) ============================
        if (lane == 0 && PAIR && crank != 0) {
            // peer CTA of a pair: no MMA to issue; relay "my A tile and my half of B for stage s are in place" to the leader
            int it = 0;
            for (int tile = cid; tile < ntiles; tile += ncl) {
                for (int ks = 0; ks < ksteps_total; ks++, it++) {
                    const int s = it % NSTAGE;
                    const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
                    mbar_wait(&a_full[s], ph);
                    mbar_wait(&b_full[s], ph);
                    mbar_arrive_remote(&peer_full[s], 0);
                }
            }
        } else if (lane == 0) {
            int it = 0, tcount = 0;
            for (int tile = cid; tile < ntiles; tile += ncl, tcount++) {
                const int ab = tcount % NACC;
                const int nb_valid = min(NB, p.n_tok_tiles - (tile / n_rbg) * NB);
                if (PAIR) mbar_wait_cluster(&acc_empty[ab], ((uint32_t)(tcount / NACC) & 1u) ^ 1u);   // both CTAs' epilogues
                else mbar_wait(&acc_empty[ab], ((uint32_t)(tcount / NACC) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(ab * ACC_COLS);
                for (int ks = 0; ks < ksteps_total; ks++, it++) {
                    const int s = it % NSTAGE;
                    const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
                    mbar_wait(&a_full[s], ph);
                    mbar_wait(&b_full[s], ph);
                    if (PAIR) mbar_wait_cluster(&peer_full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + s * A_STAGE_BYTES);
                    const uint32_t b_addr = smem_u32(sB + s * B_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < KSTEP / 16; kk++) {
                        const uint32_t acc = (ks > 0 || kk > 0) ? 1u : 0u;
                        if (PAIR) {
                            // 256 weight rows (128 per CTA) x BN tokens (BN / 2 rows of B from each CTA's shared memory)
                            umma_f16_pair(d_tmem, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), IDESC, acc);
                            if (NB == 2 && nb_valid == 2)
                                umma_f16_pair(d_tmem + ACC_COLS, umma_desc_sw128(a_addr + kk * 32),
                                              umma_desc_sw128(b_addr + B_TILE_LOCAL + kk * 32), IDESC, acc);
                        } else {
                            umma_f16(d_tmem, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), IDESC, acc);
                            if (NB == 2 && nb_valid == 2)
                                umma_f16(d_tmem + ACC_COLS, umma_desc_sw128(a_addr + kk * 32),
                                         umma_desc_sw128(b_addr + B_TILE_BYTES + kk * 32), IDESC, acc);
                        }
                    }
                    if (PAIR) umma_commit_pair(&ab_empty[s], CL_MASK);       // frees the stage in both CTAs
                    else if (CL == 1) umma_commit(&ab_empty[s]);              // frees the A/B stage when these MMAs retire
                    else umma_commit_mc(&ab_empty[s], CL_MASK);               // ... in every CTA of the cluster (they multicast into it)
                }
                if (PAIR) umma_commit_pair(&acc_full[ab], CL_MASK);          // accumulators complete in both CTAs
                else umma_commit(&acc_full[ab]);
            }
        }
        __syncwarp();
    } else if (warp < kFirstUnpackWarp) {
        // ============================ epilogue (4 warps = 4 TMEM lane quadrants) ============================
        const int quad = warp & 3;
        T *Y = static_cast<T *>(p.Y);
        if (p.pdl) ggb_pdl_wait();                             // Y may still be read by the kernel before the image kernel
        int tcount = 0;
        for (int tile = cid; tile < ntiles; tile += ncl, tcount++) {
            const int rb = (tile % n_rbg) * CL + crank, tg = tile / n_rbg;
            const int nb_valid = min(NB, p.n_tok_tiles - tg * NB);
            const int ab = tcount % NACC;
            mbar_wait(&acc_full[ab], (uint32_t)(tcount / NACC) & 1u);
            tc_fence_after();
            const int64_t tok0 = (int64_t)tg * NB * BN;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * ACC_COLS);
            constexpr int NCH = (BN + 31) / 32;
            constexpr int VPT = 16 / (int)sizeof(T);          // values per 16-byte vector
            constexpr int VPR = BM / VPT;                     // vectors per staged token row (128 weight rows)
            constexpr int TS = 16 * VPT / 4;                  // tokens per staging pass: an 8 KB buffer (32 for fp16/bf16, 16 for fp32)
            T *stage = reinterpret_cast<T *>(smem + p.off_epi);               // [TS tokens][128 weight rows]
            const int et = tid - kFirstEpiWarp * 32;          // 0..127
            const int64_t row0 = (int64_t)rb * BM;
            // tokens stored divided by a power of two (out-of-range magnitudes, quantize_b_image_kernel): rare, so one vote per tile
            const int ntok = (int)min((int64_t)NB * BN, p.M - tok0);
            uint32_t plain = 1u;
            for (int i = et; i < ntok; i += 128) plain &= (__ldg(p.row_scale + tok0 + i) == 1.0f) ? 1u : 0u;
            asm volatile("{\n\t.reg .pred pi, po;\n\tsetp.ne.u32 pi, %1, 0;\n\tbar.red.and.pred po, 2, 128, pi;\n\tselp.u32 %0, 1, 0, po;\n\t}"
                         : "=r"(plain) : "r"(plain) : "memory");
#pragma unroll 1
            for (int ch = 0; ch < NCH * nb_valid; ch++) {
                uint32_t v[32];
                tmem_ld32(taddr + ch * 32, v);        // NB == 2: the second accumulator follows at column ACC_COLS == BN
                tmem_ld_wait();
                if (!plain) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const int jt = ch * 32 + j;
                        const float sc = jt < ntok ? __ldg(p.row_scale + tok0 + jt) : 1.0f;
                        v[j] = __float_as_uint(__uint_as_float(v[j]) * sc);
                    }
                }
                // The accumulator is weight-row major (one TMEM lane per row): stores straight from it are 64-byte
                // pieces at a 2 * ldy stride.  Transposed through shared memory, every token row of the tile leaves
                // as 256 contiguous bytes (fp16) in 16-byte vectors.
#pragma unroll
                for (int h = 0; h < 32 / TS; h++) {
#pragma unroll
                    for (int j = 0; j < TS; j++) stage[j * BM + quad * 32 + lane] = from_f32<T>(__uint_as_float(v[h * TS + j]));
                    asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
                    for (int it = 0; it < TS * VPR / 128; it++) {
                        const int vid = et + it * 128, j = vid / VPR, c = vid % VPR;
                        const int jt = ch * 32 + h * TS + j;
                        const int64_t tok = tok0 + jt;
                        const int64_t r = row0 + c * VPT;
                        if ((NB == 2 || jt < BN) && tok < p.M && r < p.N) {
                            const uint4 val = *reinterpret_cast<const uint4 *>(stage + j * BM + c * VPT);
                            T *dst = Y + tok * p.ldy + r;
                            if (p.y_vec && r + VPT <= p.N) {
                                *reinterpret_cast<uint4 *>(dst) = val;
                            } else {
                                const T *e = reinterpret_cast<const T *>(&val);
#pragma unroll
                                for (int k = 0; k < VPT; k++)
                                    if (r + k < p.N) dst[k] = e[k];
                            }
                        }
                    }
                    asm volatile("bar.sync 2, 128;" ::: "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR && crank != 0) mbar_arrive_remote(&acc_empty[ab], 0);     // the leader issues the next tile's MMAs
                else mbar_arrive(&acc_empty[ab]);
            }
        }
    } else {
        // ============================ unpack warps: raw GGUF bytes -> fp16 swizzled A tile ============================
        const int u = tid - kFirstUnpackWarp * 32;             // 0..255
        int it = 0, rit = 0;
        for (int tile = cid; tile < ntiles; tile += ncl) {
            const int rb = (tile % n_rbg) * CL + crank;
            const int64_t row0 = (int64_t)rb * BM;
            for (int sb = 0; sb < p.nsb; sb++, rit++) {
                const int rs = rit % NRAW;
                mbar_wait(&raw_full[rs], (uint32_t)(rit / NRAW) & 1u);
                const uint8_t *raw = sRaw + rs * (BM * p.raw_slot);
                const int slot_off = p.w_fast ? ((sb * SB_BYTES) & 15) : 0;     // rows are 16-byte aligned in the TMA path
                using HU = HalfUnpack<TYPE>;
                uint32_t *tab = sTab + rs * (BM * HU::NSUB);
                if constexpr (HU::HAS) {
                    // decode every sub-block's (scale, min) pair of this raw stage once; the table is double-buffered
                    // with the raw ring, so this one barrier per raw stage also orders its reuse
                    for (int e = u; e < BM * HU::NSUB; e += kUnpackWarps * 32)
                        tab[e] = HU::sub_scale(raw + (e / HU::NSUB) * p.raw_slot + slot_off, e % HU::NSUB);
                    asm volatile("bar.sync 1, %0;" ::"n"(kUnpackWarps * 32) : "memory");
                }
                for (int q = 0; q < RAWK / KSTEP; q++, it++) {
                    const int s = it % NSTAGE;
                    mbar_wait(&ab_empty[s], ((uint32_t)(it / NSTAGE) & 1u) ^ 1u);
                    uint8_t *a_stage = sA + s * A_STAGE_BYTES;
                    if constexpr (HU::FAST32) {
                        // one 128-bit load per thread: row r, nibble bytes 16*half..16*half+15 of this k-step.  A quarter
                        // warp = 8 consecutive rows of the same half: its 128-bit loads (row stride 144 B) and its swizzled
                        // 128-bit stores (chunk ^ (row & 7)) both touch 8 distinct 16-byte bank groups.
                        const int r = (u & 7) | ((u >> 4) << 3), half = (u >> 3) & 1;
                        // the two table words this thread's 32 values use: Q4_K sub-blocks 2q, 2q+1; Q6_K (16 per row, one per
                        // 16 values) sub-blocks 4q + half and 4q + 2 + half
                        uint2 sc;
                        if constexpr (HU::NSUB == 8) sc = *reinterpret_cast<const uint2 *>(tab + r * HU::NSUB + 2 * q);
                        else sc = make_uint2(tab[r * HU::NSUB + 4 * q + half], tab[r * HU::NSUB + 4 * q + 2 + half]);
                        uint4 o[4];
                        HU::run32(raw + r * p.raw_slot + slot_off, q, half, sc, o);
                        uint8_t *row = a_stage + r * 128;
                        const int sw = r & 7;
                        *reinterpret_cast<uint4 *>(row + (((2 * half) ^ sw) << 4)) = o[0];
                        *reinterpret_cast<uint4 *>(row + (((2 * half + 1) ^ sw) << 4)) = o[1];
                        *reinterpret_cast<uint4 *>(row + (((4 + 2 * half) ^ sw) << 4)) = o[2];
                        *reinterpret_cast<uint4 *>(row + (((5 + 2 * half) ^ sw) << 4)) = o[3];
                    } else
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int item = i * 256 + u;
                        const int r = item >> 3, c = item & 7;
                        const int e = q * KSTEP + c * 8;                       // element offset inside the raw stage
                        uint8_t *dst = a_stage + r * 128 + ((c ^ (r & 7)) << 4);
                        if constexpr (HU::HAS) {
                            const int pp = e >> 3;                             // K-quants: one block per raw stage
                            *reinterpret_cast<uint4 *>(dst) = HU::run(raw + r * p.raw_slot + slot_off, pp, tab[r * HU::NSUB + HU::sub_of(pp)]);
                        } else {
                            float y[8];
                            if (sb * RAWK + e < p.K) {
                                const uint8_t *blk = raw + r * p.raw_slot + slot_off + (e / QK) * BS;
                                Unpack8<TYPE>::run(blk, (e % QK) >> 3, y);
                            } else {
#pragma unroll
                                for (int k = 0; k < 8; k++) y[k] = 0.0f;
                            }
                            store8(reinterpret_cast<__half *>(dst), y);
                        }
                    }
                    fence_proxy_async();                       // generic-proxy stores -> visible to the UMMA (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[s]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_empty[rs]);
            }
        }
    }

    // A tail launch did not wait for the main launch it overlaps with; the kernel AFTER it only waits for the tail, so the
    // tail must not complete first: its CTAs wait for the main launch here, at their end.
    if (p.pdl == 0) ggb_pdl_wait();
    tc_fence_before();
    if (CL == 1) __syncthreads();
    else cluster_sync_all();                                   // no CTA leaves while a peer may still signal or copy into it
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ---------------------------------------------------------------- host side
int g_mmq_cfg = 0;           // tuning bits: 1 = one token tile per weight stage, 2 = no clusters, 4 = no tail launch, 8 = no CTA-pair MMA
int pick_bn(int64_t M) {
    if (M > 128) return 256;
    if (M > 64) return 128;
    if (M > 32) return 64;
    if (M > 16) return 32;
    return 16;
}
int64_t padded_k(int64_t K) { return (K + RAWK - 1) / RAWK * RAWK; }
int64_t padded_m(int64_t M) { const int bn = pick_bn(M); return (M + bn - 1) / bn * bn; }

template <typename T>
int launch_quantize(const void *X, int64_t M, int64_t K, uint8_t *img, float *row_scale, bool d_half, cudaStream_t stream) {
    const int64_t Kp = padded_k(K), Mp = padded_m(M);
    const bool vec_ok = (reinterpret_cast<uintptr_t>(X) % 16 == 0) && (K % 8 == 0);
    const int64_t nblk = M;                       // one CTA per token
    if (nblk > 0x7fffffffLL) {
        ggb_set_error("ggb_mmq: activation matrix too large");
        return GGB_E_SHAPE;
    }
    const T *x = static_cast<const T *>(X);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nblk);
    cfg.blockDim = dim3(256);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<quantize_b_image_kernel<T, true>>(ggb_device_info().dev)));
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<quantize_b_image_kernel<T, false>>(ggb_device_info().dev)));
    if (d_half) GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, quantize_b_image_kernel<T, true>, x, img, row_scale, M, (int)K, (int)Kp, (int)Mp, vec_ok));
    else        GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, quantize_b_image_kernel<T, false>, x, img, row_scale, M, (int)K, (int)Kp, (int)Mp, vec_ok));
    return GGB_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point query (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// returns 1 when the configuration does not fit in shared memory (the caller tries the next one)
// img_plane = token atoms per K-plane of the image (0: this problem's own padded M); tile_limit = cluster-level tiles to
// run (-1: all)
template <int TYPE, int BN, int NB, int CL, typename T, bool PAIR = false>
int launch_gemm(const void *W, const uint8_t *img, const float *rscale, void *Y, int64_t M, int64_t K, int64_t N, int64_t ldy, cudaStream_t stream,
                int img_plane = 0, int tile_limit = -1, bool is_tail = false) {
    constexpr int NSTAGE = PAIR ? 3 : nstage_of(NB);
    constexpr int QK = BlockTraits<TYPE>::QK, BS = BlockTraits<TYPE>::BS;
    constexpr int SB_BYTES = RAWK / QK * BS;
    const DeviceInfo &di = ggb_device_info();
    Params p;
    p.W = static_cast<const uint8_t *>(W);
    p.Ximg = img;
    p.row_scale = rscale;
    p.Y = Y;
    p.N = N;
    p.M = M;
    p.ldy = ldy;
    p.K = (int)K;
    p.row_bytes = (int)(K / QK * BS);
    p.nsb = (int)(padded_k(K) / RAWK);
    p.Mp = (int)((M + BN - 1) / BN * BN);
    p.img_plane = img_plane > 0 ? img_plane : (int)(padded_m(M) >> 3);
    p.ntiles_run = tile_limit >= 0 ? tile_limit : 0x7fffffff;
    p.pdl = is_tail ? 0 : (ggb_static_weights() ? 2 : 1);
    p.n_row_blocks = (int)((N + BM - 1) / BM);
    p.n_tok_tiles = p.Mp / BN;
    p.n_tok_groups = (p.n_tok_tiles + NB - 1) / NB;
    constexpr int BOX_INNER = (SB_BYTES + ((SB_BYTES % 16) ? 14 : 0) + 15) & ~15;
    p.raw_slot = BOX_INNER;
    p.w_fast = ((reinterpret_cast<uintptr_t>(W) & 15) == 0) && (p.row_bytes % 16 == 0);
    p.w_lo = 0;
    alignas(64) CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (p.w_fast) {
        EncodeTiledFn enc = get_encode_fn();
        if (!enc) {
            ggb_set_error("ggb_mmq: cuTensorMapEncodeTiled is not available from this driver");
            return GGB_E_CUDA;
        }
        // W as a 2-D tensor of uint16 [N rows][row_bytes / 2], box = [128 rows][BOX_INNER / 2]
        const cuuint64_t gdim[2] = {(cuuint64_t)(p.row_bytes / 2), (cuuint64_t)N};
        const cuuint64_t gstride[1] = {(cuuint64_t)p.row_bytes};
        const cuuint32_t box[2] = {(cuuint32_t)(BOX_INNER / 2), (cuuint32_t)BM};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(W), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            ggb_set_error("ggb_mmq: cuTensorMapEncodeTiled failed (%d)", (int)r);
            return GGB_E_CUDA;
        }
    }
    p.off_b = NSTAGE * A_STAGE_BYTES;
    p.off_raw = p.off_b + NSTAGE * NB * BN * (PAIR ? 64 : 128);
    p.off_raw = (p.off_raw + 1023) & ~1023;
    p.off_tab = (p.off_raw + NRAW * BM * p.raw_slot + 127) & ~127;
    p.off_epi = (p.off_tab + (HalfUnpack<TYPE>::HAS ? NRAW * BM * HalfUnpack<TYPE>::NSUB * 4 : 0) + 127) & ~127;
    p.off_bar = p.off_epi + 8192;
    p.y_vec = ((reinterpret_cast<uintptr_t>(Y) & 15) == 0) && ((ldy * (int64_t)sizeof(T)) % 16 == 0);
    const int smem = p.off_bar + 256 + 1024;     // + slack so the kernel can align its base to 1024
    if (smem > di.max_smem_optin) {
        if (NB > 1) return 1;
        ggb_set_error("ggb_mmq: tile needs %d bytes of shared memory (limit %d)", smem, di.max_smem_optin);
        return GGB_E_SHAPE;
    }
    static bool attr_set[64] = {};
    if (!attr_set[di.dev]) {
        GGB_CUDA_TRY(cudaFuncSetAttribute(mmq_kernel<TYPE, BN, NB, CL, T, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.max_smem_optin));
        attr_set[di.dev] = true;
    }
    int64_t ntiles = (int64_t)((p.n_row_blocks + CL - 1) / CL) * p.n_tok_groups;             // cluster-level tiles
    if (ntiles > p.ntiles_run) ntiles = p.ntiles_run;
    if (ntiles <= 0) return GGB_OK;
    const int64_t max_cl = di.sm_count / CL;
    const int grid = (int)(ntiles < max_cl ? ntiles : max_cl) * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    GGB_CUDA_TRY((ggb_prefer_max_smem_once<mmq_kernel<TYPE, BN, NB, CL, T, PAIR>>(di.dev)));
    GGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, mmq_kernel<TYPE, BN, NB, CL, T, PAIR>, p, tmap));
    return GGB_OK;
}

template <int TYPE, typename T>
int run_typed(const void *W, const void *X, void *Y, int64_t M, int64_t K, int64_t N, int64_t ldy, uint8_t *img, cudaStream_t stream) {
    float *rs = reinterpret_cast<float *>(img + padded_k(K) * padded_m(M) * 2);      // per-token scales follow the image
    int rc = launch_quantize<T>(X, M, K, img, rs, ggb_mmq_need_sum(TYPE), stream);
    if (rc != GGB_OK) return rc;
    switch (pick_bn(M)) {
        case 256:
            // enough tokens for two 256-token tiles per unpacked weight stage, and enough tiles to fill the GPU
            // enough tokens for two 256-token tiles per unpacked weight stage, and enough tiles to fill the GPU
            if (M > 256 && (g_mmq_cfg & 1) == 0 &&
                ((N + BM - 1) / BM) * ((padded_m(M) / 256 + 1) / 2) >= ggb_device_info().sm_count) {
                if (g_mmq_cfg & 2) {
                    rc = launch_gemm<TYPE, 256, 2, 1, T>(W, img, rs, Y, M, K, N, ldy, stream);
                    if (rc != 1) return rc;
                } else {
                    // Wave quantisation: with 74 clusters, the last round of 128 x 512 tiles is often nearly empty (M = 2048:
                    // 224 cluster tiles = 3 rounds + 2 tiles).  Those last `rem` tiles are a rectangle (the last row blocks of
                    // the last token group): run them as 128 x 64 tiles in a second launch that fills the GPU for a fraction
                    // of a round instead.
                    const int sms = ggb_device_info().sm_count, ncl = sms / 2;
                    const int n_rbg = (int)(((N + BM - 1) / BM + 1) / 2), n_tg = (int)((padded_m(M) / 256 + 1) / 2);
                    const int ntiles = n_rbg * n_tg, rem = ntiles % ncl;
                    const bool tail = (g_mmq_cfg & 4) == 0 && ntiles > ncl && rem > 0 && rem <= n_rbg && rem * 16 <= sms + sms / 4 &&
                                      (K % 256 == 0 || BlockTraits<TYPE>::QK == 32);
                    rc = (g_mmq_cfg & 8) ? launch_gemm<TYPE, 256, 2, 2, T>(W, img, rs, Y, M, K, N, ldy, stream, 0, tail ? ntiles - rem : -1)
                                         : launch_gemm<TYPE, 256, 2, 2, T, true>(W, img, rs, Y, M, K, N, ldy, stream, 0, tail ? ntiles - rem : -1);
                    if (rc != 1) {
                        if (rc != GGB_OK || !tail) return rc;
                        const int64_t row0 = (int64_t)(n_rbg - rem) * 2 * BM, tok0 = (int64_t)(n_tg - 1) * 512;
                        const int64_t row_bytes = K / BlockTraits<TYPE>::QK * BlockTraits<TYPE>::BS;
                        return launch_gemm<TYPE, 64, 1, 1, T>(static_cast<const uint8_t *>(W) + row0 * row_bytes, img + (tok0 >> 3) * 1024, rs + tok0,
                                                              static_cast<T *>(Y) + tok0 * ldy + row0, M - tok0, K, N - row0, ldy, stream,
                                                              (int)(padded_m(M) >> 3), -1, /*is_tail=*/true);
                    }
                }
            }
            if ((g_mmq_cfg & 2) == 0 && (N + BM - 1) / BM * (padded_m(M) / 256) >= ggb_device_info().sm_count)
                return launch_gemm<TYPE, 256, 1, 2, T>(W, img, rs, Y, M, K, N, ldy, stream);
            return launch_gemm<TYPE, 256, 1, 1, T>(W, img, rs, Y, M, K, N, ldy, stream);
        case 128: return launch_gemm<TYPE, 128, 1, 1, T>(W, img, rs, Y, M, K, N, ldy, stream);
        case 64:  return launch_gemm<TYPE, 64, 1, 1, T>(W, img, rs, Y, M, K, N, ldy, stream);
        case 32:  return launch_gemm<TYPE, 32, 1, 1, T>(W, img, rs, Y, M, K, N, ldy, stream);
        default:  return launch_gemm<TYPE, 16, 1, 1, T>(W, img, rs, Y, M, K, N, ldy, stream);
    }
}

template <int TYPE>
int run_dtype(const void *W, const void *X, void *Y, int dt, int64_t M, int64_t K, int64_t N, int64_t ldy, uint8_t *img, cudaStream_t stream) {
    switch (dt) {
        case GGB_F32:  return run_typed<TYPE, float>(W, X, Y, M, K, N, ldy, img, stream);
        case GGB_F16:  return run_typed<TYPE, __half>(W, X, Y, M, K, N, ldy, img, stream);
        case GGB_BF16: return run_typed<TYPE, __nv_bfloat16>(W, X, Y, M, K, N, ldy, img, stream);
        default:
            ggb_set_error("ggb_mmq: unknown dtype %d", dt);
            return GGB_E_DTYPE;
    }
}

}  // namespace

void ggb_mmq_tune(int cfg) { g_mmq_cfg = cfg; }

int64_t ggb_mmq_workspace_bytes_impl(int, int64_t M, int64_t K, int64_t) {
    return padded_k(K) * padded_m(M) * 2 + padded_m(M) * 4;       // the fp16 B-operand image + one fp32 scale per token
}

int ggb_mmq_impl(const void *W, const void *X, void *Y, int type, int xy_dtype, int64_t M, int64_t K, int64_t N,
                 int64_t ldy, void *ws, int64_t, cudaStream_t stream) {
    if (K > (1 << 24) || M > (1LL << 30) || N > (1LL << 30)) {
        ggb_set_error("ggb_mmq: shape too large");
        return GGB_E_SHAPE;
    }
    uint8_t *img = static_cast<uint8_t *>(ws);
    switch (type) {
        case T_Q4_0: return run_dtype<T_Q4_0>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q4_1: return run_dtype<T_Q4_1>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q5_0: return run_dtype<T_Q5_0>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q5_1: return run_dtype<T_Q5_1>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q8_0: return run_dtype<T_Q8_0>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q2_K: return run_dtype<T_Q2_K>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q3_K: return run_dtype<T_Q3_K>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q4_K: return run_dtype<T_Q4_K>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q5_K: return run_dtype<T_Q5_K>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        case T_Q6_K: return run_dtype<T_Q6_K>(W, X, Y, xy_dtype, M, K, N, ldy, img, stream);
        default:
            ggb_set_error("ggb_mmq: unsupported ggml type %d", type);
            return GGB_E_TYPE;
    }
}
